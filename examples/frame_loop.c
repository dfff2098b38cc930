/*
 * frame_loop.c -- a plain C caller of libsgm_b200.so.
 *
 * Part 1 is the reference's own call sequence (SemiGlobalMatching/SemiGlobalMatching/main.c:48-83): fill an
 * SGMOption, SGM_Initialize, SGM_Match -- unchanged signatures, so a reference caller only re-links.
 * Part 2 is the board's frame loop (ZedBoard/Vitis/lwip_tcp_perf_client/src/main.c:214-241) on the additive API: a
 * frame as the server sends it ('<BiHH' header, 80-byte calibration, six planes B,G,R / B,G,R) goes in, the reply
 * message the board would put on the wire (type 3, frame id, size, float32 depth rows) comes out.
 *
 * Build: make example      Run: examples/frame_loop   (needs one B200; there is no CPU fallback)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sgm_b200.h"

enum { W = 640, H = 360, D = 64 };

static uint32_t lcg(uint32_t* s) { *s = *s * 1664525u + 1013904223u; return *s >> 24; }

int main(void)
{
    /* a synthetic pair: random texture, right view shifted by 9 columns */
    uint8_t* left = malloc((size_t)W * H);
    uint8_t* right = malloc((size_t)W * H);
    uint8_t* tex = malloc((size_t)(W + D) * H);
    float* disp = malloc(sizeof(float) * W * H);
    uint32_t seed = 0xB200;
    for (size_t i = 0; i < (size_t)(W + D) * H; ++i) tex[i] = (uint8_t)lcg(&seed);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            left[y * W + x] = tex[y * (W + D) + x];
            right[y * W + x] = tex[y * (W + D) + x + 9];
        }

    /* ---- part 1: the reference's three functions (main.c:48-65 option values) */
    SGMOption opt = { .num_paths = 8, .min_disparity = 0, .max_disparity = D, .is_check_unique = true,
                      .uniqueness_ratio = 0.99f, .is_check_lr = true, .lrcheck_thres = 1.0f,
                      .is_remove_speckles = true, .min_speckle_area = 50, .p1 = 10, .p2_init = 150 };
    if (!SGM_Initialize(W, H, &opt)) { fprintf(stderr, "SGM_Initialize: %s\n", SGMB_LastError()); return 1; }
    if (!SGM_Match(left, right, disp)) { fprintf(stderr, "SGM_Match: %s\n", SGMB_LastError()); return 1; }
    size_t valid = 0, close = 0;
    for (size_t i = 0; i < (size_t)W * H; ++i)
        if (disp[i] != INVALID_FLOAT) { ++valid; close += fabsf(disp[i] - 9.0f) <= 0.5f; }
    printf("SGM_Match: %zu of %d pixels valid, %zu within 0.5 px of the true shift 9\n", valid, W * H, close);

    /* ---- part 2: one frame in the wire format through the additive API */
    SGMB_Context* ctx = NULL;
    if (SGMB_Create(&ctx, 0, 1) || SGMB_Configure(ctx, W, H, &opt)) { fprintf(stderr, "%s\n", SGMB_LastError()); return 1; }
    const size_t n = (size_t)W * H;
    uint8_t* msg = malloc(9 + 80 + 6 * n);
    msg[0] = 1;                                                   /* type 1: frame with calibration (server.py:114) */
    const int32_t seq = 7; memcpy(msg + 1, &seq, 4);
    const uint16_t w16 = W, h16 = H; memcpy(msg + 5, &w16, 2); memcpy(msg + 7, &h16, 2);
    float calib[20] = {0};
    calib[0] = 1733.74f; calib[18] = 0.0f; calib[19] = 536.62f;  /* cam0 fx, doffs, baseline (stereo_calibration.py:177-194) */
    memcpy(msg + 9, calib, 80);
    for (int c = 0; c < 3; ++c) {                                 /* grey texture in all three colour planes */
        memcpy(msg + 9 + 80 + (size_t)c * n, left, n);
        memcpy(msg + 9 + 80 + (size_t)(3 + c) * n, right, n);
    }
    int type; int32_t seqIn; uint16_t wi, hi; size_t payload;
    if (SGMB_ParseFrameHeader(msg, &type, &seqIn, &wi, &hi, &payload)) { fprintf(stderr, "%s\n", SGMB_LastError()); return 1; }
    float* depth = malloc(sizeof(float) * n);
    if (SGMB_MatchFrame(ctx, msg + 9 + 80, type == 1 ? (const float*)(msg + 9) : NULL, depth)) { fprintf(stderr, "%s\n", SGMB_LastError()); return 1; }
    uint8_t* reply = malloc(SGMB_DepthReplyBytes(wi, hi));
    SGMB_PackDepthReply((uint32_t)seqIn, wi, hi, depth, reply, SGMB_DepthReplyBytes(wi, hi));
    printf("frame %d (%ux%u, %zu payload bytes) -> reply of %zu bytes, depth at the centre %.1f mm (baseline*fx/9 = %.1f)\n",
           seqIn, wi, hi, payload, SGMB_DepthReplyBytes(wi, hi), depth[(H / 2) * W + W / 2], 536.62f * 1733.74f / 9.0f);
    SGMB_Destroy(ctx);
    free(left); free(right); free(tex); free(disp); free(msg); free(depth); free(reply);
    return 0;
}
