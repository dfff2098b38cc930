/*
 * sgm_b200.h -- additive C-ABI of libsgm_b200.so (plain pointers and sizes only).
 *
 * The three reference entry points live in SemiGlobalMatching.h and keep the reference's signatures.
 * Everything here is NEW surface that the reference does not have (last row):
 * explicit contexts (the reference has one global instance, SemiGlobalMatching.c:27), device
 * selection, stage taps (the reference exposes its global buffers instead, SemiGlobalMatching.h:67-72),
 * device-pointer and batched entry points.  SGM_Initialize/SGM_Reset/SGM_Match are thin wrappers over
 * a process-global context driven through these functions.
 *
 * All functions return 0 on success and a negative SGMB_E_* code on failure; none aborts, throws or
 * falls back to a CPU implementation.  SGMB_LastError() gives a human-readable reason (thread-local).
 */
#ifndef SGM_B200_H
#define SGM_B200_H

#include <stddef.h>
#include <stdint.h>
#include "SemiGlobalMatching.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct SGMB_Context SGMB_Context;

enum {
    SGMB_OK = 0,
    SGMB_E_ARG = -1,        /* the argument errors the reference rejects (SemiGlobalMatching.c:43-48,70-75) */
    SGMB_E_UNSUPPORTED = -2,/* disparity range > 256, negative P1/P2, more than 65535 rows/columns        */
    SGMB_E_CUDA = -3,       /* no device, allocation failure, launch failure                               */
    SGMB_E_STATE = -4       /* not configured / stage not retained / wrong buffer size                     */
};

/* Pipeline switches (SGMB_SetPipeline).  Default = SGMB_PIPE_REFERENCE: what SGM_Match does. */
enum {
    SGMB_PIPE_SPECKLE = 1u << 0,  /* run the speckle filter when SGMOption.is_remove_speckles (SemiGlobalMatching.c:113-117) */
    SGMB_PIPE_MEDIAN  = 1u << 1,  /* run the in-place 3x3 median (SemiGlobalMatching.c:120)                                  */
    SGMB_PIPE_TAPS    = 1u << 2,  /* retain every intermediate stage for SGMB_GetStage (allocates the uint16 S volume)       */
    SGMB_PIPE_REFERENCE = SGMB_PIPE_SPECKLE | SGMB_PIPE_MEDIAN,
    SGMB_PIPE_HOTPATH = 0u        /* census .. LR check only: the north-star hot path                                         */
};

/* Stages readable with SGMB_GetStage after a Match (element type, element count; N = W*H, D = range). */
enum {
    SGMB_STAGE_CENSUS_LEFT   = 0,  /* uint32 [N]     census_transform_5x5       SemiGlobalMatching.c:134-159 (uint64 [N] with the 9x7 window) */
    SGMB_STAGE_CENSUS_RIGHT  = 1,  /* uint32 [N]                                                             */
    SGMB_STAGE_AGGR          = 2,  /* uint16 [N*D]   S(p,d), reference layout   SemiGlobalMatching.c:198-372 (needs SGMB_PIPE_TAPS) */
    SGMB_STAGE_DISP_LEFT_WTA = 3,  /* float  [N]     left view before LR check  SemiGlobalMatching.c:374-443 (needs SGMB_PIPE_TAPS) */
    SGMB_STAGE_DISP_RIGHT    = 4,  /* float  [N]     right view                 SemiGlobalMatching.c:105     (needs SGMB_PIPE_TAPS) */
    SGMB_STAGE_DISP_LR       = 5,  /* float  [N]     after LR check             SemiGlobalMatching.c:445-470 */
    SGMB_STAGE_DISP_SPECKLE  = 6,  /* float  [N]     after speckle removal      SemiGlobalMatching.c:585-642 (needs SGMB_PIPE_TAPS) */
    SGMB_STAGE_DISP_FINAL    = 7,  /* float  [N]     what SGM_Match returns                                   */
    SGMB_STAGE_SPECKLE_LABELS = 8, /* int32  [2*N]   speckle filter scratch: component root per pixel (-1: invalid), then size per root
                                      (exact below min_speckle_area, "at least min_speckle_area" above: counting stops there) */
    SGMB_STAGE_GREY_LEFT     = 9,  /* uint8  [N]     grey images the path ran on (after SGMB_MatchFrame: the fused conversion's output) */
    SGMB_STAGE_GREY_RIGHT    = 10,
    SGMB_STAGE_PATH_PLANE_0  = 16  /* uint8  [N*D]   +r: L_r(p,d) of direction r (order of SemiGlobalMatching.c:213-220) as written by
                                      its regular paths; pixels on an irregular path hold 0 (needs SGMB_PIPE_TAPS) */
};

const char* SGMB_LastError(void);

/* Number of CUDA devices visible, or a negative error. */
int SGMB_DeviceCount(void);

/* Create / destroy a context bound to one CUDA device.  `slots` (>= 1) is the number of frames that may
 * be in flight at once in SGMB_MatchBatch*; each slot owns a stream and a full set of device buffers. */
int  SGMB_Create(SGMB_Context** out, int device, int slots);
void SGMB_Destroy(SGMB_Context* ctx);

/* == SGM_Initialize on an explicit context. */
int SGMB_Configure(SGMB_Context* ctx, uint16_t width, uint16_t height, const SGMOption* option);
int SGMB_SetPipeline(SGMB_Context* ctx, unsigned flags);

/* Census window (additive extension).  5x5 (default) is the reference's
 * census_transform_5x5 (SemiGlobalMatching.c:134-159, uint32 descriptors) and the only mode with a reference
 * to compare against.  9x7 (9 columns, 7 rows) builds 63-bit descriptors in 64-bit words and costs them with
 * popcll(xor): same conventions generalised (neighbour < centre -> 1, rows outer / columns inner, first comparison
 * in the top used bit, border of 3 rows / 4 columns = 0, stage skipped for W <= 9 or H <= 7, out-of-row cost
 * 127).  Takes effect at the next SGMB_Configure; changing it un-configures the context.  The census stage taps
 * then hold 8-byte descriptors.  SGMB_SetGlobalCensusWindow (or env SGM_B200_CENSUS=9x7) selects it for the
 * context behind SGM_Initialize / SGM_Match. */
int SGMB_SetCensusWindow(SGMB_Context* ctx, int width, int height);
int SGMB_SetGlobalCensusWindow(int width, int height);

/* == SGM_Match: host pointers, blocking; H2D + kernels + D2H. */
int SGMB_Match(SGMB_Context* ctx, const uint8_t* img_left, const uint8_t* img_right, float* disp_left);

/* Device-resident variant: all three pointers are device memory on the context's device.  The work is
 * enqueued on slot 0's stream; with `sync` != 0 the call returns after completion. */
int SGMB_MatchDevice(SGMB_Context* ctx, const uint8_t* d_left, const uint8_t* d_right, float* d_disp, int sync);
int SGMB_Synchronize(SGMB_Context* ctx);

/* Batch of n independent pairs, host pointers, frames pipelined over the context's slots: copy-in, kernels and
 * copy-out of different frames overlap.  Page-locked buffers (SGMB_HostAlloc / cudaHostAlloc / cudaHostRegister) are
 * copied by the copy engine directly; pageable ones (malloc, static arrays) are staged through page-locked buffers
 * owned by the slot.  All pointers are validated before anything is enqueued; after an error every slot is drained
 * before the call returns, so no copy into a caller buffer is still in flight. */
int SGMB_MatchBatch(SGMB_Context* ctx, const uint8_t* const* lefts, const uint8_t* const* rights,
                    float* const* disps, int n);
/* Same with device-resident inputs and outputs (no PCIe traffic). */
int SGMB_MatchBatchDevice(SGMB_Context* ctx, const uint8_t* const* d_lefts, const uint8_t* const* d_rights,
                          float* const* d_disps, int n);

/* Batch sharded over several GPUs of one box: pair k goes to device devices[k * ndev / n] (contiguous
 * shards), one host worker thread and one private context per device, no inter-GPU communication. */
int SGMB_MatchBatchMultiGPU(const int* devices, int ndev, int slots_per_device, uint16_t width, uint16_t height,
                            const SGMOption* option, unsigned pipeline_flags, const uint8_t* const* lefts,
                            const uint8_t* const* rights, float* const* disps, int n);

/* Persistent variant of the above: one context per device kept alive between batches (no allocation per call).
 * SGMB_PoolConfigure == SGMB_SetPipeline + SGMB_Configure on every device (one host thread each);
 * SGMB_PoolMatchBatch shards the pairs like SGMB_MatchBatchMultiGPU and runs SGMB_MatchBatch per device concurrently.
 * SGMB_PoolContext gives access to a member context (e.g. for SGMB_SetCensusWindow before SGMB_PoolConfigure). */
typedef struct SGMB_Pool SGMB_Pool;
int  SGMB_PoolCreate(SGMB_Pool** out, const int* devices, int ndev, int slots_per_device);
void SGMB_PoolDestroy(SGMB_Pool* pool);
int  SGMB_PoolSize(SGMB_Pool* pool);
SGMB_Context* SGMB_PoolContext(SGMB_Pool* pool, int index);
int  SGMB_PoolConfigure(SGMB_Pool* pool, uint16_t width, uint16_t height, const SGMOption* option, unsigned pipeline_flags);
int  SGMB_PoolMatchBatch(SGMB_Pool* pool, const uint8_t* const* lefts, const uint8_t* const* rights, float* const* disps, int n);

/* ---- The steps either side of the path in the reference system ---- */

/* Colour -> grey weights used by SGMB_MatchFrame*: grey = (wR*R + 150*G + 29*B) >> 8. */
enum {
    SGMB_GREY_BOARD = 0,   /* wR = 76: the board's convert_to_gray, ZedBoard/.../src/stereo_matching.c:18-24 (default)      */
    SGMB_GREY_STB   = 1    /* wR = 77: stb_image's conversion used by the demo's loader, stb_image.h:1746-1749 / main.c:25  */
};
int SGMB_SetGreyFormula(SGMB_Context* ctx, int formula);

/* One frame in the board's layout: six planes of width*height bytes, left B,G,R then right B,G,R
 * (SteroPairImg_t, ZedBoard/.../src/frame_buffer.h:29-41; sent plane by plane by HostScript_Server/server.py:126-131).
 * The colour -> grey conversion is fused into the census kernel's staging pass.  calib20 == NULL: `out` receives the
 * disparity map (as SGM_Match).  Otherwise calib20 is the 20-float wire calibration (cam0[9], cam1[9], doffs,
 * baseline; HostScript_Server/stereo_calibration.py:177-194, frame_buffer.h:16-22) and `out` receives the depth map
 * baseline*fx/(disparity+doffs) in float32 (HostScript_Server/depth_image.py:138-165), NaN where the disparity is
 * invalid -- the float32 rows the board replies with (ZedBoard/.../src/tcp_perf_client.c:92-143). */
int SGMB_MatchFrame(SGMB_Context* ctx, const uint8_t* planes6, const float* calib20, float* out);
int SGMB_MatchFrameDevice(SGMB_Context* ctx, const uint8_t* d_planes6, const float* calib20, float* d_out, int sync);

/* Wire messages (host only, no CUDA call).  Reply: type byte 3, frame id u32 LE, width u16 LE, height u16 LE, then
 * width*height float32 (tcp_perf_client.c:106-131; read back by server.py:148-177).  Frame header: '<BiHH' = type
 * (1: followed by the 80-byte calibration, 2: images only), sequence, width, height (server.py:114); payload_bytes =
 * what follows the 9 header bytes. */
size_t SGMB_DepthReplyBytes(uint16_t width, uint16_t height);
int SGMB_PackDepthReply(uint32_t frame_id, uint16_t width, uint16_t height, const float* depth, uint8_t* dst, size_t capacity);
int SGMB_ParseFrameHeader(const uint8_t* bytes9, int* type, int32_t* seq, uint16_t* width, uint16_t* height, size_t* payload_bytes);

/* Evaluation step after the path.  depth = baseline*fx/(disp+doffs), float32, numpy's evaluation order
 * (depth_image.py:138-165); +inf disparities (INVALID_FLOAT) become NaN. */
int SGMB_DisparityToDepth(SGMB_Context* ctx, const float* disp, size_t n, float baseline, float fx, float doffs, float* depth);
int SGMB_DisparityToDepthDevice(SGMB_Context* ctx, const float* d_disp, size_t n, float baseline, float fx, float doffs, float* d_depth);
/* compare_img (depth_image.py:276-319): over pixels where both maps are finite, RMSE and the fraction with
 * |test - gt| > abs_thresh (the reference's default: 10 mm); (NaN, NaN, 0) when no pixel is valid. */
int SGMB_CompareDepth(SGMB_Context* ctx, const float* gt, const float* test, size_t n, float abs_thresh, double* rmse,
                      double* bpr, long long* n_valid);
int SGMB_CompareDepthDevice(SGMB_Context* ctx, const float* d_gt, const float* d_test, size_t n, float abs_thresh, double* rmse,
                            double* bpr, long long* n_valid);

/* Copy a retained stage of slot 0 to host memory; `bytes` must equal the stage's size. */
int SGMB_GetStage(SGMB_Context* ctx, int stage, void* host_dst, size_t bytes);

/* Page-locked host memory.  SGMB_HostAlloc / SGMB_HostFree allocate it; SGMB_HostRegister page-locks memory the caller
 * already owns (the static arrays of the reference's demo, main.c:25-26,81) until SGMB_HostUnregister, which must precede
 * freeing it.  With page-locked images the copy engine reads them directly, and with a page-locked disparity buffer and the
 * reference pipeline (median last, no taps) the last kernel writes the caller's buffer itself: no device-to-host copy is
 * left at the end of SGM_Match / SGMB_Match / SGMB_MatchBatch (0.61 ms per call at 1242x375, D=128 against 0.81 ms with
 * pageable buffers). */
int  SGMB_HostAlloc(void** out, size_t bytes);
void SGMB_HostFree(void* p);
int  SGMB_HostRegister(void* p, size_t bytes);
int  SGMB_HostUnregister(void* p);

/* Introspection for benchmarks: kernels launched per frame by the current plan, algorithmic and
 * actually-moved DRAM byte models, device-side time of the last SGMB_Match* in milliseconds. */
int    SGMB_KernelLaunchesPerFrame(SGMB_Context* ctx);
double SGMB_ModelBytesPerFrame(SGMB_Context* ctx);   /* W*H*(4*P*D + 6)                    */
double SGMB_PlanBytesPerFrame(SGMB_Context* ctx);    /* bytes this implementation's plan moves (planes) */
float  SGMB_LastDeviceMs(SGMB_Context* ctx);
/* Time `iters` back-to-back device-resident frames on slot 0 with CUDA events on its stream, after
 * `warmup` untimed ones; writes per-iteration milliseconds of the whole frame and of the dominant
 * (aggregation) kernel.  flush_l2 != 0 overwrites a >L2-sized scratch buffer between iterations. */
int SGMB_TimeDevice(SGMB_Context* ctx, const uint8_t* d_left, const uint8_t* d_right, float* d_disp, int warmup,
                    int iters, int flush_l2, float* frame_ms, float* aggr_kernel_ms);

/* Run `iters` device-resident frames back to back on slot 0 without host synchronisation and time the whole region
 * with CUDA events on that stream (*total_ms); agg_ms (optional, [iters]) receives the duration of every
 * aggregation-kernel launch inside the region.  The frames are recorded into one CUDA graph (kernels, the side-buffer
 * memsets, external event records for agg_ms) and replayed with a single launch (enqueued kernel by kernel when the
 * environment variable SGM_B200_NO_GRAPH is set). */
int SGMB_RunDevice(SGMB_Context* ctx, const uint8_t* d_left, const uint8_t* d_right, float* d_disp, int iters,
                   float* total_ms, float* agg_ms);

/* Same, replayed `replays` times (one graph launch each, every replay timed on its own: replay_ms[replays]); agg_ms
 * (optional, [iters]) holds the aggregation-kernel durations of the last replay. */
int SGMB_RunDeviceReplays(SGMB_Context* ctx, const uint8_t* d_left, const uint8_t* d_right, float* d_disp, int iters,
                          int replays, float* replay_ms, float* agg_ms);

/* Per-kernel durations of one device-resident frame of the current pipeline: CUDA events around every launch (direct
 * launches, no graph), averaged over `iters` frames after `warmup` untimed ones, written to kernel_ms[0..capacity).
 * Returns the number of kernels of the frame (or a negative error); SGMB_KernelName(ctx, k) names kernel k of the last call. */
int SGMB_TimeKernels(SGMB_Context* ctx, const uint8_t* d_left, const uint8_t* d_right, float* d_disp, int warmup, int iters,
                     float* kernel_ms, int capacity);
const char* SGMB_KernelName(SGMB_Context* ctx, int index);

/* The batch sharding rule of SGMB_MatchBatchMultiGPU / SGMB_PoolMatchBatch (host only, no CUDA call): device g of
 * ndev processes the contiguous pairs [*lo, *hi) of n. */
int SGMB_ShardRange(int n, int ndev, int g, int* lo, int* hi);

/* The context behind SGM_Initialize/SGM_Match (NULL before the first SGM_Initialize), and the device it
 * will use (default 0, or env SGM_B200_DEVICE). */
SGMB_Context* SGMB_GlobalContext(void);
int SGMB_SetGlobalDevice(int device);

/* Host-only views of the aggregation path topology (no CUDA call; usable without a GPU): the pixel indices
 * one path visits (index outside [0, W*H) == the reference's out-of-bounds visit, which is skipped), and
 * which paths of a direction (0..7, order of SemiGlobalMatching.c:213-220) leave their toroidal diagonal.
 * Both return the number of entries written or a negative error. */
int SGMB_DebugWalkPath(int width, int height, int direction, int path, int* positions, int capacity);
int SGMB_DebugClassifyPaths(int width, int height, int direction, uint8_t* irregular, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* SGM_B200_H */
