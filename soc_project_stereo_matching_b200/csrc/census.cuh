// census.cuh -- K1: census transform of both images in one launch.
//
// sgm_census<5, 5, uint32_t> restates census_transform_5x5 (SemiGlobalMatching.c:134-159): interior pixels
// only (2-pixel border stays 0), 25 comparisons "neighbour < centre", rows outer / columns inner, each
// shifted in from the LSB (first comparison ends at bit 24; the centre's own bit 12 is always 0).
// sgm_census<9, 7, desc64_t> is the EXTENSION named by the task (9 columns x 7 rows, 63 comparisons in a
// 64-bit word); the reference has no such code, so it follows the same conventions generalised: border of
// (3 rows, 4 columns) stays 0, stage skipped for W <= 9 or H <= 7 (parity pinned only by the oracle's
// own generalisation, see oracle/sgm_oracle.h).
//
// Layout produced (DT = descriptor type, K = 32 / sizeof(DT) descriptors per 256-bit load):
//   left  : DT [H*W]
//   right : K copies, copy a shifted right by a elements inside a zero-padded array:
//           right[a][padF + a + p] = census(p).  K2 needs, per lane, 2..16 consecutive right descriptors
//           cR[q-k]; picking the copy with (padF + a + q - (n-1)) % min(n, K) == 0 turns that window into
//           aligned 64/128/256-bit loads for every pixel position q.
// The image tile (+ halo) is staged in shared memory.
//
// PLANAR = true fuses the colour -> grey conversion of the board's frame format into the staging pass: the
// input of each view is three planes B, G, R of N bytes (zb/frame_buffer.h:29-41, sent plane by plane by
// host/server.py:126-131), grey = (wR*R + wG*G + wB*B) >> 8 with the board's weights 76/150/29
// (zb/stereo_matching.c:18-24) or stb_image's 77/150/29 (stb_image.h:1746-1749, what main.c's loader applies);
// the grey tile feeds the census directly and its interior is also written out (K2 needs the left grey image).
#pragma once

#include <stdint.h>

namespace sgmb {

struct CensusParams {
    const uint8_t* img[2];   // left, right: grey [N], or planar B,G,R [3][N] when PLANAR
    uint8_t* grey[2];        // PLANAR only: converted grey images [N]
    uint32_t wR, wG, wB;     // PLANAR only: grey = (wR*R + wG*G + wB*B) >> 8
    void* left;              // DT [N]
    void* pixL;              // {left descriptor, grey value} per pixel of the LEFT image: uint2 (32-bit descriptors) or
                             // uint4 {lo, hi, grey, 0} (64-bit descriptors)
    void* right4;            // DT [K][copyStride]
    size_t copyStride;       // elements per copy (padF + N + padB, multiple of 8)
    int padF;
    int W, H;
};

constexpr int kCensusTileW = 64;
constexpr int kCensusTileH = 8;
constexpr int kCensusPerThread = 4;      // horizontally adjacent outputs per thread
constexpr int kCensusThreads = kCensusTileW * kCensusTileH / kCensusPerThread;

template <int CW, int CH, typename DT, bool PLANAR>
__global__ void __launch_bounds__(kCensusThreads)
sgm_census(CensusParams P)
{
    // tile of 64 x 8 outputs, (64 + CW - 1) x (8 + CH - 1) inputs; every thread produces four horizontally adjacent outputs:
    // their windows span 4 + CW - 1 columns, fetched as 32-bit words (2 per row for 5x5, 3 for 9x7: 10 loads instead of the
    // 100 byte loads of four separate windows), the bytes picked with PRMT at compile-time positions
    constexpr int RX = CW / 2, RY = CH / 2, K = 32 / (int)sizeof(DT), PX = kCensusPerThread;
    constexpr int TW = kCensusTileW + CW - 1, TH = kCensusTileH + CH - 1, TP = (TW + 3 + 3) & ~3, NWORD = (PX + CW - 1 + 3) / 4;
    __shared__ __align__(16) uint8_t tile[TH][TP];
    const int which = blockIdx.z;
    const uint8_t* __restrict__ img = which ? P.img[1] : P.img[0];
    const int W = P.W, H = P.H;
    const int x0 = blockIdx.x * kCensusTileW, y0 = blockIdx.y * kCensusTileH;

    // all loads of the staging pass are issued before the first shared-memory store (one round trip to memory, not ITER)
    constexpr int ITER = (TH * TW + kCensusThreads - 1) / kCensusThreads;
    uint8_t staged[ITER];
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
        const int i = threadIdx.x + k * kCensusThreads;
        const int ty = i / TW, tx = i % TW;
        const int y = y0 + ty - RY, x = x0 + tx - RX;
        uint8_t v = 0;
        if (i < TH * TW && y >= 0 && y < H && x >= 0 && x < W) {
            const size_t p = (size_t)y * W + x;
            if (PLANAR) {
                const size_t N = (size_t)W * H;
                v = (uint8_t)((P.wB * __ldg(img + p) + P.wG * __ldg(img + N + p) + P.wR * __ldg(img + 2 * N + p)) >> 8);
                // each pixel belongs to the interior of exactly one tile
                if (ty >= RY && ty < RY + kCensusTileH && tx >= RX && tx < RX + kCensusTileW) P.grey[which][p] = v;
            } else {
                v = __ldg(img + p);
            }
        }
        staged[k] = v;
    }
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
        const int i = threadIdx.x + k * kCensusThreads;
        if (i < TH * TW) tile[i / TW][i % TW] = staged[k];
    }
    __syncthreads();

    const int ty = threadIdx.x / (kCensusTileW / PX);
    const int tx = (threadIdx.x % (kCensusTileW / PX)) * PX;
    const int y = y0 + ty;
    // Right image, 32-bit descriptors: the tile's descriptors are parked in shared memory and every shifted copy is then
    // written as aligned 128-bit pieces (below); all threads of the block stay for that pass.
    constexpr bool kStagedCopies = sizeof(DT) == 4;
    __shared__ uint32_t parked[kStagedCopies ? kCensusTileH : 1][kStagedCopies ? kCensusTileW + 4 : 1];   // + 4: neighbouring rows start four banks apart
    const bool viaShared = kStagedCopies && which == 1;
    if (y >= H && !viaShared) return;
    if (y < H) {
        uint32_t b[CH][PX + CW - 1];                              // b[r][c] = tile[ty + r][tx + c]: every byte is extracted once for all four windows
#pragma unroll
        for (int r = 0; r < CH; ++r)
#pragma unroll
            for (int k = 0; k < NWORD; ++k) {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(&tile[ty + r][tx + 4 * k]);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (4 * k + j < PX + CW - 1) b[r][4 * k + j] = __byte_perm(w, 0u, 0x4440u | (uint32_t)j);
            }
        auto px = [&](int r, int c) -> uint32_t { return b[r][c]; };
#pragma unroll
        for (int o = 0; o < PX; ++o) {
            const int x = x0 + tx + o;
            if (x >= W) break;
            DT bits = 0;
            const uint32_t centre = px(RY, o + RX);
            if (y >= RY && y < H - RY && x >= RX && x < W - RX && W > CW && H > CH) {
                // "neighbour < centre" is the sign of (neighbour - centre); a funnel shift moves it into the descriptor: two
                // instructions per comparison.  64-bit descriptors are built as two 32-bit words (the last 32 comparisons: low word).
                uint32_t acc[2] = {0u, 0u};
#pragma unroll
                for (int r = 0; r < CH; ++r)
#pragma unroll
                    for (int c = 0; c < CW; ++c) {
                        constexpr int total = CW * CH;
                        const int word = (sizeof(DT) == 8 && r * CW + c < total - 32) ? 0 : 1;
                        acc[word] = __funnelshift_l(px(r, o + c) - centre, acc[word], 1);
                    }
                bits = (DT)acc[1];
                if (sizeof(DT) == 8) bits |= (DT)((unsigned long long)acc[0] << 32);
            }
            const size_t p = (size_t)y * W + x;
            if (which == 0) {
                static_cast<DT*>(P.left)[p] = bits;
                const uint32_t grey = centre;
                if (sizeof(DT) == 4) static_cast<uint2*>(P.pixL)[p] = make_uint2((uint32_t)bits, grey);
                else static_cast<uint4*>(P.pixL)[p] = make_uint4((uint32_t)bits, (uint32_t)((unsigned long long)bits >> 32), grey, 0u);
            } else if (kStagedCopies) {
                parked[ty][tx + o] = (uint32_t)bits;
            } else {
#pragma unroll
                for (int a = 0; a < K; ++a) static_cast<DT*>(P.right4)[a * P.copyStride + P.padF + a + p] = bits;
            }
        }
    }
    if (!viaShared) return;
    // The K shifted copies of the tile's rows: copy a, tile row r = elements [e0, e0 + cols) of that copy with
    // e0 = padF + a + (y0 + r) * W + x0.  A tile row belongs to 16 consecutive threads for all copies: thread s writes the s-th
    // aligned 128-bit piece of each copy; the <= 3 elements before the first and after the last piece go to six of them.
    // (One 4-byte store per copy and output - 32 per thread - made this kernel 19.5 us at C2 with eight copies.)
    __syncthreads();
    if constexpr (kStagedCopies) {
        const int cols = min(kCensusTileW, W - x0);
        const int slot = threadIdx.x & 15;
        uint32_t* const copies = static_cast<uint32_t*>(P.right4);
        // thread -> (tile row r, slot) for every copy a in turn: eight (row, copy) pairs per thread, no loop over elements
        const int r = threadIdx.x >> 4;                                // kCensusThreads / 16 == kCensusTileH rows
        static_assert(kCensusThreads / 16 == kCensusTileH, "one group of 16 threads per tile row");
        if (y0 + r < H) {
            const size_t rowStart = (size_t)P.padF + (size_t)(y0 + r) * W + x0;
            const uint32_t* const src = &parked[r][0];
#pragma unroll
            for (int a = 0; a < K; ++a) {
                const size_t e0 = rowStart + a;
                uint32_t* const dst = copies + a * P.copyStride + e0;
                const int lead = min((int)((0 - e0) & 3), cols);      // elements before the first aligned piece
                const int nfull = (cols - lead) >> 2;                  // aligned 128-bit pieces (<= 16)
                if (slot < nfull) {
                    const uint32_t* q = src + lead + 4 * slot;
                    *reinterpret_cast<uint4*>(dst + lead + 4 * slot) = make_uint4(q[0], q[1], q[2], q[3]);
                }
                // the <= 3 elements in front and the <= 3 behind the pieces: one thread each (slots 0-2 and 3-5)
                if (slot < lead) dst[slot] = src[slot];
                const int t = lead + 4 * nfull + slot - 3;
                if (slot >= 3 && slot < 6 && t < cols) dst[t] = src[t];
            }
        }
    }
}

}  // namespace sgmb
