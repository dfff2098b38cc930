// wta.cuh -- K3: sum the per-direction planes into S (uint16) and, in the same pass over the data,
// select disparities for BOTH views, refine them to sub-pixel precision and run the left-right check.
//
// Restates ComputeDisparity() (SemiGlobalMatching.c:374-443, both `inverse` modes) and LRCheck()
// (SemiGlobalMatching.c:445-470).  S(p,d) = sum_r L_r(p,d) is exactly the reference's cost_aggr after
// CostAggregation() (SemiGlobalMatching.c:198-221,345); it only ever exists in shared memory unless the taps
// are enabled, in which case it is also written out in the reference layout [N][D].
//
// The kernel streams 8 (or 4) byte planes once and is meant to run at HBM speed, so the loads of the NEXT tile
// (128 bytes per thread, straight into the registers phase A has just consumed) are in flight during phase B.
// One CTA owns one image row and walks it from right to left in tiles of TW columns:
//   phase A : CPP lanes share a pixel, each owning 16 consecutive disparities (one 128-bit load per plane): add
//             the planes as packed 16-bit fields (+ the side buffer of irregular paths) and store S into a
//             shared-memory ring of 2*TW + D columns;
//   phase B : (after one __syncthreads) half of the CTA scans the LEFT view of the tile's columns, the other half
//             the RIGHT view of the same number of pixels, those whose first column lies in this tile (the right view
//             reads the ring along the diagonal S[x + d][d], SemiGlobalMatching.c:397-399; walking right to left, all
//             later columns of such a pixel are already in the ring).  Four lanes share a pixel; keys
//             (cost << 16 | d) make "lowest d wins ties" and "second = min over d != best" one integer min / max
//             each (:390-393,412-419); the four partial results are combined with xor shuffles and parked,
//             together with the two costs next to the best one, in a 16-byte record per pixel;
//   epilogue: every thread finishes whole pixels from the records (uniqueness, border, sub-pixel fit), then
//             LRCheck runs over the row.
// The ring spans 2*TW + D columns so that phase A of the next tile never overwrites a column phase B still reads:
// one barrier per tile.  Its row stride (16*CPP + 2 halfwords, an odd number of 32-bit words) spreads row-wise
// and diagonal accesses over the banks.
#pragma once

#include <math.h>
#include <stdint.h>

namespace sgmb {

struct WtaParams {
    const uint8_t* planes;      // [P][planeStride]
    size_t planeStride;
    int nPlanes;                // 4 or 8
    const uint16_t* side;       // [E][Dp]
    const int32_t* entryOf;     // [N]
    int hasSide;                // 0: no irregular paths (4-path mode)
    uint16_t* S;                // optional tap [N][D]
    float* dispLeftWta;         // optional tap [N]
    float* dispRight;           // optional tap [N]
    float* dispOut;             // [N] left disparity after the LR check (or plain WTA when LR is off)
    float* rightRow;            // [N] scratch: right-view disparities
    uint4* records;             // [2][N] scratch: parked (best, second best, neighbour costs) of the left / right view
    int W, H, D, Dp, dmin;
    int checkUnique;  float oneMinusRatio;   // (1 - uniqueness_ratio), formed in float like the reference
    int checkLR;      float lrThres;
    int ringCols;               // 2 * TW + D
};

__device__ __forceinline__ float sgm_invalid() { return __int_as_float(0x7f800000); }

struct WtaPair { uint32_t kmin, ksec; };   // smallest and second smallest key; key = cost << 16 | disparity index

__device__ __forceinline__ void wta_push2(WtaPair& w, uint32_t ka, uint32_t kb)
{
    const uint32_t lo = min(ka, kb), hi = max(ka, kb);
    w.ksec = min(min(w.ksec, max(w.kmin, lo)), hi);
    w.kmin = min(w.kmin, lo);
}

__device__ __forceinline__ void wta_push1(WtaPair& w, uint32_t k)
{
    w.ksec = min(w.ksec, max(w.kmin, k));
    w.kmin = min(w.kmin, k);
}

// combine the partial results of the four lanes that share a pixel (lane ids differ in bits 0 and 1)
__device__ __forceinline__ void wta_reduce4(WtaPair& w)
{
#pragma unroll
    for (int o = 2; o > 0; o >>= 1) {
        const uint32_t k2 = __shfl_xor_sync(0xffffffffu, w.kmin, o), s2 = __shfl_xor_sync(0xffffffffu, w.ksec, o);
        w.ksec = min(min(w.ksec, s2), max(w.kmin, k2));
        w.kmin = min(w.kmin, k2);
    }
}

// Uniqueness / border tests and the sub-pixel fit (SemiGlobalMatching.c:412-440) from a parked record:
// x = smallest key, y = second smallest key, z = cost(best - 1) | cost(best + 1) << 16 (65535 where the
// reference sees UINT16_MAX).
__device__ __forceinline__ float wta_finish(uint4 rec, int D, int dmin, int checkUnique, float oneMinusRatio)
{
    const int best = (int)(rec.x & 0xFFFFu);
    const int cmin = (int)(rec.x >> 16);
    const int second = (int)(rec.y >> 16);
    if (cmin == 0xFFFF) return sgm_invalid();          // no candidate at all (right view, x + dmin >= W)
    if (checkUnique) {
        const float lim = __fmul_rn((float)cmin, oneMinusRatio);
        const int ilim = (int)(__float2uint_rz(lim) & 0xFFFFu);            // (uint16_t)(min * (1 - ratio))  :422
        if (second - cmin <= ilim) return sgm_invalid();
    }
    if (best == 0 || best == D - 1) return sgm_invalid();                  // :428-431
    const int c1 = (int)(int16_t)(uint16_t)(rec.z & 0xFFFFu);             // :434
    const int c2 = (int)(int16_t)(uint16_t)(rec.z >> 16);                 // :435 (65535 -> -1)
    int denom = (int)(int16_t)(c1 + c2 - 2 * cmin);                        // :437
    if (denom < 1) denom = 1;
    return __fadd_rn((float)(best + dmin), __fdiv_rn((float)(c1 - c2), __fmul_rn((float)denom, 2.0f)));   // :440
}

template <int CPP>
struct WtaShape {
    static constexpr int kThreads = CPP == 16 ? 512 : 256;
    static constexpr int kTW = kThreads / CPP;          // columns per tile
    static constexpr int kRS = 16 * CPP + 2;            // ring row stride in halfwords
};

template <int CPP, int NP, bool TAPS>
__global__ void __launch_bounds__(WtaShape<CPP>::kThreads, CPP == 16 ? 1 : 3)
sgm_reduce_wta_lr(const __grid_constant__ WtaParams P)
{
    constexpr int THREADS = WtaShape<CPP>::kThreads, TW = WtaShape<CPP>::kTW, RS = WtaShape<CPP>::kRS;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int W = P.W, D = P.D, Dp = P.Dp;
    const int RB = P.ringCols;                    // 2*TW + D columns
    uint16_t* ring = reinterpret_cast<uint16_t*>(smem_raw);

    const int y = blockIdx.x;
    const size_t rowBase = (size_t)y * W;
    const int tid = threadIdx.x;
    float* rowL = P.dispOut + rowBase;
    float* rowR = P.rightRow + rowBase;
    uint4* recL = P.records + rowBase;
    uint4* recR = P.records + (size_t)P.H * W + rowBase;

    // ---- phase A mapping: (pixel of the tile, 16-disparity chunk)
    const int pix = tid / CPP, v = tid % CPP;
    const bool chunkOk = 16 * v < Dp;             // CPP is the power of two >= Dp / 16
    const uint8_t* srcLane = P.planes + rowBase * Dp + 16 * v;
    // ---- phase B mapping: first half of the CTA = left view, second half = right view; four lanes per pixel
    constexpr int HALF = THREADS / 2, GROUPS = HALF / 4;
    const bool rightRole = tid >= HALF;
    const int grp = (tid % HALF) / 4, sub = tid % 4;
    const int CH = (((D + 3) / 4) + 1) & ~1;      // disparities per lane, even
    const int k0 = sub * CH;                      // this lane scans indices [k0, min(k0 + CH, n))

    uint4 cur[NP];
    int eCur = -1;
    const size_t planeStride = P.planeStride;
    auto load_tile = [&](int c0, uint4 (&q)[NP], int& e) {
        const int c = c0 + pix;
        e = -1;
        if (c < W && chunkOk) {
            const uint8_t* src = srcLane + (size_t)c * Dp;
#pragma unroll
            for (int r = 0; r < NP; ++r) { q[r] = __ldcs(reinterpret_cast<const uint4*>(src)); src += planeStride; }
            if (P.hasSide) e = __ldg(P.entryOf + rowBase + c);
        }
    };

    // The row is walked from its RIGHT end to the left: a right-view pixel x needs the columns x + dmin .. x + dmin + D - 1,
    // so with this order it is complete as soon as its own tile has been summed - every tile finishes TW left-view and TW
    // right-view pixels, and both halves of the CTA have the same amount of work in every tile (walking left to right made
    // half of the warps idle for the first D columns and left them five tiles' worth of pixels after the last one).
    const int lastStart = ((W - 1) / TW) * TW;    // first tile processed = rightmost
    int slotTile = lastStart % RB;                // c0 % RB
    load_tile(lastStart, cur, eCur);
    for (int c0 = lastStart; c0 >= 0; c0 -= TW) {
        const int cols = min(TW, W - c0);
        const bool firstTile = (c0 == lastStart);
        // ------------------------------------------------------------------ phase A
        const int c = c0 + pix;
        if (c < W && chunkOk) {
            uint32_t ev[4] = {0, 0, 0, 0}, od[4] = {0, 0, 0, 0};     // even / odd disparities as 16-bit fields
#pragma unroll
            for (int r = 0; r < NP; ++r) {
                const uint32_t w[4] = {cur[r].x, cur[r].y, cur[r].z, cur[r].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ev[i] += w[i] & 0x00FF00FFu;
                    od[i] += __byte_perm(w[i], 0, 0x4341);
                }
            }
            uint32_t out[8];                                        // natural order: (d0,d1),(d2,d3),...
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                out[2 * i] = __byte_perm(ev[i], od[i], 0x5410);
                out[2 * i + 1] = __byte_perm(ev[i], od[i], 0x7632);
            }
            if (eCur >= 0) {
                const uint4* sp = reinterpret_cast<const uint4*>(P.side + (size_t)eCur * Dp + 16 * v);
                const uint4 a = __ldg(sp), b = __ldg(sp + 1);
                out[0] += a.x; out[1] += a.y; out[2] += a.z; out[3] += a.w;
                out[4] += b.x; out[5] += b.y; out[6] += b.z; out[7] += b.w;
            }
            if (16 * v + 16 > D) {                                  // padding disparities never win
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int d = 16 * v + 2 * i;
                    if (d >= D) out[i] |= 0x0000FFFFu;
                    if (d + 1 >= D) out[i] |= 0xFFFF0000u;
                }
            }
            int slot = slotTile + pix; if (slot >= RB) slot -= RB;
            uint32_t* dst = reinterpret_cast<uint32_t*>(ring + slot * RS + 16 * v);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = out[i];
            if (TAPS && P.S) {
                uint16_t* g = P.S + (rowBase + c) * D + 16 * v;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int d = 16 * v + 2 * i;
                    if (d < D) g[2 * i] = (uint16_t)(out[i] & 0xFFFFu);
                    if (d + 1 < D) g[2 * i + 1] = (uint16_t)(out[i] >> 16);
                }
            }
        }
        // the loads of the next tile fly during phase B (issued only now: the sums above must not wait on them)
        if (c0 > 0) load_tile(c0 - TW, cur, eCur);
        __syncthreads();
        // ------------------------------------------------------------------ phase B
        if (!rightRole) {
            for (int g0 = 0; g0 < cols; g0 += GROUPS) {                // uniform per warp: shuffles inside
                const int gi = g0 + grp;
                WtaPair w{0xFFFFFFFFu, 0xFFFFFFFFu};
                int slot = slotTile + gi; if (slot >= RB) slot -= RB;
                const uint16_t* row = ring + slot * RS;
                if (gi < cols) {
                    const uint32_t* q = reinterpret_cast<const uint32_t*>(row + k0);
                    const int n2 = max(0, min(CH, D - k0)) / 2;        // whole pairs; an odd tail is handled below
#pragma unroll 4
                    for (int i = 0; i < n2; ++i) {
                        const uint32_t pr = q[i];
                        wta_push2(w, pr * 65536u + (uint32_t)(k0 + 2 * i), (pr & 0xFFFF0000u) | (uint32_t)(k0 + 2 * i + 1));
                    }
                    if (k0 + 2 * n2 < min(k0 + CH, D)) wta_push1(w, (uint32_t)row[k0 + 2 * n2] * 65536u + (uint32_t)(k0 + 2 * n2));
                }
                wta_reduce4(w);
                if (sub == 0 && gi < cols) {
                    const int best = (int)(w.kmin & 0xFFFFu);
                    const uint32_t c1 = best > 0 ? row[best - 1] : 0u, c2 = best + 1 < D ? row[best + 1] : 0xFFFFu;
                    recL[c0 + gi] = make_uint4(w.kmin, w.ksec, c1 | (c2 << 16), 0u);
                }
            }
        } else if (P.checkLR) {
            // right pixels whose first column x + dmin lies in this tile (all their other columns are further right and
            // already summed); the rightmost tile also takes the pixels without any candidate (x + dmin >= W)
            const int rightBegin = max(0, c0 - P.dmin);
            const int rightEnd = firstTile ? W : min(W, c0 + cols - P.dmin);
            for (int x0 = rightBegin; x0 < rightEnd; x0 += GROUPS) {
                const int x = x0 + grp;
                const int first = x + P.dmin;                         // left column of disparity index 0
                const int n = (x < rightEnd) ? max(0, min(D, W - first)) : 0;
                WtaPair w{0xFFFFFFFFu, 0xFFFFFFFFu};
                int slot0 = slotTile + (first - c0);                  // first % RB; 0 <= first - c0 < TW wherever n > 0
                slot0 -= (slot0 >= RB) ? RB : 0;
                {
                    const int kEnd = min(k0 + CH, n);
                    int slot = slot0 + k0; if (slot >= RB) slot -= RB;
                    const int run = min(kEnd - k0, RB - slot);         // elements before the ring wraps
                    const uint16_t* q = ring + slot * RS + k0;
                    int i = 0;
#pragma unroll 4
                    for (; i < run; ++i) wta_push1(w, (uint32_t)q[i * (RS + 1)] * 65536u + (uint32_t)(k0 + i));
                    q -= RB * RS;
                    for (; k0 + i < kEnd; ++i) wta_push1(w, (uint32_t)q[i * (RS + 1)] * 65536u + (uint32_t)(k0 + i));
                }
                wta_reduce4(w);
                if (sub == 0 && x < rightEnd) {
                    const int best = (int)(w.kmin & 0xFFFFu);
                    auto cost = [&](int k) -> uint32_t {
                        if (k < 0 || k >= n) return 0xFFFFu;
                        int sl = slot0 + k; if (sl >= RB) sl -= RB;
                        return ring[sl * RS + k];
                    };
                    recR[x] = make_uint4(w.kmin, w.ksec, cost(best - 1) | (cost(best + 1) << 16), 0u);
                }
            }
        }
        slotTile -= TW; if (slotTile < 0) slotTile += RB;
    }
    // ---------------------------------------------------------------------- finish whole pixels from the records
    __syncthreads();
    for (int x = tid; x < W; x += THREADS) {
        const float d = wta_finish(recL[x], D, P.dmin, P.checkUnique, P.oneMinusRatio);
        rowL[x] = d;
        if (TAPS && P.dispLeftWta) P.dispLeftWta[rowBase + x] = d;
        if (P.checkLR) {
            const float r = wta_finish(recR[x], D, P.dmin, P.checkUnique, P.oneMinusRatio);
            rowR[x] = r;
            if (TAPS && P.dispRight) P.dispRight[rowBase + x] = r;
        }
    }
    if (!P.checkLR) return;
    // ---------------------------------------------------------------------- LR check (:445-470)
    __syncthreads();
    for (int x = tid; x < W; x += THREADS) {
        const float d = rowL[x];
        if (d != sgm_invalid()) {
            const float shifted = __fsub_rn((float)x, d);
            const int xr = __double2int_rz(__dadd_rn((double)shifted, 0.5));
            bool keep = true;
            if (xr < 0 || xr >= W) keep = false;
            else {
                const float dr = rowR[xr];
                if (dr != sgm_invalid() && fabsf(__fsub_rn(d, dr)) > P.lrThres) keep = false;
            }
            if (!keep) rowL[x] = sgm_invalid();
        }
    }
}

}  // namespace sgmb
