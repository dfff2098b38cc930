// wta.cuh -- K3: sum the per-direction planes into S (uint16) and, in the same pass over the data,
// select disparities for BOTH views, refine them to sub-pixel precision and run the left-right check.
//
// Restates ComputeDisparity() (SemiGlobalMatching.c:374-443, both `inverse` modes) and LRCheck()
// (SemiGlobalMatching.c:445-470).  S(p,d) = sum_r L_r(p,d) is exactly the reference's cost_aggr after
// CostAggregation() (SemiGlobalMatching.c:198-221,345); it only ever exists in shared memory unless the
// taps are enabled, in which case it is also written out in the reference layout [N][D].
//
// One CTA owns one image row and streams it left to right in tiles of TW columns:
//   sum phase : all threads load the P planes of the tile with 128-bit loads (P independent loads in
//               flight per thread), add them as packed 16-bit fields, add the side-buffer contribution of
//               irregular paths, and store S into a shared-memory ring of TW + D columns;
//   WTA phase : thread t < TW scans the D costs of left pixel (c0 + t); thread TW + t scans those of a right
//               pixel whose last contributing column (x + D - 1) has just arrived: the right view reads the
//               ring along the diagonal S[x + d][d] (SemiGlobalMatching.c:397-399).  One thread per pixel with
//               a sequential scan needs ~4 instructions per cost and no cross-lane reduction;
//   epilogue  : both disparity rows sit in shared memory, so LRCheck runs in the same kernel.
// The ring row stride (Dp + 2 halfwords, an odd number of 32-bit words) makes row-wise (left view) and
// diagonal (right view) accesses of consecutive threads hit distinct banks.
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
    int W, H, D, Dp, dmin;
    int checkUnique;  float oneMinusRatio;   // (1 - uniqueness_ratio), formed in float like the reference
    int checkLR;      float lrThres;
    int ringCols;               // TW + D
};

__device__ __forceinline__ float sgm_invalid() { return __int_as_float(0x7f800000); }

// Scan `n` costs cost(d) = base[d * stride] (d = 0..n-1; costs with d >= n count as 65535), apply the
// uniqueness / border tests and the sub-pixel fit.  Lowest d wins ties because (cost << 16 | d) is
// minimised; second = second order statistic of the multiset == min over d != best (:412-419).
template <typename Fetch>
__device__ __forceinline__ float wta_scan(Fetch fetch, int n, int D, int dmin, int checkUnique, float oneMinusRatio)
{
    uint32_t kmin = 0xFFFFFFFFu, second = 0xFFFFu;
#pragma unroll 4
    for (int d = 0; d < n; ++d) {
        const uint32_t s = fetch(d);
        second = min(second, max(s, kmin >> 16));
        kmin = min(kmin, (s << 16) | (uint32_t)d);
    }
    const int best = (int)(kmin & 0xFFFFu);
    const int cmin = (int)(kmin >> 16);
    if (cmin == 0xFFFF) return sgm_invalid();          // no candidate at all (right view, x + dmin >= W)
    if (checkUnique) {
        const float lim = __fmul_rn((float)cmin, oneMinusRatio);
        const int ilim = (int)(__float2uint_rz(lim) & 0xFFFFu);            // (uint16_t)(min * (1 - ratio))  :422
        if ((int)second - cmin <= ilim) return sgm_invalid();
    }
    if (best == 0 || best == D - 1) return sgm_invalid();                  // :428-431
    const int c1 = (int)(int16_t)(uint16_t)fetch(best - 1);               // :434
    const int c2 = (int)(int16_t)(uint16_t)((best + 1 < n) ? fetch(best + 1) : 0xFFFFu);   // :435 (65535 -> -1)
    int denom = (int)(int16_t)(c1 + c2 - 2 * cmin);                        // :437
    if (denom < 1) denom = 1;
    return __fadd_rn((float)(best + dmin), __fdiv_rn((float)(c1 - c2), __fmul_rn((float)denom, 2.0f)));   // :440
}

template <int TW>
__global__ void __launch_bounds__(2 * TW)
sgm_reduce_wta_lr(WtaParams P)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int W = P.W, D = P.D, Dp = P.Dp;
    const int RS = Dp + 2;                        // ring row stride in halfwords
    const int RB = P.ringCols;
    uint16_t* ring = reinterpret_cast<uint16_t*>(smem_raw);
    float* rowL = reinterpret_cast<float*>(smem_raw + (((size_t)RB * RS * 2 + 15) & ~(size_t)15));
    float* rowR = rowL + W;

    const int y = blockIdx.x;
    const size_t rowBase = (size_t)y * W;
    const int tid = threadIdx.x;
    const int vecPerPix = Dp / 16;
    int rightDone = 0;                            // right pixels [0, rightDone) are finished

    for (int c0 = 0; c0 < W; c0 += TW) {
        const int cols = min(TW, W - c0);
        // ------------------------------------------------------------------ sum phase
        for (int task = tid; task < cols * vecPerPix; task += 2 * TW) {
            const int cl = task / vecPerPix, v = task - cl * vecPerPix;
            const size_t p = rowBase + c0 + cl;
            const uint8_t* src = P.planes + p * Dp + 16 * v;
            uint32_t ev[4] = {0, 0, 0, 0}, od[4] = {0, 0, 0, 0};     // even / odd disparities as 16-bit fields
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (r < P.nPlanes) {
                    const uint4 q = __ldcs(reinterpret_cast<const uint4*>(src + (size_t)r * P.planeStride));
                    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ev[i] += w[i] & 0x00FF00FFu;
                        od[i] += __byte_perm(w[i], 0, 0x4341);
                    }
                }
            }
            uint32_t out[8];                                        // natural order: (d0,d1),(d2,d3),...
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                out[2 * i] = __byte_perm(ev[i], od[i], 0x5410);
                out[2 * i + 1] = __byte_perm(ev[i], od[i], 0x7632);
            }
            if (P.hasSide) {
                const int e = __ldg(P.entryOf + p);
                if (e >= 0) {
                    const uint4* sp = reinterpret_cast<const uint4*>(P.side + (size_t)e * Dp + 16 * v);
                    const uint4 a = __ldg(sp), b = __ldg(sp + 1);
                    out[0] += a.x; out[1] += a.y; out[2] += a.z; out[3] += a.w;
                    out[4] += b.x; out[5] += b.y; out[6] += b.z; out[7] += b.w;
                }
            }
            if (16 * v + 16 > D) {                                  // padding disparities never win
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int d = 16 * v + 2 * i;
                    if (d >= D) out[i] |= 0x0000FFFFu;
                    if (d + 1 >= D) out[i] |= 0xFFFF0000u;
                }
            }
            uint32_t* dst = reinterpret_cast<uint32_t*>(ring + (size_t)((c0 + cl) % RB) * RS + 16 * v);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = out[i];
            if (P.S) {
                uint16_t* g = P.S + p * D + 16 * v;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int d = 16 * v + 2 * i;
                    if (d < D) g[2 * i] = (uint16_t)(out[i] & 0xFFFFu);
                    if (d + 1 < D) g[2 * i + 1] = (uint16_t)(out[i] >> 16);
                }
            }
        }
        __syncthreads();
        // ------------------------------------------------------------------ WTA phase
        const bool lastTile = (c0 + TW >= W);
        if (tid < TW) {
            if (tid < cols) {
                const uint16_t* base = ring + (size_t)((c0 + tid) % RB) * RS;
                const float d = wta_scan([&](int k) { return (uint32_t)base[k]; }, D, D, P.dmin, P.checkUnique, P.oneMinusRatio);
                rowL[c0 + tid] = d;
                if (P.dispLeftWta) P.dispLeftWta[rowBase + c0 + tid] = d;
            }
        } else if (P.checkLR) {
            // right pixel x is complete once column x + dmin + D - 1 has been summed (or the row has ended)
            const int rightEnd = lastTile ? W : max(0, c0 + cols - (P.dmin + D - 1));
            for (int x = rightDone + (tid - TW); x < rightEnd; x += TW) {
                const int first = x + P.dmin;                       // left column of disparity index 0
                const int n = max(0, min(D, W - first));
                int slot = first % RB;
                const float d = wta_scan(
                    [&](int k) {
                        int s = slot + k; if (s >= RB) s -= RB;
                        return (uint32_t)ring[(size_t)s * RS + k];
                    }, n, D, P.dmin, P.checkUnique, P.oneMinusRatio);
                rowR[x] = d;
                if (P.dispRight) P.dispRight[rowBase + x] = d;
            }
        }
        if (P.checkLR) rightDone = lastTile ? W : max(rightDone, max(0, c0 + cols - (P.dmin + D - 1)));
        __syncthreads();
    }
    // ---------------------------------------------------------------------- LR check (:445-470)
    for (int x = tid; x < W; x += 2 * TW) {
        float d = rowL[x];
        if (P.checkLR && d != sgm_invalid()) {
            const float shifted = __fsub_rn((float)x, d);
            const int xr = __double2int_rz(__dadd_rn((double)shifted, 0.5));
            if (xr < 0 || xr >= W) d = sgm_invalid();
            else {
                const float dr = rowR[xr];
                if (dr != sgm_invalid() && fabsf(__fsub_rn(d, dr)) > P.lrThres) d = sgm_invalid();
            }
        }
        P.dispOut[rowBase + x] = d;
    }
}

}  // namespace sgmb
