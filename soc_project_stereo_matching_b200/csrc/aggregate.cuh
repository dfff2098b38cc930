// aggregate.cuh -- K2: multi-path cost aggregation, one warp per path, disparities across lanes.
//
// Restates CostAggregate() (SemiGlobalMatching.c:229-372) with the matching cost of ComputeCost()/
// Hamming32() (SemiGlobalMatching.c:161-196) computed on the fly as popc(xor), so C(p,d) is never
// written to memory:
//   first pixel of a path :  L = C,  minPrev = min_d C                                   (:266-275)
//   later pixels          :  L[d] = (uint8)( C + min(Lp[d], Lp[d-1]+P1, Lp[d+1]+P1,
//                                              minPrev + max(P1, P2_init/(|g-gPrev|+1))) - minPrev )
//                            with Lp[-1] = Lp[D] = 255, uint16 candidates, wrap mod 256  (:325-353)
//
// Mapping: lane l owns the 2*NR consecutive disparity indices [2*NR*l, 2*NR*(l+1)) as NR registers of
// two 16-bit fields (index 2r in the low half of register r).  The 16-bit fields make every DP step a
// native DPX instruction (VIADDMNMX.U16x2 / VIMNMX.U16x2; the 8-bit SIMD intrinsics __vminu4/__vaddus4
// are emulated with 5-8 ALU ops each on sm_100a) while values stay <= 255, so the reference's mod-256
// wrap is one AND.  Lp[d-1]/Lp[d+1] across lane edges come from __shfl_up_sync/__shfl_down_sync, min_d
// from __reduce_min_sync; all path state lives in registers.  Candidates are allowed to exceed 255
// (they can never win against Lp[d] <= 255), only the final C + m - minPrev is wrapped.
//
// Output: each direction r owns a uint8 plane [N][Dp]; the warp stores L_r(p, .) with one 64*NR-byte
// coalesced store per pixel.  No direction reads or modifies another one's data, so all directions
// run concurrently in ONE launch without atomics; K3 sums the planes into the uint16 S while it
// computes the disparities.  Paths flagged irregular (the reference's anomalous diagonal walks, which
// visit some pixels twice) add their L into a small uint16 side buffer addressed through entryOf[].
#pragma once

#include <stdint.h>
#include "path_walker.h"

namespace sgmb {

struct PathWork {           // one warp's job
    int path;               // path index inside its direction
    uint8_t dir;            // 0..7, order of SemiGlobalMatching.c:213-220
    uint8_t irregular;      // 1: walk leaves the toroidal diagonal -> results go to the side buffer
    uint16_t pad;
};

struct AggParams {
    const uint8_t* img;         // left image [N]
    const uint32_t* censusL;    // [N]
    const uint32_t* censusR4;   // [4][copyStride], see census.cuh
    size_t copyStride;
    int padF;
    uint8_t* planes;            // [8][planeStride]
    size_t planeStride;         // bytes per plane = N * Dp
    uint32_t* side;             // uint16 [E][Dp] viewed as packed pairs, zeroed per frame
    const int32_t* entryOf;     // [N] -> side-buffer entry or -1
    const uint16_t* p2tab;      // [256] min(256, max(P1, P2_init/(delta+1)))
    const PathWork* work;
    int nWork;
    int W, H, D, Dp, dmin;
    int p1;                     // min(P1, 256)
};

constexpr int kAggWarpsPerBlock = 4;

template <int NR>
struct StepInput {
    uint32_t v[2 * NR];   // right census descriptors: v[j] = cR[q_top - (2*NR-1) + j]
    uint32_t cl;          // left census descriptor of the pixel
    uint32_t g;           // grey value of the pixel
    int pos, tcol;
    bool inside;
};

template <int NR>
__device__ __forceinline__ void load_step(const AggParams& P, const PathWalker& wk, int lane, StepInput<NR>& in)
{
    constexpr int DPL = 2 * NR;
    constexpr int VEC = DPL < 4 ? DPL : 4;
    in.pos = wk.pos; in.tcol = wk.tcol; in.inside = wk.inside();
    if (!in.inside) return;
    in.g = __ldg(P.img + wk.pos);
    in.cl = __ldg(P.censusL + wk.pos);
    // window cR[q_top-(DPL-1) .. q_top], q_top = pos - dmin - DPL*lane, fetched from the copy that makes it aligned
    const int y0 = wk.pos - P.dmin - DPL * lane - (DPL - 1) + P.padF;
    const int a = (-y0) & (VEC - 1);
    const uint32_t* src = P.censusR4 + (size_t)a * P.copyStride + (y0 + a);
    if (VEC == 2) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(src));
        in.v[0] = t.x; in.v[1] = t.y;
    } else {
#pragma unroll
        for (int j = 0; j < DPL / 4; ++j) {
            const uint4 t = __ldg(reinterpret_cast<const uint4*>(src) + j);
            in.v[4 * j + 0] = t.x; in.v[4 * j + 1] = t.y; in.v[4 * j + 2] = t.z; in.v[4 * j + 3] = t.w;
        }
    }
}

template <int NR>
__global__ void __launch_bounds__(kAggWarpsPerBlock * 32)
sgm_aggregate_paths(AggParams P)
{
    constexpr int DPL = 2 * NR;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint16_t s_p2[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_p2[i] = P.p2tab[i];
    __syncthreads();

    const int widx = blockIdx.x * kAggWarpsPerBlock + (threadIdx.x >> 5);
    if (widx >= P.nWork) return;
    const int lane = threadIdx.x & 31;
    const PathWork job = P.work[widx];
    const Dir dir = direction(job.dir);

    // 0x00FF in every 16-bit field whose disparity index is >= D: those fields are pinned to 255, which is
    // exactly the reference's Lp[D] = 255 sentinel (SemiGlobalMatching.c:260-263,357) for the last real d.
    uint32_t padm[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int i0 = DPL * lane + 2 * r;
        padm[r] = (i0 >= P.D ? 0x000000FFu : 0u) | (i0 + 1 >= P.D ? 0x00FF0000u : 0u);
    }
    const uint32_t p1x2 = (uint32_t)P.p1 * 0x00010001u;
    const int dbase = P.dmin + DPL * lane;              // absolute disparity of this lane's first index
    const int dmax_warp = P.dmin + 64 * NR - 1;         // largest absolute disparity any lane may hold

    uint8_t* plane = P.planes + (size_t)job.dir * P.planeStride;
    const bool lane_stores = DPL * lane < P.Dp;

    PathWalker wk;
    wk.start(P.W, P.H, dir.dx, dir.dy, job.path);
    const int len = wk.length();

    StepInput<NR> cur, nxt;
    load_step<NR>(P, wk, lane, cur);
    nxt = cur;

    uint32_t L[NR];
    uint32_t minPrev = 255, gPrev = 0;
    bool first = true;

    for (int s = 0; s < len; ++s) {
        if (s + 1 < len) {               // issue the next pixel's loads before this pixel's dependent chain
            wk.advance();
            load_step<NR>(P, wk, lane, nxt);
        }
        if (cur.inside) {
            // ---- matching cost C(p, d) = popc(cl ^ cR[p - d]), 127 where the right column would be < 0 (:170-177)
            uint32_t c[DPL];
#pragma unroll
            for (int k = 0; k < DPL; ++k) c[k] = __popc(cur.cl ^ cur.v[DPL - 1 - k]);
            if (cur.tcol < dmax_warp) {  // warp-uniform: only the first columns of a row need the per-disparity test
#pragma unroll
                for (int k = 0; k < DPL; ++k) c[k] = (dbase + k > cur.tcol) ? 127u : c[k];
            }
            uint32_t C[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) C[r] = c[2 * r] | (c[2 * r + 1] << 16);

            if (first) {
#pragma unroll
                for (int r = 0; r < NR; ++r) L[r] = C[r] | padm[r];
                first = false;
            } else {
                int dg = (int)cur.g - (int)gPrev;
                dg = dg < 0 ? -dg : dg;
                const uint32_t p2x2 = (uint32_t)s_p2[dg] * 0x00010001u;                 // min(256, max(P1, P2/(dg+1)))
                const uint32_t negmin = ((0u - minPrev) & 0xFFFFu) * 0x00010001u;       // -minPrev in both fields
                const uint32_t minx2 = minPrev * 0x00010001u; (void)minx2; (void)negmin;
                uint32_t up = __shfl_up_sync(FULL, L[NR - 1], 1);
                uint32_t dn = __shfl_down_sync(FULL, L[0], 1);
                if (lane == 0) up = 0x00FF00FFu;      // Lp[-1] = 255
                if (lane == 31) dn = 0x00FF00FFu;     // Lp[64*NR] = 255
                uint32_t Ln[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const uint32_t below = (r == 0) ? up : L[r - 1];
                    const uint32_t above = (r == NR - 1) ? dn : L[r + 1];
                    const uint32_t lm1 = __byte_perm(below, L[r], 0x5432);   // (Lp[d-1]) for both fields
                    const uint32_t lp1 = __byte_perm(L[r], above, 0x5432);   // (Lp[d+1]) for both fields
                    uint32_t t = __viaddmin_u16x2(lm1, p1x2, L[r]);          // min(Lp[d-1]+P1, Lp[d])
                    t = __viaddmin_u16x2(lp1, p1x2, t);                      // min(Lp[d+1]+P1, .)
#ifdef SGMB_NO_WRAPPING_ADDMIN
                    t = __vsub2(__vminu2(t, __vadd2(p2x2, minx2)), minx2);   // same value without relying on the 16-bit wrap
#else
                    t = __viaddmin_u16x2(t, negmin, p2x2);                   // min(. - minPrev, P2')  (>= 0, <= 255)
#endif
                    Ln[r] = (__vadd2(C[r], t) & 0x00FF00FFu) | padm[r];      // (uint8)(C + m - minPrev)
                }
#pragma unroll
                for (int r = 0; r < NR; ++r) L[r] = Ln[r];
            }
            // ---- min over all disparities of the new L (:347,353)
            uint32_t m = L[0];
#pragma unroll
            for (int r = 1; r < NR; ++r) m = __vminu2(m, L[r]);
            m = min(m & 0xFFFFu, m >> 16);
            minPrev = __reduce_min_sync(FULL, m);
            gPrev = cur.g;

            // ---- emit L_r(p, .)
            if (!job.irregular) {
                if (lane_stores) {
                    uint8_t* dst = plane + (size_t)cur.pos * P.Dp + DPL * lane;
                    if (NR == 1) {
                        *reinterpret_cast<uint16_t*>(dst) = (uint16_t)__byte_perm(L[0], 0, 0x4420);
                    } else if (NR == 2) {
                        *reinterpret_cast<uint32_t*>(dst) = __byte_perm(L[0], L[1], 0x6420);
                    } else {
                        *reinterpret_cast<uint2*>(dst) = make_uint2(__byte_perm(L[0], L[1], 0x6420),
                                                                    __byte_perm(L[NR - 2], L[NR - 1], 0x6420));
                    }
                }
            } else {
                const int e = __ldg(P.entryOf + cur.pos);
                if (e >= 0 && lane_stores) {
                    uint32_t* dst = P.side + ((size_t)e * P.Dp + DPL * lane) / 2;
#pragma unroll
                    for (int r = 0; r < NR; ++r) atomicAdd(dst + r, L[r] & ~padm[r]);
                }
            }
        }
        cur = nxt;
    }
}

}  // namespace sgmb
