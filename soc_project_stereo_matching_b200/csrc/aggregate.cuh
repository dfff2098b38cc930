// aggregate.cuh -- K2: multi-path cost aggregation; disparities across lanes, path state in registers.
//
// Restates CostAggregate() (SemiGlobalMatching.c:229-372) with the matching cost of ComputeCost()/
// Hamming32() (SemiGlobalMatching.c:161-196) computed on the fly as popc(xor), so C(p,d) is never
// written to memory:
//   first pixel of a path :  L = C,  minPrev = min_d C                                   (:266-275)
//   later pixels          :  L[d] = (uint8)( C + min(Lp[d], Lp[d-1]+P1, Lp[d+1]+P1,
//                                              minPrev + max(P1, P2_init/(|g-gPrev|+1))) - minPrev )
//                            with Lp[-1] = Lp[D] = 255, uint16 candidates, wrap mod 256  (:325-353)
//
// Mapping: a path is owned by a group of LPP lanes (32, 16 or 8); lane s of the group owns the 2*NR
// consecutive disparity indices [2*NR*s, 2*NR*(s+1)) as NR registers of two 16-bit fields (index 2r in
// the low half of register r).  16-bit fields make every DP step a native DPX instruction
// (VIADDMNMX.U16x2 / VIMNMX.U16x2; the 8-bit SIMD intrinsics __vminu4/__vaddus4 are emulated with 5-8
// ALU ops each on sm_100a) while values stay <= 255, so the reference's mod-256 wrap is one AND.
// Lp[d-1]/Lp[d+1] across lane edges come from __shfl_up_sync/__shfl_down_sync, min_d from
// __reduce_min_sync (LPP == 32) or a shuffle butterfly; all path state lives in registers.  Candidates may
// exceed 255 (they can never win against Lp[d] <= 255); only the final C + m - minPrev is wrapped.
//
// The kernel is bound by its instruction count times the latency of dependent instructions, not by HBM or by
// any pipe (ncu: profiles/), so the per-step overhead that does not depend on the
// number of disparities (position update, the loads, penalty lookup, the two shuffles, the min reduction, the
// store) is amortised by giving each lane MORE disparities and putting 32/LPP paths of the same direction in
// one warp.  The horizontal directions have only H paths of W steps each, so they keep more lanes per path.
// Templates: DT = census descriptor type (uint32_t: the reference's 5x5 census; desc64_t: the 9x7 extension),
// PAD = whether some lanes hold disparity slots beyond D (false for D = 64 / 128 / 256).
//
// Output: each direction r owns a uint8 plane [N][Dp]; a group stores L_r(p, .) with one coalesced store
// per pixel.  No direction reads or modifies another one's data, so all directions run concurrently in ONE
// launch without atomics; K3 sums the planes into the uint16 S while it computes the disparities.  Paths
// flagged irregular (the reference's anomalous diagonal walks, which visit some pixels twice and are
// followed with the generic walker of path_walker.h) add their L into a small uint16 side buffer addressed
// through entryOf[].
#pragma once

#include <stdint.h>
#include "path_walker.h"

namespace sgmb {

struct WarpWork {           // one warp's job
    int firstPath;          // first path index inside its direction
    uint8_t dir;            // 0..7, order of SemiGlobalMatching.c:213-220
    uint8_t count;          // paths handled by this warp (<= 32 / LPP; 1 for irregular paths)
    uint16_t pad;
};

// Census descriptors are uint32_t (5x5 window, the reference's census_transform_5x5) or 64-bit words (9x7
// window extension); the kernels are templates over the descriptor type DT.
typedef unsigned long long desc64_t;
__device__ __forceinline__ uint32_t desc_popc(uint32_t x) { return (uint32_t)__popc(x); }
__device__ __forceinline__ uint32_t desc_popc(desc64_t x) { return (uint32_t)__popcll(x); }

struct AggParams {
    const uint8_t* img;         // left image [N]
    const void* censusL;        // DT [N]
    const void* pixL;           // {left descriptor, grey value} per pixel - uint2 for 32-bit descriptors, uint4 {lo, hi, grey, 0} for
                                // 64-bit ones: what a column visit needs besides its right-census window, fetched with ONE vector
                                // load (see load_step)
    const void* censusR4;       // DT [32 / sizeof(DT)][copyStride], see census.cuh
    uint32_t copyStride;        // elements per copy (< 2^31)
    int padF;
    uint8_t* planes;            // [8][planeStride]
    size_t planeStride;         // bytes per plane = N * Dp
    uint32_t* side;             // uint16 [E][Dp] viewed as packed pairs, zeroed per frame
    const int32_t* entryOf;     // [N] -> side-buffer entry or -1
    const WarpWork* work;       // [nIrregularWarps] irregular jobs, then [nRegularWarps] regular jobs
    int nIrregularWarps, nRegularWarps;
    int W, H, D, Dp, dmin;
    int wrapInterior;           // 1: a visit without out-of-image costs can still exceed 255 (largest census cost + largest P2 > 255)
    uint32_t p1x2;              // min(P1, 256) in both 16-bit fields
    uint32_t p2x2[256];         // min(256, max(P1, P2_init/(delta+1))) in both fields, indexed by |g - gPrev|
};

#ifndef SGM_AGG_WPB
#define SGM_AGG_WPB 4
#endif
constexpr int kAggWarpsPerBlock = SGM_AGG_WPB;

// ------------------------------------------------------------------------------------------------ shared pieces
template <int NR, typename DT>
struct StepInput {
    DT v[2 * NR];         // right census descriptors: v[j] = cR[q_top - (2*NR-1) + j]
    DT cl;                // left census descriptor of the pixel
    uint32_t g;           // grey value of the pixel
};

// Loads of one pixel visit: grey value, left descriptor and the window cR[q_top-(DPL-1) .. q_top] with
// q_top = pos - dmin - DPL*sub, fetched from the copy that makes the window 8/16-byte aligned.
template <int NR, typename DT>
__device__ __forceinline__ void load_step(const AggParams& P, uint32_t pos, int sub, StepInput<NR, DT>& in)
{
    constexpr int DPL = 2 * NR;
    constexpr int PER32 = 32 / (int)sizeof(DT);           // descriptors per 256-bit load
    constexpr int VEC = DPL < PER32 ? DPL : PER32;        // alignment unit of the window, in descriptors
    if constexpr (sizeof(DT) == 4) {
        // One 64-bit load instead of two scalar ones.  With scalar loads ptxas let the load overwrite its own address
        // register and parked the value with a register move a few instructions later - a full L2 latency on the
        // in-order issue stream (two such moves carried 17 % of the kernel's stall samples, ncu source view of round
        // r1_e); a vector destination stays where it is loaded.  What remains (8 % of the samples on the first use of
        // one of the three buffers) is structural: ptxas tracks all prefetch loads of the loop with ONE hardware
        // scoreboard - the other five rotate among POPCs, shuffles and table lookups (decoded from the control words)
        // - so the wait before a buffer's first use also waits for the request issued ~70 instructions earlier for
        // another buffer.  Measured alternatives, all slower or equal: two buffers with the request right behind the
        // wait (0.288 ms), reloading a buffer inside its consuming visit three visits ahead (0.284 ms), staging through
        // shared memory with cp.async (0.409 ms: the L1 data pipe, already the busiest unit, then carries every
        // window three times); this version: 0.282 ms at C2.
        const uint2 px = __ldg(static_cast<const uint2*>(P.pixL) + pos);
        in.cl = px.x; in.g = px.y;
    } else {
        const uint4 px = __ldg(static_cast<const uint4*>(P.pixL) + pos);
        in.cl = ((desc64_t)px.y << 32) | px.x; in.g = px.z;
    }
    // The window of a lane is DPL consecutive descriptors, and the lanes of a path sit DPL descriptors apart, so ONE load
    // instruction of a warp touches (lanes) pieces that are DPL * sizeof(DT) bytes apart: the L1 data pipe needs a pass
    // ("wavefront") per 128-byte line touched, not per byte delivered.  With 128-bit loads a visit of the 8 x 16 layout cost
    // 4 instructions x ~17 lines; ncu (round 2, r2_y): l1tex data-pipe wavefronts 78 % of peak, 77 per visit, and the SMs that
    // hold only column-like warps finished last, in proportion to their wavefront count.  256-bit loads (LDG.E.256, new on
    // sm_100; the address must be 32-byte aligned, hence 32 / sizeof(DT) shifted copies of the right census instead of
    // 16 / sizeof(DT)) halve the instructions while each still touches the same lines: ~36 wavefronts per visit.
    const uint32_t y0 = pos - (uint32_t)(P.dmin + DPL * sub + (DPL - 1)) + (uint32_t)P.padF;   // >= 0 by the front padding
    const uint32_t al = (y0 + (VEC - 1)) & ~(uint32_t)(VEC - 1);
    const DT* src = static_cast<const DT*>(P.censusR4) + ((al - y0) * P.copyStride + al);
    if constexpr (sizeof(DT) == 4 && VEC == 2) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(src));
        in.v[0] = t.x; in.v[1] = t.y;
    } else if constexpr (sizeof(DT) == 4 && VEC == 4) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(src));
        in.v[0] = t.x; in.v[1] = t.y; in.v[2] = t.z; in.v[3] = t.w;
    } else if constexpr (sizeof(DT) == 4) {
#pragma unroll
        for (int j = 0; j < DPL / 8; ++j)
            asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(in.v[8 * j]), "=r"(in.v[8 * j + 1]), "=r"(in.v[8 * j + 2]), "=r"(in.v[8 * j + 3]),
                           "=r"(in.v[8 * j + 4]), "=r"(in.v[8 * j + 5]), "=r"(in.v[8 * j + 6]), "=r"(in.v[8 * j + 7])
                         : "l"(src + 8 * j));
    } else if constexpr (VEC == 2) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(src));
        in.v[0] = ((desc64_t)t.y << 32) | t.x; in.v[1] = ((desc64_t)t.w << 32) | t.z;
    } else {
#pragma unroll
        for (int j = 0; j < DPL / 4; ++j)
            asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(in.v[4 * j]), "=l"(in.v[4 * j + 1]), "=l"(in.v[4 * j + 2]), "=l"(in.v[4 * j + 3]) : "l"(src + 4 * j));
    }
}

// Which two of the lane's 2*NR disparity indices share register r.  Natural pairing: (2r, 2r+1).  Paired ("strided") layout,
// used by all regular paths when every layout of the launch owns whole 8-disparity units (NRH and NRV multiples of 4):
// inside each unit of 8 indices register j of the unit holds (j, j + 4), so that the neighbours d-1 / d+1 of BOTH fields of
// a register are simply the previous / next register - only the first and last register of a unit need a PRMT (2 per unit
// instead of one per register: 4 instead of 9 per visit of the 8 x 16 layout, 2 instead of 5 per horizontal step).  The
// plane bytes are stored as the registers lie (store_plane is the same code): unit byte order 0,4,1,5,2,6,3,7; K3 sums the
// words as even / odd byte fields, which with this order ARE the natural pairs (wta.cuh), and SGMB_GetStage undoes it.
__host__ __device__ constexpr bool agg_paired_layout(int nrh, int nrv) { return nrh % 4 == 0 && nrv % 4 == 0; }
template <bool STR> __device__ __forceinline__ constexpr int pair_lo(int r) { return STR ? 8 * (r >> 2) + (r & 3) : 2 * r; }
template <bool STR> __device__ __forceinline__ constexpr int pair_hi(int r) { return STR ? 8 * (r >> 2) + (r & 3) + 4 : 2 * r + 1; }

// Matching cost of the lane's 2*NR disparities, packed two per register (SemiGlobalMatching.c:170-177).
// BORDER: the pixel is so close to the left image edge that some right columns are negative -> cost 127
// for disparity indices >= nvalid.
template <int NR, bool BORDER, typename DT, bool STR>
__device__ __forceinline__ void pack_cost(const StepInput<NR, DT>& in, int nvalid, uint32_t (&C)[NR])
{
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        uint32_t c0 = desc_popc(in.cl ^ in.v[2 * NR - 1 - pair_lo<STR>(r)]);
        uint32_t c1 = desc_popc(in.cl ^ in.v[2 * NR - 1 - pair_hi<STR>(r)]);
        if (BORDER) {
            c0 = (pair_lo<STR>(r) < nvalid) ? c0 : 127u;
            c1 = (pair_hi<STR>(r) < nvalid) ? c1 : 127u;
        }
        C[r] = __byte_perm(c0, c1, 0x5410);
    }
}

// One DP step on the lane's registers.  up / dn: the neighbouring lanes' edge registers (sentinel 0x00FF00FF
// at the ends of the disparity range).
// WRAP = false: the caller guarantees C + (m - minPrev) <= 255 in every field (no cost of 127 in this visit and
// max cost + largest P2 <= 255, AggParams::wrapInterior == 0), so the reference's uint8 truncation is the identity.
template <int NR, bool WRAP, bool STR>
__device__ __forceinline__ void dp_step(uint32_t (&L)[NR], const uint32_t (&C)[NR], const uint32_t (&padm)[NR],
                                        uint32_t up, uint32_t dn, uint32_t p1x2, uint32_t p2x2, uint32_t negmin)
{
    if constexpr (STR) {
        static_assert(NR % 4 == 0, "paired layout: whole units of 8 disparities");
        uint32_t o[NR];                                                    // the previous pixel's values (L is overwritten below)
#pragma unroll
        for (int r = 0; r < NR; ++r) o[r] = L[r];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int u = r >> 2, j = r & 3;
            // Lp[d-1] / Lp[d+1] of both fields: the neighbouring register, except at the ends of a unit, where the low field
            // continues in the previous unit's last register (high field) / the high field in the next unit's first (low field)
            const uint32_t lm1 = j > 0 ? o[r - 1] : __byte_perm(u == 0 ? up : o[4 * u - 1], o[4 * u + 3], 0x5432);
            const uint32_t lp1 = j < 3 ? o[r + 1] : __byte_perm(o[4 * u], u == NR / 4 - 1 ? dn : o[4 * u + 4], 0x5432);
            uint32_t t = __viaddmin_u16x2(lm1, p1x2, o[r]);
            t = __viaddmin_u16x2(lp1, p1x2, t);
            t = __viaddmin_u16x2(t, negmin, p2x2);
            const uint32_t sum = C[r] + t;
            L[r] = (WRAP ? (sum & 0x00FF00FFu) : sum) | padm[r];
        }
    } else {
        uint32_t lm1 = __byte_perm(up, L[0], 0x5432);                      // Lp[d-1] of both fields of register 0
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const uint32_t above = (r == NR - 1) ? dn : L[r + 1];
            const uint32_t lp1 = __byte_perm(L[r], above, 0x5432);         // Lp[d+1]; also Lp[d-1] of register r+1
            uint32_t t = __viaddmin_u16x2(lm1, p1x2, L[r]);                // min(Lp[d-1]+P1, Lp[d])
            t = __viaddmin_u16x2(lp1, p1x2, t);                            // min(Lp[d+1]+P1, .)
            t = __viaddmin_u16x2(t, negmin, p2x2);                         // min(. - minPrev, P2')   in [0, 255]
            const uint32_t sum = C[r] + t;                                 // fields <= 127 + 255: no carry between them
            L[r] = (WRAP ? (sum & 0x00FF00FFu) : sum) | padm[r];           // (uint8)(C + m - minPrev)
            lm1 = lp1;
        }
    }
}

template <int NR>
__device__ __forceinline__ uint32_t lane_min_x2(const uint32_t (&L)[NR])
{
    uint32_t m = L[0];
#pragma unroll
    for (int r = 1; r < NR; ++r) m = __vminu2(m, L[r]);
    return __vminu2(m, __byte_perm(m, 0, 0x1032));                         // both fields = min of the lane
}

template <int LPP>
__device__ __forceinline__ uint32_t group_min_x2(uint32_t m)               // m: both fields equal
{
    if (LPP == 32) return __reduce_min_sync(0xffffffffu, m);
    if (LPP == 16) {
        // two warp-wide CREDUX.MIN with the other half-warp masked to 0xFFFFFFFF: 44 cycles in a dependent chain
        // against 132 for the four-round shuffle butterfly (scripts/micro/groupmin.cu, measured on B200)
        const bool hi = (threadIdx.x & 16) != 0;
        const uint32_t a = __reduce_min_sync(0xffffffffu, hi ? 0xffffffffu : m);
        const uint32_t b = __reduce_min_sync(0xffffffffu, hi ? m : 0xffffffffu);
        return hi ? b : a;
    }
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) m = __vminu2(m, __shfl_xor_sync(0xffffffffu, m, o));
    return m;
}

template <int NR>
__device__ __forceinline__ void store_plane(uint8_t* dst, const uint32_t (&L)[NR])
{
    if (NR == 1) {
        __stcs(reinterpret_cast<unsigned short*>(dst), (unsigned short)__byte_perm(L[0], 0, 0x4420));
    } else if (NR == 2) {
        __stcs(reinterpret_cast<uint32_t*>(dst), __byte_perm(L[0], L[1], 0x6420));
    } else if (NR == 4) {
        __stcs(reinterpret_cast<uint2*>(dst), make_uint2(__byte_perm(L[0], L[1], 0x6420), __byte_perm(L[2], L[3], 0x6420)));
    } else {
#pragma unroll
        for (int j = 0; j < NR / 8; ++j)
            __stcs(reinterpret_cast<uint4*>(dst) + j, make_uint4(__byte_perm(L[8 * j + 0], L[8 * j + 1], 0x6420), __byte_perm(L[8 * j + 2], L[8 * j + 3], 0x6420),
                                                          __byte_perm(L[8 * j + 4], L[8 * j + 5], 0x6420), __byte_perm(L[8 * j + 6], L[8 * j + 7], 0x6420)));
    }
}

// ------------------------------------------------------------------------------------------------ horizontal paths
// 32/LPP paths (image rows) per warp, LPP lanes per path, 2*NR disparities per lane.  Only H paths exist per
// direction and each is W dependent steps long, so this loop is latency-critical.  It contains NO global load and
// is software pipelined: while the dependent chain of step s runs (shuffle -> DPX min/add -> group min), the
// inputs of step s+1 are prepared in the same basic block.
//  * every LPP steps the lanes of a path fetch the next LPP pixels of their row with three coalesced loads (lane j
//    of the group holds the grey value, the left descriptor and the one new right descriptor of step LPP*b+1+j;
//    fetched one block ahead) and every step broadcasts its pixel with a width-LPP __shfl_sync;
//  * the right-census window cR[x-d] slides by one element per step, so it lives in registers and is shifted
//    across the lanes of the group with one __shfl_up/down per step.
template <int NR, bool FWD, typename DT>
struct HorizontalState {
    DT w[2 * NR];            // w[k] = cR[x - dmin - DPL*sub - k] for the column x being prepared
    uint32_t L[NR];
    uint32_t C[NR];          // cost of the prepared step
    uint32_t p2x2;           // penalty of the prepared step
    uint32_t g;              // grey value of the prepared step
    uint32_t minx2;
};

// Prepare step with column x: broadcast its pixel from lane j of the group's block registers, slide the window, cost.
// ROT >= 0: the window is kept as a ring in its registers -- logical element k lives in register (k -+ t) mod DPL at
// chunk time t -- so that inside a fully unrolled chunk of DPL steps the slide costs no register moves; ROT is the
// chunk time BEFORE this slide (after DPL slides logical == physical again).  ROT < 0: plain shifting registers.
template <int NR, int LPP, bool FWD, bool BORDER, typename DT, bool STR>
__device__ __forceinline__ void horizontal_prepare(const AggParams& P, HorizontalState<NR, FWD, DT>& st, uint32_t gBlk, DT clBlk,
                                                   DT crBlk, int j, int x, int sub, int dbase, int ROT)
{
    constexpr int DPL = 2 * NR;
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t g = __shfl_sync(FULL, gBlk, j, LPP);
    const DT cl = __shfl_sync(FULL, clBlk, j, LPP);
    const DT fresh = __shfl_sync(FULL, crBlk, j, LPP);
    int base = 0;                                       // register of logical element 0 after the slide
    if (ROT < 0) {
        if (FWD) {
            DT t = __shfl_up_sync(FULL, st.w[DPL - 1], 1, LPP);
            if (sub == 0) t = fresh;
#pragma unroll
            for (int k = DPL - 1; k > 0; --k) st.w[k] = st.w[k - 1];
            st.w[0] = t;
        } else {
            DT t = __shfl_down_sync(FULL, st.w[0], 1, LPP);
            if (sub == LPP - 1) t = fresh;
#pragma unroll
            for (int k = 0; k < DPL - 1; ++k) st.w[k] = st.w[k + 1];
            st.w[DPL - 1] = t;
        }
    } else if (FWD) {
        const int last = (2 * DPL - 1 - ROT) % DPL;     // register of logical element DPL-1 at time ROT == of element 0 at ROT+1
        DT t = __shfl_up_sync(FULL, st.w[last], 1, LPP);
        if (sub == 0) t = fresh;
        st.w[last] = t;
        base = last;
    } else {
        const int first = ROT % DPL;                    // register of logical element 0 at time ROT == of element DPL-1 at ROT+1
        DT t = __shfl_down_sync(FULL, st.w[first], 1, LPP);
        if (sub == LPP - 1) t = fresh;
        st.w[first] = t;
        base = (first + 1) % DPL;
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        uint32_t c0 = desc_popc(cl ^ st.w[(base + pair_lo<STR>(r)) % DPL]);
        uint32_t c1 = desc_popc(cl ^ st.w[(base + pair_hi<STR>(r)) % DPL]);
        if (BORDER) {                                   // right column x - d < 0  ->  cost 127 (SemiGlobalMatching.c:170-172)
            c0 = (dbase + pair_lo<STR>(r) <= x) ? c0 : 127u;
            c1 = (dbase + pair_hi<STR>(r) <= x) ? c1 : 127u;
        }
        st.C[r] = __byte_perm(c0, c1, 0x5410);
    }
    int dg = (int)g - (int)st.g;
    dg = dg < 0 ? -dg : dg;
    st.p2x2 = P.p2x2[dg];
    st.g = g;
}

// n steps whose inputs come from one block of registers: step i of the block consumes the prepared inputs and
// prepares the following step from lane i of the group.  WRAP: see dp_step.
template <int NR, int LPP, bool FWD, bool BORDER, bool WRAP, typename DT, bool STR, bool FULL_BLOCK = false>
__device__ __forceinline__ void horizontal_block(const AggParams& P, HorizontalState<NR, FWD, DT>& st, const uint32_t (&padm)[NR], int n,
                                                 int xnext, uint32_t gBlk, DT clBlk, DT crBlk, int sub, int dbase,
                                                 uint8_t*& out, long long outStride, bool stores)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int DPL = 2 * NR;
    auto chain = [&]() {                                // dependent chain of the current step
        uint32_t Ccur[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) Ccur[r] = st.C[r];
        const uint32_t p2 = st.p2x2;
        uint32_t up = __shfl_up_sync(FULL, st.L[NR - 1], 1, LPP);
        uint32_t dn = __shfl_down_sync(FULL, st.L[0], 1, LPP);
        if (sub == 0) up = 0x00FF00FFu;
        if (sub == LPP - 1) dn = 0x00FF00FFu;
        dp_step<NR, WRAP, STR>(st.L, Ccur, padm, up, dn, P.p1x2, p2, __vneg2(st.minx2));
        st.minx2 = group_min_x2<LPP>(lane_min_x2<NR>(st.L));
        if (stores) store_plane<NR>(out, st.L);
        out += outStride;
    };
    if (DPL <= 8) {
        // chunks of DPL steps, fully unrolled: the window rotates through its registers without moves.  Only the last
        // block of a row can end inside a chunk (n < LPP), and nothing reads the window after it, so leaving the
        // chunk early there is harmless.  (Windows of 16 descriptors, D = 256: unrolling 16 steps costs more in
        // registers and code size than the moves it saves.)
        for (int i = 0; i < n; i += DPL) {
#pragma unroll
            for (int t = 0; t < DPL; ++t) {
                if (!FULL_BLOCK && i + t >= n) break;
                chain();
                // inputs of the next step (independent of the chain above)
                horizontal_prepare<NR, LPP, FWD, BORDER, DT, STR>(P, st, gBlk, clBlk, crBlk, i + t, FWD ? xnext + i + t : xnext - i - t, sub, dbase, t);
            }
        }
    } else {
        for (int i = 0; i < n; ++i) {
            chain();
            horizontal_prepare<NR, LPP, FWD, BORDER, DT, STR>(P, st, gBlk, clBlk, crBlk, i, FWD ? xnext + i : xnext - i, sub, dbase, -1);
        }
    }
}

template <int NR, int LPP, bool FWD, typename DT, bool PAD, bool STR>
__device__ __forceinline__ void aggregate_horizontal(const AggParams& P, const WarpWork job, int lane)
{
    constexpr int DPL = 2 * NR;
    constexpr unsigned FULL = 0xffffffffu;
    const int grp = lane / LPP, sub = lane % LPP;
    const bool active = grp < (int)job.count;
    const int W = P.W, row = job.firstPath + (active ? grp : (int)job.count - 1);   // idle groups shadow the last row
    const uint32_t rowBase = (uint32_t)row * (uint32_t)W;
    const DT* cR = static_cast<const DT*>(P.censusR4) + P.padF;     // copy 0: cR[p], zero padding in front
    const DT* cL = static_cast<const DT*>(P.censusL);
    uint32_t padm[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int i0 = DPL * sub + pair_lo<STR>(r), i1 = DPL * sub + pair_hi<STR>(r);
        padm[r] = PAD ? ((i0 >= P.D ? 0x000000FFu : 0u) | (i1 >= P.D ? 0x00FF0000u : 0u)) : 0u;
    }
    const int dbase = P.dmin + DPL * sub;
    const int dlast = P.dmin + LPP * DPL - 1;
    // without padding every lane holds real disparities; idle groups shadow the last path and store the same bytes again
    const bool stores = PAD ? (active && (DPL * sub < P.Dp)) : true;
    const long long outStride = FWD ? (long long)P.Dp : -(long long)P.Dp;
    uint8_t* out = P.planes + (size_t)job.dir * P.planeStride + ((size_t)rowBase + (FWD ? 0 : W - 1)) * P.Dp + DPL * sub;

    auto column = [&](int s) { return FWD ? s : W - 1 - s; };
    // block b holds steps LPP*b+1 .. LPP*b+LPP (step 0 is set up below): lane j of the group <-> step LPP*b + 1 + j
    auto load_block = [&](int b, uint32_t& gB, DT& clB, DT& crB) {
        const int s = LPP * b + 1 + sub;
        gB = 0; clB = 0; crB = 0;
        if (s < W) {
            const int x = column(s);
            gB = __ldg(P.img + rowBase + x);
            clB = __ldg(cL + rowBase + x);
            const int xin = FWD ? x - P.dmin : x - P.dmin - (LPP * DPL - 1);   // element entering the window on arrival at x
            if (xin >= 0) crB = __ldg(cR + rowBase + xin);
        }
    };

    HorizontalState<NR, FWD, DT> st;
    uint32_t gA, gB;
    DT clA, crA, clB, crB;
    load_block(0, gA, clA, crA);
    // ---- step 0: window, cost, L = C (SemiGlobalMatching.c:266-275)
    {
        const int x0 = column(0);
        const DT cl = __ldg(cL + rowBase + x0);
        st.g = __ldg(P.img + rowBase + x0);
#pragma unroll
        for (int k = 0; k < DPL; ++k) {
            const int xr = x0 - dbase - k;
            st.w[k] = (xr >= 0) ? __ldg(cR + rowBase + xr) : (DT)0;
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const uint32_t c0 = (dbase + pair_lo<STR>(r) <= x0) ? desc_popc(cl ^ st.w[pair_lo<STR>(r)]) : 127u;
            const uint32_t c1 = (dbase + pair_hi<STR>(r) <= x0) ? desc_popc(cl ^ st.w[pair_hi<STR>(r)]) : 127u;
            st.L[r] = (c1 * 65536u + c0) | padm[r];
        }
        st.minx2 = group_min_x2<LPP>(lane_min_x2<NR>(st.L));
        if (stores) store_plane<NR>(out, st.L);
        out += outStride;
    }
    if (W == 1) return;
    // prepare step 1 from lane 0 of block 0; afterwards every block iteration consumes one step and prepares the next
    horizontal_prepare<NR, LPP, FWD, true, DT, STR>(P, st, gA, clA, crA, 0, column(1), sub, dbase, -1);
    // steps 1 .. W-1: step s is consumed in block (s-1)/LPP at i = (s-1)%LPP, where step s+1 is prepared from lane i+1
    // of the same block, or lane 0 of the next one.  To keep one loop body, rotate the block registers by one lane:
    // consuming position i prepares from lane i of registers that hold steps LPP*b+2 .. LPP*b+LPP+1.
    const int nsteps = W - 1;                                        // steps still to consume
    int done = 0;
    auto shifted = [&](auto cur, auto nxt) {
        const auto v = __shfl_down_sync(FULL, cur, 1, LPP);
        const auto first = __shfl_sync(FULL, nxt, 0, LPP);
        return sub == LPP - 1 ? first : v;
    };
    bool prevBorder = true;                                          // step 1 was prepared with the border handling
    for (int b = 0; done < nsteps; ++b) {
        load_block(b + 1, gB, clB, crB);                             // one block ahead
        const uint32_t gS = shifted(gA, gB);
        const DT clS = shifted(clA, clB), crS = shifted(crA, crB);
        const int n = min(LPP, nsteps - done);                       // consume steps done+1 .. done+n, prepare done+2 .. done+n+1
        const int sPrepFirst = done + 2;
        const int xa = column(min(sPrepFirst, W - 1)), xb = column(min(sPrepFirst + n - 1, W - 1));
        const bool border = min(xa, xb) < dlast;                     // warp-uniform
        // the first step a block consumes was prepared by the previous block: no truncation only if neither saw the border
        // (two code variants only: the border one is also correct, just slower, for interior columns)
        const bool slow = border || prevBorder || P.wrapInterior != 0;
        if (slow) horizontal_block<NR, LPP, FWD, true, true, DT, STR>(P, st, padm, n, column(sPrepFirst), gS, clS, crS, sub, dbase, out, outStride, stores);
        else if (n == LPP && LPP % DPL == 0)   // a whole block of interior steps: no end-of-row test inside the unrolled chunk
                  horizontal_block<NR, LPP, FWD, false, false, DT, STR, true>(P, st, padm, n, column(sPrepFirst), gS, clS, crS, sub, dbase, out, outStride, stores);
        else      horizontal_block<NR, LPP, FWD, false, false, DT, STR>(P, st, padm, n, column(sPrepFirst), gS, clS, crS, sub, dbase, out, outStride, stores);
        prevBorder = border;
        done += n;
        gA = gB; clA = clB; crA = crB;
    }
}

// ------------------------------------------------------------------------------------------------ vertical / diagonal paths
// A regular vertical or diagonal path follows a column or the toroidal diagonal, so its position is updated
// with a constant stride (plus one column wrap for diagonals); 32/LPP paths of one direction share a warp and
// every lane owns 2*NR disparities.  The loads of visit s+1 are issued before the dependent chain of visit s
// (two input buffers, loop unrolled by two).
template <int NR, int LPP, bool DIAG, typename DT, bool PAD, bool STR>
__device__ __forceinline__ void aggregate_column_like(const AggParams& P, const WarpWork job, int lane)
{
    static_assert(NR == 1 || NR == 2 || NR == 4 || NR == 8, "NR");
    constexpr int DPL = 2 * NR;
    constexpr unsigned FULL = 0xffffffffu;
    const int grp = lane / LPP, sub = lane % LPP;
    const bool active = grp < (int)job.count;
    const int path = job.firstPath + (active ? grp : (int)job.count - 1);      // idle groups shadow the last path

    const Dir dir = direction(job.dir);
    const bool fwd = dir.dy > 0;
    const int W = P.W, H = P.H, len = H;
    int tcol = path;
    uint32_t pos = fwd ? (uint32_t)path : (uint32_t)((H - 1) * W + path);
    const uint32_t dpos = (uint32_t)(dir.dy * W + dir.dx);
    const int dcol = dir.dx;

    uint32_t padm[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int i0 = DPL * sub + pair_lo<STR>(r), i1 = DPL * sub + pair_hi<STR>(r);
        padm[r] = PAD ? ((i0 >= P.D ? 0x000000FFu : 0u) | (i1 >= P.D ? 0x00FF0000u : 0u)) : 0u;
    }
    const int dbase = P.dmin + DPL * sub;                 // absolute disparity of this lane's first index
    const int dlast = P.dmin + DPL * LPP - 1;             // largest absolute disparity the group may hold
    // without padding every lane holds real disparities; idle groups shadow the last path and store the same bytes again
    const bool stores = PAD ? (active && (DPL * sub < P.Dp)) : true;
    uint8_t* const planeLane = P.planes + (size_t)job.dir * P.planeStride + DPL * sub;
    const uint32_t p1x2 = P.p1x2;
    // vertical paths keep their column: whether the left image border matters is decided once per warp
    const bool colBorder = DIAG ? false : (__any_sync(FULL, tcol < dlast) != 0);

    uint32_t L[NR], C[NR];
    uint32_t minx2, gPrev, posPrev;

    auto advance = [&]() {
        pos += dpos;
        if (DIAG) {
            tcol += dcol;
            if (tcol >= W) { tcol -= W; pos -= (uint32_t)W; }
            if (tcol < 0)  { tcol += W; pos += (uint32_t)W; }
        }
    };
    auto cost = [&](const StepInput<NR, DT>& in, int tc) {
        const bool border = DIAG ? (__any_sync(FULL, tc < dlast) != 0) : colBorder;
        if (border) pack_cost<NR, true, DT, STR>(in, tc - dbase + 1, C);
        else pack_cost<NR, false, DT, STR>(in, 0, C);
    };
    auto visit = [&](const StepInput<NR, DT>& in, uint32_t p, int tc) {
        const bool slow = DIAG ? (__any_sync(FULL, tc < dlast) != 0) : colBorder;
        if (slow) pack_cost<NR, true, DT, STR>(in, tc - dbase + 1, C);
        else pack_cost<NR, false, DT, STR>(in, 0, C);
        int dg = (int)in.g - (int)gPrev;
        dg = dg < 0 ? -dg : dg;
        gPrev = in.g;
        const uint32_t p2x2 = P.p2x2[dg];
        const uint32_t negmin = __vneg2(minx2);
        uint32_t up = __shfl_up_sync(FULL, L[NR - 1], 1, LPP);
        uint32_t dn = __shfl_down_sync(FULL, L[0], 1, LPP);
        if (sub == 0) up = 0x00FF00FFu;           // Lp[-1] = 255
        if (sub == LPP - 1) dn = 0x00FF00FFu;     // Lp[DPL*LPP] = 255
        if (stores) store_plane<NR>(planeLane + (size_t)posPrev * P.Dp, L);   // emit the previous visit before overwriting L
        // (always with the uint8 truncation: a second copy of the step without it, selected per visit, measured 9 % slower)
        dp_step<NR, true, STR>(L, C, padm, up, dn, p1x2, p2x2, negmin);
        minx2 = group_min_x2<LPP>(lane_min_x2<NR>(L));
        posPrev = p;
    };

    // Loads run two visits ahead of the dependent chain (three input buffers, loop unrolled by three): with ~4
    // warps per scheduler a visit takes several hundred cycles, so two visits cover an L2 miss.
    StepInput<NR, DT> in0, in1, in2;
    uint32_t q0 = pos, q1 = 0, q2 = 0;
    int t0 = tcol, t1 = 0, t2 = 0;
    load_step<NR, DT>(P, pos, sub, in0);
    if (1 < len) { advance(); q1 = pos; t1 = tcol; load_step<NR, DT>(P, pos, sub, in1); }
    if (2 < len) { advance(); q2 = pos; t2 = tcol; load_step<NR, DT>(P, pos, sub, in2); }
    // ---- first pixel: L = C (SemiGlobalMatching.c:266-275)
    cost(in0, t0);
#pragma unroll
    for (int r = 0; r < NR; ++r) L[r] = C[r] | padm[r];
    minx2 = group_min_x2<LPP>(lane_min_x2<NR>(L));
    gPrev = in0.g;
    posPrev = q0;

    int s = 1;      // next visit to process: its inputs are in in1, those of visit s+1 in in2
    // main loop: three visits per iteration with the buffers rotating statically (no register moves) and every
    // prefetch unconditional: the last one fetches visit s+4 <= len-1
    for (; s + 5 <= len; s += 3) {
        advance(); q0 = pos; t0 = tcol; load_step<NR, DT>(P, pos, sub, in0);
        visit(in1, q1, t1);
        advance(); q1 = pos; t1 = tcol; load_step<NR, DT>(P, pos, sub, in1);
        visit(in2, q2, t2);
        advance(); q2 = pos; t2 = tcol; load_step<NR, DT>(P, pos, sub, in2);
        visit(in0, q0, t0);
    }
    // tail: at most four visits, one at a time
    while (s < len) {
        const bool more = s + 2 < len;
        if (more) { advance(); q0 = pos; t0 = tcol; load_step<NR, DT>(P, pos, sub, in0); }
        visit(in1, q1, t1);
        ++s;
        in1 = in2; q1 = q2; t1 = t2;
        if (more) { in2 = in0; q2 = q0; t2 = t0; }
    }
    if (stores) store_plane<NR>(planeLane + (size_t)posPrev * P.Dp, L);
}

// ------------------------------------------------------------------------------------------------ irregular paths
// Generic walker (path_walker.h), one path per warp; results are added to the side buffer because an
// irregular path visits pixels that a regular path also writes.
template <int NR, typename DT, bool PAD>
__device__ __forceinline__ void aggregate_irregular(const AggParams& P, const WarpWork job, int lane)
{
    constexpr int DPL = 2 * NR;
    constexpr unsigned FULL = 0xffffffffu;
    const Dir dir = direction(job.dir);
    uint32_t padm[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int i0 = DPL * lane + 2 * r;
        padm[r] = PAD ? ((i0 >= P.D ? 0x000000FFu : 0u) | (i0 + 1 >= P.D ? 0x00FF0000u : 0u)) : 0u;
    }
    const int dbase = P.dmin + DPL * lane;
    const bool lane_stores = DPL * lane < P.Dp;

    PathWalker wk;
    wk.start(P.W, P.H, dir.dx, dir.dy, job.firstPath);
    const int len = wk.length();
    uint32_t L[NR];
    uint32_t minx2 = 0x00FF00FFu, gPrev = 0;
    bool first = true;

    // The walk does not depend on data and neither do the matching costs, so the path is processed in blocks of kBlock
    // visits: wait for the block's inputs, form the costs of all its visits, issue the loads of the NEXT block, then run
    // the kBlock dependent DP steps while those loads fly.  No load is issued between a load and its first use (ptxas
    // tracks them with one scoreboard), and a whole block of DP steps covers the memory latency: these four warps run alone
    // on their schedulers, and with one visit fetched per visit processed (round 1) they took 237 us at C2 - as long as
    // all regular paths together (ncu r2_f).
    constexpr int kBlock = 4;
    StepInput<NR, DT> ring[kBlock];
    int tcR[kBlock], eR[kBlock];
    bool inR[kBlock];
    auto fetch = [&](int slot) {
        tcR[slot] = wk.tcol; inR[slot] = wk.inside(); eR[slot] = -1;
        if (inR[slot]) { load_step<NR, DT>(P, (uint32_t)wk.pos, lane, ring[slot]); eR[slot] = __ldg(P.entryOf + wk.pos); }
    };
    int fetched = 0;                              // visits whose loads have been issued
#pragma unroll
    for (int k = 0; k < kBlock; ++k)
        if (fetched < len) { if (fetched) wk.advance(); fetch(k); ++fetched; }
    for (int s = 0; s < len; s += kBlock) {
        uint32_t Cb[kBlock][NR], gB[kBlock];
        int eB[kBlock];
        bool inB[kBlock];
#pragma unroll
        for (int k = 0; k < kBlock; ++k) {
            inB[k] = (s + k < len) && inR[k];     // the reference's out-of-bounds visit is skipped (warp-uniform)
            eB[k] = eR[k];
            gB[k] = ring[k].g;
            if (inB[k]) pack_cost<NR, true, DT, false>(ring[k], tcR[k] - dbase + 1, Cb[k]);
        }
#pragma unroll
        for (int k = 0; k < kBlock; ++k)
            if (fetched < len) { wk.advance(); fetch(k); ++fetched; }
#pragma unroll
        for (int k = 0; k < kBlock; ++k) {
            if (!inB[k]) continue;
            if (first) {
#pragma unroll
                for (int r = 0; r < NR; ++r) L[r] = Cb[k][r] | padm[r];
                first = false;
            } else {
                int dg = (int)gB[k] - (int)gPrev;
                dg = dg < 0 ? -dg : dg;
                uint32_t up = __shfl_up_sync(FULL, L[NR - 1], 1);
                uint32_t dn = __shfl_down_sync(FULL, L[0], 1);
                if (lane == 0) up = 0x00FF00FFu;
                if (lane == 31) dn = 0x00FF00FFu;
                dp_step<NR, true, false>(L, Cb[k], padm, up, dn, P.p1x2, P.p2x2[dg], __vneg2(minx2));
            }
            minx2 = group_min_x2<32>(lane_min_x2<NR>(L));
            gPrev = gB[k];
            if (eB[k] >= 0 && lane_stores) {
                uint32_t* dst = P.side + ((size_t)eB[k] * P.Dp + DPL * lane) / 2;
#pragma unroll
                for (int r = 0; r < NR; ++r) atomicAdd(dst + r, L[r] & ~padm[r]);
            }
        }
    }
}

// NRH/LPPH: layout of the horizontal directions (latency-critical: H paths of W steps);
// NRV/LPPV: layout of the vertical and diagonal directions; NRI: layout of irregular paths (32 lanes).
// PAD = false: the disparity range fills every lane of every layout exactly (D == LPPH*2*NRH == LPPV*2*NRV ==
// 64*NRI), so the masks that park unused disparity slots at 255 vanish (they cost ~50 of ~220 instructions per
// visit when the compiler rematerialises them in the loop).
#ifdef SGM_AGG_TRACE
// Profiling build only (scripts/micro/agg_trace.py): start / end time, SM and direction of every warp job.
constexpr int kAggTraceWarps = 65536;
__device__ unsigned long long g_aggTrace[kAggTraceWarps * 3];
struct AggTraceScope {
    int widx, dir; unsigned long long t0;
    __device__ AggTraceScope(int w, int d) : widx(w), dir(d) { asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0)); }
    __device__ ~AggTraceScope() {
        unsigned long long t1; unsigned smid;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        if ((threadIdx.x & 31) == 0 && widx < kAggTraceWarps) {
            g_aggTrace[3 * widx] = t0; g_aggTrace[3 * widx + 1] = t1; g_aggTrace[3 * widx + 2] = ((unsigned long long)dir << 32) | smid;
        }
    }
};
#endif

template <int NRH, int LPPH, int NRV, int LPPV, int NRI, typename DT, bool PAD>
__global__ void __launch_bounds__(kAggWarpsPerBlock * 32, 16 / kAggWarpsPerBlock)
sgm_aggregate_paths(const __grid_constant__ AggParams P)
{
    const int widx = blockIdx.x * kAggWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (widx >= P.nIrregularWarps + P.nRegularWarps) return;
    const WarpWork job = P.work[widx];
#ifdef SGM_AGG_TRACE
    AggTraceScope trace(widx, widx < P.nIrregularWarps ? 8 + job.dir : job.dir);
#endif
    constexpr bool STR = agg_paired_layout(NRH, NRV);     // register pairing of the regular paths = byte order of the planes
    if (widx < P.nIrregularWarps) aggregate_irregular<NRI, DT, PAD>(P, job, lane);
    else if (job.dir == 0)        aggregate_horizontal<NRH, LPPH, true, DT, PAD, STR>(P, job, lane);
    else if (job.dir == 1)        aggregate_horizontal<NRH, LPPH, false, DT, PAD, STR>(P, job, lane);
    else if (job.dir < 4)         aggregate_column_like<NRV, LPPV, false, DT, PAD, STR>(P, job, lane);
    else                          aggregate_column_like<NRV, LPPV, true, DT, PAD, STR>(P, job, lane);
}

}  // namespace sgmb
