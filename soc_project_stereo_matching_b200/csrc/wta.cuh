// wta.cuh -- K3: sum the per-direction planes into S (uint16) and, in the same pass over the data,
// select disparities for BOTH views, refine them to sub-pixel precision and run the left-right check.
//
// Restates ComputeDisparity() (SemiGlobalMatching.c:374-443, both `inverse` modes) and LRCheck()
// (SemiGlobalMatching.c:445-470).  S(p,d) = sum_r L_r(p,d) is exactly the reference's cost_aggr after
// CostAggregation() (SemiGlobalMatching.c:198-221,345); it only ever exists in registers and shared memory unless the
// taps are enabled, in which case it is also written out in the reference layout [N][D].
//
// The kernel streams 8 (or 4) byte planes once and is meant to run at HBM speed.  One CTA owns one image row and walks it
// from right to left in tiles of TW columns; CPP lanes share a pixel, each owning 16 consecutive disparities (one 128-bit
// load per plane):
//   phase A : add the planes as packed 16-bit fields (+ the side buffer of irregular paths); every plane register is
//             re-loaded for the NEXT tile as soon as it has been consumed, so 128 bytes per thread are always in flight.
//             The LEFT view is scanned right here, from the registers that hold the sums: keys (S << 4 | index) make
//             "lowest d wins ties" and "second = min over d != best" (:390-393,412-419) packed 16-bit min / max
//             (S <= 12 * 255 < 4096), three DPX instructions per two disparities; the CPP partial results are combined with
//             xor shuffles.  The sums also go to a shared-memory ring, SKEWED so that the right view becomes a row scan
//             as well: word i of the lane's chunk (disparities 2i, 2i+1 of the chunk) of pixel c goes to ring row c - 2i.
//   phase B : (after one __syncthreads) the right-view pixel x reads S[x + d][d] (:397-399); with the skew the even
//             disparities of chunk v lie in ring row x + dmin + 16v and the odd ones in the next row, eight words each -
//             two 128-bit loads per row and one PRMT per word rebuild the same packed layout the left view scans, and
//             the same scan code runs on it.  Walking right to left, all columns a right pixel needs have been summed
//             when its first column's tile is; every tile finishes TW left and TW right pixels.
//   epilogue: every thread finishes whole pixels from the parked 16-byte records (uniqueness, border, sub-pixel fit),
//             then LRCheck runs over the row.
// Cells that do not exist (disparities beyond D in the last chunk, right-view columns beyond the image) hold the marker
// 0x0FFF, larger than any sum, so they lose every comparison without a test.  The ring has a multiple of TW rows and at
// least 2*TW + D, so that phase A of the next tile never touches a row phase B still reads (one barrier per tile) and a
// tile's rows wrap around the ring's end only in one tile out of RB / TW (a separate code path).
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
    int ringCols;               // rows of the shared-memory ring: WtaShape<CPP>::ring_rows(D)
    // work split (host: wta_plan): blocks [0, fullRows) own one whole row each; block fullRows + p owns piece p = up to two
    // segments of the remaining rows, so that the last partial wave of rows is spread evenly over all SMs
    int fullRows;
    const int4* segments;       // [2 * pieces] {row, colBegin (multiple of TW), colEnd, segments of that row}; row < 0: none
    int* rowDone;               // [H] arrival counters of the split rows (zero between launches: the last arrival resets)
    // Which block takes what is decided when the block starts, from the SM it landed on: the first rowsPerSm blocks of an SM
    // take whole rows, the next one a piece (falling back to the other kind when one has run out), so that every SM ends up
    // with the same amount of work wherever the hardware places the blocks.  Zero between launches (the last block resets).
    int rowsPerSm, pieces;
    int* sched;                 // [kWtaSchedInts]: blocks seen per SM [kWtaSchedSms], row ticket, piece ticket, blocks done
};

constexpr int kWtaSchedSms = 1024, kWtaSchedInts = kWtaSchedSms + 3;

__device__ __forceinline__ float sgm_invalid() { return __int_as_float(0x7f800000); }

struct WtaPair { uint32_t kmin, ksec; };   // smallest and second smallest key; key = cost << 16 | disparity index

constexpr uint32_t kWtaMissing = 0x0FFFu;  // ring value of a cell that does not exist; every real sum is smaller (<= 12 * 255)

// Smallest and second smallest of the 16 cells of one chunk: w[i] holds the sums of disparity indices d0 + 2i (low half)
// and d0 + 2i + 1 (high half).  Packed 16-bit keys (sum << 4 | position in the chunk) keep both halves of a register busy;
// the result is widened to (sum << 16 | disparity index) keys; a cell that does not exist keeps the marker as its sum, which
// still loses against every real cell (wta_finish turns it into the reference's UINT16_MAX).
__device__ __forceinline__ WtaPair wta_scan16(const uint32_t (&w)[8], int d0)
{
    uint32_t m = w[0] * 16u + 0x00010000u, s = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        const uint32_t k = w[i] * 16u + (uint32_t)(2 * i) * 0x00010001u + 0x00010000u;   // fields <= 0x0FFF: no carry between them
        s = __vminu2(s, __vmaxu2(m, k));
        m = __vminu2(m, k);
    }
    const uint32_t a = m & 0xFFFFu, b = m >> 16;
    const uint32_t lo = min(a, b), hi = max(a, b);
    const uint32_t sec = min(hi, min(s & 0xFFFFu, s >> 16));
    // (sum << 4 | i) -> (sum << 16 | d0 + i) = k * 4096 + d0 - 4095 * i: one AND and two multiply-adds
    auto widen = [&](uint32_t k16) -> uint32_t { return (k16 * 4096u + (uint32_t)d0) - 4095u * (k16 & 15u); };
    return WtaPair{widen(lo), widen(sec)};
}

// combine the partial results of the CPP lanes that share a pixel (consecutive lane ids)
template <int CPP>
__device__ __forceinline__ void wta_reduce(WtaPair& w)
{
#pragma unroll
    for (int o = CPP / 2; o > 0; o >>= 1) {
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
    int second = (int)(rec.y >> 16);
    if (cmin >= (int)kWtaMissing) return sgm_invalid();   // no candidate at all (right view, x + dmin >= W)
    if (second >= (int)kWtaMissing) second = 0xFFFF;      // a single candidate: the reference's second minimum stays UINT16_MAX (:381)
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

// base + 16 * idx with a 32-bit index: one IMAD.WIDE (left to itself the compiler widens every term of the index to 64 bits
// first, four instructions per address)
__device__ __forceinline__ const uint4* wta_chunk_ptr(const uint4* base, uint32_t idx)
{
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(idx), "l"((unsigned long long)base));
    return reinterpret_cast<const uint4*>(a);
}

template <int CPP>
struct WtaShape {
    static constexpr int kThreads = CPP == 16 ? 512 : 256;
    static constexpr int kTW = kThreads / CPP;          // columns per tile
    static constexpr int kRW = 8 * CPP + 4;             // ring row stride in 32-bit words (a multiple of 4: rows are 16-byte aligned)
    // byte order of the planes: disparity ranges above 64 (CPP >= 8) run K2 with the paired register layout (aggregate.cuh,
    // agg_paired_layout; checked against the launch table in enqueue_frame)
    static constexpr bool kPairedPlanes = CPP >= 8;
    // ring rows: a multiple of the tile width, at least 2 * TW + D
    __host__ __device__ static constexpr int ring_rows(int D) { return 2 * kTW + ((D + kTW - 1) / kTW) * kTW; }
    __host__ __device__ static constexpr size_t ring_bytes(int D) { return (size_t)ring_rows(D) * kRW * 4; }
};

// One segment [colBegin, colEnd) of row y (the whole row when rowSegs == 1): left-view pixels of these columns, right-view
// pixels whose first candidate column lies in them.  A segment that does not end at the row's end first sums the D - 1 columns
// to its right into the ring ("halo" tiles: plane sum + store only), which is all the right view needs from them.
template <int CPP, int NP, bool TAPS>
__device__ __forceinline__ void wta_segment(const WtaParams& P, uint32_t* ring, const int y, const int colBegin, const int colEnd, const int rowSegs)
{
    constexpr int THREADS = WtaShape<CPP>::kThreads, TW = WtaShape<CPP>::kTW, RW = WtaShape<CPP>::kRW;
    const int W = P.W, D = P.D, Dp = P.Dp;
    const int RB = P.ringCols;                    // ring rows
    const size_t rowBase = (size_t)y * W;
    const int tid = threadIdx.x;
    float* rowL = P.dispOut + rowBase;
    float* rowR = P.rightRow + rowBase;
    uint4* recL = P.records + rowBase;
    uint4* recR = P.records + (size_t)P.H * W + rowBase;

    // thread (pix, v): pixel `pix` of the tile (phase A: left image column; phase B: right-view pixel), chunk v of 16 disparities
    const int pix = tid / CPP, v = tid % CPP;
    const bool chunkOk = 16 * v < Dp;             // CPP is the power of two >= Dp / 16
    // plane r, column c, this lane's chunk = rowChunks[c * dp16 + r * ps16] in 16-byte units: 32-bit indices (the host
    // refuses frames with 7 * planeStride / 16 >= 2^32), one multiply-add per load address
    const uint4* rowChunks = reinterpret_cast<const uint4*>(P.planes + rowBase * Dp) + v;
    const uint32_t dp16 = (uint32_t)Dp >> 4, ps16 = (uint32_t)(P.planeStride >> 4);

    // ---- ring: every cell "missing" until a sum is stored; right pixels without any candidate (x + dmin >= W)
    for (int i = tid; i < RB * RW; i += THREADS) ring[i] = kWtaMissing * 0x00010001u;
    if (P.checkLR && colEnd == W)
        for (int x = max(0, W - P.dmin) + tid; x < W; x += THREADS) recR[x] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);

    // cost of the cell (ring row rc of its column - may exceed RB by less than RB -, disparity index d in [0, D)):
    // row (rc - 2 * ((d >> 1) & 7)) mod RB, word d >> 1, half d & 1.  Branch-free: every lane of the warp calls it.
    auto ring_cost = [&](int rc, int d) -> uint32_t {
        int r = rc - 2 * ((d >> 1) & 7);
        r -= r >= RB ? RB : 0;
        r += r < 0 ? RB : 0;
        return __byte_perm(ring[r * RW + (d >> 1)], 0u, (d & 1) ? 0x4432u : 0x4410u);
    };
    // the same for a neighbour k of the best index that may lie outside [0, D): `below` / `above` are what the reference's
    // cost_local holds there (:433-435 read index -1 / D only for pixels that were already declared invalid)
    auto neighbour_cost = [&](int rc0, bool rowFollowsIndex, int k) -> uint32_t {
        const int kc = min(max(k, 0), D - 1);
        const uint32_t cst = ring_cost(rowFollowsIndex ? rc0 + kc : rc0, kc);
        return (unsigned)k >= (unsigned)D ? 0xFFFFu : cst;
    };

    uint4 cur[NP];
    int eCur = -1;
    const int haloEnd = min(W, colEnd + D - 1);   // columns [colEnd, haloEnd): summed for the right view only
    const int lastStart = ((haloEnd - 1) / TW) * TW;   // first tile processed = rightmost; tiles are aligned to TW
    {
        const int c = lastStart + pix;
        if (c < W && chunkOk) {
            const uint32_t src = (uint32_t)c * dp16;
#pragma unroll
            for (int r = 0; r < NP; ++r) cur[r] = __ldcs(wta_chunk_ptr(rowChunks, src + (uint32_t)r * ps16));
            if (P.hasSide) eCur = __ldg(P.entryOf + rowBase + c);
        }
    }
    __syncthreads();

    int s0 = lastStart % RB;                      // ring row of column c0 (a multiple of TW)
    uint32_t nsrc = (uint32_t)(lastStart + pix) * dp16;
    for (int c0 = lastStart; c0 >= colBegin; c0 -= TW) {
        const int cols = min(TW, W - c0);
        const bool halo = c0 >= colEnd;           // block-uniform
        // ------------------------------------------------------------------ phase A: plane sum, left view, skewed store
        const int c = c0 + pix;
        const bool active = c < W && chunkOk;     // this thread sums a chunk of column c
        const bool nextOk = c0 > colBegin && chunkOk;   // ... and one of column c - TW in the next tile (always inside the row)
        const int rTop = s0 + pix;                // ring row of column c (< RB: s0 is a multiple of TW)
        WtaPair wl{0xFFFFFFFFu, 0xFFFFFFFFu};
        {
            nsrc -= (uint32_t)TW * dp16;                              // index of column c - TW (unused, and possibly wrapped, when !nextOk)
            // od: the odd disparities of a word as 16-bit fields (one PRMT per plane); all: the plain 32-bit sum of the words,
            // = even fields + 256 * odd fields (mod 2^32), so the even fields cost no extraction per plane at all
            uint32_t all[4] = {0, 0, 0, 0}, od[4] = {0, 0, 0, 0};
#pragma unroll
            for (int r = 0; r < NP; r += 2) {                        // two planes per step: one 3-input add per field pair
                const uint32_t a[4] = {cur[r].x, cur[r].y, cur[r].z, cur[r].w};
                const uint32_t b[4] = {cur[r + 1].x, cur[r + 1].y, cur[r + 1].z, cur[r + 1].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    all[i] = all[i] + a[i] + b[i];
                    od[i] = od[i] + __byte_perm(a[i], 0, 0x4341) + __byte_perm(b[i], 0, 0x4341);
                }
                // the registers are free now: their loads for the next tile fly during the rest of this tile (issued after
                // the adds so that the old values need no copies)
                if (nextOk) {
                    cur[r] = __ldcs(wta_chunk_ptr(rowChunks, nsrc + (uint32_t)r * ps16));
                    cur[r + 1] = __ldcs(wta_chunk_ptr(rowChunks, nsrc + (uint32_t)(r + 1) * ps16));
                }
            }
            const int eNow = eCur;
            if (P.hasSide) eCur = nextOk ? __ldg(P.entryOf + rowBase + c - TW) : -1;
            if (active) {
                uint32_t out[8];                                    // natural order: (d0,d1),(d2,d3),...
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t ev = all[i] - (od[i] << 8);       // both even fields <= 8 * 255: exact mod 2^32
                    if constexpr (WtaShape<CPP>::kPairedPlanes) {
                        // K2's paired register layout stores each unit of 8 disparities as bytes 0,4,1,5 | 2,6,3,7: the even
                        // byte fields of word i ARE the pair (4(i/2)... ) in natural order and the odd ones the pair four further
                        out[4 * (i >> 1) + (i & 1)] = ev;            // word 0: (d0,d1)  1: (d2,d3)  2: (d8,d9)  3: (d10,d11)
                        out[4 * (i >> 1) + (i & 1) + 2] = od[i];     //         (d4,d5)     (d6,d7)     (d12,d13)   (d14,d15)
                    } else {
                        out[2 * i] = __byte_perm(ev, od[i], 0x5410);
                        out[2 * i + 1] = __byte_perm(ev, od[i], 0x7632);
                    }
                }
                if (eNow >= 0) {
                    const uint4* sp = reinterpret_cast<const uint4*>(P.side + (size_t)eNow * Dp + 16 * v);
                    const uint4 a = __ldg(sp), b = __ldg(sp + 1);
                    out[0] += a.x; out[1] += a.y; out[2] += a.z; out[3] += a.w;
                    out[4] += b.x; out[5] += b.y; out[6] += b.z; out[7] += b.w;
                }
                if (TAPS && P.S && !halo) {
                    uint16_t* g = P.S + (rowBase + c) * D + 16 * v;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int d = 16 * v + 2 * i;
                        if (d < D) g[2 * i] = (uint16_t)(out[i] & 0xFFFFu);
                        if (d + 1 < D) g[2 * i + 1] = (uint16_t)(out[i] >> 16);
                    }
                }
                if (16 * v + 16 > D) {                              // disparity indices beyond the range do not exist
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int d = 16 * v + 2 * i;
                        if (d >= D) out[i] = (out[i] & 0xFFFF0000u) | kWtaMissing;
                        if (d + 1 >= D) out[i] = (out[i] & 0x0000FFFFu) | (kWtaMissing << 16);
                    }
                }
                // skewed store: word i of the chunk -> ring row (c - 2i) mod RB, word 8v + i
                uint32_t* dst = ring + rTop * RW + 8 * v;
                if (s0 != 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) dst[i * (1 - 2 * RW)] = out[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) dst[i * (1 - 2 * RW) + (pix < 2 * i ? RB * RW : 0)] = out[i];
                }
                if (!halo) wl = wta_scan16(out, 16 * v);            // left view of column c: this chunk, in registers
            }
        }
        if (halo) {                                                 // nobody reads the ring before the segment's first real tile
            s0 -= TW; if (s0 < 0) s0 += RB;
            continue;
        }
        wta_reduce<CPP>(wl);                                        // all lanes (idle ones carry "nothing")
        // Costs next to the best disparity (sub-pixel fit).  With four or more lanes per pixel and the right view enabled
        // both views fetch theirs side by side after phase B (lanes v = 0, 1: left view, v = 2, 3: right view); otherwise
        // lane v = 0 fetches S[best - 1] and lane v = 1 (or 0 again when the pixel has one lane) S[best + 1] right here.
        constexpr bool kMergedFetch = CPP >= 4;
        if (!(kMergedFetch && P.checkLR)) {
            __syncwarp();                                           // the pixel's words (written by lanes of this warp) are visible
            const int best = (int)(wl.kmin & 0xFFFFu);
            const uint32_t nb = neighbour_cost(rTop, false, best + (CPP > 1 && v == 1 ? 1 : -1));
            const uint32_t c2 = CPP > 1 ? __shfl_down_sync(0xffffffffu, nb, 1) : neighbour_cost(rTop, false, best + 1);
            if (active && v == 0) recL[c] = make_uint4(wl.kmin, wl.ksec, nb | (c2 << 16), 0u);
        }
        __syncthreads();
        // ------------------------------------------------------------------ phase B: right view of the pixels whose first column lies in this tile
        if (P.checkLR) {
            const int x = c - P.dmin;                               // c = x + dmin: left column of disparity index 0
            const bool scan = pix < cols && x >= 0 && chunkOk;
            WtaPair wr{0xFFFFFFFFu, 0xFFFFFFFFu};
            if (scan) {
                int r = rTop + 16 * v;                              // even disparities of chunk v: ring row of column c + 16v, odd ones: the next row
                if (r >= RB) r -= RB;
                const int r1 = r + 1 == RB ? 0 : r + 1;
                const uint4* pe = reinterpret_cast<const uint4*>(ring + r * RW + 8 * v);
                const uint4* po = reinterpret_cast<const uint4*>(ring + r1 * RW + 8 * v);
                const uint4 e0 = pe[0], e1 = pe[1], o0 = po[0], o1 = po[1];
                const uint32_t wv[8] = {__byte_perm(e0.x, o0.x, 0x7610), __byte_perm(e0.y, o0.y, 0x7610), __byte_perm(e0.z, o0.z, 0x7610),
                                        __byte_perm(e0.w, o0.w, 0x7610), __byte_perm(e1.x, o1.x, 0x7610), __byte_perm(e1.y, o1.y, 0x7610),
                                        __byte_perm(e1.z, o1.z, 0x7610), __byte_perm(e1.w, o1.w, 0x7610)};
                wr = wta_scan16(wv, 16 * v);
            }
            wta_reduce<CPP>(wr);
            // 65535 where the reference sees UINT16_MAX: index outside the range or column x + k beyond the image (:407);
            // a left-view cell with k in [0, D) is never the marker
            auto cost = [&](bool right, int k) -> uint32_t {
                const uint32_t cst = neighbour_cost(rTop, right, k);
                return cst == kWtaMissing ? 0xFFFFu : cst;
            };
            if (kMergedFetch) {
                const bool right = (v & 2) != 0;
                const uint32_t kmin = right ? wr.kmin : wl.kmin, ksec = right ? wr.ksec : wl.ksec;
                const uint32_t nb = cost(right, (int)(kmin & 0xFFFFu) + ((v & 1) ? 1 : -1));
                const uint32_t c2 = __shfl_down_sync(0xffffffffu, nb, 1);
                uint4* rec = (v == 0 && active) ? recL + c : ((v == 2 && scan) ? recR + x : nullptr);
                if (rec) *rec = make_uint4(kmin, ksec, nb | (c2 << 16), 0u);
            } else {
                const int best = (int)(wr.kmin & 0xFFFFu);
                const uint32_t nb = cost(true, best + (CPP > 1 && v == 1 ? 1 : -1));
                const uint32_t c2 = CPP > 1 ? __shfl_down_sync(0xffffffffu, nb, 1) : cost(true, best + 1);
                if (scan && v == 0) recR[x] = make_uint4(wr.kmin, wr.ksec, nb | (c2 << 16), 0u);
            }
        }
        s0 -= TW; if (s0 < 0) s0 += RB;
    }
    // ---------------------------------------------------------------------- finish whole pixels from the records
    __syncthreads();
    for (int x = colBegin + tid; x < colEnd; x += THREADS) {
        const float d = wta_finish(recL[x], D, P.dmin, P.checkUnique, P.oneMinusRatio);
        rowL[x] = d;
        if (TAPS && P.dispLeftWta) P.dispLeftWta[rowBase + x] = d;
    }
    if (!P.checkLR) return;
    {   // right-view pixels whose first candidate column x + dmin lies in the segment (+ those without any, at the row's end)
        const int xEnd = colEnd == W ? W : colEnd - P.dmin;
        for (int x = max(0, colBegin - P.dmin) + tid; x < xEnd; x += THREADS) {
            const float r = wta_finish(recR[x], D, P.dmin, P.checkUnique, P.oneMinusRatio);
            rowR[x] = r;
            if (TAPS && P.dispRight) P.dispRight[rowBase + x] = r;
        }
    }
    // ---------------------------------------------------------------------- LR check (:445-470) of the whole row, by the block
    // that completes it (a split row: the last of its segments to arrive; nobody waits)
    if (rowSegs > 1) {
        __shared__ int lastArrival;
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int n = atomicAdd(P.rowDone + y, 1);
            lastArrival = n == rowSegs - 1;
            if (n == rowSegs - 1) P.rowDone[y] = 0;
        }
        __syncthreads();
        if (!lastArrival) return;
        __threadfence();
    } else {
        __syncthreads();
    }
    for (int x = tid; x < W; x += THREADS) {
        const float d = __ldcg(rowL + x);
        if (d != sgm_invalid()) {
            const float shifted = __fsub_rn((float)x, d);
            const int xr = __double2int_rz(__dadd_rn((double)shifted, 0.5));
            bool keep = true;
            if (xr < 0 || xr >= W) keep = false;
            else {
                const float dr = __ldcg(rowR + xr);
                if (dr != sgm_invalid() && fabsf(__fsub_rn(d, dr)) > P.lrThres) keep = false;
            }
            if (!keep) rowL[x] = sgm_invalid();
        }
    }
}

template <int CPP, int NP, bool TAPS>
__global__ void __launch_bounds__(WtaShape<CPP>::kThreads, CPP == 16 ? 1 : 3)
sgm_reduce_wta_lr(const __grid_constant__ WtaParams P)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw);
    __shared__ int item;                           // >= 0: whole row; < 0: piece -1 - item
    if (P.pieces == 0) {
        wta_segment<CPP, NP, TAPS>(P, ring, (int)blockIdx.x, 0, P.W, 1);
        return;
    }
    if (threadIdx.x == 0) {
        unsigned sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        const int seen = atomicAdd(P.sched + (sm & (kWtaSchedSms - 1)), 1);
        int* rowTicket = P.sched + kWtaSchedSms, *pieceTicket = rowTicket + 1;
        int it;
        if (seen % (P.rowsPerSm + 1) < P.rowsPerSm) {
            const int r = atomicAdd(rowTicket, 1);
            it = r < P.fullRows ? r : -1 - atomicAdd(pieceTicket, 1);
        } else {
            const int q = atomicAdd(pieceTicket, 1);
            it = q < P.pieces ? -1 - q : atomicAdd(rowTicket, 1);
        }
        item = it;
    }
    __syncthreads();
    const int it = item;
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
        int4 sg;
        if (it >= 0) {
            if (k) break;
            sg = make_int4(it, 0, P.W, 1);
        } else {
            sg = __ldg(P.segments + 2 * (-1 - it) + k);
            if (sg.x < 0) break;
        }
        if (k) __syncthreads();                   // the ring is re-initialised: everybody has left the previous segment
        wta_segment<CPP, NP, TAPS>(P, ring, sg.x, sg.y, sg.z, sg.w);
    }
    // the last block to finish clears the scheduling state for the next launch
    __syncthreads();
    if (threadIdx.x == 0) item = atomicAdd(P.sched + kWtaSchedSms + 2, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (item)
        for (int i = threadIdx.x; i < kWtaSchedInts; i += blockDim.x) P.sched[i] = 0;
}

}  // namespace sgmb
