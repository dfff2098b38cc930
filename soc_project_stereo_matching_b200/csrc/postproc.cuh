// postproc.cuh -- the two post-processing stages SGM_Match runs after the hot path:
//   K4  RemoveSpeckles   (SemiGlobalMatching.c:585-642)  as GPU connected-component labelling
//   K5  MedianFilter     (SemiGlobalMatching.c:525-557, called IN PLACE at :120) as a skewed wavefront
// Both are bit-exact restatements: the speckle criterion is a symmetric edge relation, so components do
// not depend on the visiting order; the median selects one of its nine inputs, so any correct selection
// network returns the same bits (+inf takes part like any other value; there are no NaNs and no -0).
#pragma once

#include <math.h>
#include <stdint.h>

namespace sgmb {


__device__ __forceinline__ bool pp_valid(float d) { return d != __int_as_float(0x7f800000); }

// ------------------------------------------------------------------------------------------------ K4 speckles
// Connected components over the 8-neighbourhood; two valid neighbours are connected when |d_a - d_b| <= diff
// (SemiGlobalMatching.c:620-624); components smaller than min_speckle_area become invalid (:633-638).
//
// Union-find on pixel indices (init links every pixel to the first pixel of its run; unions link ROOTS only, the root with
// the larger scrambled index below the other one, so the forest is acyclic and shallow), organised so
// that a single huge component - the normal case for a good disparity map - costs almost nothing:
//   init   every pixel is linked to the first pixel of its horizontal run (one ballot per 32-pixel segment, the
//          run that crosses a segment boundary carried along the row by the warp that owns the row); so rows are
//          merged without a single atomic and every find starts one link away from a run's first pixel;
//   merge  only the unions that are not implied by others are executed: the vertical link below a pixel is
//          skipped when the same two rows are already linked one column to the left, a diagonal link when the
//          corresponding horizontal + vertical links exist; finds split the path they walk;
//   count  one atomicAdd per warp segment (its length) instead of one per pixel;
//   apply  look up the size of the pixel's component.
__device__ __forceinline__ bool pp_edge(float a, float b, float diff)
{
    return pp_valid(a) && pp_valid(b) && fabsf(__fsub_rn(a, b)) <= diff;
}

// Lock-free union-find in the standard form: only ROOTS are ever linked (compare-and-swap root -> smaller node), so a
// link that a finished union relies on is never replaced by anything but a link to one of its ancestors.
// Root of x with path splitting: every node on the walked path is re-pointed to its grandparent.  The stores are
// plain: the node written is a non-root (forever: roots are only ever linked, never unlinked) and the value is one of its ancestors, so
// concurrent writers can only disagree about WHICH ancestor.  Loads bypass L1 (another SM may have linked the node).
#ifdef SGM_SPECKLE_DEBUG
__device__ unsigned long long g_ufStats[8];       // finds, hops, longest walk, unions asked, CAS failures
#endif
__device__ __forceinline__ int uf_find(int* lab, int x)
{
    int p = __ldcg(lab + x);
#ifdef SGM_SPECKLE_DEBUG
    unsigned hops = 0;
#endif
    while (p != x) {
        const int gp = __ldcg(lab + p);
        if (gp != p) lab[x] = gp;
        x = p; p = gp;
#ifdef SGM_SPECKLE_DEBUG
        ++hops;
#endif
    }
#ifdef SGM_SPECKLE_DEBUG
    atomicAdd(&g_ufStats[0], 1ull); atomicAdd(&g_ufStats[1], (unsigned long long)hops); atomicMax(&g_ufStats[2], (unsigned long long)hops);
#endif
    return x;
}

// Root of x without touching the structure (used once no more unions happen: every thread then stores the ROOT of
// its own pixel, and concurrent readers see either the old parent or the root).
__device__ __forceinline__ int uf_root(const int* lab, int x)
{
    int p = __ldcg(lab + x);
    while (p != x) { x = p; p = __ldcg(lab + x); }
    return x;
}

__device__ __forceinline__ unsigned uf_key(int x) { return (unsigned)x * 0x9E3779B1u; }

__device__ __forceinline__ void uf_union(int* lab, int a, int b)
{
#ifdef SGM_SPECKLE_DEBUG
    atomicAdd(&g_ufStats[3], 1ull);
#endif
    for (;;) {
        a = uf_find(lab, a);
        b = uf_find(lab, b);
        if (a == b) return;
        // The root with the larger KEY is linked below the other one.  Any strict total order keeps the forest acyclic; with
        // the pixel index as the key every row's run linked to the run above it and the big component of a good map became a
        // chain hundreds of links deep that the first finds had to walk (longest walk measured at C2: 51 links with eight
        // paths, 102 with four - and speckle_merge took about 0.45 us per link of that walk).  A scrambled index (an odd
        // multiplier is a bijection on 32 bits) links in random order: expected depth O(log n).
        if (uf_key(a) < uf_key(b)) { const int t = a; a = b; b = t; }
        const int old = atomicCAS(lab + a, a, b);
#ifdef SGM_SPECKLE_DEBUG
        if (old != a) atomicAdd(&g_ufStats[4], 1ull);
#endif
        if (old == a) return;
        a = old;                                            // a had been linked meanwhile: continue from its parent
    }
}

// grid ceil(H / 4), block (32, 4): ONE WARP PER IMAGE ROW, walking its 32-pixel segments from left to right, so that every
// pixel is linked DIRECTLY to the first pixel of its horizontal run however many segments the run spans (the first pixel
// of the run that reaches the end of a segment is carried to the next one in a register).  Before, a segment was chained
// to the previous one through its first pixel, and a find from the far end of a run of k segments walked 2k links - each an
// L2 round trip; speckle_merge and speckle_count, which are nothing but finds, took 24 + 16 us at C2.  The loads of a
// batch of eight segments are issued together (they do not depend on the carried value).
constexpr int kSpeckleRowsPerBlock = 8;            // speckle_merge / speckle_count: rows per CTA
constexpr int kSpeckleInitRowsPerBlock = 4;
__global__ void __launch_bounds__(32 * kSpeckleInitRowsPerBlock) speckle_init(const float* __restrict__ disp, int* __restrict__ lab,
                                                                            int* __restrict__ size, int W, int H, float diff)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x, y = blockIdx.x * kSpeckleInitRowsPerBlock + threadIdx.y;
    if (y >= H) return;                                              // whole warps leave together
    const float inf = __int_as_float(0x7f800000);
    const int rowBase = y * W;
    int carry = -1;                                                  // first pixel of the run that reaches the end of the previous segment
    float prevLast = inf;                                            // disparity of the previous segment's last pixel
    constexpr int B = 8;                                             // segments per batch: their loads are issued together
    for (int xb = 0; xb < W; xb += 32 * B) {
        float dv[B];
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const int x = xb + 32 * k + lane;
            dv[k] = x < W ? __ldg(disp + rowBase + x) : inf;
        }
#pragma unroll
        for (int k = 0; k < B; ++k) {
            const int x0 = xb + 32 * k;
            if (x0 >= W) break;                                      // warp-uniform
            const int x = x0 + lane, p = rowBase + x;
            const bool in = x < W;
            const float d = dv[k];
            float left = __shfl_up_sync(FULL, d, 1);
            if (lane == 0) left = prevLast;
            const bool valid = in && pp_valid(d);
            const bool joined = valid && pp_edge(d, left, diff);     // connected to the pixel on its left
            const unsigned starts = __ballot_sync(FULL, valid && !joined);
            int label = -1;
            if (valid) {
                const unsigned below = starts & (0xffffffffu >> (31 - lane));    // run starts at lanes <= mine
                label = below ? rowBase + x0 + (31 - __clz(below)) : carry;      // no start up to here: the run came in from the left
            }
            if (in) { size[p] = 0; lab[p] = label; }
            carry = __shfl_sync(FULL, label, 31);
            prevLast = __shfl_sync(FULL, d, 31);
        }
    }
}

__global__ void speckle_merge(const float* __restrict__ disp, int* lab, int W, int H, float diff)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y + 1 >= H) return;
    const int p = y * W + x, q = p + W;
    const float inf = __int_as_float(0x7f800000);
    const float c = disp[p];
    if (!pp_valid(c)) return;
    const float l = x > 0 ? disp[p - 1] : inf, r = x + 1 < W ? disp[p + 1] : inf;
    const float bl = x > 0 ? disp[q - 1] : inf, b = disp[q], br = x + 1 < W ? disp[q + 1] : inf;
    const bool eL = pp_edge(c, l, diff), eR = pp_edge(c, r, diff);
    // vertical link, unless the path  c - l - bl - b  already exists
    if (pp_edge(c, b, diff) && !(eL && pp_edge(l, bl, diff) && pp_edge(bl, b, diff))) uf_union(lab, p, q);
    // diagonal links, unless  c - l - bl  /  c - r - br  exist (those vertical links are made by the neighbours)
    if (pp_edge(c, bl, diff) && !(eL && pp_edge(l, bl, diff)) && !(pp_edge(c, b, diff) && pp_edge(b, bl, diff))) uf_union(lab, p, q - 1);
    if (pp_edge(c, br, diff) && !(eR && pp_edge(r, br, diff)) && !(pp_edge(c, b, diff) && pp_edge(b, br, diff))) uf_union(lab, p, q + 1);
}

// grid (ceil(W/32), ceil(H/8)), block (32, 8): flatten and add each warp segment's length to its root.  Only the
// comparison size < minArea is ever made (SemiGlobalMatching.c:633), so a root whose count has already reached minArea
// receives no further additions: a good disparity map is one huge component, and ~15 000 atomics on its single counter
// were most of this kernel's time.  size[root] is therefore exact below minArea and "at least minArea" above.
__global__ void __launch_bounds__(32 * kSpeckleRowsPerBlock) speckle_count(int* lab, int* size, int W, int H, int minArea)
{
    const int lane = threadIdx.x, x = blockIdx.x * 32 + lane, y = blockIdx.y * kSpeckleRowsPerBlock + threadIdx.y;
    if (y >= H) return;
    const bool in = x < W;
    const int p = y * W + x;
    int root = -1;
    if (in && lab[p] >= 0) { root = uf_root(lab, p); lab[p] = root; }
    // lanes with the same root inside the warp add once
    const unsigned peers = __match_any_sync(0xffffffffu, root);
    if (root >= 0 && lane == __ffs(peers) - 1 && __ldcg(&size[root]) < minArea) atomicAdd(&size[root], __popc(peers));
}

// Returns the number of kernels launched.  The component sizes are applied by the consumer (K5a below, or
// speckle_apply when the median is switched off).
constexpr int kSpeckleLabelLaunches = 3;

// `mark(name)` is called after every launch (per-kernel timing hook of the host layer; a no-op otherwise).
template <typename Mark>
static int launch_speckle_labels(const float* in, int32_t* scratch /* [2N] */, int W, int H, float diff, int minArea, cudaStream_t st, Mark mark)
{
    const int n = W * H;
    int* lab = scratch;
    int* size = scratch + n;
    dim3 b(32, kSpeckleRowsPerBlock), g((W + 31) / 32, (H + kSpeckleRowsPerBlock - 1) / kSpeckleRowsPerBlock);
    speckle_init<<<(H + kSpeckleInitRowsPerBlock - 1) / kSpeckleInitRowsPerBlock, dim3(32, kSpeckleInitRowsPerBlock), 0, st>>>(in, lab, size, W, H, diff);
    mark("speckle_init");
    speckle_merge<<<g, b, 0, st>>>(in, lab, W, H, diff);
    mark("speckle_merge");
    speckle_count<<<g, b, 0, st>>>(lab, size, W, H, minArea);
    mark("speckle_count");
    return kSpeckleLabelLaunches;
}

// SemiGlobalMatching.c:633-638: a pixel of a component smaller than minArea becomes invalid.
__device__ __forceinline__ float speckle_filtered(const float* __restrict__ in, const int* __restrict__ lab,
                                                  const int* __restrict__ size, int p, int minArea)
{
    float d = __ldg(in + p);
    if (lab) {
        const int r = __ldg(lab + p);                                // flattened by speckle_count
        if (r >= 0 && __ldg(size + r) < minArea) d = __int_as_float(0x7f800000);
    }
    return d;
}

__global__ void speckle_apply(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ lab,
                              const int* __restrict__ size, int n, int minArea)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) out[p] = speckle_filtered(in, lab, size, p, minArea);
}

// ------------------------------------------------------------------------------------------------ K5 in-place median
// The reference filters in place in raster order (SemiGlobalMatching.c:120 calls MedianFilter with in == out),
// so for interior pixels out(i,j) is the median (5th smallest, :496-523) of
//     out(i-1,j-1) out(i-1,j) out(i-1,j+1)        <- already filtered        "B" = {a, b, c, left}
//     out(i,  j-1) in (i,  j) in (i,  j+1)
//     in (i+1,j-1) in (i+1,j) in (i+1,j+1)        <- five unfiltered inputs  "A"
// while border pixels pass through (:531-540).  (i,j) depends on (i,j-1) and (i-1,j+1), so all pixels with the
// same s = 2i + j are independent and the critical path is W + 2H dependent medians long.  Everything that does
// not depend on filtered values is therefore moved off that path:
//
//   K5a median_prepare (fully parallel, fused with the speckle filter's last step): per pixel, sort the five
//       unfiltered inputs and store them in the order the wavefront will read them: row group g = i / 32,
//       lane = i % 32, step u = j + 2*lane; per step 160 floats: 32 x float4 {A1..A4} then 32 x A5, so the wavefront
//       reads its five inputs with one 128-bit and one 32-bit shared-memory load (two instead of five: 222 -> 217 us).
//       A border pixel stores its own value five times: the 5th smallest of {v,v,v,v,v} + any four values is v,
//       so the wavefront needs no border test.
//   K5b median_wavefront: one warp per 32 rows, ONE WARP PER CTA (each warp has an SM sub-partition to itself),
//       thread of row i handles column s - 2*lane at step s.  With A sorted and (a,b) sorted one step earlier the
//       median is the 5th smallest of two sorted lists, min_i max(A[5-i], B[i]):  16 min/max (3-input FMNMX3 at
//       the end), five deep behind the arrival of c = out(i-1,j+1), which comes from the lane above by
//       __shfl_up_sync exactly one step after it was produced.  Between warps (row 32g-1 -> row 32g) the
//       producer's last lane publishes (tag, value) as one 64-bit word per column in a global exchange row; the
//       consumer fetches 32 columns per coalesced load, one batch ahead of use, and re-polls only if a tag is
//       stale.  Producers never wait and every dependency points to the previous ROW GROUP.  A CTA takes its row group
//       from a ticket counter (atomicAdd when it starts running), not from blockIdx: whatever order the hardware
//       dispatches CTAs in, the group a CTA waits for was taken by a CTA that is already running, so the chain of
//       waits always ends at a running CTA even when not all CTAs are co-resident (40 KB of shared memory per CTA
//       allow ~5 per SM, ~740 per device; H = 65535 has 2048 groups).  The poll loop is bounded: a CTA that has
//       waited ~2 s traps instead of hanging the device.
constexpr int kMedianLaunches = 2;
constexpr int kMedianTileW = 64;      // K5a: columns per block (32 for frames whose grid would not fill the device, see launch_median3_inplace)
// One bulk copy per 32 steps: the per-block bookkeeping (mbarrier wait, warp syncs, proxy fence, re-arming the copy) costs
// several hundred cycles of a lone warp, so the block is as long as the exchange batch allows (measured at C2: 4 steps per
// block 281 us, 8: 216 us, 16: 197 us, 32: 176 us).
constexpr int kMedianBlockSteps = 32; // K5b: steps per bulk-copy block (20 KB)
constexpr int kMedianBatch = 1;       // K5b: blocks per exchange batch / super-block (32 steps)
constexpr int kMedianRing = 2;        // K5b: blocks in the shared-memory ring (the next block's copy runs during the current block)
constexpr int kMedianFrontPad = 8;    // K5b starts kMedianFrontPad steps early (feeds a, b of the first interior column)
constexpr int kMedianStepBytes = 5 * 32 * 4;

__host__ __device__ __forceinline__ int median_steps_padded(int W)
{
    constexpr int Q = kMedianBlockSteps * kMedianBatch;
    return ((W + 62 + kMedianFrontPad + Q - 1) / Q) * Q;
}

static size_t median_prep_floats(int W, int H) { return (size_t)((H + 31) / 32) * median_steps_padded(W) * 5 * 32; }

__device__ __forceinline__ void cswap(float& a, float& b)
{
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo; b = hi;
}

// grid (ceil(W / 64), ceil(H / 32)), block 256.  `filtered` (optional tap) receives the speckle-filtered map.
template <int TW>
__global__ void __launch_bounds__(256)
median_prepare(const float* __restrict__ in, const int* __restrict__ lab, const int* __restrict__ size, int minArea,
               float* __restrict__ filtered, float* __restrict__ prep, int W, int H)
{
    constexpr int TS = TW + 3;                           // row stride 67 / 35 = 3 (mod 32): lane l reads column s - 2l -> bank (l + s) % 32
    __shared__ float tile[33][TS];                        // rows 32g .. 32g+32, columns j0-1 .. j0+TW
    const int g = blockIdx.y, j0 = blockIdx.x * TW, i0 = 32 * g;
    for (int k = threadIdx.x; k < 33 * (TW + 2); k += 256) {
        const int r = k / (TW + 2), c = k - r * (TW + 2);
        const int i = i0 + r, j = j0 - 1 + c;
        float v = 0.f;
        if (i < H && j >= 0 && j < W) {
            v = speckle_filtered(in, lab, size, i * W + j, minArea);
            if (filtered && r < 32 && c >= 1 && c <= TW) filtered[i * W + j] = v;
        }
        tile[r][c] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = i0 + lane;
    const int SP = median_steps_padded(W);
    float* dst = prep + ((size_t)g * SP * 5) * 32;       // per step: 32 x float4 {A1..A4} (one 128-bit read per lane), then 32 x A5
    // steps whose column s - 2*lane falls into this tile for some lane: s in [j0, j0 + TW + 62)
    for (int s = j0 + warp; s < j0 + TW + 62; s += 8) {
        const int j = s - 2 * lane;
        if (j < j0 || j >= j0 + TW || j >= W || i >= H) continue;
        const int c = j - j0 + 1;
        float a0 = tile[lane][c], a1, a2, a3, a4;
        if (i == 0 || i == H - 1 || j == 0 || j == W - 1) {
            a1 = a2 = a3 = a4 = a0;
        } else {
            a1 = tile[lane][c + 1]; a2 = tile[lane + 1][c - 1]; a3 = tile[lane + 1][c]; a4 = tile[lane + 1][c + 1];
            cswap(a0, a1); cswap(a3, a4); cswap(a2, a4); cswap(a2, a3); cswap(a0, a3);       // 9-comparator sort of five
            cswap(a0, a2); cswap(a1, a4); cswap(a1, a3); cswap(a1, a2);
        }
        float* q = dst + (size_t)(s + kMedianFrontPad) * 5 * 32;
        reinterpret_cast<float4*>(q)[lane] = make_float4(a0, a1, a2, a3);
        q[128 + lane] = a4;
    }
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned lo, unsigned hi)
{
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(lo), "r"(hi) : "memory");
}

// State carried from step to step by every lane.  Only c = out(i-1, j+1) arrives late (one shuffle after the lane
// above produced it), so everything else is folded BEFORE it arrives: e4 <= e5 are the 4th and 5th smallest of the
// eight values known early (the five sorted unfiltered inputs A and the sorted {a, b, left}); the median of all
// nine is then clamp(c, e4, e5) = min(max(c, e4), e5): two dependent operations behind the shuffle.
struct MedianLane {
    float a, b, c, left;
};

__device__ __forceinline__ void median_fold(const float* A, int lane4, float a, float b, float left, float& e4, float& e5)
{
    // A: this step's 160 floats; `lane4` = 4 * lane: {A1..A4} as one 128-bit shared-memory read, A5 as a second read
    const float4 A14 = *reinterpret_cast<const float4*>(A + lane4);
    const float A1 = A14.x, A2 = A14.y, A3 = A14.z, A4 = A14.w, A5 = A[128 + (lane4 >> 2)];
    const float p = fminf(a, b), q = fmaxf(a, b);
    const float B1 = fminf(p, left), B3 = fmaxf(q, left), B2 = fmaxf(p, fminf(q, left));     // sorted {a, b, left}
    // k-th smallest of two sorted lists = min over i + j = k of max(A_i, B_j)
    e4 = fminf(fminf(A4, fmaxf(A3, B1)), fminf(fmaxf(A2, B2), fmaxf(A1, B3)));
    e5 = fminf(fminf(A5, fmaxf(A4, B1)), fminf(fmaxf(A3, B2), fmaxf(A2, B3)));
}

// Values of the last row of the group above (row 32g - 1), kMedianFeed columns at a time.  The producer publishes every column as
// it is computed; the consumer takes them in pieces of kMedianFeed columns - lanes 0 .. kMedianFeed-1 hold the current
// piece, the next piece is requested when the current one is taken over and verified (tag) when its turn comes.  The
// piece size sets how far a group must trail its predecessor: 62 steps by construction (lane 31 of the group above is 62
// steps behind its lane 0) + one piece + one piece of prefetch distance + the store-to-poll latency.  With 32-column
// pieces fetched 32 steps ahead (round 1) a group trailed by ~170 steps = 7.6 us, and the whole filter was groups x that.
// Piece size, measured at C2 with the tile-staged output (median_wavefront, us): 4: 170, 8: 129, 16: 126, 32: 137; C3: 8: 493, 16: 452.
constexpr int kMedianFeed = 16;
struct AboveFeed {
    const unsigned long long* aboveX;
    int W, lane;
    unsigned tagBase;
    unsigned long long next;     // requested word of the next piece (lanes < kMedianFeed)
    float cur;                   // lanes < kMedianFeed: out(32g - 1, first column of the piece + lane)
#ifdef SGM_MEDIAN_DEBUG
    long long polls = 0, pieces = 0, waited = 0; unsigned long long tFirst = 0;
#endif

    __device__ __forceinline__ unsigned long long request(int col) const
    {
        return (lane < kMedianFeed && col >= 0 && col < W) ? ld_relaxed_gpu_u64(aboveX + col) : 0ull;
    }
    // make the piece starting at column col0 current (wait until the producer has published all of it), request the next one
    __device__ __forceinline__ void advance(int col0)
    {
        const int col = col0 + lane;
        const bool need = lane < kMedianFeed && col >= 0 && col < W;
        unsigned long long v = next;
        unsigned spins = 0;
#ifdef SGM_MEDIAN_DEBUG
        const long long w0 = clock64(); ++pieces;
#endif
        while (!__all_sync(0xffffffffu, !need || (unsigned)(v >> 32) == (tagBase | (unsigned)(col + 1)))) {
            __nanosleep(64);                             // do not hammer the L2 line the producer is storing to
            if (++spins > (1u << 24)) __trap();          // ~2 s without progress: fail the launch instead of hanging the device
            v = request(col);
#ifdef SGM_MEDIAN_DEBUG
            ++polls;
#endif
        }
#ifdef SGM_MEDIAN_DEBUG
        if (col0 >= 8) waited += clock64() - w0;
        if (col0 >= 0 && !tFirst) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tFirst));
#endif
        cur = __uint_as_float((unsigned)v);
        next = request(col + kMedianFeed);
    }
};

// 32 steps (kMedianBatch bulk-copy blocks); ringLane: the lane's view of the first of them.  PRED: some lane's column may fall outside the row.
// Results go to a 32 x 32 shared-memory tile (one conflict-free store per step); a second warp of the block writes the tile out
// as 32 coalesced row pieces of 128 bytes while the wavefront warp fills the other tile (named barriers, producer /
// consumer).  A step's 32 results lie in 32 different rows: stored directly they are 32 separate 4-byte transactions, which
// device memory absorbs but a page-locked HOST buffer does not (SGM_Match at C2 with the caller's buffer written directly:
// 1.08 ms with scattered stores against 0.65 ms with a device buffer + copy) - and writing the caller's buffer during the
// filter is what removes the device-to-host copy from the end of SGMB_Match.
constexpr int kMedianTileStride = 33;
template <bool HAS_ABOVE, bool PRED, typename Wait, typename Refill>
__device__ __forceinline__ void median_superblock(MedianLane& st, const float* ringLane, AboveFeed& feed, int colAbove0, float* tile, unsigned long long* xp,
                                                  int jBase, int Wrow, bool publishes, unsigned tag0, int lane, Wait wait, Refill refill)
{
    float* tileLane = tile + lane * kMedianTileStride;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int BS = kMedianBlockSteps, NB = kMedianBatch;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        wait(k);
#pragma unroll
        for (int e = 0; e < BS; ++e) {
            const int off = k * BS + e;
            if (HAS_ABOVE && off % kMedianFeed == 0) feed.advance(colAbove0 + off);
            float e4, e5;
            median_fold(ringLane + (k * BS + e) * 5 * 32, 4 * lane, st.a, st.b, st.left, e4, e5);
            const float o = fminf(fmaxf(st.c, e4), e5);
            tileLane[off] = o;
            if (publishes && (!PRED || (unsigned)(jBase + off) < (unsigned)Wrow)) st_relaxed_gpu_u64(xp + off, __float_as_uint(o), tag0 + (unsigned)off);
            st.left = o;
            float up = __shfl_up_sync(FULL, o, 1);       // out(i-1, j+2): the c of the next step
            if (HAS_ABOVE) {
                const float fromAbove = __shfl_sync(FULL, feed.cur, off % kMedianFeed);
                if (lane == 0) up = fromAbove;
            }
            st.a = st.b; st.b = st.c; st.c = up;
        }
        __syncwarp();                                     // every lane has consumed slot k
        refill(k);
    }
}

// named barriers of the wavefront block: kMedianBarFull + b: tile b is complete, kMedianBarFree + b: tile b has been written out
constexpr int kMedianBarFull = 1, kMedianBarFree = 3;
__device__ __forceinline__ void bar_sync64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive64(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// The second warp: writes tile after tile to the output rows.  Row r of the group produced columns s0 - 2r .. s0 - 2r + 31
// in the 32 steps that start at step s0.
__device__ __forceinline__ void median_flush_tiles(const int g, const float* tiles, float* __restrict__ out, int W, int H)
{
    constexpr int Q = kMedianBlockSteps * kMedianBatch;
    static_assert(Q == 32, "a tile holds the 32 steps of one super-block");
    const int lane = threadIdx.x & 31;
    const int nSuper = median_steps_padded(W) / Q;
    const int rows = min(32, H - 32 * g);
    float* outRows = out + (size_t)(32 * g) * W;
    for (int sb = 0; sb < nSuper; ++sb) {
        const int b = sb & 1;
        const int s0 = sb * Q - kMedianFrontPad;
        const float* tile = tiles + b * 32 * kMedianTileStride;
        bar_sync64(kMedianBarFull + b);
        float* dst = outRows + s0 + lane;
        if (s0 >= 62 && s0 + Q <= W) {
#pragma unroll 8
            for (int r = 0; r < rows; ++r) dst[(size_t)r * W - 2 * r] = tile[r * kMedianTileStride + lane];
        } else {
            for (int r = 0; r < rows; ++r)
                if ((unsigned)(s0 - 2 * r + lane) < (unsigned)W) dst[(size_t)r * W - 2 * r] = tile[r * kMedianTileStride + lane];
        }
        bar_arrive64(kMedianBarFree + b);
    }
}

template <bool HAS_ABOVE>
__device__ __forceinline__ void median_wavefront_body(const int g, const float* __restrict__ prep, float* __restrict__ out,
                                                      unsigned long long* xchg, int W, int H, unsigned epoch,
                                                      float (*ring)[kMedianBlockSteps][5][32], unsigned long long* mbar, float* tiles)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int BS = kMedianBlockSteps, NB = kMedianBatch, NR = kMedianRing;
    static_assert(NR % NB == 0 && NR >= NB, "ring must hold whole batches");
    constexpr unsigned kBlockBytes = BS * kMedianStepBytes;
    const int lane = threadIdx.x;
    const int Wrow = W;                                                       // rows beyond the image (last group) run like the others; their results stay in the tile
    (void)out;
    const bool publishes = (lane == 31) && (32 * (g + 1) < H);                // someone consumes this row
    unsigned long long* myX = xchg + (size_t)g * W;
    const unsigned long long* aboveX = HAS_ABOVE ? xchg + (size_t)(g - 1) * W : nullptr;
    const unsigned tagBase = epoch << 16;
    const int SP = median_steps_padded(W);
    const int nBlocks = SP / BS;
    const char* src = reinterpret_cast<const char*>(prep + ((size_t)g * SP * 5) * 32);

    auto issue = [&](int blk, int slot) {                 // lane 0: bulk copy of block blk into ring slot
        const unsigned bar = smem_u32(&mbar[slot]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kBlockBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(&ring[slot][0][0][0])), "l"(src + (size_t)blk * kBlockBytes), "r"(kBlockBytes), "r"(bar) : "memory");
    };

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[k])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k) if (k < nBlocks) issue(k, k);
    }
    __syncwarp();

#ifdef SGM_MEDIAN_DEBUG
    unsigned long long tStart, tEnd;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tStart));
#endif
    MedianLane st{0.f, 0.f, 0.f, 0.f};
    // exchange with the group above (AboveFeed): step s needs out(32g - 1, s + 2) for its successor
    AboveFeed feed{aboveX, W, lane, tagBase, 0ull, 0.f};
    if (HAS_ABOVE) feed.next = feed.request(2 - kMedianFrontPad + lane);
    const int nSuper = nBlocks / NB;
    for (int sb = 0; sb < nSuper; ++sb) {
        const int slot0 = (sb * NB) % NR;                // ring slots of this super-block: slot0 .. slot0 + NB - 1
        const float* ringLane = &ring[slot0][0][0][0];
        const int s0 = sb * NB * BS - kMedianFrontPad;   // first step of the super-block
        const int jBase = s0 - 2 * lane;
        unsigned long long* xp = myX + jBase;
        const unsigned tag0 = tagBase + (unsigned)(jBase + 1);      // == tagBase | (j + 1): j + 1 < 65536
        const unsigned parity = (unsigned)((sb * NB) / NR) & 1u;
        auto wait = [&](int k) {
            const unsigned bar = smem_u32(&mbar[slot0 + k]);
            unsigned done;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            } while (!__all_sync(FULL, done != 0));
        };
        auto refill = [&](int k) {
            const int nextBlk = sb * NB + k + NR;
            if (lane == 0 && nextBlk < nBlocks) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(nextBlk, slot0 + k);
            }
            __syncwarp();
        };
        // all 32 lanes are inside their rows for every step of the super-block <=> s0 >= 62 and s0 + 31 < W (and the row exists)
        const bool interior = s0 >= 62 && s0 + NB * BS <= W;
        const int tb = sb & 1;
        float* tile = tiles + tb * 32 * kMedianTileStride;
        if (sb >= 2) bar_sync64(kMedianBarFree + tb);     // the flush warp has written out what this tile held two super-blocks ago
        if (interior) median_superblock<HAS_ABOVE, false>(st, ringLane, feed, s0 + 2, tile, xp, jBase, Wrow, publishes, tag0, lane, wait, refill);
        else          median_superblock<HAS_ABOVE, true>(st, ringLane, feed, s0 + 2, tile, xp, jBase, Wrow, publishes, tag0, lane, wait, refill);
        bar_arrive64(kMedianBarFull + tb);
    }
#ifdef SGM_MEDIAN_DEBUG
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tEnd));
    if (lane == 0) printf("median g=%d start=%llu first=%llu end=%llu dur_us=%.1f pieces=%lld polls=%lld waited_cyc_after_first=%lld\n", g, tStart % 100000000ull,
                          feed.tFirst % 100000000ull, tEnd % 100000000ull, (tEnd - tStart) * 1e-3, feed.pieces, feed.polls, feed.waited);
#endif
}

// `ticket`: zeroed before every launch (it lives behind the exchange rows and is cleared with them).
__global__ void __launch_bounds__(64)
median_wavefront(const float* __restrict__ prep, float* __restrict__ out, unsigned long long* xchg, int* ticket,
                 int W, int H, unsigned epoch)
{
    __shared__ __align__(128) float ring[kMedianRing][kMedianBlockSteps][5][32];
    __shared__ __align__(8) unsigned long long mbar[kMedianRing];
    extern __shared__ float tiles[];                       // kMedianTileBytes: two output tiles (beyond the 48 KB of static shared memory)
    __shared__ int group;
    if (threadIdx.x == 0) group = atomicAdd(ticket, 1);
    __syncthreads();
    const int g = group;
    if (threadIdx.x >= 32) { median_flush_tiles(g, tiles, out, W, H); return; }
    if (g == 0) median_wavefront_body<false>(g, prep, out, xchg, W, H, epoch, ring, mbar, tiles);
    else        median_wavefront_body<true>(g, prep, out, xchg, W, H, epoch, ring, mbar, tiles);
}

constexpr int kMedianTileBytes = 2 * 32 * kMedianTileStride * (int)sizeof(float);
// once per device (SGMB_Configure): ring + tiles exceed the default shared-memory limit of a block
static cudaError_t median_configure() { return cudaFuncSetAttribute(median_wavefront, cudaFuncAttributeMaxDynamicSharedMemorySize, kMedianTileBytes); }

// argument list of median_wavefront, for the host code that re-points `out` in a recorded graph
constexpr int kMedianWavefrontArgs = 7, kMedianWavefrontOutArg = 1;

// Bytes of the exchange buffer: one 64-bit (tag, value) word per column and row group, then the ticket counter.
static size_t median_xchg_bytes(int W, int H) { return ((size_t)((H + 31) / 32) * W + 1) * sizeof(unsigned long long); }

// `epoch` must differ between consecutive launches on the same exchange buffer (never 0: the buffer is
// zero-initialised), so stale tags of the previous frame are never taken for current ones.
// in: disparity map before the speckle decision; lab/size: speckle labels (or NULL: no speckle filter).
template <typename Mark>
static int launch_median3_inplace(const float* in, const int* lab, const int* size, int minArea, float* filteredTap, float* prep,
                                  float* out, unsigned long long* xchg, unsigned* epoch, int W, int H, cudaStream_t st,
                                  Mark mark)
{
    *epoch = (*epoch % 65535u) + 1u;
    const int groups = (H + 31) / 32;
    // a block covers 32 rows x TW columns; with 64 columns a KITTI-sized frame has only 240 blocks for 148 SMs and the kernel
    // is one round of dependent loads deep: 32 columns there (14.2 -> 12.3 us at C2); large frames keep 64 (C3: 46 vs 55 us)
    if (((W + kMedianTileW - 1) / kMedianTileW) * groups < 4 * 148)
        median_prepare<32><<<dim3((W + 31) / 32, groups), 256, 0, st>>>(in, lab, size, minArea, filteredTap, prep, W, H);
    else
        median_prepare<kMedianTileW><<<dim3((W + kMedianTileW - 1) / kMedianTileW, groups), 256, 0, st>>>(in, lab, size, minArea, filteredTap, prep, W, H);
    mark("median_prepare");
    int* ticket = reinterpret_cast<int*>(xchg + (size_t)groups * W);
    median_wavefront<<<groups, 64, kMedianTileBytes, st>>>(prep, out, xchg, ticket, W, H, *epoch);
    mark("median_wavefront");
    return kMedianLaunches;
}

}  // namespace sgmb
