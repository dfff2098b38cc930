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

constexpr int kSpeckleLaunches = 4;
constexpr int kMedianLaunches = 1;

__device__ __forceinline__ bool pp_valid(float d) { return d != __int_as_float(0x7f800000); }

// ------------------------------------------------------------------------------------------------ K4 speckles
// Connected components over the 8-neighbourhood; two valid neighbours are connected when |d_a - d_b| <= diff
// (SemiGlobalMatching.c:620-624); components smaller than min_speckle_area become invalid (:633-638).
//
// Union-find on pixel indices (links always point to a smaller index, so the forest is acyclic), organised so
// that a single huge component - the normal case for a good disparity map - costs almost nothing:
//   init   every pixel is linked to the first pixel of its horizontal run inside its 32-pixel warp segment
//          (one ballot), segments of one run are chained through their first pixel; so rows are merged without
//          a single atomic;
//   merge  only the unions that are not implied by others are executed: the vertical link below a pixel is
//          skipped when the same two rows are already linked one column to the left, a diagonal link when the
//          corresponding horizontal + vertical links exist; finds compress the path they walk;
//   count  one atomicAdd per warp segment (its length) instead of one per pixel;
//   apply  look up the size of the pixel's component.
__device__ __forceinline__ bool pp_edge(float a, float b, float diff)
{
    return pp_valid(a) && pp_valid(b) && fabsf(__fsub_rn(a, b)) <= diff;
}

// Root of x with path compression (every node on the walked path is re-linked to the root found; values
// written are ancestors read from the structure, so concurrent unions stay consistent).
__device__ __forceinline__ int uf_find(int* lab, int x)
{
    int r = x, p;
    while ((p = lab[r]) < r) r = p;
    int cur = x;
    while ((p = lab[cur]) > r) { lab[cur] = r; cur = p; }
    return r;
}

__device__ __forceinline__ void uf_union(int* lab, int a, int b)
{
    bool done = false;
    do {
        a = uf_find(lab, a);
        b = uf_find(lab, b);
        if (a < b)      { const int old = atomicMin(&lab[b], a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(&lab[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

// grid (ceil(W/32), H), block 32: one warp per 32-pixel row segment.
__global__ void __launch_bounds__(32) speckle_init(const float* __restrict__ disp, int* __restrict__ lab, int* __restrict__ size,
                                                   int W, int H, float diff)
{
    const int lane = threadIdx.x, x = blockIdx.x * 32 + lane, y = blockIdx.y;
    const bool in = x < W;
    const int p = y * W + x;
    const float d = in ? disp[p] : __int_as_float(0x7f800000);
    const float left = (in && x > 0) ? disp[p - 1] : __int_as_float(0x7f800000);
    const bool valid = in && pp_valid(d);
    const bool joined = valid && pp_edge(d, left, diff);             // connected to the pixel on its left
    const unsigned starts = __ballot_sync(0xffffffffu, valid && !joined);
    if (!in) return;
    size[p] = 0;
    if (!valid) { lab[p] = -1; return; }
    const unsigned below = starts & (0xffffffffu >> (31 - lane));    // run starts at lanes <= mine
    const int segBase = p - lane;
    if (below)          lab[p] = segBase + (31 - __clz(below));      // first pixel of the run inside this segment
    else if (lane == 0) lab[p] = p - 1;                              // the run began in an earlier segment: chain to its last pixel
    else                lab[p] = segBase;                            // ... through this segment's first pixel
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

// grid (ceil(W/32), H), block 32: flatten and add each warp segment's length to its root.
__global__ void __launch_bounds__(32) speckle_count(int* lab, int* size, int W, int H)
{
    const int lane = threadIdx.x, x = blockIdx.x * 32 + lane, y = blockIdx.y;
    const bool in = x < W;
    const int p = y * W + x;
    int root = -1;
    if (in && lab[p] >= 0) { root = uf_find(lab, p); lab[p] = root; }
    // lanes with the same root inside the warp add once
    const unsigned peers = __match_any_sync(0xffffffffu, root);
    if (root >= 0 && lane == __ffs(peers) - 1) atomicAdd(&size[root], __popc(peers));
}

__global__ void speckle_apply(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ lab,
                              const int* __restrict__ size, int n, int minArea)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float d = in[p];
    const int r = lab[p];                                            // flattened by speckle_count
    if (r >= 0 && size[r] < minArea) d = __int_as_float(0x7f800000);   // SemiGlobalMatching.c:633-638
    out[p] = d;
}

// Returns the number of kernels launched.
static int launch_speckle_filter(const float* in, float* out, int32_t* scratch /* [2N] */, int W, int H, float diff,
                                 int minArea, cudaStream_t st)
{
    const int n = W * H;
    int* lab = scratch;
    int* size = scratch + n;
    dim3 gseg((W + 31) / 32, H);
    speckle_init<<<gseg, 32, 0, st>>>(in, lab, size, W, H, diff);
    dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    speckle_merge<<<g, b, 0, st>>>(in, lab, W, H, diff);
    speckle_count<<<gseg, 32, 0, st>>>(lab, size, W, H);
    speckle_apply<<<(n + 255) / 256, 256, 0, st>>>(in, out, lab, size, n, minArea);
    return kSpeckleLaunches;
}

// ------------------------------------------------------------------------------------------------ K5 in-place median
// The reference filters in place in raster order, so out(i,j) is the median of
//     out(i-1,j-1) out(i-1,j) out(i-1,j+1)        <- already filtered
//     out(i,  j-1) in (i,  j) in (i,  j+1)
//     in (i+1,j-1) in (i+1,j) in (i+1,j+1)
// for interior pixels, while border pixels pass through.  (i,j) depends on (i,j-1) and (i-1,j+1), so all
// pixels with the same t = 2i + j are independent: one thread per row, thread of row i handles column
// t - 2i at step t.  Inside a warp (32 consecutive rows) the filtered value of the row above arrives by
// __shfl_up_sync exactly when it is produced; filtered values never travel through memory, `in` is only
// read and `out` only written.  Between warps (row 32g-1 -> row 32g) the producer's last lane publishes
// (value, tag) as one 64-bit word per column in a global exchange row; the consumer warp refills 32 columns
// at a time with one coalesced load, spinning until all 32 tags are current.  Producers never wait, blocks
// are dispatched in index order and every dependency points to a lower warp index, so this cannot deadlock.
//
// Median-of-9 on the critical path: the seven inputs known early are sorted (16 compare-exchanges, off the
// dependency chain) and only their middle three can still be the answer; the two late inputs lo <= hi then
// give median = med3(e3, max(e2, lo), min(e4, hi)).
constexpr int kMedianWarpsPerBlock = 8;

__device__ __forceinline__ void cswap(float& a, float& b)
{
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo; b = hi;
}

__device__ __forceinline__ float median9_late2(float e0, float e1, float e2, float e3, float e4, float e5, float e6,
                                               float l0, float l1)
{
    cswap(e0, e6); cswap(e2, e3); cswap(e4, e5);
    cswap(e0, e2); cswap(e1, e4); cswap(e3, e6);
    cswap(e0, e1); cswap(e2, e5); cswap(e3, e4);
    cswap(e1, e2); cswap(e4, e6);
    cswap(e2, e3); cswap(e4, e5);
    cswap(e1, e2); cswap(e3, e4); cswap(e5, e6);
    const float lo = fminf(l0, l1), hi = fmaxf(l0, l1);
    const float x = fmaxf(e2, lo), y = fminf(e4, hi);
    return fmaxf(fminf(e3, x), fminf(fmaxf(e3, x), y));
}

// Prefetch distance (in steps) of the unfiltered inputs: every lane walks its own row, so its loads are
// uncoalesced L2 hits (~300 cycles); they are issued kMedianPF steps before first use and ride in a small
// shift register so a miss never stalls the warp-synchronous loop.
constexpr int kMedianPF = 6;

__global__ void __launch_bounds__(kMedianWarpsPerBlock * 32)
median3_inplace_wavefront(const float* __restrict__ in, float* __restrict__ out, unsigned long long* xchg,
                          int W, int H, unsigned epoch)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int PF = kMedianPF;
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * kMedianWarpsPerBlock + (threadIdx.x >> 5);     // group of 32 rows
    const int i = 32 * g + lane;
    if (32 * g >= H) return;
    const bool rowOk = i < H;
    const bool hasBelow = i + 1 < H;
    const float* inRow = in + (size_t)(rowOk ? i : 0) * W;
    const float* inBelow = in + (size_t)(hasBelow ? i + 1 : 0) * W;
    float* outRow = out + (size_t)(rowOk ? i : 0) * W;
    const bool produces = (lane == 31) && (32 * (g + 1) < H);                 // someone consumes this row
    unsigned long long* myX = xchg + (size_t)g * W;
    const unsigned long long* aboveX = (g > 0) ? xchg + (size_t)(g - 1) * W : nullptr;
    const unsigned tagBase = epoch << 16;

    float a = 0.f, b = 0.f, c = 0.f, left = 0.f;   // out(i-1, j-1..j+1), out(i, j-1)
    float m[PF + 2];                               // m[k] = in(i,   j + k)
    float n[PF + 3];                               // n[k] = in(i+1, j - 1 + k)
#pragma unroll
    for (int k = 0; k < PF + 2; ++k) m[k] = 0.f;
#pragma unroll
    for (int k = 0; k < PF + 3; ++k) n[k] = 0.f;
    float batch = 0.f;           // lane k holds out(32g-1, batchBase + k)
    int batchBase = -(1 << 30);

    const int tEnd = 2 * 31 + W - 1;
    for (int t = -(PF + 2); t <= tEnd; ++t) {
        const int j = t - 2 * lane;
        float o = 0.f;
        if (rowOk && j >= 0 && j < W) {
            if (i == 0 || i == H - 1 || j == 0 || j == W - 1) o = m[0];
            else o = median9_late2(a, b, m[0], m[1], n[0], n[1], n[2], left, c);
            outRow[j] = o;
            if (produces)
                *reinterpret_cast<volatile unsigned long long*>(myX + j) =
                    ((unsigned long long)(tagBase | (unsigned)(j + 1)) << 32) | __float_as_uint(o);
            left = o;
        }
        // filtered value of the row above for column j + 2: the lane below us just produced it
        float up = __shfl_up_sync(FULL, o, 1);
        if (aboveX) {                        // warp-uniform: the first lane takes it from the exchange row
            const int need = t + 2;          // lane 0's column j + 2
            if (need >= 0 && need < W) {
                if (need >= batchBase + 32) {
                    batchBase = need;
                    const int col = batchBase + lane;
                    const unsigned want = tagBase | (unsigned)(col + 1);
                    unsigned long long v = 0;
                    bool ok;
                    do {
                        ok = true;
                        if (col < W) {
                            v = *reinterpret_cast<const volatile unsigned long long*>(aboveX + col);
                            ok = (unsigned)(v >> 32) == want;
                        }
                    } while (!__all_sync(FULL, ok));
                    batch = __uint_as_float((unsigned)v);
                }
                const float fromAbove = __shfl_sync(FULL, batch, need - batchBase);
                if (lane == 0) up = fromAbove;
            }
        }
        a = b; b = c; c = up;
#pragma unroll
        for (int k = 0; k < PF + 1; ++k) m[k] = m[k + 1];
#pragma unroll
        for (int k = 0; k < PF + 2; ++k) n[k] = n[k + 1];
        const int jn = j + PF + 2;           // column entering both windows
        if (rowOk && jn >= 0 && jn < W) {
            m[PF + 1] = __ldg(inRow + jn);
            n[PF + 2] = hasBelow ? __ldg(inBelow + jn) : 0.f;
        }
    }
}

// `epoch` must differ between consecutive launches on the same exchange buffer (never 0: the buffer is
// zero-initialised), so stale tags of the previous frame are never taken for current ones.
static int launch_median3_inplace(const float* in, float* out, unsigned long long* xchg, unsigned* epoch, int W, int H,
                                  cudaStream_t st)
{
    *epoch = (*epoch % 65535u) + 1u;
    const int groups = (H + 31) / 32;
    const int blocks = (groups + kMedianWarpsPerBlock - 1) / kMedianWarpsPerBlock;
    median3_inplace_wavefront<<<blocks, kMedianWarpsPerBlock * 32, 0, st>>>(in, out, xchg, W, H, *epoch);
    return kMedianLaunches;
}

}  // namespace sgmb
