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
// Union-find with atomicMin links (labels only ever decrease, so the forest stays acyclic).
__device__ __forceinline__ int uf_find(const int* lab, int x)
{
    int p = lab[x];
    while (p != x) { x = p; p = lab[x]; }
    return x;
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

__global__ void speckle_init(const float* __restrict__ disp, int* __restrict__ lab, int* __restrict__ size, int n)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    lab[p] = pp_valid(disp[p]) ? p : -1;
    size[p] = 0;
}

// Two valid 8-neighbours are connected when |d_a - d_b| <= diff (SemiGlobalMatching.c:622-624).  Each pixel
// links to its E, SW, S and SE neighbours, which covers every unordered neighbour pair once.
__global__ void speckle_merge(const float* __restrict__ disp, int* lab, int W, int H, float diff)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int p = y * W + x;
    const float d = disp[p];
    if (!pp_valid(d)) return;
    const int ox[4] = {1, -1, 0, 1}, oy[4] = {0, 1, 1, 1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int qx = x + ox[k], qy = y + oy[k];
        if (qx < 0 || qx >= W || qy >= H) continue;
        const int q = qy * W + qx;
        const float e = disp[q];
        if (pp_valid(e) && fabsf(__fsub_rn(e, d)) <= diff) uf_union(lab, p, q);
    }
}

__global__ void speckle_count(int* lab, int* size, int n)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || lab[p] < 0) return;
    const int r = uf_find(lab, p);
    lab[p] = r;      // safe: r is an ancestor of p, so every chain through p still reaches the root
    atomicAdd(&size[r], 1);
}

__global__ void speckle_apply(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ lab,
                              const int* __restrict__ size, int n, int minArea)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float d = in[p];
    const int l = lab[p];
    if (l >= 0) {
        const int r = uf_find(lab, l);
        if (size[r] < minArea) d = __int_as_float(0x7f800000);     // SemiGlobalMatching.c:633-638
    }
    out[p] = d;
}

// Returns the number of kernels launched.
static int launch_speckle_filter(const float* in, float* out, int32_t* scratch /* [2N] */, int W, int H, float diff,
                                 int minArea, cudaStream_t st)
{
    const int n = W * H;
    int* lab = scratch;
    int* size = scratch + n;
    speckle_init<<<(n + 255) / 256, 256, 0, st>>>(in, lab, size, n);
    dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    speckle_merge<<<g, b, 0, st>>>(in, lab, W, H, diff);
    speckle_count<<<(n + 255) / 256, 256, 0, st>>>(lab, size, n);
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
