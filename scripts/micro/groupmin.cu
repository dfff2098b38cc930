// Microbenchmark: latency of the per-group minimum inside a dependent chain (cycles per step, one warp).
//   V0: xor-shuffle butterfly over 16 lanes (4 rounds)     V1: two masked full-warp CREDUX.MIN (16-lane groups)
//   V2: butterfly over 8 lanes (3 rounds)                   V3: four masked CREDUX.MIN (8-lane groups)
//   V4: __reduce_min_sync over 32 lanes                     V5: butterfly over 32 lanes (5 rounds)
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__global__ void k(unsigned* out, long long* cycles, int n, unsigned seed)
{
    const int lane = threadIdx.x;
    unsigned m = seed + lane * 7919u;
    long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
        unsigned v = __vminu2(m, m * 3u + (unsigned)it);      // stand-in for the lane-local work
        if (V == 0) { for (int o = 8; o > 0; o >>= 1) v = __vminu2(v, __shfl_xor_sync(0xffffffffu, v, o)); }
        if (V == 1) {
            const bool hi = lane & 16;
            const unsigned a = __reduce_min_sync(0xffffffffu, hi ? 0xffffffffu : v), b = __reduce_min_sync(0xffffffffu, hi ? v : 0xffffffffu);
            v = hi ? b : a;
        }
        if (V == 2) { for (int o = 4; o > 0; o >>= 1) v = __vminu2(v, __shfl_xor_sync(0xffffffffu, v, o)); }
        if (V == 3) {
            const int g = lane >> 3;
            const unsigned a = __reduce_min_sync(0xffffffffu, g == 0 ? v : 0xffffffffu), b = __reduce_min_sync(0xffffffffu, g == 1 ? v : 0xffffffffu);
            const unsigned c = __reduce_min_sync(0xffffffffu, g == 2 ? v : 0xffffffffu), d = __reduce_min_sync(0xffffffffu, g == 3 ? v : 0xffffffffu);
            v = g == 0 ? a : (g == 1 ? b : (g == 2 ? c : d));
        }
        if (V == 4) v = __reduce_min_sync(0xffffffffu, v);
        if (V == 5) { for (int o = 16; o > 0; o >>= 1) v = __vminu2(v, __shfl_xor_sync(0xffffffffu, v, o)); }
        m = v + 1u;
    }
    long long t1 = clock64();
    if (lane == 0) cycles[V] = t1 - t0;
    out[lane] = m;
}
int main()
{
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 64);
    const int n = 20000;
    for (int rep = 0; rep < 2; ++rep) {
        k<0><<<1, 32>>>(out, cyc, n, 1); k<1><<<1, 32>>>(out, cyc, n, 1); k<2><<<1, 32>>>(out, cyc, n, 1);
        k<3><<<1, 32>>>(out, cyc, n, 1); k<4><<<1, 32>>>(out, cyc, n, 1); k<5><<<1, 32>>>(out, cyc, n, 1);
        cudaDeviceSynchronize();
    }
    const char* names[] = {"butterfly16", "2xCREDUX(16)", "butterfly8", "4xCREDUX(8)", "reduce32", "butterfly32"};
    for (int v = 0; v < 6; ++v) printf("%-14s %.1f cycles/step\n", names[v], (double)cyc[v] / n);
    return 0;
}
