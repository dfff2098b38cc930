// Microbenchmark: one wavefront warp of the in-place median fed by (A) a shared-memory ring that is already filled
// (the product's bulk-copy ring, copy cost excluded) and (B) plain coalesced global loads prefetched PD steps ahead
// into registers (no ring, no mbarriers).  Cycles per step, one warp per SM like the product; the global buffer is
// 10 MB and L2-resident.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float fold5(float A1, float A2, float A3, float A4, float A5, float a, float b, float left, float c)
{
    const float p = fminf(a, b), q = fmaxf(a, b);
    const float B1 = fminf(p, left), B3 = fmaxf(q, left), B2 = fmaxf(p, fminf(q, left));
    const float e4 = fminf(fminf(A4, fmaxf(A3, B1)), fminf(fmaxf(A2, B2), fmaxf(A1, B3)));
    const float e5 = fminf(fminf(A5, fmaxf(A4, B1)), fminf(fmaxf(A3, B2), fmaxf(A2, B3)));
    return fminf(fmaxf(c, e4), e5);
}
template <int PD>
__global__ void k_ldg(const float* __restrict__ prep, float* out, unsigned long long* x, long long* cycles, int n)
{
    const int lane = threadIdx.x;
    const float* src = prep + (size_t)blockIdx.x * n * 160 + lane;
    float A[PD][5];
#pragma unroll
    for (int d = 0; d < PD; ++d)
#pragma unroll
        for (int r = 0; r < 5; ++r) A[d][r] = __ldg(src + (size_t)d * 160 + r * 32);
    float a = 1.f, b = 2.f, c = 3.f, left = 4.f, batch = 5.f + lane;
    float* op = out + (size_t)(blockIdx.x * 32 + lane) * 4096;
    long long t0 = clock64();
    for (int it = 0; it + PD <= n; it += PD) {
#pragma unroll
        for (int e = 0; e < PD; ++e) {
            const float o = fold5(A[e][0], A[e][1], A[e][2], A[e][3], A[e][4], a, b, left, c);
            const int nxt = min(it + e + PD, n - 1);
#pragma unroll
            for (int r = 0; r < 5; ++r) A[e][r] = __ldg(src + (size_t)nxt * 160 + r * 32);
            op[(it + e) & 4095] = o;
            if (lane == 31) asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(x + ((it + e) & 4095)), "r"(__float_as_uint(o)), "r"(it + e) : "memory");
            left = o;
            float up = __shfl_up_sync(0xffffffffu, o, 1);
            const float fa = __shfl_sync(0xffffffffu, batch, e & 31);
            if (lane == 0) up = fa;
            a = b; b = c; c = up;
        }
    }
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    out[lane] += a + b + c;
}
__global__ void k_lds(const float* __restrict__ prep, float* out, unsigned long long* x, long long* cycles, int n)
{
    __shared__ float ring[32][5][32];
    const int lane = threadIdx.x;
    for (int i = lane; i < 32 * 160; i += 32) (&ring[0][0][0])[i] = prep[i];
    __syncwarp();
    float a = 1.f, b = 2.f, c = 3.f, left = 4.f, batch = 5.f + lane;
    float* op = out + (size_t)(blockIdx.x * 32 + lane) * 4096;
    long long t0 = clock64();
    for (int it = 0; it < n; it += 32) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            const float* A = &ring[e][0][lane];
            const float o = fold5(A[0], A[32], A[64], A[96], A[128], a, b, left, c);
            op[(it + e) & 4095] = o;
            if (lane == 31) asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(x + ((it + e) & 4095)), "r"(__float_as_uint(o)), "r"(it + e) : "memory");
            left = o;
            float up = __shfl_up_sync(0xffffffffu, o, 1);
            const float fa = __shfl_sync(0xffffffffu, batch, e);
            if (lane == 0) up = fa;
            a = b; b = c; c = up;
        }
    }
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    out[lane] += a + b + c;
}
int main()
{
    const int groups = 12, n = 1344;
    float *prep, *out; unsigned long long* x; long long* cyc;
    cudaMalloc(&prep, (size_t)groups * n * 160 * 4); cudaMemset(prep, 0, (size_t)groups * n * 160 * 4);
    cudaMalloc(&out, (size_t)groups * 32 * 4096 * 4 + 1024); cudaMalloc(&x, 4096 * 8); cudaMallocManaged(&cyc, 64 * 8);
    auto report = [&](const char* name) {
        cudaDeviceSynchronize();
        long long mx = 0; for (int g = 0; g < groups; ++g) mx = cyc[g] > mx ? cyc[g] : mx;
        printf("%-28s %.1f cycles/step  (%s)\n", name, (double)mx / n, cudaGetErrorString(cudaGetLastError()));
    };
    for (int rep = 0; rep < 2; ++rep) {
        k_lds<<<groups, 32>>>(prep, out, x, cyc, n); report("smem ring (pre-filled)");
        k_ldg<8><<<groups, 32>>>(prep, out, x, cyc, n); report("global loads, 8 steps ahead");
        k_ldg<12><<<groups, 32>>>(prep, out, x, cyc, n); report("global loads, 12 steps ahead");
        k_ldg<16><<<groups, 32>>>(prep, out, x, cyc, n); report("global loads, 16 steps ahead");
    }
    return 0;
}
