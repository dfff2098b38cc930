// Microbenchmark: cycles per step of a single-warp dependent chain (shuffle + min/max), variants A..F.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void fold(const float* A, float a, float b, float left, float& e4, float& e5)
{
    const float A1 = A[0], A2 = A[32], A3 = A[64], A4 = A[96], A5 = A[128];
    const float p = fminf(a, b), q = fmaxf(a, b);
    const float B1 = fminf(p, left), B3 = fmaxf(q, left), B2 = fmaxf(p, fminf(q, left));
    e4 = fminf(fminf(A4, fmaxf(A3, B1)), fminf(fmaxf(A2, B2), fmaxf(A1, B3)));
    e5 = fminf(fminf(A5, fmaxf(A4, B1)), fminf(fmaxf(A3, B2), fmaxf(A2, B3)));
}
template <int V>
__global__ void k(float* out, unsigned long long* x, long long* cycles, int n, float seed)
{
    __shared__ float ring[32][5][32];
    const int lane = threadIdx.x;
    for (int i = lane; i < 32 * 5 * 32; i += 32) (&ring[0][0][0])[i] = seed + (i % 97);
    __syncwarp();
    float a = seed, b = seed + 1, c = seed + 2, left = seed + 3, batch = seed + lane;
    float* op = out + lane * 4096;
    long long t0 = clock64();
    for (int it = 0; it < n; it += 32) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            float e4 = a, e5 = b;
            if (V >= 2) fold(V >= 3 ? &ring[e][0][lane] : &ring[0][0][lane], a, b, left, e4, e5);
            const float o = fminf(fmaxf(c, e4), e5);
            if (V >= 4) op[(it + e) & 4095] = o;
            if (V >= 5 && lane == 31) asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(x + ((it + e) & 4095)), "r"(__float_as_uint(o)), "r"(it + e) : "memory");
            left = o;
            float up = __shfl_up_sync(0xffffffffu, o, 1);
            if (V >= 1) { const float fa = __shfl_sync(0xffffffffu, batch, e); if (lane == 0) up = fa; }
            a = b; b = c; c = up;
        }
    }
    long long t1 = clock64();
    if (lane == 0) cycles[V] = t1 - t0;
    out[lane] = a + b + c;
}
int main()
{
    float* out; unsigned long long* x; long long* cyc;
    cudaMalloc(&out, 32 * 4096 * 4 + 1024); cudaMalloc(&x, 4096 * 8); cudaMallocManaged(&cyc, 64);
    const int n = 32 * 400;
    for (int rep = 0; rep < 2; ++rep) {
        k<0><<<1, 32>>>(out, x, cyc, n, 1.f); k<1><<<1, 32>>>(out, x, cyc, n, 1.f); k<2><<<1, 32>>>(out, x, cyc, n, 1.f);
        k<3><<<1, 32>>>(out, x, cyc, n, 1.f); k<4><<<1, 32>>>(out, x, cyc, n, 1.f); k<5><<<1, 32>>>(out, x, cyc, n, 1.f);
        cudaDeviceSynchronize();
    }
    const char* names[] = {"A shfl+clamp", "B +2nd shfl/sel", "C +fold(regs)", "D +5 LDS", "E +STG", "F +st.relaxed.gpu lane31"};
    for (int v = 0; v < 6; ++v) printf("%-28s %.1f cycles/step\n", names[v], (double)cyc[v] / n);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
