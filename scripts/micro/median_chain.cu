// Microbenchmark: cycles per wavefront step of the in-place 3x3 median recurrence, one warp.
//   V0  current product formulation: fold {a, b, left} with the five sorted inputs (7 levels behind `left`), clamp by c
//   V1  new formulation: s3 <= s4 <= s5 of {A1..A5, a, b} prepared one step ahead (off the chain),
//       e4 = clamp(left, s3, s4), e5 = clamp(left, s4, s5), o = clamp(c, e4, e5): 4 levels behind `left`, 2 behind c
// Both include the 5 LDS, the output store, lane 31's publish store and the second shuffle feeding lane 0.
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
__device__ __forceinline__ void prep7(const float* A, float a, float b, float& s3, float& s4, float& s5)
{
    const float A1 = A[0], A2 = A[32], A3 = A[64], A4 = A[96], A5 = A[128];
    const float p = fminf(a, b), q = fmaxf(a, b);
    s3 = fminf(A3, fminf(fmaxf(A2, p), fmaxf(A1, q)));
    s4 = fminf(A4, fminf(fmaxf(A3, p), fmaxf(A2, q)));
    s5 = fminf(A5, fminf(fmaxf(A4, p), fmaxf(A3, q)));
}
template <int V>
__global__ void k(float* out, unsigned long long* x, long long* cycles, int n, float seed, float* check)
{
    __shared__ float ring[32][5][32];
    const int lane = threadIdx.x;
    for (int i = lane; i < 32 * 5 * 32; i += 32) {
        const int e = i / 160, r = (i / 32) % 5, l = i % 32;
        (&ring[0][0][0])[i] = seed + (float)((e * 37 + l * 11) % 97) + 10.f * r;     // sorted along r
    }
    __syncwarp();
    float a = seed, b = seed + 1, c = seed + 2, left = seed + 3, batch = seed + lane;
    float* op = out + lane * 4096;
    float s3 = 0, s4 = 0, s5 = 0, acc = 0;
    if (V == 1) prep7(&ring[0][0][lane], a, b, s3, s4, s5);
    long long t0 = clock64();
    for (int it = 0; it < n; it += 32) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            float o;
            if (V == 0) {
                float e4, e5;
                fold(&ring[e][0][lane], a, b, left, e4, e5);
                o = fminf(fmaxf(c, e4), e5);
            } else {
                const float e4 = fmaxf(s3, fminf(s4, left)), e5 = fmaxf(s4, fminf(s5, left));
                o = fminf(fmaxf(c, e4), e5);
                prep7(&ring[(e + 1) & 31][0][lane], b, c, s3, s4, s5);      // next step's a, b = this step's b, c
            }
            op[(it + e) & 4095] = o;
            if (lane == 31) asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(x + ((it + e) & 4095)), "r"(__float_as_uint(o)), "r"(it + e) : "memory");
            acc += o;
            left = o;
            float up = __shfl_up_sync(0xffffffffu, o, 1);
            const float fa = __shfl_sync(0xffffffffu, batch, e);
            if (lane == 0) up = fa;
            a = b; b = c; c = up;
        }
    }
    long long t1 = clock64();
    if (lane == 0) cycles[V] = t1 - t0;
    check[V * 32 + lane] = acc;
}
int main()
{
    float *out, *check; unsigned long long* x; long long* cyc;
    cudaMalloc(&out, 32 * 4096 * 4 + 1024); cudaMalloc(&x, 4096 * 8); cudaMallocManaged(&cyc, 64); cudaMallocManaged(&check, 64 * 4);
    const int n = 32 * 400;
    for (int rep = 0; rep < 2; ++rep) {
        k<0><<<1, 32>>>(out, x, cyc, n, 1.f, check); k<1><<<1, 32>>>(out, x, cyc, n, 1.f, check);
        cudaDeviceSynchronize();
    }
    printf("V0 fold+clamp          %.1f cycles/step\nV1 prepared s3..s5     %.1f cycles/step\n", (double)cyc[0] / n, (double)cyc[1] / n);
    int same = 1;
    for (int l = 0; l < 32; ++l) same &= (check[l] == check[32 + l]);
    printf("results identical: %s (%g vs %g)\n%s\n", same ? "yes" : "NO", check[5], check[37], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
