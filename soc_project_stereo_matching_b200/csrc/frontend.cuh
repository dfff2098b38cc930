// frontend.cuh -- the steps either side of the SGM path in the reference system.
//
// N4, evaluation step after the path (HostScript_Server/depth_image.py):
//   sgm_disparity_to_depth : depth = baseline * fx / (disparity + doffs)                  depth_image.py:138-165
//                            float32 throughout, in numpy's evaluation order: bf = fl(baseline * fx),
//                            den = fl(disp + doffs), depth = fl(bf / den).  Our invalid disparity (+inf,
//                            SemiGlobalMatching.h:12) becomes NaN, the "no value" the reference's evaluation
//                            skips (its ground truth marks invalid pixels NaN; compare_img keeps finite pairs only).
//   sgm_compare_depth_*    : valid = isfinite(test) & isfinite(gt); RMSE = sqrt(mean((test-gt)^2)) and bad-pixel
//                            rate = count(|test-gt| > thresh) / n_valid                    depth_image.py:276-319
//                            Differences and squares are formed in float32 like numpy does; the SUM runs in
//                            float64 in a fixed order (deterministic), so RMSE agrees with numpy's float32
//                            pairwise sum to ~1e-6 relative, counts are exact.
#pragma once

#include <math.h>
#include <stdint.h>

namespace sgmb {

__global__ void sgm_disparity_to_depth(const float* __restrict__ disp, float* __restrict__ depth, size_t n, float bf, float doffs)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float d = disp[i];
        const float r = __fdiv_rn(bf, __fadd_rn(d, doffs));
        // NaN results carry the default quiet-NaN pattern the host's arithmetic produces (the GPU's is 0x7fffffff)
        depth[i] = (d == __int_as_float(0x7f800000) || r != r) ? __int_as_float(0x7fc00000) : r;
    }
}

struct CompareAcc { double sumSq; unsigned long long nValid, nBad; };

constexpr int kCompareBlocks = 296;       // 2 per SM
constexpr int kCompareThreads = 256;

__global__ void __launch_bounds__(kCompareThreads)
sgm_compare_depth_partial(const float* __restrict__ gt, const float* __restrict__ test, size_t n, float thresh, CompareAcc* partial)
{
    double sumSq = 0.0;
    unsigned long long nValid = 0, nBad = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float g = gt[i], t = test[i];
        if (isfinite(g) && isfinite(t)) {
            const float diff = __fsub_rn(t, g);
            sumSq += (double)__fmul_rn(diff, diff);
            ++nValid;
            nBad += fabsf(diff) > thresh;
        }
    }
    __shared__ double sS[kCompareThreads];
    __shared__ unsigned long long sV[kCompareThreads], sB[kCompareThreads];
    sS[threadIdx.x] = sumSq; sV[threadIdx.x] = nValid; sB[threadIdx.x] = nBad;
    __syncthreads();
    for (int o = kCompareThreads / 2; o > 0; o >>= 1) {          // fixed tree: deterministic
        if ((int)threadIdx.x < o) {
            sS[threadIdx.x] += sS[threadIdx.x + o]; sV[threadIdx.x] += sV[threadIdx.x + o]; sB[threadIdx.x] += sB[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = CompareAcc{sS[0], sV[0], sB[0]};
}

__global__ void sgm_compare_depth_final(const CompareAcc* partial, int nPartial, CompareAcc* out)
{
    if (threadIdx.x || blockIdx.x) return;
    CompareAcc a{0.0, 0, 0};
    for (int i = 0; i < nPartial; ++i) { a.sumSq += partial[i].sumSq; a.nValid += partial[i].nValid; a.nBad += partial[i].nBad; }
    *out = a;
}

}  // namespace sgmb
