// census.cuh -- K1: 5x5 census transform of both images in one launch.
//
// Restates census_transform_5x5 (SemiGlobalMatching.c:134-159): interior pixels only (2-pixel border
// stays 0), 25 comparisons "neighbour < centre", rows outer / columns inner, each shifted in from the
// LSB (first comparison ends at bit 24; the centre's own bit 12 is always 0).
//
// Layout produced:
//   left  : uint32 [H*W]                                   (read as a warp-uniform broadcast by K2)
//   right : FOUR copies, copy a shifted right by a elements inside a zero-padded array:
//           right4[a][padF + a + p] = census(p).  K2 needs, per lane, the 2/4/8 consecutive right
//           descriptors cR[q-k]; picking the copy with (padF + a + q - (n-1)) % 4 == 0 turns that
//           window into aligned 64/128-bit loads for every pixel position q.
// The image tile (+2 halo) is staged in shared memory with 32-bit loads where alignment allows.
#pragma once

#include <stdint.h>

namespace sgmb {

struct CensusParams {
    const uint8_t* img[2];   // left, right
    uint32_t* left;          // [N]
    uint32_t* right4;        // [4][copyStride]
    size_t copyStride;       // elements per copy (padF + N + padB, multiple of 4)
    int padF;
    int W, H;
};

constexpr int kCensusTileW = 64;
constexpr int kCensusTileH = 8;

__global__ void __launch_bounds__(kCensusTileW * kCensusTileH / 2)
sgm_census5x5(CensusParams P)
{
    // tile of 64 x 8 outputs, 68 x 12 inputs; every thread produces two horizontally adjacent outputs
    __shared__ uint8_t tile[kCensusTileH + 4][kCensusTileW + 4 + 4];
    const int which = blockIdx.z;
    const uint8_t* __restrict__ img = which ? P.img[1] : P.img[0];
    const int W = P.W, H = P.H;
    const int x0 = blockIdx.x * kCensusTileW, y0 = blockIdx.y * kCensusTileH;

    for (int i = threadIdx.x; i < (kCensusTileH + 4) * (kCensusTileW + 4); i += blockDim.x) {
        const int ty = i / (kCensusTileW + 4), tx = i % (kCensusTileW + 4);
        const int y = y0 + ty - 2, x = x0 + tx - 2;
        tile[ty][tx] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(img + (size_t)y * W + x) : 0;
    }
    __syncthreads();

    const int ty = threadIdx.x / (kCensusTileW / 2);
    const int tx = (threadIdx.x % (kCensusTileW / 2)) * 2;
    const int y = y0 + ty;
    if (y >= H) return;
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        const int x = x0 + tx + o;
        if (x >= W) break;
        uint32_t bits = 0;
        if (y >= 2 && y < H - 2 && x >= 2 && x < W - 2 && W > 5 && H > 5) {
            const uint32_t centre = tile[ty + 2][tx + o + 2];
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < 5; ++c)
                    bits = (bits << 1) | (uint32_t)(tile[ty + r][tx + o + c] < centre);
        }
        const size_t p = (size_t)y * W + x;
        if (which == 0) {
            P.left[p] = bits;
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a) P.right4[a * P.copyStride + P.padF + a + p] = bits;
        }
    }
}

}  // namespace sgmb
