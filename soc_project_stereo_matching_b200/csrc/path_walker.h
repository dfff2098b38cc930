// path_walker.h -- pixel-index walker of one aggregation path, shared by host (path classification at
// SGMB_Configure) and device (the aggregation kernel follows it step by step), so there is exactly one
// statement of the path topology in the product.
//
// Semantics restated from the reference CostAggregate(): start positions SemiGlobalMatching.c:245-254,
// per-step moves :283-322, uint16_t row/col trackers :278-279 and :359-367.  The trackers are advanced
// after the wrap handling, so after a wrap the column tracker runs ahead of the true column; on the one
// "irregular" path per diagonal direction (more for portrait images) this makes the walk leave the
// toroidal diagonal, visit some pixels a second time and - for (1,1) and (-1,-1) - step outside the
// image once.  The walker reproduces the index arithmetic; callers skip visits
// whose index is outside [0, W*H).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define SGM_HD __host__ __device__ __forceinline__
#else
#define SGM_HD inline
#endif

namespace sgmb {

// Direction order of SemiGlobalMatching.c:213-220.
struct Dir { int dx, dy; };
SGM_HD Dir direction(int r)
{
    switch (r) {
        case 0: return Dir{1, 0};
        case 1: return Dir{-1, 0};
        case 2: return Dir{0, 1};
        case 3: return Dir{0, -1};
        case 4: return Dir{1, 1};
        case 5: return Dir{-1, -1};
        case 6: return Dir{1, -1};
        default: return Dir{-1, 1};
    }
}

SGM_HD bool dir_is_forward(int dx, int dy)          // SemiGlobalMatching.c:232
{
    return (dx == 1 && dy == 0) || (dx == 0 && dy == 1) || (dx == 1 && dy == 1) || (dx == -1 && dy == 1);
}

struct PathWalker {
    int W, H;
    int kind;                 // 0 horizontal, 1 vertical, 2 diagonal (dx == dy), 3 anti-diagonal (dx == -dy)
    int step;                 // +1 forward, -1 backward
    int pos;                  // pixel index of the current visit; may lie outside [0, W*H)
    int tcol;                 // true column of pos (meaningful only while pos is inside the image)
    unsigned short row, col;  // the reference's trackers

    SGM_HD void start(int w, int h, int dx, int dy, int path)
    {
        W = w; H = h;
        const bool fwd = dir_is_forward(dx, dy);
        step = fwd ? 1 : -1;
        if (dy == 0) {
            kind = 0;
            pos = fwd ? path * W : path * W + (W - 1);
            tcol = fwd ? 0 : W - 1;
        } else {
            kind = (dx == 0) ? 1 : ((dx == dy) ? 2 : 3);
            pos = fwd ? path : (H - 1) * W + path;
            tcol = path;
        }
        row = (unsigned short)(fwd ? 0 : H - 1);
        col = (unsigned short)path;
    }

    SGM_HD int length() const { return kind == 0 ? W : H; }
    SGM_HD bool inside() const { return pos >= 0 && pos < W * H; }

    SGM_HD void advance()
    {
        const bool fwd = step > 0;
        if (kind == 0) {
            pos += step; tcol += step;
        } else if (kind == 1) {
            pos += step * W;
        } else {
            if ((fwd && col == W - 1 && row < H - 1) || (!fwd && col == W - 1 && row > 0)) {
                pos = ((int)row + step) * W;            tcol = 0;      col = 0;
            } else if ((!fwd && col == 0 && row > 0) || (fwd && col == 0 && row < H - 1)) {
                pos = ((int)row + step) * W + (W - 1);  tcol = W - 1;  col = (unsigned short)(W - 1);
            } else if (kind == 2) {
                pos += step * (W + 1);  tcol += step;
            } else {
                pos += step * (W - 1);  tcol -= step;
            }
            if (tcol >= W) tcol -= W;
            if (tcol < 0)  tcol += W;
        }
        row = (unsigned short)(row + step);
        col = (unsigned short)(kind == 3 ? col - step : col + step);
    }
};

// Position a REGULAR diagonal path occupies at its s-th visit: the toroidal diagonal
// (row0 + step*s, (path + dx*s) mod W).  A path is regular iff its walk equals this for every s.
SGM_HD int regular_position(int W, int H, int dx, int dy, int path, int s)
{
    const int step = dir_is_forward(dx, dy) ? 1 : -1;
    const int r = (step > 0 ? 0 : H - 1) + step * s;
    int c = (path + dx * s) % W;
    if (c < 0) c += W;
    return r * W + c;
}

}  // namespace sgmb
