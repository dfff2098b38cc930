/*
 * sgm_oracle.c -- CPU restatement of the reference SGM pipeline.  TEST INFRASTRUCTURE ONLY
 * (see sgm_oracle.h for who may use it and for the parity status: PINNED against the compiled
 * reference in oracle/_ref and the fixtures in tests/golden/).
 *
 * "SGM.c:n" = /root/reference/SemiGlobalMatching/SemiGlobalMatching/SemiGlobalMatching.c line n.
 * The arithmetic (integer widths, truncations, float/double mix) follows those lines exactly; the
 * structure does not: sizes are run-time, there are no globals, each aggregation path is first
 * expanded into a list of pixel indices by sgmo_walk_path() and then swept by one DP routine.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -o libsgm_oracle.so sgm_oracle.c -lm
 */
#include "sgm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SGMO_INVALID ((float)INFINITY) /* SGM.h:12 */

/* ------------------------------------------------------------------------------------------------
 * Census 5x5 (SGM.c:134-159): for interior pixels compare the 25 window pixels (rows outer,
 * columns inner) with the centre, shifting one bit in from the LSB per comparison; "neighbour <
 * centre" -> 1.  The two-pixel border is never written; the whole stage is skipped for W<=5 or H<=5.
 * ---------------------------------------------------------------------------------------------- */
void sgmo_census5x5(const uint8_t* img, int W, int H, uint32_t* census)
{
    if (!img || !census || W <= 5 || H <= 5) return;
    for (int y = 2; y < H - 2; ++y) {
        for (int x = 2; x < W - 2; ++x) {
            const uint8_t centre = img[(size_t)y * W + x];
            uint32_t bits = 0;
            for (int k = 0; k < 25; ++k) {
                const int yy = y + k / 5 - 2, xx = x + k % 5 - 2;
                bits = (bits << 1) | (uint32_t)(img[(size_t)yy * W + xx] < centre);
            }
            census[(size_t)y * W + x] = bits;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Matching cost (SGM.c:161-196): C(y,x,d) = popcount(cl[y,x] ^ cr[y,x-d]) for x-d inside the row,
 * UINT8_MAX/2 = 127 otherwise.  Layout: pixel-major, disparity-minor (SGM.c:168).
 * ---------------------------------------------------------------------------------------------- */
void sgmo_cost(const uint32_t* cl, const uint32_t* cr, int W, int H, int dmin, int dmax, uint8_t* C)
{
    const int D = dmax - dmin;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            uint8_t* out = C + ((size_t)y * W + x) * D;
            const uint32_t a = cl[(size_t)y * W + x];
            for (int d = dmin; d < dmax; ++d) {
                const int xr = x - d;
                out[d - dmin] = (xr < 0 || xr >= W) ? (uint8_t)(UINT8_MAX / 2)
                                                   : (uint8_t)__builtin_popcount(a ^ cr[(size_t)y * W + xr]);
            }
        }
}

/* ------------------------------------------------------------------------------------------------
 * Generalised census / cost on 64-bit descriptors.  cw = ch = 5 restates SGM.c:134-159,161-196 (same
 * bits as sgmo_census5x5, zero-extended); cw, ch = 9, 7 is the EXTENSION (63 comparisons, centre bit 31
 * always 0, costs 0..62) -- PARITY UNPINNED, see sgm_oracle.h.
 * ---------------------------------------------------------------------------------------------- */
void sgmo_census(const uint8_t* img, int W, int H, int cw, int ch, uint64_t* census)
{
    if (!img || !census || cw < 1 || ch < 1 || cw * ch > 64 || W <= cw || H <= ch) return;
    const int rx = cw / 2, ry = ch / 2;
    for (int y = ry; y < H - ry; ++y)
        for (int x = rx; x < W - rx; ++x) {
            const uint8_t centre = img[(size_t)y * W + x];
            uint64_t bits = 0;
            for (int r = -ry; r <= ry; ++r)
                for (int c = -rx; c <= rx; ++c)
                    bits = (bits << 1) | (uint64_t)(img[(size_t)(y + r) * W + (x + c)] < centre);
            census[(size_t)y * W + x] = bits;
        }
}

void sgmo_cost64(const uint64_t* cl, const uint64_t* cr, int W, int H, int dmin, int dmax, uint8_t* C)
{
    const int D = dmax - dmin;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            uint8_t* out = C + ((size_t)y * W + x) * D;
            const uint64_t a = cl[(size_t)y * W + x];
            for (int d = dmin; d < dmax; ++d) {
                const int xr = x - d;
                out[d - dmin] = (xr < 0 || xr >= W) ? (uint8_t)(UINT8_MAX / 2)
                                                   : (uint8_t)__builtin_popcountll(a ^ cr[(size_t)y * W + xr]);
            }
        }
}

/* ------------------------------------------------------------------------------------------------
 * Path walker (SGM.c:232-256 start positions, :281-323 moves, :359-367 row/col trackers).
 * The reference keeps uint16_t row/col trackers that are advanced AFTER the wrap handling, so after
 * a wrap the column tracker runs one ahead of the true column; that drift is part of the semantics
 * and is reproduced here.  Horizontal directions have H paths of W pixels, all others W paths of H.
 * ---------------------------------------------------------------------------------------------- */
int sgmo_walk_path(int W, int H, int dx, int dy, int path, int64_t* pos)
{
    const int forward = (dx == 1 && dy == 0) || (dx == 0 && dy == 1) || (dx == 1 && dy == 1) || (dx == -1 && dy == 1);
    const int64_t step = forward ? 1 : -1;
    const int horizontal = (dy == 0);
    const int len = horizontal ? W : H;
    int64_t p;
    if (horizontal) p = forward ? (int64_t)path * W : (int64_t)path * W + (W - 1);                 /* :245-247 */
    else            p = forward ? (int64_t)path : (int64_t)(H - 1) * W + path;                     /* :252-254 */
    uint16_t row = (uint16_t)(forward ? 0 : H - 1), col = (uint16_t)path;                          /* :278-279 */
    const int anti = (dx == -1 && dy == 1) || (dx == 1 && dy == -1);
    int n = 0;
    pos[n++] = p;
    for (int k = 0; k < len - 1; ++k) {
        if (horizontal)      p += step;                                                             /* :283-288 */
        else if (dx == 0)    p += step * W;                                                         /* :289-294 */
        else if ((forward && col == W - 1 && row < H - 1) || (!forward && col == W - 1 && row > 0)) {
            p = ((int64_t)row + step) * W;             col = 0;                                     /* :297-303 */
        } else if ((!forward && col == 0 && row > 0) || (forward && col == 0 && row < H - 1)) {
            p = ((int64_t)row + step) * W + (W - 1);   col = (uint16_t)(W - 1);                     /* :304-310 */
        } else if (!anti)    p += step * (W + 1);                                                   /* :311-316 */
        else                 p += step * (W - 1);                                                   /* :317-322 */
        pos[n++] = p;
        row = (uint16_t)(row + step);                                                               /* :359 */
        col = (uint16_t)(anti ? col - step : col + step);                                           /* :360-367 */
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * One aggregation direction (SGM.c:229-372).
 *   first pixel of a path:  L = C, S += C, minPrev = min_d C                        (:266-275)
 *   every later pixel, with Lp[-1] = Lp[D] = 255                                    (:260-263,349,357)
 *      l1 = Lp[d]; l2 = Lp[d-1]+P1; l3 = Lp[d+1]+P1;                                (:332-334, uint16)
 *      l4 = minPrev + max(P1, P2_init / (|g - gPrev| + 1))                          (:335, uint16)
 *      L[d] = (uint8_t)(C + min(l1..l4) - minPrev)   -- wraps mod 256               (:343)
 *      S[d] += L[d]; minPrev = min_d L[d]; gPrev = g                                (:345-353,369)
 * The reference's out-of-bounds visit (index outside [0, W*H)) is skipped.
 * ---------------------------------------------------------------------------------------------- */
void sgmo_aggregate_dir(const uint8_t* img, const uint8_t* C, int W, int H, int D, int p1, int p2_init,
                        int dx, int dy, uint16_t* S, uint16_t* contrib)
{
    const int npaths = (dy == 0) ? H : W;
    const int maxlen = (W > H ? W : H);
    int64_t* pos = (int64_t*)malloc(sizeof(int64_t) * (size_t)maxlen);
    uint8_t* prev = (uint8_t*)malloc((size_t)D + 2);
    uint8_t* cur = (uint8_t*)malloc((size_t)D + 2);
    const int64_t npix = (int64_t)W * H;

    for (int path = 0; path < npaths; ++path) {
        const int n = sgmo_walk_path(W, H, dx, dy, path, pos);
        uint8_t min_prev = UINT8_MAX, g_prev = 0;
        for (int k = 0; k < n; ++k) {
            const int64_t p = pos[k];
            if (p < 0 || p >= npix) continue; /* elided out-of-bounds visit */
            const uint8_t* c = C + (size_t)p * D;
            const uint8_t g = img[p];
            uint8_t mn = UINT8_MAX;
            /* prev/cur are stored with one sentinel slot on each side: index d+1 <-> disparity d */
            if (k == 0) {
                for (int d = 0; d < D; ++d) { cur[d + 1] = c[d]; if (c[d] < mn) mn = c[d]; }
            } else {
                const int dg = abs((int)g - (int)g_prev);
                int pen2 = p2_init / (dg + 1);
                if (pen2 < p1) pen2 = p1;
                const uint16_t l4 = (uint16_t)(min_prev + pen2);
                for (int d = 0; d < D; ++d) {
                    uint16_t m = prev[d + 1];
                    const uint16_t l2 = (uint16_t)(prev[d] + p1);
                    const uint16_t l3 = (uint16_t)(prev[d + 2] + p1);
                    if (l2 < m) m = l2;
                    if (l3 < m) m = l3;
                    if (l4 < m) m = l4;
                    const uint8_t v = (uint8_t)(c[d] + m - min_prev);
                    cur[d + 1] = v;
                    if (v < mn) mn = v;
                }
            }
            cur[0] = UINT8_MAX; cur[D + 1] = UINT8_MAX;
            uint16_t* s = S + (size_t)p * D;
            for (int d = 0; d < D; ++d) s[d] = (uint16_t)(s[d] + cur[d + 1]);
            if (contrib) {
                uint16_t* q = contrib + (size_t)p * D;
                for (int d = 0; d < D; ++d) q[d] = (uint16_t)(q[d] + cur[d + 1]);
            }
            min_prev = mn; g_prev = g;
            uint8_t* t = prev; prev = cur; cur = t;
        }
    }
    free(pos); free(prev); free(cur);
}

/* ------------------------------------------------------------------------------------------------
 * Winner-takes-all + uniqueness + sub-pixel (SGM.c:374-443).
 *   left view  (inverse=0): costs of pixel (y,x) are S[y,x,:]
 *   right view (inverse=1): cost of d is S[y,x+d,d] when x+d < W, else 65535         (:397-408)
 *   best = lowest d with the minimum cost (strict <)                                 (:390-393,401-404)
 *   uniqueness: sec = min over d != best; invalid if sec-min <= (uint16)(min*(1-ratio)) (:412-426)
 *   invalid if best is the first or last disparity                                   (:428-431)
 *   sub-pixel with int16 neighbours (65535 -> -1), int16 denominator clamped to >=1  (:432-440)
 * ---------------------------------------------------------------------------------------------- */
void sgmo_wta(const uint16_t* S, int W, int H, int dmin, int dmax, int check_unique,
              float uniqueness_ratio, int inverse, float* disp)
{
    const int D = dmax - dmin;
    uint16_t* local = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)D);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            uint16_t best_cost = UINT16_MAX, second = UINT16_MAX, best = 0;
            for (int d = dmin; d < dmax; ++d) {
                uint16_t c;
                if (!inverse)            c = S[((size_t)y * W + x) * D + (d - dmin)];
                else if (x + d < W)      c = S[((size_t)y * W + x + d) * D + (d - dmin)];
                else { local[d - dmin] = UINT16_MAX; continue; }
                local[d - dmin] = c;
                if (c < best_cost) { best_cost = c; best = (uint16_t)d; }
            }
            float* out = disp + (size_t)y * W + x;
            if (check_unique) {
                for (int d = dmin; d < dmax; ++d)
                    if (d != best && local[d - dmin] < second) second = local[d - dmin];
                if (second - best_cost <= (uint16_t)(best_cost * (1 - uniqueness_ratio))) { *out = SGMO_INVALID; continue; }
            }
            if (best == dmin || best == dmax - 1) { *out = SGMO_INVALID; continue; }
            if (best < dmin || best >= dmax) { *out = SGMO_INVALID; continue; } /* reference: out-of-array read; not reachable with check_unique or dmin==0 */
            const int16_t c1 = (int16_t)local[best - 1 - dmin];
            const int16_t c2 = (int16_t)local[best + 1 - dmin];
            int16_t denom = (int16_t)(c1 + c2 - 2 * best_cost);
            if (denom < 1) denom = 1;
            *out = (float)best + (float)(c1 - c2) / (denom * 2.0f);
        }
    free(local);
}

/* ------------------------------------------------------------------------------------------------
 * Left-right check (SGM.c:445-470): column in the right view = (int32)(x - d + 0.5) with a float
 * subtraction followed by a double addition; out of range -> invalid; right invalid -> keep;
 * |d - dR| > thres -> invalid.
 * ---------------------------------------------------------------------------------------------- */
void sgmo_lrcheck(float* disp_left, const float* disp_right, int W, int H, float thres)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float* dl = disp_left + (size_t)y * W + x;
            const float d = *dl;
            if (d == SGMO_INVALID) continue;
            const float shifted = (float)x - d;            /* uint16 - float -> float  (:454) */
            const int32_t xr = (int32_t)((double)shifted + 0.5);
            if (xr < 0 || xr >= W) { *dl = SGMO_INVALID; continue; }
            const float dr = disp_right[(size_t)y * W + xr];
            if (dr == SGMO_INVALID) continue;
            if (fabs(d - dr) > thres) *dl = SGMO_INVALID;
        }
}

/* ------------------------------------------------------------------------------------------------
 * Speckle removal (SGM.c:585-642): connected components over the 8-neighbourhood, two valid pixels
 * being connected when |d_a - d_b| <= diff_insame; components with fewer than min_area pixels become
 * invalid.  The edge relation is symmetric, so the result does not depend on the visiting order.
 * ---------------------------------------------------------------------------------------------- */
void sgmo_remove_speckles(float* disp, int W, int H, float diff_insame, int min_area)
{
    const size_t N = (size_t)W * H;
    uint8_t* seen = (uint8_t*)calloc(N, 1);
    uint32_t* comp = (uint32_t*)malloc(sizeof(uint32_t) * N);
    for (size_t seed = 0; seed < N; ++seed) {
        if (seen[seed] || disp[seed] == SGMO_INVALID) continue;
        size_t head = 0, tail = 0;
        comp[tail++] = (uint32_t)seed; seen[seed] = 1;
        while (head < tail) {
            const uint32_t p = comp[head++];
            const int py = (int)(p / (uint32_t)W), px = (int)(p % (uint32_t)W);
            const float base = disp[p];
            for (int oy = -1; oy <= 1; ++oy)
                for (int ox = -1; ox <= 1; ++ox) {
                    const int qy = py + oy, qx = px + ox;
                    if ((oy == 0 && ox == 0) || qy < 0 || qy >= H || qx < 0 || qx >= W) continue;
                    const size_t q = (size_t)qy * W + qx;
                    if (!seen[q] && disp[q] != SGMO_INVALID && fabs(disp[q] - base) <= diff_insame) {
                        seen[q] = 1; comp[tail++] = (uint32_t)q;
                    }
                }
        }
        if (tail < (size_t)min_area)
            for (size_t k = 0; k < tail; ++k) disp[comp[k]] = SGMO_INVALID;
    }
    free(seen); free(comp);
}

/* ------------------------------------------------------------------------------------------------
 * 3x3 median, IN PLACE in raster order (SGM.c:120 passes the same buffer as input and output to
 * SGM.c:525-557): interior pixels only; each output is the 5th smallest of the 9 window values, where
 * the row above and the left neighbour have already been overwritten.  +inf takes part like any
 * other value (SGM.c:496-523).
 * ---------------------------------------------------------------------------------------------- */
static float fifth_smallest_of_9(const float* v)
{
    float s[9];
    memcpy(s, v, sizeof s);
    for (int i = 0; i < 5; ++i) {            /* partial selection sort: fix s[0..4] */
        int m = i;
        for (int j = i + 1; j < 9; ++j) if (s[j] < s[m]) m = j;
        const float t = s[i]; s[i] = s[m]; s[m] = t;
    }
    return s[4];
}

void sgmo_median3_inplace(float* disp, int W, int H)
{
    float w[9];
    for (int y = 1; y < H - 1; ++y)
        for (int x = 1; x < W - 1; ++x) {
            int n = 0;
            for (int oy = -1; oy <= 1; ++oy)
                for (int ox = -1; ox <= 1; ++ox) w[n++] = disp[(size_t)(y + oy) * W + (x + ox)];
            disp[(size_t)y * W + x] = fifth_smallest_of_9(w);
        }
}

/* ------------------------------------------------------------------------------------------------
 * Pipeline (SGM.c:37-125).  Direction order = SGM.c:213-220; num_paths == 4 keeps the first four.
 * ---------------------------------------------------------------------------------------------- */
static const int8_t kDirs[8][2] = { {1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, -1}, {1, -1}, {-1, 1} };

static int run(const sgmo_params* prm, const uint8_t* left, const uint8_t* right, float* disp_out,
               const sgmo_taps* taps, int hot_only)
{
    if (!prm || prm->width <= 0 || prm->height <= 0) return -1;             /* SGM.c:43-45 */
    if (prm->max_disparity <= prm->min_disparity) return -1;                /* SGM.c:46-48 */
    if (!left || !right) return -1;                                         /* SGM.c:73-75 */
    const int W = prm->width, H = prm->height, D = prm->max_disparity - prm->min_disparity;
    const size_t N = (size_t)W * H, V = N * (size_t)D;

    uint32_t* cl = (uint32_t*)calloc(N, sizeof(uint32_t));
    uint32_t* cr = (uint32_t*)calloc(N, sizeof(uint32_t));
    uint8_t* C = (uint8_t*)malloc(V);
    uint16_t* S = (uint16_t*)calloc(V, sizeof(uint16_t));                   /* SGM.c:57 */
    float* dl = (float*)malloc(sizeof(float) * N);
    float* dr = (float*)malloc(sizeof(float) * N);
    if (!cl || !cr || !C || !S || !dl || !dr) { free(cl); free(cr); free(C); free(S); free(dl); free(dr); return -2; }

    const int cw = prm->census_w ? prm->census_w : 5, ch = prm->census_h ? prm->census_h : 5;
    uint64_t *cl64 = NULL, *cr64 = NULL;
    if (cw == 5 && ch == 5) {
        sgmo_census5x5(left, W, H, cl);                                     /* SGM.c:82-83 */
        sgmo_census5x5(right, W, H, cr);
        sgmo_cost(cl, cr, W, H, prm->min_disparity, prm->max_disparity, C); /* SGM.c:89 */
    } else {                                                                /* extension, parity unpinned */
        cl64 = (uint64_t*)calloc(N, sizeof(uint64_t));
        cr64 = (uint64_t*)calloc(N, sizeof(uint64_t));
        if (!cl64 || !cr64 || cw * ch > 64 || !(cw & 1) || !(ch & 1)) {
            free(cl64); free(cr64); free(cl); free(cr); free(C); free(S); free(dl); free(dr);
            return -1;
        }
        sgmo_census(left, W, H, cw, ch, cl64);
        sgmo_census(right, W, H, cw, ch, cr64);
        sgmo_cost64(cl64, cr64, W, H, prm->min_disparity, prm->max_disparity, C);
    }
    const int ndir = (prm->num_paths == 4) ? 4 : 8;
    for (int r = 0; r < ndir; ++r) {                                        /* SGM.c:94,213-220 */
        uint16_t* contrib = (taps && taps->path_cost[r]) ? taps->path_cost[r] : NULL;
        if (contrib) memset(contrib, 0, V * sizeof(uint16_t));
        sgmo_aggregate_dir(left, C, W, H, D, prm->p1, prm->p2_init, kDirs[r][0], kDirs[r][1], S, contrib);
    }
    sgmo_wta(S, W, H, prm->min_disparity, prm->max_disparity, prm->check_unique, prm->uniqueness_ratio, 0, dl); /* :99 */
    if (taps) {
        if (taps->census_left)  memcpy(taps->census_left, cl, N * sizeof(uint32_t));
        if (taps->census_right) memcpy(taps->census_right, cr, N * sizeof(uint32_t));
        for (size_t i = 0; i < N; ++i) {
            if (taps->census64_left)  taps->census64_left[i] = cl64 ? cl64[i] : cl[i];
            if (taps->census64_right) taps->census64_right[i] = cr64 ? cr64[i] : cr[i];
        }
        if (taps->cost)         memcpy(taps->cost, C, V);
        if (taps->aggr)         memcpy(taps->aggr, S, V * sizeof(uint16_t));
        if (taps->disp_left_wta) memcpy(taps->disp_left_wta, dl, N * sizeof(float));
    }
    if (prm->check_lr) {                                                    /* SGM.c:102-111 */
        sgmo_wta(S, W, H, prm->min_disparity, prm->max_disparity, prm->check_unique, prm->uniqueness_ratio, 1, dr);
        if (taps && taps->disp_right) memcpy(taps->disp_right, dr, N * sizeof(float));
        sgmo_lrcheck(dl, dr, W, H, prm->lrcheck_thres);
    }
    if (taps && taps->disp_lr) memcpy(taps->disp_lr, dl, N * sizeof(float));
    if (!hot_only) {
        if (prm->remove_speckles) sgmo_remove_speckles(dl, W, H, 1.0f, prm->min_speckle_area);   /* SGM.c:113-117 */
        if (taps && taps->disp_speckle) memcpy(taps->disp_speckle, dl, N * sizeof(float));
        if (prm->median) sgmo_median3_inplace(dl, W, H);                                          /* SGM.c:120 */
    }
    if (disp_out) memcpy(disp_out, dl, N * sizeof(float));                  /* SGM.c:122 */
    free(cl64); free(cr64);
    free(cl); free(cr); free(C); free(S); free(dl); free(dr);
    return 0;
}

int sgmo_match(const sgmo_params* prm, const uint8_t* left, const uint8_t* right, float* disp_out, const sgmo_taps* taps)
{
    return run(prm, left, right, disp_out, taps, 0);
}

int sgmo_match_hotpath(const sgmo_params* prm, const uint8_t* left, const uint8_t* right, float* disp_out)
{
    return run(prm, left, right, disp_out, NULL, 1);
}
