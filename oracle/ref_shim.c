/*
 * ref_shim.c -- tap shim around the UNMODIFIED reference translation unit.  TEST INFRASTRUCTURE ONLY.
 *
 * oracle/build_ref.py compiles this file once per (W, H, D, variant) into oracle/_ref/libsgm_ref_*.so:
 *   -I<tmp>        holds a generated copy of the reference header whose three MAX_* macros are rewritten
 *                  to W, H + guard rows, D (the header has no #ifndef guards, so -D cannot do it);
 *                  including it first defines SEMI_GLOBAL_MATCHING_H, so the reference .c's own
 *                  #include of the original header becomes a no-op;
 *   -DSGM_REF_C=   path of the reference SemiGlobalMatching.c under /root/reference (or, for the
 *                  D256 / P4 variants, of a generated copy carrying the one-token patches listed in
 *                  build_ref.py).
 * Including the .c here gives access to its static stage functions, so every stage can be tapped
 * without touching the reference code.  -fno-toplevel-reorder keeps the BSS objects in source order,
 * so the reference's two out-of-bounds visits land in guard rows.
 */
#include "SemiGlobalMatching.h"
#include SGM_REF_C

#include <string.h>

int ref_max_width(void)  { return MAX_IMG_WIDTH; }
int ref_max_height(void) { return MAX_IMG_HEIGHT; }
int ref_max_disp(void)   { return MAX_DISPARITY_RANGE; }
int ref_sizeof_option(void) { return (int)sizeof(SGMOption); }

/* Distances (bytes) that prove the guard layout: see build_ref.py::check_layout. */
long ref_gap_init_to_aggr(void) { return (long)((char*)cost_aggr_buffer - (char*)cost_init_buffer); }
long ref_gap_left_to_init(void) { return (long)((char*)cost_init_buffer - (char*)census_left_buffer); }

/* Same call sequence as SGM_Match (reference .c lines 68-125) with a copy-out after every stage.
 * Any output pointer may be NULL.  Returns 1 on success like SGM_Match. */
int ref_match_staged(const uint8_t* left, const uint8_t* right,
                     uint32_t* census_l, uint32_t* census_r, uint8_t* cost, uint16_t* aggr,
                     float* disp_left_wta, float* disp_right, float* disp_lr, float* disp_speckle,
                     float* disp_final)
{
    if (!sgm.is_initialized || !left || !right) return 0;
    const size_t n = (size_t)sgm.width * sgm.height, v = n * sgm.disp_range;
    sgm.img_left = left;
    sgm.img_right = right;
    census_transform_5x5(sgm.img_left, sgm.census_left);
    census_transform_5x5(sgm.img_right, sgm.census_right);
    if (census_l) memcpy(census_l, sgm.census_left, n * sizeof(uint32_t));
    if (census_r) memcpy(census_r, sgm.census_right, n * sizeof(uint32_t));
    ComputeCost(sgm.census_left, sgm.census_right, sgm.cost_init);
    if (cost) memcpy(cost, sgm.cost_init, v);
    CostAggregation();
    if (aggr) memcpy(aggr, sgm.cost_aggr, v * sizeof(uint16_t));
    ComputeDisparity(sgm.cost_aggr, sgm.disp_left, 0);
    if (disp_left_wta) memcpy(disp_left_wta, sgm.disp_left, n * sizeof(float));
    if (sgm.option.is_check_lr) {
        ComputeDisparity(sgm.cost_aggr, sgm.disp_right, 1);
        if (disp_right) memcpy(disp_right, sgm.disp_right, n * sizeof(float));
        LRCheck(sgm.disp_left, sgm.disp_right);
    }
    if (disp_lr) memcpy(disp_lr, sgm.disp_left, n * sizeof(float));
    if (sgm.option.is_remove_speckles) RemoveSpeckles(sgm.disp_left, 1);
    if (disp_speckle) memcpy(disp_speckle, sgm.disp_left, n * sizeof(float));
    MedianFilter(sgm.disp_left, sgm.disp_left, FILTER_WINDOW_SIZE);
    if (disp_final) memcpy(disp_final, sgm.disp_left, n * sizeof(float));
    return 1;
}

/* Contribution of ONE direction to S: zero S, run CostAggregate(dx,dy) on the census/cost left by the
 * last ref_match_staged()/SGM_Match() call (whose left image must still be alive), copy out. */
int ref_aggregate_dir(int dx, int dy, uint16_t* out)
{
    if (!sgm.is_initialized || !sgm.img_left || !out) return 0;
    const size_t v = (size_t)sgm.width * sgm.height * sgm.disp_range;
    memset(sgm.cost_aggr, 0, sizeof(uint16_t) * (size_t)MAX_DISP_IMG_SIZE);
    CostAggregate(sgm.img_left, sgm.cost_init, sgm.cost_aggr, (int8_t)dx, (int8_t)dy);
    memcpy(out, sgm.cost_aggr, v * sizeof(uint16_t));
    return 1;
}

/* Hot path only (census .. LR check) for CPU-baseline timing; result stays in disp_left_buffer. */
int ref_match_hotpath(const uint8_t* left, const uint8_t* right, float* disp_out)
{
    if (!sgm.is_initialized || !left || !right) return 0;
    sgm.img_left = left;
    sgm.img_right = right;
    census_transform_5x5(sgm.img_left, sgm.census_left);
    census_transform_5x5(sgm.img_right, sgm.census_right);
    ComputeCost(sgm.census_left, sgm.census_right, sgm.cost_init);
    CostAggregation();
    ComputeDisparity(sgm.cost_aggr, sgm.disp_left, 0);
    if (sgm.option.is_check_lr) {
        ComputeDisparity(sgm.cost_aggr, sgm.disp_right, 1);
        LRCheck(sgm.disp_left, sgm.disp_right);
    }
    if (disp_out) memcpy(disp_out, sgm.disp_left, (size_t)sgm.width * sgm.height * sizeof(float));
    return 1;
}
