/*
 * SemiGlobalMatching.h -- drop-in public header of the B200-native SGM library (libsgm_b200.so).
 *
 * Same names, types, field order and calling convention as the reference's public interface
 * (/root/reference/SemiGlobalMatching/SemiGlobalMatching/SemiGlobalMatching.h), so a caller such as
 * the reference's main.c compiles against this header unchanged and links against libsgm_b200.so
 * instead of the reference's SemiGlobalMatching.c:
 *
 *   SGMOption       replaces  SemiGlobalMatching.h:24-40   (x86-64 SysV: sizeof 28, align 4)
 *   SGM_Initialize  replaces  SemiGlobalMatching.h:78 / SemiGlobalMatching.c:37-66
 *   SGM_Reset       replaces  SemiGlobalMatching.h:79 / SemiGlobalMatching.c:128-132
 *   SGM_Match       replaces  SemiGlobalMatching.h:80 / SemiGlobalMatching.c:68-125
 *
 * Differences that a caller can observe (all documented in DESIGN.md):
 *   - no compile-time size limits: device buffers are sized from (width, height, disparity range)
 *     at SGM_Initialize; the MAX_* macros below are kept only so that callers which size their own
 *     arrays with them (main.c:81) still compile;
 *   - the reference's global buffers (census_*_buffer, cost_*_buffer, disp_*_buffer) and the global
 *     `sgm` instance are not exported; use SGMB_GetStage() from sgm_b200.h to read a stage;
 *   - SGM_Match zeroes its accumulators on every call (the reference only does so in
 *     SGM_Initialize, SemiGlobalMatching.c:57);
 *   - additional `false` returns: no usable B200 / CUDA error / disparity range > 256 / negative
 *     P1 or P2 / out of memory.  There is no CPU fallback.
 */
#ifndef SEMI_GLOBAL_MATCHING_H
#define SEMI_GLOBAL_MATCHING_H

#include <stdint.h>
#include <stdbool.h>
#include <float.h>
#include <math.h> /* INFINITY */

#define INVALID_FLOAT (INFINITY) /* marks an invalid disparity; SemiGlobalMatching.h:12 */

/* Historical compile-time maxima of the reference (SemiGlobalMatching.h:14-19).  This library does
 * not use them; they are the sizes the reference demo (main.c:81,102) allocates with. */
#define MAX_IMG_WIDTH          450
#define MAX_IMG_HEIGHT         375
#define MAX_DISPARITY_RANGE    64
#define FILTER_WINDOW_SIZE     3
#define MAX_IMG_SIZE           (MAX_IMG_WIDTH * MAX_IMG_HEIGHT)
#define MAX_DISP_IMG_SIZE      (MAX_IMG_WIDTH * MAX_IMG_HEIGHT * MAX_DISPARITY_RANGE)

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint8_t   num_paths;          /* 4: the four axis-aligned directions; anything else: all 8       */
    uint16_t  min_disparity;      /* disparities searched: [min_disparity, max_disparity)            */
    uint16_t  max_disparity;

    bool      is_check_unique;    /* reject pixels whose best and second-best costs are too close    */
    float     uniqueness_ratio;

    bool      is_check_lr;        /* left-right consistency check                                    */
    float     lrcheck_thres;

    bool      is_remove_speckles; /* invalidate connected regions smaller than min_speckle_area      */
    uint16_t  min_speckle_area;

    int16_t   p1;                 /* penalty for a disparity change of 1                             */
    int16_t   p2_init;            /* penalty numerator for larger changes: max(p1, p2_init/(|dI|+1)) */
} SGMOption;

/* Copies *option, (re)allocates device buffers for width x height x (max-min) and prepares the
 * launch plan.  false: width == 0, height == 0, max_disparity <= min_disparity (as the reference),
 * or an unsupported configuration / CUDA failure (see above). */
bool SGM_Initialize(uint16_t width, uint16_t height, const SGMOption* option);

/* Forget the current configuration and initialise again. */
bool SGM_Reset(uint16_t width, uint16_t height, const SGMOption* option);

/* img_left / img_right: 8-bit grey, row-major, stride == width, width*height bytes, host memory.
 * disp_left: caller-allocated float[width*height]; receives absolute disparities (INVALID_FLOAT for
 * rejected pixels) after census -> cost -> aggregation -> WTA/sub-pixel -> LR check -> speckle
 * filter -> in-place 3x3 median, bit-identical to the reference.  Blocking.  false: not initialised,
 * NULL image pointer, or CUDA failure. */
bool SGM_Match(const uint8_t* img_left, const uint8_t* img_right, float* disp_left);

#ifdef __cplusplus
}
#endif

#endif /* SEMI_GLOBAL_MATCHING_H */
