/* hx_api.h — C-ABI of the OpenCV-free driver ("harness") that stands in for the
 * reference's main.cpp around the `int runcuda(GlobalState&)` boundary
 * (reference gipuma.h:2, called from main.cpp:973).
 *
 * TEST / BENCH INFRASTRUCTURE.  The same source (hx_harness.cu) is built twice:
 *   - oracle/_ref/libhx_ref.so    : linked with the *reference* gipuma.cu (pinned, see
 *                                   oracle/build_ref.sh) — the parity oracle and the
 *                                   "reference CUDA path" timing arm;
 *   - oracle/_ref/libhx_dropin.so : linked with gipuma_b200's runcuda adapter — proves the
 *                                   new library is a drop-in behind the unchanged boundary.
 * Both fill GlobalState exactly as main.cpp:829-933,968 does (managed structs, float
 * textures: Linear / Wrap / unnormalised / ElementType, main.cpp:607-656).
 */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hx_params {
    int   box_hsize, box_vsize;          /* algorithmparameters.h:24-25 */
    float tau_color, tau_gradient;       /* :26-27 */
    float alpha, gamma;                  /* :28-29 */
    float min_disparity, max_disparity;  /* set by main.cpp:905-906 from the depth range */
    int   iterations;                    /* :31 */
    int   n_best, cost_comb;             /* :41-42 */
    float good_factor;                   /* :40 */
    int   color_processing;              /* :32; 1: images are n_images x rows x cols x 4 floats (B, G, R, unused) */
    float depthMin, depthMax;            /* main.cpp:898-903 -> cameras[0].depthMin/Max */
} hx_params;

typedef struct hx_camera {               /* the Camera_cu fields main.cpp/cameraGeometryUtils.h fill */
    float K[9], K_inv[9], R[9], R_orig_inv[9], M_inv[9];
    float P[12];
    float t[3], C[3];
    float fx, fy, f, alpha, baseline;
} hx_camera;

enum { HX_STEP_INIT = 0, HX_STEP_BLACK_CLOSE = 1, HX_STEP_BLACK_FAR = 2, HX_STEP_BLACK_REFINE = 3,
       HX_STEP_RED_CLOSE = 4, HX_STEP_RED_FAR = 5, HX_STEP_RED_REFINE = 6, HX_STEP_COMPUTE_DISP = 7,
       /* the fused 20-neighbour kernels the reference launches when SMALLKERNEL is not defined (gipuma.cu:1714-1725,
        * 1770-1781, 1937-1939): 12 axial + 8 knight-move neighbours, then plane refinement, one launch per colour */
       HX_STEP_BLACK_FUSED = 8, HX_STEP_RED_FUSED = 9 };

/* Full run through runcuda().  images: n_images x rows x cols float (index 0 = reference).
 * subset: indices into images/cams of the selected source views (viewSelectionSubset).
 * out_norm4: rows*cols*4 floats, out_cost: rows*cols floats.
 * out_times[0] = the "Total time needed for computation" the callee printed (seconds; sweeps +
 *                final kernel, init excluded — gipuma.cu:1908-1952), or -1 if not found;
 * out_times[1] = wall milliseconds around runcuda() (CUDA events, includes init).
 * Returns 0, or a negative error. */
int hx_run(const hx_params* prm, int rows, int cols, int n_images, const float* images,
           const hx_camera* cams, int n_sel, const int* subset, unsigned long long seed,
           float* out_norm4, float* out_cost, double* out_times);

/* Step-level run (reference backend only): optionally start from a given raw state
 * (in_norm4/in_cost may be NULL -> zeroed as LineState::resize does), launch the listed
 * kernels in order, return the raw state.  out_ms (may be NULL) gets per-step milliseconds. */
int hx_steps(const hx_params* prm, int rows, int cols, int n_images, const float* images,
             const hx_camera* cams, int n_sel, const int* subset, unsigned long long seed,
             const int* steps, int n_steps, const float* in_norm4, const float* in_cost,
             float* out_norm4, float* out_cost, float* out_ms);

/* Evaluate the iteration-time multi-view cost (pmCostMultiview_cu with the shared tile,
 * gipuma.cu:720-806 via 585-680) of given planes at every pixel, without accept logic
 * (reference backend only).  planes: rows*cols*4 (n.xyz, d). */
int hx_cost_eval(const hx_params* prm, int rows, int cols, int n_images, const float* images,
                 const hx_camera* cams, int n_sel, const int* subset,
                 const float* planes, float* out_cost);

const char* hx_backend(void);   /* "reference" or "dropin" */
int hx_max_views(void);         /* 32 (stock costVector) or 64 (pin P3) */

#ifdef __cplusplus
}
#endif
