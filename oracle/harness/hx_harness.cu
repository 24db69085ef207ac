/* hx_harness.cu — OpenCV-free stand-in for the reference's main.cpp around runcuda().
 *
 * TEST / BENCH INFRASTRUCTURE (see hx_api.h).  Compiled only where /root/reference exists
 * (it needs the reference's boundary headers: globalstate.h, algorithmparameters.h,
 * cameraparameters.h, camera.h, linestate.h); the resulting .so files travel to the GPU box.
 *
 * Two backends, selected at compile time:
 *   -DHX_BACKEND_REFERENCE : this TU #includes the reference gipuma.cu *unmodified* from
 *        /root/reference (or, for > 32 views, a sed-patched copy with costVector[64] — pin P3).
 *        Pins P1/P2 of SURVEY.md §8c are applied without touching the source, by macro:
 *          P1  clock64()  -> a fixed seed              (gipuma.cu:1019, the curand_init seed)
 *          P2  cudaMalloc -> malloc + memset(0)        (gipuma.cu:1840, the never-written gs.cs)
 *   (default)              : declares runcuda() extern; linked against gipuma_b200's adapter.
 *
 * What it mirrors from main.cpp (file:line are /root/reference/main.cpp):
 *   829      new GlobalState (managed; ctor allocates CameraParameters_cu + LineState)
 *   833      getCameraParameters -> Camera_cu fields (here: copied from hx_camera PODs)
 *   888-892  viewSelectionSubset[], 898-906 depthMin/Max + min/max_disparity, 916-933 params/rows/cols/lines
 *   607-656  addImageToTextureFloatGray: cudaArray + texture object, Linear/Wrap/unnormalised
 *   973      runcuda(*gs);  976-985 host reads gs->lines->norm4
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <fcntl.h>
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "globalstate.h"          /* reference boundary types (-I/root/reference) */
#include "algorithmparameters.h"
#include "cameraparameters.h"
#include "linestate.h"
#include "helper_cuda.h"
#include "hx_api.h"

#ifdef HX_BACKEND_REFERENCE
__managed__ unsigned long long hx_pin_seed = 0xC0FFEEULL;
static cudaError_t hx_malloc_zero(void** p, size_t n)
{
    cudaError_t e = cudaMalloc(p, n);
    if (e == cudaSuccess) e = cudaMemset(*p, 0, n);
    return e;
}
#define clock64() ((long long)hx_pin_seed)                              /* pin P1 */
#define cudaMalloc(p, n) hx_malloc_zero((void**)(p), (size_t)(n))       /* pin P2 */
#include HX_GIPUMA_CU                                                   /* the reference TU */
#undef cudaMalloc
#undef clock64
#ifndef HX_MAX_VIEWS
#define HX_MAX_VIEWS 32
#endif
static const char* kBackend = "reference";
#else
int runcuda(GlobalState& gs);                                           /* gipuma.h:2 */
#ifndef HX_MAX_VIEWS
#define HX_MAX_VIEWS 512
#endif
static const char* kBackend = "dropin";
extern "C" void gpm_adapter_set_seed(unsigned long long seed);
#endif

namespace {

struct Scene {
    GlobalState* gs = nullptr;
    AlgorithmParameters* prm = nullptr;
    int n_images = 0;
};

void copy9(float* dst, const float* src) { for (int i = 0; i < 9; i++) dst[i] = src[i]; }

int build_scene(Scene& sc, const hx_params* p, int rows, int cols, int n_images, const float* images,
                const hx_camera* cams, int n_sel, const int* subset)
{
    if (n_images < 1 || n_images > MAX_IMAGES || n_sel < 0 || n_sel > HX_MAX_VIEWS) return -2;
    sc.n_images = n_images;
    sc.prm = new AlgorithmParameters;             /* managed, as main.cpp:1211 */
    sc.gs = new GlobalState;                      /* main.cpp:829 */
    GlobalState* gs = sc.gs;
    AlgorithmParameters& a = *sc.prm;
    a.box_hsize = p->box_hsize;  a.box_vsize = p->box_vsize;
    a.tau_color = p->tau_color;  a.tau_gradient = p->tau_gradient;
    a.alpha = p->alpha;          a.gamma = p->gamma;
    a.min_disparity = p->min_disparity;  a.max_disparity = p->max_disparity;
    a.iterations = p->iterations;
    a.n_best = p->n_best;  a.cost_comb = p->cost_comb;  a.good_factor = p->good_factor;
    a.color_processing = p->color_processing != 0;
    a.depthMin = p->depthMin;  a.depthMax = p->depthMax;

    CameraParameters_cu& cpc = *gs->cameras;
    for (int i = 0; i < n_images; i++) {          /* cameraGeometryUtils.h:305-346 */
        Camera_cu& c = cpc.cameras[i];
        const hx_camera& h = cams[i];
        copy9(c.K, h.K);  copy9(c.K_inv, h.K_inv);  copy9(c.R, h.R);
        copy9(c.R_orig_inv, h.R_orig_inv);  copy9(c.M_inv, h.M_inv);
        for (int k = 0; k < 12; k++) c.P[k] = h.P[k];
        c.t4 = make_float4(h.t[0], h.t[1], h.t[2], 0.f);
        c.C4 = make_float4(h.C[0], h.C[1], h.C[2], 0.f);
        c.P_col34 = make_float4(h.P[3], h.P[7], h.P[11], 0.f);
        c.fx = h.fx;  c.fy = h.fy;  c.f = h.f;  c.alpha = h.alpha;  c.baseline = h.baseline;
        c.reference = (i == 0);
    }
    cpc.f = cams[0].f;
    for (int i = 0; i < n_sel; i++) cpc.viewSelectionSubset[i] = subset[i];   /* main.cpp:888-892 */
    cpc.viewSelectionSubsetNumber = n_sel;                                    /* :918 */
    cpc.cameras[0].depthMin = p->depthMin;                                    /* :898-903 */
    cpc.cameras[0].depthMax = p->depthMax;
    gs->params = sc.prm;                                                      /* :916 */
    cpc.cols = cols;  cpc.rows = rows;  a.cols = cols;  a.rows = rows;        /* :921-924 */
    gs->lines->n = rows * cols;                                               /* :927-933 */
    gs->lines->resize(rows * cols);
    gs->lines->s = cols;
    gs->lines->l = cols;

    const int ch = p->color_processing ? 4 : 1;   /* float4 images: addImageToTextureFloatColor, main.cpp:560-605 */
    for (int i = 0; i < n_images; i++) {          /* main.cpp:607-656 */
        cudaChannelFormatDesc desc = p->color_processing ? cudaCreateChannelDesc<float4>()
                                                         : cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
        checkCudaErrors(cudaMallocArray(&gs->cuArray[i], &desc, cols, rows));
        checkCudaErrors(cudaMemcpy2DToArray(gs->cuArray[i], 0, 0, images + (size_t)i * rows * cols * ch,
                                            cols * sizeof(float) * ch, cols * sizeof(float) * ch, rows,
                                            cudaMemcpyHostToDevice));
        cudaResourceDesc res;  memset(&res, 0, sizeof(res));
        res.resType = cudaResourceTypeArray;  res.res.array.array = gs->cuArray[i];
        cudaTextureDesc tex;  memset(&tex, 0, sizeof(tex));
        tex.addressMode[0] = cudaAddressModeWrap;  tex.addressMode[1] = cudaAddressModeWrap;
        tex.filterMode = cudaFilterModeLinear;  tex.readMode = cudaReadModeElementType;
        tex.normalizedCoords = 0;
        checkCudaErrors(cudaCreateTextureObject(&gs->imgs[i], &res, &tex, NULL));
    }
    return 0;
}

void free_scene(Scene& sc)
{
    if (!sc.gs) return;
    cudaDeviceSynchronize();
    for (int i = 0; i < sc.n_images; i++) {
        cudaDestroyTextureObject(sc.gs->imgs[i]);
        cudaFreeArray(sc.gs->cuArray[i]);
    }
    delete sc.gs;
    delete sc.prm;
    sc.gs = nullptr;
}

/* Run f() with stdout redirected to a temp file; return the text. */
template <class F>
void capture_stdout(F f, char* buf, size_t cap)
{
    fflush(stdout);
    char path[] = "/tmp/hx_stdout_XXXXXX";
    int tmp = mkstemp(path);
    int saved = dup(1);
    if (tmp >= 0) dup2(tmp, 1);
    f();
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
    buf[0] = 0;
    if (tmp >= 0) {
        lseek(tmp, 0, SEEK_SET);
        ssize_t n = read(tmp, buf, cap - 1);
        buf[n > 0 ? n : 0] = 0;
        close(tmp);
        unlink(path);
    }
}

}  // namespace

extern "C" const char* hx_backend(void) { return kBackend; }
extern "C" int hx_max_views(void) { return HX_MAX_VIEWS; }

extern "C" int hx_run(const hx_params* prm, int rows, int cols, int n_images, const float* images,
                      const hx_camera* cams, int n_sel, const int* subset, unsigned long long seed,
                      float* out_norm4, float* out_cost, double* out_times)
{
    Scene sc;
    int rc = build_scene(sc, prm, rows, cols, n_images, images, cams, n_sel, subset);
    if (rc) return rc;
#ifdef HX_BACKEND_REFERENCE
    hx_pin_seed = seed;
#else
    gpm_adapter_set_seed(seed);
#endif
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);  cudaEventCreate(&e1);
    static char text[1 << 16];
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    capture_stdout([&] { runcuda(*sc.gs); }, text, sizeof(text));        /* main.cpp:973 */
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (out_times) {
        out_times[0] = -1.0;
        const char* key = "Total time needed for computation:";
        const char* at = strstr(text, key);
        if (at) out_times[0] = atof(at + strlen(key));
        out_times[1] = ms;
    }
    cudaError_t err = cudaDeviceSynchronize();
    /* main.cpp:976-985: the host reads the managed result arrays directly */
    if (out_norm4) memcpy(out_norm4, sc.gs->lines->norm4, sizeof(float) * 4 * rows * cols);
    if (out_cost) memcpy(out_cost, sc.gs->lines->c, sizeof(float) * rows * cols);
    cudaEventDestroy(e0);  cudaEventDestroy(e1);
    free_scene(sc);
    return err == cudaSuccess ? 0 : -1;
}

#ifdef HX_BACKEND_REFERENCE
/* ---- step-level drivers over the reference kernels -------------------------------------- */
namespace {

/* Tile constants exactly as gipuma<T>() sets them (gipuma.cu:1844-1856). */
void ref_setup_tiles(GlobalState& gs, dim3& grid, dim3& block, dim3& grid16, dim3& block16, int& smem_elems)
{
    int rows = gs.cameras->rows, cols = gs.cameras->cols;
    WIN_RADIUS_W = (gs.params->box_hsize + 1) / 2;
    WIN_RADIUS_H = (gs.params->box_vsize + 1) / 2;
    TILE_W = 32;
    TILE_H = 32;
    SHARED_SIZE_W_m = TILE_W + WIN_RADIUS_W * 2;
    SHARED_SIZE_H = TILE_H + WIN_RADIUS_H * 2;
    SHARED_SIZE = SHARED_SIZE_W_m * SHARED_SIZE_H;
    cudaMemcpyToSymbol(SHARED_SIZE_W, &SHARED_SIZE_W_m, sizeof(SHARED_SIZE_W_m));
    smem_elems = SHARED_SIZE;
    grid = dim3((cols + 31) / 32, ((rows / 2) + 15) / 16);
    block = dim3(32, 16);
    grid16 = dim3((cols + 15) / 16, (rows + 15) / 16);
    block16 = dim3(16, 16);
}

/* Cost of a given plane at every pixel through the reference's own iteration-time device
 * function (pmCostMultiview_cu with the shared tile).  The kernel body around the call is
 * ours: whole tile loaded by all threads before anyone leaves. */
template <typename T>
__global__ void hx_cost_eval_kernel(GlobalState& gs, const float4* planes, float* out, int color)
{
    extern __shared__ __align__(16) unsigned char hx_smem[];
    T* hx_tile = reinterpret_cast<T*>(hx_smem);
    const int rows = gs.cameras->rows, cols = gs.cameras->cols;
    int2 p = make_int2(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y * blockDim.y + threadIdx.y);
    p.y = p.y * 2 + (((threadIdx.x & 1) != 0) ^ (color != 0) ? 1 : 0);
    int2 tile_offset = make_int2(blockIdx.x * 32 - WIN_RADIUS_W, blockIdx.y * 32 - WIN_RADIUS_H);
    for (int e = threadIdx.y * 32 + threadIdx.x; e < SHARED_SIZE; e += 512) {
        int I = e % SHARED_SIZE_W, J = e / SHARED_SIZE_W;
        hx_tile[e] = tex2D<T>(gs.imgs[REFERENCE], tile_offset.x + I + 0.5f, tile_offset.y + J + 0.5f);
    }
    __syncthreads();
    if (p.x >= cols || p.y >= rows) return;
    int box_hrad = (gs.params->box_hsize - 1) / 2, box_vrad = (gs.params->box_vsize - 1) / 2;
    const int center = p.y * cols + p.x;
    out[center] = pmCostMultiview_cu<T>(gs.imgs, hx_tile, tile_offset, p, planes[center], box_vrad,
                                            box_hrad, *gs.params, *gs.cameras, planes, 0);
}

}  // namespace

extern "C" int hx_steps(const hx_params* prm, int rows, int cols, int n_images, const float* images,
                        const hx_camera* cams, int n_sel, const int* subset, unsigned long long seed,
                        const int* steps, int n_steps, const float* in_norm4, const float* in_cost,
                        float* out_norm4, float* out_cost, float* out_ms)
{
    Scene sc;
    int rc = build_scene(sc, prm, rows, cols, n_images, images, cams, n_sel, subset);
    if (rc) return rc;
    GlobalState& gs = *sc.gs;
    hx_pin_seed = seed;
    if (in_norm4) memcpy(gs.lines->norm4, in_norm4, sizeof(float) * 4 * rows * cols);
    if (in_cost) memcpy(gs.lines->c, in_cost, sizeof(float) * rows * cols);
    cudaDeviceSetCacheConfig(cudaFuncCachePreferShared);                 /* gipuma.cu:1829 */
    checkCudaErrors(hx_malloc_zero((void**)&gs.cs, (size_t)rows * cols * sizeof(curandState)));
    dim3 grid, block, grid16, block16;
    int smem_elems = 0;
    ref_setup_tiles(gs, grid, block, grid16, block16, smem_elems);
    const bool col = prm->color_processing != 0;
    size_t smem = (size_t)smem_elems * (col ? sizeof(float4) : sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);  cudaEventCreate(&e1);
    for (int s = 0; s < n_steps; s++) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        if (col) switch (steps[s]) {
        case HX_STEP_INIT:         gipuma_init_cu2<float4><<<grid16, block16>>>(gs); break;
        case HX_STEP_BLACK_CLOSE:  gipuma_black_spatialPropClose_cu<float4><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_BLACK_FAR:    gipuma_black_spatialPropFar_cu<float4><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_BLACK_REFINE: gipuma_black_planeRefine_cu<float4><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_RED_CLOSE:    gipuma_red_spatialPropClose_cu<float4><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_RED_FAR:      gipuma_red_spatialPropFar_cu<float4><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_RED_REFINE:   gipuma_red_planeRefine_cu<float4><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_COMPUTE_DISP: gipuma_compute_disp<<<grid16, block16>>>(gs); break;
        case HX_STEP_BLACK_FUSED:  gipuma_black_cu<float4><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_RED_FUSED:    gipuma_red_cu<float4><<<grid, block, smem>>>(gs, 0); break;
        default: break;
        }
        else switch (steps[s]) {
        case HX_STEP_INIT:         gipuma_init_cu2<float><<<grid16, block16>>>(gs); break;
        case HX_STEP_BLACK_CLOSE:  gipuma_black_spatialPropClose_cu<float><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_BLACK_FAR:    gipuma_black_spatialPropFar_cu<float><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_BLACK_REFINE: gipuma_black_planeRefine_cu<float><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_RED_CLOSE:    gipuma_red_spatialPropClose_cu<float><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_RED_FAR:      gipuma_red_spatialPropFar_cu<float><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_RED_REFINE:   gipuma_red_planeRefine_cu<float><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_COMPUTE_DISP: gipuma_compute_disp<<<grid16, block16>>>(gs); break;
        case HX_STEP_BLACK_FUSED:  gipuma_black_cu<float><<<grid, block, smem>>>(gs, 0); break;
        case HX_STEP_RED_FUSED:    gipuma_red_cu<float><<<grid, block, smem>>>(gs, 0); break;
        default: break;
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (out_ms) out_ms[s] = ms;
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) fprintf(stderr, "hx_steps: %s\n", cudaGetErrorString(err));
    if (out_norm4) memcpy(out_norm4, gs.lines->norm4, sizeof(float) * 4 * rows * cols);
    if (out_cost) memcpy(out_cost, gs.lines->c, sizeof(float) * rows * cols);
    cudaFree(gs.cs);
    cudaEventDestroy(e0);  cudaEventDestroy(e1);
    free_scene(sc);
    return err == cudaSuccess ? 0 : -1;
}

extern "C" int hx_cost_eval(const hx_params* prm, int rows, int cols, int n_images, const float* images,
                            const hx_camera* cams, int n_sel, const int* subset,
                            const float* planes, float* out_cost)
{
    Scene sc;
    int rc = build_scene(sc, prm, rows, cols, n_images, images, cams, n_sel, subset);
    if (rc) return rc;
    GlobalState& gs = *sc.gs;
    dim3 grid, block, grid16, block16;
    int smem_elems = 0;
    ref_setup_tiles(gs, grid, block, grid16, block16, smem_elems);
    float4* d_planes = nullptr;
    float* d_out = nullptr;
    size_t n = (size_t)rows * cols;
    checkCudaErrors(hx_malloc_zero((void**)&d_planes, n * sizeof(float4)));
    checkCudaErrors(hx_malloc_zero((void**)&d_out, n * sizeof(float)));
    cudaMemcpy(d_planes, planes, n * sizeof(float4), cudaMemcpyHostToDevice);
    for (int color = 0; color < 2; color++) {
        if (prm->color_processing) hx_cost_eval_kernel<float4><<<grid, block, smem_elems * sizeof(float4)>>>(gs, d_planes, d_out, color);
        else hx_cost_eval_kernel<float><<<grid, block, smem_elems * sizeof(float)>>>(gs, d_planes, d_out, color);
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) fprintf(stderr, "hx_cost_eval: %s\n", cudaGetErrorString(err));
    cudaMemcpy(out_cost, d_out, n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_planes);  cudaFree(d_out);
    free_scene(sc);
    return err == cudaSuccess ? 0 : -1;
}
#else
extern "C" int hx_steps(const hx_params*, int, int, int, const float*, const hx_camera*, int, const int*,
                        unsigned long long, const int*, int, const float*, const float*, float*, float*, float*)
{
    return -100;   /* the drop-in boundary has no step-level entry; use the gpm_* C-ABI */
}
extern "C" int hx_cost_eval(const hx_params*, int, int, int, const float*, const hx_camera*, int, const int*,
                            const float*, float*)
{
    return -100;
}
#endif
