// runcuda_adapter.cu — `int runcuda(GlobalState &gs)` (reference gipuma.h:2) on top of libgipuma_b200.
//
// This is the ONLY translation unit that knows the reference's managed-memory boundary types
// (globalstate.h, algorithmparameters.h, cameraparameters.h, camera.h, linestate.h); it is compiled with
// -I<gipuma checkout> by whoever links main.cpp against this library (INTEGRATION.md shows the two-line
// CMake change).  It flattens GlobalState into the plain C-ABI of include/gipuma_b200.h:
//   gs.params                     -> gpm_params                 (fields read on device, algorithmparameters.h:53-84)
//   gs.cameras->cameras[0]        -> gpm_set_reference camera   (K_inv, M_inv, P_col34, C4, fx, alpha, K, f, baseline, R_orig_inv)
//   gs.cameras->cameras[subset[v]]-> gpm_set_view camera        (K, R, t4)
//   gs.cuArray[i]                 -> linear device copy of the float (or, with -color_processing, float4) image
//                                    (the texture objects gs.imgs[] are not used)
//   gs.lines->norm4 / c           <- gpm_get_state (managed memory, visible to the host on return: main.cpp:976-985)
// Behaviour kept from the reference: synchronous, returns 0, prints the same progress lines, CUDA/argument
// failures print a message and exit(EXIT_FAILURE) like checkCudaErrors (helper_cuda.h:890-905).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "globalstate.h"
#include "algorithmparameters.h"
#include "cameraparameters.h"
#include "linestate.h"
#include "gipuma_b200.h"

static unsigned long long g_adapter_seed = 0;
static bool g_adapter_seed_set = false;

// Pin the curand_init seed (the reference draws it from clock64(), gipuma.cu:1019).  Also settable with the
// GIPUMA_SEED environment variable; otherwise a time-based seed is used, as in the reference.
extern "C" void gpm_adapter_set_seed(unsigned long long seed)
{
    g_adapter_seed = seed;
    g_adapter_seed_set = true;
}

static void die(const char* what, int rc)
{
    fprintf(stderr, "gipuma_b200 runcuda: %s failed (%d): %s\n", what, rc, gpm_last_error());
    cudaDeviceReset();
    exit(EXIT_FAILURE);
}
#define GPM_CHECK(call) do { int rc_ = (call); if (rc_ != 0) die(#call, rc_); } while (0)

static void fill_camera(gpm_camera& g, const Camera_cu& c)
{
    memset(&g, 0, sizeof(g));
    memcpy(g.K, c.K, 9 * sizeof(float));
    memcpy(g.K_inv, c.K_inv, 9 * sizeof(float));
    memcpy(g.R, c.R, 9 * sizeof(float));
    memcpy(g.M_inv, c.M_inv, 9 * sizeof(float));
    memcpy(g.R_orig_inv, c.R_orig_inv, 9 * sizeof(float));
    g.t[0] = c.t4.x;  g.t[1] = c.t4.y;  g.t[2] = c.t4.z;
    g.C[0] = c.C4.x;  g.C[1] = c.C4.y;  g.C[2] = c.C4.z;
    g.P_col34[0] = c.P_col34.x;  g.P_col34[1] = c.P_col34.y;  g.P_col34[2] = c.P_col34.z;
    g.fx = c.fx;  g.fy = c.fy;  g.f = c.f;  g.alpha = c.alpha;  g.baseline = c.baseline;
}

int runcuda(GlobalState& gs)
{
    const AlgorithmParameters& a = *gs.params;
    CameraParameters_cu& cpc = *gs.cameras;
    const int rows = cpc.rows, cols = cpc.cols, V = cpc.viewSelectionSubsetNumber;
    const bool col = a.color_processing;                      // T = float4 (gipuma.cu:1965-1966)
    const size_t texel = col ? 4 * sizeof(float) : sizeof(float);
    int device = 0;
    cudaGetDevice(&device);                                   // main.cpp:690 selected it already
    gpm_ctx* ctx = nullptr;
    GPM_CHECK(gpm_create(&ctx, device, cols, rows, V > 0 ? V : 1));

    gpm_params p;
    memset(&p, 0, sizeof(p));
    p.box_hsize = a.box_hsize;  p.box_vsize = a.box_vsize;
    p.tau_color = a.tau_color;  p.tau_gradient = a.tau_gradient;  p.alpha = a.alpha;  p.gamma = a.gamma;
    p.min_disparity = a.min_disparity;  p.max_disparity = a.max_disparity;
    p.iterations = a.iterations;  p.n_best = a.n_best;  p.cost_comb = a.cost_comb;  p.good_factor = a.good_factor;
    p.depthMin = cpc.cameras[REFERENCE].depthMin;             // ISDISPDEPTHWITHINBORDERS, gipuma.cu:829-830
    p.depthMax = cpc.cameras[REFERENCE].depthMax;
    GPM_CHECK(gpm_set_params(ctx, &p));

    float* lin = nullptr;
    if (cudaMalloc(&lin, (size_t)rows * cols * texel) != cudaSuccess) die("cudaMalloc", -2);
    gpm_camera cam;
    fill_camera(cam, cpc.cameras[REFERENCE]);
    cam.f = cpc.f;                                            // gipuma.cu:904 reads camParams.f; == cameras[0].f (cameraGeometryUtils.h:315-316)
    if (cudaMemcpy2DFromArray(lin, cols * texel, gs.cuArray[REFERENCE], 0, 0, cols * texel, rows,
                              cudaMemcpyDeviceToDevice) != cudaSuccess) die("cudaMemcpy2DFromArray", -2);
    if (col) GPM_CHECK(gpm_set_reference_color(ctx, lin, 0, 1, &cam));
    else GPM_CHECK(gpm_set_reference(ctx, lin, 0, 1, &cam));
    for (int v = 0; v < V; v++) {
        const int idx = cpc.viewSelectionSubset[v];           // gipuma.cu:743
        fill_camera(cam, cpc.cameras[idx]);
        if (cudaMemcpy2DFromArray(lin, cols * texel, gs.cuArray[idx], 0, 0, cols * texel, rows,
                                  cudaMemcpyDeviceToDevice) != cudaSuccess) die("cudaMemcpy2DFromArray", -2);
        if (col) GPM_CHECK(gpm_set_view_color(ctx, v, lin, 0, 1, &cam));
        else GPM_CHECK(gpm_set_view(ctx, v, lin, 0, 1, &cam));
        cudaDeviceSynchronize();                              // `lin` is reused for the next view
    }
    GPM_CHECK(gpm_set_num_views(ctx, V));

    unsigned long long seed = g_adapter_seed;
    if (!g_adapter_seed_set) {
        const char* env = getenv("GIPUMA_SEED");
        seed = env ? strtoull(env, nullptr, 0)
                   : (unsigned long long)std::chrono::high_resolution_clock::now().time_since_epoch().count();
    }
    GPM_CHECK(gpm_set_rng(ctx, seed, GPM_RNG_REFERENCE));
    // The reference picks its sweep at compile time: `#define SMALLKERNEL` (gipuma.cu:33, the shipped setting) runs the six
    // 4-neighbour kernels per iteration, without it the fused 20-neighbour kernels (gipuma.cu:1913-1940).  Here: build the
    // adapter with -DGPM_ADAPTER_NEIGHBOURS=20 or set GIPUMA_B200_NEIGHBOURS=20.
#ifndef GPM_ADAPTER_NEIGHBOURS
#define GPM_ADAPTER_NEIGHBOURS 8
#endif
    {
        const char* env = getenv("GIPUMA_B200_NEIGHBOURS");
        const int nb = env ? atoi(env) : GPM_ADAPTER_NEIGHBOURS;
        if (nb != 8) GPM_CHECK(gpm_set_option(ctx, "neighbours", nb));
    }

    size_t avail = 0, total = 0;
    cudaMemGetInfo(&avail, &total);
    printf("Device memory used: %fMB\n", (total - avail) / 1000000.0f);      // gipuma.cu:1898-1903
    printf("Blocksize is %dx%d\n", a.box_hsize, a.box_vsize);
    printf("Number of iterations is %d\n", a.iterations);
    printf("Iteration ");
    for (int it = 0; it < a.iterations; it++) printf("%d ", it + 1);
    printf("\n");
    float ms = 0.f;
    GPM_CHECK(gpm_run(ctx, &ms));
    printf("\t\tTotal time needed for computation: %f seconds\n", ms / 1000.f);   // gipuma.cu:1952

    // results into the caller-owned managed arrays (linestate.h:10-11); host reads them right after (main.cpp:976-985)
    GPM_CHECK(gpm_get_state(ctx, reinterpret_cast<float*>(gs.lines->norm4), gs.lines->c, 1));
    cudaDeviceSynchronize();
    cudaFree(lin);
    gpm_destroy(ctx);
    return 0;                                                  // gipuma.cu:1969
}
