/* gipuma_b200.h — C-ABI of the Blackwell-native PatchMatch hot path.
 *
 * This library replaces exactly one thing in kysucix/gipuma: the work behind
 *      int runcuda(GlobalState &gs);                     (reference gipuma.h:2, gipuma.cu:1962-1970)
 * i.e. gipuma<T>() (gipuma.cu:1825-1960): random plane initialisation, red/black checkerboard
 * spatial propagation (close +-1 px, far +-5 px), plane refinement, the multi-view
 * adaptive-support-weight photo-consistency cost, and the final depth / world-normal output.
 *
 * Plain C, no torch / CUDA types in the signatures: device pointers are passed as void* plus an
 * `on_device` flag.  Every entry point returns 0 on success or a negative GPM_E_* code;
 * gpm_last_error() returns a thread-local description.  One gpm_ctx per device / per reference view;
 * contexts are independent (one per rank in multi-GPU runs).
 *
 * The runcuda() adapter that flattens the reference's managed GlobalState into these calls lives in
 * gipuma_b200/csrc/runcuda_adapter.cu (it is the only code that knows GlobalState); see INTEGRATION.md.
 */
#ifndef GIPUMA_B200_H
#define GIPUMA_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPM_MAX_VIEWS 64          /* source views per reference view (reference: costVector[32], gipuma.cu:736) */
#define GPM_MAX_BOX   25          /* window size limit of the reference's tile loader (gipuma.cu:1513-1522) */

enum {
    GPM_OK = 0,
    GPM_E_ARG = -1,               /* bad argument / unsupported parameter */
    GPM_E_CUDA = -2,              /* a CUDA call failed; see gpm_last_error() */
    GPM_E_STATE = -3              /* call sequence error (e.g. sweep before views are set) */
};

/* cost combination — algorithmparameters.h:17 */
enum { GPM_COMB_ALL = 0, GPM_COMB_BEST_N = 1, GPM_COMB_ANGLE = 2, GPM_COMB_GOOD = 3 };

/* RNG behaviour of the plane-refinement step.
 * GPM_RNG_REFERENCE reproduces what the reference does: gs.cs is cudaMalloc'ed and never written
 * (gipuma.cu:1840, 1608, 1702), so every refinement kernel starts each pixel from an all-zero XORWOW
 * state.  GPM_RNG_STATEFUL keeps a per-pixel XORWOW state seeded at init and advanced by every draw
 * (what the code evidently intended); it has no reference output to compare against. */
enum { GPM_RNG_REFERENCE = 0, GPM_RNG_STATEFUL = 1 };

/* The AlgorithmParameters fields the device path reads (algorithmparameters.h:53-84). */
typedef struct gpm_params {
    int   box_hsize, box_vsize;         /* odd, equal, <= GPM_MAX_BOX */
    float tau_color, tau_gradient;
    float alpha, gamma;
    float min_disparity, max_disparity; /* = f*baseline/depthMax, f*baseline/depthMin (main.cpp:905-906) */
    int   iterations;
    int   n_best;
    int   cost_comb;                    /* GPM_COMB_* */
    float good_factor;
    float depthMin, depthMax;           /* cameras[REFERENCE].depthMin/Max (main.cpp:898-903) */
} gpm_params;

/* The Camera_cu fields the device path reads (camera.h:7-62), 3x3 matrices row-major. */
typedef struct gpm_camera {
    float K[9], K_inv[9], R[9], M_inv[9], R_orig_inv[9];
    float t[3], C[3], P_col34[3];
    float fx, fy, f, alpha, baseline;
} gpm_camera;

typedef struct gpm_ctx gpm_ctx;

/* Create a context on CUDA device `device` for width x height images and up to `max_views`
 * source views.  Replaces the allocations of gipuma<T>() (gipuma.cu:1840) and main.cpp:927-933. */
int gpm_create(gpm_ctx** out, int device, int width, int height, int max_views);
void gpm_destroy(gpm_ctx* ctx);

int gpm_set_params(gpm_ctx* ctx, const gpm_params* p);

/* Reference image + camera (index REFERENCE=0 of gs.imgs / gs.cameras->cameras, config.h:21).
 * `img` is row-major float, `pitch_bytes` between rows; host or device memory. */
int gpm_set_reference(gpm_ctx* ctx, const float* img, size_t pitch_bytes, int on_device, const gpm_camera* cam);

/* Source view `v` (0-based position in viewSelectionSubset, main.cpp:888-892). */
int gpm_set_view(gpm_ctx* ctx, int v, const float* img, size_t pitch_bytes, int on_device, const gpm_camera* cam);

/* Colour images: the reference's -color_processing path (T = float4, gipuma.cu:1965-1966; images built at
 * main.cpp:560-605).  `rgba` is row-major float4 per pixel (x, y, z = the three colour channels, w ignored — the
 * reference's float4 operators drop it, vector_operations.h:3-38).  A context holds either float or float4 images. */
int gpm_set_reference_color(gpm_ctx* ctx, const float* rgba, size_t pitch_bytes, int on_device, const gpm_camera* cam);
int gpm_set_view_color(gpm_ctx* ctx, int v, const float* rgba, size_t pitch_bytes, int on_device, const gpm_camera* cam);

/* Number of source views actually used (viewSelectionSubsetNumber, main.cpp:918). */
int gpm_set_num_views(gpm_ctx* ctx, int n_views);

/* Seed of the per-pixel curand_init(seed, y, x) at initialisation (gipuma.cu:1019; the reference uses
 * clock64()) and the refinement RNG mode. */
int gpm_set_rng(gpm_ctx* ctx, unsigned long long seed, int mode);

/* Overwrite / read the raw per-pixel state: planes [height*width] float4 (n.xyz in the reference
 * camera frame, w = plane distance d) and costs [height*width] float, both row-major with stride =
 * width — the LineState layout (linestate.h:8-24).  Either pointer may be NULL. */
int gpm_set_state(gpm_ctx* ctx, const float* norm4, const float* cost, int on_device);
int gpm_get_state(gpm_ctx* ctx, float* norm4, float* cost, int on_device);

/* gipuma_init_cu2 (gipuma.cu:996-1051): random plane + initial cost for every pixel. */
int gpm_init(gpm_ctx* ctx);

/* `iterations` red/black sweeps (gipuma.cu:1911-1941).  One iteration = black{close,far,refine} then
 * red{close,far,refine}; the three phases of a colour are fused into one launch. */
int gpm_sweep(gpm_ctx* ctx, int iterations);

/* One phase of one colour, for step-level parity tests: colour 0 = black, 1 = red;
 * phase_mask bit0 = close (gipuma.cu:1471-1588), bit1 = far (:1353-1468), bit2 = refine (:1590-1711). */
int gpm_phase(gpm_ctx* ctx, int colour, int phase_mask);

/* gipuma_compute_disp (gipuma.cu:1080-1103): normals to world frame, w <- depth (0 where cost == MAXCOST). */
int gpm_finalize(gpm_ctx* ctx);

/* Multi-view cost of caller-supplied planes at every pixel (no accept logic) — pmCostMultiview_cu
 * (gipuma.cu:720-806) through the iteration-time path; `planes`/`out_cost` like gpm_set_state. */
int gpm_cost_eval(gpm_ctx* ctx, const float* planes, float* out_cost, int on_device);

/* Whole job as runcuda() does it: init, params.iterations sweeps, finalize; results are left in the
 * context (gpm_get_state).  `sweep_ms` (may be NULL) receives the CUDA-event time from the first sweep
 * kernel to the end of the final kernel — the reference's own timed span (gipuma.cu:1908-1952). */
int gpm_run(gpm_ctx* ctx, float* sweep_ms);

/* ---- source-view sharding across GPUs (cost_comb = best_n only) --------------------------------
 * The reference is single-GPU (main.cpp:658-692); what is preserved is pmCostMultiview_cu's combination over ALL views
 * (gipuma.cu:742-806).  Each rank creates a context holding ALL of the state but only ITS subset of the source views
 * (gpm_set_view / gpm_set_num_views with the local views; same parameters, reference image and seed everywhere).
 *
 * High level — the whole runcuda() flow, exchange included, behind the C-ABI (a C++ host needs nothing else):
 *   gpm_shard_unique_id(id)           rank 0: 128-byte NCCL unique id; distribute it to the other ranks by any means
 *   gpm_shard_comm_init(ctx, id, r, n) every rank: ncclCommInitRank on the context's device (n = 1: no communicator)
 *   gpm_shard_comm_attach(ctx, comm, r, n)   alternative: use an existing ncclComm_t of the shard group
 *   gpm_shard_run(ctx, &ms)           init planes, initial costs, params.iterations sweeps, final kernel; per exchange
 *                                     stage one ncclAllGather of the local top-n_best lists on the context's stream; no
 *                                     host synchronisation in between.  Output bit-identical to gpm_run over all views.
 * NCCL is loaded at run time (dlopen "libnccl.so.2", override with GPM_NCCL_LIB); the library itself links only cudart.
 *
 * Low level (tests, custom transports): per stage, gpm_shard_stage applies the accept of the PREVIOUS stage from
 * `gathered_prev_dev` (rank-major concatenation of every rank's lists; ignored for stages 0 and 1) and writes this
 * rank's ascending n_best smallest per-view costs per pixel and hypothesis slot of THIS stage into `xchg_dev`
 * (gpm_shard_stage_floats(stage) floats, device memory).  Stages per colour: 1 (8 propagation candidates), then
 * 2 .. gpm_shard_num_stages()-1 (refinement steps, sequential), then gpm_shard_num_stages() (closing accept only, no
 * output).  Stage 0 evaluates the initial costs after gpm_init_planes (both colours at once; `colour` ignored) and is
 * closed by gpm_shard_finish_init. */
int gpm_init_planes(gpm_ctx* ctx);
int gpm_shard_num_stages(gpm_ctx* ctx);
long long gpm_shard_stage_floats(gpm_ctx* ctx, int stage);
int gpm_shard_stage(gpm_ctx* ctx, int colour, int stage, const float* gathered_prev_dev, int world, float* xchg_dev);
int gpm_shard_finish_init(gpm_ctx* ctx, const float* gathered_dev, int world);
int gpm_shard_unique_id(void* id128);
int gpm_shard_comm_init(gpm_ctx* ctx, const void* id128, int rank, int world);
int gpm_shard_comm_attach(gpm_ctx* ctx, void* nccl_comm, int rank, int world);
int gpm_shard_run(gpm_ctx* ctx, float* sweep_ms);
/* Fused compute + exchange over peer memory (NVLink / NVSwitch), what gpm_shard_run uses once regions are attached: every rank
 * exports one exchange region (CUDA IPC handle, 64 bytes; `local_ptr` for ranks of the same process), the handles of all
 * ranks are handed to gpm_shard_p2p_attach, and from then on one kernel per colour pass stores each pixel's lists directly
 * into the peers' regions while it samples and synchronises tile by tile with arrival flags — no collective launches.
 * Option "exchange" = 0 falls back to the NCCL all-gather flow. */
int gpm_shard_p2p_export(gpm_ctx* ctx, int world, void* handle64, void** local_ptr);
int gpm_shard_p2p_attach(gpm_ctx* ctx, const void* handles64, void* const* local_ptrs, int rank, int world);

/* Counters of the last gpm_sweep/gpm_run: [0] kernels launched, [1] hypotheses offered,
 * [2] hypotheses skipped as exact duplicates / out of depth range, [3] hypotheses cut short by the
 * exact lower bound, [4] (view,sample) evaluations done, [5] (view,sample) evaluations a full run would do,
 * [6] collectives issued by gpm_shard_run. */
int gpm_get_stats(gpm_ctx* ctx, unsigned long long stats[8]);
int gpm_reset_stats(gpm_ctx* ctx);

/* Measured ceiling of the unit that bounds this path: filtered R32F fetches per second (in 1e9) of the texture unit for
 * dense footprints on this context's own source-view texture (a ~30 ms microbenchmark; needs the views set).
 * bench.py's roofline.binding_unit divides the achieved fetch rate by it. */
int gpm_measure_fetch_peak(gpm_ctx* ctx, double* gfetch_per_s);

/* Diagnostics of the experimental "packed" sampling mode (option value 3 = sample both ways and compare): number of fetches
 * whose one-fetch gradient differed from the reference's four fetches and up to 64 records of 8 floats
 * {x, y, view, gx packed, gx reference, gy packed, gy reference, centre}. */
int gpm_debug_packed_mismatches(gpm_ctx* ctx, unsigned* count, float* records512, int reset);

/* Tuning / diagnostics; results are bit-identical for every setting.
 * "prune" (1): exact lower-bound early-out; "dedupe" (1): skip bit-identical candidate planes;
 * "trust_state" (0): treat a state loaded with gpm_set_state as cost-consistent; "nwarps" (0 = auto): warps per block;
 * "stats" (1): maintain the gpm_get_stats counters; "memo" (1): skip candidates / refinements this pixel is already
 * known to reject (exact); "cost_variant" (-1 = auto): which of the reference binary's rounding variants gpm_cost_eval
 * reproduces: bit 0 = x-term first (float: 1 in the propagation kernels, 0 at init / refinement; float4: 0 in
 * all six sweep kernels), bit 1 = float4 gradient folding (float4 initialisation = 3); auto = the propagation kernels' form;
 * "shard_async" (0): 1 makes gpm_shard_stage / gpm_shard_finish_init return after enqueueing on gpm_stream() — run the
 * collective on that stream (or order it with events) instead of paying a host synchronisation per stage;
 * "tma" (1): stage the reference window with one cp.async.bulk.tensor (TMA) per block instead of a cooperative copy;
 * "fused_warps" (16): warps per block of the fused shard kernel (8 = two resident blocks per SM: one samples while the other
 * waits for its peers' lists — measured no faster);
 * "exchange" (1): view shard over peer memory when regions are attached, 0 = NCCL all-gather per stage;
 * "async_upload" (0): 1 lets gpm_set_reference / gpm_set_view return without a host synchronisation — the caller keeps its
 * (page-locked) image buffers unchanged until the next gpm_run / gpm_sweep returns;
 * "prepass" (0; environment GPM_PREPASS=1 at gpm_create): a thread-per-pixel pre-pass of the sweep kernel lists the pixels
 * that still have work, so that converged pixels cost no warp;
 * "quadperm" (1): deal the samples of a round to the lanes as 2x2 blocks per hardware quad (texture-unit locality);
 * "neighbours" (8): 20 selects the reference's fused sweep — the kernels it launches when built without SMALLKERNEL
 * (gipuma.cu:1122-1351, 1913-1940): 12 axial + 8 knight-move neighbours, then refinement, one launch per colour; bit-exact
 * like the default (view sharding stays 8-neighbour only).
 * EXPERIMENTAL, not covered by the bit-exactness statement: "packed" (0 = off, 1 = auto, 2 = on) samples the source-view
 * gradients with one RG32F fetch instead of four R32F fetches for 8-bit images; measured 1 differing pixel in 1.92 M at cfg 2. */
int gpm_set_option(gpm_ctx* ctx, const char* name, int value);

/* The CUDA stream all work of this context is enqueued on (as a void*), for event timing by callers. */
void* gpm_stream(gpm_ctx* ctx);

/* ---- callers and data formats either side of the hot path (plain host C++, no CUDA) ------
 * gpm_prepare_cameras: n row-major 3x4 projection matrices (index 0 = reference) -> the Camera_cu field values, without
 *   OpenCV (cameraGeometryUtils.h:174-353: RQ decomposition, re-basing so that the reference is K[I|0], baseline 0.54).
 * gpm_select_views: deterministic selectViews (main.cpp:430-499); returns the number of views written to `subset`.
 * gpm_write_dmb / gpm_read_dmb / gpm_write_result_dmb: the .dmb depth / normal maps consumed by `fusibile`
 *   (fileIoUtils.h:247-368; scripts/dtu_fast.sh:57). */
int gpm_prepare_cameras(const double* P, int n, double cam_scale, gpm_camera* out);
int gpm_select_views(const gpm_camera* cams, int n, int cols, int rows, float min_angle, float max_angle, int max_views,
                     int* subset, float* depth_range);
int gpm_write_dmb(const char* path, const float* data, int rows, int cols, int channels);
int gpm_read_dmb(const char* path, float* data, size_t capacity_floats, int* rows, int* cols, int* channels);
int gpm_write_result_dmb(const char* depth_path, const char* normal_path, const float* norm4, int rows, int cols);

/* ---- reference-view batch driver: replaces the shell loop that starts one `gipuma` process per
 * reference image (scripts/dtu_fast.sh:30-55) and the view selection of main.cpp:430-499 ------------------------------
 * One process, a worker thread + gpm_ctx per listed device, ONE page-locked copy of the image set shared by all of them
 * (each image is registered once, uploaded asynchronously by whichever device needs it), reference views taken from a
 * common queue.  Per reference view: cameras re-based on it (gpm_prepare_cameras), source views selected
 * (gpm_select_views), depth range as main.cpp:480-483 / 898-906, gpm_run, result returned and / or written as
 * <out_dir>/<ref, 8 digits>/{disp,normals}.dmb (main.cpp:1002-1003) for the external fusibile. */
typedef struct gpm_batch_desc {
    int n_images, width, height;
    const float* const* images;        /* n_images host pointers, row-major float, pitch_bytes between rows (0 = width*4) */
    size_t pitch_bytes;
    const double* P;                   /* n_images row-major 3x4 projection matrices */
    double cam_scale;                  /* --cam_scale (K divided by it); <= 0 means 1 */
    gpm_params params;                 /* depthMin / depthMax <= 0: taken from the view selection; disparities are derived */
    float min_angle, max_angle;        /* degrees (scripts/dtu_fast.sh:18-19) */
    int max_views;                     /* scripts/dtu_fast.sh:21 */
    const int* ref_indices;            /* reference views to process (NULL: 0 .. n_refs-1) */
    int n_refs;
    const int* devices;                /* CUDA device ordinals, one worker each */
    int n_devices;
    unsigned long long seed;
    const char* out_dir;               /* NULL: no files */
    float* out_norm4;                  /* NULL or n_refs * height*width*4 floats (world normal + depth per job) */
    float* out_cost;                   /* NULL or n_refs * height*width floats */
} gpm_batch_desc;
typedef struct gpm_batch_stats {
    int jobs_done;
    double sweep_ms_total;             /* sum of the jobs' device times (reference's own span) */
    float* per_job_sweep_ms;           /* NULL or n_refs entries */
    int* per_job_views;                /* NULL or n_refs entries: source views selected */
    int* per_job_device;               /* NULL or n_refs entries: device that ran the job */
} gpm_batch_stats;
int gpm_batch_run(const gpm_batch_desc* desc, gpm_batch_stats* stats);
const char* gpm_batch_last_error(void);

const char* gpm_last_error(void);
const char* gpm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GIPUMA_B200_H */
