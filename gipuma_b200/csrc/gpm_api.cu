// gpm_api.cu — host side of libgipuma_b200.so: the C-ABI declared in include/gipuma_b200.h.
// Replaces the host launcher gipuma<T>() (reference gipuma.cu:1825-1960) and, through the adapter in
// runcuda_adapter.cu, int runcuda(GlobalState&) (gipuma.cu:1962-1970).
//
// Device memory layout (everything resident in HBM for the life of the context):
//   planes   float4[H*W]   (n.xyz, d) row-major, stride W   — LineState::norm4 layout (linestate.h:10)
//   cost     float [H*W]                                      — LineState::c
//   refpad   float [(H+32) x pitch]  reference image, replicate-padded by 16 px, pitch multiple of 32 floats
//   src      cudaArray (layered, R32F, W x H x max_views) + one texture object: Linear filter, the
//            reference's addressing (Wrap + unnormalised, main.cpp:642-648), element read mode
//   cams     ViewCam[max_views]  (K, R, t per source view) — copied to shared memory by every block
//   rng      uint32[H*W*6]  XORWOW state per pixel (GPM_RNG_STATEFUL only)
// No CPU fallback exists: every entry point needs a CUDA device and reports GPM_E_CUDA otherwise.
#include "../../include/gipuma_b200.h"
#include "gpm_kernels.cuh"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace gpm;

static thread_local std::string g_err;

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(GPM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
    } while (0)

struct gpm_ctx {
    int device = 0, W = 0, H = 0, maxV = 0, V = 0;
    int color = -1;                  // -1 undecided, 0 float images, 1 float4 (RGB) images — fixed by the first image upload
    gpm_params prm{};
    bool have_params = false, have_ref = false;
    std::vector<char> have_view;
    float4* planes = nullptr;
    float* cost = nullptr;
    unsigned* rng = nullptr;
    float* dispbuf = nullptr;        // view-shard mode: disp_now carried between the stages of one colour
    float4* candbuf = nullptr;       // view-shard mode: refinement candidate of the current step
    float* canddepth = nullptr;
    float4* seen = nullptr;          // [H*W*ncand] last plane offered to each pixel from each of the 8 (fused kernel: 20) propagation directions
    int seen_slots = 8;
    float4* refseen = nullptr;       // [H*W]   plane from which the last all-rejected refinement started
    unsigned* memo_mask = nullptr;   // [H*W] validity bits of seen (0-19) and refseen (GPM_MEMO_REFINE)
    unsigned char* prov = nullptr;   // per pixel: which rounding variant of the cost function produced cost[] (see k_sweep)
    float* refpad = nullptr;
    int refpitch = 0;
    float* staging = nullptr;        // W*H floats, upload scratch
    cudaArray_t srcArr = nullptr;
    cudaTextureObject_t srcTex = 0;
    cudaArray_t srcArr4 = nullptr;   // colour mode: layered R32F, 3 channel planes per view (layer 3v + ch: an RGBA32F fetch costs
                                     // 4.6x an R32F one on B200, profiles/r01_texbench.txt) + padded float4 reference + float4 staging
    float* planar = nullptr;         // [3][H][W] split target
    cudaTextureObject_t srcTex4 = 0;
    float4* refpad4 = nullptr;
    float4* staging4 = nullptr;
    cudaArray_t gradArr = nullptr;   // layered RG32F: (Gx, Gy) central differences of every source view (packed sampling mode)
    cudaTextureObject_t gradTex = 0;
    float2* gradLin = nullptr;       // W*H staging for one view's gradients
    int* d_flag = nullptr;
    std::vector<char> view_8bit;     // per view: every pixel an integer in [0,255]
    ViewCam* d_cams = nullptr;
    std::vector<ViewCam> h_cams;
    bool cams_dirty = true;
    RefCam ref{};
    unsigned long long seed = 0xC0FFEEULL;
    int rng_mode = GPM_RNG_REFERENCE;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned long long* d_stats = nullptr;
    unsigned long long launches = 0;
    int opt_prune = 1, opt_dedupe = 1, opt_trust_state = 0, opt_nwarps = 0, opt_stats = 1;
    int opt_cost_variant = -1, opt_packed = 0, opt_memo = 1, opt_quadperm = 1;
    int opt_shard_async = 0;                     // 1: gpm_shard_eval / gpm_shard_accept only enqueue on gpm_stream()
    int opt_neighbours = 8;                      // 20: the reference's fused kernel (built when SMALLKERNEL is not defined)
    int opt_site[21];                            // diagnostics: override the fused kernel's call-site variants (-1 = table)
    int smem_optin = 0, num_sms = 148;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int build_kparams(gpm_ctx* c, bool init_phase, KParams& P, bool eval_call = false)
{
    if (!c->have_params) return fail(GPM_E_STATE, "gpm_set_params has not been called");
    if (!c->have_ref) return fail(GPM_E_STATE, "gpm_set_reference has not been called");
    if (c->V < 1) return fail(GPM_E_STATE, "no source views (gpm_set_num_views)");
    for (int v = 0; v < c->V; v++)
        if (!c->have_view[v]) return fail(GPM_E_STATE, "source view " + std::to_string(v) + " has not been set");
    const gpm_params& p = c->prm;
    memset(&P, 0, sizeof(P));
    P.W = c->W;  P.H = c->H;  P.V = c->V;
    const int box = p.box_hsize;
    P.rad = init_phase ? box / 2 : (box - 1) / 2;          // gipuma.cu:1012 vs :1474
    P.nside = P.rad + 1;                                    // i = -rad, -rad+2, ... <= rad
    P.ns = P.nside * P.nside;
    P.ns_pad = (P.ns + 3) & ~3;
    P.halo = (box + 1) / 2;                                 // gipuma.cu:1844-1847
    P.tile_w = GPM_TILE + 2 * P.halo;
    // rounds of 32 consecutive samples (one per lane); the remainder forms a last, shorter round.  A window with
    // fewer than 48 samples (b <= 11) is split into two equal rounds instead.
    {
        int r = 0;
        if (P.ns > 32 && P.ns < 48) {
            P.round_end[r++] = (unsigned char)((P.ns + 1) / 2);
            P.round_end[r++] = (unsigned char)P.ns;
        } else {
            for (int e = 32; e < P.ns && r < 15; e += 32) P.round_end[r++] = (unsigned char)e;
            P.round_end[r++] = (unsigned char)P.ns;
        }
        P.nrounds = r;
    }
    // lane -> sample table of each round.  The texture unit filters a warp's fetches one hardware quad (lanes 4q..4q+3) at a
    // time and is fastest when a quad's four footprints form a compact 2x2 block (profiles/r02_texshape.txt): with
    // "quadperm" the samples of a round are dealt to the lanes as 2x2 blocks (x offset i, i+2; y offset j, j+2) where the
    // round contains them, leftovers in window order.  Only WHICH lane evaluates a sample changes; the per-view FMA chain
    // still consumes the dissimilarities in the reference's order.
    P.quadperm = c->opt_quadperm;
    {
        int s0 = 0;
        for (int r = 0; r < P.nrounds; r++) {
            const int s1 = P.round_end[r], len = s1 - s0;
            unsigned char* row = P.perm[r];
            for (int l = 0; l < 32; l++) row[l] = (unsigned char)(l < len ? l : 0);
            if (P.quadperm && len >= 4 && len <= 32) {
                bool used[32] = {false};
                int n = 0;
                for (int k = 0; k < len; k++) {
                    if (used[k]) continue;
                    const int s = s0 + k, jj = s % P.nside;
                    const int kr = k + P.nside, kd = k + 1, kx = k + P.nside + 1;
                    if (jj + 1 < P.nside && kx < len && !used[kr] && !used[kd] && !used[kx]) {
                        row[n++] = (unsigned char)k;  row[n++] = (unsigned char)kr;  row[n++] = (unsigned char)kd;  row[n++] = (unsigned char)kx;
                        used[k] = used[kr] = used[kd] = used[kx] = true;
                    }
                }
                for (int k = 0; k < len; k++) if (!used[k]) row[n++] = (unsigned char)k;
            }
            s0 = s1;
        }
    }
    P.refpitch = c->refpitch;
    P.tau_color = p.tau_color;  P.tau_gradient = p.tau_gradient;  P.alpha = p.alpha;  P.gamma = p.gamma;
    P.min_disp = p.min_disparity;  P.max_disp = p.max_disparity;
    P.n_best = p.n_best;  P.cost_comb = p.cost_comb;  P.good_factor = p.good_factor;
    P.prune = c->opt_prune;
    P.dedupe_self = c->opt_dedupe ? 1 : 0;
    // rounding variant of k_cost_eval: at initialisation the reference's binary is y-first for float, x-first for float4;
    // gpm_cost_eval defaults to the variant of the propagation kernels (x-first for float, y-first for float4)
    // For float4 the initialisation kernel additionally folds the other gradient term into the FMA (grad_variant).
    // Option cost_variant: bit 0 = x-term first, bit 1 = gradient folding (float4 only).
    const int cv = eval_call ? c->opt_cost_variant : -1;          // the option only steers gpm_cost_eval
    P.cost_rt = 0;
    if (init_phase) { P.cost_variant = c->color == 1 ? 1 : 0;  P.grad_variant = c->color == 1 ? 1 : 0; }
    else if (cv >= 0) { P.cost_variant = cv & 1;  P.grad_variant = (c->color == 1) ? ((cv >> 1) & 1) : 0;  P.cost_rt = c->color == 1 ? cv : (cv & ~2); }
    else { P.cost_variant = c->color == 1 ? 0 : 1;  P.grad_variant = 0; }
    P.ncand = c->opt_neighbours == 20 ? 20 : 8;
    // call-site variants of the fused kernels, read off the reference build (tools/fused_probe.py)
    {
        static const unsigned char kSitesFloat[21] = GPM_FUSED_SITES_FLOAT, kSitesFloat4[21] = GPM_FUSED_SITES_FLOAT4;
        for (int k = 0; k < 21; k++)
            P.site[k] = c->opt_site[k] >= 0 ? (unsigned char)c->opt_site[k] : (c->color == 1 ? kSitesFloat4[k] : kSitesFloat[k]);
    }
    P.color = c->color == 1 ? 1 : 0;
    P.memo = c->opt_memo;
    P.packed = c->opt_packed;
    for (int v = 0; v < c->V; v++) if (!c->view_8bit[v]) P.packed = 0;
    P.dedupe_cand = c->opt_dedupe ? 1 : 0;
    P.rng_mode = c->rng_mode;
    P.ref = c->ref;
    P.ref.depthMin = p.depthMin;  P.ref.depthMax = p.depthMax;
    // warps per block: as many as fit (<= 16).  The kernels use 128 registers per thread, so one 16-warp block fills an SM
    // anyway; shared memory may therefore be spent up to the per-block opt-in limit.
    const size_t per_warp = (size_t)warp_scratch_floats(P.ns_pad, P.V, P.color) * sizeof(float);
    const size_t fixed = ((size_t)fixed_smem_floats(P) + 4) * sizeof(float);
    int nw = GPM_LB_THREADS / 32;
    const size_t budget = (size_t)c->smem_optin > 16 * 1024 ? (size_t)c->smem_optin - 8 * 1024 : 40 * 1024;
    while (nw > 2 && fixed + nw * per_warp > budget) nw--;
    if (c->opt_nwarps > 0) nw = c->opt_nwarps;
    if (nw > GPM_LB_THREADS / 32) nw = GPM_LB_THREADS / 32;
    P.nwarps = nw;
    if (P.color) P.packed = 0;
    if (P.packed && c->opt_packed == 1) {
        // auto: the packed mode triples the texture bytes a warp touches per pixel (4 -> 12 B per texel and view); it only
        // pays while the block's working set stays near the L1 (measured: +11 % at cfg 2, -17 % at cfg 3)
        const size_t ws = (size_t)P.V * (box + 2) * (box + 2) * 12 * nw;
        if (ws > 640 * 1024) P.packed = 0;
    }
    if (block_smem_bytes(P) > (size_t)c->smem_optin)
        return fail(GPM_E_ARG, "configuration needs more shared memory per block than the device offers");
    return GPM_OK;
}

int sync_cams(gpm_ctx* c)
{
    if (!c->cams_dirty) return GPM_OK;
    CU(cudaMemcpyAsync(c->d_cams, c->h_cams.data(), sizeof(ViewCam) * c->maxV, cudaMemcpyHostToDevice, c->stream));
    c->cams_dirty = false;
    return GPM_OK;
}

int launch_colour(gpm_ctx* c, const KParams& P, int colour, int mask)
{
    const size_t smem = block_smem_bytes(P);
    dim3 grid((P.W + GPM_TILE - 1) / GPM_TILE, (P.H + GPM_TILE - 1) / GPM_TILE);
    // at least ~8 waves of blocks over the SMs: split each tile's pixel list into up to 8 slices for small images
    int split = 1;
    while (split < 8 && (long long)grid.x * grid.y * split < 8LL * c->num_sms) split *= 2;
    grid.z = split;
    auto kern = P.ncand == 20 ? (P.color ? k_sweep<false, true, true> : k_sweep<false, false, true>)
                              : (P.color ? k_sweep<false, true, false> : (P.packed ? k_sweep<true, false, false> : k_sweep<false, false, false>));
    kern<<<grid, P.nwarps * 32, smem, c->stream>>>(P, c->d_cams, P.color ? (const float*)c->refpad4 : c->refpad, P.color ? c->srcTex4 : c->srcTex, c->gradTex, c->planes, c->cost, c->rng,
                                                      c->prov, c->seen, c->refseen, c->memo_mask, colour, mask, c->opt_stats ? c->d_stats : nullptr);
    c->launches++;
    CU(cudaGetLastError());
    return GPM_OK;
}

int upload_image(gpm_ctx* c, const float* img, size_t pitch_bytes, int on_device, float** dev_img, size_t* dev_pitch_floats)
{
    if (pitch_bytes == 0) pitch_bytes = (size_t)c->W * sizeof(float);
    if (pitch_bytes % sizeof(float)) return fail(GPM_E_ARG, "pitch_bytes must be a multiple of 4");
    if (on_device) {
        *dev_img = const_cast<float*>(img);
        *dev_pitch_floats = pitch_bytes / sizeof(float);
        return GPM_OK;
    }
    CU(cudaMemcpy2DAsync(c->staging, (size_t)c->W * sizeof(float), img, pitch_bytes, (size_t)c->W * sizeof(float), c->H,
                         cudaMemcpyHostToDevice, c->stream));
    *dev_img = c->staging;
    *dev_pitch_floats = c->W;
    return GPM_OK;
}

}  // namespace

static int ensure_color(gpm_ctx* c, int want);

static void set_ref_camera(gpm_ctx* c, const gpm_camera* cam)
{
    RefCam& r = c->ref;
    memcpy(r.K_inv, cam->K_inv, sizeof(r.K_inv));
    memcpy(r.M_inv, cam->M_inv, sizeof(r.M_inv));
    memcpy(r.R_orig_inv, cam->R_orig_inv, sizeof(r.R_orig_inv));
    memcpy(r.P34, cam->P_col34, sizeof(r.P34));
    memcpy(r.C, cam->C, sizeof(r.C));
    r.fx = cam->fx;  r.alpha = cam->alpha;  r.K2 = cam->K[2];  r.K5 = cam->K[5];
    r.f = cam->f;  r.f_cam = cam->f;  r.baseline = cam->baseline;
    c->have_ref = true;
}

extern "C" const char* gpm_last_error(void) { return g_err.c_str(); }
extern "C" const char* gpm_version(void) { return "gipuma_b200 0.1 (sm_100a)"; }

extern "C" int gpm_create(gpm_ctx** out, int device, int width, int height, int max_views)
{
    if (!out || width < 8 || height < 8 || max_views < 1 || max_views > GPM_MAX_VIEWS)
        return fail(GPM_E_ARG, "gpm_create: bad arguments (max_views must be 1.." + std::to_string(GPM_MAX_VIEWS) + ")");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1)
        return fail(GPM_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (gipuma_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(GPM_E_ARG, "gpm_create: no such device");
    DeviceGuard g(device);
    gpm_ctx* c = new gpm_ctx;
    for (int& v : c->opt_site) v = -1;
    c->device = device;  c->W = width;  c->H = height;  c->maxV = max_views;
    c->have_view.assign(max_views, 0);
    c->view_8bit.assign(max_views, 0);
    c->h_cams.assign(max_views, ViewCam{});
    const size_t n = (size_t)width * height;
    c->refpitch = (width + 2 * GPM_APRON + 31) & ~31;
    cudaError_t err = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (err == cudaSuccess && r != cudaSuccess) err = r; return r == cudaSuccess; };
    ok(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    ok(cudaEventCreate(&c->ev0));
    ok(cudaEventCreate(&c->ev1));
    ok(cudaMalloc(&c->planes, n * sizeof(float4)));
    ok(cudaMalloc(&c->cost, n * sizeof(float)));
    ok(cudaMalloc(&c->prov, n));
    ok(cudaMalloc(&c->seen, n * 8 * sizeof(float4)));
    ok(cudaMalloc(&c->refseen, n * sizeof(float4)));
    ok(cudaMalloc(&c->memo_mask, n * sizeof(unsigned)));
    ok(cudaMalloc(&c->staging, n * sizeof(float)));
    ok(cudaMalloc(&c->d_flag, sizeof(int)));
    ok(cudaMalloc(&c->refpad, (size_t)c->refpitch * (height + 2 * GPM_APRON) * sizeof(float)));
    ok(cudaMalloc(&c->d_cams, sizeof(ViewCam) * max_views));
    ok(cudaMalloc(&c->d_stats, 8 * sizeof(unsigned long long)));
    if (err == cudaSuccess) {
        ok(cudaMemsetAsync(c->planes, 0, n * sizeof(float4), c->stream));      // LineState::resize zeroes (linestate.h:19-24)
        ok(cudaMemsetAsync(c->cost, 0, n * sizeof(float), c->stream));
        ok(cudaMemsetAsync(c->prov, GPM_PROV_UNKNOWN, n, c->stream));
        ok(cudaMemsetAsync(c->memo_mask, 0, n * sizeof(unsigned), c->stream));
        ok(cudaMemsetAsync(c->d_stats, 0, 8 * sizeof(unsigned long long), c->stream));
        cudaChannelFormatDesc desc = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
        ok(cudaMalloc3DArray(&c->srcArr, &desc, make_cudaExtent(width, height, max_views), cudaArrayLayered));
    }
    if (err == cudaSuccess) {
        cudaResourceDesc res;  memset(&res, 0, sizeof(res));
        res.resType = cudaResourceTypeArray;  res.res.array.array = c->srcArr;
        cudaTextureDesc td;  memset(&td, 0, sizeof(td));
        td.addressMode[0] = cudaAddressModeWrap;  td.addressMode[1] = cudaAddressModeWrap;   // main.cpp:644-645
        td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;  td.readMode = cudaReadModeElementType;  td.normalizedCoords = 0;
        ok(cudaCreateTextureObject(&c->srcTex, &res, &td, NULL));
        ok(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        ok(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
        ok(cudaFuncSetAttribute(k_sweep<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_sweep<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_sweep<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_sweep<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_sweep<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_cost_eval<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_cost_eval<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_cost_eval<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_shard_eval<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_shard_eval<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_shard_eval<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
    }
    if (err != cudaSuccess) {
        std::string m = std::string("gpm_create: ") + cudaGetErrorString(err);
        gpm_destroy(c);
        return fail(GPM_E_CUDA, m);
    }
    *out = c;
    return GPM_OK;
}

extern "C" void gpm_destroy(gpm_ctx* c)
{
    if (!c) return;
    DeviceGuard g(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->srcTex) cudaDestroyTextureObject(c->srcTex);
    if (c->srcArr) cudaFreeArray(c->srcArr);
    if (c->srcTex4) cudaDestroyTextureObject(c->srcTex4);
    if (c->srcArr4) cudaFreeArray(c->srcArr4);
    cudaFree(c->refpad4);  cudaFree(c->staging4);  cudaFree(c->planar);
    if (c->gradTex) cudaDestroyTextureObject(c->gradTex);
    if (c->gradArr) cudaFreeArray(c->gradArr);
    cudaFree(c->gradLin);  cudaFree(c->d_flag);
    cudaFree(c->planes);  cudaFree(c->cost);  cudaFree(c->prov);  cudaFree(c->seen);  cudaFree(c->refseen);  cudaFree(c->memo_mask);  cudaFree(c->rng);  cudaFree(c->dispbuf);  cudaFree(c->candbuf);  cudaFree(c->canddepth);  cudaFree(c->refpad);  cudaFree(c->staging);
    cudaFree(c->d_cams);  cudaFree(c->d_stats);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int gpm_set_params(gpm_ctx* c, const gpm_params* p)
{
    if (!c || !p) return fail(GPM_E_ARG, "gpm_set_params: null argument");
    if (p->box_hsize != p->box_vsize)
        return fail(GPM_E_ARG, "box_hsize != box_vsize: the reference's tile loader (gipuma.cu:1517) is only defined for square windows");
    if (p->box_hsize < 3 || p->box_hsize > GPM_MAX_BOX || !(p->box_hsize & 1))
        return fail(GPM_E_ARG, "box size must be odd, 3.." + std::to_string(GPM_MAX_BOX));
    if (p->cost_comb < 0 || p->cost_comb > 3) return fail(GPM_E_ARG, "cost_comb must be 0..3");
    if (p->iterations < 0) return fail(GPM_E_ARG, "iterations must be >= 0");
    c->prm = *p;
    c->have_params = true;
    {
        DeviceGuard g(c->device);
        CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));   // memo depends on the parameters
    }
    return GPM_OK;
}

extern "C" int gpm_set_num_views(gpm_ctx* c, int n)
{
    if (!c || n < 1 || n > c->maxV) return fail(GPM_E_ARG, "gpm_set_num_views: out of range");
    c->V = n;
    return GPM_OK;
}

extern "C" int gpm_set_rng(gpm_ctx* c, unsigned long long seed, int mode)
{
    if (!c || (mode != GPM_RNG_REFERENCE && mode != GPM_RNG_STATEFUL)) return fail(GPM_E_ARG, "gpm_set_rng: bad mode");
    DeviceGuard g(c->device);
    c->seed = seed;
    c->rng_mode = mode;
    CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
    if (mode == GPM_RNG_STATEFUL && !c->rng) {
        CU(cudaMalloc(&c->rng, (size_t)c->W * c->H * 6 * sizeof(unsigned)));
        CU(cudaMemsetAsync(c->rng, 0, (size_t)c->W * c->H * 6 * sizeof(unsigned), c->stream));
    }
    return GPM_OK;
}

extern "C" int gpm_set_reference(gpm_ctx* c, const float* img, size_t pitch_bytes, int on_device, const gpm_camera* cam)
{
    if (!c || !img || !cam) return fail(GPM_E_ARG, "gpm_set_reference: null argument");
    DeviceGuard g(c->device);
    float* d = nullptr;
    size_t pf = 0;
    int rc = ensure_color(c, 0);
    if (rc) return rc;
    rc = upload_image(c, img, pitch_bytes, on_device, &d, &pf);
    if (rc) return rc;
    dim3 b(32, 8), gr((c->W + 2 * GPM_APRON + 31) / 32, (c->H + 2 * GPM_APRON + 7) / 8);
    k_pad_reference<<<gr, b, 0, c->stream>>>(d, pf, c->W, c->H, c->refpad, c->refpitch);
    CU(cudaGetLastError());
    set_ref_camera(c, cam);
    CU(cudaMemsetAsync(c->prov, GPM_PROV_UNKNOWN, (size_t)c->W * c->H, c->stream));
    CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
    CU(cudaStreamSynchronize(c->stream));     // staging buffer is reused by the next upload
    return GPM_OK;
}

extern "C" int gpm_set_view(gpm_ctx* c, int v, const float* img, size_t pitch_bytes, int on_device, const gpm_camera* cam)
{
    if (!c || !img || !cam) return fail(GPM_E_ARG, "gpm_set_view: null argument");
    if (v < 0 || v >= c->maxV) return fail(GPM_E_ARG, "gpm_set_view: view index out of range");
    DeviceGuard g(c->device);
    float* d = nullptr;
    size_t pf = 0;
    int rc = ensure_color(c, 0);
    if (rc) return rc;
    rc = upload_image(c, img, pitch_bytes, on_device, &d, &pf);          // host images go through the staging buffer
    if (rc) return rc;
    cudaMemcpy3DParms m;  memset(&m, 0, sizeof(m));
    m.srcPtr = make_cudaPitchedPtr(d, pf * sizeof(float), c->W, c->H);
    m.dstArray = c->srcArr;
    m.dstPos = make_cudaPos(0, 0, v);
    m.extent = make_cudaExtent(c->W, c->H, 1);
    m.kind = cudaMemcpyDeviceToDevice;
    CU(cudaMemcpy3DAsync(&m, c->stream));
    c->view_8bit[v] = 0;
    if (c->opt_packed) {                         // experimental packed sampling mode: central-difference planes + "8-bit valued" test
    if (!c->gradArr) {
        cudaChannelFormatDesc desc2 = cudaCreateChannelDesc(32, 32, 0, 0, cudaChannelFormatKindFloat);
        CU(cudaMalloc3DArray(&c->gradArr, &desc2, make_cudaExtent(c->W, c->H, c->maxV), cudaArrayLayered));
        cudaResourceDesc res;  memset(&res, 0, sizeof(res));
        res.resType = cudaResourceTypeArray;  res.res.array.array = c->gradArr;
        cudaTextureDesc td;  memset(&td, 0, sizeof(td));
        td.addressMode[0] = cudaAddressModeWrap;  td.addressMode[1] = cudaAddressModeWrap;  td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;  td.readMode = cudaReadModeElementType;  td.normalizedCoords = 0;
        CU(cudaCreateTextureObject(&c->gradTex, &res, &td, NULL));
        CU(cudaMalloc(&c->gradLin, (size_t)c->W * c->H * sizeof(float2)));
    }
    const int one = 1;
    CU(cudaMemcpyAsync(c->d_flag, &one, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    dim3 b(32, 8), gr((c->W + 31) / 32, (c->H + 7) / 8);
    k_make_gradients<<<gr, b, 0, c->stream>>>(d, pf, c->W, c->H, c->gradLin, c->d_flag);
    CU(cudaGetLastError());
    m.srcPtr = make_cudaPitchedPtr(c->gradLin, (size_t)c->W * sizeof(float2), c->W, c->H);
    m.dstArray = c->gradArr;
    CU(cudaMemcpy3DAsync(&m, c->stream));
    int flag = 0;
    CU(cudaMemcpyAsync(&flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->view_8bit[v] = flag ? 1 : 0;
    }
    ViewCam& vc = c->h_cams[v];
    memcpy(vc.K, cam->K, sizeof(vc.K));
    memcpy(vc.R, cam->R, sizeof(vc.R));
    memcpy(vc.t, cam->t, sizeof(vc.t));
    c->cams_dirty = true;
    c->have_view[v] = 1;
    CU(cudaMemsetAsync(c->prov, GPM_PROV_UNKNOWN, (size_t)c->W * c->H, c->stream));
    CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
    CU(cudaStreamSynchronize(c->stream));        // staging buffers are reused; the caller may reuse its buffer
    return GPM_OK;
}


// ---- colour (float4) images: the reference's -color_processing path (T = float4, main.cpp:560-605) -------------
static int ensure_color(gpm_ctx* c, int want)
{
    if (c->color == -1) c->color = want;
    if (c->color != want) return fail(GPM_E_STATE, "a context holds either float or float4 images, not both");
    if (want == 1 && !c->srcArr4) {
        cudaChannelFormatDesc d1 = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
        CU(cudaMalloc3DArray(&c->srcArr4, &d1, make_cudaExtent(c->W, c->H, 3 * (size_t)c->maxV), cudaArrayLayered));
        CU(cudaMalloc(&c->planar, 3 * (size_t)c->W * c->H * sizeof(float)));
        cudaResourceDesc res;  memset(&res, 0, sizeof(res));
        res.resType = cudaResourceTypeArray;  res.res.array.array = c->srcArr4;
        cudaTextureDesc td;  memset(&td, 0, sizeof(td));
        td.addressMode[0] = cudaAddressModeWrap;  td.addressMode[1] = cudaAddressModeWrap;  td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;  td.readMode = cudaReadModeElementType;  td.normalizedCoords = 0;
        CU(cudaCreateTextureObject(&c->srcTex4, &res, &td, NULL));
        CU(cudaMalloc(&c->refpad4, (size_t)c->refpitch * (c->H + 2 * GPM_APRON) * sizeof(float4)));
        CU(cudaMalloc(&c->staging4, (size_t)c->W * c->H * sizeof(float4)));
    }
    return GPM_OK;
}

static int upload_image4(gpm_ctx* c, const float* img, size_t pitch_bytes, int on_device, float4** dev_img, size_t* dev_pitch_elems)
{
    if (pitch_bytes == 0) pitch_bytes = (size_t)c->W * sizeof(float4);
    if (pitch_bytes % sizeof(float4)) return fail(GPM_E_ARG, "pitch_bytes must be a multiple of 16 for float4 images");
    if (on_device) { *dev_img = (float4*)img;  *dev_pitch_elems = pitch_bytes / sizeof(float4);  return GPM_OK; }
    CU(cudaMemcpy2DAsync(c->staging4, (size_t)c->W * sizeof(float4), img, pitch_bytes, (size_t)c->W * sizeof(float4), c->H,
                         cudaMemcpyHostToDevice, c->stream));
    *dev_img = c->staging4;
    *dev_pitch_elems = c->W;
    return GPM_OK;
}

extern "C" int gpm_set_reference_color(gpm_ctx* c, const float* rgba, size_t pitch_bytes, int on_device, const gpm_camera* cam)
{
    if (!c || !rgba || !cam) return fail(GPM_E_ARG, "gpm_set_reference_color: null argument");
    DeviceGuard g(c->device);
    int rc = ensure_color(c, 1);
    if (rc) return rc;
    float4* d = nullptr;
    size_t pe = 0;
    rc = upload_image4(c, rgba, pitch_bytes, on_device, &d, &pe);
    if (rc) return rc;
    dim3 b(32, 8), gr((c->W + 2 * GPM_APRON + 31) / 32, (c->H + 2 * GPM_APRON + 7) / 8);
    k_pad_reference4<<<gr, b, 0, c->stream>>>(d, pe, c->W, c->H, c->refpad4, c->refpitch);
    CU(cudaGetLastError());
    set_ref_camera(c, cam);
    CU(cudaMemsetAsync(c->prov, GPM_PROV_UNKNOWN, (size_t)c->W * c->H, c->stream));
    CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

extern "C" int gpm_set_view_color(gpm_ctx* c, int v, const float* rgba, size_t pitch_bytes, int on_device, const gpm_camera* cam)
{
    if (!c || !rgba || !cam) return fail(GPM_E_ARG, "gpm_set_view_color: null argument");
    if (v < 0 || v >= c->maxV) return fail(GPM_E_ARG, "gpm_set_view_color: view index out of range");
    DeviceGuard g(c->device);
    int rc = ensure_color(c, 1);
    if (rc) return rc;
    float4* d = nullptr;
    size_t pe = 0;
    rc = upload_image4(c, rgba, pitch_bytes, on_device, &d, &pe);
    if (rc) return rc;
    {
        dim3 b(32, 8), gr((c->W + 31) / 32, (c->H + 7) / 8);
        k_split_channels<<<gr, b, 0, c->stream>>>(d, pe, c->W, c->H, c->planar);
        CU(cudaGetLastError());
    }
    cudaMemcpy3DParms m;  memset(&m, 0, sizeof(m));
    m.srcPtr = make_cudaPitchedPtr(c->planar, c->W * sizeof(float), c->W, c->H);
    m.dstArray = c->srcArr4;
    m.dstPos = make_cudaPos(0, 0, 3 * (size_t)v);
    m.extent = make_cudaExtent(c->W, c->H, 3);
    m.kind = cudaMemcpyDeviceToDevice;
    CU(cudaMemcpy3DAsync(&m, c->stream));
    ViewCam& vc = c->h_cams[v];
    memcpy(vc.K, cam->K, sizeof(vc.K));
    memcpy(vc.R, cam->R, sizeof(vc.R));
    memcpy(vc.t, cam->t, sizeof(vc.t));
    c->cams_dirty = true;
    c->have_view[v] = 1;
    c->view_8bit[v] = 0;
    CU(cudaMemsetAsync(c->prov, GPM_PROV_UNKNOWN, (size_t)c->W * c->H, c->stream));
    CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

extern "C" int gpm_set_state(gpm_ctx* c, const float* norm4, const float* cost, int on_device)
{
    if (!c) return fail(GPM_E_ARG, "gpm_set_state: null context");
    DeviceGuard g(c->device);
    const size_t n = (size_t)c->W * c->H;
    const cudaMemcpyKind k = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (norm4) CU(cudaMemcpyAsync(c->planes, norm4, n * sizeof(float4), k, c->stream));
    if (cost) CU(cudaMemcpyAsync(c->cost, cost, n * sizeof(float), k, c->stream));
    // provenance of the supplied costs is unknown (GPM_PROV_UNKNOWN) unless the caller vouches that they came from an
    // initialisation / refinement evaluation of exactly these planes ("trust_state": 0)
    CU(cudaMemsetAsync(c->prov, c->opt_trust_state ? (c->color == 1 ? 3 : 0) : GPM_PROV_UNKNOWN, n, c->stream));
    CU(cudaMemsetAsync(c->memo_mask, 0, n * sizeof(unsigned), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

extern "C" int gpm_get_state(gpm_ctx* c, float* norm4, float* cost, int on_device)
{
    if (!c) return fail(GPM_E_ARG, "gpm_get_state: null context");
    DeviceGuard g(c->device);
    const size_t n = (size_t)c->W * c->H;
    const cudaMemcpyKind k = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (norm4) CU(cudaMemcpyAsync(norm4, c->planes, n * sizeof(float4), k, c->stream));
    if (cost) CU(cudaMemcpyAsync(cost, c->cost, n * sizeof(float), k, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

static int do_init(gpm_ctx* c)
{
    KParams P;
    int rc = build_kparams(c, true, P);
    if (rc) return rc;
    rc = sync_cams(c);
    if (rc) return rc;
    dim3 b(16, 16), gr((c->W + 15) / 16, (c->H + 15) / 16);                   // gipuma.cu:1870-1875
    k_init_planes<<<gr, b, 0, c->stream>>>(P, c->seed, c->planes, c->rng_mode == GPM_RNG_STATEFUL ? c->rng : nullptr);
    c->launches++;
    CU(cudaGetLastError());
    dim3 grid((P.W + GPM_TILE - 1) / GPM_TILE, (P.H + GPM_TILE - 1) / GPM_TILE);
    (P.color ? k_cost_eval<false, true> : (P.packed ? k_cost_eval<true, false> : k_cost_eval<false, false>))<<<grid, P.nwarps * 32, block_smem_bytes(P), c->stream>>>(P, c->d_cams, P.color ? (const float*)c->refpad4 : c->refpad, P.color ? c->srcTex4 : c->srcTex, c->gradTex, c->planes,
                                                                         c->cost, nullptr);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(c->prov, c->color == 1 ? 3 : 0, (size_t)c->W * c->H, c->stream));   // costs now come from the init-variant evaluation (float4: x-first + gradient folding 1)
    CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
    return GPM_OK;
}

static int do_sweeps(gpm_ctx* c, int iterations)
{
    KParams P;
    int rc = build_kparams(c, false, P);
    if (rc) return rc;
    rc = sync_cams(c);
    if (rc) return rc;
    for (int it = 0; it < iterations; it++) {                                  // gipuma.cu:1911-1941
        rc = launch_colour(c, P, 0, 7);
        if (rc) return rc;
        rc = launch_colour(c, P, 1, 7);
        if (rc) return rc;
    }
    return GPM_OK;
}

static int do_finalize(gpm_ctx* c)
{
    KParams P;
    int rc = build_kparams(c, false, P);
    if (rc) return rc;
    dim3 b(16, 16), gr((c->W + 15) / 16, (c->H + 15) / 16);
    k_finalize<<<gr, b, 0, c->stream>>>(P, c->planes, c->cost);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(c->prov, GPM_PROV_UNKNOWN, (size_t)c->W * c->H, c->stream));
    CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));   // planes are world-frame outputs now
    return GPM_OK;
}

extern "C" int gpm_init(gpm_ctx* c)
{
    if (!c) return fail(GPM_E_ARG, "null context");
    DeviceGuard g(c->device);
    int rc = do_init(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

extern "C" int gpm_sweep(gpm_ctx* c, int iterations)
{
    if (!c || iterations < 0) return fail(GPM_E_ARG, "gpm_sweep: bad arguments");
    DeviceGuard g(c->device);
    int rc = do_sweeps(c, iterations);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

extern "C" int gpm_phase(gpm_ctx* c, int colour, int phase_mask)
{
    if (!c || colour < 0 || colour > 1 || phase_mask < 1 || phase_mask > 7) return fail(GPM_E_ARG, "gpm_phase: bad arguments");
    DeviceGuard g(c->device);
    KParams P;
    int rc = build_kparams(c, false, P);
    if (rc) return rc;
    rc = sync_cams(c);
    if (rc) return rc;
    rc = launch_colour(c, P, colour, phase_mask);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

extern "C" int gpm_finalize(gpm_ctx* c)
{
    if (!c) return fail(GPM_E_ARG, "null context");
    DeviceGuard g(c->device);
    int rc = do_finalize(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

extern "C" int gpm_cost_eval(gpm_ctx* c, const float* planes, float* out_cost, int on_device)
{
    if (!c || !planes || !out_cost) return fail(GPM_E_ARG, "gpm_cost_eval: null argument");
    DeviceGuard g(c->device);
    KParams P;
    int rc = build_kparams(c, false, P, true);
    if (rc) return rc;
    rc = sync_cams(c);
    if (rc) return rc;
    const size_t n = (size_t)c->W * c->H;
    float4* d_pl = nullptr;
    float* d_out = nullptr;
    if (on_device) { d_pl = (float4*)planes;  d_out = out_cost; }
    else {
        CU(cudaMalloc(&d_pl, n * sizeof(float4)));
        CU(cudaMalloc(&d_out, n * sizeof(float)));
        CU(cudaMemcpyAsync(d_pl, planes, n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    }
    dim3 grid((P.W + GPM_TILE - 1) / GPM_TILE, (P.H + GPM_TILE - 1) / GPM_TILE);
    (P.color ? k_cost_eval<false, true> : (P.packed ? k_cost_eval<true, false> : k_cost_eval<false, false>))<<<grid, P.nwarps * 32, block_smem_bytes(P), c->stream>>>(P, c->d_cams, P.color ? (const float*)c->refpad4 : c->refpad, P.color ? c->srcTex4 : c->srcTex, c->gradTex, d_pl, d_out, nullptr);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && !on_device) e = cudaMemcpyAsync(out_cost, d_out, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (!on_device) { cudaFree(d_pl);  cudaFree(d_out); }
    if (e != cudaSuccess) return fail(GPM_E_CUDA, std::string("gpm_cost_eval: ") + cudaGetErrorString(e));
    return GPM_OK;
}

extern "C" int gpm_run(gpm_ctx* c, float* sweep_ms)
{
    if (!c) return fail(GPM_E_ARG, "null context");
    DeviceGuard g(c->device);
    CU(cudaMemsetAsync(c->d_stats, 0, 8 * sizeof(unsigned long long), c->stream));
    c->launches = 0;
    int rc = do_init(c);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev0, c->stream));                                    // gipuma.cu:1908
    rc = do_sweeps(c, c->prm.iterations);
    if (rc) return rc;
    rc = do_finalize(c);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev1, c->stream));                                    // gipuma.cu:1946
    CU(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (sweep_ms) *sweep_ms = ms;
    CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}


// ---- source-view sharding (multi-GPU) -------------------------------------------------------------------------
static int shard_refine_steps(const gpm_params& p)
{
    int n = 0;
    for (float dz = p.max_disparity * 0.5f; dz >= 0.01f; dz = dz * 0.1f) n++;      // gipuma.cu:958-959
    return n;
}

extern "C" int gpm_shard_num_stages(gpm_ctx* c)
{
    if (!c || !c->have_params) { fail(GPM_E_STATE, "gpm_shard_num_stages: parameters not set");  return -1; }
    return 2 + shard_refine_steps(c->prm);
}

extern "C" long long gpm_shard_stage_floats(gpm_ctx* c, int stage)
{
    if (!c || !c->have_params || stage < 0) { fail(GPM_E_ARG, "gpm_shard_stage_floats: bad arguments");  return -1; }
    const long long half = (long long)c->H * ((c->W + 1) / 2);
    const int slots = stage == 1 ? 8 : 1;
    return (stage == 0 ? 2 : 1) * half * slots * c->prm.n_best;
}

static int shard_common(gpm_ctx* c, int stage, KParams& P)
{
    if (c->prm.cost_comb != GPM_COMB_BEST_N) return fail(GPM_E_ARG, "view sharding supports cost_comb = best_n only");
    if (c->opt_neighbours != 8) return fail(GPM_E_ARG, "view sharding supports the 8-neighbour sweep only (option neighbours = 8)");
    if (c->prm.n_best < 1 || c->prm.n_best > 32) return fail(GPM_E_ARG, "view sharding needs 1 <= n_best <= 32");
    if (stage < 0 || stage >= 2 + shard_refine_steps(c->prm)) return fail(GPM_E_ARG, "no such stage");
    int rc = build_kparams(c, stage == 0, P);
    if (rc) return rc;
    rc = sync_cams(c);
    if (rc) return rc;
    if (!c->dispbuf) {
        const size_t n = (size_t)c->W * c->H;
        CU(cudaMalloc(&c->dispbuf, n * sizeof(float)));
        CU(cudaMalloc(&c->candbuf, n * sizeof(float4)));
        CU(cudaMalloc(&c->canddepth, n * sizeof(float)));
    }
    return GPM_OK;
}

extern "C" int gpm_shard_eval(gpm_ctx* c, int colour, int stage, float* xchg_dev)
{
    if (!c || !xchg_dev || colour < 0 || colour > 1) return fail(GPM_E_ARG, "gpm_shard_eval: bad arguments");
    DeviceGuard g(c->device);
    KParams P;
    int rc = shard_common(c, stage, P);
    if (rc) return rc;
    dim3 grid((P.W + GPM_TILE - 1) / GPM_TILE, (P.H + GPM_TILE - 1) / GPM_TILE);
    int split = 1;                                        // at least ~8 waves of blocks, as in launch_colour
    while (split < 8 && (long long)grid.x * grid.y * split < 8LL * c->num_sms) split *= 2;
    grid.z = split;
    (P.color ? k_shard_eval<false, true> : (P.packed ? k_shard_eval<true, false> : k_shard_eval<false, false>))<<<grid, P.nwarps * 32, block_smem_bytes(P), c->stream>>>(P, c->d_cams, P.color ? (const float*)c->refpad4 : c->refpad, P.color ? c->srcTex4 : c->srcTex, c->gradTex, c->planes, c->cost,
                                                                          c->prov, c->dispbuf, c->candbuf, c->canddepth, c->seen, c->memo_mask, colour, stage, xchg_dev);
    c->launches++;
    CU(cudaGetLastError());
    if (!c->opt_shard_async) CU(cudaStreamSynchronize(c->stream));       // the caller's collective may run on another stream
    return GPM_OK;
}

extern "C" int gpm_shard_accept(gpm_ctx* c, int colour, int stage, const float* gathered_dev, int world)
{
    if (!c || !gathered_dev || colour < 0 || colour > 1 || world < 1 || world > 8) return fail(GPM_E_ARG, "gpm_shard_accept: bad arguments");
    DeviceGuard g(c->device);
    KParams P;
    int rc = shard_common(c, stage, P);
    if (rc) return rc;
    dim3 b(32, 8), gr(((P.W + 1) / 2 + 31) / 32, (P.H + 7) / 8);
    k_shard_accept<<<gr, b, 0, c->stream>>>(P, c->planes, c->cost, c->prov, c->dispbuf, c->candbuf, c->canddepth, colour, stage,
                                            gathered_dev, world);
    c->launches++;
    CU(cudaGetLastError());
    if (!c->opt_shard_async) CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

// random planes only (stage 0 of the sharded flow computes the initial cost over all ranks' views)
extern "C" int gpm_init_planes(gpm_ctx* c)
{
    if (!c) return fail(GPM_E_ARG, "null context");
    DeviceGuard g(c->device);
    KParams P;
    int rc = build_kparams(c, true, P);
    if (rc) return rc;
    dim3 b(16, 16), gr((c->W + 15) / 16, (c->H + 15) / 16);
    k_init_planes<<<gr, b, 0, c->stream>>>(P, c->seed, c->planes, c->rng_mode == GPM_RNG_STATEFUL ? c->rng : nullptr);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

extern "C" int gpm_get_stats(gpm_ctx* c, unsigned long long stats[8])
{
    if (!c || !stats) return fail(GPM_E_ARG, "gpm_get_stats: null argument");
    DeviceGuard g(c->device);
    CU(cudaMemcpyAsync(stats, c->d_stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    stats[ST_LAUNCH] = c->launches;
    return GPM_OK;
}

extern "C" int gpm_reset_stats(gpm_ctx* c)
{
    if (!c) return fail(GPM_E_ARG, "null context");
    DeviceGuard g(c->device);
    CU(cudaMemsetAsync(c->d_stats, 0, 8 * sizeof(unsigned long long), c->stream));
    c->launches = 0;
    return GPM_OK;
}

extern "C" int gpm_set_option(gpm_ctx* c, const char* name, int value)
{
    if (!c || !name) return fail(GPM_E_ARG, "gpm_set_option: null argument");
    const std::string n(name);
    if (n == "prune") c->opt_prune = value != 0;
    else if (n == "dedupe") c->opt_dedupe = value != 0;
    else if (n == "trust_state") c->opt_trust_state = value != 0;
    else if (n == "nwarps") c->opt_nwarps = value;
    else if (n == "stats") c->opt_stats = value != 0;
    else if (n == "cost_variant") c->opt_cost_variant = value < 0 ? -1 : (value & 15);
    else if (n.rfind("site", 0) == 0 && n.size() > 4 && n.size() <= 6 && std::isdigit((unsigned char)n[4])) {
        const int k = std::atoi(n.c_str() + 4);             // "site<k>": variant of call site k of the fused kernel (diagnostics)
        if (k < 0 || k > 20) return fail(GPM_E_ARG, "site index out of range");
        c->opt_site[k] = value < 0 ? -1 : (value & 15);
    }
    else if (n == "neighbours") {
        if (value != 8 && value != 20) return fail(GPM_E_ARG, "neighbours must be 8 (close + far kernels) or 20 (fused kernel)");
        DeviceGuard g(c->device);
        if (value == 20) c->opt_packed = 0;
        const int slots = value == 20 ? 20 : 8;
        if (slots != c->seen_slots) {
            CU(cudaStreamSynchronize(c->stream));
            cudaFree(c->seen);  c->seen = nullptr;
            CU(cudaMalloc(&c->seen, (size_t)c->W * c->H * slots * sizeof(float4)));
            c->seen_slots = slots;
        }
        CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
        c->opt_neighbours = value;
    }
    else if (n == "packed") { c->opt_packed = value;  if (value) for (auto& f : c->view_8bit) f = 0; }   // set BEFORE uploading views;            // 0 off (default), 1 auto, 2 force — EXPERIMENTAL, see DESIGN.md §5
    else if (n == "memo") c->opt_memo = value != 0;
    else if (n == "quadperm") c->opt_quadperm = value != 0;
    else if (n == "shard_async") c->opt_shard_async = value != 0;
    else return fail(GPM_E_ARG, "gpm_set_option: unknown option '" + n + "'");
    return GPM_OK;
}

extern "C" void* gpm_stream(gpm_ctx* c) { return c ? (void*)c->stream : nullptr; }
