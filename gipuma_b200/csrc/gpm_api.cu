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

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace gpm;

#ifndef GPM_PREPASS_DEFAULT
#define GPM_PREPASS_DEFAULT 0
#endif

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
    unsigned char* sflags = nullptr; // view-shard mode: per-pixel flags of the current colour pass (GPM_SF_*)
    float* xchg = nullptr;           // view-shard mode: this rank's lists of the current stage / all ranks' lists (gpm_shard_run)
    float* gath = nullptr;
    size_t xchg_floats = 0;
    int gath_world = 0;
    void* comm = nullptr;            // ncclComm_t of the shard group
    bool comm_owned = false;
    int shard_rank = 0, shard_world = 0;
    unsigned long long collectives = 0;
    // fused peer-memory exchange (k_shard_fused): this rank's region, the peers' regions as mapped here, sequence counter
    char* p2p_region = nullptr;
    size_t p2p_bytes = 0, p2p_slot_floats = 0, p2p_flag_bytes = 0;
    int p2p_world = 0, p2p_nblocks = 0;
    bool p2p_attached = false;
    void* p2p_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool p2p_ipc_opened[8] = {false, false, false, false, false, false, false, false};
    unsigned p2p_seq = 0;
    int opt_exchange = 1;
    int opt_equal_rounds = 0;
    int opt_prepass = GPM_PREPASS_DEFAULT;       // k_sweep's thread-per-pixel pre-pass over idle pixels (env GPM_PREPASS overrides at gpm_create)
    int opt_fused_warps = 16;                    // warps per block of k_shard_fused: 8 -> two blocks per SM (one samples while the other waits); measured no faster
    int opt_async_upload = 0;                    // 1: image uploads return without a host synchronisation (caller keeps its buffers alive until the next run)
    bool inputs_dirty = false;                   // an input changed: stored costs / memo are stale (cleared once, at the next launch)                        // 1: peer-memory exchange when attached; 0: NCCL all-gather per stage
    unsigned* seen = nullptr;        // [H*W*ncand] identity of the plane last offered to each pixel from each of the 8 (fused kernel: 20) propagation directions
    int seen_slots = 8;
    unsigned* refseen = nullptr;     // [H*W]   identity of the plane from which the last all-rejected refinement started
    unsigned* pid = nullptr;         // [H*W]   identity of the stored plane (gpm_kernels.cuh, struct Memo)
    unsigned* next_id = nullptr;     // device counter of fresh identities
    unsigned* memo_mask = nullptr;   // [H*W] validity bits of seen (0-19) and refseen (GPM_MEMO_REFINE)
    unsigned char* prov = nullptr;   // per pixel: which rounding variant of the cost function produced cost[] (see k_sweep)
    float* refpad = nullptr;
    int refpitch = 0, refrows = 0;   // padded reference image: refrows x refpitch texels
    CUtensorMap tmap{};              // 2-D TMA descriptor of the padded reference image for the current box size / image type
    int tmap_box = 0, tmap_color = -1;
    int opt_tma = 1;
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

// 2-D TMA descriptor (cuTensorMapEncodeTiled, fetched through the runtime so that the library links only cudart) of the
// padded reference image: box = one tile's window, tile_stride x tile_w texels (float4 images: 4 floats per texel).
int ensure_tensor_map(gpm_ctx* c, const KParams& P)
{
    const int box = c->prm.box_hsize;
    if (c->tmap_box == box && c->tmap_color == P.color) return GPM_OK;
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(GPM_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (EncodeTiled)fn;
    }
    const int comps = P.color ? 4 : 1;
    void* base = P.color ? (void*)c->refpad4 : (void*)c->refpad;
    const cuuint64_t gdim[2] = {(cuuint64_t)c->refpitch * comps, (cuuint64_t)c->refrows};
    const cuuint64_t gstride[1] = {(cuuint64_t)c->refpitch * comps * sizeof(float)};
    const cuuint32_t bdim[2] = {(cuuint32_t)(P.tile_stride * comps), (cuuint32_t)P.tile_w};
    const cuuint32_t estr[2] = {1, 1};
    if (bdim[0] > 256 || bdim[1] > 256) return fail(GPM_E_ARG, "window too large for one TMA box");
    const CUresult r = encode(&c->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(GPM_E_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    c->tmap_box = box;  c->tmap_color = P.color;
    return GPM_OK;
}

int build_kparams(gpm_ctx* c, bool init_phase, KParams& P, bool eval_call = false)
{
    if (!c->have_params) return fail(GPM_E_STATE, "gpm_set_params has not been called");
    if (!c->have_ref) return fail(GPM_E_STATE, "gpm_set_reference has not been called");
    if (c->V < 1) return fail(GPM_E_STATE, "no source views (gpm_set_num_views)");
    for (int v = 0; v < c->V; v++)
        if (!c->have_view[v]) return fail(GPM_E_STATE, "source view " + std::to_string(v) + " has not been set");
    if (c->inputs_dirty) {                 // one clearing for a whole scene upload instead of two memsets per image
        CU(cudaMemsetAsync(c->prov, GPM_PROV_UNKNOWN, (size_t)c->W * c->H, c->stream));
        CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
        c->inputs_dirty = false;
    }
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
    P.tile_stride = (P.tile_w + 15) & ~15;                  // 48 or 64 texels per staged row
    // a TMA box must start on a 16-byte boundary of the image (boxes starting at byte offsets 12, 28, 40, 44 fault on B200,
    // profiles/r02_tma_alignment.txt): tile_x0 + GPM_APRON = 32 bx + 16 - halo, so the box starts (16 - halo) mod 4 texels early
    P.tile_xo = (c->color == 1) ? 0 : ((GPM_APRON - P.halo) & 3);
    // rounds of 32 consecutive samples (one per lane); the remainder forms a last, shorter round whose (view, sample) pairs
    // are packed several views to an instruction.  (Option "equal_rounds": windows with 33..47 samples — blocksize 11 —
    // as two equal rounds, the round-1 arrangement; 32 + 4 keeps whole 2x2 quads and measured faster)
    {
        int r = 0;
        if (P.ns > 32 && P.ns < 48 && c->opt_equal_rounds) {
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
    P.prepass = c->opt_prepass;
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
    // The lower bound counts a view as valid from its PARTIAL cost; a view whose final cost reaches MAXCOST (then excluded,
    // gipuma.cu:771-774) would make the bound exceed the final value.  Unreachable while ns * max dissimilarity < MAXCOST.
    if ((float)P.ns * ((1.0f - p.alpha) * p.tau_color + p.alpha * p.tau_gradient) >= GPM_MAXCOST) P.prune = 0;
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
    P.use_tma = c->opt_tma ? 1 : 0;
    if (P.use_tma) {
        int rc = ensure_tensor_map(c, P);
        if (rc) return rc;
    }
    return GPM_OK;
}

// identities 1 .. W*H belong to the pixels' initial planes; fresh ones start above
int reset_next_id(gpm_ctx* c)
{
    const unsigned first = (unsigned)((size_t)c->W * c->H) + 1u;
    CU(cudaMemcpyAsync(c->next_id, &first, sizeof(first), cudaMemcpyHostToDevice, c->stream));
    return GPM_OK;
}

int fill_ids(gpm_ctx* c)
{
    const unsigned n = (unsigned)((size_t)c->W * c->H);
    k_fill_ids<<<(n + 255) / 256, 256, 0, c->stream>>>(c->pid, n);
    CU(cudaGetLastError());
    return reset_next_id(c);
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
    kern<<<grid, P.nwarps * 32, smem, c->stream>>>(P, c->tmap, c->d_cams, P.color ? (const float*)c->refpad4 : c->refpad, P.color ? c->srcTex4 : c->srcTex, c->gradTex, c->planes, c->cost, c->rng,
                                                      c->prov, Memo{c->pid, c->seen, c->refseen, c->memo_mask, c->next_id}, colour, mask, c->opt_stats ? c->d_stats : nullptr);
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
    if (const char* e = getenv("GPM_PREPASS")) c->opt_prepass = atoi(e) != 0;
    c->device = device;  c->W = width;  c->H = height;  c->maxV = max_views;
    c->have_view.assign(max_views, 0);
    c->view_8bit.assign(max_views, 0);
    c->h_cams.assign(max_views, ViewCam{});
    const size_t n = (size_t)width * height;
    // the staged window of the last tile row / column reaches roundup(size, 32) + halo + GPM_APRON (+ 3 texels of TMA row padding)
    c->refpitch = (((width + 31) & ~31) + 2 * GPM_APRON + 4 + 31) & ~31;
    c->refrows = ((height + 31) & ~31) + 2 * GPM_APRON;
    cudaError_t err = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (err == cudaSuccess && r != cudaSuccess) err = r; return r == cudaSuccess; };
    ok(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    ok(cudaEventCreate(&c->ev0));
    ok(cudaEventCreate(&c->ev1));
    ok(cudaMalloc(&c->planes, n * sizeof(float4)));
    ok(cudaMalloc(&c->cost, n * sizeof(float)));
    ok(cudaMalloc(&c->prov, n));
    ok(cudaMalloc(&c->seen, n * 8 * sizeof(unsigned)));
    ok(cudaMalloc(&c->refseen, n * sizeof(unsigned)));
    ok(cudaMalloc(&c->pid, n * sizeof(unsigned)));
    ok(cudaMalloc(&c->next_id, sizeof(unsigned)));
    ok(cudaMalloc(&c->memo_mask, n * sizeof(unsigned)));
    ok(cudaMalloc(&c->staging, n * sizeof(float)));
    ok(cudaMalloc(&c->d_flag, sizeof(int)));
    ok(cudaMalloc(&c->refpad, (size_t)c->refpitch * c->refrows * sizeof(float)));
    ok(cudaMalloc(&c->d_cams, sizeof(ViewCam) * max_views));
    ok(cudaMalloc(&c->d_stats, 8 * sizeof(unsigned long long)));
    if (err == cudaSuccess) {
        ok(cudaMemsetAsync(c->planes, 0, n * sizeof(float4), c->stream));      // LineState::resize zeroes (linestate.h:19-24)
        ok(cudaMemsetAsync(c->cost, 0, n * sizeof(float), c->stream));
        ok(cudaMemsetAsync(c->prov, GPM_PROV_UNKNOWN, n, c->stream));
        ok(cudaMemsetAsync(c->memo_mask, 0, n * sizeof(unsigned), c->stream));
        ok(cudaMemsetAsync(c->d_stats, 0, 8 * sizeof(unsigned long long), c->stream));
        if (fill_ids(c) != GPM_OK) err = cudaErrorUnknown;
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
        ok(cudaFuncSetAttribute(k_shard_stage<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_shard_fused<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_shard_stage<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_shard_fused<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_shard_stage<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
        ok(cudaFuncSetAttribute(k_shard_fused<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin));
    }
    if (err != cudaSuccess) {
        std::string m = std::string("gpm_create: ") + cudaGetErrorString(err);
        gpm_destroy(c);
        return fail(GPM_E_CUDA, m);
    }
    *out = c;
    return GPM_OK;
}

static void gpm_shard_comm_destroy_(gpm_ctx* c);

extern "C" void gpm_destroy(gpm_ctx* c)
{
    if (!c) return;
    DeviceGuard g(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && c->comm_owned) gpm_shard_comm_destroy_(c);
    for (int r = 0; r < 8; r++) if (c->p2p_ipc_opened[r] && c->p2p_peer[r]) cudaIpcCloseMemHandle(c->p2p_peer[r]);
    cudaFree(c->p2p_region);
    if (c->srcTex) cudaDestroyTextureObject(c->srcTex);
    if (c->srcArr) cudaFreeArray(c->srcArr);
    if (c->srcTex4) cudaDestroyTextureObject(c->srcTex4);
    if (c->srcArr4) cudaFreeArray(c->srcArr4);
    cudaFree(c->refpad4);  cudaFree(c->staging4);  cudaFree(c->planar);
    if (c->gradTex) cudaDestroyTextureObject(c->gradTex);
    if (c->gradArr) cudaFreeArray(c->gradArr);
    cudaFree(c->gradLin);  cudaFree(c->d_flag);
    cudaFree(c->planes);  cudaFree(c->cost);  cudaFree(c->prov);  cudaFree(c->seen);  cudaFree(c->refseen);  cudaFree(c->pid);  cudaFree(c->next_id);  cudaFree(c->memo_mask);  cudaFree(c->rng);  cudaFree(c->dispbuf);  cudaFree(c->candbuf);  cudaFree(c->canddepth);  cudaFree(c->sflags);  cudaFree(c->xchg);  cudaFree(c->gath);  cudaFree(c->refpad);  cudaFree(c->staging);
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
    if (n != c->V) c->inputs_dirty = true;           // the cost function changes with the view count: stored costs / memo are stale
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
    dim3 b(32, 8), gr((c->refpitch + 31) / 32, (c->refrows + 7) / 8);
    k_pad_reference<<<gr, b, 0, c->stream>>>(d, pf, c->W, c->H, c->refpad, c->refpitch, c->refrows);
    CU(cudaGetLastError());
    set_ref_camera(c, cam);
    c->inputs_dirty = true;
    if (!c->opt_async_upload) CU(cudaStreamSynchronize(c->stream));     // the caller may reuse its buffer
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
    cudaMemcpy3DParms m;  memset(&m, 0, sizeof(m));
    m.dstArray = c->srcArr;
    m.dstPos = make_cudaPos(0, 0, v);
    m.extent = make_cudaExtent(c->W, c->H, 1);
    if (!on_device && !c->opt_packed) {
        // host image straight into its layer of the array: no staging copy
        const size_t pb = pitch_bytes ? pitch_bytes : (size_t)c->W * sizeof(float);
        if (pb % sizeof(float)) return fail(GPM_E_ARG, "pitch_bytes must be a multiple of 4");
        m.srcPtr = make_cudaPitchedPtr(const_cast<float*>(img), pb, c->W, c->H);
        m.kind = cudaMemcpyHostToDevice;
    } else {
        rc = upload_image(c, img, pitch_bytes, on_device, &d, &pf);      // staging buffer (the gradient planes of the packed mode read it)
        if (rc) return rc;
        m.srcPtr = make_cudaPitchedPtr(d, pf * sizeof(float), c->W, c->H);
        m.kind = cudaMemcpyDeviceToDevice;
    }
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
    c->inputs_dirty = true;
    if (!c->opt_async_upload) CU(cudaStreamSynchronize(c->stream));        // the caller may reuse its buffer
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
        CU(cudaMalloc(&c->refpad4, (size_t)c->refpitch * c->refrows * sizeof(float4)));
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
    dim3 b(32, 8), gr((c->refpitch + 31) / 32, (c->refrows + 7) / 8);
    k_pad_reference4<<<gr, b, 0, c->stream>>>(d, pe, c->W, c->H, c->refpad4, c->refpitch, c->refrows);
    CU(cudaGetLastError());
    set_ref_camera(c, cam);
    c->inputs_dirty = true;
    if (!c->opt_async_upload) CU(cudaStreamSynchronize(c->stream));
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
    c->inputs_dirty = true;
    if (!c->opt_async_upload) CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

extern "C" int gpm_set_state(gpm_ctx* c, const float* norm4, const float* cost, int on_device)
{
    if (!c) return fail(GPM_E_ARG, "gpm_set_state: null context");
    DeviceGuard g(c->device);
    const size_t n = (size_t)c->W * c->H;
    const cudaMemcpyKind k = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (norm4) {
        CU(cudaMemcpyAsync(c->planes, norm4, n * sizeof(float4), k, c->stream));
        int rc = fill_ids(c);               // caller-supplied planes: every pixel's plane is its own (equal values are simply not recognised as such)
        if (rc) return rc;
    }
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
    k_init_planes<<<gr, b, 0, c->stream>>>(P, c->seed, c->planes, c->rng_mode == GPM_RNG_STATEFUL ? c->rng : nullptr, c->pid);
    { int rc_ = reset_next_id(c);  if (rc_) return rc_; }
    c->launches++;
    CU(cudaGetLastError());
    dim3 grid((P.W + GPM_TILE - 1) / GPM_TILE, (P.H + GPM_TILE - 1) / GPM_TILE);
    (P.color ? k_cost_eval<false, true> : (P.packed ? k_cost_eval<true, false> : k_cost_eval<false, false>))<<<grid, P.nwarps * 32, block_smem_bytes(P), c->stream>>>(P, c->tmap, c->d_cams, P.color ? (const float*)c->refpad4 : c->refpad, P.color ? c->srcTex4 : c->srcTex, c->gradTex, c->planes,
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
        cudaError_t e1 = cudaMalloc(&d_pl, n * sizeof(float4));
        if (e1 == cudaSuccess) e1 = cudaMalloc(&d_out, n * sizeof(float));
        if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(d_pl, planes, n * sizeof(float4), cudaMemcpyHostToDevice, c->stream);
        if (e1 != cudaSuccess) { cudaFree(d_pl);  cudaFree(d_out);  return fail(GPM_E_CUDA, std::string("gpm_cost_eval: ") + cudaGetErrorString(e1)); }
    }
    dim3 grid((P.W + GPM_TILE - 1) / GPM_TILE, (P.H + GPM_TILE - 1) / GPM_TILE);
    (P.color ? k_cost_eval<false, true> : (P.packed ? k_cost_eval<true, false> : k_cost_eval<false, false>))<<<grid, P.nwarps * 32, block_smem_bytes(P), c->stream>>>(P, c->tmap, c->d_cams, P.color ? (const float*)c->refpad4 : c->refpad, P.color ? c->srcTex4 : c->srcTex, c->gradTex, d_pl, d_out, nullptr);
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
    if (c->rng_mode != GPM_RNG_REFERENCE) return fail(GPM_E_ARG, "view sharding implements the reference's zero-state refinement stream only (GPM_RNG_REFERENCE)");
    const int last = 1 + shard_refine_steps(c->prm);
    if (stage < 0 || stage > last + 1) return fail(GPM_E_ARG, "no such stage");
    int rc = build_kparams(c, stage == 0, P);
    if (rc) return rc;
    rc = sync_cams(c);
    if (rc) return rc;
    if (!c->dispbuf) {
        const size_t n = (size_t)c->W * c->H;
        CU(cudaMalloc(&c->dispbuf, n * sizeof(float)));
        CU(cudaMalloc(&c->candbuf, n * sizeof(float4)));
        CU(cudaMalloc(&c->canddepth, n * sizeof(float)));
        CU(cudaMalloc(&c->sflags, n));
        CU(cudaMemsetAsync(c->sflags, 0, n, c->stream));
    }
    return GPM_OK;
}

// enqueue one stage kernel (stage 0 .. last+1) on the context's stream
static int shard_launch_stage(gpm_ctx* c, const KParams& P, int colour, int stage, const float* gathered_prev, int world, float* xchg)
{
    const int last = 1 + shard_refine_steps(c->prm);
    dim3 grid((P.W + GPM_TILE - 1) / GPM_TILE, (P.H + GPM_TILE - 1) / GPM_TILE);
    int split = 1;                                        // at least ~8 waves of blocks, as in launch_colour
    while (split < 8 && (long long)grid.x * grid.y * split < 8LL * c->num_sms) split *= 2;
    grid.z = split;
    ShardState S{c->dispbuf, c->candbuf, c->canddepth, c->sflags};
    auto kern = P.color ? k_shard_stage<false, true> : (P.packed ? k_shard_stage<true, false> : k_shard_stage<false, false>);
    kern<<<grid, P.nwarps * 32, block_smem_bytes(P), c->stream>>>(P, c->tmap, c->d_cams, P.color ? (const float*)c->refpad4 : c->refpad,
        P.color ? c->srcTex4 : c->srcTex, c->gradTex, c->planes, c->cost, c->prov, S, Memo{c->pid, c->seen, c->refseen, c->memo_mask, c->next_id}, colour, stage, last,
        gathered_prev, world, xchg, c->opt_stats ? c->d_stats : nullptr);
    c->launches++;
    CU(cudaGetLastError());
    return GPM_OK;
}

extern "C" int gpm_shard_stage(gpm_ctx* c, int colour, int stage, const float* gathered_prev_dev, int world, float* xchg_dev)
{
    if (!c || colour < 0 || colour > 1 || world < 1 || world > 8) return fail(GPM_E_ARG, "gpm_shard_stage: bad arguments");
    DeviceGuard g(c->device);
    KParams P;
    int rc = shard_common(c, stage, P);
    if (rc) return rc;
    const int last = 1 + shard_refine_steps(c->prm);
    if (stage >= 2 && !gathered_prev_dev) return fail(GPM_E_ARG, "gpm_shard_stage: stages >= 2 need the gathered lists of the previous stage");
    if (stage <= last && !xchg_dev) return fail(GPM_E_ARG, "gpm_shard_stage: no exchange buffer");
    rc = shard_launch_stage(c, P, colour, stage, gathered_prev_dev, world, xchg_dev);
    if (rc) return rc;
    if (!c->opt_shard_async) CU(cudaStreamSynchronize(c->stream));       // the caller's collective may run on another stream
    return GPM_OK;
}

extern "C" int gpm_shard_finish_init(gpm_ctx* c, const float* gathered_dev, int world)
{
    if (!c || !gathered_dev || world < 1 || world > 8) return fail(GPM_E_ARG, "gpm_shard_finish_init: bad arguments");
    DeviceGuard g(c->device);
    KParams P;
    int rc = shard_common(c, 0, P);
    if (rc) return rc;
    dim3 b(32, 8), gr(((P.W + 1) / 2 + 31) / 32, (P.H + 7) / 8);
    k_shard_accept0<<<gr, b, 0, c->stream>>>(P, c->cost, c->prov, gathered_dev, world);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
    if (!c->opt_shard_async) CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

// ---- the exchange behind the C-ABI: NCCL, loaded at run time (a C++ host needs no torch) ---------------------------
#include <dlfcn.h>
namespace {
struct NcclId { char internal[128]; };
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};
NcclApi* nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {getenv("GPM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
        }
        if (api.h) {
            api.GetUniqueId = (int (*)(NcclId*))dlsym(api.h, "ncclGetUniqueId");
            api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(api.h, "ncclCommInitRank");
            api.CommDestroy = (int (*)(void*))dlsym(api.h, "ncclCommDestroy");
            api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(api.h, "ncclAllGather");
            api.GetErrorString = (const char* (*)(int))dlsym(api.h, "ncclGetErrorString");
            api.GetVersion = (int (*)(int*))dlsym(api.h, "ncclGetVersion");
            if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather) { dlclose(api.h);  api.h = nullptr; }
        }
    }
    return api.h ? &api : nullptr;
}
}  // namespace

#define NC(call)                                                                                                        \
    do {                                                                                                                \
        int e_ = (call);                                                                                                \
        if (e_ != 0) return fail(GPM_E_CUDA, std::string(#call) + ": NCCL error " + (nccl_api() && nccl_api()->GetErrorString ? nccl_api()->GetErrorString(e_) : "?")); \
    } while (0)

static void gpm_shard_comm_destroy_(gpm_ctx* c)
{
    if (c->comm && c->comm_owned && nccl_api()) nccl_api()->CommDestroy(c->comm);
    c->comm = nullptr;  c->comm_owned = false;
}

extern "C" int gpm_shard_unique_id(void* id128)
{
    if (!id128) return fail(GPM_E_ARG, "gpm_shard_unique_id: null argument");
    NcclApi* n = nccl_api();
    if (!n) return fail(GPM_E_STATE, "NCCL (libnccl.so.2) could not be loaded; set GPM_NCCL_LIB");
    NC(n->GetUniqueId((NcclId*)id128));
    return GPM_OK;
}

static int shard_alloc_exchange(gpm_ctx* c)
{
    long long mx = 0;
    const int ns = gpm_shard_num_stages(c);
    for (int st = 0; st < ns; st++) { const long long f = gpm_shard_stage_floats(c, st);  if (f > mx) mx = f; }
    if (c->xchg && c->xchg_floats >= (size_t)mx && c->gath_world >= c->shard_world) return GPM_OK;
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(c->xchg);  cudaFree(c->gath);  c->xchg = c->gath = nullptr;
    CU(cudaMalloc(&c->xchg, (size_t)mx * sizeof(float)));
    CU(cudaMalloc(&c->gath, (size_t)mx * sizeof(float) * c->shard_world));
    c->xchg_floats = (size_t)mx;  c->gath_world = c->shard_world;
    return GPM_OK;
}

extern "C" int gpm_shard_comm_init(gpm_ctx* c, const void* id128, int rank, int world)
{
    if (!c || world < 1 || world > 8 || rank < 0 || rank >= world) return fail(GPM_E_ARG, "gpm_shard_comm_init: bad arguments (1 <= world <= 8)");
    DeviceGuard g(c->device);
    if (c->comm && c->comm_owned) { nccl_api()->CommDestroy(c->comm); }
    c->comm = nullptr;  c->comm_owned = false;
    c->shard_rank = rank;  c->shard_world = world;
    if (world > 1) {
        if (!id128) return fail(GPM_E_ARG, "gpm_shard_comm_init: null unique id");
        NcclApi* n = nccl_api();
        if (!n) return fail(GPM_E_STATE, "NCCL (libnccl.so.2) could not be loaded; set GPM_NCCL_LIB");
        NcclId id;  memcpy(&id, id128, sizeof(id));
        NC(n->CommInitRank(&c->comm, world, id, rank));
        c->comm_owned = true;
    }
    return GPM_OK;
}

extern "C" int gpm_shard_comm_attach(gpm_ctx* c, void* nccl_comm, int rank, int world)
{
    if (!c || world < 1 || world > 8 || rank < 0 || rank >= world || (world > 1 && !nccl_comm)) return fail(GPM_E_ARG, "gpm_shard_comm_attach: bad arguments");
    if (world > 1 && !nccl_api()) return fail(GPM_E_STATE, "NCCL (libnccl.so.2) could not be loaded; set GPM_NCCL_LIB");
    if (c->comm && c->comm_owned) nccl_api()->CommDestroy(c->comm);
    c->comm = nccl_comm;  c->comm_owned = false;  c->shard_rank = rank;  c->shard_world = world;
    return GPM_OK;
}

// all-gather of this rank's lists of one stage (rank-major result), on the context's stream
static int shard_exchange(gpm_ctx* c, int stage)
{
    const size_t n = (size_t)gpm_shard_stage_floats(c, stage);
    if (c->shard_world == 1) { CU(cudaMemcpyAsync(c->gath, c->xchg, n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));  return GPM_OK; }
    NC(nccl_api()->AllGather(c->xchg, c->gath, n, 7 /* ncclFloat32 */, c->comm, c->stream));
    c->collectives++;
    return GPM_OK;
}

// ---- fused peer-memory exchange: region management ------------------------------------------------------------------
static dim3 shard_grid(gpm_ctx* c)
{
    dim3 grid((c->W + GPM_TILE - 1) / GPM_TILE, (c->H + GPM_TILE - 1) / GPM_TILE);
    int split = 1;                                        // at least ~8 waves of blocks, as in launch_colour
    while (split < 8 && (long long)grid.x * grid.y * split < 8LL * c->num_sms) split *= 2;
    grid.z = split;
    return grid;
}

// Allocate this rank's exchange region for a shard group of `world` ranks and return its CUDA IPC handle (64 bytes) and its
// address (for ranks living in the same process).  Layout: [world][nblocks] arrival flags, then [2][world][slot] lists.
extern "C" int gpm_shard_p2p_export(gpm_ctx* c, int world, void* handle64, void** local_ptr)
{
    if (!c || world < 2 || world > 8) return fail(GPM_E_ARG, "gpm_shard_p2p_export: bad arguments (2 <= world <= 8)");
    if (!c->have_params) return fail(GPM_E_STATE, "gpm_shard_p2p_export: parameters not set");
    DeviceGuard g(c->device);
    long long mx = 0;
    const int ns = gpm_shard_num_stages(c);
    for (int st = 0; st < ns; st++) { const long long f = gpm_shard_stage_floats(c, st);  if (f > mx) mx = f; }
    const dim3 grid = shard_grid(c);
    const int nblocks = (int)(grid.x * grid.y * grid.z);
    const size_t flag_bytes = ((size_t)world * nblocks * sizeof(unsigned) + 1023) & ~(size_t)1023;
    const size_t bytes = flag_bytes + 2ull * world * (size_t)mx * sizeof(float);
    if (!c->p2p_region || c->p2p_bytes != bytes || c->p2p_world != world) {
        CU(cudaStreamSynchronize(c->stream));
        for (int r = 0; r < 8; r++) { if (c->p2p_ipc_opened[r] && c->p2p_peer[r]) cudaIpcCloseMemHandle(c->p2p_peer[r]);  c->p2p_peer[r] = nullptr;  c->p2p_ipc_opened[r] = false; }
        cudaFree(c->p2p_region);  c->p2p_region = nullptr;  c->p2p_attached = false;
        CU(cudaMalloc(&c->p2p_region, bytes));
        c->p2p_bytes = bytes;  c->p2p_world = world;
    }
    c->p2p_slot_floats = (size_t)mx;  c->p2p_flag_bytes = flag_bytes;  c->p2p_nblocks = nblocks;
    CU(cudaMemsetAsync(c->p2p_region, 0, flag_bytes, c->stream));          // flags: nothing has arrived
    CU(cudaStreamSynchronize(c->stream));
    c->p2p_seq = 0;
    if (handle64) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, c->p2p_region));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(handle64, &h, 64);
    }
    if (local_ptr) *local_ptr = c->p2p_region;
    return GPM_OK;
}

// Map the regions of all ranks of the group.  `handles64`: world x 64 bytes (rank-major) from gpm_shard_p2p_export of every
// rank, opened with cudaIpcOpenMemHandle (ranks in other processes); `local_ptrs` (may be NULL): world addresses, a non-NULL
// entry is used as is (ranks in this process).  Every rank must have exported before any rank attaches.
extern "C" int gpm_shard_p2p_attach(gpm_ctx* c, const void* handles64, void* const* local_ptrs, int rank, int world)
{
    if (!c || world < 2 || world > 8 || rank < 0 || rank >= world) return fail(GPM_E_ARG, "gpm_shard_p2p_attach: bad arguments");
    if (!c->p2p_region || c->p2p_world != world) return fail(GPM_E_STATE, "gpm_shard_p2p_attach: call gpm_shard_p2p_export(ctx, world, ...) first");
    DeviceGuard g(c->device);
    for (int r = 0; r < world; r++) {
        if (c->p2p_ipc_opened[r] && c->p2p_peer[r]) cudaIpcCloseMemHandle(c->p2p_peer[r]);
        c->p2p_peer[r] = nullptr;  c->p2p_ipc_opened[r] = false;
        if (r == rank) { c->p2p_peer[r] = c->p2p_region;  continue; }
        if (local_ptrs && local_ptrs[r]) { c->p2p_peer[r] = local_ptrs[r];  continue; }
        if (!handles64) return fail(GPM_E_ARG, "gpm_shard_p2p_attach: no handle for a peer");
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles64 + 64 * (size_t)r, 64);
        void* ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->p2p_peer[r] = ptr;  c->p2p_ipc_opened[r] = true;
    }
    c->shard_rank = rank;  c->shard_world = world;  c->p2p_attached = true;
    return GPM_OK;
}

static int shard_launch_fused(gpm_ctx* c, const KParams& P, int colour, int init_phase, int exchanges)
{
    const int last = 1 + shard_refine_steps(c->prm);
    const dim3 grid = shard_grid(c);
    ShardState S{c->dispbuf, c->candbuf, c->canddepth, c->sflags};
    ShardP2P X;
    memset(&X, 0, sizeof(X));
    for (int r = 0; r < c->shard_world; r++) {
        X.flags[r] = reinterpret_cast<unsigned*>(c->p2p_peer[r]);
        X.lists[r] = reinterpret_cast<float*>(reinterpret_cast<char*>(c->p2p_peer[r]) + c->p2p_flag_bytes);
    }
    X.err = reinterpret_cast<unsigned*>(c->d_stats + 7);
    X.slot_floats = c->p2p_slot_floats;  X.me = c->shard_rank;  X.world = c->shard_world;  X.nblocks = c->p2p_nblocks;
    auto kern = P.color ? k_shard_fused<false, true> : (P.packed ? k_shard_fused<true, false> : k_shard_fused<false, false>);
    KParams Pf = P;                                       // half-size blocks: two per SM (128 registers x 256 threads each)
    if (c->opt_fused_warps > 0 && c->opt_fused_warps < Pf.nwarps) Pf.nwarps = c->opt_fused_warps;
    kern<<<grid, Pf.nwarps * 32, block_smem_bytes(Pf), c->stream>>>(Pf, c->tmap, c->d_cams, P.color ? (const float*)c->refpad4 : c->refpad,
        P.color ? c->srcTex4 : c->srcTex, c->gradTex, c->planes, c->cost, c->prov, S, Memo{c->pid, c->seen, c->refseen, c->memo_mask, c->next_id}, colour, init_phase, last,
        X, c->p2p_seq, c->opt_stats ? c->d_stats : nullptr);
    c->launches++;
    c->p2p_seq += (unsigned)exchanges;
    c->collectives += (unsigned long long)exchanges;
    CU(cudaGetLastError());
    return GPM_OK;
}

// runcuda() with the source views sharded over the ranks of the attached communicator.  Every rank calls it with the same
// parameters, reference image and seed and ITS views; all work — kernels and collectives — is enqueued on the context's
// stream, one host synchronisation at the end.
extern "C" int gpm_shard_run(gpm_ctx* c, float* sweep_ms)
{
    if (!c) return fail(GPM_E_ARG, "null context");
    if (c->shard_world < 1) return fail(GPM_E_STATE, "gpm_shard_run: no communicator (gpm_shard_comm_init / gpm_shard_comm_attach)");
    DeviceGuard g(c->device);
    KParams P0, P;
    int rc = shard_common(c, 0, P0);
    if (rc) return rc;
    rc = shard_common(c, 1, P);
    if (rc) return rc;
    rc = shard_alloc_exchange(c);
    if (rc) return rc;
    const int world = c->shard_world, last = 1 + shard_refine_steps(c->prm);
    CU(cudaMemsetAsync(c->d_stats, 0, 8 * sizeof(unsigned long long), c->stream));
    c->launches = 0;  c->collectives = 0;
    if (world > 1 && c->p2p_attached && c->opt_exchange == 1) {
        // fused flow: one launch for the initial costs, one per colour pass; the exchange happens inside the kernels
        dim3 b(16, 16), gr((c->W + 15) / 16, (c->H + 15) / 16);
        k_init_planes<<<gr, b, 0, c->stream>>>(P0, c->seed, c->planes, nullptr, c->pid);
        c->launches++;
        CU(cudaGetLastError());
        rc = reset_next_id(c);
        if (rc) return rc;
        rc = shard_launch_fused(c, P0, 0, 1, 1);
        if (rc) return rc;
        CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
        CU(cudaEventRecord(c->ev0, c->stream));
        for (int it = 0; it < c->prm.iterations; it++)
            for (int colour = 0; colour < 2; colour++) {
                rc = shard_launch_fused(c, P, colour, 0, last);
                if (rc) return rc;
            }
        rc = do_finalize(c);
        if (rc) return rc;
        CU(cudaEventRecord(c->ev1, c->stream));
        CU(cudaEventSynchronize(c->ev1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (sweep_ms) *sweep_ms = ms;
        unsigned long long err = 0;
        CU(cudaMemcpyAsync(&err, c->d_stats + 7, sizeof(err), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (err & 0xffffffffull) return fail(GPM_E_STATE, "gpm_shard_run: a peer rank did not deliver its lists in time (peer-memory exchange)");
        return GPM_OK;
    }
    if (world > 1 && !c->comm) return fail(GPM_E_STATE, "gpm_shard_run: neither an NCCL communicator nor peer regions are attached");
    {
        dim3 b(16, 16), gr((c->W + 15) / 16, (c->H + 15) / 16);
        k_init_planes<<<gr, b, 0, c->stream>>>(P0, c->seed, c->planes, nullptr, c->pid);
        c->launches++;
        CU(cudaGetLastError());
        rc = reset_next_id(c);
        if (rc) return rc;
        CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
        rc = shard_launch_stage(c, P0, 0, 0, nullptr, world, c->xchg);
        if (rc) return rc;
        rc = shard_exchange(c, 0);
        if (rc) return rc;
        dim3 b2(32, 8), gr2(((P0.W + 1) / 2 + 31) / 32, (P0.H + 7) / 8);
        k_shard_accept0<<<gr2, b2, 0, c->stream>>>(P0, c->cost, c->prov, c->gath, world);
        c->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(c->ev0, c->stream));
    for (int it = 0; it < c->prm.iterations; it++)
        for (int colour = 0; colour < 2; colour++) {
            for (int stage = 1; stage <= last; stage++) {
                rc = shard_launch_stage(c, P, colour, stage, c->gath, world, c->xchg);
                if (rc) return rc;
                rc = shard_exchange(c, stage);
                if (rc) return rc;
            }
            rc = shard_launch_stage(c, P, colour, last + 1, c->gath, world, c->xchg);     // closing accept of the colour
            if (rc) return rc;
        }
    rc = do_finalize(c);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (sweep_ms) *sweep_ms = ms;
    CU(cudaStreamSynchronize(c->stream));
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
    k_init_planes<<<gr, b, 0, c->stream>>>(P, c->seed, c->planes, c->rng_mode == GPM_RNG_STATEFUL ? c->rng : nullptr, c->pid);
    { int rc_ = reset_next_id(c);  if (rc_) return rc_; }
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(c->prov, GPM_PROV_UNKNOWN, (size_t)c->W * c->H, c->stream));      // every plane is new: costs and memo are stale
    CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return GPM_OK;
}

// diagnostics of the experimental packed sampling mode ("packed" = 3): number of fetches whose one-fetch gradient differed
// from the reference's four fetches, and up to 64 records {cx, cy, view, gx packed, gx reference, gy packed, gy reference, centre}
extern "C" int gpm_debug_packed_mismatches(gpm_ctx* c, unsigned* count, float* records512, int reset)
{
    if (!c || !count) return fail(GPM_E_ARG, "gpm_debug_packed_mismatches: null argument");
    DeviceGuard g(c->device);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpyFromSymbol(count, g_packed_mismatch_n, sizeof(unsigned)));
    if (records512) CU(cudaMemcpyFromSymbol(records512, g_packed_mismatch, 64 * 8 * sizeof(float)));
    if (reset) { const unsigned z = 0;  CU(cudaMemcpyToSymbol(g_packed_mismatch_n, &z, sizeof(z))); }
    return GPM_OK;
}

extern "C" int gpm_measure_fetch_peak(gpm_ctx* c, double* gfetch_per_s)
{
    if (!c || !gfetch_per_s) return fail(GPM_E_ARG, "gpm_measure_fetch_peak: null argument");
    if (c->V < 1 || c->color < 0) return fail(GPM_E_STATE, "gpm_measure_fetch_peak: no source views");
    DeviceGuard g(c->device);
    float* sink = reinterpret_cast<float*>(c->d_stats + 7);
    const int blocks = c->num_sms * 16, threads = 256, reps = 2048;
    const cudaTextureObject_t tex = c->color == 1 ? c->srcTex4 : c->srcTex;      // colour: three R32F channel planes per view
    const int layers = c->color == 1 ? 3 * c->V : c->V;
    // random positions inside a window of at most 512 x 512 texels of at most 4 layers: the UNIT's ceiling is wanted, so the
    // working set must stay in L2 whatever the image size (a 3200 x 2400 x 64-layer array would make this a DRAM test)
    int xmask = 1, ymask = 1;
    while (2 * xmask + 1 < c->W - 48 && xmask < 511) xmask = 2 * xmask + 1;
    while (2 * ymask + 1 < c->H - 24 && ymask < 511) ymask = 2 * ymask + 1;
    k_fetch_peak<<<blocks, threads, 0, c->stream>>>(tex, xmask, ymask, layers, 64, sink);        // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; rep++) {
        CU(cudaEventRecord(c->ev0, c->stream));
        k_fetch_peak<<<blocks, threads, 0, c->stream>>>(tex, xmask, ymask, layers, reps, sink);
        CU(cudaEventRecord(c->ev1, c->stream));
        CU(cudaEventSynchronize(c->ev1));
        CU(cudaGetLastError());
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        const double g = (double)blocks * threads * reps * 5.0 / (ms * 1e6);
        if (g > best) best = g;
    }
    *gfetch_per_s = best;
    return GPM_OK;
}

extern "C" int gpm_get_stats(gpm_ctx* c, unsigned long long stats[8])
{
    if (!c || !stats) return fail(GPM_E_ARG, "gpm_get_stats: null argument");
    DeviceGuard g(c->device);
    CU(cudaMemcpyAsync(stats, c->d_stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    stats[ST_LAUNCH] = c->launches;
    stats[6] = c->collectives;
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
            CU(cudaMalloc(&c->seen, (size_t)c->W * c->H * slots * sizeof(unsigned)));
            c->seen_slots = slots;
        }
        CU(cudaMemsetAsync(c->memo_mask, 0, (size_t)c->W * c->H * sizeof(unsigned), c->stream));
        c->opt_neighbours = value;
    }
    else if (n == "packed") { c->opt_packed = value;  if (value) for (auto& f : c->view_8bit) f = 0; }   // set BEFORE uploading views;            // 0 off (default), 1 auto, 2 force — EXPERIMENTAL, see DESIGN.md §5
    else if (n == "memo") c->opt_memo = value != 0;
    else if (n == "quadperm") c->opt_quadperm = value != 0;
    else if (n == "tma") c->opt_tma = value != 0;
    else if (n == "exchange") c->opt_exchange = value != 0;
    else if (n == "async_upload") c->opt_async_upload = value != 0;
    else if (n == "equal_rounds") c->opt_equal_rounds = value != 0;
    else if (n == "prepass") c->opt_prepass = value != 0;
    else if (n == "fused_warps") c->opt_fused_warps = value;
    else if (n == "shard_async") c->opt_shard_async = value != 0;
    else return fail(GPM_E_ARG, "gpm_set_option: unknown option '" + n + "'");
    return GPM_OK;
}

extern "C" void* gpm_stream(gpm_ctx* c) { return c ? (void*)c->stream : nullptr; }
