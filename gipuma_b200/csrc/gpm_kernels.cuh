// gpm_kernels.cuh — __global__ kernels of the PatchMatch hot path (sm_100a).  See gpm_device.cuh.
#pragma once
#include "gpm_device.cuh"
#include <cuda.h>                 // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <curand_kernel.h>

namespace gpm {

__host__ __device__ inline int tile_floats(const KParams& P) { return P.tile_stride * P.tile_w * (P.color ? 4 : 1); }
__host__ __device__ inline int fixed_smem_floats(const KParams& P) { return (tile_floats(P) + P.V * GPM_VIEWCAM_FLOATS + 3) & ~3; }
__device__ __forceinline__ unsigned long long* block_mbar(const KParams& P, float* smem_base)
{
    return reinterpret_cast<unsigned long long*>(smem_base + fixed_smem_floats(P) + (size_t)P.nwarps * warp_scratch_floats(P.ns_pad, P.V, P.color) + 2);
}
// list of the tile's pixels that have work to do (k_sweep's pre-pass): 512 x uint16 behind the lane table
__device__ __forceinline__ unsigned short* block_active(const KParams& P, float* smem_base)
{
    return reinterpret_cast<unsigned short*>(smem_base + fixed_smem_floats(P) + (size_t)P.nwarps * warp_scratch_floats(P.ns_pad, P.V, P.color) + 4 + GPM_MAX_ROUNDS * 8);
}
__device__ __forceinline__ unsigned char* block_perm(const KParams& P, float* smem_base)
{
    return reinterpret_cast<unsigned char*>(smem_base + fixed_smem_floats(P) + (size_t)P.nwarps * warp_scratch_floats(P.ns_pad, P.V, P.color) + 4);
}

// ---- TMA (cp.async.bulk.tensor) + mbarrier primitives -------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned done = 0;
    for (int spin = 0; !done; spin++) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spin > (1 << 22)) __trap();          // a lost TMA transaction must fail loudly, not hang the GPU
    }
}

// Stage the (32+2*halo)^2 reference window of tile (bx, by) and the per-view camera table in shared memory.
// Unlike the reference's loader (gipuma.cu:1510-1525, skipped by threads that left early — SURVEY.md §7),
// the whole window is always defined.  With P.use_tma one elected thread issues a single 2-D TMA box copy
// (cp.async.bulk.tensor) from the replicate-padded reference image and the block waits on its mbarrier while the other
// threads copy the camera and lane tables; otherwise the window is a cooperative copy.  `phase` = parity of the barrier's
// current phase (0 for the first / only tile of a block).
__device__ __forceinline__ void stage_block(const KParams& P, const CUtensorMap* tmap, const float* __restrict__ refpad,
                                            const ViewCam* __restrict__ cams, float* tile, float* sCam,
                                            int tile_x0, int tile_y0, unsigned phase = 0u)
{
    const int tw = P.tile_w, ts = P.tile_stride;
    unsigned long long* bar = block_mbar(P, tile);
    if (P.use_tma) {
        if (threadIdx.x == 0) {
            if (phase == 0u) mbar_init(bar, 1);
            mbar_arrive_expect_tx(bar, (unsigned)(tile_floats(P) * sizeof(float)));
            tma_load_2d(tile, tmap, (tile_x0 + GPM_APRON - P.tile_xo) * (P.color ? 4 : 1), tile_y0 + GPM_APRON, bar);
        }
    } else if (P.color) {                           // float4 texels: refpitch counts float4 elements
        const float4* src4 = reinterpret_cast<const float4*>(refpad) + (size_t)(tile_y0 + GPM_APRON) * P.refpitch + (tile_x0 + GPM_APRON);
        float4* tile4 = reinterpret_cast<float4*>(tile);
        for (int e = threadIdx.x; e < tw * tw; e += blockDim.x) {
            const int J = e / tw, I = e - J * tw;
            tile4[J * ts + I] = src4[(size_t)J * P.refpitch + I];
        }
    } else {
        const float* src = refpad + (size_t)(tile_y0 + GPM_APRON) * P.refpitch + (tile_x0 + GPM_APRON);
        for (int e = threadIdx.x; e < tw * tw; e += blockDim.x) {
            const int J = e / tw, I = e - J * tw;
            tile[J * ts + P.tile_xo + I] = src[(size_t)J * P.refpitch + I];
        }
    }
    const float* c = reinterpret_cast<const float*>(cams);
    for (int e = threadIdx.x; e < P.V * GPM_VIEWCAM_FLOATS; e += blockDim.x) sCam[e] = c[e];
    // lane -> sample table of the sampling rounds (KParams::perm), word by word
    unsigned* sp = reinterpret_cast<unsigned*>(block_perm(P, tile));
    const unsigned* gp = reinterpret_cast<const unsigned*>(&P.perm[0][0]);
    for (int e = threadIdx.x; e < GPM_MAX_ROUNDS * 8; e += blockDim.x) sp[e] = gp[e];
    __syncthreads();                                // tables visible; mbarrier initialised before anyone waits on it
    if (P.use_tma) mbar_wait(bar, phase & 1u);
}

__device__ __forceinline__ void flush_stats(unsigned long long* stats, const WarpStats& st, unsigned lane)
{
    if (lane == 0 && stats) {
        atomicAdd(stats + ST_HYP, (unsigned long long)st.hyp);
        atomicAdd(stats + ST_SKIP, (unsigned long long)st.skip);
        atomicAdd(stats + ST_PRUNED, (unsigned long long)st.pruned);
        atomicAdd(stats + ST_PAIRS, st.pairs);
        atomicAdd(stats + ST_PAIRS_FULL, st.pairs_full);
    }
}

// shared memory: [tile tile_stride*tile_w][cams V*21][pad to 16 B][nwarps * warp_scratch][4 ints: work counter, active count, mbarrier (8 B)][lane permutation table, GPM_MAX_ROUNDS x 32 bytes][active pixel list, 512 x uint16]
__host__ __device__ inline size_t block_smem_bytes(const KParams& P)
{
    size_t fl = (size_t)fixed_smem_floats(P) + (size_t)P.nwarps * warp_scratch_floats(P.ns_pad, P.V, P.color) + 4 + GPM_MAX_ROUNDS * 8 + GPM_TILE * GPM_TILE / 4;
    return fl * sizeof(float);
}

// ---- exact memo of rejected work: plane identities ------------------------------------------------------------
// Every plane carries a 32-bit identity: pid[p] = p + 1 at initialisation, a fresh number (atomic counter) whenever a
// refinement creates a new plane, and COPIED when a plane propagates to a neighbour (gipuma.cu:867-871 copies the plane
// verbatim).  Equal identity therefore implies bit-identical planes, so "this neighbour still holds the plane I was offered
// before" is one 4-byte compare instead of a 16-byte one, the per-pixel memo shrinks from 8 x 16 + 16 to 8 x 4 + 4 bytes,
// and a neighbour whose identity is remembered is skipped without even loading its plane.  (Planes that are equal by value
// but were created independently have different identities: a skip is lost, never a decision.)
struct Memo {
    unsigned* pid;        // [H*W]       identity of the stored plane
    unsigned* seen;       // [H*W*NC]    identity last offered to each pixel from each propagation direction
    unsigned* refseen;    // [H*W]       identity of the plane from which the last all-rejected refinement started
    unsigned* mask;       // [H*W]       validity bits of seen (0..19) and refseen (GPM_MEMO_REFINE)
    unsigned* next_id;    // device counter of fresh identities
};

// ---- random plane initialisation — gipuma_init_cu2, gipuma.cu:996-1036 ----------------------
__global__ void k_fill_ids(unsigned* __restrict__ pid, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pid[i] = i + 1u;
}

__global__ void k_init_planes(const __grid_constant__ KParams P, unsigned long long seed, float4* __restrict__ planes,
                              unsigned* __restrict__ rng_state, unsigned* __restrict__ pid)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.W || y >= P.H) return;
    curandState st;
    curand_init(seed, y, x, &st);                                    // :1019 (reference seed: clock64())
    Xorwow r;
    r.v0 = st.v[0];  r.v1 = st.v[1];  r.v2 = st.v[2];  r.v3 = st.v[3];  r.v4 = st.v[4];  r.d = st.d;
    const float px = __int2float_rn(x), py = __int2float_rn(y);
    float vx, vy, vz;
    view_vector(P.ref, px, py, vx, vy, vz);                         // :1025
    // disp_now = curand_between(mind, maxd)                           :1028
    const float disp = ffma(fsub(P.max_disp, P.min_disp), xorwow_uniform(r), P.min_disp);
    // rndUnitVectorSphereMarsaglia_cu, :148-164
    float a, b, sum;
    do {
        a = ffma(xorwow_uniform(r), 2.0f, -1.0f);
        b = ffma(xorwow_uniform(r), 2.0f, -1.0f);
        sum = ffma(a, a, fmul(b, b));
    } while (sum >= 1.0f);
    const float sq = fsqrt_(fsub(1.0f, sum));
    float nx = fmul(fadd(a, a), sq), ny = fmul(fadd(b, b), sq), nz = fsub(1.0f, fadd(sum, sum));
    if (dot3(nx, ny, nz, vx, vy, vz) > 0.0f) { nx = -nx;  ny = -ny;  nz = -nz; }   // vecOnHemisphere_cu :131-137
    // depth = f * baseline / disp (:1031), d = getD_cu (:1034)
    const float depth = fmul(fmul(P.ref.f_cam, P.ref.baseline), frcp(disp));
    const float d = plane_d(P.ref, nx, ny, nz, px, py, depth);
    planes[(size_t)y * P.W + x] = make_float4(nx, ny, nz, d);
    pid[(size_t)y * P.W + x] = (unsigned)(y * P.W + x) + 1u;
    if (rng_state) {
        unsigned* o = rng_state + ((size_t)y * P.W + x) * 6;
        o[0] = r.v0;  o[1] = r.v1;  o[2] = r.v2;  o[3] = r.v3;  o[4] = r.v4;  o[5] = r.d;
    }
}

// ---- cost of the stored (or supplied) plane at every pixel — gipuma.cu:1040-1049 and gpm_cost_eval ----
#ifndef GPM_LB_THREADS
#define GPM_LB_THREADS 512       // max threads per block of the warp-per-pixel kernels (16 warps)
#define GPM_LB_BLOCKS 1
#endif
template <bool PACKED, bool COLOR>
__global__ void __launch_bounds__(GPM_LB_THREADS, GPM_LB_BLOCKS)
k_cost_eval(const __grid_constant__ KParams P, const __grid_constant__ CUtensorMap tmap, const ViewCam* __restrict__ cams, const float* __restrict__ refpad,
            cudaTextureObject_t src, cudaTextureObject_t grad, const float4* __restrict__ planes, float* __restrict__ cost,
            unsigned long long* __restrict__ stats)
{
    extern __shared__ __align__(128) float smem[];
    float* tile = smem;
    float* sCam = tile + tile_floats(P);
    float* scratch = smem + fixed_smem_floats(P);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_x0 = blockIdx.x * GPM_TILE - P.halo, tile_y0 = blockIdx.y * GPM_TILE - P.halo;
    stage_block(P, &tmap, refpad, cams, tile, sCam, tile_x0, tile_y0);
    const WarpScratch ws = carve(scratch + (size_t)warp * warp_scratch_floats(P.ns_pad, P.V, P.color), P.ns_pad, P.V, P.color, block_perm(P, smem));
    WarpStats st = {0, 0, 0, 0, 0};
    for (int idx = warp; idx < GPM_TILE * GPM_TILE; idx += P.nwarps) {
        const int px = blockIdx.x * GPM_TILE + (idx & 31), py = blockIdx.y * GPM_TILE + (idx >> 5);
        if (px >= P.W || py >= P.H) continue;
        setup_window<COLOR>(P, tile + P.tile_xo, ws, px, py, tile_x0, tile_y0, lane);
        const float4 n = planes[(size_t)py * P.W + px];
        const float inf = __int_as_float(0x7f800000);
        const float c = (P.cost_rt & 12) ? eval_plane<2, PACKED, COLOR>(P, sCam, ws, src, grad, n.x, n.y, n.z, n.w, inf, lane, st, nullptr, nullptr, (unsigned)P.cost_rt)
                        : P.cost_variant ? eval_plane<1, PACKED, COLOR>(P, sCam, ws, src, grad, n.x, n.y, n.z, n.w, inf, lane, st)
                                         : eval_plane<0, PACKED, COLOR>(P, sCam, ws, src, grad, n.x, n.y, n.z, n.w, inf, lane, st);
        if (lane == 0) cost[(size_t)py * P.W + px] = c;
    }
    flush_stats(stats, st, lane);
}

// ---- one checkerboard colour: close + far propagation + refinement, fused --------------------
// gipuma_{black,red}_spatialPropClose_cu / spatialPropFar_cu / planeRefine_cu, gipuma.cu:1353-1823.
// colour 0 = black, 1 = red; phase_mask bit0 close, bit1 far, bit2 refine.
// Candidate table of the fused 20-neighbour kernel gipuma_checkerboard_cu (gipuma.cu:1236-1330): offset of candidate k
// and the reference's own border guards, written as margins  x >= gx0, x <= W-1-gx1, y >= gy0, y <= H-1-gy1
// (they are not the tightest ones: e.g. `left - cols*2` is guarded by p.y > 2, gipuma.cu:1313-1317).
__device__ const signed char kFusedDx[20]  = { 0,  0,  0, 0, 0, 0, -1, -3, -5, 1, 3, 5,  2, 2, -2, -2, -1,  1, -1, 1};
__device__ const signed char kFusedDy[20]  = {-1, -3, -5, 1, 3, 5,  0,  0,  0, 0, 0, 0, -1, 1, -1,  1, -2, -2,  2, 2};
__device__ const signed char kFusedGx0[20] = { 0,  0,  0, 0, 0, 0,  1,  3,  5, 0, 0, 0,  0, 0,  2,  2,  1,  0,  1, 0};
__device__ const signed char kFusedGx1[20] = { 0,  0,  0, 0, 0, 0,  0,  0,  0, 1, 3, 5,  2, 2,  0,  0,  0,  1,  0, 1};
__device__ const signed char kFusedGy0[20] = { 1,  3,  5, 0, 0, 0,  0,  0,  0, 0, 0, 0,  1, 0,  1,  0,  3,  3,  0, 0};
__device__ const signed char kFusedGy1[20] = { 0,  0,  0, 1, 3, 5,  0,  0,  0, 0, 0, 0,  0, 1,  0,  1,  0,  0,  2, 2};

template <bool PACKED, bool COLOR, bool FUSED>
__global__ void __launch_bounds__(GPM_LB_THREADS, GPM_LB_BLOCKS)
k_sweep(const __grid_constant__ KParams P, const __grid_constant__ CUtensorMap tmap, const ViewCam* __restrict__ cams, const float* __restrict__ refpad,
        cudaTextureObject_t src, cudaTextureObject_t grad, float4* __restrict__ planes, float* __restrict__ cost,
        unsigned* __restrict__ rng_state, unsigned char* __restrict__ prov, Memo M, int colour, int phase_mask,
        unsigned long long* __restrict__ stats)
{
    extern __shared__ __align__(128) float smem[];
    float* tile = smem;
    float* sCam = tile + tile_floats(P);
    float* scratch = smem + fixed_smem_floats(P);
    int* counter = reinterpret_cast<int*>(scratch + (size_t)P.nwarps * warp_scratch_floats(P.ns_pad, P.V, P.color));
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_x0 = blockIdx.x * GPM_TILE - P.halo, tile_y0 = blockIdx.y * GPM_TILE - P.halo;
    if (threadIdx.x == 0) { counter[0] = P.nwarps;  counter[1] = 0; }
    stage_block(P, &tmap, refpad, cams, tile, sCam, tile_x0, tile_y0);
    const WarpScratch ws = carve(scratch + (size_t)warp * warp_scratch_floats(P.ns_pad, P.V, P.color), P.ns_pad, P.V, P.color, block_perm(P, smem));
    const RefCam& cam = P.ref;
    WarpStats st = {0, 0, 0, 0, 0};
    const int W = P.W, H = P.H;
    constexpr int NC = FUSED ? 20 : 8;

    // gridDim.z slices of the tile's 512 pixels of this colour (small images: more, smaller work units -> no idle SMs in
    // the last wave); inside a slice the warps take pixels dynamically
    const int per_slice = (GPM_TILE * GPM_TILE / 2 + (int)gridDim.z - 1) / (int)gridDim.z;
    const int idx_begin = (int)blockIdx.z * per_slice;
    const int idx_end = min(GPM_TILE * GPM_TILE / 2, idx_begin + per_slice);

    // Pre-pass, one THREAD per pixel: which pixels have anything to do?  A pixel whose every candidate is a memo hit (the
    // neighbour still carries the identity it was offered before) and whose refinement is known to reject every step needs no
    // warp at all — the warp path would read its state, skip everything and write the same state back.  In the late
    // iterations that is most pixels; the warps then only walk the compact list of the others.  (Other exact shortcuts —
    // history, duplicates, depth range — still run in the warp path: a pixel they would clear is merely listed.)
    unsigned short* act = block_active(P, smem);
    if (P.prepass) {
        unsigned nskip = 0;
        for (int i = idx_begin + (int)threadIdx.x; i < idx_end; i += (int)blockDim.x) {
            const int tx = i & 31, ty = i >> 5;
            const int px = blockIdx.x * GPM_TILE + tx;
            const int py = blockIdx.y * GPM_TILE + 2 * ty + (((tx & 1) ^ colour) & 1);
            if (px >= W || py >= H) continue;
            bool idle = false;
            unsigned ncand = 0;
            if (P.memo) {
                const size_t center = (size_t)py * W + px;
                const unsigned mmask = M.mask[center];
                bool all_hit = true;
                for (int k = 0; k < NC; k++) {
                    bool ok;
                    size_t at;
                    if (FUSED) {
                        ok = (phase_mask & 1) && px >= kFusedGx0[k] && px <= W - 1 - kFusedGx1[k] && py >= kFusedGy0[k] && py <= H - 1 - kFusedGy1[k];
                        at = (size_t)(py + kFusedDy[k]) * W + (px + kFusedDx[k]);
                    } else {
                        const int dist = k < 4 ? 1 : 5, dir = k & 3;
                        int qx = px, qy = py;
                        if (dir == 0)      { ok = py > dist - 1;  qy = py - dist; }
                        else if (dir == 1) { ok = py < H - dist;  qy = py + dist; }
                        else if (dir == 2) { ok = px > dist - 1;  qx = px - dist; }
                        else               { ok = px < W - dist;  qx = px + dist; }
                        ok = ok && ((phase_mask >> (k >> 2)) & 1);
                        at = (size_t)qy * W + qx;
                    }
                    if (!ok) continue;
                    ncand++;
                    if (!((mmask >> k) & 1) || M.seen[center * NC + k] != M.pid[at]) { all_hit = false;  break; }
                }
                const bool refine = (phase_mask & 4) != 0;
                const bool refine_idle = !refine || (P.rng_mode == 0 && (mmask & GPM_MEMO_REFINE) && M.refseen[center] == M.pid[center]);
                idle = all_hit && refine_idle;
                if (idle) nskip += ncand + (refine ? 3u : 0u);
            }
            if (!idle) act[atomicAdd(&counter[1], 1)] = (unsigned short)i;
        }
        nskip = __reduce_add_sync(GPM_FULL, nskip);
        st.skip += nskip;                       // every lane holds the warp's sum; flush_stats takes lane 0's
    }
    __syncthreads();
    const int nactive = P.prepass ? counter[1] : idx_end - idx_begin;      // without the pre-pass: every pixel of the slice, in order
    int pos = (int)warp;
    while (pos < nactive) {
        const int idx = P.prepass ? (int)act[pos] : idx_begin + pos;
        // pixel of this colour — gipuma.cu:1730-1734 / 1786-1790
        const int tx = idx & 31, ty = idx >> 5;
        const int px = blockIdx.x * GPM_TILE + tx;
        const int py = blockIdx.y * GPM_TILE + 2 * ty + (((tx & 1) ^ colour) & 1);
        if (px < W && py < H) {
            const size_t center = (size_t)py * W + px;
            bool window_ready = false;          // the per-pixel window data is only built if something gets evaluated
            const float fpx = __int2float_rn(px), fpy = __int2float_rn(py);
            float4 norm_now = planes[center];
            float cost_now = cost[center];
            float disp_now = plane_depth(cam, norm_now.x, norm_now.y, norm_now.z, norm_now.w, fpx, fpy);   // :1530
            // which rounding variant of the cost function produced cost_now: eval_plane's rt (0 = y-first, 1 = x-first, ...), 0xff = unknown
            int prov_now = prov[center];

            // candidate k lives in lane k: 0..3 = up, down, left, right at 1 px (:1571-1582), 4..7 at 5 px (:1450-1462);
            // fused kernel: the 20 candidates of gipuma.cu:1236-1330 in source order
            constexpr unsigned NCMASK = (1u << NC) - 1u;
            float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
            bool mine_ok = false;
            size_t mine_at = center;
            if (FUSED) {
                if (lane < NC) {
                    mine_ok = (phase_mask & 1) && px >= kFusedGx0[lane] && px <= W - 1 - kFusedGx1[lane] &&
                              py >= kFusedGy0[lane] && py <= H - 1 - kFusedGy1[lane];
                    mine_at = (size_t)(py + kFusedDy[lane]) * W + (px + kFusedDx[lane]);
                }
            } else if (lane < 8) {
                const int dist = lane < 4 ? 1 : 5;
                const int dir = lane & 3;
                const bool phase_on = (phase_mask >> (lane >> 2)) & 1;
                int qx = px, qy = py;
                bool ok;
                if (dir == 0)      { ok = py > dist - 1;      qy = py - dist; }
                else if (dir == 1) { ok = py < H - dist;      qy = py + dist; }
                else if (dir == 2) { ok = px > dist - 1;      qx = px - dist; }
                else               { ok = px < W - dist;      qx = px + dist; }
                mine_ok = ok && phase_on;
                mine_at = (size_t)qy * W + qx;
            }
            const unsigned mine_id = mine_ok ? M.pid[mine_at] : 0u;
            unsigned own_id = M.pid[center];
            // Memo of rejected work.  A pixel's cost only ever decreases (gipuma.cu:867,986), and cost(p, plane) is a pure
            // function, so a neighbour plane this pixel has already been offered — accepted or not — can never be
            // accepted later: seen[p][k] holds the last plane offered from direction k; an unchanged neighbour is skipped
            // without evaluation, in every later iteration.  Exact, not a heuristic.
            // the cost function of the fused kernel differs between its inlined call sites (P.site): the duplicate / memo
            // shortcuts below only pair candidates of sites with the same rounding variant
            const unsigned my_site = FUSED ? (unsigned)P.site[lane < NC ? lane : 0] : (COLOR ? 0u : 1u);
            unsigned mmask = P.memo ? M.mask[center] : 0u;
            const bool old_ok = P.memo && lane < NC && ((mmask >> lane) & 1);
            const unsigned old_id = old_ok ? M.seen[center * NC + lane] : 0u;
            const bool memo_hit = old_ok && mine_ok && old_id == mine_id;
            if (mine_ok && !memo_hit) mine = planes[mine_at];          // a remembered neighbour is skipped below without its plane
            const unsigned cand_mask = __ballot_sync(GPM_FULL, mine_ok);
            for (int k = 0; k < NC; k++) {
                if (!((cand_mask >> k) & 1)) continue;
                // rounding variant of this call site: x-first for float images, y-first for float4; the fused
                // kernel's inlined sites are looked up (bit 0 = x-first, bit 1 = gradient folding)
                const unsigned site = FUSED ? (unsigned)P.site[k] : (COLOR ? 0u : 1u);
                float4 nb;
                nb.x = __shfl_sync(GPM_FULL, mine.x, k);  nb.y = __shfl_sync(GPM_FULL, mine.y, k);
                nb.z = __shfl_sync(GPM_FULL, mine.z, k);  nb.w = __shfl_sync(GPM_FULL, mine.w, k);
                const unsigned cand_id = __shfl_sync(GPM_FULL, mine_id, k);
                // already offered to this pixel before, from ANY direction (the 8 memo entries double as a history)?
                const bool known = old_ok && my_site == site && cand_id == old_id;
                if (__any_sync(GPM_FULL, known)) { st.skip++; continue; }
                // spatialPropagation_cu, gipuma.cu:832-874
                const float disp_before = plane_depth(cam, nb.x, nb.y, nb.z, nb.w, fpx, fpy);
                const bool in_range = disp_before >= cam.depthMin && disp_before <= cam.depthMax;   // :829-830, :865
                // exact duplicates: cost(p, plane) is a pure function, so a plane equal to the current one or to an
                // earlier candidate of this pixel cannot be accepted (`cost_before < *cost_now` is false)
                const bool same_now = P.dedupe_self && prov_now == (int)site && cand_id == own_id;
                const bool same_mine = P.dedupe_cand && mine_ok && (int)lane < k && my_site == site && cand_id == mine_id;
                const bool dup = __any_sync(GPM_FULL, same_mine);
                if (!in_range || same_now || dup) { st.skip++; continue; }
                if (!window_ready) { setup_window<COLOR>(P, tile + P.tile_xo, ws, px, py, tile_x0, tile_y0, lane);  window_ready = true; }
                const float c = FUSED ? eval_plane<2, PACKED, COLOR>(P, sCam, ws, src, grad, nb.x, nb.y, nb.z, nb.w, cost_now, lane, st, nullptr, nullptr, site)
                                      : eval_plane<(COLOR ? 0 : 1), PACKED, COLOR>(P, sCam, ws, src, grad, nb.x, nb.y, nb.z, nb.w, cost_now, lane, st);
                if (c < cost_now) {                                                              // :867-871
                    disp_now = disp_before;
                    norm_now = nb;
                    cost_now = c;
                    prov_now = (int)site;
                    own_id = cand_id;                                                            // a copy keeps its identity
                }
            }

            if (P.memo && mine_ok && !memo_hit) {            // every candidate offered above is now known to this pixel
                M.seen[center * NC + lane] = mine_id;
            }
            unsigned new_mask = mmask | (cand_mask & NCMASK);
            // The refinement candidates are a pure function of (pixel, plane at refinement start) in GPM_RNG_REFERENCE
            // mode (zero-state stream; disp_now == plane_depth(norm_now) here), so a refinement that rejected all its steps
            // from exactly this plane before would reject them again.
            bool refine = (phase_mask & 4) != 0;
            if (refine && P.memo && P.rng_mode == 0 && (mmask & GPM_MEMO_REFINE)) {
                if (M.refseen[center] == own_id) {
                    refine = false;
                    st.skip += 3;
                }
            }
            if (refine) {
                if (!window_ready) { setup_window<COLOR>(P, tile + P.tile_xo, ws, px, py, tile_x0, tile_y0, lane);  window_ready = true; }
                const unsigned id_start = own_id;
                bool any_accept = false;
                // planeRefinement_cu, gipuma.cu:928-994 with getRndDispAndUnitVector_cu, :890-927
                Xorwow r = {0u, 0u, 0u, 0u, 0u, 0u};       // GPM_RNG_REFERENCE: gs.cs is never written (gipuma.cu:1840,1608)
                if (P.rng_mode == 1) {
                    const unsigned* in = rng_state + center * 6;
                    r.v0 = in[0];  r.v1 = in[1];  r.v2 = in[2];  r.v3 = in[3];  r.v4 = in[4];  r.d = in[5];
                }
                float vx, vy, vz;
                view_vector(cam, fpx, fpy, vx, vy, vz);                                          // :948
                const float fb = fmul(cam.baseline, cam.f);
                float deltaN = 1.0f;
                for (float deltaZ = fmul(P.max_disp, 0.5f); deltaZ >= 0.01f; deltaZ = fmul(deltaZ, 0.1f)) {   // :958-959
                    const float rdisp = frcp(disp_now);                   // disp = f*baseline/depth (:904), kept fused
                    const float hi = fmin_(ffma(rdisp, -fb, P.max_disp), deltaZ);        // maxDelta (:910)
                    const float lo = fmin_(ffma(rdisp, fb, P.min_disp), deltaZ);         // -minDelta (:909)
                    const float dz = ffma(xorwow_uniform(r), fadd(lo, hi), -lo);        // curand_between(minDelta, maxDelta) :914
                    float dnew = ffma(rdisp, fb, dz);                                    // disp + deltaZ
                    dnew = fmin_(P.max_disp, fmax_(P.min_disp, dnew));                   // :916
                    const float depth_new = fmul(frcp(dnew), fb);                        // :918
                    const float twoN = fadd(deltaN, deltaN);
                    const float ax = fadd(norm_now.x, ffma(xorwow_uniform(r), twoN, -deltaN));   // :921-923
                    const float ay = fadd(norm_now.y, ffma(twoN, xorwow_uniform(r), -deltaN));
                    const float az = fadd(norm_now.z, ffma(twoN, xorwow_uniform(r), -deltaN));
                    const float rs = frsq(ffma(az, az, ffma(ax, ax, fmul(ay, ay))));     // normalize_cu :113-120
                    float4 cand;
                    cand.x = fmul(ax, rs);  cand.y = fmul(ay, rs);  cand.z = fmul(az, rs);
                    if (dot3(cand.x, cand.y, cand.z, vx, vy, vz) > 0.0f) { cand.x = -cand.x;  cand.y = -cand.y;  cand.z = -cand.z; }
                    cand.w = plane_d(cam, cand.x, cand.y, cand.z, fpx, fpy, depth_new);   // :969
                    const unsigned rsite = FUSED ? (unsigned)P.site[20] : 0u;
                    const float c = FUSED ? eval_plane<2, PACKED, COLOR>(P, sCam, ws, src, grad, cand.x, cand.y, cand.z, cand.w, cost_now, lane, st, nullptr, nullptr, rsite)
                                          : eval_plane<0, PACKED, COLOR>(P, sCam, ws, src, grad, cand.x, cand.y, cand.z, cand.w, cost_now, lane, st);
                    if (c < cost_now) {                                                          // :986-990 (no depth-range test)
                        prov_now = (int)rsite;
                        any_accept = true;
                        cost_now = c;
                        disp_now = depth_new;
                        norm_now = cand;
                    }
                    deltaN = fmul(deltaN, 0.25f);                                                // :992
                }
                if (P.rng_mode == 1 && lane == 0) {
                    unsigned* o = rng_state + center * 6;
                    o[0] = r.v0;  o[1] = r.v1;  o[2] = r.v2;  o[3] = r.v3;  o[4] = r.v4;  o[5] = r.d;
                }
                if (any_accept) {                                 // a new plane: one fresh identity for the pixel's final plane
                    unsigned fresh = 0u;
                    if (lane == 0) fresh = atomicAdd(M.next_id, 1u);
                    own_id = __shfl_sync(GPM_FULL, fresh, 0);
                }
                if (P.memo && P.rng_mode == 0) {
                    if (!any_accept) { if (lane == 0) M.refseen[center] = id_start;  new_mask |= GPM_MEMO_REFINE; }
                    else new_mask &= ~GPM_MEMO_REFINE;
                }
            }
            if (P.memo && lane == 0 && new_mask != mmask) M.mask[center] = new_mask;
            if (lane == 0) {                                                                     // :1585-1587
                cost[center] = cost_now;
                planes[center] = norm_now;
                prov[center] = (unsigned char)prov_now;
                M.pid[center] = own_id;
            }
        }
        if (lane == 0) pos = atomicAdd(&counter[0], 1);
        pos = __shfl_sync(GPM_FULL, pos, 0);
    }
    flush_stats(stats, st, lane);
}

// ============================================================================================================
// Source-view sharding across GPUs (BASELINE configs 4 / 5, north_star's strong-scaling axis).
// Every rank holds the full plane/cost state and a SUBSET of the source views.  Per-view costs are independent
// (gipuma.cu:742-778); only the combination needs all views.  For COMB_BEST_N the n_best smallest costs over all views
// are the n_best smallest of the union of every rank's n_best smallest, so each rank exports, per pixel and hypothesis
// slot, its ascending local top-n_best; the lists of all ranks are exchanged (NCCL all-gather, or peer stores over NVLink
// in the fused kernel below) and every rank merges them, sums in ascending order — the reference's order
// (gipuma.cu:779-797) — and applies the accept logic redundantly.  All ranks therefore keep bit-identical state,
// identical to a single-GPU run over all views.
//   stage 0        initial cost of the stored planes        (1 slot,  radius box/2, both colours)
//   stage 1        close + far propagation candidates       (8 slots)
//   stage 2 + s    refinement step s                         (1 slot; sequential: step s+1 depends on step s' accept)
// The accept of stage s is FUSED into the evaluation of stage s+1 (shard_stage_pixel): a colour costs 1 + S evaluation
// passes plus one closing accept instead of 2 (1 + S) launches.  The exact shortcuts of k_sweep that depend only on the
// shared state are kept — direction memo, history, duplicates (incl. the current plane), refinement memo — so every rank
// skips the same work and the skipped slots simply carry +inf lists.
// Exchange index of pixel (x, y): y * ceil(W/2) + x/2 (pixels of one colour have distinct x/2 within a row).
// ============================================================================================================
#define GPM_SF_SKIP_REFINE 1u     // sflags: the refinement of this colour pass is known to reject every step (memo)
#define GPM_SF_ACCEPTED    2u     // sflags: some refinement step of this colour pass was accepted

struct ShardState {
    float* dispbuf;               // disp_now carried between the stages of one colour
    float4* candbuf;              // refinement candidate of the step being exchanged
    float* canddepth;
    unsigned char* sflags;
};

// Peer-memory exchange (k_shard_fused): every rank owns one exchange region = [2 parities][world source ranks][slot_floats]
// lists + [world][nblocks] arrival flags; all regions are mapped into every rank's address space (CUDA IPC over NVLink).
// A rank writes its lists straight into slot `me` of every peer's region and then raises flag [me][block] there.
struct ShardP2P {
    float* lists[8];              // list area of rank r's region as mapped here (lists[me] is local memory)
    unsigned* flags[8];           // flag area of rank r's region
    unsigned* err;                // local sticky error word (a peer did not arrive in time)
    unsigned long long slot_floats;
    int me, world, nblocks;
};

// merged cost of one slot: the n_best smallest over all ranks' lists, summed ascending, / count  (gipuma.cu:779-803)
__device__ __forceinline__ float shard_merge(const float* __restrict__ g, size_t rank_stride, int world, int nb)
{
    int pos[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float sum = 0.0f;
    int taken = 0;
    for (int i = 0; i < nb; i++) {
        float best = __int_as_float(0x7f800000);
        int br = -1;
        for (int r = 0; r < world; r++) {
            if (pos[r] < nb) { const float v = __ldcg(g + r * rank_stride + pos[r]); if (v < best) { best = v;  br = r; } }
        }
        if (br < 0 || !(best < GPM_MAXCOST)) break;        // only views with c < MAXCOST count (numValidViews)
        pos[br]++;
        sum = fadd(sum, best);
        taken++;
    }
    float cost = fmul(frcp(__int2float_rn(taken)), sum);
    if (taken < 1) cost = GPM_MAXCOST;
    if (cost != cost || cost > GPM_MAXCOST || cost < 0.0f) cost = GPM_MAXCOST;
    return cost;
}

__device__ __forceinline__ bool same_bits(const float4& a, const float4& b)
{
    return __float_as_uint(a.x) == __float_as_uint(b.x) && __float_as_uint(a.y) == __float_as_uint(b.y) &&
           __float_as_uint(a.z) == __float_as_uint(b.z) && __float_as_uint(a.w) == __float_as_uint(b.w);
}

// refinement candidate of step `sidx` from (norm_now, disp_now) — getRndDispAndUnitVector_cu, gipuma.cu:890-927, with the
// reference's zero-state stream (pin P2) advanced by 4 draws per earlier step
__device__ __forceinline__ float4 shard_refine_candidate(const KParams& P, const float4& norm_now, float disp_now, float fpx, float fpy,
                                                         int sidx, float& depth_new)
{
    const RefCam& cam = P.ref;
    Xorwow r = {0u, 0u, 0u, 0u, 0u, 0u};
    float deltaZ = fmul(P.max_disp, 0.5f), deltaN = 1.0f;
    for (int i = 0; i < sidx; i++) { for (int j = 0; j < 4; j++) xorwow_next(r);  deltaZ = fmul(deltaZ, 0.1f);  deltaN = fmul(deltaN, 0.25f); }
    float vx, vy, vz;
    view_vector(cam, fpx, fpy, vx, vy, vz);
    const float fb = fmul(cam.baseline, cam.f);
    const float rdisp = frcp(disp_now);
    const float hi = fmin_(ffma(rdisp, -fb, P.max_disp), deltaZ);
    const float lo = fmin_(ffma(rdisp, fb, P.min_disp), deltaZ);
    const float dz = ffma(xorwow_uniform(r), fadd(lo, hi), -lo);
    float dnew = ffma(rdisp, fb, dz);
    dnew = fmin_(P.max_disp, fmax_(P.min_disp, dnew));
    depth_new = fmul(frcp(dnew), fb);
    const float twoN = fadd(deltaN, deltaN);
    const float ax = fadd(norm_now.x, ffma(xorwow_uniform(r), twoN, -deltaN));
    const float ay = fadd(norm_now.y, ffma(twoN, xorwow_uniform(r), -deltaN));
    const float az = fadd(norm_now.z, ffma(twoN, xorwow_uniform(r), -deltaN));
    const float rs = frsq(ffma(az, az, ffma(ax, ax, fmul(ay, ay))));
    float4 cand;
    cand.x = fmul(ax, rs);  cand.y = fmul(ay, rs);  cand.z = fmul(az, rs);
    if (dot3(cand.x, cand.y, cand.z, vx, vy, vz) > 0.0f) { cand.x = -cand.x;  cand.y = -cand.y;  cand.z = -cand.z; }
    cand.w = plane_d(cam, cand.x, cand.y, cand.z, fpx, fpy, depth_new);
    return cand;
}

// Accept of stage `stage` at pixel (px, py) from the exchanged lists (one warp; every lane returns the same state).
// stage 1: the 8 propagation slots in the reference's order (gipuma.cu:1571-1582, 1450-1462; accept rule :867-871);
// stage >= 2: one refinement step (:986-990).
__device__ __forceinline__ void shard_accept_pixel(const KParams& P, const float4* __restrict__ planes, const ShardState& S, const Memo& M,
                                                   int px, int py, int stage, const float* __restrict__ gathered, int world, size_t rank_stride,
                                                   float4& norm_now, float& cost_now, float& disp_now, int& prov_now, unsigned& sf,
                                                   unsigned& own_id, unsigned lane)
{
    const int W = P.W, H = P.H, Wh = (W + 1) >> 1, nb = P.n_best;
    const size_t center = (size_t)py * W + px;
    // list layout: NCCL flow — dense per stage (8 slots per pixel in stage 1, 1 otherwise); fused flow (rank_stride != 0) — a
    // fixed 8 slots per pixel in every stage, so that a pixel's lists only ever alias lists of the same pixel, i.e. of the
    // same block (blocks of one launch run stages apart from each other; only same-index blocks are ordered by the flags)
    const int slots = (stage == 1 || rank_stride) ? 8 : 1;
    const size_t per_rank = rank_stride ? rank_stride : (size_t)H * Wh * slots * nb;
    const float* g = gathered + ((size_t)py * Wh + (px >> 1)) * slots * nb;
    const float fpx = __int2float_rn(px), fpy = __int2float_rn(py);
    if (stage == 1) {
        float mine = GPM_MAXCOST;
        if (lane < 8) mine = shard_merge(g + lane * nb, per_rank, world, nb);
        for (int k = 0; k < 8; k++) {
            const float c = __shfl_sync(GPM_FULL, mine, k);
            if (c < cost_now) {
                const int dist = k < 4 ? 1 : 5, dir = k & 3;
                const int qx = px + (dir == 2 ? -dist : dir == 3 ? dist : 0), qy = py + (dir == 0 ? -dist : dir == 1 ? dist : 0);
                norm_now = planes[(size_t)qy * W + qx];                 // other colour: not written during this colour's stages
                own_id = M.pid[(size_t)qy * W + qx];                    // a copy keeps its identity
                disp_now = plane_depth(P.ref, norm_now.x, norm_now.y, norm_now.z, norm_now.w, fpx, fpy);
                cost_now = c;
                prov_now = P.color ? 0 : 1;
            }
        }
    } else if (!(sf & GPM_SF_SKIP_REFINE)) {
        const float c = shard_merge(g, per_rank, world, nb);
        if (c < cost_now) {
            cost_now = c;  norm_now = S.candbuf[center];  disp_now = S.canddepth[center];  prov_now = 0;
            sf |= GPM_SF_ACCEPTED;
        }
    }
}

// One stage of one pixel: [accept of stage-1's exchange] + evaluation of `stage` on this rank's views -> local top-n lists.
template <bool PACKED, bool COLOR>
__device__ __forceinline__ void shard_stage_pixel(const KParams& P, const float* __restrict__ sCam, const float* __restrict__ tile,
                                                  const WarpScratch& ws, cudaTextureObject_t src, cudaTextureObject_t grad,
                                                  float4* __restrict__ planes, float* __restrict__ cost, unsigned char* __restrict__ prov,
                                                  const ShardState& S, const Memo& M, int px, int py, int tile_x0, int tile_y0,
                                                  int stage, int last_stage, const float* __restrict__ gathered_prev, int world,
                                                  float* __restrict__ xchg, unsigned lane, WarpStats& st,
                                                  size_t rank_stride = 0, const ShardP2P* X = nullptr, size_t slot_off = 0)
{
    // fused exchange: copy this pixel's finished lists (count floats at `o`, just written to the local slot) into slot `me` of
    // every peer's region — plain stores over NVLink; the block raises the arrival flags after its last pixel
    auto publish = [&](const float* o, int count) {
        if (!X) return;
        __syncwarp();
        const size_t off = slot_off + (size_t)(o - xchg);
        for (int i = (int)lane; i < count; i += 32) {
            const float v = o[i];
            for (int d = 0; d < X->world; d++)
                if (d != X->me) X->lists[d][off + i] = v;
        }
    };
    const RefCam& cam = P.ref;
    const int W = P.W, H = P.H, Wh = (W + 1) >> 1, nb = P.n_best;
    const size_t center = (size_t)py * W + px;
    const float inf = __int_as_float(0x7f800000);
    const float fpx = __int2float_rn(px), fpy = __int2float_rn(py);
    float c0, c1;
    float4 norm_now = planes[center];
    if (stage == 0) {                                       // initial costs (gipuma.cu:1040-1049), both colours: two half-resolution planes
        float* out = xchg + ((size_t)((px + py) & 1) * H * Wh + (size_t)py * Wh + (px >> 1)) * nb;
        setup_window<COLOR>(P, tile + P.tile_xo, ws, px, py, tile_x0, tile_y0, lane);
        eval_plane<(COLOR ? 1 : 0), PACKED, COLOR>(P, sCam, ws, src, grad, norm_now.x, norm_now.y, norm_now.z, norm_now.w, inf, lane, st, &c0, &c1);
        local_topn(P, c0, c1, lane, out);
        publish(out, nb);
        return;
    }
    float cost_now = cost[center];
    int prov_now = prov[center];
    unsigned sf = stage >= 2 ? (unsigned)S.sflags[center] : 0u;
    float disp_now = stage >= 2 ? S.dispbuf[center] : plane_depth(cam, norm_now.x, norm_now.y, norm_now.z, norm_now.w, fpx, fpy);
    unsigned mmask = P.memo ? M.mask[center] : 0u;
    unsigned new_mask = mmask;
    unsigned own_id = M.pid[center];
    if (stage >= 2) {
        shard_accept_pixel(P, planes, S, M, px, py, stage - 1, gathered_prev, world, rank_stride, norm_now, cost_now, disp_now, prov_now, sf, own_id, lane);
        if (stage == 2) {
            // refinement memo (see k_sweep): the S candidates are a pure function of (pixel, plane at refinement start)
            sf = 0u;
            if (P.memo && P.rng_mode == 0) {
                if ((mmask & GPM_MEMO_REFINE) && M.refseen[center] == own_id) sf = GPM_SF_SKIP_REFINE;
                else { if (lane == 0) M.refseen[center] = own_id;  new_mask &= ~GPM_MEMO_REFINE; }
            }
        }
    }
    float* out = xchg + ((size_t)py * Wh + (px >> 1)) * ((stage == 1 || rank_stride) ? 8 : 1) * nb;      // see shard_accept_pixel
    if (stage > last_stage) {                               // closing pass of a colour: only the accept of the last refinement step
        if (P.memo && P.rng_mode == 0 && !(sf & (GPM_SF_SKIP_REFINE | GPM_SF_ACCEPTED))) new_mask |= GPM_MEMO_REFINE;
        if (sf & GPM_SF_ACCEPTED) {                         // the refinement created a new plane: one fresh identity
            unsigned fresh = 0u;
            if (lane == 0) fresh = atomicAdd(M.next_id, 1u);
            own_id = __shfl_sync(GPM_FULL, fresh, 0);
        }
    } else if (stage == 1) {
        float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
        bool mine_ok = false;
        size_t mine_at = center;
        if (lane < 8) {
            const int dist = lane < 4 ? 1 : 5, dir = lane & 3;
            int qx = px, qy = py;
            bool ok;
            if (dir == 0)      { ok = py > dist - 1;  qy = py - dist; }
            else if (dir == 1) { ok = py < H - dist;  qy = py + dist; }
            else if (dir == 2) { ok = px > dist - 1;  qx = px - dist; }
            else               { ok = px < W - dist;  qx = px + dist; }
            mine_ok = ok;
            mine_at = (size_t)qy * W + qx;
        }
        const unsigned mine_id = mine_ok ? M.pid[mine_at] : 0u;
        const bool old_ok = P.memo && lane < 8 && ((mmask >> lane) & 1);
        const unsigned old_id = old_ok ? M.seen[center * 8 + lane] : 0u;
        const bool memo_hit = old_ok && mine_ok && old_id == mine_id;
        if (mine_ok && !memo_hit) mine = planes[mine_at];
        const unsigned cand_mask = __ballot_sync(GPM_FULL, mine_ok);
        const int site = COLOR ? 0 : 1;
        bool window_ready = false;
        for (int k = 0; k < 8; k++) {
            float4 nbp;
            nbp.x = __shfl_sync(GPM_FULL, mine.x, k);  nbp.y = __shfl_sync(GPM_FULL, mine.y, k);
            nbp.z = __shfl_sync(GPM_FULL, mine.z, k);  nbp.w = __shfl_sync(GPM_FULL, mine.w, k);
            const unsigned cand_id = __shfl_sync(GPM_FULL, mine_id, k);
            bool skip = !((cand_mask >> k) & 1);
            if (!skip) skip = __any_sync(GPM_FULL, old_ok && cand_id == old_id);            // offered before, from any direction
            if (!skip) {
                const float d = plane_depth(cam, nbp.x, nbp.y, nbp.z, nbp.w, fpx, fpy);
                const bool in_range = d >= cam.depthMin && d <= cam.depthMax;
                // a plane equal to the current one (cost of the same rounding variant) or to an earlier slot can never pass
                // `cost < cost_now`: cost_now only decreases while the slots are accepted in order
                const bool same_now = P.dedupe_self && prov_now == site && cand_id == own_id;
                const bool same_mine = P.dedupe_cand && mine_ok && (int)lane < k && cand_id == mine_id;
                skip = !in_range || same_now || __any_sync(GPM_FULL, same_mine);
            }
            if (skip) { st.skip++;  if ((int)lane < nb) out[k * nb + lane] = inf;  __syncwarp();  continue; }
            if (!window_ready) { setup_window<COLOR>(P, tile + P.tile_xo, ws, px, py, tile_x0, tile_y0, lane);  window_ready = true; }
            eval_plane<(COLOR ? 0 : 1), PACKED, COLOR>(P, sCam, ws, src, grad, nbp.x, nbp.y, nbp.z, nbp.w, inf, lane, st, &c0, &c1);
            local_topn(P, c0, c1, lane, out + k * nb);
        }
        if (P.memo && mine_ok && !memo_hit) M.seen[center * 8 + lane] = mine_id;
        new_mask |= (cand_mask & 0xffu);
        publish(out, 8 * nb);
    } else {
        if (sf & GPM_SF_SKIP_REFINE) {
            st.skip++;
            if ((int)lane < nb) out[lane] = inf;
        } else {
            float depth_new;
            const float4 cand = shard_refine_candidate(P, norm_now, disp_now, fpx, fpy, stage - 2, depth_new);
            if (lane == 0) { S.candbuf[center] = cand;  S.canddepth[center] = depth_new; }
            setup_window<COLOR>(P, tile + P.tile_xo, ws, px, py, tile_x0, tile_y0, lane);
            eval_plane<0, PACKED, COLOR>(P, sCam, ws, src, grad, cand.x, cand.y, cand.z, cand.w, inf, lane, st, &c0, &c1);
            local_topn(P, c0, c1, lane, out);
        }
        publish(out, nb);
    }
    if (lane == 0) {
        if (stage >= 2) { planes[center] = norm_now;  cost[center] = cost_now;  prov[center] = (unsigned char)prov_now;  M.pid[center] = own_id; }
        if (stage >= 1) { S.dispbuf[center] = disp_now;  S.sflags[center] = (unsigned char)sf; }
        if (P.memo && new_mask != mmask) M.mask[center] = new_mask;
    }
}

// One exchange stage over the whole image (NCCL flow: the all-gather runs between two launches of this kernel).
// stage 0: both colours; 1 .. last_stage: one colour; last_stage + 1: closing accept of the colour.
template <bool PACKED, bool COLOR>
__global__ void __launch_bounds__(GPM_LB_THREADS, GPM_LB_BLOCKS)
k_shard_stage(const __grid_constant__ KParams P, const __grid_constant__ CUtensorMap tmap, const ViewCam* __restrict__ cams, const float* __restrict__ refpad,
              cudaTextureObject_t src, cudaTextureObject_t grad, float4* __restrict__ planes, float* __restrict__ cost,
              unsigned char* __restrict__ prov, ShardState S, Memo M, int colour, int stage, int last_stage, const float* __restrict__ gathered_prev,
              int world, float* __restrict__ xchg, unsigned long long* __restrict__ stats)
{
    extern __shared__ __align__(128) float smem[];
    float* tile = smem;
    float* sCam = tile + tile_floats(P);
    float* scratch = smem + fixed_smem_floats(P);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_x0 = blockIdx.x * GPM_TILE - P.halo, tile_y0 = blockIdx.y * GPM_TILE - P.halo;
    if (stage <= last_stage) stage_block(P, &tmap, refpad, cams, tile, sCam, tile_x0, tile_y0);
    const WarpScratch ws = carve(scratch + (size_t)warp * warp_scratch_floats(P.ns_pad, P.V, P.color), P.ns_pad, P.V, P.color, block_perm(P, smem));
    WarpStats st = {0, 0, 0, 0, 0};
    // stage 0 visits every pixel (both colours, like gipuma_init_cu2); the others one colour
    const int npix = (stage == 0) ? GPM_TILE * GPM_TILE : GPM_TILE * GPM_TILE / 2;
    // gridDim.z slices of the tile's pixel list, as in k_sweep
    const int per_slice = (npix + (int)gridDim.z - 1) / (int)gridDim.z;
    const int idx_end = min(npix, ((int)blockIdx.z + 1) * per_slice);
    for (int idx = (int)blockIdx.z * per_slice + (int)warp; idx < idx_end; idx += P.nwarps) {
        int px, py;
        if (stage == 0) { px = blockIdx.x * GPM_TILE + (idx & 31);  py = blockIdx.y * GPM_TILE + (idx >> 5); }
        else { const int tx = idx & 31, ty = idx >> 5;  px = blockIdx.x * GPM_TILE + tx;  py = blockIdx.y * GPM_TILE + 2 * ty + (((tx & 1) ^ colour) & 1); }
        if (px >= P.W || py >= P.H) continue;
        shard_stage_pixel<PACKED, COLOR>(P, sCam, tile, ws, src, grad, planes, cost, prov, S, M, px, py,
                                         tile_x0, tile_y0, stage, last_stage, gathered_prev, world, xchg, lane, st);
    }
    flush_stats(stats, st, lane);
}

// ---- fused compute + exchange over peer memory (NVLink / NVSwitch) ---------------------------------------------------
// Polling load: relaxed, system scope (served by L2, the point of coherence for peer stores into this GPU's memory).  The
// acquire is ONE fence after the poll succeeds: an acquiring load (or a system-scope fence) invalidates the SM's whole L1 —
// the texture cache of every resident block — so it must not sit inside the spin loop (measured: with ld.acquire.sys in the
// loop two resident blocks per SM ran 7 % SLOWER than one, profiles/r02_shard_*_4gpu.json).
__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acquire_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until every peer has delivered its lists of exchange `seq` for this block (one spinning thread per peer)
__device__ __forceinline__ void p2p_wait(const ShardP2P& X, unsigned bid, unsigned seq)
{
    if ((int)threadIdx.x < X.world && (int)threadIdx.x != X.me) {
        const unsigned* f = X.flags[X.me] + (size_t)threadIdx.x * X.nblocks + bid;
        const long long t0 = clock64();
        while ((int)(ld_relaxed_sys(f) - seq) < 0) {
            if (*(volatile unsigned*)X.err) break;
            if (clock64() - t0 > (1LL << 33)) { atomicExch(X.err, 1u);  break; }      // ~4 s: fail loudly on the host, never hang
            __nanosleep(200);
        }
        fence_acquire_sys();
    }
    __syncthreads();
}
// all lists of this block are stored: make them visible system-wide, then raise flag [me][bid] = seq at every peer
__device__ __forceinline__ void p2p_signal(const ShardP2P& X, unsigned bid, unsigned seq)
{
    __syncthreads();
    if ((int)threadIdx.x < X.world && (int)threadIdx.x != X.me) {
        __threadfence_system();
        st_release_sys(X.flags[threadIdx.x] + (size_t)X.me * X.nblocks + bid, seq);
    }
}

// One whole colour pass (all 1 + S exchange stages and the closing accept) — or, with init_phase, the initial-cost stage and
// its accept — of one tile slice per block, the exchange fused in: after every stage the block's lists are already in every
// peer's memory (stored there pixel by pixel while the block was still sampling) and only the arrival flags remain to be
// raised; before the next stage the block waits for the same block of every peer — never for another block of its own GPU,
// so blocks need not be co-resident and the transfer of one tile overlaps the sampling of all the others.  The reference
// window is staged once per colour instead of once per stage.  `seq0` = number of exchanges completed before this launch.
template <bool PACKED, bool COLOR>
__global__ void __launch_bounds__(GPM_LB_THREADS, GPM_LB_BLOCKS)
k_shard_fused(const __grid_constant__ KParams P, const __grid_constant__ CUtensorMap tmap, const ViewCam* __restrict__ cams,
              const float* __restrict__ refpad, cudaTextureObject_t src, cudaTextureObject_t grad, float4* __restrict__ planes,
              float* __restrict__ cost, unsigned char* __restrict__ prov, ShardState S, Memo M, int colour, int init_phase, int last_stage,
              const __grid_constant__ ShardP2P X, unsigned seq0, unsigned long long* __restrict__ stats)
{
    extern __shared__ __align__(128) float smem[];
    float* tile = smem;
    float* sCam = tile + tile_floats(P);
    float* scratch = smem + fixed_smem_floats(P);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_x0 = blockIdx.x * GPM_TILE - P.halo, tile_y0 = blockIdx.y * GPM_TILE - P.halo;
    const unsigned bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    stage_block(P, &tmap, refpad, cams, tile, sCam, tile_x0, tile_y0);
    const WarpScratch ws = carve(scratch + (size_t)warp * warp_scratch_floats(P.ns_pad, P.V, P.color), P.ns_pad, P.V, P.color, block_perm(P, smem));
    WarpStats st = {0, 0, 0, 0, 0};
    const int W = P.W, H = P.H, Wh = (W + 1) >> 1, nb = P.n_best;
    const int npix = init_phase ? GPM_TILE * GPM_TILE : GPM_TILE * GPM_TILE / 2;
    const int per_slice = (npix + (int)gridDim.z - 1) / (int)gridDim.z;
    const int idx0 = (int)blockIdx.z * per_slice, idx_end = min(npix, idx0 + per_slice);
    const int first = init_phase ? 0 : 1, lastS = init_phase ? 0 : last_stage;
    const size_t slot = (size_t)X.slot_floats;
    for (int stage = first; stage <= lastS + 1; stage++) {
        const bool closing = stage > lastS;
        const unsigned seq = seq0 + (unsigned)(stage - first) + 1u;          // exchange produced by `stage`
        if (stage > first) p2p_wait(X, bid, seq - 1u);
        const float* prev = X.lists[X.me] + (size_t)((seq - 1u) & 1u) * X.world * slot;
        const size_t slot_off = ((size_t)(seq & 1u) * X.world + X.me) * slot;
        float* outb = X.lists[X.me] + slot_off;
        for (int idx = idx0 + (int)warp; idx < idx_end; idx += P.nwarps) {
            int px, py;
            if (init_phase) { px = blockIdx.x * GPM_TILE + (idx & 31);  py = blockIdx.y * GPM_TILE + (idx >> 5); }
            else { const int tx = idx & 31, ty = idx >> 5;  px = blockIdx.x * GPM_TILE + tx;  py = blockIdx.y * GPM_TILE + 2 * ty + (((tx & 1) ^ colour) & 1); }
            if (px >= W || py >= H) continue;
            if (init_phase && closing) {                                      // merged initial cost (gipuma.cu:1040-1049 over all views)
                if (lane == 0) {
                    const float* g = prev + ((size_t)((px + py) & 1) * H * Wh + (size_t)py * Wh + (px >> 1)) * nb;
                    cost[(size_t)py * W + px] = shard_merge(g, slot, X.world, nb);
                    prov[(size_t)py * W + px] = P.color ? 3 : 0;
                }
                continue;
            }
            shard_stage_pixel<PACKED, COLOR>(P, sCam, tile, ws, src, grad, planes, cost, prov, S, M, px, py,
                                             tile_x0, tile_y0, stage, lastS, prev, X.world, outb, lane, st, slot, &X, slot_off);
        }
        if (!closing) p2p_signal(X, bid, seq);
    }
    flush_stats(stats, st, lane);
}

// closing accept of stage 0: merged initial costs for both colours (thread per half-row position)
__global__ void k_shard_accept0(const __grid_constant__ KParams P, float* __restrict__ cost, unsigned char* __restrict__ prov,
                                const float* __restrict__ gathered, int world)
{
    const int W = P.W, H = P.H, Wh = (W + 1) >> 1, nb = P.n_best;
    const int hx = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (hx >= Wh || py >= H) return;
    const size_t per_rank = (size_t)2 * H * Wh * nb;
    for (int col = 0; col < 2; col++) {
        const int px = 2 * hx + ((py + col) & 1);                      // the pixel of this colour with x/2 == hx in row py
        if (px >= W) continue;
        const size_t center = (size_t)py * W + px;
        const float* g = gathered + ((size_t)col * H * Wh + (size_t)py * Wh + hx) * nb;
        cost[center] = shard_merge(g, per_rank, world, nb);
        prov[center] = P.color ? 3 : 0;          // initialisation variant
    }
}

// ---- final depth / world normal — gipuma_compute_disp, gipuma.cu:1080-1103 ------------------
__global__ void k_finalize(const __grid_constant__ KParams P, float4* __restrict__ planes, const float* __restrict__ cost)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.W || y >= P.H) return;
    const size_t center = (size_t)y * P.W + x;
    const float4 n = planes[center];
    const float* R = P.ref.R_orig_inv;
    float4 o;
    o.x = ffma(n.z, R[2], ffma(n.x, R[0], fmul(n.y, R[1])));
    o.y = ffma(n.z, R[5], ffma(n.x, R[3], fmul(n.y, R[4])));
    o.z = ffma(n.z, R[8], ffma(n.x, R[6], fmul(n.y, R[7])));
    o.w = (cost[center] != GPM_MAXCOST) ? plane_depth(P.ref, n.x, n.y, n.z, n.w, __int2float_rn(x), __int2float_rn(y)) : 0.0f;
    planes[center] = o;
}

// Ceiling of the unit that bounds this path, measured in process: filtered R32F fetches per second of the texture unit in its
// best case — 32 lanes = 16 x 2 adjacent texels (every hardware quad a 2x2 block; 1143 Gfetch/s in tools/texbench.cu), the five
// taps of the reference's sample pattern, one layer per warp — on the context's own layered texture.  bench.py divides the
// fetch rate it achieves by this number (roofline.binding_unit).
__global__ void k_fetch_peak(cudaTextureObject_t src, int xmask, int ymask, int V, int reps, float* __restrict__ sink)
{
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const float dx = (float)(lane >> 1), dy = (float)(lane & 1);
    const int v = warp % V;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
    unsigned h = (unsigned)warp * 2654435761u + 12345u;
    for (int r = 0; r < reps; r++) {
        h = h * 1664525u + 1013904223u;
        const float x = 8.37f + (float)((h >> 8) & (unsigned)xmask) + dx;
        const float y = 8.61f + (float)((h >> 20) & (unsigned)ymask) + dy;
        a0 += tex2DLayered<float>(src, x, y, v);
        a1 += tex2DLayered<float>(src, x + 1.0f, y, v);
        a2 += tex2DLayered<float>(src, x - 1.0f, y, v);
        a3 += tex2DLayered<float>(src, x, y + 1.0f, v);
        a4 += tex2DLayered<float>(src, x, y - 1.0f, v);
    }
    const float a = a0 + a1 + a2 + a3 + a4;
    if (a == 12345.678f) sink[0] = a;
}

// Colour source views are sampled from three R32F planes (layer 3v + channel) instead of one RGBA32F layer.
__global__ void k_split_channels(const float4* __restrict__ in, size_t pitch_elems, int W, int H, float* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float4 p = in[(size_t)y * pitch_elems + x];
    const size_t plane = (size_t)W * H, o = (size_t)y * W + x;
    out[o] = p.x;  out[plane + o] = p.y;  out[2 * plane + o] = p.z;
}

// G[x,y] = (I[clamp(x+1),y] - I[clamp(x-1),y],  I[x,clamp(y+1)] - I[x,clamp(y-1)]) for the packed sampling mode, and a flag
// that stays 1 only if every pixel is an integer in [0, 255] (the exactness condition of that mode).
__global__ void k_make_gradients(const float* __restrict__ img, size_t pitch_floats, int W, int H, float2* __restrict__ out,
                                 int* __restrict__ all_8bit)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float* row = img + (size_t)y * pitch_floats;
    const float v = row[x];
    const float gx = fsub(row[min(x + 1, W - 1)], row[max(x - 1, 0)]);
    const float gy = fsub(img[(size_t)min(y + 1, H - 1) * pitch_floats + x], img[(size_t)max(y - 1, 0) * pitch_floats + x]);
    out[(size_t)y * W + x] = make_float2(gx, gy);
    if (!(v >= 0.0f && v <= 255.0f && v == rintf(v))) *all_8bit = 0;
}

__global__ void k_pad_reference4(const float4* __restrict__ img, size_t pitch_elems, int W, int H, float4* __restrict__ out, int out_pitch, int out_rows)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out_pitch || y >= out_rows) return;
    const int sx = min(max(x - GPM_APRON, 0), W - 1), sy = min(max(y - GPM_APRON, 0), H - 1);
    out[(size_t)y * out_pitch + x] = img[(size_t)sy * pitch_elems + sx];
}

// replicate-pad the reference image by GPM_APRON on the left / top and out to the end of the allocation on the right / bottom
// (== the texture's clamp addressing, main.cpp:644-645): the window of the last tile row / column reaches
// roundup(size, 32) + halo + GPM_APRON, past size + 2 * GPM_APRON whenever size % 32 is small
__global__ void k_pad_reference(const float* __restrict__ img, size_t pitch_floats, int W, int H,
                                float* __restrict__ out, int out_pitch, int out_rows)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= out_pitch || y >= out_rows) return;
    const int sx = min(max(x - GPM_APRON, 0), W - 1), sy = min(max(y - GPM_APRON, 0), H - 1);
    out[(size_t)y * out_pitch + x] = img[(size_t)sy * pitch_floats + sx];
}

}  // namespace gpm
