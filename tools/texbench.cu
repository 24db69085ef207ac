// tools/texbench.cu — B200 texture-path microbenchmark + exactness probe (design evidence, not product).
//
// Questions it answers (results summarised in profiles/ and DESIGN.md):
//   1. Filtered-fetch rate of the texture unit for the reference's access pattern (5 bilinear fetches of an
//      R32F texture per sample: centre, x+-1, y+-1) versus ONE fetch of a packed texel (I, Gx, Gy, .) in
//      RGBA32F / RGBA16F, where Gx(x,y)=I(x+1,y)-I(x-1,y), Gy likewise (clamped indices).
//   2. Whether the packed fetch is BIT-IDENTICAL to the 5-fetch arithmetic for integer-valued (8-bit)
//      images — bilinear filtering is linear and, with 8-bit fractional weights and integer texels
//      |v| <= 255, every product/sum is exactly representable in fp32 and fp16 texels hold the inputs exactly.
// Build: nvcc -O3 --use_fast_math -gencode arch=compute_100a,code=sm_100a -o tools/texbench tools/texbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static cudaTextureObject_t make_tex(cudaArray_t arr)
{
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = cudaAddressModeWrap;      // as the reference: Wrap + unnormalised == clamp
    td.addressMode[1] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t t;
    CK(cudaCreateTextureObject(&t, &res, &td, NULL));
    return t;
}

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// mode 0: 5 x R32F; 1: 1 x RGBA32F; 2: 1 x RGBA16F; 3: 4 x gather R32F; 4: 1 x R32F (centre only)
// pattern 0: warp = 32 horizontally adjacent pixels (reference mapping), samples walk a 8x8 stride-2 patch
// pattern 1: warp = one patch, lanes = 32 of the 8x8 stride-2 samples (warp-per-pixel mapping)
template <int MODE, int PATTERN>
__global__ void bench_kernel(cudaTextureObject_t t1, cudaTextureObject_t t4, cudaTextureObject_t th,
                             int W, int H, int reps, float* sink)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = tid >> 5;
    float acc = 0.f;
    for (int r = 0; r < reps; r++) {
        uint32_t h = hash32((PATTERN == 0 ? (uint32_t)warp : (uint32_t)warp) * 977u + r * 131071u);
        // a smooth mapping: base point + ~1.03 scale + small shear, sub-pixel offset from the hash
        float bx = 40.f + (float)(h % (uint32_t)(W - 120)) + (float)((h >> 20) & 255) * (1.f / 256.f);
        float by = 40.f + (float)((h >> 8) % (uint32_t)(H - 120)) + (float)((h >> 12) & 255) * (1.f / 256.f);
#pragma unroll 1
        for (int s = 0; s < (PATTERN == 0 ? 64 : 2); s++) {
            float dx, dy;
            if (PATTERN == 0) { dx = (float)lane + 2.f * (float)(s >> 3); dy = (float)(lane & 1) + 2.f * (float)(s & 7); }
            else { int q = lane + 32 * s; dx = 2.f * (float)(q >> 3); dy = 2.f * (float)(q & 7); }
            float x = bx + 1.03f * dx + 0.05f * dy, y = by - 0.04f * dx + 0.98f * dy;
            if (MODE == 0) {
                float c = tex2D<float>(t1, x + 0.5f, y + 0.5f);
                float gx = tex2D<float>(t1, x + 1 + 0.5f, y + 0.5f) - tex2D<float>(t1, x - 1 + 0.5f, y + 0.5f);
                float gy = tex2D<float>(t1, x + 0.5f, y + 1 + 0.5f) - tex2D<float>(t1, x + 0.5f, y - 1 + 0.5f);
                acc += c + fabsf(gx) + fabsf(gy);
            } else if (MODE == 1) {
                float4 v = tex2D<float4>(t4, x + 0.5f, y + 0.5f);
                acc += v.x + fabsf(v.y) + fabsf(v.z);
            } else if (MODE == 2) {
                float4 v = tex2D<float4>(th, x + 0.5f, y + 0.5f);
                acc += v.x + fabsf(v.y) + fabsf(v.z);
            } else if (MODE == 3) {
                float4 a = tex2Dgather<float4>(t1, x - 1 + 0.5f, y + 0.5f, 0);
                float4 b = tex2Dgather<float4>(t1, x + 1 + 0.5f, y + 0.5f, 0);
                float4 c = tex2Dgather<float4>(t1, x + 0.5f, y - 1 + 0.5f, 0);
                float4 d = tex2Dgather<float4>(t1, x + 0.5f, y + 1 + 0.5f, 0);
                acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
            } else {
                acc += tex2D<float>(t1, x + 0.5f, y + 0.5f);
            }
        }
    }
    if (acc == 12345.678f) sink[0] = acc;
}

// lane-arrangement probe: 32 lanes = COLS x ROWS samples (ROWS = 1 << LR) at `stride` pixels, 5 R32F fetches each
template <int LR>
__global__ void shape_kernel(cudaTextureObject_t t1, int W, int H, int reps, float stride, float* sink)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = tid >> 5;
    const float dx = stride * (float)(lane >> LR), dy = stride * (float)(lane & ((1 << LR) - 1));
    float acc = 0.f;
    for (int r = 0; r < reps; r++) {
        uint32_t h = hash32((uint32_t)warp * 977u + r * 131071u);
        float bx = 40.f + (float)(h % (uint32_t)(W - 160)) + (float)((h >> 20) & 255) * (1.f / 256.f);
        float by = 40.f + (float)((h >> 8) % (uint32_t)(H - 160)) + (float)((h >> 12) & 255) * (1.f / 256.f);
        float x = bx + 1.03f * dx + 0.05f * dy, y = by - 0.04f * dx + 0.98f * dy;
        float c = tex2D<float>(t1, x + 0.5f, y + 0.5f);
        float gx = tex2D<float>(t1, x + 1 + 0.5f, y + 0.5f) - tex2D<float>(t1, x - 1 + 0.5f, y + 0.5f);
        float gy = tex2D<float>(t1, x + 0.5f, y + 1 + 0.5f) - tex2D<float>(t1, x + 0.5f, y - 1 + 0.5f);
        acc += c + fabsf(gx) + fabsf(gy);
    }
    if (acc == 12345.678f) sink[0] = acc;
}

template <int LR>
static void run_shape(cudaTextureObject_t t1, int W, int H, float stride, float* sink)
{
    const int blocks = 148 * 16, threads = 256, reps = 512;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    shape_kernel<LR><<<blocks, threads>>>(t1, W, H, reps, stride, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 3; i++) shape_kernel<LR><<<blocks, threads>>>(t1, W, H, reps, stride, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fetches = (double)blocks * threads * reps * 5 * 3;
    printf("shape %2d cols x %2d rows, stride %.0f px : %7.2f Gfetch/s  %.2f fetch/clk/SM  (%.1f clk per warp-fetch @1.9GHz)\n",
           32 >> LR, 1 << LR, stride, fetches / (ms * 1e6), fetches / (ms * 1e-3) / 148.0 / 1.9e9,
           32.0 / (fetches / (ms * 1e-3) / 148.0 / 1.9e9));
}

// layered-array variant of the 4 cols x 8 rows shape (what gpm::eval_plane issues), with a homography-like scale
template <bool LAYERED>
__global__ void layered_kernel(cudaTextureObject_t t1, cudaTextureObject_t tl, int W, int H, int reps, float scale, int nlayers, float* sink)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = tid >> 5;
    const float dx = 2.f * (float)(lane >> 3), dy = 2.f * (float)(lane & 7);
    float acc = 0.f;
    for (int r = 0; r < reps; r++) {
        uint32_t h = hash32((uint32_t)warp * 977u + r * 131071u);
        float bx = 60.f + (float)(h % (uint32_t)(W - 200)) + (float)((h >> 20) & 255) * (1.f / 256.f);
        float by = 60.f + (float)((h >> 8) % (uint32_t)(H - 200)) + (float)((h >> 12) & 255) * (1.f / 256.f);
        float x = bx + scale * (1.0f * dx + 0.15f * dy), y = by + scale * (-0.12f * dx + 0.98f * dy);
        const int layer = r % nlayers;
        float c, xp, xm, yp, ym;
        if (LAYERED) {
            c = tex2DLayered<float>(tl, x + 0.5f, y + 0.5f, layer);
            xp = tex2DLayered<float>(tl, x + 1 + 0.5f, y + 0.5f, layer);  xm = tex2DLayered<float>(tl, x - 1 + 0.5f, y + 0.5f, layer);
            yp = tex2DLayered<float>(tl, x + 0.5f, y + 1 + 0.5f, layer);  ym = tex2DLayered<float>(tl, x + 0.5f, y - 1 + 0.5f, layer);
        } else {
            c = tex2D<float>(t1, x + 0.5f, y + 0.5f);
            xp = tex2D<float>(t1, x + 1 + 0.5f, y + 0.5f);  xm = tex2D<float>(t1, x - 1 + 0.5f, y + 0.5f);
            yp = tex2D<float>(t1, x + 0.5f, y + 1 + 0.5f);  ym = tex2D<float>(t1, x + 0.5f, y - 1 + 0.5f);
        }
        acc += c + fabsf(xp - xm) + fabsf(yp - ym);
    }
    if (acc == 12345.678f) sink[0] = acc;
}

template <bool LAYERED>
static void run_layered(cudaTextureObject_t t1, cudaTextureObject_t tl, int W, int H, float scale, int nlayers, float* sink)
{
    const int blocks = 148 * 16, threads = 256, reps = 512;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    layered_kernel<LAYERED><<<blocks, threads>>>(t1, tl, W, H, reps, scale, nlayers, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 3; i++) layered_kernel<LAYERED><<<blocks, threads>>>(t1, tl, W, H, reps, scale, nlayers, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fetches = (double)blocks * threads * reps * 5 * 3;
    printf("%s 4x8 stride-2 patch, scale %.2f, %2d layers : %7.2f Gfetch/s  (%.1f clk per warp-fetch @1.9GHz)\n", LAYERED ? "layered" : "plain  ",
           scale, nlayers, fetches / (ms * 1e6), 32.0 / (fetches / (ms * 1e-3) / 148.0 / 1.9e9));
}

// exactness probe: random coordinates (interior), compare bitwise
__global__ void exact_kernel(cudaTextureObject_t t1, cudaTextureObject_t t4, cudaTextureObject_t th,
                             int W, int H, int n, unsigned long long* counts, float* examples)
{
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    uint32_t h1 = hash32(tid * 2654435761u + 17u), h2 = hash32(h1 + 0x9e3779b9u), h3 = hash32(h2 ^ 0x85ebca6bu);
    // full-precision random coordinate inside [3, W-4] x [3, H-4]
    float x = 3.f + (float)(h1 >> 8) * (1.f / 16777216.f) * (float)(W - 7);
    float y = 3.f + (float)(h2 >> 8) * (1.f / 16777216.f) * (float)(H - 7);
    if ((h3 & 7) == 0) x = floorf(x) + (float)((h3 >> 8) & 1023) * (1.f / 1024.f);   // some on a 1/1024 grid
    if ((h3 & 7) == 1) y = floorf(y) + 0.5f;                                          // some on half-texel centres
    const float cx = x + 0.5f, cy = y + 0.5f;
    const float xp = x + 1 + 0.5f, xm = x - 1 + 0.5f, yp = y + 1 + 0.5f, ym = y - 1 + 0.5f;
    float c = tex2D<float>(t1, cx, cy);
    float gx = tex2D<float>(t1, xp, cy) - tex2D<float>(t1, xm, cy);
    float gy = tex2D<float>(t1, cx, yp) - tex2D<float>(t1, cx, ym);
    float4 v4 = tex2D<float4>(t4, cx, cy);
    float4 vh = tex2D<float4>(th, cx, cy);
    bool aligned = (xp - cx == 1.f) && (cx - xm == 1.f) && (yp - cy == 1.f) && (cy - ym == 1.f);
    atomicAdd(&counts[0], 1ULL);
    if (!aligned) atomicAdd(&counts[1], 1ULL);
    bool bad4 = (__float_as_uint(v4.x) != __float_as_uint(c)) || (v4.y != gx) || (v4.z != gy);
    bool badh = (__float_as_uint(vh.x) != __float_as_uint(c)) || (vh.y != gx) || (vh.z != gy);
    if (bad4) atomicAdd(&counts[2], 1ULL);
    if (badh) atomicAdd(&counts[3], 1ULL);
    if (bad4 && aligned) atomicAdd(&counts[4], 1ULL);
    if (badh && aligned) { unsigned long long k = atomicAdd(&counts[5], 1ULL);
        if (k < 8) { float* e = examples + k * 8; e[0] = x; e[1] = y; e[2] = c; e[3] = vh.x; e[4] = gx; e[5] = vh.y; e[6] = gy; e[7] = vh.z; } }
    // software model of the filter: 8-bit fractional weights
    {
        float xb = cx - 0.5f, yb = cy - 0.5f;
        float fx = floorf(xb), fy = floorf(yb);
        // candidates for the hardware's weight quantisation: round-to-nearest of frac*256
        float a = rintf((xb - fx) * 256.f) * (1.f / 256.f), b = rintf((yb - fy) * 256.f) * (1.f / 256.f);
        float t00 = tex2D<float>(t1, fx + 0.5f, fy + 0.5f), t10 = tex2D<float>(t1, fx + 1.5f, fy + 0.5f);
        float t01 = tex2D<float>(t1, fx + 0.5f, fy + 1.5f), t11 = tex2D<float>(t1, fx + 1.5f, fy + 1.5f);
        float m = (1.f - a) * (1.f - b) * t00 + a * (1.f - b) * t10 + (1.f - a) * b * t01 + a * b * t11;
        if (m != c) atomicAdd(&counts[6], 1ULL);
    }
}

template <int MODE, int PATTERN>
static void run_bench(const char* name, cudaTextureObject_t t1, cudaTextureObject_t t4, cudaTextureObject_t th,
                      int W, int H, float* sink)
{
    const int blocks = 148 * 16, threads = 256, reps = (PATTERN == 0 ? 8 : 256);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; i++) bench_kernel<MODE, PATTERN><<<blocks, threads>>>(t1, t4, th, W, H, reps, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int launches = 5;
    for (int i = 0; i < launches; i++) bench_kernel<MODE, PATTERN><<<blocks, threads>>>(t1, t4, th, W, H, reps, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double samples = (double)blocks * threads * reps * (PATTERN == 0 ? 64 : 2) * launches;
    static const int fetches_per_sample[5] = {5, 1, 1, 4, 1};
    printf("%-34s pattern %d : %8.3f ms  %8.2f Gsamples/s  %8.2f Gfetch/s  (%.2f fetch/clk/SM @1.9GHz)\n", name, PATTERN, ms / launches,
           samples / (ms * 1e6), samples * fetches_per_sample[MODE] / (ms * 1e6),
           samples * fetches_per_sample[MODE] / (ms * 1e-3) / 148.0 / 1.9e9);
}

int main()
{
    const int W = 1600, H = 1200;
    std::vector<float> img((size_t)W * H);
    uint32_t s = 12345u;
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        s = s * 1664525u + 1013904223u;
        float v = 127.5f + 60.f * sinf(0.21f * x + 0.07f * y) + 40.f * sinf(0.05f * x - 0.13f * y) + (float)((s >> 24) & 15) - 7.5f;
        img[(size_t)y * W + x] = fminf(fmaxf(rintf(v), 0.f), 255.f);
    }
    auto at = [&](int x, int y) { x = x < 0 ? 0 : (x >= W ? W - 1 : x); y = y < 0 ? 0 : (y >= H ? H - 1 : y); return img[(size_t)y * W + x]; };
    std::vector<float4> img4((size_t)W * H);
    std::vector<ushort4> imgh((size_t)W * H);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        float4 v = make_float4(at(x, y), at(x + 1, y) - at(x - 1, y), at(x, y + 1) - at(x, y - 1), 0.f);
        img4[(size_t)y * W + x] = v;
        ushort4 u;
        u.x = __half_as_ushort(__float2half(v.x)); u.y = __half_as_ushort(__float2half(v.y));
        u.z = __half_as_ushort(__float2half(v.z)); u.w = 0;
        imgh[(size_t)y * W + x] = u;
    }
    cudaArray_t a1, a4, ah;
    cudaChannelFormatDesc d1 = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
    cudaChannelFormatDesc d4 = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat);
    cudaChannelFormatDesc dh = cudaCreateChannelDesc(16, 16, 16, 16, cudaChannelFormatKindFloat);
    CK(cudaMallocArray(&a1, &d1, W, H)); CK(cudaMallocArray(&a4, &d4, W, H)); CK(cudaMallocArray(&ah, &dh, W, H));
    CK(cudaMemcpy2DToArray(a1, 0, 0, img.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice));
    CK(cudaMemcpy2DToArray(a4, 0, 0, img4.data(), W * 16, W * 16, H, cudaMemcpyHostToDevice));
    CK(cudaMemcpy2DToArray(ah, 0, 0, imgh.data(), W * 8, W * 8, H, cudaMemcpyHostToDevice));
    cudaTextureObject_t t1 = make_tex(a1), t4 = make_tex(a4), th = make_tex(ah);
    float* sink; CK(cudaMalloc(&sink, 64));

    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);

    // exactness
    unsigned long long* counts; float* examples;
    CK(cudaMalloc(&counts, 8 * sizeof(unsigned long long))); CK(cudaMemset(counts, 0, 8 * sizeof(unsigned long long)));
    CK(cudaMalloc(&examples, 64 * sizeof(float))); CK(cudaMemset(examples, 0, 64 * sizeof(float)));
    const int n = 1 << 26;
    exact_kernel<<<(n + 255) / 256, 256>>>(t1, t4, th, W, H, n, counts, examples);
    CK(cudaDeviceSynchronize());
    unsigned long long hc[8]; float he[64];
    CK(cudaMemcpy(hc, counts, sizeof(hc), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(he, examples, sizeof(he), cudaMemcpyDeviceToHost));
    printf("exactness over %llu random interior coordinates (integer-valued image):\n", hc[0]);
    printf("  taps not exactly +-1 from centre (fp32 coordinate rounding): %llu\n", hc[1]);
    printf("  RGBA32F packed != 5-fetch: %llu   (of which with aligned taps: %llu)\n", hc[2], hc[4]);
    printf("  RGBA16F packed != 5-fetch: %llu   (of which with aligned taps: %llu)\n", hc[3], hc[5]);
    printf("  software bilinear (rint(frac*256)/256 weights) != hardware centre fetch: %llu\n", hc[6]);
    for (int k = 0; k < 8 && k < (int)hc[5]; k++)
        printf("   ex: x=%.6f y=%.6f  c=%.8f/%.8f gx=%.8f/%.8f gy=%.8f/%.8f\n", he[k * 8], he[k * 8 + 1], he[k * 8 + 2], he[k * 8 + 3],
               he[k * 8 + 4], he[k * 8 + 5], he[k * 8 + 6], he[k * 8 + 7]);

    // lane arrangements (what a warp-level texture instruction costs as a function of the lanes' footprint)
    for (float stride : {2.f, 1.f}) {
        run_shape<0>(t1, W, H, stride, sink); run_shape<1>(t1, W, H, stride, sink); run_shape<2>(t1, W, H, stride, sink);
        run_shape<3>(t1, W, H, stride, sink); run_shape<4>(t1, W, H, stride, sink); run_shape<5>(t1, W, H, stride, sink);
    }
    {
        const int NL = 10;
        cudaArray_t al;
        CK(cudaMalloc3DArray(&al, &d1, make_cudaExtent(W, H, NL), cudaArrayLayered));
        for (int l = 0; l < NL; l++) {
            cudaMemcpy3DParms m = {};
            m.srcPtr = make_cudaPitchedPtr(img.data(), W * 4, W, H);
            m.dstArray = al;  m.dstPos = make_cudaPos(0, 0, l);  m.extent = make_cudaExtent(W, H, 1);  m.kind = cudaMemcpyHostToDevice;
            CK(cudaMemcpy3D(&m));
        }
        cudaTextureObject_t tl = make_tex(al);
        for (float sc : {1.0f, 1.3f, 0.8f}) {
            run_layered<false>(t1, tl, W, H, sc, 1, sink);
            run_layered<true>(t1, tl, W, H, sc, 1, sink);
            run_layered<true>(t1, tl, W, H, sc, NL, sink);
        }
    }
    // throughput
    run_bench<0, 0>("5 x R32F bilinear", t1, t4, th, W, H, sink);
    run_bench<4, 0>("1 x R32F bilinear", t1, t4, th, W, H, sink);
    run_bench<1, 0>("1 x RGBA32F packed", t1, t4, th, W, H, sink);
    run_bench<2, 0>("1 x RGBA16F packed", t1, t4, th, W, H, sink);
    run_bench<3, 0>("4 x R32F gather", t1, t4, th, W, H, sink);
    run_bench<0, 1>("5 x R32F bilinear", t1, t4, th, W, H, sink);
    run_bench<4, 1>("1 x R32F bilinear", t1, t4, th, W, H, sink);
    run_bench<1, 1>("1 x RGBA32F packed", t1, t4, th, W, H, sink);
    run_bench<2, 1>("1 x RGBA16F packed", t1, t4, th, W, H, sink);
    run_bench<3, 1>("4 x R32F gather", t1, t4, th, W, H, sink);
    return 0;
}
