// tools/texprobe.cu — how does the B200 texture unit turn an unnormalised fp32 coordinate into (texel index, 8-bit weight)?
// Design evidence for DESIGN.md §5 (root cause of the "packed" mode's rare mismatches); not product code.
//
//  (A) read-out: a ramp texture I(x, y) = x filtered bilinearly returns i + alpha, i.e. the unit's own fixed-point
//      coordinate.  Every fp32 value of a few coordinate ranges is fetched; positions where the result changes are
//      recorded ("transitions") and three candidate rules are compared on all of them:
//        H1  q = RN_even(x * 256 - 128) / 256          H2  q = floor(x * 256 - 128) / 256
//        H3  q = floor(x * 256 - 128 + 0.5) / 256
//  (B) exactness of the whole filter for 8-bit-valued texels: integer-exact software bilinear with each rule versus the
//      hardware fetch at random coordinates (inside, at the border, outside the image).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/texprobe tools/texprobe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static cudaTextureObject_t make_tex(cudaArray_t arr)
{
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = cudaAddressModeWrap;      // the reference's setting (main.cpp:644-645): Wrap + unnormalised == clamp
    td.addressMode[1] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t t;
    CK(cudaCreateTextureObject(&t, &res, &td, NULL));
    return t;
}

__device__ __forceinline__ float rule(int h, float x)
{
    const float s = x * 256.0f - 128.0f;          // exact for 0.5 <= x < 65536
    float r;
    if (h == 0) r = rintf(s);
    else if (h == 1) r = floorf(s);
    else r = floorf(s + 0.5f);
    return r * (1.0f / 256.0f);
}

struct Transition { uint32_t bits; float before, after; };

// every fp32 bit pattern in [b0, b1): q(x) = tex(ramp); axis 0 = x ramp, 1 = y ramp
__global__ void readout(cudaTextureObject_t ramp, int axis, uint32_t b0, uint32_t b1, float limit, Transition* list, unsigned* nlist,
                        unsigned cap, unsigned long long* mism)
{
    const uint32_t b = b0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= b1) return;
    const float x = __uint_as_float(b), xp = __uint_as_float(b - 1);
    const float q = axis ? tex2D<float>(ramp, 1.5f, x) : tex2D<float>(ramp, x, 1.5f);
    const float qp = axis ? tex2D<float>(ramp, 1.5f, xp) : tex2D<float>(ramp, xp, 1.5f);
    if (q != qp) {
        const unsigned k = atomicAdd(nlist, 1u);
        if (k < cap) { list[k].bits = b; list[k].before = qp; list[k].after = q; }
    }
    for (int h = 0; h < 3; h++) {
        float r = rule(h, x);
        r = fminf(fmaxf(r, 0.0f), limit);          // clamp addressing: i + alpha saturates at 0 and at size-1
        if (r != q) atomicAdd(&mism[h], 1ULL);
    }
}

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ void split(int h, float c, int size, int& i, int& a)
{
    const float s = c * 256.0f - 128.0f;
    float r;
    if (h == 0) r = rintf(s); else if (h == 1) r = floorf(s); else r = floorf(s + 0.5f);
    r = fminf(fmaxf(r, -1024.0f * 256.0f), 70000.0f * 256.0f);
    const int f = (int)r;
    i = f >> 8;  a = f & 255;
    (void)size;
}

// integer-exact bilinear of an 8-bit-valued image with clamp addressing
__device__ __forceinline__ float soft(const float* img, int W, int H, int h, float cx, float cy)
{
    int i, j, a, b;
    split(h, cx, W, i, a);  split(h, cy, H, j, b);
    const int i0 = min(max(i, 0), W - 1), i1 = min(max(i + 1, 0), W - 1);
    const int j0 = min(max(j, 0), H - 1), j1 = min(max(j + 1, 0), H - 1);
    const int t00 = (int)img[j0 * W + i0], t10 = (int)img[j0 * W + i1], t01 = (int)img[j1 * W + i0], t11 = (int)img[j1 * W + i1];
    const int v = (256 - a) * (256 - b) * t00 + a * (256 - b) * t10 + (256 - a) * b * t01 + a * b * t11;
    return (float)v * (1.0f / 65536.0f);
}

__global__ void exact(cudaTextureObject_t t, const float* img, int W, int H, int n, unsigned long long* counts, float* ex)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    const uint32_t h1 = hash32(tid * 2654435761u + 17u), h2 = hash32(h1 + 0x9e3779b9u), h3 = hash32(h2 ^ 0x85ebca6bu);
    float x = -6.f + (float)(h1 >> 8) * (1.f / 16777216.f) * (float)(W + 12);
    float y = -6.f + (float)(h2 >> 8) * (1.f / 16777216.f) * (float)(H + 12);
    if ((h3 & 7) == 0) x = floorf(x) + (float)((h3 >> 8) & 1023) * (1.f / 1024.f);     // on a 1/1024 grid
    if ((h3 & 7) == 1) y = floorf(y) + (float)((h3 >> 8) & 511) * (1.f / 512.f);       // on a 1/512 grid: exact ties of the 8-bit weight
    if ((h3 & 7) == 2) x = floorf(x) + 0.5f + (float)((h3 >> 8) & 511) * (1.f / 512.f);
    const float c = tex2D<float>(t, x, y);
    atomicAdd(&counts[3], 1ULL);
    for (int h = 0; h < 3; h++) {
        const float s = soft(img, W, H, h, x, y);
        if (__float_as_uint(s) != __float_as_uint(c)) {
            const unsigned long long k = atomicAdd(&counts[h], 1ULL);
            if (k < 6) { float* e = ex + (h * 6 + k) * 4; e[0] = x; e[1] = y; e[2] = c; e[3] = s; }
        }
    }
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
    const int RW = 4096, RH = 4096;
    cudaChannelFormatDesc d1 = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
    cudaArray_t ax, ay;
    {
        std::vector<float> rx((size_t)RW * 4), ry((size_t)4 * RH);
        for (int y = 0; y < 4; y++) for (int x = 0; x < RW; x++) rx[(size_t)y * RW + x] = (float)x;
        for (int y = 0; y < RH; y++) for (int x = 0; x < 4; x++) ry[(size_t)y * 4 + x] = (float)y;
        CK(cudaMallocArray(&ax, &d1, RW, 4));  CK(cudaMallocArray(&ay, &d1, 4, RH));
        CK(cudaMemcpy2DToArray(ax, 0, 0, rx.data(), RW * 4, RW * 4, 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy2DToArray(ay, 0, 0, ry.data(), 4 * 4, 4 * 4, RH, cudaMemcpyHostToDevice));
    }
    cudaTextureObject_t tx = make_tex(ax), ty = make_tex(ay);
    const unsigned cap = 1u << 20;
    Transition* dl; unsigned* dn; unsigned long long* dm;
    CK(cudaMalloc(&dl, cap * sizeof(Transition)));  CK(cudaMalloc(&dn, 4));  CK(cudaMalloc(&dm, 3 * 8));
    std::vector<Transition> hl(cap);
    struct Range { float lo, hi; } ranges[] = {{0.25f, 4.0f}, {255.0f, 257.0f}, {1000.0f, 1002.0f}, {1598.0f, 1601.0f}, {4093.0f, 4097.0f}};
    for (int axis = 0; axis < 2; axis++)
        for (const Range& r : ranges) {
            uint32_t b0, b1;
            memcpy(&b0, &r.lo, 4);  memcpy(&b1, &r.hi, 4);
            CK(cudaMemset(dn, 0, 4));  CK(cudaMemset(dm, 0, 24));
            const uint32_t n = b1 - b0;
            readout<<<(n + 255) / 256, 256>>>(axis ? ty : tx, axis, b0, b1, (float)((axis ? RH : RW) - 1), dl, dn, cap, dm);
            CK(cudaDeviceSynchronize());
            unsigned hn; unsigned long long hm[3];
            CK(cudaMemcpy(&hn, dn, 4, cudaMemcpyDeviceToHost));  CK(cudaMemcpy(hm, dm, 24, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hl.data(), dl, (size_t)(hn < cap ? hn : cap) * sizeof(Transition), cudaMemcpyDeviceToHost));
            printf("axis %c range [%g, %g): %u floats, %u transitions; mismatches RN-even %llu  floor %llu  half-up %llu\n",
                   axis ? 'y' : 'x', r.lo, r.hi, n, hn, hm[0], hm[1], hm[2]);
            // the transition points tell the rule: print x*256 - 128 at the first float on the new value, for a sample of them
            std::vector<Transition> v(hl.begin(), hl.begin() + (hn < cap ? hn : cap));
            std::sort(v.begin(), v.end(), [](const Transition& a, const Transition& b) { return a.bits < b.bits; });
            int shown = 0;
            double minfrac = 1e9, maxfrac = -1e9;
            unsigned odd_steps = 0;
            for (size_t k = 0; k < v.size(); k++) {
                float x;  memcpy(&x, &v[k].bits, 4);
                const double s = (double)x * 256.0 - 128.0;
                const double frac = s - floor(s);
                if (frac < minfrac) minfrac = frac;
                if (frac > maxfrac) maxfrac = frac;
                if (fabs((v[k].after - v[k].before) * 256.0 - 1.0) > 1e-6) odd_steps++;
                if (shown < 6 || (k + 3 >= v.size())) { printf("    x=%.9g  s=x*256-128=%.6f  q: %.8f -> %.8f  (q*256: %.4f -> %.4f)\n", x, s, v[k].before, v[k].after, v[k].before * 256.0, v[k].after * 256.0); shown++; }
            }
            printf("    frac(x*256-128) at the transitions: min %.6f max %.6f; steps that are not exactly +1/256: %u\n", minfrac, maxfrac, odd_steps);
        }

    // (B) exactness of the filter on an 8-bit-valued image
    const int W = 1600, H = 1200;
    std::vector<float> img((size_t)W * H);
    uint32_t s = 12345u;
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        s = s * 1664525u + 1013904223u;
        img[(size_t)y * W + x] = (float)((s >> 24) & 255);          // white noise: every weight bit matters
    }
    cudaArray_t a1;
    CK(cudaMallocArray(&a1, &d1, W, H));
    CK(cudaMemcpy2DToArray(a1, 0, 0, img.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice));
    cudaTextureObject_t t1 = make_tex(a1);
    float* dimg;  CK(cudaMalloc(&dimg, img.size() * 4));  CK(cudaMemcpy(dimg, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
    unsigned long long* dc;  float* dex;
    CK(cudaMalloc(&dc, 4 * 8));  CK(cudaMemset(dc, 0, 32));  CK(cudaMalloc(&dex, 3 * 6 * 4 * 4));  CK(cudaMemset(dex, 0, 3 * 6 * 4 * 4));
    const int n = 1 << 26;
    exact<<<(n + 255) / 256, 256>>>(t1, dimg, W, H, n, dc, dex);
    CK(cudaDeviceSynchronize());
    unsigned long long hc[4];  float hex_[72];
    CK(cudaMemcpy(hc, dc, 32, cudaMemcpyDeviceToHost));  CK(cudaMemcpy(hex_, dex, sizeof(hex_), cudaMemcpyDeviceToHost));
    const char* names[3] = {"RN-even", "floor", "half-up"};
    printf("integer-exact software bilinear vs hardware fetch, %llu random coordinates in [-6, size+6) on a white-noise 8-bit image:\n", hc[3]);
    for (int h = 0; h < 3; h++) {
        printf("  rule %-8s: %llu mismatches\n", names[h], hc[h]);
        for (int k = 0; k < 6 && k < (int)hc[h]; k++)
            printf("      x=%.9g y=%.9g  hw=%.9g soft=%.9g\n", hex_[(h * 6 + k) * 4], hex_[(h * 6 + k) * 4 + 1], hex_[(h * 6 + k) * 4 + 2], hex_[(h * 6 + k) * 4 + 3]);
    }
    return 0;
}
