/* gipuma_oracle.c — single-thread CPU restatement of Gipuma's PatchMatch hot path (reference gipuma.cu).
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build,
 * load or call this file; the product (gipuma_b200/) never does.
 *
 * What it is: a plain-C, scalar restatement of the reference algorithm, function by function, each citing the
 * gipuma.cu lines it follows, with a software model of the texture unit (bilinear, clamp addressing, 8-bit
 * fractional weights — CUDA C Programming Guide, "Texture Fetching / Linear Filtering").
 * What it is not: bit-exact with the GPU.  The reference is compiled with --use_fast_math (approximate
 * reciprocal / exp2 / rsqrt, flush-to-zero, ptxas FMA fusion) and filters through texture hardware whose exact
 * arithmetic is undocumented; a CPU cannot reproduce those bit patterns.  PARITY PINNING: bit-level parity of the
 * CUDA product is pinned against the *compiled reference itself* (oracle/_ref, built by oracle/build_ref.sh) and
 * against golden outputs of that build (tests/golden/); this C restatement is pinned to the compiled reference
 * within tolerance by tests/test_oracle_vs_golden.py (cost level: |dc| <= 2e-3 * max(1, c)) and serves as the
 * readable specification, the function-level checker, and the CPU timing baseline.
 * Third-party arithmetic on the path, cuRAND's XORWOW (CUDA toolkit 12.9 curand_kernel.h): curand_init's seed scrambling and
 * skip-ahead, the generator step and curand_uniform are restated here too (gpo_curand_init, gpo_init_planes) and pinned to
 * known answers computed by the toolkit header itself on the host (tests/golden/xorwow_init_kat.json).
 *
 * Build: gcc -O2 -fopenmp -fPIC -shared -o oracle/libgipuma_oracle.so oracle/gipuma_oracle.c -lm   (oracle/Makefile)
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gipuma_b200.h" /* gpm_params / gpm_camera PODs (the boundary's field list) */

#define MAXCOST 1000.0f /* config.h:22 */

typedef struct gpo_scene {
    int W, H, V;
    const gpm_params* prm;
    const gpm_camera* ref;    /* cameras[REFERENCE] */
    const gpm_camera* views;  /* [V] cameras[viewSelectionSubset[v]] */
    const float* ref_img;     /* H*W (gray) or H*W*4 interleaved (colour) */
    const float* const* view_imgs; /* [V], same layout each */
    int color;                /* params.color_processing: T = float4 (gipuma.cu:1965-1966) */
} gpo_scene;

/* Colour mode of the entry points below (off by default).  The images then hold 4 interleaved floats per pixel, the
 * 4th unused, as main.cpp:560-605 uploads them. */
static int g_color = 0;
void gpo_set_color(int on) { g_color = on != 0; }

/* Rows of one colour are independent (a colour only reads the other colour's planes), and so are cost evaluations:
 * the row loops below run on `g_threads` OpenMP threads (default 1; results do not depend on it). */
static int g_threads = 1;
void gpo_set_threads(int n) { g_threads = n > 0 ? n : 1; }
int gpo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- texture unit model: tex2D<float>(tex, x, y), Linear filter, clamp, unnormalised (main.cpp:642-648) ---- */
static float texel(const float* img, int W, int H, int i, int j)
{
    if (i < 0) i = 0; else if (i >= W) i = W - 1;
    if (j < 0) j = 0; else if (j >= H) j = H - 1;
    return img[(size_t)j * W + i];
}
static float tex2d(const float* img, int W, int H, float x, float y)
{
    if (!(x == x) || !(y == y)) return texel(img, W, H, 0, 0);
    float xb = x - 0.5f, yb = y - 0.5f;
    if (xb < -4.0f) xb = -4.0f; else if (xb > (float)W + 4.0f) xb = (float)W + 4.0f;   /* far outside == clamped edge */
    if (yb < -4.0f) yb = -4.0f; else if (yb > (float)H + 4.0f) yb = (float)H + 4.0f;
    const float fi = floorf(xb), fj = floorf(yb);
    const float a = floorf((xb - fi) * 256.0f + 0.5f) * (1.0f / 256.0f);    /* 1.8 fixed-point weights */
    const float b = floorf((yb - fj) * 256.0f + 0.5f) * (1.0f / 256.0f);
    const int i = (int)fi, j = (int)fj;
    const float t00 = texel(img, W, H, i, j), t10 = texel(img, W, H, i + 1, j);
    const float t01 = texel(img, W, H, i, j + 1), t11 = texel(img, W, H, i + 1, j + 1);
    return (1.f - a) * (1.f - b) * t00 + a * (1.f - b) * t10 + (1.f - a) * b * t01 + a * b * t11;
}

/* float4 texture: the same filter per channel (one set of weights) */
static float texel_c(const float* img, int W, int H, int i, int j, int ch)
{
    if (i < 0) i = 0; else if (i >= W) i = W - 1;
    if (j < 0) j = 0; else if (j >= H) j = H - 1;
    return img[((size_t)j * W + i) * 4 + ch];
}
static void texel3(const float* img, int W, int H, int i, int j, float out[3])
{
    for (int ch = 0; ch < 3; ch++) out[ch] = texel_c(img, W, H, i, j, ch);
}
static void tex2d3(const float* img, int W, int H, float x, float y, float out[3])
{
    if (!(x == x) || !(y == y)) { texel3(img, W, H, 0, 0, out); return; }
    float xb = x - 0.5f, yb = y - 0.5f;
    if (xb < -4.0f) xb = -4.0f; else if (xb > (float)W + 4.0f) xb = (float)W + 4.0f;
    if (yb < -4.0f) yb = -4.0f; else if (yb > (float)H + 4.0f) yb = (float)H + 4.0f;
    const float fi = floorf(xb), fj = floorf(yb);
    const float a = floorf((xb - fi) * 256.0f + 0.5f) * (1.0f / 256.0f);
    const float b = floorf((yb - fj) * 256.0f + 0.5f) * (1.0f / 256.0f);
    const int i = (int)fi, j = (int)fj;
    for (int ch = 0; ch < 3; ch++) {
        const float t00 = texel_c(img, W, H, i, j, ch), t10 = texel_c(img, W, H, i + 1, j, ch);
        const float t01 = texel_c(img, W, H, i, j + 1, ch), t11 = texel_c(img, W, H, i + 1, j + 1, ch);
        out[ch] = (1.f - a) * (1.f - b) * t00 + a * (1.f - b) * t10 + (1.f - a) * b * t01 + a * b * t11;
    }
}
/* l1_norm(float4), gipuma.cu:173-178: the 4th channel does not take part */
static float l1_norm3(const float a[3], const float b[3])
{
    return (fabsf(a[0] - b[0]) + fabsf(a[1] - b[1]) + fabsf(a[2] - b[2])) * 0.3333333f;
}

/* ---- geometry ---------------------------------------------------------------------------------------- */
/* getDepthFromPlane3_cu / getDisparity_cu, gipuma.cu:694-715 */
static float plane_depth(const gpm_camera* c, const float n[4], int px, int py)
{
    if (n[3] != n[3]) return 1000.0f;
    return -n[3] * c->fx / ((n[0] * ((float)px - c->K[2])) + (n[1] * ((float)py - c->K[5])) * c->alpha + n[2] * c->fx);
}
/* getD_cu, gipuma.cu:96-111 */
static float plane_d(const gpm_camera* c, const float n[3], int px, int py, float depth)
{
    const float X = depth * (float)px - c->P_col34[0], Y = depth * (float)py - c->P_col34[1], Z = depth - c->P_col34[2];
    const float* M = c->M_inv;
    const float wx = M[0] * X + M[1] * Y + M[2] * Z, wy = M[3] * X + M[4] * Y + M[5] * Z, wz = M[6] * X + M[7] * Y + M[8] * Z;
    return -(n[0] * wx + n[1] * wy + n[2] * wz);
}
/* getViewVector_cu, gipuma.cu:122-130 */
static void view_vector(const gpm_camera* c, int px, int py, float v[3])
{
    const float a = (float)px - c->P_col34[0], b = (float)py - c->P_col34[1], e = 1.0f - c->P_col34[2];
    const float* M = c->M_inv;
    float x = M[0] * a + M[1] * b + M[2] * e - c->C[0];
    float y = M[3] * a + M[4] * b + M[5] * e - c->C[1];
    float z = M[6] * a + M[7] * b + M[8] * e - c->C[2];
    const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
    v[0] = x * inv; v[1] = y * inv; v[2] = z * inv;
}
/* getHomography_cu, gipuma.cu:339-356: H = K_to (R - t n^T / d) K_ref^-1 */
static void homography(const gpm_camera* ref, const gpm_camera* to, const float n[4], float H[9])
{
    float A[9], T[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A[3 * i + j] = to->R[3 * i + j] - to->t[i] * n[j] / n[3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            T[3 * i + j] = A[3 * i] * ref->K_inv[j] + A[3 * i + 1] * ref->K_inv[3 + j] + A[3 * i + 2] * ref->K_inv[6 + j];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            H[3 * i + j] = to->K[3 * i] * T[j] + to->K[3 * i + 1] * T[3 + j] + to->K[3 * i + 2] * T[6 + j];
}

/* ---- photo-consistency cost --------------------------------------------------------------------------- */
/* pmCost_shared / pmCost + pmCostComputation(_shared) + weight_cu: gipuma.cu:585-680, 455-518, 223-320, 186-193.
 * The reference window is read with clamp addressing, exactly what the shared tile / texture deliver. */
/* Diagnostics (tools/prune_study.py): when set, view_cost() also stores every sample's term w*dis, [v][sample] with
 * row length g_tap_stride, in window order. */
static float* g_tap = NULL;
static int g_tap_stride = 0;

static float view_cost(const gpo_scene* s, int v, int px, int py, const float n[4], int rad)
{
    const gpm_params* p = s->prm;
    int tap_k = 0;
    float H[9];
    homography(s->ref, &s->views[v], n, H);
    const float* L = s->ref_img;
    const float* R = s->view_imgs[v];
    const int W = s->W, Hh = s->H;
    const float center = texel(L, W, Hh, px, py);
    float cost = 0.0f;
    for (int i = -rad; i < rad + 1; i += 2) {          /* WIN_INCREMENT 2, x offset outer (gipuma.cu:633) */
        for (int j = -rad; j < rad + 1; j += 2) {      /* y offset inner (:634) */
            const int x = px + i, y = py + j;
            const float left = texel(L, W, Hh, x, y);
            const float w = expf(-fabsf(left - center) / p->gamma);                       /* weight_cu :186-193 */
            const float fx = (float)x, fy = (float)y;                                     /* getCorrespondingPoint_cu :207-217 */
            const float X = H[0] * fx + H[1] * fy + H[2], Y = H[3] * fx + H[4] * fy + H[5], Z = H[6] * fx + H[7] * fy + H[8];
            const float qx = X / Z, qy = Y / Z;
            const float gx2 = tex2d(R, W, Hh, qx + 1 + 0.5f, qy + 0.5f) - tex2d(R, W, Hh, qx - 1 + 0.5f, qy + 0.5f);   /* :251 */
            const float gy2 = tex2d(R, W, Hh, qx + 0.5f, qy + 1 + 0.5f) - tex2d(R, W, Hh, qx + 0.5f, qy - 1 + 0.5f);   /* :252 */
            const float colDiff = fabsf(left - tex2d(R, W, Hh, qx + 0.5f, qy + 0.5f));                                  /* :253 */
            const float gx1 = texel(L, W, Hh, x + 1, y) - texel(L, W, Hh, x - 1, y);                                    /* :258 */
            const float gy1 = texel(L, W, Hh, x, y + 1) - texel(L, W, Hh, x, y - 1);                                    /* :259 */
            const float gradDis = fminf((fabsf(gx1 - gx2) + fabsf(gy1 - gy2)) * 0.0625f, p->tau_gradient);              /* :267 */
            const float colDis = fminf(colDiff, p->tau_color);                                                          /* :271 */
            cost = cost + w * ((1.f - p->alpha) * colDis + p->alpha * gradDis);                                         /* :272-274, :674 */
            if (g_tap) g_tap[(size_t)v * g_tap_stride + tap_k++] = w * ((1.f - p->alpha) * colDis + p->alpha * gradDis);
        }
    }
    return cost;
}

/* The same with T = float4: every image difference becomes l1_norm(float4) (gipuma.cu:173-178), i.e. the mean absolute
 * difference of the three colour channels; gradients are per-channel differences (:251-252, :258-259). */
static float view_cost_color(const gpo_scene* s, int v, int px, int py, const float n[4], int rad)
{
    const gpm_params* p = s->prm;
    float H[9];
    homography(s->ref, &s->views[v], n, H);
    const float* L = s->ref_img;
    const float* R = s->view_imgs[v];
    const int W = s->W, Hh = s->H;
    float center[3];
    texel3(L, W, Hh, px, py, center);
    float cost = 0.0f;
    for (int i = -rad; i < rad + 1; i += 2) {
        for (int j = -rad; j < rad + 1; j += 2) {
            const int x = px + i, y = py + j;
            float left[3], a[3], b[3], c[3], d[3], gx1[3], gy1[3], gx2[3], gy2[3], right[3];
            texel3(L, W, Hh, x, y, left);
            const float w = expf(-l1_norm3(left, center) / p->gamma);                     /* weight_cu<float4> :186-190 */
            const float fx = (float)x, fy = (float)y;
            const float X = H[0] * fx + H[1] * fy + H[2], Y = H[3] * fx + H[4] * fy + H[5], Z = H[6] * fx + H[7] * fy + H[8];
            const float qx = X / Z, qy = Y / Z;
            tex2d3(R, W, Hh, qx + 1 + 0.5f, qy + 0.5f, a);  tex2d3(R, W, Hh, qx - 1 + 0.5f, qy + 0.5f, b);
            tex2d3(R, W, Hh, qx + 0.5f, qy + 1 + 0.5f, c);  tex2d3(R, W, Hh, qx + 0.5f, qy - 1 + 0.5f, d);
            tex2d3(R, W, Hh, qx + 0.5f, qy + 0.5f, right);
            for (int k = 0; k < 3; k++) { gx2[k] = a[k] - b[k];  gy2[k] = c[k] - d[k]; }
            texel3(L, W, Hh, x + 1, y, a);  texel3(L, W, Hh, x - 1, y, b);
            texel3(L, W, Hh, x, y + 1, c);  texel3(L, W, Hh, x, y - 1, d);
            for (int k = 0; k < 3; k++) { gx1[k] = a[k] - b[k];  gy1[k] = c[k] - d[k]; }
            const float colDiff = l1_norm3(left, right);                                                   /* :253 */
            const float gradDis = fminf((l1_norm3(gx1, gx2) + l1_norm3(gy1, gy2)) * 0.0625f, p->tau_gradient);   /* :267 */
            const float colDis = fminf(colDiff, p->tau_color);
            cost = cost + w * ((1.f - p->alpha) * colDis + p->alpha * gradDis);
        }
    }
    return cost;
}

/* sort_small, gipuma.cu:684-693 */
static void sort_small(float* d, int n)
{
    for (int i = 1; i < n; i++) {
        float tmp = d[i];
        int j;
        for (j = i; j >= 1 && tmp < d[j - 1]; j--) d[j] = d[j - 1];
        d[j] = tmp;
    }
}

/* pmCostMultiview_cu, gipuma.cu:720-806 */
static float multiview_cost(const gpo_scene* s, int px, int py, const float n[4], int rad)
{
    const gpm_params* p = s->prm;
    float cv[GPM_MAX_VIEWS];
    int numValid = 0;
    for (int v = 0; v < s->V; v++) {
        float c = s->color ? view_cost_color(s, v, px, py, n, rad) : view_cost(s, v, px, py, n, rad);
        if (c < MAXCOST) numValid++; else c = MAXCOST;                  /* :771-774 */
        cv[v] = c;
    }
    sort_small(cv, s->V);                                               /* :779 */
    int numBest = numValid;
    if (p->cost_comb == GPM_COMB_BEST_N) numBest = numBest < p->n_best ? numBest : p->n_best;   /* :783-784 */
    if (p->cost_comb == GPM_COMB_GOOD) numBest = s->V;                                           /* :785-786 */
    const float thresh = cv[0] * p->good_factor;
    float cost = 0.0f;
    int considered = 0;
    for (int i = 0; i < numBest; i++) {                                 /* :790-797 */
        float c = cv[i];
        considered++;
        if (p->cost_comb == GPM_COMB_GOOD) c = fminf(c, thresh);
        cost = cost + c;
    }
    cost = cost / (float)considered;
    if (considered < 1) cost = MAXCOST;
    if (cost != cost || cost > MAXCOST || cost < 0) cost = MAXCOST;     /* :799-803 */
    return cost;
}

/* ---- RNG: XORWOW (curand_kernel.h curand()), curand_uniform, gipuma.cu:138-169 ---------------------------- */
typedef struct { uint32_t v[5], d; } gpo_xorwow;
static uint32_t xorwow(gpo_xorwow* s)
{
    const uint32_t t = s->v[0] ^ (s->v[0] >> 2);
    s->v[0] = s->v[1]; s->v[1] = s->v[2]; s->v[2] = s->v[3]; s->v[3] = s->v[4];
    s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
    s->d += 362437u;
    return s->v[4] + s->d;
}
static float uniform01(gpo_xorwow* s) { return (float)xorwow(s) * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }
static float between(gpo_xorwow* s, float lo, float hi) { return uniform01(s) * (hi - lo) + lo; }   /* curand_between :138-141 */

/* Random plane of gipuma_init_cu2 (gipuma.cu:1021-1034) from a given XORWOW state (6 words: v[0..4], d). */
void gpo_random_plane(const gpm_params* p, const gpm_camera* ref, int px, int py, const uint32_t state[6], float out[4])
{
    gpo_xorwow r;
    memcpy(r.v, state, 5 * sizeof(uint32_t));
    r.d = state[5];
    float vv[3];
    view_vector(ref, px, py, vv);
    const float disp = between(&r, p->min_disparity, p->max_disparity);                /* :1028 */
    float x = 1.f, y = 1.f, sum = 2.f;                                                  /* Marsaglia :148-164 */
    while (sum >= 1.0f) { x = between(&r, -1.f, 1.f); y = between(&r, -1.f, 1.f); sum = x * x + y * y; }
    const float sq = sqrtf(1.0f - sum);
    float n[3] = {2.0f * x * sq, 2.0f * y * sq, 1.0f - 2.0f * sum};
    if (n[0] * vv[0] + n[1] * vv[1] + n[2] * vv[2] > 0.0f) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }   /* :131-137 */
    const float depth = ref->f * ref->baseline / disp;                                  /* :1031 */
    out[0] = n[0]; out[1] = n[1]; out[2] = n[2];
    out[3] = plane_d(ref, n, px, py, depth);                                            /* :1034 */
}

/* ---- curand_init(seed, subsequence, offset) for XORWOW — CUDA toolkit 12.9, curand_kernel.h:772-856 ---------------------
 * The reference calls curand_init(seed, p.y, p.x, &state) per pixel (gipuma.cu:1019).  cuRAND defines it as: scramble the
 * seed into (v[0..4], d) (:836-847); advance by `subsequence` * 2^67 steps (skipahead_sequence); advance by `offset` steps
 * (skipahead; d += 362437 * offset).  The toolkit does the jumps with precomputed matrices; the restatement below builds the
 * 2^67-step matrix itself: the five v words evolve linearly over GF(2) (the step of curand(), :863-871, without the Weyl
 * counter d), so one step is a 160 x 160 bit matrix M and the subsequence jump is M^(2^67), 67 squarings.  d is untouched by
 * subsequence jumps (2^67 * 362437 = 0 mod 2^32, :697, :736).  Pinned by tests/golden/xorwow_init_kat.json, generated from
 * the toolkit header compiled for the host (tools/make_xorwow_kat.cu). */
static void xorwow_linear_step(uint32_t v[5])
{
    const uint32_t t = v[0] ^ (v[0] >> 2);
    v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
    v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
}
typedef struct { uint32_t col[160][5]; } xorwow_mat;         /* column j = image of basis vector e_j */
static void xorwow_matvec(const xorwow_mat* m, const uint32_t in[5], uint32_t out[5])
{
    uint32_t r[5] = {0, 0, 0, 0, 0};
    for (int j = 0; j < 160; j++)
        if ((in[j >> 5] >> (j & 31)) & 1u)
            for (int k = 0; k < 5; k++) r[k] ^= m->col[j][k];
    memcpy(out, r, sizeof(r));
}
static const xorwow_mat* xorwow_sequence_matrix(void)        /* M^(2^67), built once */
{
    static xorwow_mat a, b;
    static int ready = 0;
    if (!ready) {
        for (int j = 0; j < 160; j++) {
            uint32_t e[5] = {0, 0, 0, 0, 0};
            e[j >> 5] = 1u << (j & 31);
            xorwow_linear_step(e);
            memcpy(a.col[j], e, sizeof(e));
        }
        for (int sq = 0; sq < 67; sq++) {                    /* a <- a * a */
            for (int j = 0; j < 160; j++) xorwow_matvec(&a, a.col[j], b.col[j]);
            a = b;
        }
        ready = 1;
    }
    return &a;
}
/* state6 = (v[0..4], d) after curand_init(seed, subsequence, offset) */
void gpo_curand_init(uint64_t seed, uint64_t subsequence, uint64_t offset, uint32_t state6[6])
{
    const uint32_t s0 = (uint32_t)seed ^ 0xaad26b49u, s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;   /* curand_kernel.h:836-847 */
    const uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    uint32_t v[5] = {123456789u + t0, 362436069u ^ t0, 521288629u + t1, 88675123u ^ t1, 5783321u + t0};
    uint32_t d = 6615241u + t1 + t0;
    const xorwow_mat* A = xorwow_sequence_matrix();
    xorwow_mat P = *A, T;                                    /* v <- A^subsequence v by binary powers of A */
    for (uint64_t n = subsequence; n; n >>= 1) {
        if (n & 1) xorwow_matvec(&P, v, v);
        if (n >> 1) { for (int j = 0; j < 160; j++) xorwow_matvec(&P, P.col[j], T.col[j]);  P = T; }
    }
    for (uint64_t k = 0; k < offset; k++) xorwow_linear_step(v);     /* offsets on this path are pixel columns: step them */
    d += 362437u * (uint32_t)offset;
    memcpy(state6, v, sizeof(v));
    state6[5] = d;
}
/* gipuma_init_cu2's planes for the whole image (gipuma.cu:1019-1034): curand_init(seed, y, x) per pixel, then the draws of
 * gpo_random_plane.  Row states by one subsequence jump per row, pixels of a row by single steps. */
void gpo_random_plane(const gpm_params* p, const gpm_camera* ref, int px, int py, const uint32_t state[6], float out[4]);
int gpo_init_planes(int W, int H, const gpm_params* prm, const gpm_camera* ref, uint64_t seed, float* planes)
{
    uint32_t row[6];
    gpo_curand_init(seed, 0, 0, row);
    const xorwow_mat* A = xorwow_sequence_matrix();
    for (int y = 0; y < H; y++) {
        uint32_t st[6];
        memcpy(st, row, sizeof(st));
        for (int x = 0; x < W; x++) {
            gpo_random_plane(prm, ref, x, y, st, planes + 4 * ((size_t)y * W + x));
            xorwow_linear_step(st);                            /* offset x + 1 */
            st[5] += 362437u;
        }
        xorwow_matvec(A, row, row);                            /* subsequence y + 1 (d unchanged) */
    }
    return 0;
}

/* ---- entry points ---------------------------------------------------------------------------------------- */
static void make_scene(gpo_scene* s, int W, int H, int V, const gpm_params* prm, const gpm_camera* ref,
                       const gpm_camera* views, const float* ref_img, const float* const* view_imgs)
{
    s->W = W; s->H = H; s->V = V; s->prm = prm; s->ref = ref; s->views = views; s->ref_img = ref_img; s->view_imgs = view_imgs;
    s->color = g_color;
}

/* Cost of the given planes at the pixels of rows [y0, y1) (all columns).  init_radius != 0 uses box/2
 * (gipuma.cu:1012) instead of (box-1)/2 (:1474). */
int gpo_cost_eval(int W, int H, int V, const gpm_params* prm, const gpm_camera* ref, const gpm_camera* views,
                  const float* ref_img, const float* const* view_imgs, const float* planes, float* cost,
                  int y0, int y1, int init_radius)
{
    gpo_scene s;
    make_scene(&s, W, H, V, prm, ref, views, ref_img, view_imgs);
    const int rad = init_radius ? prm->box_hsize / 2 : (prm->box_hsize - 1) / 2;
#pragma omp parallel for schedule(dynamic, 1) num_threads(g_threads)
    for (int y = y0; y < y1; y++)
        for (int x = 0; x < W; x++)
            cost[(size_t)y * W + x] = multiview_cost(&s, x, y, planes + 4 * ((size_t)y * W + x), rad);
    return 0;
}

/* spatialPropagation_cu, gipuma.cu:832-874 */
static void propagate(const gpo_scene* s, int px, int py, const float* nb, float* cost_now, float* norm_now, float* disp_now, int rad)
{
    const float disp_before = plane_depth(s->ref, nb, px, py);
    const float cost_before = multiview_cost(s, px, py, nb, rad);
    if (disp_before >= s->prm->depthMin && disp_before <= s->prm->depthMax) {            /* :829-830, :865 */
        if (cost_before < *cost_now) {
            *disp_now = disp_before;
            memcpy(norm_now, nb, 4 * sizeof(float));
            *cost_now = cost_before;
        }
    }
}

/* One phase set of one checkerboard colour, in place: colour 0 = black, 1 = red; phase_mask bit0 close
 * (gipuma.cu:1471-1588), bit1 far (:1353-1468), bit3 the 20 candidates of the fused kernel (:1122-1351, built when
 * SMALLKERNEL is not defined), bit2 refine (:1590-1711, planeRefinement_cu :928-994 with an
 * all-zero XORWOW state per pixel — pin P2).  Pixel (x, y) is black iff (x + y) is even (:1730-1734).
 * A colour only reads the other colour's planes, so the pixel order inside a colour is irrelevant.
 * Restricted to rows [y0, y1) for bounded CPU timing. */
int gpo_phase(int W, int H, int V, const gpm_params* prm, const gpm_camera* ref, const gpm_camera* views,
              const float* ref_img, const float* const* view_imgs, float* planes, float* cost,
              int colour, int phase_mask, int y0, int y1)
{
    gpo_scene s;
    make_scene(&s, W, H, V, prm, ref, views, ref_img, view_imgs);
    const int rad = (prm->box_hsize - 1) / 2;
#pragma omp parallel for schedule(dynamic, 1) num_threads(g_threads)
    for (int py = y0; py < y1; py++) {
        for (int px = 0; px < W; px++) {
            if (((px + py) & 1) != colour) continue;
            const size_t center = (size_t)py * W + px;
            float norm_now[4];
            memcpy(norm_now, planes + 4 * center, sizeof(norm_now));
            float cost_now = cost[center];
            float disp_now = plane_depth(ref, norm_now, px, py);
            for (int far = 0; far < 2; far++) {
                if (!((phase_mask >> far) & 1)) continue;
                const int d = far ? 5 : 1;                                   /* :1439-1446, :1560-1567 */
                if (py > d - 1) propagate(&s, px, py, planes + 4 * (center - (size_t)d * W), &cost_now, norm_now, &disp_now, rad);
                if (py < H - d) propagate(&s, px, py, planes + 4 * (center + (size_t)d * W), &cost_now, norm_now, &disp_now, rad);
                if (px > d - 1) propagate(&s, px, py, planes + 4 * (center - d), &cost_now, norm_now, &disp_now, rad);
                if (px < W - d) propagate(&s, px, py, planes + 4 * (center + d), &cost_now, norm_now, &disp_now, rad);
            }
            if (phase_mask & 8) {
                /* the fused kernel's 20 candidates in source order with the reference's own guards,
                 * gipuma_checkerboard_cu, gipuma.cu:1236-1330 (EXTRAPOINT, EXTRAPOINTFAR, EXTRAPOINT2 all defined, :36-38) */
                static const int dx[20] = {0, 0, 0, 0, 0, 0, -1, -3, -5, 1, 3, 5, 2, 2, -2, -2, -1, 1, -1, 1};
                static const int dy[20] = {-1, -3, -5, 1, 3, 5, 0, 0, 0, 0, 0, 0, -1, 1, -1, 1, -2, -2, 2, 2};
                const int guard[20] = {
                    py > 0, py > 2, py > 4, py < H - 1, py < H - 3, py < H - 5,                 /* up, upup, up 5; down ... :1236-1264 */
                    px > 0, px > 2, px > 4, px < W - 1, px < W - 3, px < W - 5,                 /* left ...; right ...      :1266-1292 */
                    py > 0 && px < W - 2, py < H - 1 && px < W - 2, py > 0 && px > 1, py < H - 1 && px > 1,      /* :1295-1311 */
                    px > 0 && py > 2, px < W - 1 && py > 2, px > 0 && py < H - 2, px < W - 1 && py < H - 2 };   /* :1312-1329 */
                for (int k = 0; k < 20; k++)
                    if (guard[k])
                        propagate(&s, px, py, planes + 4 * ((size_t)(py + dy[k]) * W + (px + dx[k])), &cost_now, norm_now, &disp_now, rad);
            }
            if (phase_mask & 4) {
                gpo_xorwow r;
                memset(&r, 0, sizeof(r));
                float vv[3];
                view_vector(ref, px, py, vv);
                float deltaN = 1.0f;
                for (float deltaZ = prm->max_disparity / 2.0f; deltaZ >= 0.01f; deltaZ = deltaZ / 10.0f) {   /* :958-959 */
                    /* getRndDispAndUnitVector_cu :890-927 */
                    float disp = ref->f * ref->baseline / disp_now;
                    const float minDelta = -fminf(deltaZ, prm->min_disparity + disp);
                    const float maxDelta = fminf(deltaZ, prm->max_disparity - disp);
                    const float dz = between(&r, minDelta, maxDelta);
                    float dnew = fminf(fmaxf(disp + dz, prm->min_disparity), prm->max_disparity);
                    const float depth_new = ref->f * ref->baseline / dnew;
                    float cand[4];
                    cand[0] = norm_now[0] + between(&r, -deltaN, deltaN);
                    cand[1] = norm_now[1] + between(&r, -deltaN, deltaN);
                    cand[2] = norm_now[2] + between(&r, -deltaN, deltaN);
                    const float inv = 1.0f / sqrtf(cand[0] * cand[0] + cand[1] * cand[1] + cand[2] * cand[2]);
                    cand[0] *= inv; cand[1] *= inv; cand[2] *= inv;
                    if (cand[0] * vv[0] + cand[1] * vv[1] + cand[2] * vv[2] > 0.0f) { cand[0] = -cand[0]; cand[1] = -cand[1]; cand[2] = -cand[2]; }
                    cand[3] = plane_d(ref, cand, px, py, depth_new);                            /* :969 */
                    const float c = multiview_cost(&s, px, py, cand, rad);
                    if (c < cost_now) { cost_now = c; disp_now = depth_new; memcpy(norm_now, cand, sizeof(cand)); }   /* :986-990 */
                    deltaN = deltaN / 4.0f;                                                     /* :992 */
                }
            }
            cost[center] = cost_now;
            memcpy(planes + 4 * center, norm_now, sizeof(norm_now));
        }
    }
    return 0;
}

/* gipuma_compute_disp, gipuma.cu:1080-1103 */
int gpo_finalize(int W, int H, const gpm_camera* ref, float* planes, const float* cost)
{
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float* n = planes + 4 * ((size_t)y * W + x);
            const float* R = ref->R_orig_inv;
            const float o0 = R[0] * n[0] + R[1] * n[1] + R[2] * n[2], o1 = R[3] * n[0] + R[4] * n[1] + R[5] * n[2];
            const float o2 = R[6] * n[0] + R[7] * n[1] + R[8] * n[2];
            const float w = (cost[(size_t)y * W + x] != MAXCOST) ? plane_depth(ref, n, x, y) : 0.0f;
            n[0] = o0; n[1] = o1; n[2] = o2; n[3] = w;
        }
    return 0;
}

/* Whole sweeps on rows [y0, y1): `iterations` x (black{close,far,refine}, red{...}) — gipuma.cu:1911-1941. */
int gpo_sweep(int W, int H, int V, const gpm_params* prm, const gpm_camera* ref, const gpm_camera* views,
              const float* ref_img, const float* const* view_imgs, float* planes, float* cost, int iterations, int y0, int y1)
{
    for (int it = 0; it < iterations; it++)
        for (int colour = 0; colour < 2; colour++)
            for (int ph = 1; ph <= 4; ph <<= 1)
                gpo_phase(W, H, V, prm, ref, views, ref_img, view_imgs, planes, cost, colour, ph, y0, y1);
    return 0;
}

/* Combined cost of one plane at one pixel plus every sample term of every view (diagnostics). */
float gpo_multiview_terms(int W, int H, int V, const gpm_params* prm, const gpm_camera* ref, const gpm_camera* views,
                          const float* ref_img, const float* const* view_imgs, int px, int py, const float* plane,
                          float* terms, int stride)
{
    gpo_scene s;
    make_scene(&s, W, H, V, prm, ref, views, ref_img, view_imgs);
    g_tap = terms;  g_tap_stride = stride;
    const float c = multiview_cost(&s, px, py, plane, (prm->box_hsize - 1) / 2);
    g_tap = NULL;
    return c;
}

float gpo_tex2d(const float* img, int W, int H, float x, float y) { return tex2d(img, W, H, x, y); }
float gpo_plane_depth(const gpm_camera* c, const float n[4], int px, int py) { return plane_depth(c, n, px, py); }
float gpo_plane_d(const gpm_camera* c, const float n[3], int px, int py, float depth) { return plane_d(c, n, px, py, depth); }
void gpo_homography(const gpm_camera* ref, const gpm_camera* to, const float n[4], float H[9]) { homography(ref, to, n, H); }
void gpo_view_vector(const gpm_camera* c, int px, int py, float v[3]) { view_vector(c, px, py, v); }
