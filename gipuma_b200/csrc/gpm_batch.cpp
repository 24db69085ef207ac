// gpm_batch.cpp — reference-view batch driver: what the reference does with a shell loop that starts
// one `gipuma` process per reference image (scripts/dtu_fast.sh:30-55 — each of them re-reading every image from disk,
// main.cpp:741-745) done by ONE process: a pool of worker threads, one gpm_ctx per device, that share a single page-locked
// copy of the image set (the "shared image cache": every image is registered once and uploaded with asynchronous copies by
// whichever device needs it) and take reference views from a common queue.  Per reference view the driver re-bases the
// cameras on it (gpm_prepare_cameras, cameraGeometryUtils.h:174-353), selects the source views (gpm_select_views,
// main.cpp:430-499, deterministic), derives the depth / disparity range as main.cpp:480-483 and :898-906 do, runs the job
// (gpm_run) and hands back / writes the result (disp.dmb + normals.dmb, main.cpp:1002-1003, for the external fusibile).
// Host C++ over the C-ABI; the only CUDA calls are cudaHostRegister / cudaHostUnregister.
#include "../../include/gipuma_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>

namespace {
std::mutex g_err_mutex;
std::string g_batch_err;
void set_err(const std::string& m) { std::lock_guard<std::mutex> l(g_err_mutex);  if (g_batch_err.empty()) g_batch_err = m; }
}  // namespace

extern "C" const char* gpm_batch_last_error(void)
{
    std::lock_guard<std::mutex> l(g_err_mutex);
    static thread_local std::string copy;
    copy = g_batch_err;
    return copy.c_str();
}

extern "C" int gpm_batch_run(const gpm_batch_desc* d, gpm_batch_stats* stats)
{
    { std::lock_guard<std::mutex> l(g_err_mutex);  g_batch_err.clear(); }
    if (!d || !d->images || !d->P || d->n_images < 2 || d->width < 8 || d->height < 8 || d->n_devices < 1 || !d->devices ||
        d->n_refs < 1 || d->max_views < 1 || d->max_views > GPM_MAX_VIEWS) {
        set_err("gpm_batch_run: bad arguments");
        return GPM_E_ARG;
    }
    const int n = d->n_images, W = d->width, H = d->height;
    const size_t pitch = d->pitch_bytes ? d->pitch_bytes : (size_t)W * sizeof(float);
    const size_t npix = (size_t)W * H;
    // shared image cache: page-lock every image once (cudaHostRegisterPortable: usable from every device's context)
    std::vector<char> registered(n, 0);
    for (int i = 0; i < n; i++)
        registered[i] = cudaHostRegister(const_cast<float*>(d->images[i]), pitch * H, cudaHostRegisterPortable) == cudaSuccess;
    cudaGetLastError();                     // an image that was already pinned (or cannot be) is simply used as it is

    std::atomic<int> next(0), failed(0);
    std::vector<double> sweep_ms(d->n_refs, 0.0);
    std::vector<int> views_used(d->n_refs, 0), device_of(d->n_refs, -1);
    auto worker = [&](int dev) {
        gpm_ctx* ctx = nullptr;
        if (gpm_create(&ctx, dev, W, H, d->max_views) != GPM_OK) { set_err(std::string("gpm_create: ") + gpm_last_error());  failed++;  return; }
        std::vector<double> P((size_t)n * 12);
        std::vector<gpm_camera> cams(n);
        std::vector<int> order(n), subset(n);
        std::vector<float> n4, cost;
        for (;;) {
            const int job = next.fetch_add(1);
            if (job >= d->n_refs || failed.load()) break;
            const int ref = d->ref_indices ? d->ref_indices[job] : job;
            if (ref < 0 || ref >= n) { set_err("gpm_batch_run: reference index out of range");  failed++;  break; }
            // cameras with the reference first (main.cpp: img_filenames[0] is the reference image)
            order[0] = ref;
            for (int i = 0, k = 1; i < n; i++) if (i != ref) order[k++] = i;
            for (int i = 0; i < n; i++) memcpy(&P[(size_t)i * 12], d->P + (size_t)order[i] * 12, 12 * sizeof(double));
            int rc = gpm_prepare_cameras(P.data(), n, d->cam_scale > 0 ? d->cam_scale : 1.0, cams.data());
            float range[2] = {0.f, 0.f};
            const int nv = rc == GPM_OK ? gpm_select_views(cams.data(), n, W, H, d->min_angle, d->max_angle, d->max_views, subset.data(), range) : -1;
            if (nv < 1) { set_err("gpm_batch_run: no source view passes the angle filter for reference " + std::to_string(ref));  failed++;  break; }
            gpm_params p = d->params;
            if (p.depthMin <= 0) p.depthMin = range[0];                                   // main.cpp:480-483
            if (p.depthMax <= 0) p.depthMax = range[1];
            p.min_disparity = cams[0].f * cams[0].baseline / p.depthMax;                 // main.cpp:905-906
            p.max_disparity = cams[0].f * cams[0].baseline / p.depthMin;
            rc = gpm_set_params(ctx, &p);
            if (rc == GPM_OK) rc = gpm_set_reference(ctx, d->images[ref], pitch, 0, &cams[0]);
            for (int v = 0; v < nv && rc == GPM_OK; v++) rc = gpm_set_view(ctx, v, d->images[order[subset[v]]], pitch, 0, &cams[subset[v]]);
            if (rc == GPM_OK) rc = gpm_set_num_views(ctx, nv);
            if (rc == GPM_OK) rc = gpm_set_rng(ctx, d->seed, GPM_RNG_REFERENCE);
            float ms = 0.f;
            if (rc == GPM_OK) rc = gpm_run(ctx, &ms);
            if (rc != GPM_OK) { set_err(std::string("reference view ") + std::to_string(ref) + ": " + gpm_last_error());  failed++;  break; }
            sweep_ms[job] = ms;  views_used[job] = nv;  device_of[job] = dev;
            float* o4 = d->out_norm4 ? d->out_norm4 + (size_t)job * npix * 4 : nullptr;
            float* oc = d->out_cost ? d->out_cost + (size_t)job * npix : nullptr;
            if (d->out_dir && !o4) { n4.resize(npix * 4);  o4 = n4.data(); }
            if (o4 || oc) rc = gpm_get_state(ctx, o4, oc, 0);
            if (rc == GPM_OK && d->out_dir) {
                // <out_dir>/<ref as 8 digits>/disp.dmb, normals.dmb — what the per-image processes of the script leave for fusibile
                char dir[1024];
                snprintf(dir, sizeof(dir), "%s/%08d", d->out_dir, ref);
                mkdir(d->out_dir, 0777);  mkdir(dir, 0777);
                const std::string dp = std::string(dir) + "/disp.dmb", np_ = std::string(dir) + "/normals.dmb";
                rc = gpm_write_result_dmb(dp.c_str(), np_.c_str(), o4, H, W);
            }
            if (rc != GPM_OK) { set_err("gpm_batch_run: could not store the result of reference view " + std::to_string(ref));  failed++;  break; }
        }
        gpm_destroy(ctx);
    };
    std::vector<std::thread> pool;
    for (int k = 0; k < d->n_devices; k++) pool.emplace_back(worker, d->devices[k]);
    for (auto& t : pool) t.join();
    for (int i = 0; i < n; i++) if (registered[i]) cudaHostUnregister(const_cast<float*>(d->images[i]));
    if (stats) {
        stats->jobs_done = 0;  stats->sweep_ms_total = 0.0;
        for (int j = 0; j < d->n_refs; j++) if (device_of[j] >= 0) { stats->jobs_done++;  stats->sweep_ms_total += sweep_ms[j]; }
        if (stats->per_job_sweep_ms) for (int j = 0; j < d->n_refs; j++) stats->per_job_sweep_ms[j] = (float)sweep_ms[j];
        if (stats->per_job_views) for (int j = 0; j < d->n_refs; j++) stats->per_job_views[j] = views_used[j];
        if (stats->per_job_device) for (int j = 0; j < d->n_refs; j++) stats->per_job_device[j] = device_of[j];
    }
    return failed.load() ? GPM_E_STATE : GPM_OK;
}
