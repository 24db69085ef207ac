// examples/shard_host.cpp — a C++ host that runs ONE reference view with its source views sharded over the GPUs of a node,
// using nothing but the C-ABI of include/gipuma_b200.h (no Python, no torch, no CUDA headers): the multi-GPU counterpart of
// what main.cpp does around runcuda() (main.cpp:829-985).  One thread per GPU; rank 0 creates the NCCL unique id
// (gpm_shard_unique_id), every thread joins the communicator (gpm_shard_comm_init) and calls gpm_shard_run — kernels and
// ncclAllGather exchanges all happen behind the ABI.  Input: a scene file written by tools/dump_scene.py; output: the
// LineState arrays (norm4, cost) of rank 0, identical to a single-GPU gpm_run over all views.
//   g++ -std=c++17 -Iinclude examples/shard_host.cpp -Lgipuma_b200 -lgipuma_b200 -lpthread -Wl,-rpath,'$ORIGIN/../gipuma_b200' -o examples/shard_host
//   examples/shard_host scene.bin out.bin [n_gpus]
#include "gipuma_b200.h"

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct SceneFile {                       // layout written by tools/dump_scene.py
    int W = 0, H = 0, n_images = 0, n_views = 0;
    unsigned long long seed = 0;
    gpm_params params{};
    std::vector<int> subset;             // n_views indices into images / cameras (index 0 = reference)
    std::vector<gpm_camera> cameras;     // n_images
    std::vector<float> images;           // n_images * H * W
};

static bool load(const char* path, SceneFile& s)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    int hdr[4];
    bool ok = fread(hdr, sizeof(int), 4, f) == 4 && fread(&s.seed, sizeof(s.seed), 1, f) == 1 && fread(&s.params, sizeof(s.params), 1, f) == 1;
    if (ok) {
        s.W = hdr[0];  s.H = hdr[1];  s.n_images = hdr[2];  s.n_views = hdr[3];
        s.subset.resize(s.n_views);  s.cameras.resize(s.n_images);  s.images.resize((size_t)s.n_images * s.W * s.H);
        ok = fread(s.subset.data(), sizeof(int), s.n_views, f) == (size_t)s.n_views &&
             fread(s.cameras.data(), sizeof(gpm_camera), s.n_images, f) == (size_t)s.n_images &&
             fread(s.images.data(), sizeof(float), s.images.size(), f) == s.images.size();
    }
    fclose(f);
    return ok;
}

struct Rendezvous {                      // what MPI_Bcast / a socket would do between processes
    std::mutex m;
    std::condition_variable cv;
    bool ready = false;
    char id[128];
};

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.bin out.bin [n_gpus]\n", argv[0]);  return 2; }
    SceneFile sc;
    if (!load(argv[1], sc)) { fprintf(stderr, "cannot read %s\n", argv[1]);  return 2; }
    const int world = argc > 3 ? atoi(argv[3]) : 1;
    if (world < 1 || world > 8 || world > sc.n_views) { fprintf(stderr, "bad GPU count\n");  return 2; }
    const size_t npix = (size_t)sc.W * sc.H;
    std::vector<float> norm4(npix * 4), cost(npix);
    std::vector<int> rc(world, 0);
    std::vector<float> ms(world, 0.f);
    std::vector<std::string> err(world);
    Rendezvous rv;

    auto rank_main = [&](int rank) {
        auto fail = [&](const char* what) { rc[rank] = 1;  err[rank] = std::string(what) + ": " + gpm_last_error(); };
        // contiguous, balanced share of viewSelectionSubset (main.cpp:888-892)
        const int base = sc.n_views / world, extra = sc.n_views % world;
        const int first = rank * base + (rank < extra ? rank : extra), count = base + (rank < extra ? 1 : 0);
        gpm_ctx* ctx = nullptr;
        if (gpm_create(&ctx, rank, sc.W, sc.H, count) != GPM_OK) {
            fail("gpm_create");
            if (rank == 0) { std::lock_guard<std::mutex> l(rv.m);  rv.ready = true;  rv.cv.notify_all(); }      // do not leave the others waiting
            return;
        }
        bool ok = gpm_set_params(ctx, &sc.params) == GPM_OK &&
                  gpm_set_reference(ctx, sc.images.data(), 0, 0, &sc.cameras[0]) == GPM_OK;
        for (int v = 0; ok && v < count; v++) {
            const int idx = sc.subset[first + v];
            ok = gpm_set_view(ctx, v, sc.images.data() + (size_t)idx * npix, 0, 0, &sc.cameras[idx]) == GPM_OK;
        }
        ok = ok && gpm_set_num_views(ctx, count) == GPM_OK && gpm_set_rng(ctx, sc.seed, GPM_RNG_REFERENCE) == GPM_OK;
        // rendezvous on the 128-byte NCCL id
        if (rank == 0) {
            if (ok && world > 1) ok = gpm_shard_unique_id(rv.id) == GPM_OK;
            std::lock_guard<std::mutex> l(rv.m);
            rv.ready = true;
            rv.cv.notify_all();
        } else {
            std::unique_lock<std::mutex> l(rv.m);
            rv.cv.wait(l, [&] { return rv.ready; });
        }
        if (!ok) { fail("set-up");  gpm_destroy(ctx);  return; }
        if (gpm_shard_comm_init(ctx, world > 1 ? rv.id : nullptr, rank, world) != GPM_OK) { fail("gpm_shard_comm_init");  gpm_destroy(ctx);  return; }
        if (gpm_shard_run(ctx, &ms[rank]) != GPM_OK) { fail("gpm_shard_run");  gpm_destroy(ctx);  return; }
        if (rank == 0 && gpm_get_state(ctx, norm4.data(), cost.data(), 0) != GPM_OK) fail("gpm_get_state");
        gpm_destroy(ctx);
    };
    std::vector<std::thread> ranks;
    for (int r = 0; r < world; r++) ranks.emplace_back(rank_main, r);
    for (auto& t : ranks) t.join();
    for (int r = 0; r < world; r++)
        if (rc[r]) { fprintf(stderr, "rank %d failed: %s\n", r, err[r].c_str());  return 1; }
    float worst = 0.f;
    for (float v : ms) worst = v > worst ? v : worst;
    FILE* f = fopen(argv[2], "wb");
    if (!f || fwrite(norm4.data(), sizeof(float), norm4.size(), f) != norm4.size() || fwrite(cost.data(), sizeof(float), cost.size(), f) != cost.size()) {
        fprintf(stderr, "cannot write %s\n", argv[2]);
        return 1;
    }
    fclose(f);
    printf("{\"world\": %d, \"views\": %d, \"sweep_ms_max_over_ranks\": %.3f, \"mpixel_iters_per_s\": %.3f}\n", world, sc.n_views, worst,
           (double)npix * sc.params.iterations / 1e3 / worst);
    return 0;
}
