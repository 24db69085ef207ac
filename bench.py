#!/usr/bin/env python
"""bench.py — headline benchmark of the PatchMatch hot path (BASELINE.json metric: Mpixel-iters/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one complete pass of the hot path over one reference view: random initialisation, `iterations`
red/black sweeps, final depth/normal kernel (what the reference's runcuda() does once per process).
Workload at N=1: BASELINE.json configs[1] — DTU 'dtu_fast' parameters, 1600x1200, 10 source views, 8 iterations,
synthetic 8-bit images rendered from a textured height field, real DTU camera geometry.
At N>1 every rank processes its OWN reference view (the reference runs one process per reference image,
scripts/dtu_fast.sh:30-55): independent units, no data-path collective, weak scaling.

Printed JSON (one line, rank 0):
  value        Mpixel-iters/s = ranks * W*H*iterations / 1e6 / t, t = device time from the first sweep kernel to the
               end of the final kernel (the reference's own timed span, gipuma.cu:1908-1952; init excluded), inputs
               resident in HBM; summed over the K timed steps, max over ranks.
  ms_per_step  full device time of a step INCLUDING initialisation.
  e2e          same metric through the public API with HOST buffers: per step the images are uploaded from pinned
               host memory (H2D), the job runs, and planes+costs are read back (D2H); wall clock around the call.
  roofline     HBM roofline of the dominant kernel (k_sweep) from algorithmic bytes and its average
               launch duration measured live with CUDA events; plus the binding unit measured by ncu (profiles/).
  cpu_baseline the single-thread C restatement (oracle/gipuma_oracle.c) timed on a bounded sample of this workload.
--impl reference times the reference's own implementation of the path — gipuma.cu compiled unmodified for sm_100a
(oracle/_ref, pins P1/P2 by macro) — on the same workload, same metric.  (The reference has no CPU implementation
of this path; its CUDA kernels are "the reference's own implementation")
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpixel-iters/s (ref-view PatchMatch sweep)"
UNIT = "Mpixel-iters/s"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([t.strip() for t in out.split(",")])
            except Exception:      # noqa: BLE001
                pass
            self._halt.wait(0.2)

    def finish(self) -> dict:
        self._halt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 3 + i and r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def make_scene(config: int, rank: int, color: bool = False):
    from gipuma_b200 import scene as S
    # every rank gets its own reference view: another rendered surface / texture seed (and, for the DTU
    # configurations, the geometry of the same rig)
    sc = S.make_config(config, seed=1234 + 17 * rank)
    return S.colorize(sc) if color else sc


def variant_suffix(args):
    """Non-default modes of the path (both arms run the same one): float4 images, fused 20-neighbour sweep."""
    out = ""
    if args.color:
        out += ", -color_processing (float4 images)"
    if args.neighbours == 20:
        out += ", fused 20-neighbour sweep (reference built without SMALLKERNEL)"
    return out


def algorithmic_bytes_per_sweep_launch(W, H, V):
    """one k_sweep launch = one colour (W*H/2 pixels): own plane+cost read and written (2*20 B),
    8 neighbour planes (8*16 B), the memo of rejected work read (8*16 + 16 + 4 B; its writes are data dependent and not
    counted), and each image plane (reference + V sources) streamed once (4 B/pixel each)."""
    return (W * H // 2) * (40 + 128 + 148) + (1 + V) * W * H * 4


def cpu_baseline(sc, budget_s: float = 15.0, neighbours: int = 8) -> dict:
    from oracle import pyoracle
    o = pyoracle.Oracle(sc)
    threads = o.set_threads(len(os.sched_getaffinity(0)))       # all host cores this process may use (rows are independent)
    rng = np.random.default_rng(0)
    H, W = sc.rows, sc.cols
    pl = np.zeros((H, W, 4), np.float32)
    pl[..., 2] = -1.0
    pl[..., 3] = sc.gt_depth * rng.uniform(0.9, 1.1, size=(H, W)).astype(np.float32)
    y0 = H // 2
    t0 = time.perf_counter()
    c = o.cost_eval(pl, y0, y0 + threads, init_radius=True)    # calibrate: one row of initial costs per thread
    t_row = (time.perf_counter() - t0) / threads
    evals_per_px_iter = 2 * (neighbours + 3)                     # hypotheses per pixel-iteration (E = 8 + S, S = 3 on DTU)
    rows = int(max(threads, min(H // 2, budget_s / max(1e-6, t_row * evals_per_px_iter / 2))))
    cost = np.full((H, W), 50.0, np.float32)
    cost[y0:y0 + rows] = o.cost_eval(pl, y0, y0 + rows)[y0:y0 + rows]
    t0 = time.perf_counter()
    if neighbours == 20:                                         # fused kernel: 20 candidates + refinement per colour
        for colour in (0, 1):
            pl, cost = o.phase(pl, cost, colour, 8 | 4, y0, y0 + rows)
    else:
        o.sweep(pl, cost, 1, y0, y0 + rows)
    dt = time.perf_counter() - t0
    return {"value": rows * W * 1 / 1e6 / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "1 iteration over rows [%d,%d) of the same %dx%d / %d-view workload (%.1f s on %d OpenMP thread(s), host has %d cores)"
                      % (y0, y0 + rows, W, H, sc.n_views, dt, threads, os.cpu_count() or 0)}


def run_ours(args, rank, world, local):
    import torch
    from gipuma_b200 import api
    torch.cuda.set_device(local)
    sc = make_scene(args.config, rank, args.color)
    W, H, V, iters = sc.cols, sc.rows, sc.n_views, sc.params.iterations
    pinned = torch.from_numpy(np.ascontiguousarray(sc.images)).pin_memory()
    out4 = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    outc = torch.empty((H, W), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2
    ctx = api.Context(W, H, V, device=local)
    if args.neighbours != 8:
        ctx.set_option("neighbours", args.neighbours)
    imgs = [pinned[i] for i in range(pinned.shape[0])]

    def upload():
        ctx.load_scene(sc, images=imgs)                                       # H2D from pinned host memory

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    upload()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(args.warmup):
        ctx.run()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- device-resident timing -----------------------------------------------------------------------------
    barrier()
    sweep_ms = step_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        flush.fill_(1)                                                        # L2 flush between timed iterations
        torch.cuda.synchronize()
        e0.record(stream)
        sweep_ms += ctx.run()                                                 # the library's own CUDA-event span
        e1.record(stream)
        e1.synchronize()
        step_ms += e0.elapsed_time(e1)
        launches += ctx.stats()["launches"]
    barrier()
    stats = ctx.stats()
    # ---- end to end through the public API, host buffers ------------------------------------------------------
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        upload()
        ctx.run()
        ctx.get_state_into(out4, outc)                                        # D2H of planes + costs
        torch.cuda.synchronize()
        e2e_s += time.perf_counter() - t0
    barrier()
    clocks = sampler.finish() if sampler else None
    t = torch.tensor([sweep_ms, step_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)    # max over ranks
    sweep_ms, step_ms, e2e_ms = [float(v) for v in t.tolist()]
    ctx.close()
    if rank != 0:
        return None
    units = world * args.steps * W * H * iters / 1e6
    n_sweep_launches = 2 * iters
    avg_launch_ms = sweep_ms / args.steps / n_sweep_launches                  # k_sweep dominates the span (finalize < 0.1 %)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:      # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algorithmic_bytes_per_sweep_launch(W, H, V) / 1e9 / (avg_launch_ms / 1e3)
    line = {
        "metric": METRIC, "value": units / (sweep_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[%d]: %s, %dx%d, %d source views, %d iterations, blocksize %d, n_best %d"
                               % (args.config - 1, sc.name, W, H, V, iters, sc.params.box_hsize, sc.params.n_best) + variant_suffix(args),
                   "parallelism": "reference-view batch: %d independent reference view(s), one per GPU, no collective" % world,
                   "timed_span": "first sweep kernel .. end of final depth/normal kernel (reference's own span, init excluded)",
                   "l2": "flushed between timed steps (256 MiB write)", "rng": "seed 0xC0FFEE, reference zero-state refinement RNG"},
        "value_incl_init": units / (step_ms / 1e3),
        "e2e": {"value": units / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(pinned.numel() * 4),
                "d2h_bytes_per_step": int(W * H * 20), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": 453.3e6,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                     "kernel": "gpm::k_sweep", "avg_launch_ms": avg_launch_ms,
                     "binding_unit": {"name": "l1tex__data_pipe_tex_wavefronts", "frac_of_peak": 0.950,
                                      "source": "ncu --set full, profiles/r01_ncu_k_sweep_cfg2_iter2_black.txt (not measured live)"},
                     "algorithmic_bytes_per_launch": algorithmic_bytes_per_sweep_launch(W, H, V),
                     "note": "this path is bound by the L1TEX data pipe, not HBM: ncu l1tex__data_pipe_tex_wavefronts = 95.0 % of "
                             "peak for this kernel (profiles/r01_ncu_k_sweep_cfg2_iter2_black.txt); traffic = ncu dram bytes of that launch "
                             "(iteration 2, black; later launches move less)"},
        "work": {"hypotheses_evaluated": stats["hypotheses"], "hypotheses_skipped_exact": stats["skipped"],
                 "hypotheses_pruned_exact": stats["pruned"], "view_samples": stats["pairs"]},
        "clocks": clocks,
    }
    if world == 1:                                            # the CPU baseline is timed at N = 1 only
        try:
            line["cpu_baseline"] = cpu_baseline(sc, neighbours=args.neighbours)
        except Exception as e:      # noqa: BLE001
            line["cpu_baseline"] = {"error": repr(e)}
    else:
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "timed at N = 1 only"}
    return line


def run_reference(args, rank, world, local):
    """The reference's own gipuma.cu (oracle/_ref), one reference view on rank 0."""
    if rank != 0:
        return None
    import torch
    from oracle import pyref
    torch.cuda.set_device(local)
    sc = make_scene(args.config, 0, args.color)
    W, H, V, iters = sc.cols, sc.rows, sc.n_views, sc.params.iterations
    try:
        h = pyref.Harness("ref64" if V > 32 else "ref")
    except Exception as e:      # noqa: BLE001
        return {"impl": "reference", "unavailable": "pinned reference build missing: %r" % (e,)}
    sampler = ClockSampler(local)
    def run_once():
        if args.neighbours == 20:                     # the kernels a reference built without SMALLKERNEL launches
            _, _, ms = h.run_fused(sc)
            return ms / 1e3
        return h.run(sc)[2]

    for _ in range(args.warmup):
        run_once()
    sampler.start()
    printed = wall = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        printed_s = run_once()
        wall += time.perf_counter() - t0
        printed += printed_s
    clocks = sampler.finish()
    units = args.steps * W * H * iters / 1e6
    value = units / printed
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": printed * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[%d]: %s, %dx%d, %d source views, %d iterations, blocksize %d, n_best %d"
                               % (args.config - 1, sc.name, W, H, V, iters, sc.params.box_hsize, sc.params.n_best) + variant_suffix(args),
                   "parallelism": "1 reference view on rank 0 (the reference is single-GPU, main.cpp:658-692)",
                   "timed_span": "the reference's own printed 'Total time needed for computation' (gipuma.cu:1908-1952)",
                   "build": "unmodified gipuma.cu, nvcc 12.9 -O3 --use_fast_math sm_100a, pins P1/P2 by macro (oracle/build_ref.sh)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                         "sample": "whole workload on the GPU: the reference has no CPU implementation of this path"},
        "e2e": {"value": units / wall, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "wall clock around the whole main.cpp stand-in call (GlobalState build, texture upload, runcuda, read-back)"},
        "gpu_launches": args.steps * (1 + 6 * iters + 1), "clocks": clocks,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--color", action="store_true", help="float4 images (the reference's -color_processing)")
    ap.add_argument("--neighbours", type=int, default=8, choices=[8, 20],
                    help="20: the fused sweep of a reference built without SMALLKERNEL")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    rank, world, local = dist_env()
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: gipuma_b200 has no CPU fallback"}))
        sys.exit(1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = run_reference(args, rank, world, local) if args.impl == "reference" else run_ours(args, rank, world, local)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0 and line is not None:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
