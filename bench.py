#!/usr/bin/env python
"""bench.py -- Voronoi cells/s of the hot path on N B200s of one node.

Metric (BASELINE.json): Voronoi cells/sec, device-timed.  A *step* is one CVT Lloyd
iteration over the whole seed set: counting-sort grid build -> kNN -> warp-per-cell
clipping with the security radius (+ re-launches of uncertified cells) -> fused fp64
volume/centroid integration -> seed update.  Workload at N=1: BASELINE configs[1]
("10M uniform seeds, periodic unit cube, CVT Lloyd iterations on 1 B200"); at N>1 the
weak-scaling shape of configs[4] (fixed seeds per GPU, every rank holds all seeds, cells
sharded by seed-index range, updated seeds all-gathered over NCCL each iteration).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--seeds S] [--impl ours|reference]

`--impl reference` times the reference's own CPU implementation (oracle/_ref, the
unmodified geogram sources: PeriodicDelaunay3d + copy_Laguerre_cell_from_Delaunay +
compute_mg) with all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "voronoi_cells_per_sec"
UNIT = "cells/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        threading.Thread.__init__(self)
        self.daemon = True
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.check_output(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                     "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names)
                   if any(s[3 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples),
                "samples": len(sm), "reasons": reasons}


def cpu_reference_rate(sample_n, steps, warmup, seed=12345):
    """Reference CPU path on a bounded sample: periodic Lloyd iterations through oracle A
    (unmodified geogram).  Falls back to the restatement oracle if oracle/_ref is absent."""
    import numpy as np
    import oracles
    xyz = oracles.uniform_seeds(sample_n, seed)
    port = oracles.Port()
    if oracles.have_reference():
        ref = oracles.Reference()
        kind, cores = "reference", ref.num_threads()

        def one(x):
            A = ref.cells_delaunay(x, periodic=True)
            return port.lloyd_update(x, A.m, A.mg, True, 1.0)
    else:
        kind, cores = "port", port.num_threads()

        def one(x):
            A = port.cells(x, periodic=True, k0=20)
            return port.lloyd_update(x, A.m, A.mg, True, 1.0)
    for _ in range(warmup):
        xyz = one(xyz)
    t0 = time.time()
    for _ in range(steps):
        xyz = one(xyz)
    dt = time.time() - t0
    return sample_n * steps / dt, dt / steps, kind, cores


def run_reference(args, rank):
    if rank != 0:
        return
    sample_n = args.cpu_sample
    rate, sps, kind, cores = cpu_reference_rate(sample_n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "uniform seeds, periodic unit cube, CVT Lloyd iterations "
                               "(BASELINE configs[1]); bounded sample per step",
                   "seeds_per_step": sample_n},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d uniform seeds per step, periodic; PeriodicDelaunay3d + "
                                   "copy_Laguerre_cell_from_Delaunay + compute_mg + update" % sample_n},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--seeds", type=int, default=0, help="seeds per GPU (default 10M at N=1, 12.5M at N>1)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=1000000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--k0", type=int, default=0)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--ppc", type=float, default=0.0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    # stdout carries exactly ONE JSON line: libraries that chat on stdout (NCCL prints its version
    # there) are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    from geogram_b200 import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n_per = args.seeds or (10000000 if world == 1 else 12500000)
    n_total = n_per * world

    # synthetic seeds: uniform in the periodic unit cube, generated on the device
    gen = torch.Generator(device=dev)
    gen.manual_seed(12345 + rank)
    mine = torch.rand((n_per, 3), dtype=torch.float64, device=dev, generator=gen)
    full = torch.empty((n_total, 3), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(full, mine)
    else:
        full.copy_(mine)
    del mine

    stream = torch.cuda.current_stream()
    ctx = Context(local_rank)
    ctx.set_domain_periodic(1.0)
    if args.ppc > 0:
        ctx.set_points_per_cell(args.ppc)
    from geogram_b200.sharding import ShardedLloyd
    job = ShardedLloyd(ctx, n_total, dist if world > 1 else None, dev, stream)
    lo, hi = job.first, job.last
    # N>1: rows re-ordered once into x-quantile slabs, so that rank r owns a spatial slab
    job.set_seeds_device(full, slabs=world > 1)
    del full

    def step():
        job.step(k0=args.k0, mode=args.mode)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    launches0 = ctx.stats()["gpu_launches"]
    job.phase = {k: 0.0 for k in job.phase}
    job.comm_ms()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.summary() if sampler else None
    launches = ctx.stats()["gpu_launches"] - launches0
    last = ctx.stats()
    phase = dict(job.phase)
    phase["ms_allgather"] = job.comm_ms()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n_total * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers (pinned): every step uploads this
    #      rank's rows (H2D), exchanges them over NVLink (N>1), runs the Lloyd iteration and reads
    #      this rank's updated rows back (D2H) ----
    e2e_steps = min(args.steps, 5)
    host_in = torch.empty((hi - lo, 3), dtype=torch.float64).pin_memory()
    host_out = torch.empty((hi - lo, 3), dtype=torch.float64).pin_memory()
    job.get_shard_host(host_in)        # the seeds as relaxed by the timed steps

    def e2e_step():
        job.set_shard_host(host_in)
        job.step(k0=args.k0, mode=args.mode)
        job.get_shard_host(host_out)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = n_total * e2e_steps / e2e_s
    del host_in, host_out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per launch), HBM peak from MEASURED_PEAKS.json ----
    # The dominant kernel is cells_fused_kernel (one launch per step).  Its duration is
    # measured live with CUDA events recorded by the library on the launching stream
    # (stats ms_clip).  Algorithmic HBM bytes per cell: read Seed 32 + perm 4
    # + cell_start 4/ppc, write m 8 + mg 24 + status 1.
    peak, peak_src = load_peaks()
    per_step = {k: v / args.steps for k, v in phase.items()}
    n_cells = n_per
    ppc_used = args.ppc if args.ppc > 0 else 8.0
    bytes_per_cell = 32.0 + 4.0 + 4.0 / ppc_used + 8.0 + 24.0 + 1.0
    fused = last["k_pass"][0] == 0
    if fused:
        kernel, kms = "cells_fused_kernel", per_step["ms_clip"]
        alg_bytes = bytes_per_cell * n_cells
    else:
        k_eff = last["k_pass"][0] + 1
        dom = max(("ms_knn", "ms_clip"), key=lambda k: per_step[k])
        kernel = {"ms_knn": "knn_kernel", "ms_clip": "cells_knn_kernel"}[dom]
        kms = per_step[dom]
        alg_bytes = ((32.0 + 4 * k_eff) if dom == "ms_knn" else (69.0 + 4 * k_eff)) * n_cells
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    traffic = None
    issue = None
    tpath = os.path.join(ROOT, "profiles", "r01_fused_traffic.json")
    if fused and os.path.exists(tpath):
        tj = json.load(open(tpath))
        if int(tj.get("seeds", 0)) == n_cells:
            traffic = tj["dram_bytes_per_launch"]     # ncu --set full, dram__bytes_read+write
        if "warp_inst_per_launch" in tj:
            # issue-slot roofline (north_star: "FP32 issue-rate roofline"): warp instructions per
            # cell from the committed ncu capture x the cells of one launch / the LIVE kernel
            # time, against 4 schedulers x SMs x max SM clock (one warp instruction per cycle each)
            props = torch.cuda.get_device_properties(local_rank)
            ipc = tj["warp_inst_per_launch"] / float(tj["seeds"])
            peak_issue = 4.0 * props.multi_processor_count * (clocks or {}).get("sm_max_mhz", 1965.0) * 1e6
            ach = ipc * n_cells / (kms * 1e-3)
            issue = {"warp_inst_per_cell": ipc, "achieved_ginst_s": ach / 1e9,
                     "peak_ginst_s": peak_issue / 1e9, "frac": ach / peak_issue,
                     "thread_inst_per_warp_inst": tj.get("thread_inst_per_warp_inst"),
                     "source": "instruction count: " + tj["source"] + "; time: live CUDA events"}
    # flop model of SURVEY 8(d): 12 P + 7 C + 64 V + 45 I per cell, from the event counters of
    # one untimed compute_cells() call on the same seeds
    ctx.compute_cells()
    ev = ctx.stats()
    flops = 12.0 * ev["sum_planes"] + 7.0 * ev["sum_tests"] + 64.0 * ev["sum_created"] + 45.0 * ev["sum_tets"]
    roofline = {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms_per_launch": kms, "algorithmic_bytes_per_cell": bytes_per_cell,
                "flop_model": {"sum_planes": ev["sum_planes"], "sum_tests": ev["sum_tests"],
                               "sum_created": ev["sum_created"], "sum_tets": ev["sum_tets"],
                               "grid_scans_per_cell": ev["sum_scans"] / max(n_cells, 1),
                               "flop_per_cell": flops / max(n_cells, 1),
                               "achieved_tflops": flops / (kms * 1e-3) / 1e12,
                               "fp64_peak_tflops_nominal": 37.0},
                "issue_rate": issue,
                "note": "issue/latency-bound kernel (see issue_rate; fp64 pipe ~10%); neither DRAM nor "
                        "tensor bound: the hbm fraction is reported because the contract asks for one, "
                        "see DESIGN.md 4.2 and profiles/"}

    cpu = None
    if world == 1 and not args.no_cpu:
        rate, sps, kind, cores = cpu_reference_rate(args.cpu_sample, 2, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d uniform seeds, periodic, 2 Lloyd iterations "
                         "(Delaunay + cell extraction + compute_mg + update)" % args.cpu_sample}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: uniform seeds, periodic unit cube, one CVT "
                               "Lloyd iteration per step" if world == 1 else
                               "BASELINE configs[4]: weak scaling, uniform periodic, one Lloyd "
                               "iteration per step, x-quantile slab per rank, seeds all-gathered over NCCL",
                   "seeds_per_gpu": n_per, "seeds_total": n_total,
                   "k_pass": last["k_pass"][: last["passes"]],
                   "cells_per_pass": last["n_pass"][: last["passes"]],
                   "grid": last["grid_dim"], "mode": args.mode,
                   "l2": "inputs (>= 240 MB/step) exceed the 126 MB L2"},
        "phase_ms_per_step": per_step,
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps,
                "h2d_bytes_per_step": 24 * n_total, "d2h_bytes_per_step": 24 * n_total,
                "note": "whole job: every rank moves its own rows (24 B/seed each way)"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
