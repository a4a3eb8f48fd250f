#!/usr/bin/env python
"""bench.py -- Voronoi cells/s of the hot path on N B200s of one node.

Metric (BASELINE.json): Voronoi cells/sec, device-timed.  A *step* is one CVT Lloyd
iteration over the whole seed set: counting-sort grid build -> neighbour walk + per-seed
cell clipping with the security radius (+ re-launches of the cells a kernel flags) -> fp64
volume/centroid integration -> seed update.  Workload at N=1: BASELINE configs[1]
("10M uniform seeds, periodic unit cube, 10 CVT Lloyd iterations on 1 B200") AS WRITTEN: the
warm-up iterations run on a scratch copy, then the RAW uniform seeds are loaded again and the
timed steps are Lloyd iterations 1..K from them (`ms_iter` lists every iteration, so the
dear first iteration and the relaxed steady state are both visible).  At N>1: the
weak-scaling shape of configs[4] (12.5 M seeds per GPU; every rank keeps one x-slab plus a halo of
ghost seeds exchanged with ncclSend/Recv inside the library, `--sharding replicated` = the
round-1 all-gather of all seeds).  At N=1 the line also carries short
sub-results for configs[2] (Laguerre pass + OT solve), configs[3] (clustered seeds) and the
12.5 M-seed point of the weak-scaling curve under `extra_configs`.

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
        self.samples = []       # (perf_counter, fields)
        self.t_begin = None
        self.proc = None
        self.stop_flag = False
        self.first = threading.Event()

    def start(self):
        """ONE nvidia-smi process for the whole run (the recipe's `-lms` loop), started and past its
        initialisation BEFORE the timed region: a fresh nvidia-smi per sample initialises NVML on every
        GPU of the box each time, which can hold up the CUDA calls of the ranks for milliseconds."""
        import shutil
        cmd = ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
               "--format=csv,noheader,nounits", "-lms", "100"]
        if shutil.which("stdbuf"):
            cmd = ["stdbuf", "-oL"] + cmd      # line-buffered although stdout is a pipe
        try:
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        threading.Thread.start(self)
        if self.proc is not None:
            t0 = time.perf_counter()
            while not self.first.wait(timeout=0.1):
                if self.proc.poll() is not None or time.perf_counter() - t0 > 8.0:
                    # no line came through (missing tool, buffered pipe): one nvidia-smi call per sample instead
                    self.proc.kill()
                    break

    def begin(self):
        """the timed region starts now: only samples taken from here on count"""
        self.t_begin = time.perf_counter()

    def run(self):
        if self.proc is not None:
            for line in self.proc.stdout:
                f = [x.strip() for x in line.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append((time.perf_counter(), f))
                    self.first.set()
            if self.first.is_set():
                return
        while not self.stop_flag:
            try:
                out = subprocess.check_output(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                     "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.samples.append((time.perf_counter(), [x.strip() for x in out.split(",")]))
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        t_end = time.perf_counter()
        self.stop_flag = True
        if self.proc is not None and self.proc.poll() is None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        self.join(timeout=6)
        t0 = self.t_begin if self.t_begin is not None else 0.0
        inside = [f for (t, f) in self.samples if t0 <= t <= t_end]
        self.samples = inside
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


def status_hist(st):
    """per-status cell counts from a status byte array (torch tensor on any device)"""
    names = {1: "uncertified", 2: "uncertain", 4: "empty", 8: "boxface", 16: "overflow",
             32: "duplicate", 64: "degenerate"}
    out = {"ok": int((st == 0).sum().item())}
    for bit, nm in names.items():
        out[nm] = int(((st & bit) != 0).sum().item())
    return out


def clustered_seeds_device(torch, n, dev, seed=12345, nclusters=32):
    """BASELINE configs[3] generator on the device: mixture of 32 isotropic Gaussians, centres
    0.1 + 0.8 U, sigma 0.005 + 0.045 U, rejected to (0,1)^3."""
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    c = 0.1 + 0.8 * torch.rand((nclusters, 3), dtype=torch.float64, device=dev, generator=gen)
    s = 0.005 + 0.045 * torch.rand((nclusters,), dtype=torch.float64, device=dev, generator=gen)
    out = []
    have = 0
    while have < n:
        k = torch.randint(0, nclusters, (n,), device=dev, generator=gen)
        p = c[k] + s[k, None] * torch.randn((n, 3), dtype=torch.float64, device=dev, generator=gen)
        p = p[((p > 0) & (p < 1)).all(dim=1)]
        out.append(p)
        have += p.shape[0]
    return torch.cat(out)[:n].contiguous()


def extra_configs(torch, dev, ctx, xyz_relaxed, n, args):
    """Short driver-visible sub-results for the configs the headline does not time (N=1 only)."""
    from geogram_b200 import Context
    import numpy as np
    res = {}
    # ---- configs[2]: Laguerre cells with random weights on the relaxed seeds: volumes + weight
    #      gradient g_i = vol_i - 1/n (one pass), then the full semi-discrete OT solve ----
    h = n ** (-1.0 / 3.0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    w = 0.25 * h * h * torch.rand((n,), dtype=torch.float64, device=dev, generator=gen)
    ctx.set_points_device(n, xyz_relaxed.data_ptr(), w.data_ptr())
    ctx.compute_cells(want_faces=True)
    ctx.compute_cells(want_faces=True)              # second pass = warm
    s = ctx.stats()
    m = torch.empty((n,), dtype=torch.float64, device=dev)
    st = torch.empty((n,), dtype=torch.uint8, device=dev)
    ctx.get_cells_device(m.data_ptr(), None, st.data_ptr())
    grad = m - 1.0 / n
    lag = {"workload": "BASELINE configs[2]: %d seeds after the timed Lloyd iterations, weights 0.25 h^2 U(0,1), "
                       "periodic; volumes + facet areas + gradient vol_i - 1/n, one pass" % n,
           "ms_grid": s["ms_grid"], "ms_cells": s["ms_clip"] + s["ms_knn"],
           "cells_per_s": n / (1e-3 * (s["ms_grid"] + s["ms_clip"] + s["ms_knn"])),
           "status": status_hist(st), "sum_m_minus_1": float(m.sum().item() - 1.0),
           "grad_sum": float(grad.sum().item()), "grad_norm2": float(torch.sqrt((grad * grad).sum()).item()),
           "cells_per_pass": s["n_pass"][: s["passes"]]}
    res["configs2_laguerre_pass"] = lag
    if not args.no_otm:
        ctx.set_points_device(n, xyz_relaxed.data_ptr())
        w0 = w.cpu().numpy()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, r = ctx.otm_newton(None, w0=w0, epsilon=1e-3)
        wall = time.perf_counter() - t0
        r["wall_s"] = wall
        r["workload"] = "BASELINE configs[2]: semi-discrete OT solve, uniform target, start weights 0.25 h^2 U(0,1), epsilon 1e-3"
        res["configs2_otm_solve"] = r
    del w, m, st, grad
    # ---- the public NearestNeighborSearch path: grid build + exact kNN, k = 21
    #      (Delaunay_NearestNeighbors' default 20 + the point itself), HBM roofline of SURVEY 8(d):
    #      (76 + 4k) algorithmic bytes per seed over (t_bin + t_knn) ----
    if not args.no_knn:
        k = 21
        c3 = Context(dev.index)
        c3.set_domain_box([0, 0, 0, 1, 1, 1])
        xyz_raw = torch.rand((n, 3), dtype=torch.float64, device=dev, generator=gen)
        nbr = torch.empty((n, k), dtype=torch.int32, device=dev)
        c3.set_points_device(n, xyz_raw.data_ptr())
        c3.knn_device(k, nbr.data_ptr())
        c3.set_points_device(n, xyz_raw.data_ptr())
        g_ms = c3.stats()["ms_grid"]
        c3.knn_device(k, nbr.data_ptr())
        s = c3.stats()
        hbm_peak, _ = load_peaks()
        bw = (76.0 + 4.0 * k) * n / ((g_ms + s["ms_knn"]) * 1e-3) / 1e9
        res["knn_k21"] = {"workload": "grid build + exact kNN, k=21, %d uniform seeds in the unit box (ids only)" % n,
                          "ms_grid": g_ms, "ms_knn": s["ms_knn"], "queries_per_s": n / ((g_ms + s["ms_knn"]) * 1e-3),
                          "algorithmic_bytes_per_seed": 76.0 + 4.0 * k, "achieved_gbs": bw, "hbm_peak_gbs": hbm_peak,
                          "frac": bw / hbm_peak}
        c3.close()
        del xyz_raw, nbr
    # ---- configs[3]: clustered Gaussian-mixture seeds in the unit box (security-radius retries,
    #      two-level grid), one cell pass ----
    nc = args.clustered
    if nc > 0:
        xyz = clustered_seeds_device(torch, nc, dev)
        c2 = Context(dev.index)
        c2.set_domain_box([0, 0, 0, 1, 1, 1])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c2.set_points_device(nc, xyz.data_ptr())
        c2.compute_cells()
        wall = time.perf_counter() - t0
        s = c2.stats()
        m = torch.empty((nc,), dtype=torch.float64, device=dev)
        st = torch.empty((nc,), dtype=torch.uint8, device=dev)
        c2.get_cells_device(m.data_ptr(), None, st.data_ptr())
        res["configs3_clustered"] = {
            "workload": "BASELINE configs[3] on one GPU: %d Gaussian-mixture seeds (32 clusters), unit box, one cell pass" % nc,
            "ms_grid": s["ms_grid"], "ms_cells": s["ms_clip"] + s["ms_knn"], "wall_s": wall,
            "cells_per_s": nc / (1e-3 * (s["ms_grid"] + s["ms_clip"] + s["ms_knn"])),
            "status": status_hist(st), "sum_m_minus_1": float(m.sum().item() - 1.0),
            "cells_per_pass": s["n_pass"][: s["passes"]], "k_pass": s["k_pass"][: s["passes"]],
            "grid": s["grid_dim"]}
        c2.close()
    return res


def run_clustered(args, torch, dist, dev, rank, world, json_fd):
    """BASELINE configs[3]: one cell pass over n Gaussian-mixture seeds in the unit box, the cells
    sharded over the ranks by x-slab (b200vor_set_query_range); device time = max over ranks."""
    from geogram_b200 import Context
    from geogram_b200.sharding import shard_range, slab_order
    n = args.seeds or 10000000
    xyz = clustered_seeds_device(torch, n, dev)          # same generator seed on every rank: same seeds
    xyz = xyz[slab_order(xyz)].contiguous()
    first, last = shard_range(n, rank, world)
    ctx = Context(dev.index)
    ctx.set_domain_box([0, 0, 0, 1, 1, 1])
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    times = []
    for it in range(args.warmup + args.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ctx.set_points_device(n, xyz.data_ptr())
        if world > 1:
            ctx.set_query_range(first, last)
        ctx.compute_cells()
        b.record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            times.append(a.elapsed_time(b))
    s = ctx.stats()
    m = torch.empty((n,), dtype=torch.float64, device=dev)
    st = torch.empty((n,), dtype=torch.uint8, device=dev)
    ctx.get_cells_device(m.data_ptr(), None, st.data_ptr())
    m, st = m[first:last], st[first:last]
    names = {1: "uncertified", 2: "uncertain", 4: "empty", 8: "boxface", 16: "overflow", 32: "duplicate", 64: "degenerate"}
    counts = torch.tensor([float((st == 0).sum().item())] + [float(((st & bit) != 0).sum().item()) for bit in names]
                          + [float(m.sum().item())], dtype=torch.float64, device=dev)
    t = torch.tensor([sum(times) / len(times), s["ms_grid"], s["ms_clip"] + s["ms_knn"]], dtype=torch.float64, device=dev)
    tmin = t.clone()
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    if rank == 0:
        c = counts.tolist()
        line = {"metric": METRIC, "value": n / (float(t[0]) * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t[0]), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "BASELINE configs[3]: %d Gaussian-mixture seeds (32 clusters, sigma 0.005..0.05), unit box, "
                                       "one pass = grid build + cells (volume, centroid); cells sharded by x-slab over %d GPU(s), "
                                       "every rank bins all seeds" % (n, world),
                           "seeds_total": n, "grid": s["grid_dim"], "cells_per_pass_rank0": s["n_pass"][: s["passes"]]},
                "phase_ms_per_step": {"ms_grid_max": float(t[1]), "ms_cells_max": float(t[2]), "ms_cells_min": float(tmin[2])},
                "status": dict([("ok", int(c[0]))] + [(nm, int(c[1 + k])) for k, nm in enumerate(names.values())]),
                "sum_m_minus_1": c[-1] - 1.0,
                "cells_per_s_per_gpu": n / (float(t[0]) * 1e-3) / world,
                "gpu_launches": int(s["gpu_launches"])}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def load_inst_counts(n_cells):
    """Per-iteration instruction counts of the cell kernels from the committed ncu capture
    (profiles/r02_inst_by_iter.json, made by tools/ncu_inst_by_iter.sh from THIS bench command:
    same seeds, same iterations -> same counts)."""
    p = os.path.join(ROOT, "profiles", "r02_inst_by_iter.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    if int(d.get("seeds", 0)) != n_cells:
        return None
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--seeds", type=int, default=0, help="seeds per GPU (default 10M at N=1, 12.5M at N>1)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=1000000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs sub-results")
    ap.add_argument("--no-otm", action="store_true", help="skip the OT solve of extra_configs")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-knn", action="store_true", help="skip the kNN sub-result of extra_configs")
    ap.add_argument("--clustered", type=int, default=10000000, help="seeds of the configs[3] sub-result (0 = skip)")
    ap.add_argument("--k0", type=int, default=0)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--ppc", type=float, default=0.0)
    ap.add_argument("--workload", default="lloyd", choices=["lloyd", "clustered"],
                    help="lloyd: the headline (BASELINE configs[1] / [4]); clustered: BASELINE configs[3], one cell pass over "
                         "--seeds Gaussian-mixture seeds IN TOTAL, the cells sharded over the GPUs by x-slab (every rank "
                         "holds all seeds: far-reaching fringe cells need them)")
    ap.add_argument("--sharding", default="slab", choices=["slab", "replicated"],
                    help="N>1: x-slabs with halo exchange inside the library (default) or every rank holds all seeds")
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

    import numpy as np  # noqa: F401
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
    if args.workload == "clustered":
        run_clustered(args, torch, dist, dev, rank, world, json_fd)
        return
    n_per = args.seeds or (10000000 if world == 1 else 12500000)
    n_total = n_per * world

    # synthetic seeds: uniform in the periodic unit cube, generated on the device
    def raw_seeds(n_per_rank):
        gen = torch.Generator(device=dev)
        gen.manual_seed(12345 + rank)
        mine = torch.rand((n_per_rank, 3), dtype=torch.float64, device=dev, generator=gen)
        if world == 1:
            return mine
        full = torch.empty((n_per_rank * world, 3), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(full, mine)
        return full

    full0 = raw_seeds(n_per)       # pristine raw seeds: the timed run starts from these

    stream = torch.cuda.current_stream()
    ctx = Context(local_rank)
    ctx.set_domain_periodic(1.0)
    if args.ppc > 0:
        ctx.set_points_per_cell(args.ppc)
    peaks = ctx.microbench() if rank == 0 else None
    from geogram_b200.sharding import ShardedLloyd, SlabLloyd
    use_slab = world > 1 and args.sharding == "slab"
    if use_slab:
        job = SlabLloyd(ctx, n_total, dist, dev, stream)
    else:
        job = ShardedLloyd(ctx, n_total, dist if world > 1 else None, dev, stream)
    lo, hi = job.first, job.last

    def step():
        job.step(k0=args.k0, mode=args.mode)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up on a scratch copy: Lloyd iterations 1..W from the raw seeds ----
    # N>1: rows re-ordered once into x-quantile slabs, so that rank r owns a spatial slab
    job.set_seeds_device(full0, slabs=world > 1)
    for _ in range(args.warmup):
        step()
    # ---- the timed run: the raw seeds again, Lloyd iterations 1..K ----
    job.set_seeds_device(full0, slabs=world > 1)
    launches0 = ctx.stats()["gpu_launches"]
    job.phase = {k: 0.0 for k in job.phase}
    job.comm_ms()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    if sampler:
        sampler.begin()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    clip_iter = []
    evs[0].record(stream)
    for k in range(args.steps):
        before = job.phase["ms_clip"]
        step()
        clip_iter.append(job.phase["ms_clip"] - before)
        evs[k + 1].record(stream)
    barrier()
    ms = evs[0].elapsed_time(evs[-1])
    ms_iter = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    clocks = sampler.summary() if sampler else None
    launches = ctx.stats()["gpu_launches"] - launches0
    last = ctx.stats()
    phase = dict(job.phase)
    phase["ms_exchange" if use_slab else "ms_allgather"] = job.comm_ms()
    if world > 1:
        t = torch.tensor([ms] + ms_iter, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0].item())
        ms_iter = [float(x) for x in t[1:].tolist()]
    value = n_total * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers (pinned): every step uploads this
    #      rank's rows (H2D), exchanges them over NVLink (N>1), runs the Lloyd iteration and reads
    #      this rank's updated rows back (D2H).  Iterations K+1.. of the same run. ----
    e2e_value, e2e_steps = None, 0
    if not args.no_e2e:
        e2e_steps = args.steps          # the same iterations 1..K as the device-timed `value`
        host_in = torch.empty((hi - lo, 3), dtype=torch.float64).pin_memory()
        host_out = torch.empty((hi - lo, 3), dtype=torch.float64).pin_memory()
        # start again from the raw seeds: this rank's rows of the (slab-ordered) raw array
        raw_rows = full0 if job.order is None else full0[job.order]
        host_in.copy_(raw_rows[lo:hi])
        del raw_rows
        torch.cuda.synchronize()

        bufs = [host_in, host_out]

        def e2e_step():
            job.set_shard_host(bufs[0])
            job.step(k0=args.k0, mode=args.mode)
            job.get_shard_host(bufs[1])
            bufs.reverse()               # the next iteration starts from this one's result (no host copy)

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

    # ---- rooflines of the dominant kernel (the first-pass cell kernel, one launch per step) ----
    # Kernel time: CUDA events recorded by the library on the launching stream (stats ms_clip,
    # summed over the K timed iterations).  Every numerator below is taken at the SAME seed
    # states as the timed launches: the flop-model event counts come from an untimed replay of
    # iterations 1..K (event-counting instance of the same kernel), the instruction counts from
    # the committed ncu capture of this very command (profiles/r02_inst_by_iter.json).
    hbm_peak, hbm_src = load_peaks()
    per_step = {k: v / args.steps for k, v in phase.items()}
    n_cells = n_per
    ppc_used = args.ppc if args.ppc > 0 else 8.0
    bytes_per_cell = 32.0 + 4.0 + 4.0 / ppc_used + 8.0 + 24.0 + 1.0
    kms_total = sum(clip_iter)
    kernel = "cells_tpc_kernel" if last.get("tpc_cells", 0) else "cells_fused_kernel"
    flop_iter = []
    ev_tot = {"sum_planes": 0, "sum_tests": 0, "sum_created": 0, "sum_tets": 0, "sum_scans": 0}
    if world == 1:
        job.set_seeds_device(full0)
        for k in range(args.steps):
            ctx.compute_cells(mode=args.mode)
            ev = ctx.stats()
            flop_iter.append(12.0 * ev["sum_planes"] + 7.0 * ev["sum_tests"] + 64.0 * ev["sum_created"]
                             + 45.0 * ev["sum_tets"])
            for key in ev_tot:
                ev_tot[key] += ev[key]
            step()
    flops = sum(flop_iter)
    ach_tf = flops / (kms_total * 1e-3) / 1e12 if flops else None
    issue = None
    inst = load_inst_counts(n_cells) if world == 1 else None
    if inst and len(inst["iters"]) >= args.steps:
        wi = sum(it["warp_inst"] for it in inst["iters"][: args.steps])
        ti = sum(it["thread_inst"] for it in inst["iters"][: args.steps])
        ach = wi / (kms_total * 1e-3)
        issue = {"warp_inst_per_cell": wi / (float(n_cells) * args.steps),
                 "thread_inst_per_warp_inst": ti / float(wi),
                 "achieved_ginst_s": ach / 1e9, "peak_ginst_s": peaks["issue_ginst_s"],
                 "frac": ach / 1e9 / peaks["issue_ginst_s"],
                 "source": "instruction counts per iteration: " + inst.get("source", "profiles/r02_inst_by_iter.json")
                           + " (same seeds and iterations as this run); time: live CUDA events"}
    traffic = inst.get("dram_bytes_per_launch") if inst else None
    ach_bw = bytes_per_cell * n_cells * args.steps / (kms_total * 1e-3) / 1e9
    roofline = {
        "kernel": kernel, "bound": "fp64", "unit": "TFLOP/s",
        "achieved": ach_tf, "peak": peaks["fp64_tflops"],
        "frac": (ach_tf / peaks["fp64_tflops"]) if ach_tf else None,
        "traffic": traffic,
        "peak_source": "fp64/fp32: FMA-chain microbench in this process (b200vor_microbench); hbm: " + hbm_src,
        "kernel_ms_per_launch": kms_total / args.steps, "kernel_ms_by_iter": clip_iter,
        "flop_model": {"formula": "12 P + 7 C + 64 V + 45 I (SURVEY 8(d)), summed over the K timed iterations "
                                  "(untimed replay with the event-counting kernel instance)",
                       "sum_planes": ev_tot["sum_planes"], "sum_tests": ev_tot["sum_tests"],
                       "sum_created": ev_tot["sum_created"], "sum_tets": ev_tot["sum_tets"],
                       "grid_scans_per_cell": ev_tot["sum_scans"] / max(n_cells * args.steps, 1),
                       "flop_per_cell_by_iter": [f / n_cells for f in flop_iter],
                       "frac_of_fp32_peak": (ach_tf / peaks["fp32_tflops"]) if ach_tf else None},
        "peaks_measured": peaks,
        "hbm": {"achieved": ach_bw, "peak": hbm_peak, "unit": "GB/s", "frac": ach_bw / hbm_peak,
                "algorithmic_bytes_per_cell": bytes_per_cell},
        "issue_rate": issue,
        "note": "the cell kernel is issue/latency bound: `frac` is the SURVEY 8(d) flop model against the measured "
                "FP64 FMA peak, `issue_rate` the share of the 4 x SMs x clock issue slots it fills, `hbm` the "
                "algorithmic bytes against the measured copy bandwidth (DESIGN.md section 4)"}

    cpu = None
    if world == 1 and not args.no_cpu:
        rate, sps, kind, cores = cpu_reference_rate(args.cpu_sample, 2, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d uniform seeds, periodic, 2 Lloyd iterations "
                         "(Delaunay + cell extraction + compute_mg + update)" % args.cpu_sample}

    extra = None
    if world == 1 and not args.no_extra:
        # the seeds as relaxed by the K iterations of the replay above
        xyz_relaxed = job.full.clone()
        ctx.lloyd_device(0, xyz_relaxed.data_ptr())
        extra = extra_configs(torch, dev, ctx, xyz_relaxed, n_cells, args)
        del xyz_relaxed
        # 12.5 M-seed point of the weak-scaling curve (configs[4] at N=1): the SAME iterations 1..K from raw
        # seeds as an N>1 run times per GPU, so that value(N) / (N * this rate) is the weak-scaling efficiency
        n_w = 12500000
        if n_cells != n_w:
            full_w = raw_seeds(n_w)
            job_w = ShardedLloyd(ctx, n_w, None, dev, stream)
            job_w.set_seeds_device(full_w)
            job_w.step()
            job_w.set_seeds_device(full_w)
            torch.cuda.synchronize()
            ev_w = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
            ev_w[0].record(stream)
            for k in range(args.steps):
                job_w.step()
                ev_w[k + 1].record(stream)
            torch.cuda.synchronize()
            ms_w = ev_w[0].elapsed_time(ev_w[-1])
            extra["configs4_weak_n1"] = {"workload": "12.5 M uniform periodic seeds on 1 GPU, Lloyd iterations 1..%d from raw "
                                                     "seeds (the per-GPU load of the N>1 runs)" % args.steps,
                                         "seeds": n_w, "ms_per_step": ms_w / args.steps,
                                         "ms_iter": [ev_w[k].elapsed_time(ev_w[k + 1]) for k in range(args.steps)],
                                         "cells_per_s": n_w * args.steps / (ms_w * 1e-3)}
            del full_w, job_w

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "ms_iter": ms_iter,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: %d uniform seeds, periodic unit cube, CVT Lloyd iterations "
                               "1..%d from the raw seeds (warm-up ran on a scratch copy)" % (n_total, args.steps)
                               if world == 1 else
                               "BASELINE configs[4]: weak scaling, %d uniform periodic seeds per GPU, Lloyd iterations "
                               "1..%d from the raw seeds, x-quantile slab per rank, %s"
                               % (n_per, args.steps,
                                  "halo of ghost seeds exchanged with ncclSend/Recv inside the library, slab-local grid"
                                  if use_slab else "seeds all-gathered over NCCL, every rank bins all seeds"),
                   "seeds_per_gpu": n_per, "seeds_total": n_total,
                   "k_pass": last["k_pass"][: last["passes"]],
                   "cells_per_pass": last["n_pass"][: last["passes"]],
                   "grid": last["grid_dim"], "mode": args.mode,
                   "sharding": (args.sharding if world > 1 else None),
                   "slab": (job.last_stats if use_slab else None),
                   "halo_retries": (job.retries if use_slab else None),
                   "slab_steps_rank0": (job.steps[-(args.steps + e2e_steps):][: args.steps] if use_slab else None),
                   "l2": "inputs (>= 240 MB/step) exceed the 126 MB L2"},
        "phase_ms_per_step": per_step,
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": None if e2e_value is None else
               {"value": e2e_value, "unit": UNIT, "steps": e2e_steps,
                "h2d_bytes_per_step": 24 * n_total, "d2h_bytes_per_step": 24 * n_total,
                "note": "whole job, Lloyd iterations 1..%d from the raw seeds: every rank uploads its own rows from "
                        "pinned host memory and reads them back each step (24 B/seed each way)" % e2e_steps},
        "extra_configs": extra,
        "gpu_launches": int(launches), "clocks": clocks,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
