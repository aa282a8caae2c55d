#!/usr/bin/env python
"""Bench of the tl.infercnv hot path (BASELINE.json metric: cells/s through tl.infercnv).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dense100|dense250|csr100]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...          # CPU arm: the oracle port on the host cores

One "step" = one full pass of the hot path over this rank's row shard of the synthetic matrix
(SURVEY.md §8d): reference profile (column means, one all-reduce when N > 1) -> centre / clip /
pyramid smoothing / row-median (the dominant kernel) -> per-chunk noise threshold.  Weak scaling:
every rank holds ``--cells`` rows (default 100 000 x 20 000 fp32 = configs[1] of BASELINE.json).

``value``   : whole-job cells/s with the input resident in HBM (device-timed, max over ranks).
``e2e``     : the same through the public API ``cnv.tl.infercnv(adata)`` with the matrix in pinned
              HOST memory: H2D of the input and D2H of the CSR result are inside the timed region.
``roofline``: algorithmic bytes of the smoothing kernel (4*G + 4*K per cell) divided
              by its CUDA-event duration, against MEASURED_PEAKS.json.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (window, step, container)
    "dense100": (100, 10, "dense"),
    "dense250": (250, 10, "dense"),
    "csr100": (100, 10, "csr"),
}
G_GENES = 20000
CHUNK = 5000
DYN = 1.5
LFC = 3.0
METRIC = "cells/sec through tl.infercnv"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dense100", choices=list(WORKLOADS))
    ap.add_argument("--cells", type=int, default=100_000, help="cells per GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-worker", default=None, help=argparse.SUPPRESS)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port).  Runs in its own interpreter so the fork pool never sees a CUDA context.
def cpu_worker(spec_json: str):
    from infercnvpy_b200.datasets import synthetic_counts, synthetic_var
    from oracle import infercnv_oracle as orc

    spec = json.loads(spec_json)
    window, step, container = WORKLOADS[spec["workload"]]
    rows, chunk, workers = spec["rows"], spec["chunk"], spec["workers"]
    var = synthetic_var(G_GENES, seed=0)
    X = synthetic_counts(rows, G_GENES, seed=1000)
    if container == "csr":
        import scipy.sparse as sp

        X = sp.csr_matrix(X)
    times = []
    for _ in range(spec["warmup"] + spec["steps"]):
        t0 = time.perf_counter()
        orc.infercnv(
            X, var["chromosome"].values, var["start"].values, window_size=window, step=step, lfc_clip=LFC,
            dynamic_threshold=DYN, chunksize=chunk, n_jobs=workers,
        )
        times.append(time.perf_counter() - t0)
    timed = times[spec["warmup"] :]
    print("CPU_WORKER_RESULT " + json.dumps({"seconds": timed, "rows": rows}))


def run_cpu_arm(workload: str, steps: int, warmup: int):
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 20))  # the reference parallelises over chunks only: 100k rows = 20 chunks
    chunk = 1250
    rows = workers * chunk
    spec = dict(workload=workload, rows=rows, chunk=chunk, workers=workers, steps=steps, warmup=warmup)
    r = subprocess.run(
        [sys.executable, str(ROOT / "bench.py"), "--cpu-worker", json.dumps(spec)], capture_output=True, text=True
    )
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("CPU_WORKER_RESULT ")]
    if not line:
        raise RuntimeError(f"cpu worker failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    res = json.loads(line[-1][len("CPU_WORKER_RESULT ") :])
    secs = res["seconds"]
    return dict(
        value=rows * len(secs) / sum(secs),
        unit="cells/s",
        cores=workers,
        kind="port",
        sample=f"{rows} cells x {G_GENES} genes ({workers} chunks of {chunk}), oracle/infercnv_oracle.py "
        f"with a {workers}-process pool like the reference's process_map; host has {cores} cores",
        seconds_per_step=sum(secs) / len(secs),
    )


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", float("inf"))
        inside = [ln for (t, ln) in self.lines if t0 <= t <= t1 + 0.15]
        for ln in inside or [ln for (_, ln) in self.lines]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * max(mx, default=1)] or sm
        return {
            "sm_mhz": float(np.median(busy)) if busy else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


def main():
    args = parse()
    if args.cpu_worker:
        cpu_worker(args.cpu_worker)
        return

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    window, step, container = WORKLOADS[args.workload]
    workload_name = {
        "dense100": "synthetic 100k cells x 20k genes fp32 dense, window=100 (BASELINE configs[1])",
        "dense250": "synthetic cells x 20k genes fp32 dense, window=250 (BASELINE configs[2] shape per GPU)",
        "csr100": "synthetic cells x 20k genes CSR input densify-on-load, window=100 (BASELINE configs[3] shape per GPU)",
    }[args.workload]
    config = {
        "workload": workload_name,
        "cells_per_gpu": args.cells,
        "genes": G_GENES,
        "window": window,
        "step": step,
        "chunksize": CHUNK,
        "dynamic_threshold": DYN,
        "reference": "mean of all cells (computed every step, all-reduced when N>1)",
        "sharding": f"rows, {world} rank(s), shard boundaries multiples of chunksize",
        "l2": f"input {args.cells * G_GENES * 4 / 1e9:.0f} GB per GPU per step >> 126 MB L2 (no flush needed)",
    }

    # ---------------- reference arm: the CPU path on the host cores ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        cb = run_cpu_arm(args.workload, min(args.steps, 10), min(args.warmup, 1))
        out = {
            "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "cells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(out))
        return

    # ---------------- B200 arm ----------------
    import torch
    import torch.distributed as dist

    import infercnvpy_b200 as cnv
    from infercnvpy_b200._engine import DevicePlan, allreduce_sums
    from infercnvpy_b200._layout import build_layout

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    var = cnv.datasets.synthetic_var(G_GENES, seed=0)
    layout = build_layout(var, window, step)
    n_local = args.cells
    Xd = cnv.datasets.device_counts(n_local, G_GENES, dev, seed=1000 + rank)
    if container == "csr":
        csr = Xd.to_sparse_csr()
        Xin = (csr.crow_indices().to(torch.int64).contiguous(), csr.col_indices().to(torch.int32).contiguous(), csr.values().contiguous())
        del Xd, csr
        torch.cuda.empty_cache()
    else:
        Xin = Xd

    plan = DevicePlan(layout, dev)
    K = plan.K
    out = torch.empty((n_local, K), dtype=torch.float32, device=dev)
    stats = torch.empty((n_local, 2), dtype=torch.float64, device=dev)
    tmp_holder = {}
    launches = {"n": 0}

    ev_s0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + args.warmup)]
    ev_s1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + args.warmup)]

    def one_step(i):
        sums, counts = plan.colsum(Xin)                     # 3 kernels (dense) / 2 (csr)
        sums, counts = allreduce_sums(sums, counts)         # the one collective of the path
        ref = plan.mean_from_sums(sums, counts)             # 1
        plan.set_reference(ref)                             # 1
        if "tmp" not in tmp_holder:
            tmp_holder["tmp"] = torch.empty((n_local, plan.tmp_width()), dtype=torch.float64, device=dev)
        ev_s0[i].record()
        plan.smooth(Xin, LFC, tmp=tmp_holder["tmp"])        # 1  <- dominant kernel (steps 1-3)
        ev_s1[i].record()
        plan.center(tmp_holder["tmp"], out=out, row_stats=stats)             # 1  (step 4: exact row median)
        thr, row_abs, row_nnz = plan.threshold(out, stats, CHUNK, DYN)       # 2  (step 5)
        launches["n"] += 9 if container == "dense" else 8
        return row_abs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        one_step(i)
    barrier()
    launches["n"] = 0
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    t0.record()
    for i in range(args.warmup, args.warmup + args.steps):
        one_step(i)
    t1.record()
    barrier()
    sampler.window(wall0, time.time())
    clocks = sampler.stop() if rank == 0 else None
    ms_total = t0.elapsed_time(t1)
    smooth_ms = [ev_s0[i].elapsed_time(ev_s1[i]) for i in range(args.warmup, args.warmup + args.steps)]
    t = torch.tensor([ms_total, float(np.mean(smooth_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, smooth_avg_ms = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    value = world * n_local / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (per launch, this rank; worst rank's duration)
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    if container == "dense":
        bytes_per_cell = 4 * G_GENES + 4 * K
    else:
        nnz_row = float(Xin[2].numel()) / n_local
        bytes_per_cell = 8 * nnz_row + 4 + 4 * K
    achieved = n_local * bytes_per_cell / (smooth_avg_ms * 1e-3) / 1e9
    traffic = None
    tfile = ROOT / "profiles" / "smooth_traffic.json"
    if tfile.exists():
        try:
            # measured on a launch of 100 000 cells (ncu --set full, tools/one_step.py); the kernel streams, so DRAM
            # bytes scale with the rows of the launch
            traffic = json.load(open(tfile)).get(args.workload)
            if traffic is not None:
                traffic = float(traffic) * n_local / 100_000
        except Exception:
            traffic = None
    roofline = {
        "kernel": "icnv::smooth_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "bytes_per_cell": bytes_per_cell, "cells_per_launch": n_local, "ms_per_launch": smooth_avg_ms,
        "launch_info": plan.launch_info(),
    }

    # ---- end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        n_e2e = n_local
        if container == "csr":
            import scipy.sparse as sp

            def pinned_np(t):  # like the dense arm: the host container lives in pinned memory
                h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                h.copy_(t)
                return h.numpy()

            ip, ix, dv = (pinned_np(x) for x in Xin)
            Xhost = sp.csr_matrix((dv, ix, ip), shape=(n_local, G_GENES), copy=False)
            Xhost.has_canonical_format = True  # produced by torch's to_sparse_csr: sorted, no duplicates
        else:
            host = torch.empty((n_e2e, G_GENES), dtype=torch.float32, pin_memory=True)
            host.copy_(Xd)
            torch.cuda.synchronize()
            Xhost = host.numpy()
        adata = cnv.AnnData(Xhost, var=var)
        e2e_steps = max(2, min(args.steps, 5))
        res = None
        for _ in range(1):
            chr_pos, res, _ = cnv.tl.infercnv(adata, window_size=window, step=step, lfc_clip=LFC, dynamic_threshold=DYN, chunksize=CHUNK, inplace=False)
        barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            chr_pos, res, _ = cnv.tl.infercnv(adata, window_size=window, step=step, lfc_clip=LFC, dynamic_threshold=DYN, chunksize=CHUNK, inplace=False)
        torch.cuda.synchronize()
        w = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        sec = float(w[0]) / e2e_steps
        h2d = n_e2e * G_GENES * 4 if container == "dense" else int(Xhost.data.nbytes + Xhost.indices.nbytes + Xhost.indptr.nbytes * 2)
        d2h = int(res.nnz * 8 + (res.shape[0] + 1) * 8)
        e2e = {
            "value": world * n_e2e / sec, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "seconds_per_step": sec, "api": "infercnvpy_b200.tl.infercnv(adata) with adata.X in pinned host memory; result scipy CSR on host",
        }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = run_cpu_arm(args.workload, steps=1, warmup=0)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": launches["n"],
            "roofline": roofline, "cpu_baseline": cpu_baseline, "impl": "b200",
            "dtype_note": "fp32 input/centring/output, fp64 window accumulation (reference computes the convolution in float64)",
        }
        print(json.dumps(line))
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
