#!/usr/bin/env python
"""Bench of the tl.infercnv hot path (BASELINE.json metric: cells/s through tl.infercnv, 1M x 20k fp32, 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--cells-total 1000000] [--workloads dense100,dense250,csr100,graph]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...          # CPU arm: the oracle port on the host cores

The job is the metric's own: ONE synthetic matrix of ``--cells-total`` (default 1 000 000) cells x 20 000 genes
(SURVEY.md §8d), row-sharded over the N ranks at multiples of ``chunksize`` -> STRONG scaling (1M cells on one GPU at
N=1: 80 GB of the 180 GB HBM; 125k cells per GPU at N=8).

One "step" = one full pass of the hot path over this rank's row shard: reference profile (column sums, ONE all-reduce
when N > 1, mean) -> centre / clip / pyramid smoothing (the dominant kernel) -> exact row median -> per-chunk noise
threshold -> dense-to-CSR compaction (the reference's ``:455``).  The headline line is window 100 on dense input
(``dense100``); ``sub`` carries the same measurement for BASELINE configs[2] (``dense250``: window 250) and configs[3]
(``csr100``: CSR input) on the same cells, and ``graph`` = configs[4] (a cells x 1792 CNV matrix -> PCA(50) -> kNN(15) ->
Leiden, wall-clock per stage).

``value``   : whole-job cells/s with the input resident in HBM (device-timed, max over ranks).
``e2e``     : the same through the public API ``cnv.tl.infercnv(adata)`` with the matrix in pinned HOST memory: H2D of
              the input and D2H of the CSR result are inside the timed region (on a stated row slice per rank when the
              whole shard would not fit in host memory next to the other ranks').
``roofline``: algorithmic bytes of the smoothing kernel (4*G + 4*K per cell; CSR: 8*nnz + 4 + 4*K)
              divided by its CUDA-event duration, against MEASURED_PEAKS.json.
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
    # name: (window, step, container, BASELINE config it stands for)
    "dense100": (100, 10, "dense", "synthetic cells x 20k genes fp32 dense, window=100 (BASELINE configs[1] shape, the metric's 1M cells)"),
    "dense250": (250, 10, "dense", "synthetic cells x 20k genes fp32 dense, window=250 (BASELINE configs[2])"),
    "csr100": (100, 10, "csr", "synthetic cells x 20k genes CSR input, window=100 (BASELINE configs[3])"),
}
G_GENES = 20000
CHUNK = 5000
DYN = 1.5
LFC = 3.0
METRIC = "cells/sec through tl.infercnv"
E2E_MAX_ROWS = 250_000  # per rank: 20 GB of pinned host memory


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workloads", default="dense100,dense250,csr100,graph",
                    help="first one is the headline line, the others go to `sub` (graph = BASELINE configs[4]: PCA + kNN + Leiden)")
    ap.add_argument("--workload", default=None, help="shorthand for --workloads <one>")
    ap.add_argument("--cells-total", type=int, default=1_000_000, help="cells of the whole job (sharded over the ranks)")
    ap.add_argument("--cells", type=int, default=None, help="cells per GPU (overrides --cells-total: weak scaling)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=150.0, help="wall-clock budget of the reference arm")
    ap.add_argument("--cpu-worker", default=None, help=argparse.SUPPRESS)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port).  Runs in its own interpreter so the fork pool never sees a CUDA context.
def cpu_worker(spec_json: str):
    from infercnvpy_b200.datasets import synthetic_counts, synthetic_var
    from oracle import infercnv_oracle as orc

    spec = json.loads(spec_json)
    window, step, container, _ = WORKLOADS[spec["workload"]]
    workers, steps, warmup, budget = spec["workers"], spec["steps"], spec["warmup"], spec["budget_s"]
    var = synthetic_var(G_GENES, seed=0)

    def make(rows):
        X = synthetic_counts(rows, G_GENES, seed=1000)
        if container == "csr":
            import scipy.sparse as sp

            X = sp.csr_matrix(X)
        return X

    def run(X, chunk):
        t0 = time.perf_counter()
        orc.infercnv(
            X, var["chromosome"].values, var["start"].values, window_size=window, step=step, lfc_clip=LFC,
            dynamic_threshold=DYN, chunksize=chunk, n_jobs=workers,
        )
        return time.perf_counter() - t0

    # calibration (untimed): one chunk of 250 rows per worker -> rows per step that fit the budget over ALL steps
    chunk = 250
    cal = run(make(workers * chunk), chunk)
    rate = workers * chunk / cal
    per_step = max(0.5, (budget - cal) / max(1, steps + warmup))
    chunk = int(min(CHUNK, max(64, rate * per_step / workers)))
    rows = workers * chunk
    X = make(rows)
    times = [run(X, chunk) for _ in range(warmup + steps)]
    print("CPU_WORKER_RESULT " + json.dumps({"seconds": times[warmup:], "rows": rows, "chunk": chunk, "calibration_s": cal}))


def run_cpu_arm(workload: str, steps: int, warmup: int, budget_s: float):
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 32))  # the reference parallelises over chunks only (process_map, :132-139)
    spec = dict(workload=workload, workers=workers, steps=steps, warmup=warmup, budget_s=budget_s)
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--cpu-worker", json.dumps(spec)], capture_output=True, text=True)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("CPU_WORKER_RESULT ")]
    if not line:
        raise RuntimeError(f"cpu worker failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    res = json.loads(line[-1][len("CPU_WORKER_RESULT ") :])
    secs, rows, chunk = res["seconds"], res["rows"], res["chunk"]
    return dict(
        value=rows * len(secs) / sum(secs),
        unit="cells/s",
        cores=workers,
        kind="port",
        sample=f"{len(secs)} timed steps (+{warmup} warm-up) of {rows} cells x {G_GENES} genes each = {workers} chunks of {chunk} rows, "
        f"one per worker process (oracle/infercnv_oracle.py with a {workers}-process pool like the reference's process_map; "
        f"the reference's chunksize only sets the std partition, not the arithmetic per cell); host has {cores} cores",
        seconds_per_step=sum(secs) / len(secs),
        steps_ran=len(secs),
        warmup_ran=warmup,
        rows_per_step=rows,
        chunksize_ran=chunk,
    )


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons sampled in-process through NVML every few ms while the timed region runs
    (B200_PROFILING.md's clocks line; nvidia-smi -lms cannot resolve a region shorter than its poll)."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, gpu_index: int, period_s: float = 0.004):
        self.period = period_s
        self.samples = []  # (t, sm_mhz, reasons bitmask)
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                try:
                    phys = int(vis.split(",")[gpu_index])
                except (ValueError, IndexError):
                    phys = gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.nv = None
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.time(), float(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nv is None or self.thread is None:  # never started (ranks other than 0 do not sample)
            return
        self.stop_flag.set()
        self.thread.join(timeout=1)

    def summary(self, t0, t1):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")], "samples": 0}
        inside = [(m, r) for (t, m, r) in self.samples if t0 <= t <= t1]
        reasons = set()
        for _, r in inside:
            for bit, name in self.REASONS.items():
                if r & bit:
                    reasons.add(name)
        sm = [m for m, _ in inside]
        return {
            "sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
            "samples": len(sm), "how": f"NVML in-process, {self.period * 1e3:.0f} ms period, samples inside the timed regions",
        }


def main():
    args = parse()
    if args.cpu_worker:
        cpu_worker(args.cpu_worker)
        return

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    names = [args.workload] if args.workload else [w for w in args.workloads.split(",") if w]
    for w in names:
        if w not in WORKLOADS and w != "graph":
            raise SystemExit(f"unknown workload {w}")
    head = [w for w in names if w != "graph"][0]
    weak = args.cells is not None
    cells_total = args.cells * world if weak else args.cells_total

    def config_for(name):
        window, step, container, label = WORKLOADS[name]
        return {
            "workload": label, "cells_total": cells_total, "genes": G_GENES, "window": window, "step": step, "container": container,
            "chunksize": CHUNK, "dynamic_threshold": DYN, "lfc_clip": LFC,
            "reference": "mean of all cells (computed every step, all-reduced when N>1)",
            "sharding": f"rows of ONE {cells_total}-cell matrix over {world} rank(s), shard boundaries multiples of chunksize",
            "step_contents": "colsum + all-reduce + mean + set_reference + smooth + row-median centring + chunk threshold + filter count + indptr scan + filtering dense->CSR compaction",
            "l2": f"input {cells_total // world * G_GENES * 4 / 1e9:.0f} GB per GPU per step >> 126 MB L2 (no flush needed)",
        }

    # ---------------- reference arm: the CPU path on the host cores ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        cb = run_cpu_arm(head, args.steps, args.warmup, args.cpu_seconds)
        out = {
            "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "cells/s", "n_gpus": args.gpus,
            "steps": cb["steps_ran"], "warmup": cb["warmup_ran"], "ms_per_step": cb["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_for(head),
            "ran": {"steps": cb["steps_ran"], "warmup": cb["warmup_ran"], "rows_per_step": cb["rows_per_step"], "chunksize": cb["chunksize_ran"],
                    "note": "each step is a bounded sample of the workload (rate is per cell; the CPU path is linear in cells)"},
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
        # NCCL prints its version banner to STDOUT at the first communicator: keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    var = cnv.datasets.synthetic_var(G_GENES, seed=0)
    if weak:
        r0, r1 = rank * args.cells, (rank + 1) * args.cells
    else:
        r0, r1 = cnv.shard_rows(cells_total, CHUNK, rank, world)
    n_local = r1 - r0
    Xd = cnv.datasets.device_counts(n_local, G_GENES, dev, seed=1000 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    traffic_table = {}
    tfile = ROOT / "profiles" / "smooth_traffic.json"
    if tfile.exists():
        try:
            traffic_table = json.load(open(tfile))
        except Exception:
            traffic_table = {}

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    regions = []

    def csr_input():
        """CSR triple of this rank's shard, built blockwise on the device from the dense matrix."""
        ip = torch.zeros((n_local + 1,), dtype=torch.int64, device=dev)
        cols, vals = [], []
        blk = 50_000
        for a in range(0, n_local, blk):
            sub = Xd[a : a + blk]
            mask = sub != 0
            ip[a + 1 : a + 1 + sub.shape[0]] = mask.sum(dim=1)
            nz = mask.nonzero()
            cols.append(nz[:, 1].to(torch.int32))
            vals.append(sub[mask])
            del mask, nz
        torch.cumsum(ip, dim=0, out=ip)
        return ip, torch.cat(cols), torch.cat(vals)

    def measure(name, Xin):
        """Device-timed steps of one workload on this rank's shard -> dict (value, roofline, ...)."""
        window, step, container, _ = WORKLOADS[name]
        layout = build_layout(var, window, step)
        plan = DevicePlan(layout, dev)
        K = plan.K
        out = torch.empty((n_local, K), dtype=torch.float32, device=dev)
        stats = torch.empty((n_local, 2), dtype=torch.float64, device=dev)
        tmp = torch.empty((n_local, plan.tmp_width()), dtype=torch.float64, device=dev)
        n_it = args.steps + args.warmup
        ev_s0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_it)]
        ev_s1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_it)]
        csr_buf = {}
        launches = {"n": 0}

        def one_step(i):
            sums, counts = plan.colsum(Xin)                     # 3 kernels (dense) / 2 (csr)
            sums, counts = allreduce_sums(sums, counts)         # the one collective of the path
            ref = plan.mean_from_sums(sums, counts)             # 1
            plan.set_reference(ref)                             # 1 (+1 when the plan keeps a second table set)
            ev_s0[i].record()
            plan.smooth(Xin, LFC, tmp=tmp)                      # 1  <- dominant kernel (steps 1-3)
            ev_s1[i].record()
            plan.center(tmp, out=out, row_stats=stats)          # 1  (step 4: exact row median)
            # step 5 + :455: chunk thresholds (1), counting pass (1), indptr scan (1), filtering compaction (1)
            if "indices" not in csr_buf:                        # first warm-up step sizes the CSR buffers (+25 %)
                thr, row_abs, row_nnz, (indptr, indices, data) = plan.filter_to_csr(out, stats, CHUNK, DYN)
                cap = int(indices.numel() * 1.25) + 1024
                csr_buf["indices"] = torch.empty((cap,), dtype=torch.int32, device=dev)
                csr_buf["data"] = torch.empty((cap,), dtype=torch.float32, device=dev)
                csr_buf["indptr"] = torch.empty((n_local + 1,), dtype=torch.int64, device=dev)
                del indices, data, indptr
            else:
                thr, row_abs, row_nnz, _ = plan.filter_to_csr(out, stats, CHUNK, DYN, indptr=csr_buf["indptr"],
                                                              indices=csr_buf["indices"], data=csr_buf["data"])
            return row_abs

        for i in range(args.warmup):
            one_step(i)
        barrier()
        plan.launches = 0
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        wall0 = time.time()
        t0.record()
        for i in range(args.warmup, n_it):
            one_step(i)
        t1.record()
        barrier()
        regions.append((wall0, time.time()))
        launches["n"] = plan.launches  # kernels of libicnv.so launched inside the timed region (counted per C-ABI call)
        ms_total = t0.elapsed_time(t1)
        smooth_ms = [ev_s0[i].elapsed_time(ev_s1[i]) for i in range(args.warmup, n_it)]
        nnz_in = float(Xin[2].numel()) if container == "csr" else 0.0
        t = torch.tensor([ms_total, float(np.mean(smooth_ms))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, smooth_avg_ms = float(t[0]), float(t[1])
        ms_per_step = ms_total / args.steps
        if container == "dense":
            bytes_per_cell = 4 * G_GENES + 4 * K
        else:
            bytes_per_cell = 8 * nnz_in / max(1, n_local) + 4 + 4 * K
        achieved = n_local * bytes_per_cell / (smooth_avg_ms * 1e-3) / 1e9
        traffic = traffic_table.get(name)
        if traffic is not None:  # ncu --set full on a 100 000-cell launch; the kernel streams, DRAM bytes scale with rows
            traffic = float(traffic) * n_local / 100_000
        res = {
            "value": cells_total / (ms_per_step * 1e-3), "unit": "cells/s", "ms_per_step": ms_per_step, "gpu_launches": launches["n"],
            "config": config_for(name), "K": K,
            "roofline": {
                "kernel": "icnv::smooth_kernel", "input": container, "bound": "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_cell": bytes_per_cell, "cells_per_launch": n_local, "ms_per_launch": smooth_avg_ms,
                "launch_info": plan.launch_info(),
            },
            "whole_step_frac_of_peak": n_local * bytes_per_cell / (ms_per_step * 1e-3) / 1e9 / peak,
        }
        plan.close()
        del out, stats, tmp, csr_buf
        torch.cuda.empty_cache()
        return res

    def measure_graph():
        """BASELINE configs[4]: CNV matrix -> PCA(50) -> kNN(15) + fuzzy graph -> Leiden, cells sharded over the ranks
        (K x K Gram all-reduce, all-gather of the N x 50 coordinates and of the kNN lists; clustering replicated)."""
        from infercnvpy_b200.pp._neighbors import neighbors_device
        from infercnvpy_b200.tl._leiden import leiden_device
        from infercnvpy_b200.tl._pca import pca_device

        K, n_clu = 1792, 12
        gen = torch.Generator(device=dev)
        gen.manual_seed(7)  # same cluster profiles on every rank
        centers = (torch.rand((n_clu, K), generator=gen, device=dev) < 0.08).float() * torch.randn((n_clu, K), generator=gen, device=dev) * 0.15
        gen.manual_seed(100 + rank)
        lab = torch.randint(0, n_clu, (n_local,), generator=gen, device=dev)
        X = torch.empty((n_local, K), dtype=torch.float32, device=dev)
        for a in range(0, n_local, 100_000):
            b = min(n_local, a + 100_000)
            noise = torch.randn((b - a, K), generator=gen, device=dev) * 0.05
            noise *= (torch.rand((b - a, K), generator=gen, device=dev) < 0.15)
            X[a:b] = centers[lab[a:b]] + noise
        steps_g, warm_g = max(1, min(args.steps, 2)), 1
        stage = {"pca": [], "neighbors": [], "leiden": []}
        total = []
        n_clusters = 0
        for it in range(warm_g + steps_g):
            barrier()
            t0 = time.perf_counter()
            Y, _, _ = pca_device(X, 50)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            g = neighbors_device(Y, 15)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            r, c, w = g["coo"]
            n_tot = g["n_total"]
            order = torch.argsort(r * n_tot + c)
            indptr = torch.zeros(n_tot + 1, dtype=torch.int64, device=dev)
            indptr[1:] = torch.cumsum(torch.bincount(r, minlength=n_tot), 0)
            labels = leiden_device(indptr, c[order].to(torch.int32), w[order])
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            if it >= warm_g:
                stage["pca"].append(t1 - t0)
                stage["neighbors"].append(t2 - t1)
                stage["leiden"].append(t3 - t2)
                total.append(t3 - t0)
            n_clusters = int(labels.max().item()) + 1
        tt = torch.tensor([float(np.mean(total))] + [float(np.mean(stage[k])) for k in ("pca", "neighbors", "leiden")], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt[0])
        # purity of the recovered clusters on this rank's rows (sanity: the planted structure is found)
        row0 = g["row0"]
        mine = labels[row0 : row0 + n_local]
        joint = torch.zeros((n_clusters, n_clu), dtype=torch.int64, device=dev)
        joint.index_put_((mine, lab), torch.ones_like(lab), accumulate=True)
        purity = float(joint.max(dim=1).values.sum()) / max(1, n_local)
        flops_exec = 2.0 * n_local * cells_total * 56 * 3  # 3xTF32 passes over ceil(50/8)*8 = 56 dims, this rank's queries
        return {
            "value": cells_total / sec, "unit": "cells/s", "seconds_per_step": sec, "steps": steps_g, "warmup": warm_g,
            "stage_seconds": {"pca": float(tt[1]), "neighbors": float(tt[2]), "leiden": float(tt[3])},
            "clusters": n_clusters, "purity": purity,
            "config": {"workload": "synthetic CNV matrix (cells x 1792 windows, 12 planted clones) -> PCA(50) + kNN(15) + Leiden (BASELINE configs[4])",
                       "cells_total": cells_total, "windows": K, "n_comps": 50, "n_neighbors": 15, "sharding": f"rows over {world} rank(s)"},
            "knn_executed_tflop_per_rank": flops_exec / 1e12,
        }

    results = {}
    want_graph = "graph" in names
    names = [w for w in names if w != "graph"]
    head = names[0]
    dense_names = [w for w in names if WORKLOADS[w][2] == "dense"]
    csr_names = [w for w in names if WORKLOADS[w][2] == "csr"]
    for w in dense_names:
        results[w] = measure(w, Xd)

    # ---- end to end through the public API with host buffers (headline workload): stage the host copy now
    e2e = None
    Xhost = None
    window, step, container, _ = WORKLOADS[head]
    n_e2e = min(n_local, E2E_MAX_ROWS) // CHUNK * CHUNK or n_local

    def pinned_np(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        torch.cuda.synchronize()
        return h.numpy()

    if not args.no_e2e and container == "dense":
        Xhost = pinned_np(Xd[:n_e2e])
    Xcsr = csr_input() if (csr_names or (container == "csr" and not args.no_e2e)) else None
    del Xd
    torch.cuda.empty_cache()
    if not args.no_e2e and container == "csr":
        import scipy.sparse as sp

        ip, ix, dv = Xcsr
        e1 = int(ip[n_e2e].item())
        Xhost = sp.csr_matrix((pinned_np(dv[:e1]), pinned_np(ix[:e1]), pinned_np(ip[: n_e2e + 1])), shape=(n_e2e, G_GENES), copy=False)
        Xhost.has_canonical_format = True  # built row by row from the dense matrix: sorted, no duplicates
    for w in csr_names:
        results[w] = measure(w, Xcsr)
    del Xcsr
    torch.cuda.empty_cache()

    if not args.no_e2e:
        adata = cnv.AnnData(Xhost, var=var)
        e2e_steps = max(2, min(args.steps, 3))
        kw = dict(window_size=window, step=step, lfc_clip=LFC, dynamic_threshold=DYN, chunksize=CHUNK, inplace=False)
        chr_pos, res, _ = cnv.tl.infercnv(adata, **kw)  # warm-up (plan tables, pinned result buffers)
        barrier()
        wall0 = time.time()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            chr_pos, res, _ = cnv.tl.infercnv(adata, **kw)
        torch.cuda.synchronize()
        w = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
        regions.append((wall0, time.time()))
        if world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        sec = float(w[0]) / e2e_steps
        from infercnvpy_b200.tl._infercnv import LAST_TRANSFER

        h2d, d2h = int(LAST_TRANSFER["h2d_bytes"]), int(LAST_TRANSFER["d2h_bytes"])  # counted where the copies are issued
        # the link's own ceiling: the same input bytes as ONE plain pinned->device copy, all ranks at the same time
        link_sec = None
        if container == "dense":
            src = torch.from_numpy(Xhost)
            dst = torch.empty(src.shape, dtype=src.dtype, device=dev)
            dst.copy_(src, non_blocking=True)
            barrier()
            l0 = time.perf_counter()
            for _ in range(2):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            lt = torch.tensor([(time.perf_counter() - l0) / 2], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(lt, op=dist.ReduceOp.MAX)
            link_sec = float(lt[0])
            del dst
        rows_all = torch.tensor([n_e2e], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(rows_all)
        e2e = {
            "value": int(rows_all[0]) / sec, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "seconds_per_step": sec, "cells_per_step_all_ranks": int(rows_all[0]), "cells_per_step_per_rank": n_e2e, "steps": e2e_steps,
            "h2d_only_seconds": link_sec, "h2d_only_GBps_per_rank": (h2d / link_sec / 1e9 if link_sec else None),
            "h2d_only_note": "plain pinned->device copy of the same input, all ranks concurrently (max over ranks): the share of seconds_per_step the host link alone takes",
            "api": "infercnvpy_b200.tl.infercnv(adata) with adata.X in pinned host memory; result scipy CSR float64 on host; "
                   f"every rank runs its {'whole shard' if n_e2e == n_local else f'first {n_e2e} rows of its shard (host memory bound)'}",
        }

    graph_res = None
    if want_graph:
        if world > 1:
            graph_res = measure_graph()  # collectives inside: an exception on one rank must not leave the others waiting
        else:
            try:
                graph_res = measure_graph()
            except Exception as e:  # single process: the sub-line must not take the headline measurement down with it
                graph_res = {"error": repr(e)[:400]}
    sampler.stop()
    clocks = None
    if rank == 0:
        t_lo, t_hi = min(a for a, _ in regions), max(b for _, b in regions)
        clocks = sampler.summary(t_lo, t_hi)
        inside = [(t, m, r) for (t, m, r) in sampler.samples if any(a <= t <= b for a, b in regions)]
        if inside:
            clocks["sm_mhz"] = float(np.median([m for _, m, _ in inside]))
            clocks["samples"] = len(inside)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = run_cpu_arm(head, steps=2, warmup=0, budget_s=25.0)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        h = results[head]
        line = {
            "metric": METRIC, "value": h["value"], "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": h["ms_per_step"], "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": h["config"], "clocks": clocks, "e2e": e2e, "gpu_launches": h["gpu_launches"],
            "roofline": h["roofline"], "cpu_baseline": cpu_baseline, "impl": "b200",
            "whole_step_frac_of_peak": h["whole_step_frac_of_peak"],
            "sub": {**{k: v for k, v in results.items() if k != head}, **({"graph": graph_res} if graph_res else {})},
            "dtype_note": "fp32 input/centring/output, fp64 window accumulation (the reference computes the convolution in float64)",
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
