#!/usr/bin/env python
"""bench.py — Lloyd-iteration samples/sec of the B200 KMeans engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): synthetic blobs 10M x 64 float32, k = 256, one chunk per
GPU, fixed init = first k rows.  A *step* is one full Lloyd iteration over the resident chunk:
fused E+M kernel -> (N>1: one all-reduce of [k*d sums | k counts | inertia]) -> centre update + shift.
With N>1 every rank holds its own 10M-row chunk (weak scaling, no data-path collective besides the
per-iteration all-reduce).  Rank 0 prints ONE JSON line.

Numbers reported:
  value        whole-job samples/s with X resident in HBM (CUDA events, max over ranks)
  e2e          same metric through ``lloyd_iteration_host`` with X in pinned HOST memory: every step
               copies X host->device (double-buffered row blocks) and reads the new centres back
  roofline     the fused chunk kernel against the tensor (dense 16-bit) and HBM roofs, algorithmic work
               2*d*k flops and d*4+4 bytes per sample
  cpu_baseline the dask-ml path restated without dask (oracle/: scikit-learn E-step + C scatter-add,
               thread pool over os.cpu_count() row blocks) on a bounded row sample of the same workload
``--impl reference`` times only that CPU path and prints the same line shape.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# the CPU baseline runs one BLAS thread per row-block task (set before numpy/scipy load their BLAS)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

import numpy as np  # noqa: E402

N_ROWS = 10_000_000
N_FEAT = 64
N_CLUST = 256
METRIC = "kmeans_lloyd_iter_samples_per_sec"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            m = json.load(f)
        src = "measured"
    else:
        m = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
        src = "fallback"
    t = os.path.join(ROOT, "profiles", "tf32_peak.json")
    tf32 = None
    if os.path.exists(t):
        with open(t) as f:
            tf32 = json.load(f)
    return m, src, tf32


def synth_blobs_device(n, d, k_true, seed, device, dtype):
    """k_true isotropic blobs, centres ~U(-10,10)^d, sigma=1, generated on the device in row blocks."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cent = (torch.rand((k_true, d), generator=g, device=device, dtype=torch.float32) * 20.0 - 10.0)
    X = torch.empty((n, d), device=device, dtype=dtype)
    blk = 1 << 20
    for s in range(0, n, blk):
        m = min(blk, n - s)
        idx = torch.randint(0, k_true, (m,), generator=g, device=device)
        X[s:s + m] = (cent[idx] + torch.randn((m, d), generator=g, device=device, dtype=torch.float32)).to(dtype)
    return X


def synth_blobs_host(n, d, k_true, seed):
    rng = np.random.default_rng(seed)
    cent = rng.uniform(-10, 10, size=(k_true, d)).astype(np.float32)
    idx = rng.integers(0, k_true, size=n)
    return cent[idx] + rng.standard_normal((n, d), dtype=np.float32)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                pynvml.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            while not self._stop_evt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.05)
        except Exception as e:  # pragma: no cover
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------
# CPU baseline (oracle) — the only place bench.py touches oracle/
# ------------------------------------------------------------------------------------------
def cpu_lloyd_baseline(sample_rows, iters, warm):
    from oracle import kmeans_oracle as ok
    import subprocess

    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle_c.so")):
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    # OpenBLAS in this image is built for at most 128 threads and aborts beyond that; 64 worker threads
    # (one BLAS thread each) keep a safe margin on the many-core GPU hosts.  `cores` reports what was used.
    cores = min(os.cpu_count() or 1, 64)
    X = synth_blobs_host(sample_rows, N_FEAT, N_CLUST, 0)
    init = X[:N_CLUST].copy()
    blocks = ok.to_blocks(X, max(1, sample_rows // cores))
    pool = ok.make_pool(cores)
    limiter = None                                 # BLAS/OpenMP are pinned to 1 thread per task via the env above
    centers = init
    times = []
    for i in range(warm + iters):
        t0 = time.perf_counter()
        _, _, centers = ok.lloyd_iteration(blocks, centers, N_CLUST, pool)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    pool.shutdown()
    t = float(np.median(times))
    return {"value": sample_rows / t, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "%d of %d rows of the same workload, %d Lloyd iterations (median), %d row blocks on %d threads, "
                      "scikit-learn float64 E-step + C scatter-add" % (sample_rows, N_ROWS, iters, len(blocks), cores),
            "ms_per_iter": t * 1e3}


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample = 1_000_000
    res = cpu_lloyd_baseline(sample, max(1, args.steps), max(1, min(args.warmup, 2)))
    line = {
        "metric": METRIC, "value": res["value"], "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_iter"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": "C2 blobs 10M x 64 float32, k=256 (bounded sample of %d rows on host cores)" % sample,
                   "n_features": N_FEAT, "n_clusters": N_CLUST},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS, help="rows per GPU (default: the named workload)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the hot path is sm_100a CUDA; there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from dask_ml_b200 import _lib
    from dask_ml_b200.cluster.k_means import LloydState, lloyd_iteration_host
    from dask_ml_b200.engine import Comm, CudaBackend, DeviceData

    be = CudaBackend(dev)
    comm = Comm()
    n = args.rows
    X = synth_blobs_device(n, N_FEAT, N_CLUST, 1000 + rank, dev, torch.float32)
    data = DeviceData([X], be, comm)
    init = X[:N_CLUST].cpu().numpy().astype(np.float64)
    if world > 1:
        init = comm.bcast_obj(init)
    st = LloydState(data, init)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(args.warmup):
        st.step()
        st.accept()
    barrier()

    # ---- timed region: K Lloyd iterations, device-resident X (2.56 GB per GPU >> 126 MB L2) ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = be.launch_count()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev0.record()
    for i in range(args.steps):
        st.step(kernel_events=kev[i])
        st.accept()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = be.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([ms_total, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kern_ms = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    value = n * world / (ms_per_step * 1e-3)
    shift = float(st.shift.item())

    # ---- e2e: host-resident X, H2D inside the timed region, through the public host-buffer call ----
    e2e = None
    if not args.no_e2e:
        e_rows = n
        Xh = torch.empty((e_rows, N_FEAT), dtype=torch.float32, pin_memory=True)
        blk = 1 << 20
        for s in range(0, e_rows, blk):
            Xh[s:s + blk].copy_(X[s:s + blk])
        torch.cuda.synchronize()
        centers = init.copy()
        e_steps = max(2, min(args.steps, 4))
        lloyd_iteration_host(Xh, centers, backend=be, comm=comm)      # warm-up (allocations, pinned pages)
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(e_steps):
            centers, inertia_h, shift_h = lloyd_iteration_host(Xh, centers, backend=be, comm=comm)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ems = max(e0.elapsed_time(e1), wall * 1e3) / e_steps
        te = torch.tensor([ems], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ems = float(te[0])
        e2e = {"value": e_rows * world / (ems * 1e-3), "unit": "samples/s",
               "h2d_bytes_per_step": int(e_rows * N_FEAT * 4 + N_CLUST * N_FEAT * 8),
               "d2h_bytes_per_step": int(N_CLUST * N_FEAT * 8 + 16), "ms_per_step": ems, "steps": e_steps,
               "api": "dask_ml_b200.cluster.k_means.lloyd_iteration_host (pinned host X, double-buffered H2D)"}
        del Xh

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src, tf32 = _peaks()
    flops = 2.0 * N_FEAT * N_CLUST * n
    bytes_alg = (N_FEAT * 4 + 4) * n
    ach_tf = flops / (kern_ms * 1e-3) / 1e12
    ach_gbs = bytes_alg / (kern_ms * 1e-3) / 1e9
    # The product runs on the kind::f16 tensor pipe (split-fp16, 3 products + the ||c||^2 step = 3.25x the
    # algorithmic flops), so the roof is the measured dense 16-bit peak: the sustained figure, because the
    # kernel is timed inside a long step.
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    peak_note = "%s dense bf16/fp16 tensor peak, sustained (MEASURED_PEAKS.json)" % peak_src
    issued_ratio = 3.25
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roofline = {
        "bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
        "traffic": traffic, "peak_source": peak_note, "kernel_ms": kern_ms,
        "kernel": "tc_chunk_kernel<true,false> (tcgen05 split-fp16 fused E+M) + tc_recheck + reduce_partials",
        "issued": {"tflops": ach_tf * issued_ratio, "frac": ach_tf * issued_ratio / peak_tf,
                   "note": "tensor-pipe work actually issued: 3 fp16 products + ||c||^2 step per algorithmic product"},
        "hbm": {"achieved": ach_gbs, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                "frac": ach_gbs / float(peaks["hbm_gbs"]), "peak_source": peak_src},
        "algorithmic": {"flops_per_sample": 2 * N_FEAT * N_CLUST, "bytes_per_sample": N_FEAT * 4 + 4,
                        "samples_per_launch": n},
    }
    cpu = None
    if not args.no_cpu:
        cpu = cpu_lloyd_baseline(500_000, 3, 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: synthetic blobs %d x %d float32 per GPU, k=%d, one chunk per GPU, fixed init (first k rows)"
                               % (n, N_FEAT, N_CLUST),
                   "n_samples_per_gpu": n, "n_features": N_FEAT, "n_clusters": N_CLUST,
                   "arithmetic": "split-fp16 (hi,lo) x3 product on tcgen05 kind::f16, fp32 accumulate, float64 re-check of near-ties + float64 centre update",
                   "l2": "inputs (%.2f GB per GPU) are larger than L2 (126 MB); no explicit flush" % (n * N_FEAT * 4 / 1e9),
                   "kernel_family": int(be.kernel_family(N_FEAT, N_CLUST, torch.float32)),
                   "final_shift": shift},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
