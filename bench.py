#!/usr/bin/env python
"""bench.py — the measurement contract for the wgebra hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extras]

Workload (BASELINE.json): at N = 1, `configs[2]` — bf16 GEMM 4096^3, the configuration the
headline target ("≥80 % of bf16 tensor-core peak on 4096³ at 1 GPU") is quoted on.  At N > 1 the
cube grows with the box: (4096·N)^3 bf16, row-sharded over N GPUs with the all-gather of C over
NVLink — N = 8 is exactly `configs[4]` (32768^3).  One "step" = one pass of the path: one GEMM
(N = 1) or one row-sharded GEMM + all-gather (N > 1).  `value` = 2·M·N·K·steps / time, whole job.

Timing: W untimed warm-up steps, then exactly K steps between CUDA events on the launching stream,
a barrier + device synchronise on both sides, max over ranks.  The three operand sets (A, B, C)
rotate through NSETS copies whose total footprint exceeds the 126 MB L2, so no step finds its inputs
cached from the previous one.  Clocks and throttle reasons are sampled through NVML during every
timed region.

The JSON line also carries: `e2e` (same metric through the C ABI with HOST buffers: H2D of A and B
and D2H of C inside the timed region), `roofline` (dominant kernel vs MEASURED_PEAKS.json),
`cpu_baseline` (the oracle's restatement of gemm.wgsl on the host cores, bounded sample), and at
N = 1 `extra`: the other BASELINE configs (f32 GEMM sweep, GEMV / level-1 GB/s) measured the same way.

`--impl reference`: the reference's own implementation cannot be built here (Rust + wgpu; see
DESIGN.md), so this arm times the oracle port of the reference's WGSL algorithm on all host cores,
on a bounded row-strip sample of the same workload, and prints the same JSON shape.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_BYTES = 126 * 1000 * 1000
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {k: float(d[k]) for k in FALLBACK_PEAKS if k in d} | {"source": "measured"}
        except Exception:
            pass
    return dict(FALLBACK_PEAKS) | {"source": "fallback"}


# ------------------------------------------------------------------------------- NVML clocks
class ClockSampler:
    """Samples SM clock + throttle reasons during timed regions (B200_PROFILING.md clocks line, via NVML)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._active = threading.Event()
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        except Exception as e:  # NVML missing: report it, do not guess
            self.err = repr(e)

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                time.sleep(0.002)
            else:
                time.sleep(0.0005)

    def __enter__(self):
        self._active.set()
        return self

    def __exit__(self, *a):
        self._active.clear()

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------- CPU baseline (oracle)
def _cpu_operand(O, seed, rows, cols):
    """Seeded U[0,1) operand; beyond 4096 x 4096 the seeded 4096-square block is tiled (values do not affect CPU timing and
    generating 10^9 hashed elements in numpy would take longer than the measurement)."""
    if rows <= 4096 and cols <= 4096:
        return O.uniform(seed, rows, cols)
    blk = O.uniform(seed, min(rows, 4096), min(cols, 4096)).reshape(min(cols, 4096), min(rows, 4096))
    reps = (-(-cols // blk.shape[0]), -(-rows // blk.shape[1]))
    return np.ascontiguousarray(np.tile(blk, reps)[:cols, :rows]).reshape(-1)


def cpu_gemm_sample(target_s: float, n: int = 4096, repeat: int = 1):
    """Times the oracle's restatement of gemm.wgsl:81-113 on a row strip of the n^3 product (f32: the reference
    has no bf16).  Returns (tflops, cores, description, seconds_per_run, rows)."""
    from oracle import oracle as O
    cores = O.use_all_cores()
    b = _cpu_operand(O, O.SEED_BASE + 2, n, n)
    # calibrate on 64 rows
    rows = 64
    a = _cpu_operand(O, O.SEED_BASE + 1, rows, n)
    out = np.zeros(rows * n, np.float32)
    t0 = time.perf_counter()
    O.gemm(O.GEMM, out, O.shape(rows, n), a, O.shape(rows, n), b, O.shape(n, n))
    dt = time.perf_counter() - t0
    rate = 2.0 * rows * n * n / dt
    rows = int(min(n, max(64, (target_s * rate / (2.0 * n * n)) // 64 * 64)))
    a = _cpu_operand(O, O.SEED_BASE + 1, rows, n)
    out = np.zeros(rows * n, np.float32)
    times = []
    for _ in range(repeat):
        t0 = time.perf_counter()
        O.gemm(O.GEMM, out, O.shape(rows, n), a, O.shape(rows, n), b, O.shape(n, n))
        times.append(time.perf_counter() - t0)
    dt = float(np.median(times))
    desc = (f"rows 0..{rows} of the {n}^3 product ({rows}x{n}x{n}), f32, oracle port of gemm.wgsl:81-113 "
            f"(wgpu fallback adapter unavailable: no Rust/Vulkan in the image), OpenMP over invocations")
    return 2.0 * rows * n * n / dt / 1e12, cores, desc, dt, rows


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = 4096 * max(1, args.gpus)          # the same cube as the native arm's config at this N
    from oracle import oracle as O
    # size the per-step sample so that (steps + warmup) steps finish in ~2 minutes at most
    budget = 100.0 / max(1, args.steps + args.warmup)
    tf, cores, desc, dt, rows = cpu_gemm_sample(min(2.0, budget), n, repeat=1)
    b = _cpu_operand(O, O.SEED_BASE + 2, n, n)
    a = _cpu_operand(O, O.SEED_BASE + 1, rows, n)
    out = np.zeros(rows * n, np.float32)
    for _ in range(args.warmup):
        O.gemm(O.GEMM, out, O.shape(rows, n), a, O.shape(rows, n), b, O.shape(n, n))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.gemm(O.GEMM, out, O.shape(rows, n), a, O.shape(rows, n), b, O.shape(n, n))
    dt = (time.perf_counter() - t0) / args.steps
    val = 2.0 * rows * n * n / dt / 1e12
    line = {"impl": "reference", "metric": "gemm_tflops", "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus) | {"cpu_sample": desc},
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(ngpu: int):
    n = 4096 * ngpu
    if ngpu == 1:
        wl = "bf16 GEMM 4096x4096x4096 on 1xB200 (BASELINE configs[2]), f32 accumulate, bf16 out"
    else:
        wl = (f"bf16 GEMM {n}x{n}x{n} row-sharded across {ngpu}xB200 with all-gather of C over NVLink "
              f"(BASELINE configs[4] family: (4096*P)^3; P=8 is 32768^3)")
    return {"workload": wl, "M": n, "N": n, "K": n, "parallelism": f"row-shard x{ngpu}" if ngpu > 1 else "single",
            "l2_policy": "operand sets rotate through copies totalling > 126 MB L2" if ngpu == 1 else "operands (>= 2 GiB) exceed L2"}


# ------------------------------------------------------------------------------- GPU arm
def main_gpu(args):
    import wgmath_b200 as w
    from oracle import oracle as O
    from wgmath_b200._lib import check, lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ngpu = args.gpus
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peaks = load_peaks()
    gpu = w.GpuInstance.new(local_rank)
    dev = gpu.device()
    shapes = w.ViewShapeBuffers.new()
    L = lib()
    sampler = ClockSampler(local_rank)
    U = w.BufferUsages
    ST = U.STORAGE | U.COPY_SRC | U.COPY_DST

    def barrier_sync():
        dev.poll_wait()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def timed(fn, steps, warmup, before_end=None):
        """fn(pass, i) enqueues step i.  Returns (seconds for `steps` steps (max over ranks), launches in region).
        before_end(), if given, runs after the last step and before the closing event is recorded (device-side joins)."""
        e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
        check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
        check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
        enc = dev.create_command_encoder()
        p = enc.compute_pass("bench", None)
        for i in range(warmup):
            fn(p, i)
        if before_end is not None:
            before_end()
        barrier_sync()
        n0 = dev.launch_count()
        with sampler:
            check(L.wgb_event_record(e0, p._h))
            for i in range(steps):
                fn(p, warmup + i)
            if before_end is not None:
                before_end()
            check(L.wgb_event_record(e1, p._h))
            p.end()
            gpu.queue().submit(enc.finish())
            barrier_sync()
        ms = ctypes.c_float()
        check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
        L.wgb_event_destroy(e0)
        L.wgb_event_destroy(e1)
        sec = ms.value / 1e3
        if dist is not None:
            import torch
            t = torch.tensor([sec], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec, dev.launch_count() - n0

    def timed_graph(fn, steps, warmup):
        """Like timed(), but the `steps` dispatches are recorded once into a CUDA graph (wgb_graph_*) and replayed with one
        launch, so kernels of a few microseconds are not hidden behind the Python/ctypes cost of issuing them."""
        e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
        check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
        check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
        enc = dev.create_command_encoder()
        with enc.compute_pass("warm", None) as p:
            for i in range(warmup):
                fn(p, i)
        with dev.capture() as cap:
            with enc.compute_pass("rec", None) as p:
                for i in range(steps):
                    fn(p, warmup + i)
        barrier_sync()
        with sampler:
            check(L.wgb_event_record(e0, None))
            cap.graph.launch()
            check(L.wgb_event_record(e1, None))
            barrier_sync()
        ms = ctypes.c_float()
        check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
        L.wgb_event_destroy(e0)
        L.wgb_event_destroy(e1)
        return ms.value / 1e3, 0

    gemm = w.Gemm.from_device(dev)
    extra = {}

    if ngpu == 1:
        n = 4096
        flops = 2.0 * n * n * n
        nsets = 4  # 4 x (32 + 32 + 32 MiB) = 384 MiB > L2
        sets = []
        enc = dev.create_command_encoder()
        with enc.compute_pass("init", None) as p:
            for s in range(nsets):
                a = w.TensorBuilder.matrix(n, n, ST).build(dev, "bf16")
                b = w.TensorBuilder.matrix(n, n, ST).build(dev, "bf16")
                c = w.TensorBuilder.matrix(n, n, ST).build(dev, "bf16")
                w.fill_uniform(dev, p, a, O.SEED_BASE + 1)
                w.fill_uniform(dev, p, b, O.SEED_BASE + 2)
                sets.append((a, b, c))
        dev.poll_wait()

        def step(p, i):
            a, b, c = sets[i % nsets]
            gemm.dispatch(dev, shapes, p, c, a, b)
        sec, launches = timed(step, args.steps, args.warmup)
        ms_step = sec * 1e3 / args.steps
        value = flops * args.steps / sec / 1e12
        enc = dev.create_command_encoder()
        with enc.compute_pass("path", None) as p:
            step(p, 0)
            gemm_path = p.last_gemm_path()
        dev.poll_wait()

        # ---- e2e: HOST buffers in, HOST buffer out, through the C ABI; every step uploads A and B and downloads C.
        # (a) wgb_gemm_host_enqueue: products queued back to back (wgpu's submit-now / read-later model); the download of
        #     product i overlaps the upload of product i+1, outputs alternate between two host buffers
        # (b) wgb_gemm_host: one blocking call per product (column-panel pipeline inside the call)
        # (c) the reference tests' sequence with separate calls: write A, write B, dispatch, blocking read (gemm.rs:156-193)
        hbytes = n * n * 2
        ha, hb, hc, hc2 = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        for h in (ha, hb, hc, hc2):
            check(L.wgb_host_alloc(hbytes, ctypes.byref(h)))
        a0, b0, c0 = sets[0]
        check(L.wgb_buffer_read(dev._h, a0.buffer()._h, 0, ha, hbytes))
        check(L.wgb_buffer_read(dev._h, b0.buffer()._h, 0, hb, hbytes))
        e2e_steps = max(3, min(args.steps, 20))

        def e2e_enqueue_step(p, i):
            gemm.enqueue_host(dev, n, n, n, hc if i % 2 == 0 else hc2, ha, hb, in_dtype="bf16", out_dtype="bf16")
        e2e_sec, _ = timed(e2e_enqueue_step, e2e_steps, 2, before_end=lambda: gemm.flush_host(dev))
        e2e_val = flops * e2e_steps / e2e_sec / 1e12

        def e2e_host_step(p, i):
            gemm.dispatch_host(dev, n, n, n, hc, ha, hb, in_dtype="bf16", out_dtype="bf16")
        e2e_sec1, _ = timed(e2e_host_step, e2e_steps, 2)
        e2e_val_blocking = flops * e2e_steps / e2e_sec1 / 1e12

        def e2e_step(p, i):
            a, b, c = sets[i % nsets]
            check(L.wgb_buffer_write(dev._h, a.buffer()._h, 0, ha, hbytes))
            check(L.wgb_buffer_write(dev._h, b.buffer()._h, 0, hb, hbytes))
            gemm.dispatch(dev, shapes, p, c, a, b)
            check(L.wgb_buffer_read(dev._h, c.buffer()._h, 0, hc, hbytes))   # blocking D2H of the result
        e2e_sec2, _ = timed(e2e_step, e2e_steps, 2)
        e2e_val_seq = flops * e2e_steps / e2e_sec2 / 1e12
        for h in (ha, hb, hc, hc2):
            L.wgb_host_free(h)

        dtype = "bf16"
        peak = peaks["bf16_tflops"] if sec < 1.0 else peaks["bf16_tflops_sustained"]
        roof = {"bound": "tensor", "kernel": {2: "gemm_tc<bf16> (tcgen05)", 1: "gemm_simt (FFMA)"}.get(gemm_path, f"path {gemm_path}"),
                "achieved": value, "peak": peak, "unit": "TFLOP/s", "frac": value / peak,
                "peak_source": f"{peaks['source']} ({'burst' if sec < 1.0 else 'sustained'})",
                # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel, from the committed ncu --set full
                # capture profiles/r1_gemm_tc_bf16_4096_ncu.txt (88.5 MB + 14.3 MB; algorithmic A + B + C = 100.7 MB)
                "traffic": 102.8e6 if gemm_path == 2 else None, "traffic_unit": "bytes per launch (ncu)",
                "algorithmic": "2*M*N*K flop per launch"}
        e2e = {"value": e2e_val, "unit": "TFLOP/s", "h2d_bytes_per_step": 2 * hbytes, "d2h_bytes_per_step": hbytes,
               "steps": e2e_steps,
               "call": "wgb_gemm_host_enqueue per step (pinned host buffers; upload / panel GEMMs / download on three streams, "
                       "product i's download under product i+1's upload), closed by wgb_gemm_host_flush",
               "ms_per_step": e2e_sec * 1e3 / e2e_steps,
               "blocking_call": {"value": e2e_val_blocking, "ms_per_step": e2e_sec1 * 1e3 / e2e_steps,
                                 "call": "wgb_gemm_host (one blocking call per product)"},
               "separate_calls": {"value": e2e_val_seq, "ms_per_step": e2e_sec2 * 1e3 / e2e_steps,
                                  "call": "wgb_buffer_write x2 + wgb_gemm_ex + wgb_buffer_read"}}
        if not args.no_extras:
            sets.clear()
            extra = run_extras(w, O, gpu, shapes, timed, peaks, timed_graph)
        cpu_tf, cores, desc, _, _ = cpu_gemm_sample(10.0)
        cpu = {"value": cpu_tf, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": desc}
    else:
        from wgmath_b200 import sharded
        res = sharded.bench_row_sharded(w, O, gpu, shapes, dist, ngpu, rank, args, timed, peaks)
        value, ms_step, launches, roof, e2e, dtype = res
        cpu = None

    if rank == 0:
        line = {"metric": "gemm_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": ngpu, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": workload_config(ngpu),
                "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roof}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extra:
            line["extra"] = extra
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_extras(w, O, gpu, shapes, timed, peaks, timed_graph):
    """The other BASELINE configs, each timed like the headline: configs[1] f32 GEMM sweep, configs[3] GEMV +
    level-1 GB/s.  Operands larger than L2 (or rotated) so no run is served from cache."""
    dev = gpu.device()
    U = w.BufferUsages
    ST = U.STORAGE | U.COPY_SRC | U.COPY_DST
    out = {"gemm_f32": [], "level12": []}
    gemm = w.Gemm.from_device(dev)
    # ---- f32 GEMM sweep
    for n in (256, 512, 1024, 2048, 4096, 8192):
        nsets = max(1, min(256, int(np.ceil(L2_BYTES * 1.5 / (3 * n * n * 4)))))   # rotating operand sets > L2
        sets = []
        enc = dev.create_command_encoder()
        with enc.compute_pass("init", None) as p:
            for _ in range(nsets):
                a, b, c = (w.TensorBuilder.matrix(n, n, ST).build(dev) for _ in range(3))
                w.fill_uniform(dev, p, a, O.SEED_BASE + 1)
                w.fill_uniform(dev, p, b, O.SEED_BASE + 2)
                sets.append((a, b, c))
        dev.poll_wait()
        for mode, name in ((w.F32Mode.X3Tf32, "3xtf32"), (w.F32Mode.Tf32, "tf32")):
            steps = 20 if n <= 2048 else (8 if n == 4096 else 4)
            path = []

            def step(p, i, mode=mode):
                a, b, c = sets[i % nsets]
                gemm.dispatch_generic(dev, shapes, p, c, a, b, w.GemmVariant.Gemm, f32_mode=mode)
                if not path:
                    path.append(p.last_gemm_path())
            sec, _ = timed_graph(step, steps, 3)          # device time of `steps` back-to-back GEMMs, one graph launch
            tf = 2.0 * n ** 3 * steps / sec / 1e12
            # TF32 dense peak is half the bf16 peak; 3xTF32 issues three MMAs per product
            peak = peaks["bf16_tflops"] / 2 / (3 if name == "3xtf32" else 1)
            out["gemm_f32"].append({"n": n, "mode": name, "path": path[0] if path else None, "tflops": tf,
                                    "ms": sec * 1e3 / steps, "frac_of_tf32_peak": tf / peak, "peak": peak})
        del sets
    # ---- GEMV 65536 x 4096 (1 GiB matrix: larger than L2) and level-1 at n = 2^26
    M, K = 65536, 4096
    m = w.TensorBuilder.matrix(M, K, ST).build(dev)
    x, y = w.TensorBuilder.vector(K, ST).build(dev), w.TensorBuilder.vector(M, ST).build(dev)
    xm, yk = w.TensorBuilder.vector(M, ST).build(dev), w.TensorBuilder.vector(K, ST).build(dev)
    enc = dev.create_command_encoder()
    with enc.compute_pass("init", None) as p:
        w.fill_uniform(dev, p, m, O.SEED_BASE + 1)
        w.fill_uniform(dev, p, x, O.SEED_BASE + 3)
        w.fill_uniform(dev, p, xm, O.SEED_BASE + 3)
    dev.poll_wait()
    gemv = w.Gemv.from_device(dev)
    hbm = peaks["hbm_gbs"]

    def rec(name, nbytes, fn, steps=20):
        sec, _ = timed(fn, steps, 3)
        gbs = nbytes * steps / sec / 1e9
        out["level12"].append({"op": name, "bytes": nbytes, "ms": sec * 1e3 / steps, "gbs": gbs, "frac_of_hbm": gbs / hbm})
    rec("gemv 65536x4096", 4 * (M * K + K + M), lambda p, i: gemv.dispatch(dev, shapes, p, y, m, x))
    rec("gemv_tr 65536x4096", 4 * (M * K + K + M), lambda p, i: gemv.dispatch_tr(dev, shapes, p, yk, m, xm))
    colsum = w.Reduce.new(dev, w.ReduceOp.Sum)
    rec("column-reduce(sum) 65536x4096", 4 * (M * K + K), lambda p, i: colsum.dispatch_columns(dev, shapes, p, m, yk))
    # gemv_tr on the transpose-shaped 4096 x 65536 view of the same buffer
    mt = m.reshape((K, M))
    rec("gemv_tr 4096x65536", 4 * (M * K + K + M), lambda p, i: gemv.dispatch_tr(dev, shapes, p, y, mt, x))
    rec("gemv 4096x65536", 4 * (M * K + K + M), lambda p, i: gemv.dispatch(dev, shapes, p, yk, mt, xm))
    del m
    n = 1 << 26
    a, b = w.TensorBuilder.vector(n, ST).build(dev), w.TensorBuilder.vector(n, ST).build(dev)
    res = w.TensorBuilder.scalar(ST).build(dev)
    enc = dev.create_command_encoder()
    with enc.compute_pass("init", None) as p:
        w.fill_uniform(dev, p, a, O.SEED_BASE + 1)
        w.fill_uniform(dev, p, b, O.SEED_BASE + 2)
    dev.poll_wait()
    add, cpy = w.OpAssign.new(dev, w.OpAssignVariant.Add), w.OpAssign.new(dev, w.OpAssignVariant.Copy)
    rsum, rsq, dot = w.Reduce.new(dev, w.ReduceOp.Sum), w.Reduce.new(dev, w.ReduceOp.SqNorm), w.Dot.new(dev)
    rec("op_assign add n=2^26", 12 * n, lambda p, i: add.dispatch(dev, shapes, p, a, b))
    rec("op_assign copy n=2^26", 8 * n, lambda p, i: cpy.dispatch(dev, shapes, p, a, b))
    rec("reduce sum n=2^26", 4 * n, lambda p, i: rsum.dispatch(dev, shapes, p, a, res))
    rec("reduce sqnorm n=2^26", 4 * n, lambda p, i: rsq.dispatch(dev, shapes, p, a, res))
    rec("dot n=2^26", 8 * n, lambda p, i: dot.dispatch(dev, shapes, p, a, b, res))
    del a, b
    # ---- integer primitives next to the linalg path: exclusive scan and key/value radix sort, n = 2^26
    out["scan_sort"] = []
    rng = np.random.default_rng(O.SEED_BASE)
    keys_h = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    keys = w.TensorBuilder.vector(n, ST).build_init(dev, keys_h, "u32")
    vals = w.TensorBuilder.vector(n, ST).build_init(dev, np.arange(n, dtype=np.uint32), "u32")
    okeys, ovals = w.TensorBuilder.vector(n, ST).build(dev, "u32"), w.TensorBuilder.vector(n, ST).build(dev, "u32")
    scan_data = w.TensorBuilder.vector(n, ST).build_init(dev, keys_h & np.uint32(0xFF), "u32")
    del keys_h
    nsort = w.TensorBuilder.scalar(ST).build_init(dev, np.array([n], np.uint32), "u32")
    psum, rsort = w.WgPrefixSum.from_device(dev), w.RadixSort.from_device(dev)
    psw, rsw = w.PrefixSumWorkspace.new(), w.RadixSortWorkspace.new(dev)

    def rec2(name, nbytes, model, fn, steps=10):
        sec, _ = timed(fn, steps, 3)
        gbs = nbytes * steps / sec / 1e9
        out["scan_sort"].append({"op": name, "bytes": nbytes, "traffic_model": model, "ms": sec * 1e3 / steps, "gbs": gbs,
                                 "frac_of_hbm": gbs / hbm, "gelems_per_s": n * steps / sec / 1e9})
    rec2("prefix_sum u32 n=2^26 (in place)", 8 * n, "read + write, single pass", lambda p, i: psum.dispatch(dev, p, psw, scan_data))
    # onesweep traffic: 4 B/pair histogram + per 8-bit digit (8 B read + 8 B written); the reference's 4-bit passes: 20 B x 8
    rec2("radix_sort (u32 key, u32 value) n=2^26, 32 bits", (4 + 4 * 16) * n, "4 + 4 digits x 16 B per pair",
         lambda p, i: rsort.dispatch(dev, p, rsw, keys, vals, nsort, 32, okeys, ovals))
    rec2("radix_sort (u32 key, u32 value) n=2^26, 16 bits", (4 + 2 * 16) * n, "4 + 2 digits x 16 B per pair",
         lambda p, i: rsort.dispatch(dev, p, rsw, keys, vals, nsort, 16, okeys, ovals))
    del keys, vals, okeys, ovals, scan_data
    # ---- batched small-matrix factorizations: out[i] = f(in[i]) over 2^22 (2x2: 2^24) matrices, one per thread.
    # algorithmic bytes = (input matrix + output struct) per element; the iterative ones (eig3/4, svd3) are FP32-pipe bound
    from wgmath_b200 import geometry as G
    out["geometry"] = []
    for dim in (2, 3, 4):
        ng, lg = (1 << 24, 24) if dim == 2 else (1 << 22, 22)   # >= 192 MB of input either way (> the 126 MB L2)
        # inputs generated on the device: U[0,1) words, then A^T A + I written over them would need another kernel, so the
        # symmetric ops are timed on s = (a + a^T) / 2 + dim * I built once on the host for 2^16 matrices and tiled
        base = np.random.default_rng(O.SEED_BASE + dim).random((1 << 16, dim, dim)).astype(np.float32)
        sym = ((base + np.transpose(base, (0, 2, 1))) * 0.5 + dim * np.eye(dim, dtype=np.float32)).astype(np.float32)
        gen_in = w.TensorBuilder.vector(ng, ST).build_init(dev, np.tile(G.pack(base), ng >> 16), f"mat{dim}")
        sym_in = w.TensorBuilder.vector(ng, ST).build_init(dev, np.tile(G.pack(sym), ng >> 16), f"mat{dim}")
        ops = [("cholesky", getattr(w, f"WgCholesky{dim}"), sym_in, f"mat{dim}"), ("lu", getattr(w, f"WgLU{dim}"), gen_in, f"lu{dim}"),
               ("qr", getattr(w, f"WgQR{dim}"), gen_in, f"qr{dim}"), ("symmetric_eigen", getattr(w, f"WgSymmetricEigen{dim}"), sym_in, f"eig{dim}")]
        if dim < 4:
            ops.append(("svd", getattr(w, f"WgSvd{dim}"), gen_in, f"svd{dim}"))
        for name, cls, src, odt in ops:
            dst = w.TensorBuilder.vector(ng, ST).build(dev, odt)
            sh = cls.from_device(dev)
            nbytes = ng * (G.Matrix[dim].itemsize + cls.OUT_TYPE.itemsize)
            sec, _ = timed(lambda p, i: sh.dispatch(dev, p, src, dst), 10, 3)
            gbs = nbytes * 10 / sec / 1e9
            out["geometry"].append({"op": f"{name}{dim} n=2^{lg}", "bytes": nbytes, "ms": sec * 1e2, "gbs": gbs, "frac_of_hbm": gbs / hbm,
                                    "gmat_per_s": ng * 10 / sec / 1e9})
            del dst
        inv_dst = w.TensorBuilder.vector(ng, ST).build(dev, f"mat{dim}")
        winv = w.WgInv.from_device(dev)
        nbytes = ng * 2 * G.Matrix[dim].itemsize
        sec, _ = timed(lambda p, i: winv.dispatch(dev, p, dim, sym_in, inv_dst), 10, 3)
        out["geometry"].append({"op": f"inv{dim} n=2^{lg}", "bytes": nbytes, "ms": sec * 1e2, "gbs": nbytes * 10 / sec / 1e9,
                                "frac_of_hbm": nbytes * 10 / sec / 1e9 / hbm, "gmat_per_s": ng * 10 / sec / 1e9})
        del gen_in, sym_in, inv_dst
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
