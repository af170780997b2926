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


def traffic_from_profile(kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, read from the committed summary of
    the `ncu --set full` capture under profiles/ (newest round first).  Returns (bytes, file) or (None, None)."""
    import glob
    import re
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_tc_bf16_4096_ncu.txt")), reverse=True):
        rd = wr = None
        in_kernel = False
        for line in open(path):
            if line.startswith("## "):
                if rd is not None and wr is not None:
                    break
                # (the kernel template has trailing defaulted parameters: match the instantiation up to where the name given ends)
                stem = kernel_substr.rstrip(">")
                in_kernel = (stem + ">") in line or (stem + ",") in line
                continue
            if not in_kernel:
                continue
            m = re.match(r"\s+DRAM bytes (read|written)\s+([0-9.]+)\s+(\w+)", line)
            if m:
                v = float(m.group(2)) * unit.get(m.group(3), 1.0)
                if m.group(1) == "read":
                    rd = v
                else:
                    wr = v
        if rd is not None and wr is not None:
            return rd + wr, os.path.relpath(path, ROOT)
    return None, None


# ------------------------------------------------------------------------------- NVML clocks
class ClockSampler:
    """Samples SM clock + throttle reasons during timed regions (B200_PROFILING.md clocks line, via NVML)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.samples, self.power, self.reasons, self.max_mhz, self.ok = [], [], set(), None, False
        self._active = threading.Event()
        self._stop = threading.Event()
        self._lock = threading.Lock()
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

    def sample_now(self):
        """One sample, taken by the caller's thread.  The timing loops call it between queueing the timed steps and waiting for
        them, so that a window has at least one sample under load even when the polling thread is starved (seen once at 8
        ranks: NVML answered none of the thread's calls inside a 141 ms window)."""
        if not self.ok:
            return
        nv = self._nv
        try:
            mhz = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
            watts = nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        except Exception:
            return
        with self._lock:
            self.samples.append(mhz)
            self.power.append(watts)
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            if self._active.is_set():
                self.sample_now()
                time.sleep(0.002)
            else:
                time.sleep(0.0005)

    def __enter__(self):
        self._active.set()
        return self

    def __exit__(self, *a):
        self._active.clear()

    def mark(self):
        """Start of a window: summary(since=mark) describes only the samples taken after it."""
        return len(self.samples), set(self.reasons)

    def summary(self, since=None):
        if since is not None and self.ok:
            n0, before = since
            smp, pw = self.samples[n0:], self.power[n0:]
            return {"sm_mhz": float(np.median(smp)) if smp else None, "sm_mhz_min": float(min(smp)) if smp else None,
                    "sm_max_mhz": self.max_mhz, "power_w_median": float(np.median(pw)) if pw else None,
                    "power_w_max": float(max(pw)) if pw else None, "reasons": sorted(self.reasons), "samples": len(smp)}
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "power_w_max": float(max(self.power)) if self.power else None,
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
            f"(reproduces the shader's output bit for bit: tests/test_reference_vectors.py; the reference's own wgpu fallback "
            f"adapter cannot run here: no Rust/Vulkan in the image), OpenMP over invocations")
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
            sampler.sample_now()          # the device is still working through the queued steps here
            barrier_sync()
        ms = ctypes.c_float()
        check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
        L.wgb_event_destroy(e0)
        L.wgb_event_destroy(e1)
        sec = ms.value / 1e3
        if dist is not None:
            import torch
            mine = torch.tensor([sec], device="cuda")
            every = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
            timed.per_rank_ms = [float(x.item()) * 1e3 / max(steps, 1) for x in every]   # of the last timed region
            sec = max(float(x.item()) for x in every)
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

        def make_sets():
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
        make_sets()

        def step(p, i):
            a, b, c = sets[i % nsets]
            gemm.dispatch(dev, shapes, p, c, a, b)
        main_mark = sampler.mark()
        sec, launches = timed(step, args.steps, args.warmup)
        clocks_main = sampler.summary(since=main_mark)
        ms_step = sec * 1e3 / args.steps
        value = flops * args.steps / sec / 1e12
        enc = dev.create_command_encoder()
        with enc.compute_pass("path", None) as p:
            step(p, 0)
            gemm_path = p.last_gemm_path()
            gemm_cfg = p.last_gemm_config()
        dev.poll_wait()
        # ---- parity of what was just timed: sampled rows of C (set 0: bf16 in, bf16 OUT, random operands) against float64 on
        # the same bf16-rounded inputs, the 1e-2 bound of BASELINE.json; a mismatch fails the run
        a0, b0, c0 = sets[0]
        rows = np.array([0, 1, 127, 128, 2049, 4095])
        got = O.bf16_from_bits(c0.read()).reshape(n, n).T[rows].astype(np.float64)
        A64 = np.stack([O.to_bf16_rne(O.uniform(O.SEED_BASE + 1, 1, n, row0=int(r))) for r in rows]).astype(np.float64)
        B64 = O.to_bf16_rne(O.uniform(O.SEED_BASE + 2, n, n)).reshape(n, n).T.astype(np.float64)
        ref = A64 @ B64
        perr = float(np.max(np.abs(got - ref) / np.abs(ref)))
        parity = {"checked": f"{len(rows)} rows x {n} columns of the timed product (bf16 out) vs float64 on the same bf16-rounded inputs",
                  "max_rel_err": perr, "tol": 1e-2, "ok": bool(perr < 1e-2)}
        del A64, B64, ref, got

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
        kname = (f"gemm_tc_kernel<{gemm_cfg['kind']}, {gemm_cfg['a_mn']}, {gemm_cfg['b_mn']}, {gemm_cfg['bn']}, {gemm_cfg['passes']}, "
                 f"{'__nv_bfloat16' if gemm_cfg['out_dtype'] == 1 else 'float'}, {gemm_cfg['cg']}>") if gemm_path >= 2 else "gemm_simt (FFMA)"
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of exactly this instantiation, parsed from the committed
        # summary of the `ncu --set full` capture (algorithmic A + B + C = 100.7 MB)
        traffic, traffic_src = traffic_from_profile(kname) if gemm_path >= 2 else (None, None)
        roof = {"bound": "tensor", "kernel": kname + (" (tcgen05)" if gemm_path >= 2 else ""),
                "achieved": value, "peak": peak, "unit": "TFLOP/s", "frac": value / peak,
                "peak_source": f"{peaks['source']} ({'burst' if sec < 1.0 else 'sustained'})",
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu)", "traffic_source": traffic_src,
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
        secondary = None
        if not args.no_extras:
            sets.clear()
            extra = run_extras(w, O, gpu, shapes, timed, peaks, timed_graph)
            # the other BASELINE configs in top-level form (the driver keeps top-level keys): configs[3] GEMV GB/s vs the
            # measured copy bandwidth, configs[1] the f32 sweep's extremes
            l12 = {r["op"]: r for r in extra["level12"]}
            x3 = [r for r in extra["gemm_f32"] if r["mode"] == "3xtf32"]
            secondary = {
                "gemv_65536x4096": {"bound": "hbm", "achieved": l12["gemv 65536x4096"]["gbs"], "unit": "GB/s", "peak": peaks["hbm_gbs"],
                                    "frac": l12["gemv 65536x4096"]["frac_of_hbm"], "ms": l12["gemv 65536x4096"]["ms"]},
                "gemv_tr_65536x4096": {"bound": "hbm", "achieved": l12["gemv_tr 65536x4096"]["gbs"], "unit": "GB/s", "peak": peaks["hbm_gbs"],
                                       "frac": l12["gemv_tr 65536x4096"]["frac_of_hbm"], "ms": l12["gemv_tr 65536x4096"]["ms"]},
                "level1_worst": min(({"op": r["op"], "gbs": r["gbs"], "frac": r["frac_of_hbm"]} for r in extra["level12"]), key=lambda r: r["frac"]),
                "gemm_f32_3xtf32": {str(r["n"]): {"tflops": r["tflops"], "frac_of_tf32_peak_div3": r["frac_of_tf32_peak"], "path": r["path"]} for r in x3},
                "gemm_f32_3xtf32_worst_frac": min(r["frac_of_tf32_peak"] for r in x3),
                "peak_source": peaks["source"]}
        # ---- sustained: the same GEMM back to back for >= 2 s (power-limited regime), against the sustained cuBLAS peak.  Last
        # of the GPU work, so that no other measurement of this run starts on a part that is already at its power limit.
        if not sets:
            make_sets()
        sus_steps = max(200, int(2.2 / (ms_step / 1e3)))
        mark = sampler.mark()
        sus_sec, _ = timed(step, sus_steps, 3)
        sus_val = flops * sus_steps / sus_sec / 1e12
        roof["sustained"] = {"value": sus_val, "unit": "TFLOP/s", "seconds": sus_sec, "steps": sus_steps,
                             "peak": peaks["bf16_tflops_sustained"], "frac": sus_val / peaks["bf16_tflops_sustained"],
                             "peak_source": f"{peaks['source']} (sustained)", "clocks": sampler.summary(since=mark)}
        sets.clear()
        cpu_tf, cores, desc, _, _ = cpu_gemm_sample(10.0)
        cpu = {"value": cpu_tf, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": desc}
    else:
        res = bench_row_sharded(w, O, gpu, shapes, dist, ngpu, rank, args, timed, peaks, sampler)
        value, ms_step, launches, roof, e2e, dtype, parity, wl_extra, clocks_main = res
        cpu = None
        secondary = None

    if rank == 0:
        line = {"metric": "gemm_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": ngpu, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": workload_config(ngpu) | (wl_extra if ngpu > 1 else {}),
                "clocks": clocks_main | {"window": "the timed region of `value`", "all_timed_regions": sampler.summary()}, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roof, "parity": parity}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if secondary is not None:
            line["roofline_secondary"] = secondary
        if extra:
            line["extra"] = extra
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if not parity["ok"]:
        print(f"bench.py: PARITY FAILURE on rank {rank}: {parity}", file=sys.stderr)
        return 3
    return 0


def bench_row_sharded(w, O, gpu, shapes, dist, ngpu, rank, args, timed, peaks, sampler):
    """N > 1: bf16 (4096*P)^3, rank p owns the row block p of A and of C, B replicated, the all-gather of C fused into the GEMM
    epilogue (TMA bulk stores into every rank's gathered buffer over NVLink).  Two timings of the same K steps:
      pipelined (the headline `value`): three rotating gathered buffers, the wait for step e's panels queued after the GEMM of
          step e + 1 (wgb_peer_gather_wait), the last one inside the timed region — every step's GEMM and gather complete before
          the closing event, but a momentarily slow rank no longer stalls the other seven at every step;
      lock step (`lockstep`): one gathered buffer, every call waits for all panels before the next GEMM starts.
    Then parity: sampled elements of the gathered C of the timed run on EVERY rank against float64."""
    from wgmath_b200 import sharded
    from wgmath_b200._lib import check, lib
    dev = gpu.device()
    n = 4096 * ngpu
    m_local = n // ngpu
    ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
    mode = os.environ.get("WGB_SHARD_MODE", "fused")                     # "fused" (peer stores) or "nccl"
    a = w.TensorBuilder.matrix(m_local, n, ST).build(dev, "bf16")       # my row block of A
    b = w.TensorBuilder.matrix(n, n, ST).build(dev, "bf16")             # B replicated
    enc = dev.create_command_encoder()
    with enc.compute_pass("init", None) as p:
        w.fill_uniform(dev, p, a, O.SEED_BASE + 1, row0=rank * m_local)  # element (i, j) independent of the sharding
        w.fill_uniform(dev, p, b, O.SEED_BASE + 2)
    dev.poll_wait()
    op = sharded.RowShardedGemm(dev)
    flops = 2.0 * n * n * n
    lock = None
    cfg = {}
    if mode == "fused":
        # gathered buffers in symmetric memory with an NVSwitch multicast mapping when the box offers one (every output block is
        # then stored once and replicated by the switch); otherwise CUDA-IPC peer mappings and one TMA bulk store per rank
        symmetric = os.environ.get("WGB_SHARD_SYMMETRIC", "1") not in ("", "0")
        sym_note = None

        def make_group(depth):
            nonlocal symmetric, sym_note
            if symmetric:
                try:
                    return sharded.PeerGather(dev, dist, rank, ngpu, m_local, n, "bf16", depth=depth, symmetric=True)
                except Exception as e:      # no symmetric-memory support here: every rank takes the same fallback
                    symmetric, sym_note = False, repr(e)[:200]
            return sharded.PeerGather(dev, dist, rank, ngpu, m_local, n, "bf16", depth=depth)
        group = make_group(3)
        calls = [0]

        def step_fn(p, i):
            op.dispatch_fused(dev, shapes, p, group, a, b, wait=False)
            if not cfg:
                cfg.update(p.last_gemm_config())
            if calls[0] > 0:
                group.wait(p, 1)          # the gather of the previous call, queued behind this call's GEMM
            calls[0] += 1
        def closing_wait():
            # the last call's panels: part of the timed region
            enc2 = dev.create_command_encoder()
            with enc2.compute_pass("gather-wait", None) as p2:
                group.wait(p2, 0)
        mark = sampler.mark()
        sec, launches = timed(step_fn, args.steps, args.warmup, before_end=closing_wait)
        clocks_main = sampler.summary(since=mark)
        by_rank = getattr(timed, "per_rank_ms", None)
        c = group.tensor_at(0)
        multicast = bool(group.multicast)
    else:
        group = None
        by_rank = None
        sharded.init_comm(dev, dist, rank, ngpu)
        c = w.TensorBuilder.tensor((m_local, n, ngpu), ST).build(dev, "bf16")
        step_fn = lambda p, i: op.dispatch(dev, shapes, p, c, a, b)             # noqa: E731
        mark = sampler.mark()
        sec, launches = timed(step_fn, args.steps, args.warmup)
        clocks_main = sampler.summary(since=mark)
    value = flops * args.steps / sec / 1e12
    ms_step = sec * 1e3 / args.steps
    if mode == "fused":
        # the same steps in lock step (one gathered buffer, the wait inside every call): what the deferral is worth.  Measured
        # after the headline, i.e. on parts that are already warm.
        group1 = make_group(1)

        def lock_step(p, i):
            op.dispatch_fused(dev, shapes, p, group1, a, b)
        lsec, _ = timed(lock_step, args.steps, args.warmup)
        lock = {"value": flops * args.steps / lsec / 1e12, "unit": "TFLOP/s", "ms_per_step": lsec * 1e3 / args.steps,
                "ms_per_step_by_rank": getattr(timed, "per_rank_ms", None),
                "what": "one gathered buffer, every call waits for all peers' panels before the next GEMM starts; measured after the "
                        "headline run (warm parts)"}
        group1.close()
    # ---- parity, on every rank: elements (row r, column j) of the gathered cube for rows spread over every rank's panel
    L = lib()
    rows = sorted({0, 1, m_local - 1, m_local, n // 2 + 17, n - m_local - 1, n - 1, (5 * m_local + 4095) % n})
    cols = sorted({0, 1, 255, 256, n // 2 - 1, n // 2, n - 257, n - 1} | {int(x) for x in np.random.default_rng(7).integers(0, n, 24)})
    Bc = np.stack([O.to_bf16_rne(O.uniform(O.SEED_BASE + 2, n, 1, col0=j)) for j in cols]).astype(np.float64)      # [col][k]
    Ar = np.stack([O.to_bf16_rne(O.uniform(O.SEED_BASE + 1, 1, n, row0=r)) for r in rows]).astype(np.float64)      # [row][k]
    ref = Ar @ Bc.T
    got = np.zeros_like(ref)
    seg = np.zeros(m_local, np.uint16)
    for cj, j in enumerate(cols):
        for q in sorted({r // m_local for r in rows}):
            off = (q * m_local * n + j * m_local) * 2
            check(L.wgb_buffer_read(dev._h, c.buffer()._h, off, seg.ctypes.data_as(ctypes.c_void_p), seg.nbytes))
            vals = O.bf16_from_bits(seg)
            for ri, r in enumerate(rows):
                if r // m_local == q:
                    got[ri, cj] = vals[r % m_local]
    perr = float(np.max(np.abs(got - ref) / np.abs(ref)))
    ok_local = bool(perr < 1e-2)
    worst, ok_all = perr, ok_local
    if dist is not None:
        import torch
        t = torch.tensor([perr, 0.0 if ok_local else 1.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst, ok_all = float(t[0].item()), bool(t[1].item() == 0.0)
    parity = {"checked": f"{len(rows)} rows (every rank's panel) x {len(cols)} columns of the gathered C of the timed run, on every "
                         f"one of the {ngpu} ranks, vs float64 on the same bf16-rounded inputs", "max_rel_err": worst, "tol": 1e-2,
              "ok": ok_all, "ranks_checked": ngpu}
    long_run = sec > 1.0
    peak = (peaks["bf16_tflops_sustained"] if long_run else peaks["bf16_tflops"]) * ngpu
    comm_bytes = (ngpu - 1) * m_local * n * 2
    kname = ((f"gemm_tc_kernel<{cfg.get('kind')}, {cfg.get('a_mn')}, {cfg.get('b_mn')}, {cfg.get('bn')}, {cfg.get('passes')}, __nv_bfloat16, "
              f"{cfg.get('cg')}> (tcgen05) with the all-gather of C fused into the epilogue: "
              f"{ {0: 'per-lane stores to', 1: 'TMA bulk stores to', 2: 'multimem.st (NVSwitch multicast), one store for'}[cfg.get('epi_tma', 0)]}"
              f" {cfg.get('dests')} gathered buffers over NVLink")
             if group is not None else "gemm_tc<bf16> (tcgen05) + chunked all-gather of C (NCCL send/recv over NVLink)")
    flop_ms = flops / ngpu / (peaks["bf16_tflops_sustained" if long_run else "bf16_tflops"] * 1e12) * 1e3
    link_ms = comm_bytes / 770e9 * 1e3
    roof = {"bound": "tensor", "kernel": kname,
            "achieved": value, "peak": peak, "unit": "TFLOP/s", "frac": value / peak,
            "peak_source": f"{peaks['source']} ({'sustained' if long_run else 'burst'}) x {ngpu} GPUs", "traffic": None,
            "algorithmic": "2*M*N*K flop per step over all ranks", "nvlink_bytes_in_per_gpu_per_step": comm_bytes,
            "nvlink_floor_ms": link_ms, "tensor_floor_ms": flop_ms,
            "fused_target_ms": max(flop_ms, link_ms), "frac_of_fused_target": max(flop_ms, link_ms) / ms_step,
            "ms_per_step_by_rank": by_rank}
    if lock is not None:
        roof["lockstep"] = lock
    wl_extra = {"gather": ("fused into the GEMM epilogue; three rotating gathered buffers, the wait for step e queued behind the GEMM of "
                           "step e + 1, the last wait inside the timed region" if group is not None else "NCCL send/recv, chunked")}
    if group is not None:
        wl_extra["gathered_memory"] = ("symmetric memory + NVSwitch multicast mapping" if multicast else
                                       "symmetric memory, no multicast mapping" if symmetric else "cudaMalloc + CUDA IPC peer mappings")
        if sym_note:
            wl_extra["symmetric_memory_unavailable"] = sym_note
    # e2e: HOST buffers in, HOST buffer out, through the C ABI, every step: this rank's A block and its 1/P column slice of B go
    # up (the slices are all-gathered over NVLink: B crosses the host links once per box), the sharded GEMM + gather runs, this
    # rank's panel comes down — the ranks of the box assemble C in host memory, every byte of C crosses a host link once.
    # `separate_calls`: the reference tests' sequence (write A, write B, dispatch, read the whole gathered cube on every rank).
    abytes, bbytes, cbytes = m_local * n * 2, n * n * 2, n * n * 2
    pbytes = m_local * n * 2
    ha, hb, hc, hp0, hp1 = (ctypes.c_void_p() for _ in range(5))
    for h, nb in ((ha, abytes), (hb, bbytes), (hc, cbytes), (hp0, pbytes), (hp1, pbytes)):
        check(L.wgb_host_alloc(nb, ctypes.byref(h)))
    check(L.wgb_buffer_read(dev._h, a.buffer()._h, 0, ha, abytes))
    check(L.wgb_buffer_read(dev._h, b.buffer()._h, 0, hb, bbytes))
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_seq_step(p, i):
        check(L.wgb_buffer_write(dev._h, a.buffer()._h, 0, ha, abytes))
        check(L.wgb_buffer_write(dev._h, b.buffer()._h, 0, hb, bbytes))
        if group is not None:
            op.dispatch_fused(dev, shapes, p, group, a, b)
            check(L.wgb_buffer_read(dev._h, group.tensor_at(0).buffer()._h, 0, hc, cbytes))
        else:
            step_fn(p, i)
            check(L.wgb_buffer_read(dev._h, c.buffer()._h, 0, hc, cbytes))
    seq_sec, _ = timed(e2e_seq_step, e2e_steps, 1)
    seq = {"value": flops * e2e_steps / seq_sec / 1e12, "ms_per_step": seq_sec * 1e3 / e2e_steps,
           "h2d_bytes_per_step": (abytes + bbytes) * ngpu, "d2h_bytes_per_step": cbytes * ngpu,
           "call": "wgb_buffer_write x2 + sharded dispatch + wgb_buffer_read of the whole gathered cube on every rank"}
    if group is not None:
        split_b = os.environ.get("WGB_SHARD_B_UPLOAD", "1") not in ("", "0")
        if split_b:
            sharded.init_comm(dev, dist, rank, ngpu)

        def e2e_step(p, i):
            op.enqueue_host_fused(dev, group, m_local, n, n, hp0 if i % 2 == 0 else hp1, ha, hb)
        e2e_sec, _ = timed(e2e_step, e2e_steps, 2, before_end=lambda: check(L.wgb_gemm_host_flush(dev._h)))
        # the panel this rank downloaded in the last e2e step must be its panel of the gathered cube, bit for bit
        dev.poll_wait()
        last = hp0 if (2 + e2e_steps - 1) % 2 == 0 else hp1
        mine = np.zeros(pbytes // 2, np.uint16)
        check(L.wgb_buffer_read(dev._h, group.tensor_at(0).buffer()._h, rank * pbytes, mine.ctypes.data_as(ctypes.c_void_p), pbytes))
        host = np.ctypeslib.as_array(ctypes.cast(last, ctypes.POINTER(ctypes.c_uint16)), shape=(pbytes // 2,))
        e2e_same = bool(np.array_equal(host, mine))
        if dist is not None:
            import torch
            t = torch.tensor([1.0 if e2e_same else 0.0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            e2e_same = bool(t.item() == 1.0)
        e2e = {"value": flops * e2e_steps / e2e_sec / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": abytes * ngpu + (bbytes if split_b else bbytes * ngpu),
               "d2h_bytes_per_step": pbytes * ngpu, "steps": e2e_steps, "ms_per_step": e2e_sec * 1e3 / e2e_steps,
               "b_upload": "1/P column slice per rank + all-gather over NVLink (ncclAllGather on the queue)" if split_b else "whole B on every rank",
               "downloaded_panel_equals_device_result": e2e_same,
               "call": "wgb_gemm_row_sharded_fused_host_enqueue per step on every rank (pinned host buffers; A block + B slice up, "
                       "all-gather of B, fused GEMM + all-gather of C, this rank's panel of C down: the box's host memory ends with "
                       "all of C), closed by wgb_gemm_host_flush", "separate_calls": seq}
        if not e2e_same:
            parity["ok"] = False
            parity["e2e_panel_mismatch"] = True
    else:
        split_b = False
        e2e = {"value": seq["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": seq["h2d_bytes_per_step"],
               "d2h_bytes_per_step": seq["d2h_bytes_per_step"], "steps": e2e_steps, "ms_per_step": seq["ms_per_step"], "call": seq["call"]}
    for h in (ha, hb, hc, hp0, hp1):
        L.wgb_host_free(h)
    dev.poll_wait()
    if dist is not None:
        dist.barrier()
    if group is None or split_b:
        lib().wgb_comm_destroy(dev._h)
    if group is not None:
        group.close()
    return value, ms_step, launches, roof, e2e, "bf16", parity, wl_extra, clocks_main


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
