"""f32 GEMM timing probe (BASELINE configs[1]): square N = 256 ... 8192, each form recorded once into a CUDA graph of `steps`
dispatches and replayed (device time per GEMM): 3xTF32 with the operand split inside the GEMM (default), 3xTF32 with the split
kernels (WGB_TF32_FUSED_SPLIT=0), single-pass TF32, and the FFMA kernel for the small sizes.  Also max relative error vs float64 on
sampled rows for the parity-gated forms."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgmath_b200 as w  # noqa: E402
from oracle import oracle as O  # noqa: E402
from wgmath_b200._lib import check, lib  # noqa: E402

ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
gpu = w.GpuInstance.new(0)
dev = gpu.device()
shapes = w.ViewShapeBuffers.new()
gemm = w.Gemm.from_device(dev)
L = lib()
PEAK_TF32_DIV3 = float(os.environ.get("PEAK_BF16", "1669.7")) / 2 / 3


def timed_graph(fn, steps):
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
    check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
    enc = dev.create_command_encoder()
    with enc.compute_pass("warm", None) as p:
        for i in range(3):
            fn(p, i)
    with dev.capture() as cap:
        with enc.compute_pass("rec", None) as p:
            for i in range(steps):
                fn(p, 3 + i)
    dev.poll_wait()
    check(L.wgb_event_record(e0, None))
    cap.graph.launch()
    check(L.wgb_event_record(e1, None))
    dev.poll_wait()
    ms = ctypes.c_float()
    check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
    return ms.value / steps


sizes = [int(x) for x in sys.argv[1:]] or [256, 512, 1024, 2048, 4096, 8192]
for n in sizes:
    nsets = max(1, min(64, int(np.ceil(126e6 * 1.5 / (3 * n * n * 4)))))
    sets = []
    enc = dev.create_command_encoder()
    with enc.compute_pass("init", None) as p:
        for _ in range(nsets):
            a, b, c = (w.TensorBuilder.matrix(n, n, ST).build(dev) for _ in range(3))
            w.fill_uniform(dev, p, a, O.SEED_BASE + 1)
            w.fill_uniform(dev, p, b, O.SEED_BASE + 2)
            sets.append((a, b, c))
    dev.poll_wait()
    steps = 20 if n <= 2048 else (8 if n == 4096 else 4)
    rows = np.array([0, 3, n // 2 + 1, n - 1])
    A64 = np.stack([O.uniform(O.SEED_BASE + 1, 1, n, row0=int(r)) for r in rows]).astype(np.float64)
    B64 = O.uniform(O.SEED_BASE + 2, n, n).reshape(n, n).T.astype(np.float64) if n <= 4096 else None
    forms = [("3xtf32 split in the GEMM", w.F32Mode.X3Tf32, {"WGB_TF32_FUSED_SPLIT": "1"}),
             ("3xtf32 split kernels first", w.F32Mode.X3Tf32, {"WGB_TF32_FUSED_SPLIT": "0"}),
             ("tf32 single pass", w.F32Mode.Tf32, {}), ("auto (default)", w.F32Mode.Auto, {})]
    if n <= 1024:
        forms.append(("ffma", w.F32Mode.Simt, {}))
    for name, mode, env in forms:
        for k, v in env.items():
            os.environ[k] = v
        cfgs = []

        def step(p, i, mode=mode):
            a, b, c = sets[i % nsets]
            gemm.dispatch_generic(dev, shapes, p, c, a, b, w.GemmVariant.Gemm, f32_mode=mode)
            if not cfgs:
                cfgs.append((p.last_gemm_path(), p.last_gemm_config()))
        ms = timed_graph(step, steps)
        err = float("nan")
        if B64 is not None:
            got = sets[0][2].read().reshape(n, n).T[rows].astype(np.float64)
            ref = A64 @ B64
            err = float(np.max(np.abs(got - ref) / np.abs(ref)))
        tf = 2.0 * n ** 3 / ms / 1e9
        path, cfg = cfgs[0]
        print(f"F32PROBE n={n:5d} {name:28s}: {ms * 1e3:9.2f} us  {tf:8.1f} TFLOP/s  ({tf / PEAK_TF32_DIV3:5.3f} of TF32 peak / 3)  path {path} "
              f"fs={cfg['fused_split']} nsplit={cfg['nsplit']} units={cfg['units']} cg={cfg['cg']}  max rel err {err:.2e}", flush=True)
        for k in env:
            os.environ.pop(k, None)
    del sets
