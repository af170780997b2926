"""Device-timed batched small GEMMs: f32 (FFMA tiles 128 / 64 / 32, 3xTF32) and bf16 (tcgen05) over
[M x N x K] x batch.  Usage: python tools/batched_probe.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgmath_b200 as w  # noqa: E402
from wgmath_b200._lib import check, lib  # noqa: E402

ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
gpu = w.GpuInstance.new(0)
dev = gpu.device()
shapes = w.ViewShapeBuffers.new()
L = lib()
gemm = w.Gemm.from_device(dev)


def timed(M, N, K, T, dtype, mode, steps=20):
    a = w.TensorBuilder.tensor((M, K, T), ST).build(dev, dtype)
    b = w.TensorBuilder.tensor((K, N, T), ST).build(dev, dtype)
    c = w.TensorBuilder.tensor((M, N, T), ST).build(dev, dtype)
    enc = dev.create_command_encoder()
    with enc.compute_pass("init", None) as p:
        w.fill_uniform(dev, p, a, 1)
        w.fill_uniform(dev, p, b, 2)
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
    check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
    enc = dev.create_command_encoder()
    with enc.compute_pass("t", None) as p:
        for it in range(steps + 3):
            if it == 3:
                check(L.wgb_event_record(e0, p._h))
            gemm.dispatch_generic(dev, shapes, p, c, a, b, w.GemmVariant.Gemm, f32_mode=mode)
        check(L.wgb_event_record(e1, p._h))
        path = p.last_gemm_path()
    ms = ctypes.c_float()
    check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
    t = ms.value / steps
    es = 2 if dtype == "bf16" else 4
    gbs = (M * K + K * N + M * N) * T * es / t / 1e6
    return t, 2.0 * M * N * K * T / t / 1e9, gbs, path


CASES = [(16, 16, 16, 16384), (32, 32, 32, 8192), (64, 64, 64, 4096), (48, 80, 100, 2048), (128, 128, 64, 2048), (128, 128, 128, 1024),
         (256, 256, 256, 128), (500, 500, 500, 1), (1000, 1000, 1000, 1)]
for (M, N, K, T) in CASES:
    for tile in ("128", "64", "32", ""):
        os.environ["WGB_SIMT_TILE"] = tile
        t, tf, gbs, path = timed(M, N, K, T, "f32", w.F32Mode.Simt)
        print(f"BATCH f32 FFMA tile={tile or 'auto':>4} {M}x{N}x{K} x{T}: {t * 1e3:9.1f} us  {tf:7.2f} TFLOP/s  {gbs:7.0f} GB/s  path {path}", flush=True)
    os.environ["WGB_SIMT_TILE"] = ""
    t, tf, gbs, path = timed(M, N, K, T, "f32", None)
    print(f"BATCH f32 auto           {M}x{N}x{K} x{T}: {t * 1e3:9.1f} us  {tf:7.2f} TFLOP/s  {gbs:7.0f} GB/s  path {path}", flush=True)
    if K % 8 == 0 and M % 8 == 0:
        for cg in ("1", "2"):
            os.environ["WGB_TC_CG"] = cg
            t, tf, gbs, path = timed(M, N, K, T, "bf16", None)
            print(f"BATCH bf16 cg={cg}          {M}x{N}x{K} x{T}: {t * 1e3:9.1f} us  {tf:7.2f} TFLOP/s  {gbs:7.0f} GB/s  path {path}", flush=True)
        os.environ.pop("WGB_TC_CG")
