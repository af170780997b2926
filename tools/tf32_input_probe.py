"""How does tcgen05.mma kind::tf32 read an f32 operand from shared memory: does it truncate the low 13 mantissa bits or round?
A * I with single-pass TF32 returns exactly what the tensor core saw of A.  (Decides how a GEMM that splits f32 operands into
hi + lo inside the kernel must define `hi`.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgmath_b200 as w  # noqa: E402

ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
gpu = w.GpuInstance.new(0)
dev = gpu.device()
shapes = w.ViewShapeBuffers.new()
gemm = w.Gemm.from_device(dev)
n = 256
rng = np.random.default_rng(1)
A = (rng.random((n, n), dtype=np.float32) * np.float32(4.0) - np.float32(2.0)).astype(np.float32)
bits = A.view(np.uint32)
trunc = (bits & np.uint32(0xFFFFE000)).view(np.float32)
rna = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)          # round half away (cvt.rna.tf32.f32)
rne = ((bits + np.uint32(0x0FFF) + ((bits >> np.uint32(13)) & np.uint32(1))) & np.uint32(0xFFFFE000)).view(np.float32)
eye = np.eye(n, dtype=np.float32)
for tr in (False, True):
    src = A.T if tr else A               # the kernel computes tr(m1) * m2 for the Tr variant
    ta = w.TensorBuilder.matrix(n, n, ST).build_init(dev, np.asfortranarray(src).reshape(-1, order="F"))
    tb = w.TensorBuilder.matrix(n, n, ST).build_init(dev, eye.reshape(-1))
    tc = w.TensorBuilder.matrix(n, n, ST).build(dev)
    enc = dev.create_command_encoder()
    with enc.compute_pass("probe", None) as p:
        gemm.dispatch_generic(dev, shapes, p, tc, ta, tb, w.GemmVariant.GemmTr if tr else w.GemmVariant.Gemm, f32_mode=w.F32Mode.Tf32)
        path = p.last_gemm_path()
    got = tc.read().reshape(n, n).T
    print(f"TF32INPUT tr={int(tr)} path={path}: == truncation {np.array_equal(got, trunc)}  == rna {np.array_equal(got, rna)}  "
          f"== rne {np.array_equal(got, rne)}  == exact f32 {np.array_equal(got, A)}  "
          f"(mismatches vs trunc {int((got != trunc).sum())}, vs rna {int((got != rna).sum())})", flush=True)
