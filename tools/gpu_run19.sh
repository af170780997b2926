#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_host.py tests/test_gpu_orderings.py -m gpu -q -x 2>&1 | tail -15
timeout 200 python - <<'PY'
# gemv fused vs chain timing at BASELINE configs[3]
import ctypes, numpy as np, sys
sys.path.insert(0, '.')
import wgmath_b200 as w
from wgmath_b200._lib import check, lib
from oracle import oracle as O
L = lib(); gpu = w.GpuInstance.new(0); dev = gpu.device(); shapes = w.ViewShapeBuffers.new()
ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
M, K = 65536, 4096
m = w.TensorBuilder.matrix(M, K, ST).build(dev); x = w.TensorBuilder.vector(K, ST).build(dev); y = w.TensorBuilder.vector(M, ST).build(dev); r = w.TensorBuilder.vector(M, ST).build(dev)
enc = dev.create_command_encoder()
with enc.compute_pass("init", None) as p:
    w.fill_uniform(dev, p, m, 1); w.fill_uniform(dev, p, x, 2); w.fill_uniform(dev, p, r, 3)
dev.poll_wait()
gemv = w.Gemv.from_device(dev); add = w.OpAssign.new(dev, w.OpAssignVariant.Add)
def timed(fn, steps=20):
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.wgb_event_create(dev._h, ctypes.byref(e0))); check(L.wgb_event_create(dev._h, ctypes.byref(e1)))
    enc = dev.create_command_encoder()
    with enc.compute_pass("t", None) as p:
        for _ in range(3): fn(p)
        check(L.wgb_event_record(e0, p._h))
        for _ in range(steps): fn(p)
        check(L.wgb_event_record(e1, p._h))
    dev.poll_wait()
    ms = ctypes.c_float(); check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms))); return ms.value / steps
t1 = timed(lambda p: gemv.dispatch(dev, shapes, p, y, m, x))
t2 = timed(lambda p: (gemv.dispatch(dev, shapes, p, y, m, x), add.dispatch(dev, shapes, p, r, y)))
t3 = timed(lambda p: gemv.dispatch_op(dev, shapes, p, r, m, x, w.OpAssignVariant.Add, r))
print(f"GEMVOP gemv {t1*1e3:.1f} us | gemv + op_assign chain {t2*1e3:.1f} us | fused {t3*1e3:.1f} us")
PY
