"""Reproduces the exit-time crash seen in the probe scripts (module-level tensors alive at interpreter shutdown)."""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import wgmath_b200 as w
from wgmath_b200._lib import check, lib
L = lib(); gpu = w.GpuInstance.new(0); dev = gpu.device(); shapes = w.ViewShapeBuffers.new()
ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
x = w.TensorBuilder.vector(1024, ST).build(dev)
y = w.TensorBuilder.vector(1024, ST).build(dev)
add = w.OpAssign.new(dev, w.OpAssignVariant.Add)
if mode in ("event", "all"):
    e0 = ctypes.c_void_p(); check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
enc = dev.create_command_encoder()
with enc.compute_pass("t", None) as p:
    add.dispatch(dev, shapes, p, x, y)
    if mode in ("event", "all"):
        check(L.wgb_event_record(e0, p._h))
dev.poll_wait()
if mode in ("lambda", "all"):
    f = lambda q: add.dispatch(dev, shapes, q, x, y)
print("done", mode, flush=True)
