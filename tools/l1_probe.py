"""Runs dot / reduce / op_assign at n = 2^26 a few times (target for an ncu capture)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgmath_b200 as w  # noqa: E402

ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
gpu = w.GpuInstance.new(0)
dev = gpu.device()
shapes = w.ViewShapeBuffers.new()
n = 1 << 26
a, b = w.TensorBuilder.vector(n, ST).build(dev), w.TensorBuilder.vector(n, ST).build(dev)
res = w.TensorBuilder.scalar(ST).build(dev)
enc = dev.create_command_encoder()
with enc.compute_pass("l1", None) as p:
    w.fill_uniform(dev, p, a, 1)
    w.fill_uniform(dev, p, b, 2)
    for _ in range(2):
        w.Dot.new(dev).dispatch(dev, shapes, p, a, b, res)
        w.Reduce.new(dev, w.ReduceOp.Sum).dispatch(dev, shapes, p, a, res)
        w.OpAssign.new(dev, w.OpAssignVariant.Add).dispatch(dev, shapes, p, a, b)
dev.poll_wait()
print("done", res.read())
