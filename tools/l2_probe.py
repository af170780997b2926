"""One launch of each level-1/2 kernel at the BASELINE configs[3] sizes (target for an ncu capture)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgmath_b200 as w  # noqa: E402

ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
gpu = w.GpuInstance.new(0)
dev = gpu.device()
shapes = w.ViewShapeBuffers.new()
M, K = 65536, 4096
m = w.TensorBuilder.matrix(M, K, ST).build(dev)
x, y = w.TensorBuilder.vector(K, ST).build(dev), w.TensorBuilder.vector(M, ST).build(dev)
xm, yk = w.TensorBuilder.vector(M, ST).build(dev), w.TensorBuilder.vector(K, ST).build(dev)
n = 1 << 26
a, b = w.TensorBuilder.vector(n, ST).build(dev), w.TensorBuilder.vector(n, ST).build(dev)
res = w.TensorBuilder.scalar(ST).build(dev)
enc = dev.create_command_encoder()
gemv = w.Gemv.from_device(dev)
with enc.compute_pass("l2", None) as p:
    for t, s in ((m, 1), (x, 3), (xm, 3), (a, 1), (b, 2)):
        w.fill_uniform(dev, p, t, s)
    for _ in range(2):
        gemv.dispatch(dev, shapes, p, y, m, x)
        gemv.dispatch_tr(dev, shapes, p, yk, m, xm)
        w.Reduce.new(dev, w.ReduceOp.Sum).dispatch_columns(dev, shapes, p, m, yk)
        w.Dot.new(dev).dispatch(dev, shapes, p, a, b, res)
        w.Reduce.new(dev, w.ReduceOp.Sum).dispatch(dev, shapes, p, a, res)
        w.OpAssign.new(dev, w.OpAssignVariant.Add).dispatch(dev, shapes, p, a, b)
dev.poll_wait()
print("done", res.read())
