"""Multi-GPU timing probe (run under torch.distributed.run, one rank per GPU): the bench's row-sharded (4096*P)^3 bf16 product in
several forms on the SAME box back to back, so they can be compared (box-to-box variation is several percent):
epilogue store form x gathered-buffer depth / wait placement x destination rotation, plus the GEMM with local stores only
(the box's 8-GPU power envelope without NVLink traffic).  Prints one line per form: ms per step (max over ranks), TFLOP/s."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import wgmath_b200 as w
    from wgmath_b200 import sharded
    from wgmath_b200._lib import check, lib

    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    gpu = w.GpuInstance.new(local_rank)
    dev = gpu.device()
    shapes = w.ViewShapeBuffers.new()
    L = lib()
    ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
    n = 4096 * world
    m_local = 4096
    steps = int(os.environ.get("PROBE_STEPS", "20"))
    a = w.TensorBuilder.matrix(m_local, n, ST).build(dev, "bf16")
    b = w.TensorBuilder.matrix(n, n, ST).build(dev, "bf16")
    enc = dev.create_command_encoder()
    with enc.compute_pass("init", None) as p:
        w.fill_uniform(dev, p, a, 1, row0=rank * m_local)
        w.fill_uniform(dev, p, b, 2)
    dev.poll_wait()
    op = sharded.RowShardedGemm(dev)
    groups = {d: sharded.PeerGather(dev, dist, rank, world, m_local, n, "bf16", depth=d) for d in (1, 3)}
    groups["mc"] = sharded.PeerGather(dev, dist, rank, world, m_local, n, "bf16", depth=3, symmetric=True)
    if rank == 0:
        print(f"SHARDPROBE symmetric group: multicast mapping {groups['mc'].multicast}", flush=True)
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    check(L.wgb_event_create(dev._h, ctypes.byref(e0)))
    check(L.wgb_event_create(dev._h, ctypes.byref(e1)))

    def run(name, env, depth, deferred):
        for k, v in env.items():
            os.environ[k] = v
        g = groups[depth]
        enc = dev.create_command_encoder()
        calls = 0
        with enc.compute_pass("probe", None) as p:
            for phase in range(2):
                if phase == 1:
                    dev.poll_wait()
                    torch.cuda.synchronize()
                    dist.barrier()
                    check(L.wgb_event_record(e0, p._h))
                for _ in range(5 if phase == 0 else steps):
                    op.dispatch_fused(dev, shapes, p, g, a, b, wait=not deferred)
                    if deferred and calls > 0:
                        g.wait(p, 1)
                    calls += 1
                if deferred:
                    g.wait(p, 0)
            check(L.wgb_event_record(e1, p._h))
        dev.poll_wait()
        ms = ctypes.c_float()
        check(L.wgb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
        t = torch.tensor([ms.value], device="cuda")
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        for k in env:
            os.environ.pop(k, None)
        # EVERY rank waits for the collectives before it may queue the next form: a persistent GEMM that already spins on a peer's
        # flag leaves no SM for the NCCL kernel of a rank that is still here, and that peer would never arrive
        per, fastest = float(tmax.item()) / steps, float(tmin.item()) / steps
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            print(f"SHARDPROBE P={world} {name:44s}: {per:7.3f} ms/step (fastest rank {fastest:7.3f})  "
                  f"{2.0 * n ** 3 / per / 1e9:8.1f} TFLOP/s  {2.0 * n ** 3 / per / 1e9 / world:7.1f} per GPU", flush=True)

    forms = [
        ("local stores only (no NVLink), lock step", {"WGB_FUSED_DEBUG_STORE_MASK": "-2", "WGB_TC_EPI": "0"}, 1, False),
        ("per-lane, depth 1, lock step, no rotation", {"WGB_TC_EPI": "0", "WGB_FUSED_ROTATE": "0"}, 1, False),
        ("per-lane, depth 1, lock step, rotated", {"WGB_TC_EPI": "0"}, 1, False),
        ("TMA, depth 1, lock step, no rotation", {"WGB_TC_EPI": "1", "WGB_FUSED_ROTATE": "0"}, 1, False),
        ("TMA, depth 1, lock step, rotated", {"WGB_TC_EPI": "1"}, 1, False),
        ("TMA, depth 3, deferred wait, rotated", {"WGB_TC_EPI": "1"}, 3, True),
        ("per-lane, depth 3, deferred wait, rotated", {"WGB_TC_EPI": "0"}, 3, True),
        ("local stores only (no NVLink), deferred", {"WGB_FUSED_DEBUG_STORE_MASK": "-2", "WGB_TC_EPI": "0"}, 3, True),
        ("multicast (multimem.st), lock step", {"WGB_TC_EPI": "2"}, "mc", False),
        ("multicast (multimem.st), deferred wait", {"WGB_TC_EPI": "2"}, "mc", True),
        ("TMA on symmetric memory, deferred wait", {"WGB_TC_EPI": "1"}, "mc", True),
    ]
    if os.environ.get("PROBE_FORMS"):
        forms = [forms[int(i)] for i in os.environ["PROBE_FORMS"].split(",")]
    for rep in range(int(os.environ.get("PROBE_REPS", "2"))):
        for f in forms:
            run(*f)
    dist.barrier()
    for g in groups.values():
        g.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
