"""Multi-GPU parity check for the row-sharded GEMM (run under torch.distributed.run, one rank per GPU):
both exchange paths (NCCL send/recv, fused peer-store epilogue) against float64 on the same seeded inputs.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import wgmath_b200 as w
    from oracle import oracle as O
    from wgmath_b200 import sharded

    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    gpu = w.GpuInstance.new(local_rank)
    dev = gpu.device()
    shapes = w.ViewShapeBuffers.new()
    ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
    ok = True
    for (m_local, N, K, tr) in ((1024, 2048, 1024, False), (512, 1000, 520, True), (2048, 4096, 4096, False)):
        M = m_local * world
        a_blk = O.to_bf16_rne(O.uniform(O.SEED_BASE + 1, K, m_local, col0=rank * m_local) if tr else
                              O.uniform(O.SEED_BASE + 1, m_local, K, row0=rank * m_local))
        b_all = O.to_bf16_rne(O.uniform(O.SEED_BASE + 2, K, N))
        ta = w.TensorBuilder.matrix(*((K, m_local) if tr else (m_local, K)), ST).build_init(dev, O.bf16_bits(a_blk), "bf16")
        tb = w.TensorBuilder.matrix(K, N, ST).build_init(dev, O.bf16_bits(b_all), "bf16")
        op = sharded.RowShardedGemm(dev)
        var = w.GemmVariant.GemmTr if tr else w.GemmVariant.Gemm
        results = {}
        for mode in ("nccl", "fused"):
            if mode == "nccl":
                sharded.init_comm(dev, dist, rank, world)
                c = w.TensorBuilder.tensor((m_local, N, world), ST).build_init(dev, np.zeros(m_local * N * world, np.uint16), "bf16")
                group = None
            else:
                group = sharded.PeerGather(dev, dist, rank, world, m_local, N, "bf16")
                c = group.tensor
            for rep in range(3):                       # repeated steps exercise the epoch / ready / done handshakes
                enc = dev.create_command_encoder()
                with enc.compute_pass("sharded", None) as p:
                    if group is None:
                        op.dispatch(dev, shapes, p, c, ta, tb, var, n_chunks=3)
                    else:
                        op.dispatch_fused(dev, shapes, p, group, ta, tb, var)
                gpu.queue().submit(enc.finish())
                got = c.read()
            results[mode] = got
            if group is not None:
                # host-operand, enqueued form (wgb_gemm_row_sharded_fused_host_enqueue): same kernel on the same inputs, so this
                # rank's downloaded panel must equal its panel of the gathered cube bit for bit; three products through the two
                # alternating device slots and host buffers, then one whole-cube download
                # with WGB_SHARD_B_UPLOAD=1 (experimental sliced upload of B + NCCL all-gather) the path needs a communicator
                split_b = os.environ.get("WGB_SHARD_B_UPLOAD", "0") not in ("", "0")
                if split_b:
                    sharded.init_comm(dev, dist, rank, world)
                ha, hb = np.ascontiguousarray(O.bf16_bits(a_blk)), np.ascontiguousarray(O.bf16_bits(b_all))
                outs = [np.zeros(m_local * N, np.uint16) for _ in range(3)]
                for o_ in outs:
                    op.enqueue_host_fused(dev, group, m_local, N, K, o_, ha, hb, var)
                cube = np.zeros(m_local * N * world, np.uint16)
                op.enqueue_host_fused(dev, group, m_local, N, K, cube, ha, hb, var, download_all=True)
                dev.poll_wait()
                mine = got[rank * m_local * N:(rank + 1) * m_local * N]
                host_ok = all(np.array_equal(o_, mine) for o_ in outs) and np.array_equal(cube, got)
                results["host_ok"] = host_ok
                if split_b:
                    w.lib().wgb_comm_destroy(dev._h)
            dist.barrier()
            if group is None:
                w.lib().wgb_comm_destroy(dev._h)
            else:
                group.close()
        # the two paths may pick different split-K / chunk shapes, so they agree to rounding, not bitwise
        fa, fb = O.bf16_from_bits(results["nccl"]).astype(np.float64), O.bf16_from_bits(results["fused"]).astype(np.float64)
        same = bool(np.max(np.abs(fa - fb) / np.abs(fb)) < 1e-2)
        full = sharded.panels_to_matrix(fb, m_local, N, world)
        full_nccl = sharded.panels_to_matrix(fa, m_local, N, world)
        # float64 reference on sampled rows of the *global* product (every rank checks rows owned by every rank)
        rows = np.array(sorted({0, 1, m_local - 1, m_local % M, M // 2, M - 1, (7 * m_local + 13) % M}))
        if tr:
            a_rows = np.stack([O.to_bf16_rne(O.uniform(O.SEED_BASE + 1, K, 1, col0=int(r))) for r in rows]).astype(np.float64)
        else:
            a_rows = np.stack([O.to_bf16_rne(O.uniform(O.SEED_BASE + 1, 1, K, row0=int(r))) for r in rows]).astype(np.float64)
        ref = a_rows @ b_all.reshape(N, K).T.astype(np.float64)
        err = max(float(np.max(np.abs(full[rows] - ref) / np.abs(ref))), float(np.max(np.abs(full_nccl[rows] - ref) / np.abs(ref))))
        good = same and err < 1e-2 and results.get("host_ok", False)
        ok &= good
        print(f"[rank {rank}] {M}x{N}x{K} tr={int(tr)} world={world}: nccl~fused {same}, host-enqueue == device path {results.get('host_ok')}, rel err vs f64 {err:.3e} -> {'OK' if good else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED CHECK", "ALL OK" if int(flag.item()) == 1 else "FAILED", flush=True)
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
