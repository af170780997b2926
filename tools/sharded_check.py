"""Multi-GPU parity check for the row-sharded GEMM (run under torch.distributed.run, one rank per GPU).

For every case: the NCCL exchange path and the fused peer-store path in each of its forms — epilogue store form (TMA bulk
stores / per-lane stores) x gathered-buffer depth (1 lock step, 2 look-ahead, 3 with the wait deferred by one call) — over
several steps with different operands per step, every step's gathered cube snapshotted on the queue and compared afterwards:
  * all fused forms agree bit for bit (same MMA plan, only the way the results leave the SM differs),
  * NCCL path ~ fused path (the NCCL path leaves SMs to the collective, so its tile plan may differ: 1e-2, not bitwise),
  * sampled rows of the GLOBAL product against float64 on the same bf16-rounded inputs (1e-2, bf16 output),
  * the host-operand enqueue form (wgb_gemm_row_sharded_fused_host_enqueue), with B uploaded whole on every rank and with
    the 1/P-slice upload + NVLink all-gather of B, equals the device path bit for bit; a large product followed by a small one
    goes through the alternating operand slots without corrupting the product in flight.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, WGB_TC_EPI, depth, deferred wait, symmetric memory + multicast mapping)
FUSED_FORMS = (("tma-d1", 1, 1, False, False), ("stg-d1", 0, 1, False, False), ("tma-d2", 1, 2, False, False),
               ("tma-d3-deferred", 1, 3, True, False), ("stg-d3-deferred", 0, 3, True, False),
               ("mc-d1", 2, 1, False, True), ("mc-d3-deferred", 2, 3, True, True), ("tma-d2-symmetric", 1, 2, False, True))


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
    quick = "quick" in sys.argv
    cases = ((1024, 2048, 1024, False), (512, 1000, 520, True), (2048, 4096, 4096, False))
    if quick:
        cases = cases[:2]
    steps = 4
    ok = True
    op = sharded.RowShardedGemm(dev)
    for (m_local, N, K, tr) in cases:
        M = m_local * world
        var = w.GemmVariant.GemmTr if tr else w.GemmVariant.Gemm
        panel = m_local * N

        def a_block(step):
            seed = O.SEED_BASE + 1 + 16 * step
            return O.to_bf16_rne(O.uniform(seed, K, m_local, col0=rank * m_local) if tr else O.uniform(seed, m_local, K, row0=rank * m_local))
        a_h = [a_block(s) for s in range(steps)]
        b_all = O.to_bf16_rne(O.uniform(O.SEED_BASE + 2, K, N))
        ta = [w.TensorBuilder.matrix(*((K, m_local) if tr else (m_local, K)), ST).build_init(dev, O.bf16_bits(a), "bf16") for a in a_h]
        tb = w.TensorBuilder.matrix(K, N, ST).build_init(dev, O.bf16_bits(b_all), "bf16")
        results = {}
        # ---- NCCL exchange (plain GEMM chunks + grouped send / recv)
        sharded.init_comm(dev, dist, rank, world)
        c = w.TensorBuilder.tensor((m_local, N, world), ST).build_init(dev, np.zeros(panel * world, np.uint16), "bf16")
        cubes = []
        for s in range(steps):
            enc = dev.create_command_encoder()
            with enc.compute_pass("sharded", None) as p:
                op.dispatch(dev, shapes, p, c, ta[s], tb, var, n_chunks=3)
            gpu.queue().submit(enc.finish())
            cubes.append(c.read())
        results["nccl"] = cubes
        dist.barrier()
        w.lib().wgb_comm_destroy(dev._h)
        # ---- fused forms
        for name, epi, depth, deferred, symmetric in FUSED_FORMS:
            os.environ["WGB_TC_EPI"] = str(epi)
            group = sharded.PeerGather(dev, dist, rank, world, m_local, N, "bf16", depth=depth, symmetric=symmetric)
            if symmetric and epi == 2 and not group.multicast:
                if rank == 0:
                    print(f"(no multicast mapping on this box: form {name} falls back to TMA stores)", flush=True)
            snaps = [w.TensorBuilder.tensor((m_local, N, world), ST).build(dev, "bf16") for _ in range(steps)]
            enc = dev.create_command_encoder()
            with enc.compute_pass("fused", None) as p:
                for s in range(steps):
                    op.dispatch_fused(dev, shapes, p, group, ta[s], tb, var, wait=not deferred)
                    cfg = p.last_gemm_config()
                    if not deferred:
                        snaps[s].copy_from(None, group.tensor_at(0))
                    elif s > 0:
                        group.wait(p, 1)
                        snaps[s - 1].copy_from(None, group.tensor_at(1))
                if deferred:
                    group.wait(p, 0)
                    snaps[steps - 1].copy_from(None, group.tensor_at(0))
            gpu.queue().submit(enc.finish())
            results[name] = [t.read() for t in snaps]
            results[name + ":cfg"] = cfg
            dist.barrier()
            if name == "tma-d1":
                # ---- host-operand enqueue form on this group: same kernel on the same inputs => bit-identical panels
                for split_b in (0, 1):
                    os.environ["WGB_SHARD_B_UPLOAD"] = str(split_b)
                    if split_b:
                        sharded.init_comm(dev, dist, rank, world)
                    ha = [np.ascontiguousarray(O.bf16_bits(a)) for a in a_h]
                    hb = np.ascontiguousarray(O.bf16_bits(b_all))
                    outs = [np.zeros(panel, np.uint16) for _ in range(steps)]
                    for s in range(steps):
                        op.enqueue_host_fused(dev, group, m_local, N, K, outs[s], ha[s], hb, var)
                    cube = np.zeros(panel * world, np.uint16)
                    op.enqueue_host_fused(dev, group, m_local, N, K, cube, ha[steps - 1], hb, var, download_all=True)
                    # a smaller product right behind the larger ones: the operand slots must not move under a product in flight
                    ms, ns, ks = m_local // 2, N // 2 // world * world, K // 2 // 8 * 8
                    small_a = O.to_bf16_rne(O.uniform(O.SEED_BASE + 77, ks, ms, col0=rank * ms) if tr else O.uniform(O.SEED_BASE + 77, ms, ks, row0=rank * ms))
                    small_b = O.to_bf16_rne(O.uniform(O.SEED_BASE + 78, ks, ns))
                    small_out = np.zeros(ms * ns, np.uint16)
                    hsa, hsb = np.ascontiguousarray(O.bf16_bits(small_a)), np.ascontiguousarray(O.bf16_bits(small_b))
                    op.enqueue_host_fused(dev, group, ms, ns, ks, small_out, hsa, hsb, var)
                    dev.poll_wait()
                    sa64 = (small_a.reshape(ms, ks) if tr else small_a.reshape(ks, ms).T).astype(np.float64)   # [m][k]
                    sref = sa64[:4] @ small_b.reshape(ns, ks).T.astype(np.float64)
                    sgot = O.bf16_from_bits(small_out).reshape(ns, ms).T.astype(np.float64)[:4]
                    small_ok = float(np.max(np.abs(sgot - sref) / np.abs(sref))) < 1e-2
                    mine = [results[name][s][rank * panel:(rank + 1) * panel] for s in range(steps)]
                    host_ok = all(np.array_equal(outs[s], mine[s]) for s in range(steps)) and np.array_equal(cube, results[name][steps - 1])
                    results[f"host_ok_splitb{split_b}"] = host_ok and small_ok
                    dist.barrier()
                    if split_b:
                        w.lib().wgb_comm_destroy(dev._h)
                os.environ.pop("WGB_SHARD_B_UPLOAD", None)
            group.close()
        os.environ.pop("WGB_TC_EPI", None)
        base = results["tma-d1"]
        forms_equal = all(np.array_equal(results[name][s], base[s]) for name, *_ in FUSED_FORMS for s in range(steps))
        epi_ran = results["tma-d1:cfg"]["epi_tma"] == 1 and results["stg-d1:cfg"]["epi_tma"] == 0 and results["tma-d1:cfg"]["dests"] == world
        mc_ran = results["mc-d1:cfg"]["epi_tma"] == 2 and results["mc-d3-deferred:cfg"]["epi_tma"] == 2
        err = 0.0
        same = True
        rows = np.array(sorted({0, 1, m_local - 1, m_local % M, M // 2, M - 1, (7 * m_local + 13) % M}))
        b64 = b_all.reshape(N, K).T.astype(np.float64)
        for s in range(steps):
            fa = O.bf16_from_bits(results["nccl"][s]).astype(np.float64)
            fb = O.bf16_from_bits(base[s]).astype(np.float64)
            same &= bool(np.max(np.abs(fa - fb) / np.abs(fb)) < 1e-2)
            full = sharded.panels_to_matrix(fb, m_local, N, world)
            # float64 reference on sampled rows of the *global* product (every rank checks rows owned by every rank)
            seed = O.SEED_BASE + 1 + 16 * s
            if tr:
                a_rows = np.stack([O.to_bf16_rne(O.uniform(seed, K, 1, col0=int(r))) for r in rows]).astype(np.float64)
            else:
                a_rows = np.stack([O.to_bf16_rne(O.uniform(seed, 1, K, row0=int(r))) for r in rows]).astype(np.float64)
            ref = a_rows @ b64
            err = max(err, float(np.max(np.abs(full[rows] - ref) / np.abs(ref))))
        good = forms_equal and epi_ran and same and err < 1e-2 and results.get("host_ok_splitb0", False) and results.get("host_ok_splitb1", False)
        ok &= good
        print(f"[rank {rank}] {M}x{N}x{K} tr={int(tr)} world={world} steps={steps}: fused forms bit-identical {forms_equal} "
              f"(tma/stg epilogues ran: {epi_ran}, multicast epilogue ran: {mc_ran}), nccl~fused {same}, host-enqueue == device path (whole B / sliced B + all-gather): "
              f"{results.get('host_ok_splitb0')} / {results.get('host_ok_splitb1')}, rel err vs f64 {err:.3e} -> {'OK' if good else 'FAIL'}",
              flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED CHECK", "ALL OK" if int(flag.item()) == 1 else "FAILED", flush=True)
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
