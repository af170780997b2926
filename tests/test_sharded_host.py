"""CPU tests (-m "not gpu") for the N > 1 path's host logic: the row partition, the exchange-chunk
plan (same rule as comm.cu), the gathered-panel layout (a reference GpuCube view) and the
communicator-id distribution, run as a real world_size-2 job over the gloo backend."""
import os
import socket
import sys

import numpy as np
import pytest

from wgmath_b200 import sharded
from wgmath_b200._lib import COMM_ID_BYTES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_partition():
    assert sharded.row_partition(32768, 8) == [(p * 4096, 4096) for p in range(8)]
    assert sharded.row_partition(4096, 1) == [(0, 4096)]
    with pytest.raises(ValueError):
        sharded.row_partition(100, 8)


@pytest.mark.parametrize("N,nch,P", [(32768, 0, 8), (4096, 0, 1), (8192, 4, 2), (1000, 3, 2), (256, 8, 8), (257, 0, 4)])
def test_column_chunks_cover_exactly_in_whole_tiles(N, nch, P):
    ch = sharded.column_chunks(N, nch, P)
    assert ch[0][0] == 0 and sum(c for _, c in ch) == N
    for (a, ca), (b, _) in zip(ch, ch[1:]):
        assert a + ca == b and ca % 256 == 0            # only the last chunk may be ragged
    if P == 1 and nch == 0:
        assert len(ch) == 1


def test_gathered_view_is_a_reference_cube_view():
    import wgmath_b200 as w
    t = w.GpuTensor((512, 64, 4), buffer=None, dtype="bf16")
    v = sharded.gathered_view(t, 512, 64, 4)
    s = v.shape()
    assert (s.size, s.stride, s.stride_mat, s.offset) == ((512, 64, 4), 512, 512 * 64, 0)
    m2 = v.matrix(2).shape()                                 # tensor.rs:466-481: rank 2's rows
    assert (m2.size, m2.offset) == ((512, 64, 1), 2 * 512 * 64)


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        from oracle import oracle as O
        dist.init_process_group("gloo", rank=rank, world_size=world)
        M, N, K = 64 * world, 48, 32
        (row0, m_local), = [sharded.row_partition(M, world)[rank]]
        # every rank builds its shard from the seeded generator: element (i, j) is independent of the sharding
        a_blk = O.uniform(O.SEED_BASE + 1, m_local, K, row0=row0).reshape(K, m_local).T.astype(np.float64)
        b = O.uniform(O.SEED_BASE + 2, K, N).reshape(N, K).T.astype(np.float64)
        panel = np.ascontiguousarray((a_blk @ b).T.reshape(-1))            # column-major [m_local x N]
        gathered = np.zeros(world * m_local * N)
        # the exchange plan of wgb_gemm_row_sharded, chunk by chunk, emulated with gloo
        for n0, nc in sharded.column_chunks(N, 2, world):
            lo, hi = n0 * m_local, (n0 + nc) * m_local
            parts = [torch.zeros(hi - lo, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(panel[lo:hi].copy()))
            for p in range(world):
                gathered[p * m_local * N + lo: p * m_local * N + hi] = parts[p].numpy()
        full = sharded.panels_to_matrix(gathered, m_local, N, world)
        a_full = O.uniform(O.SEED_BASE + 1, M, K).reshape(K, M).T.astype(np.float64)
        ok = np.allclose(full, a_full @ b, rtol=1e-12)
        # id distribution: rank 0's 128-byte payload reaches everyone
        payload = [bytes(range(128)) if rank == 0 else b""]
        dist.broadcast_object_list(payload, src=0)
        ok = ok and payload[0] == bytes(range(128)) and len(payload[0]) == COMM_ID_BYTES
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok)))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))


def test_world_size_2_gloo_exchange_plan():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: True, 1: True}, results
