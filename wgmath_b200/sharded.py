"""Row-sharded GEMM across the GPUs of one box (BASELINE configs[4]).

One process per GPU.  Rank p owns rows [p*M/P, (p+1)*M/P) of m1 and of the product; m2 is
replicated.  Each rank computes its block into a contiguous column-major [M/P x N] panel and the
panels are exchanged over NVLink (chunked by column panel, overlapped with the remaining tiles);
the gathered result is a cube view size=[M/P, N, P], stride=M/P, stride_mat=(M/P)*N.

`torch.distributed` is used only as plumbing (rendezvous, distributing the communicator id,
barriers); the data path is wgb_gemm_row_sharded in libwgebra_b200.so."""
from __future__ import annotations

import ctypes
from typing import List, Tuple

import numpy as np

from ._lib import BF16, COMM_ID_BYTES, F32, IPC_HANDLE_BYTES, check, lib
from .linalg import F32Mode, GemmVariant
from .shapes import ViewShape
from .tensor import Buffer, GpuTensor, GpuTensorView, as_view

_DTYPE_CODE = {"f32": F32, "bf16": BF16}


def row_partition(M: int, P: int) -> List[Tuple[int, int]]:
    """(first_row, nrows) per rank.  Equal blocks: the gathered panels must all have the same size."""
    if P < 1 or M % P != 0:
        raise ValueError(f"row-sharding needs M ({M}) divisible by the number of ranks ({P})")
    b = M // P
    return [(p * b, b) for p in range(P)]


def column_chunks(N: int, n_chunks: int, P: int) -> List[Tuple[int, int]]:
    """(first_col, ncols) of each exchange chunk — same rule as wgb_gemm_row_sharded (comm.cu):
    whole 256-column tiles, default 8 chunks when there is an exchange, 1 otherwise."""
    if N == 0:
        return []
    nch = n_chunks if n_chunks > 0 else (8 if P > 1 else 1)
    width = -(-N // nch)
    width = (width + 255) & ~255
    return [(n0, min(width, N - n0)) for n0 in range(0, N, width)]


def gathered_view(out: GpuTensor, m_local: int, N: int, P: int) -> GpuTensorView:
    """The gathered result as the reference's GpuCubeView (tensor.rs:465-481): matrix p = rows of rank p."""
    return GpuTensorView(ViewShape((m_local, N, P), m_local, m_local * N, 0), out.buffer(), out.dtype, 3)


def panels_to_matrix(flat: np.ndarray, m_local: int, N: int, P: int) -> np.ndarray:
    """Host-side: [P][N][m_local] gathered panels -> the full (P*m_local) x N matrix."""
    return np.concatenate([flat[p * m_local * N:(p + 1) * m_local * N].reshape(N, m_local).T for p in range(P)], axis=0)


def init_comm(device, dist=None, rank: int = 0, world: int = 1) -> None:
    """Create the NCCL communicator for this context: rank 0 makes the id, everyone gets it through
    torch.distributed (any backend), then wgb_comm_init_rank."""
    idbuf = (ctypes.c_char * COMM_ID_BYTES)()
    if rank == 0:
        check(lib().wgb_comm_get_unique_id(idbuf))
    payload = [bytes(idbuf)]
    if dist is not None and world > 1:
        dist.broadcast_object_list(payload, src=0)
    idbuf = (ctypes.c_char * COMM_ID_BYTES).from_buffer_copy(payload[0])
    check(lib().wgb_comm_init_rank(device._h, world, rank, idbuf))


class PeerGather:
    """The gathered output buffer of every rank, mapped into every other rank through CUDA IPC, plus the flag block
    the fused GEMM + all-gather kernel signals through (wgb_peer_gather_*).  `tensor` is the local gathered cube
    [m_local, N, P] as a GpuTensor."""

    def __init__(self, device, dist, rank: int, world: int, m_local: int, N: int, dtype: str = "bf16"):
        self._device, self.rank, self.world = device, rank, world
        nbytes = world * m_local * N * (2 if dtype == "bf16" else 4)
        h = ctypes.c_void_p()
        check(lib().wgb_peer_gather_create(device._h, world, rank, nbytes, ctypes.byref(h)))
        self._h = h
        mine = (ctypes.c_char * IPC_HANDLE_BYTES)()
        check(lib().wgb_peer_gather_export(h, mine))
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, bytes(mine))
        else:
            handles = [bytes(mine)]
        blob = (ctypes.c_char * (IPC_HANDLE_BYTES * world)).from_buffer_copy(b"".join(handles))
        check(lib().wgb_peer_gather_connect(h, blob))
        bh = ctypes.c_void_p()
        check(lib().wgb_peer_gather_buffer(h, ctypes.byref(bh)))
        # the group owns the wgb_buffer: never destroyed from the Python side
        self.tensor = GpuTensor((m_local, N, world), _BorrowedBuffer(device, bh, nbytes), dtype)

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().wgb_peer_gather_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _BorrowedBuffer(Buffer):
    """A wgb_buffer owned by the library (wgb_peer_gather_buffer): never destroyed from Python."""

    def __del__(self):
        self._h = None


class RowShardedGemm:
    """out_gathered[p] = (m1 * m2)[rows of rank p]  for every p, on every rank."""

    def __init__(self, device):
        self._device = device

    def dispatch(self, device, shapes, pass_, out_gathered: GpuTensor, m1_local, m2, variant=GemmVariant.Gemm,
                 f32_mode=F32Mode.Auto, n_chunks: int = 0) -> None:
        m1, b = as_view(m1_local, 3), as_view(m2, 3)
        s1, s2 = shapes.get(device, m1.shape()).to_c(), shapes.get(device, b.shape()).to_c()
        check(lib().wgb_gemm_row_sharded(pass_._h, int(variant), out_gathered.buffer()._h, m1.buffer()._h, ctypes.byref(s1),
                                         b.buffer()._h, ctypes.byref(s2), _DTYPE_CODE[m1.dtype], _DTYPE_CODE[out_gathered.dtype],
                                         int(f32_mode), n_chunks))

    def dispatch_fused(self, device, shapes, pass_, group: "PeerGather", m1_local, m2, variant=GemmVariant.Gemm,
                       f32_mode=F32Mode.Auto) -> None:
        """Same result in group.tensor, the all-gather fused into the GEMM epilogue (peer stores over NVLink)."""
        m1, b = as_view(m1_local, 3), as_view(m2, 3)
        s1, s2 = shapes.get(device, m1.shape()).to_c(), shapes.get(device, b.shape()).to_c()
        check(lib().wgb_gemm_row_sharded_fused(pass_._h, int(variant), group._h, m1.buffer()._h, ctypes.byref(s1), b.buffer()._h,
                                               ctypes.byref(s2), _DTYPE_CODE[m1.dtype], _DTYPE_CODE[group.tensor.dtype], int(f32_mode)))


    def enqueue_host_fused(self, device, group: "PeerGather", m_local: int, N: int, K: int, out_host, m1_local_host, m2_host,
                           variant=GemmVariant.Gemm, in_dtype: str = "bf16", f32_mode=F32Mode.Auto, download_all: bool = False) -> None:
        """Host operands in, host result out, enqueued (wgb_gemm_row_sharded_fused_host_enqueue): upload of the next product
        overlaps the GEMM + gather and the download of this one.  `*_host` are numpy arrays or raw host pointers; `out_host`
        receives this rank's [m_local x N] panel, or the whole gathered cube with download_all.  Close a batch with
        Gemm.flush_host(device) / device.poll_wait()."""
        def ptr(x):
            return x if isinstance(x, ctypes.c_void_p) else x.ctypes.data_as(ctypes.c_void_p)
        check(lib().wgb_gemm_row_sharded_fused_host_enqueue(device._h, int(variant), group._h, m_local, N, K, ptr(out_host),
                                                            ptr(m1_local_host), ptr(m2_host), _DTYPE_CODE[in_dtype],
                                                            _DTYPE_CODE[group.tensor.dtype], int(f32_mode), 1 if download_all else 0))


def bench_row_sharded(w, O, gpu, shapes, dist, ngpu, rank, args, timed, peaks):
    """bench.py's N > 1 leg: bf16 (4096*P)^3, row-sharded, all-gather of C; returns the JSON fields."""
    dev = gpu.device()
    n = 4096 * ngpu
    m_local = n // ngpu
    ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST
    import os
    mode = os.environ.get("WGB_SHARD_MODE", "fused")                     # "fused" (peer stores) or "nccl"
    a = w.TensorBuilder.matrix(m_local, n, ST).build(dev, "bf16")       # my row block of A
    b = w.TensorBuilder.matrix(n, n, ST).build(dev, "bf16")             # B replicated
    group = None
    if mode == "fused":
        group = PeerGather(dev, dist, rank, ngpu, m_local, n, "bf16")
        c = group.tensor
    else:
        init_comm(dev, dist, rank, ngpu)
        c = w.TensorBuilder.tensor((m_local, n, ngpu), ST).build(dev, "bf16")
    enc = dev.create_command_encoder()
    with enc.compute_pass("init", None) as p:
        w.fill_uniform(dev, p, a, O.SEED_BASE + 1, row0=rank * m_local)  # element (i, j) independent of the sharding
        w.fill_uniform(dev, p, b, O.SEED_BASE + 2)
    dev.poll_wait()
    op = RowShardedGemm(dev)
    if group is not None:
        step_fn = lambda p, i: op.dispatch_fused(dev, shapes, p, group, a, b)    # noqa: E731
    else:
        step_fn = lambda p, i: op.dispatch(dev, shapes, p, c, a, b)             # noqa: E731
    sec, launches = timed(step_fn, args.steps, args.warmup)
    flops = 2.0 * n * n * n
    value = flops * args.steps / sec / 1e12
    ms_step = sec * 1e3 / args.steps
    long_run = sec > 1.0
    peak = (peaks["bf16_tflops_sustained"] if long_run else peaks["bf16_tflops"]) * ngpu
    comm_bytes = (ngpu - 1) * m_local * n * 2
    kname = ("gemm_tc<bf16> (tcgen05) with the all-gather of C fused into the epilogue (peer stores over NVLink)" if group is not None
             else "gemm_tc<bf16> (tcgen05) + chunked all-gather of C (NCCL send/recv over NVLink)")
    roof = {"bound": "tensor", "kernel": kname,
            "achieved": value, "peak": peak, "unit": "TFLOP/s", "frac": value / peak,
            "peak_source": f"{peaks['source']} ({'sustained' if long_run else 'burst'}) x {ngpu} GPUs", "traffic": None,
            "algorithmic": "2*M*N*K flop per step over all ranks", "nvlink_bytes_in_per_gpu_per_step": comm_bytes,
            "nvlink_floor_ms": comm_bytes / 770e9 * 1e3}
    # e2e: HOST buffers in, HOST buffer out, through the C ABI, every step: this rank's A block and B go up, the sharded GEMM +
    # gather runs, the result comes down.
    # (a) fused mode: wgb_gemm_row_sharded_fused_host_enqueue — products queued back to back, upload of product i + 1 under the
    #     GEMM / download of product i; each rank downloads its own [M/P x N] panel, so the ranks of the box assemble C in host
    #     memory and every byte of C crosses a host link once
    # (b) separate blocking calls (write A, write B, dispatch, read the whole gathered cube on every rank) — the sequence of the
    #     reference's tests, kept as `separate_calls`
    L = lib()
    abytes, bbytes, cbytes = m_local * n * 2, n * n * 2, n * n * 2
    pbytes = m_local * n * 2
    ha, hb, hc, hp0, hp1 = (ctypes.c_void_p() for _ in range(5))
    for h, nb in ((ha, abytes), (hb, bbytes), (hc, cbytes), (hp0, pbytes), (hp1, pbytes)):
        check(L.wgb_host_alloc(nb, ctypes.byref(h)))
    check(L.wgb_buffer_read(dev._h, a.buffer()._h, 0, ha, abytes))
    check(L.wgb_buffer_read(dev._h, b.buffer()._h, 0, hb, bbytes))
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_seq_step(p, i):
        check(L.wgb_buffer_write(dev._h, a.buffer()._h, 0, ha, abytes))
        check(L.wgb_buffer_write(dev._h, b.buffer()._h, 0, hb, bbytes))
        step_fn(p, i)
        check(L.wgb_buffer_read(dev._h, c.buffer()._h, 0, hc, cbytes))
    seq_sec, _ = timed(e2e_seq_step, e2e_steps, 1)
    seq = {"value": flops * e2e_steps / seq_sec / 1e12, "ms_per_step": seq_sec * 1e3 / e2e_steps,
           "h2d_bytes_per_step": (abytes + bbytes) * ngpu, "d2h_bytes_per_step": cbytes * ngpu,
           "call": "wgb_buffer_write x2 + sharded dispatch + wgb_buffer_read of the whole gathered cube on every rank"}
    if group is not None:
        # experimental (off by default, see comm.cu): every rank uploads 1/P of B, the slices are all-gathered over NVLink
        split_b = os.environ.get("WGB_SHARD_B_UPLOAD", "0") not in ("", "0")
        if split_b:
            init_comm(dev, dist, rank, ngpu)

        def e2e_step(p, i):
            op.enqueue_host_fused(dev, group, m_local, n, n, hp0 if i % 2 == 0 else hp1, ha, hb)
        e2e_sec, _ = timed(e2e_step, e2e_steps, 2, before_end=lambda: check(L.wgb_gemm_host_flush(dev._h)))
        e2e = {"value": flops * e2e_steps / e2e_sec / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": abytes * ngpu + (bbytes if split_b else bbytes * ngpu),
               "d2h_bytes_per_step": pbytes * ngpu, "steps": e2e_steps, "ms_per_step": e2e_sec * 1e3 / e2e_steps,
               "b_upload": "1/P slice per rank + NCCL all-gather over NVLink (WGB_SHARD_B_UPLOAD=1)" if split_b else "whole B on every rank",
               "call": "wgb_gemm_row_sharded_fused_host_enqueue per step on every rank (pinned host buffers; A block + B up, fused "
                       "GEMM + all-gather, this rank's panel of C down: the box's host memory ends with all of C), closed by "
                       "wgb_gemm_host_flush", "separate_calls": seq}
    else:
        e2e = {"value": seq["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": seq["h2d_bytes_per_step"],
               "d2h_bytes_per_step": seq["d2h_bytes_per_step"], "steps": e2e_steps, "ms_per_step": seq["ms_per_step"], "call": seq["call"]}
    for h in (ha, hb, hc, hp0, hp1):
        L.wgb_host_free(h)
    if group is not None and split_b:
        dev.poll_wait()
        lib().wgb_comm_destroy(dev._h)
    if dist is not None:
        dist.barrier()
    if group is not None:
        group.close()
    else:
        lib().wgb_comm_destroy(dev._h)
    return value, ms_step, launches, roof, e2e, "bf16"
