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
    """The gathered output buffers of every rank, mapped into every other rank, plus the flag block the fused GEMM + all-gather
    kernel signals through (wgb_peer_gather_*).  `tensor` is the gathered cube [m_local, N, P] of the most recent call.

    depth = number of gathered buffers the calls rotate through (include/wgb200.h: 1 = lock step, 2 = peers may run one step
    ahead, 3 = additionally the wait may trail the GEMMs by one call).  One process per GPU: handles travel through `dist`
    (torch.distributed, any backend).  `PeerGather.local_group` builds all ranks inside one process instead."""

    def __init__(self, device, dist, rank: int, world: int, m_local: int, N: int, dtype: str = "bf16", depth: int = 1, _connect=True,
                 symmetric: bool = False):
        self._device, self.rank, self.world, self.depth = device, rank, world, depth
        self._dist = dist if (_connect and world > 1) else None
        self._shape, self._dtype = (m_local, N, world), dtype
        self._nbytes = world * m_local * N * (2 if dtype == "bf16" else 4)
        self.multicast = False
        h = ctypes.c_void_p()
        if symmetric:
            # Symmetric allocation + NVSwitch multicast mapping through torch's symmetric-memory allocator (plumbing: it creates
            # the cuMem allocations, exchanges the handles between the processes and binds the multicast object); the library
            # only receives pointers (wgb_peer_gather_create_external).  With a multicast mapping the fused epilogue stores each
            # output block once and the switch replicates it.
            import torch
            import torch.distributed._symmetric_memory as symm_mem
            region = lib().wgb_peer_gather_region_bytes(self._nbytes, depth)
            t = symm_mem.empty(region, dtype=torch.uint8, device=f"cuda:{device.ordinal}")
            hdl = symm_mem.rendezvous(t, dist.group.WORLD)
            bases = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
            mc = int(hdl.multicast_ptr or 0)
            check(lib().wgb_peer_gather_create_external(device._h, world, rank, self._nbytes, depth, bases, ctypes.c_void_p(mc or None),
                                                        ctypes.byref(h)))
            self._h, self._symm, self.multicast = h, (t, hdl), mc != 0
            torch.cuda.synchronize()
            dist.barrier()          # every rank's flag block is cleared before anyone's first product
            return
        check(lib().wgb_peer_gather_create_ex(device._h, world, rank, self._nbytes, depth, ctypes.byref(h)))
        self._h = h
        if not _connect:
            return
        mine = (ctypes.c_char * IPC_HANDLE_BYTES)()
        check(lib().wgb_peer_gather_export(h, mine))
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, bytes(mine))
        else:
            handles = [bytes(mine)]
        blob = (ctypes.c_char * (IPC_HANDLE_BYTES * world)).from_buffer_copy(b"".join(handles))
        check(lib().wgb_peer_gather_connect(h, blob))

    @classmethod
    def local_group(cls, devices, m_local: int, N: int, dtype: str = "bf16", depth: int = 1) -> "List[PeerGather]":
        """All ranks in this process, one device handle (context) per rank — wgb_peer_gather_connect_local.  The devices may be
        one and the same GPU: the complete protocol (peer stores, ready / done flags, epochs) then runs on a single-GPU box."""
        world = len(devices)
        groups = [cls(dev, None, r, world, m_local, N, dtype, depth, _connect=False) for r, dev in enumerate(devices)]
        arr = (ctypes.c_void_p * world)(*[g._h for g in groups])
        for g in groups:
            check(lib().wgb_peer_gather_connect_local(g._h, arr))
        return groups

    def tensor_at(self, calls_back: int = 0) -> GpuTensor:
        """The gathered cube written by the call `calls_back` calls ago (while the group still keeps it)."""
        bh = ctypes.c_void_p()
        check(lib().wgb_peer_gather_buffer_at(self._h, calls_back, ctypes.byref(bh)))
        # the group owns the wgb_buffer: never destroyed from the Python side
        return GpuTensor(self._shape, _BorrowedBuffer(self._device, bh, self._nbytes), self._dtype)

    @property
    def tensor(self) -> GpuTensor:
        return self.tensor_at(0)

    def wait(self, pass_, calls_back: int = 0) -> None:
        """Queue the wait for every peer's panel of the call `calls_back` calls ago (wgb_peer_gather_wait)."""
        check(lib().wgb_peer_gather_wait(pass_._h, self._h, calls_back))

    def debug_flags(self) -> dict:
        """Snapshot of this rank's flag block (wgb_peer_gather_debug_flags): readable while the queues are busy or stuck."""
        out = (ctypes.c_uint32 * 18)()
        check(lib().wgb_peer_gather_debug_flags(self._h, out))
        v = list(out)
        return {"ready": v[:self.world], "done": v[8:8 + self.world], "cta_counter": v[16], "calls": v[17]}

    def close(self, collective: bool = True):
        """Collective across the ranks of a multi-process group: everyone unmaps its peers, meets at a barrier, then frees
        (CUDA IPC requires the importers to let go before the exporter frees)."""
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().wgb_peer_gather_disconnect(h)
            if collective and self._dist is not None:
                self._dist.barrier()
            lib().wgb_peer_gather_destroy(h)
            self._symm = None

    def __del__(self):
        try:
            self.close(collective=False)
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
                       f32_mode=F32Mode.Auto, wait: bool = True) -> None:
        """Same result in group.tensor, the all-gather fused into the GEMM epilogue (peer stores over NVLink).
        wait=False (WGB_GATHER_NO_WAIT): only this rank's GEMM + peer stores are queued; call group.wait(pass_, calls_back)
        before the gathered cube is read."""
        m1, b = as_view(m1_local, 3), as_view(m2, 3)
        s1, s2 = shapes.get(device, m1.shape()).to_c(), shapes.get(device, b.shape()).to_c()
        check(lib().wgb_gemm_row_sharded_fused_ex(pass_._h, int(variant), group._h, m1.buffer()._h, ctypes.byref(s1), b.buffer()._h,
                                                  ctypes.byref(s2), _DTYPE_CODE[m1.dtype], _DTYPE_CODE[group._dtype], int(f32_mode),
                                                  0 if wait else 1))


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
                                                            _DTYPE_CODE[group._dtype], int(f32_mode), 1 if download_all else 0))
