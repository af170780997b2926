"""NVLink micro-benchmark across the GPUs of one box (run under torch.distributed.run): every rank streams a local buffer into
the other ranks' memory at the same time, with per-lane 16-byte stores, TMA bulk stores, and NVSwitch multicast stores
(multimem.st to a multicast mapping).  Memory comes from torch's symmetric-memory allocator (plumbing: it exchanges the
handles and binds the multicast object); the kernels are wgb_debug_link_stream in libwgebra_b200.so.
Sizes the fused all-gather: which store form moves a rank's [M/P x N] panel to the 7 peers fastest."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    import wgmath_b200 as w
    from wgmath_b200._lib import check, lib

    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = w.GpuInstance.new(local_rank).device()
    L = lib()
    nbytes = 256 << 20
    t = symm_mem.empty(nbytes, dtype=torch.uint8, device=f"cuda:{local_rank}")
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    mc = int(hdl.multicast_ptr or 0)
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    if rank == 0:
        print(f"LINK symmetric memory: world {world}, multicast ptr {mc:#x} (0 = no multicast support), "
              f"buffer ptrs {[hex(p) for p in ptrs]}", flush=True)
    src = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{local_rank}")
    src.copy_(torch.arange(nbytes, device=src.device, dtype=torch.int32).to(torch.uint8) + rank)
    torch.cuda.synchronize()

    def run(name, mode, dsts, payload_dsts, iters=5):
        arr = (ctypes.c_void_p * len(dsts))(*dsts)
        ms = ctypes.c_float()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        check(L.wgb_debug_link_stream(dev._h, mode, arr, len(dsts), ctypes.c_void_p(src.data_ptr()), nbytes, 0, iters, ctypes.byref(ms)))
        tt = torch.tensor([ms.value], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        per = float(tt.item())
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            print(f"LINK P={world} {name:52s}: {per:8.3f} ms per 256 MiB pass  -> {nbytes * payload_dsts / per / 1e6:8.1f} GB/s delivered per GPU "
                  f"({nbytes / per / 1e6:7.1f} GB/s of source)", flush=True)

    peers = [ptrs[(rank + 1 + i) % world] for i in range(world - 1)]
    for rep in range(2):
        run("local copy (st.v4 to own HBM)", 0, [ptrs[rank]], 1)
        run("st.v4 to ONE peer", 0, peers[:1], 1)
        run(f"st.v4 to all {world - 1} peers", 0, peers, world - 1)
        run("TMA bulk (16 KiB) to ONE peer", 2, peers[:1], 1)
        run(f"TMA bulk (16 KiB) to all {world - 1} peers", 2, peers, world - 1)
        if mc:
            run(f"multimem.st to the multicast address ({world} GPUs)", 1, [mc], world)
    if mc:
        # correctness of the multicast path: rank 0 alone stores its pattern, every rank must find it in its own buffer
        t.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            arr = (ctypes.c_void_p * 1)(mc)
            ms = ctypes.c_float()
            check(L.wgb_debug_link_stream(dev._h, 1, arr, 1, ctypes.c_void_p(src.data_ptr()), nbytes, 0, 1, ctypes.byref(ms)))
        torch.cuda.synchronize()
        dist.barrier()
        want = (torch.arange(nbytes, device=src.device, dtype=torch.int32).to(torch.uint8) + 0)
        okk = bool(torch.equal(t, want))
        flag = torch.tensor([1 if okk else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"LINK multicast store from rank 0 arrived intact in every rank's buffer: {bool(flag.item())}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
