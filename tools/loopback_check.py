"""The fused GEMM + all-gather protocol on ONE GPU: P ranks, each with its own context (queue) on device 0, their gathered
buffers connected in-process (wgb_peer_gather_connect_local).  Every piece of the multi-GPU path runs — peer stores (direct or
TMA bulk stores) into the other ranks' buffers, ready / done flags, epochs, buffer rotation (depth 1 / 2 / 3), deferred waits —
only the wires are missing.  Launched by tests/test_gpu_loopback.py with CUDA_DEVICE_MAX_CONNECTIONS=32 so that the ranks'
streams do not share a hardware queue.

The ranks share the SMs of one device and the GEMM is persistent, so the problems here are kept to a few tiles: a grid that
filled every SM while waiting for a peer's 32-thread flag kernel would wait for ever.

    python tools/loopback_check.py [quick]
"""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("WGB_PEER_TIMEOUT_MS", "20000")
# every kernel loaded up front: with lazy loading a first launch may synchronise the context, which never happens while another
# rank of this same process spins on a flag that a later launch must set
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgmath_b200 as w  # noqa: E402
from oracle import oracle as O  # noqa: E402
from wgmath_b200 import sharded  # noqa: E402

ST = w.BufferUsages.STORAGE | w.BufferUsages.COPY_SRC | w.BufferUsages.COPY_DST


def run_case(P, depth, epi, tr, m_local, N, K, steps, deferred, out_dtype="bf16", in_dtype="bf16"):
    print(f"... P={P} depth={depth} epi={epi} tr={int(tr)} {m_local}x{N}x{K} steps={steps} deferred={deferred}", flush=True)
    os.environ["WGB_TC_EPI"] = str(epi)
    gpus = [w.GpuInstance.new(0) for _ in range(P)]
    devs = [g.device() for g in gpus]
    shapes = w.ViewShapeBuffers.new()
    groups = sharded.PeerGather.local_group(devs, m_local, N, out_dtype, depth)
    op = [sharded.RowShardedGemm(d) for d in devs]
    gemm = [w.Gemm.from_device(d) for d in devs]
    var = w.GemmVariant.GemmTr if tr else w.GemmVariant.Gemm
    ar, ac = (K, m_local) if tr else (m_local, K)

    def operand(seed, rows, cols):
        x = O.uniform(seed, rows, cols)
        return O.bf16_bits(x) if in_dtype == "bf16" else x
    B_h = operand(O.SEED_BASE + 2, K, N)
    A, B, snap, plain = [], [], [], []
    for r, d in enumerate(devs):
        A.append([w.TensorBuilder.matrix(ar, ac, ST).build_init(d, operand(O.SEED_BASE + 100 * s + r, ar, ac), in_dtype) for s in range(steps)])
        B.append(w.TensorBuilder.matrix(K, N, ST).build_init(d, B_h, in_dtype))
        snap.append([w.TensorBuilder.tensor((m_local, N, P), ST).build(d, out_dtype) for _ in range(steps)])
        plain.append([w.TensorBuilder.matrix(m_local, N, ST).build(d, out_dtype) for _ in range(steps)])
    # the unfused product of every rank's block, step by step (same kernel configuration apart from the epilogue form)
    os.environ["WGB_TC_EPI"] = "0"
    for r, d in enumerate(devs):
        enc = d.create_command_encoder()
        with enc.compute_pass("plain", None) as p:
            for s in range(steps):
                gemm[r].dispatch_generic(d, shapes, p, plain[r][s], A[r][s], B[r], var)
        d.poll_wait()
    os.environ["WGB_TC_EPI"] = str(epi)
    encs = [d.create_command_encoder() for d in devs]
    passes = [e.compute_pass("loopback", None) for e in encs]
    for s in range(steps):
        for r, d in enumerate(devs):
            op[r].dispatch_fused(d, shapes, passes[r], groups[r], A[r][s], B[r], var, wait=not deferred)
            if not deferred:
                snap[r][s].copy_from(None, groups[r].tensor_at(0))
            elif s > 0:
                groups[r].wait(passes[r], 1)                       # the gather of the previous call, one GEMM later
                snap[r][s - 1].copy_from(None, groups[r].tensor_at(1))
    if deferred:
        for r in range(P):
            groups[r].wait(passes[r], 0)
            snap[r][steps - 1].copy_from(None, groups[r].tensor_at(0))
    for p in passes:
        p.end()
    if os.environ.get("LOOPBACK_DEBUG"):
        import time
        time.sleep(float(os.environ["LOOPBACK_DEBUG"]))
        for r, g in enumerate(groups):
            print(f"  [debug] rank {r} flags after {os.environ['LOOPBACK_DEBUG']} s: {g.debug_flags()}", flush=True)
    for d in devs:
        d.poll_wait()
    ok = True
    worst = 0.0
    panel = m_local * N
    ref_panels = [[plain[q][s].read() for s in range(steps)] for q in range(P)]
    for r in range(P):
        for s in range(steps):
            got = snap[r][s].read()
            for q in range(P):
                if not np.array_equal(got[q * panel:(q + 1) * panel], ref_panels[q][s]):
                    ok = False
    # one float64 check so that "identical to the plain kernel" cannot mean "identically wrong"
    a0 = O.uniform(O.SEED_BASE + 100 * (steps - 1) + (P - 1), ar, ac)
    b0 = O.uniform(O.SEED_BASE + 2, K, N)
    if in_dtype == "bf16":
        a0, b0 = O.to_bf16_rne(a0), O.to_bf16_rne(b0)
    a64 = a0.reshape(ac, ar).T.astype(np.float64)
    ref = (a64.T if tr else a64) @ b0.reshape(N, K).T.astype(np.float64)
    got = ref_panels[P - 1][steps - 1]
    got = (O.bf16_from_bits(got) if out_dtype == "bf16" else got).reshape(N, m_local).T.astype(np.float64)
    worst = float(np.max(np.abs(got - ref) / np.abs(ref)))
    tol = 1e-2 if out_dtype == "bf16" else 1e-4
    ok = ok and worst < tol
    for g in groups:
        g.close()
    tag = (f"P={P} depth={depth} epi={'tma' if epi else 'stg'} tr={int(tr)} {m_local}x{N}x{K} {in_dtype}->{out_dtype} steps={steps} "
           f"{'deferred-wait' if deferred else 'lock-step'}")
    print(f"{'OK  ' if ok else 'FAIL'} {tag}: every rank's cube == the unfused panels bit for bit: {ok}; rel err vs f64 {worst:.2e}", flush=True)
    return ok


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    cases = []
    for epi in (1, 0):
        for depth, deferred in ((1, False), (2, False), (3, True), (3, False)):
            cases.append((2, depth, epi, False, 256, 520, 320, 5, deferred, "bf16", "bf16"))
    cases += [
        (4, 3, 1, True, 200, 328, 136, 6, True, "bf16", "bf16"),      # ragged M / N (TMA clips), transposed A
        (4, 1, 1, False, 128, 256, 192, 3, False, "f32", "bf16"),     # f32 panels, one CTA per tile (M <= 128)
        (3, 2, 1, False, 264, 200, 96, 4, False, "bf16", "bf16"),
        (2, 3, 1, False, 256, 384, 256, 4, True, "f32", "f32"),       # 3xTF32 panels through the same protocol
        (8, 3, 1, False, 128, 136, 128, 4, True, "bf16", "bf16"),     # eight ranks
    ]
    if quick:
        cases = cases[:2] + cases[8:10]
    if os.environ.get("LOOPBACK_CASES"):
        cases = [cases[int(i)] for i in os.environ["LOOPBACK_CASES"].split(",")]
    ok = True
    for c in cases:
        ok &= run_case(*c)
    print("LOOPBACK CHECK", "ALL OK" if ok else "FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
