"""GPU parity tests (-m gpu) that FORCE every tcgen05 kernel instantiation instead of taking what the heuristics pick.

An instantiation is gemm_tc_kernel<KIND, A_MN, B_MN, BLOCK_N, PASSES, TOut, CG> (wgmath_b200/csrc/gemm_tc_kernel.cuh); the
44 of them are listed in INSTANTIATIONS and each is a test id.  Inside a test every tail plan (none / N-split / split-K) and both
epilogue forms (per-lane stores / TMA bulk stores) run on ragged, batched, offset views, and wgb_pass_last_gemm_config proves that
the forced configuration is the one that ran.  Reference semantics: gemm.rs:78-126 (out = m1 * m2 or tr(m1) * m2, overwrite,
batched over size[2]); expected values from the oracle's restatement of the view addressing (oracle/wgsl_oracle.c orc_gemm_ord).

Tolerances (BASELINE.json north_star): 1e-5 relative for 3xTF32, 1e-2 for bf16 output; bf16 -> f32 output shows the exact product
(1e-4), single-pass TF32 is TF32-accurate (2e-3, informational mode)."""
import itertools

import pytest

import wgmath_b200 as w
from tests.test_gpu_orderings import gemm_ord_case

pytestmark = pytest.mark.gpu

# (kind, a_mn, b_mn, bn, passes, out, cg)
INSTANTIATIONS = (
    [("bf16", a, b, bn, 1, out, cg) for a in (0, 1) for b in (0, 1) for out in ("bf16", "f32") for bn in (128, 256) for cg in (1, 2)]
    + [("tf32", a, 0, bn, 1, "f32", cg) for a in (0, 1) for bn in (128, 256) for cg in (1, 2)]
    + [("tf32", a, 0, 128, 3, "f32", cg) for a in (0, 1) for cg in (1, 2)])
assert len(INSTANTIATIONS) == 44


def inst_id(i):
    kind, a, b, bn, passes, out, cg = i
    return f"kind={kind}-a_mn={a}-b_mn={b}-bn={bn}-passes={passes}-out={out}-cg={cg}"


TAIL_PLANS = {"none": {"WGB_TC_NSPLIT": "0", "WGB_TC_SPLITK": "0"}, "nsplit": {"WGB_TC_NSPLIT": "1", "WGB_TC_SPLITK": "0"},
              "splitk": {"WGB_TC_NSPLIT": "0", "WGB_TC_SPLITK": "1"}}
# ragged sizes with a K long enough for a K split (>= 8 k-blocks of 64 bf16 / 32 tf32), batched, padded parents, non-zero offsets
SHAPES = [dict(M=264, N=200, K=520, T=2, pad=(8, 8, 8), off=(8, 16, 24)), dict(M=384, N=520, K=136, T=1, pad=(0, 0, 0), off=(0, 0, 0))]


@pytest.mark.parametrize("inst", INSTANTIATIONS, ids=inst_id)
def test_tc_instantiation(gpu, shapes, monkeypatch, inst):
    kind, a_mn, b_mn, bn, passes, out, cg = inst
    monkeypatch.setenv("WGB_TC_BN", str(bn))
    monkeypatch.setenv("WGB_TC_CG", str(cg))
    f32 = kind == "tf32"
    mode = None if not f32 else (w.F32Mode.X3Tf32 if passes == 3 else w.F32Mode.Tf32)
    tol = (1e-5 if passes == 3 else 2e-3) if f32 else (1e-2 if out == "bf16" else 1e-4)
    # how a kernel with a K-major A is reached: the transposed variant; for f32 also the non-transposed product through the
    # transposing operand prep (WGB_TF32_MN_DIRECT=0)
    routes = [dict(tr=not a_mn, env={"WGB_TF32_MN_DIRECT": "1", "WGB_TF32_FUSED_SPLIT": "0"}, fs=0)]
    if f32 and not a_mn:
        routes.append(dict(tr=False, env={"WGB_TF32_MN_DIRECT": "0", "WGB_TF32_FUSED_SPLIT": "1"}, fs=0))   # transposing prep: no in-kernel split
    if passes == 3:
        # the FS kernels (operand split inside the GEMM, raw f32 tiles by TMA) are separate instantiations of the same family
        routes.append(dict(tr=not a_mn, env={"WGB_TF32_MN_DIRECT": "1", "WGB_TF32_FUSED_SPLIT": "1"}, fs=1))
    seen = set()
    for route, (plan, plan_env), epi, sh in itertools.product(routes, TAIL_PLANS.items(), (0, 1), SHAPES):
        for k, v in {**route["env"], **plan_env, "WGB_TC_EPI": str(epi)}.items():
            monkeypatch.setenv(k, v)
        cfg = []
        path = gemm_ord_case(gpu, shapes, sh["M"], sh["N"], sh["K"], route["tr"], 0, 0, b_mn, T=sh["T"], dtype="f32" if f32 else "bf16",
                             out_dtype=out, mode=mode, pad=sh["pad"], off=sh["off"], tol=tol, cfg_out=cfg, out_extra=16)
        assert path == (2 if not f32 else 4 if passes == 3 else 3)
        c = cfg[0]
        got = (("tf32" if c["kind"] else "bf16"), c["a_mn"], c["b_mn"], c["bn"], c["passes"], "bf16" if c["out_dtype"] == 1 else "f32", c["cg"])
        assert got == inst, f"forced {inst_id(inst)} but {got} ran"
        assert c["epi_tma"] == epi
        assert c["fused_split"] == route["fs"]
        if plan == "none":
            assert c["nsplit"] == 1 and c["splitk"] == 1
        seen.add((plan, c["nsplit"] > 1, c["splitk"] > 1))
    # the tail plans were not only requested but taken on at least one shape
    # (an MN-major B is loaded in whole 128-byte atoms of N: 64-column halves of a CTA pair's 128-column tile cannot be split further)
    if not (b_mn and bn == 128 and cg == 2):
        assert ("nsplit", True, False) in seen, seen
    assert ("splitk", False, True) in seen, seen


@pytest.mark.parametrize("epi", [0, 1])
@pytest.mark.parametrize("fake_peers", [1, 3, 8])
def test_epilogue_store_forms_with_several_destinations(gpu, shapes, monkeypatch, epi, fake_peers):
    """The fused all-gather's epilogue loop over destinations, on one GPU: WGB_TC_DEBUG_FAKE_PEERS=n makes the kernel store every
    output block n times (replica 0 is the real output, the others scratch panels).  Both store forms, every destination count."""
    monkeypatch.setenv("WGB_TC_EPI", str(epi))
    monkeypatch.setenv("WGB_TC_DEBUG_FAKE_PEERS", str(fake_peers))
    cfg = []
    gemm_ord_case(gpu, shapes, 512, 776, 192, False, 0, 0, 0, dtype="bf16", out_dtype="bf16", tol=1e-2, cfg_out=cfg)
    assert cfg[0]["dests"] == fake_peers and cfg[0]["epi_tma"] == epi
    cfg = []
    gemm_ord_case(gpu, shapes, 300, 264, 96, True, 0, 0, 0, dtype="bf16", out_dtype="f32", tol=1e-4, cfg_out=cfg)
    assert cfg[0]["dests"] == fake_peers


def test_tma_epilogue_falls_back_on_unaligned_output(gpu, shapes, monkeypatch):
    """A leading dimension that is not a multiple of 16 bytes cannot be a TMA store: the per-lane form must run (and be right)."""
    monkeypatch.setenv("WGB_TC_EPI", "1")
    cfg = []
    gemm_ord_case(gpu, shapes, 256, 136, 128, False, 0, 0, 0, dtype="bf16", out_dtype="bf16", tol=1e-2, pad=(0, 0, 3), off=(0, 0, 1), cfg_out=cfg)
    assert cfg[0]["epi_tma"] == 0
