"""Golden vectors from the reference's own shaders: runs the unmodified WGSL of /root/reference through
tests/golden/wgsl_interp.py and stores the resulting buffers in tests/golden/ref_wgsl_*.npz.

This is the generating script of those fixtures.  It runs in the build container only (it reads /root/reference, which does
not exist on the GPU box); tests read the committed .npz files.  The shader text is composed the way the reference's Rust
side composes it — each step below cites the Rust lines it mirrors — and dispatched with the reference's grid rules.

    python tests/golden/make_reference_vectors.py [shape] [linalg] [geometry] [scan_sort] [check]    (default: all families; `check`
    re-runs and compares with the committed files instead of writing them; about 12 minutes for everything)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import reference_cases as C  # noqa: E402
import wgsl_interp as W  # noqa: E402

REF = "/root/reference/crates"


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def u8(a):
    return a.view(np.uint8)


CHECK = False


def save(name, out):
    """writes the fixture, or with `check` on the command line compares a fresh run with the committed file"""
    path = os.path.join(HERE, name)
    if not CHECK:
        np.savez_compressed(path, **out)
        return
    old = np.load(path)
    assert sorted(old.files) == sorted(out), f"{name}: case lists differ"
    bad = [k for k in out if not np.array_equal(np.asarray(out[k]).view(np.uint32), old[k].view(np.uint32))]
    print(f"check {name}: {len(out) - len(bad)} of {len(out)} buffers reproduce bit for bit" + (f"; DIFFERENT: {bad}" if bad else ""), flush=True)
    if bad:
        raise SystemExit(1)


# ------------------------------------------------------------------------------------------------ wgebra::linalg
def linalg_program(main_rel):
    """Shape::composer() + the kernel's own source (#[shader(derive(Shape), src = "gemm.wgsl")], gemm.rs:10, gemv.rs:10;
    Shape::composer()?.make_naga_module(..), op_assign.rs:53-57, reduce.rs:71-75)"""
    pr = W.Program()
    pr.add_module(read("wgebra/src/linalg/shape.wgsl"))
    pr.set_main(read(main_rel))
    return pr


def div_ceil(a, b):
    return (a + b - 1) // b


def run_gemm(case):
    pr = linalg_program("wgebra/src/linalg/gemm.wgsl")
    b = C.inputs(case)
    so, s1, s2 = case["so"], case["s1"], case["s2"]
    out_rows, out_mats = so[0], so[2]
    # gemm.rs:111-117: one invocation per 4 rows but out_rows.div_ceil(64) workgroups for gemm / gemm_tr; one workgroup per
    # 4 rows for the *_fast variants; grid = [dispatch, out_mats, 1] (gemm.rs:128)
    gx = div_ceil(out_rows, 4) if case["variant"].endswith("fast") else div_ceil(out_rows, 64)
    # bind0 order: out shape, m1 shape, m2 shape, out, m1, m2 (gemm.rs:120-127)
    pr.dispatch(case["variant"], {(0, 0): C.shape_bytes(so), (0, 1): C.shape_bytes(s1), (0, 2): C.shape_bytes(s2),
                                  (0, 3): u8(b["out"]), (0, 4): u8(b["m1"]), (0, 5): u8(b["m2"])}, (gx, out_mats, 1))
    return b["out"]


def run_gemv(case):
    pr = linalg_program("wgebra/src/linalg/gemv.wgsl")
    b = C.inputs(case)
    so, sm, sv = case["so"], case["sm"], case["sv"]
    out_nrows, out_ncols, out_nmats = so[0], so[1], so[2]
    # gemv.rs:114-127: div_ceil(32) workgroups for gemv / gemv_tr, one workgroup per 4 rows for the fast variants;
    # grid = [dispatch, out_ncols, out_nmats] (gemv.rs:138)
    gx = div_ceil(out_nrows, 4) if case["variant"].endswith("fast") else div_ceil(out_nrows, 32)
    pr.dispatch(case["variant"], {(0, 0): C.shape_bytes(so), (0, 1): C.shape_bytes(sm), (0, 2): C.shape_bytes(sv),
                                  (0, 3): u8(b["out"]), (0, 4): u8(b["m"]), (0, 5): u8(b["v"])}, (gx, out_ncols, out_nmats))
    return b["out"]


OP_ASSIGN_FN = {"add": "add_f32", "sub": "sub_f32", "mul": "mul_f32", "div": "div_f32", "copy": "copy_f32"}   # op_assign.rs:28-37
REDUCE_FNS = {  # reduce.rs:30-58: (init_fn, workspace_fn, reduce_fn)
    "min": ("init_max_f32", "reduce_min_f32", "reduce_min_f32"), "max": ("init_min_f32", "reduce_max_f32", "reduce_max_f32"),
    "sum": ("init_zero", "reduce_sum_f32", "reduce_sum_f32"), "prod": ("init_one", "reduce_prod_f32", "reduce_prod_f32"),
    "sqnorm": ("init_zero", "reduce_sqnorm_f32", "reduce_sum_f32")}


def run_op_assign(case):
    pr = linalg_program("wgebra/src/linalg/op_assign.wgsl")
    pr.redirect_function("placeholder", OP_ASSIGN_FN[case["op"]])                       # op_assign.rs:58-61
    b = C.inputs(case)
    # bind0: a shape, b shape, a, b; ceil(n / 64) workgroups (op_assign.rs:91-93)
    pr.dispatch("main", {(0, 0): C.shape_bytes(case["sa"]), (0, 1): C.shape_bytes(case["sb"]), (0, 2): u8(b["a"]), (0, 3): u8(b["b"])},
                (div_ceil(case["sa"][0], 64), 1, 1))
    return b["a"]


def run_reduce(case):
    pr = linalg_program("wgebra/src/linalg/reduce.wgsl")
    init_fn, workspace_fn, reduce_fn = REDUCE_FNS[case["op"]]
    pr.redirect_function("workspace_placeholder", workspace_fn)                          # reduce.rs:77-90
    pr.redirect_function("init_placeholder", init_fn)
    pr.redirect_function("reduce_placeholder", reduce_fn)
    b = C.inputs(case)
    # bind0: shape, input, output; exactly one workgroup (reduce.rs:108-112)
    pr.dispatch("main", {(0, 0): C.shape_bytes(case["s"]), (0, 1): u8(b["x"]), (0, 2): u8(b["out"])}, (1, 1, 1))
    return b["out"]


def make_linalg():
    out = {}
    run = {"gemm": run_gemm, "gemv": run_gemv, "op_assign": run_op_assign, "reduce": run_reduce}
    for case in C.all_linalg_cases():
        t = time.time()
        out[case["name"]] = run[case["kind"]](case)
        print(f"{case['kind']:10s} {case['name']:32s} {time.time() - t:7.1f} s", flush=True)
    save("ref_wgsl_linalg.npz", out)


# ------------------------------------------------------------------------------------------------ wgebra::geometry
def test_kernel_of(rs_rel):
    """the WGSL test kernel embedded in the reference's Rust test module (`let test_kernel = r#"..."#;`)"""
    import re
    m = re.search(r'let test_kernel = r#"(.*?)"#;', read(rs_rel), re.S)
    if not m:
        raise RuntimeError(f"no test kernel in {rs_rel}")
    return m.group(1)


# src_fn = "substituteN" of cholesky.rs:3-19 and lu.rs:5-27, applied in the same order
CHOLESKY_SUBST = {d: (("DIM", str(d)), ("MAT", f"mat{d}x{d}<f32>"), ("IMPORT_PATH", f"wgebra::cholesky{d}")) for d in (2, 3, 4)}
LU_SUBST = {d: (("NROWS", f"{d}u"), ("NCOLS", f"{d}u"), ("PERM", f"vec{d}<u32>"), ("MAT", f"mat{d}x{d}<f32>"), ("IMPORT_PATH", f"wgebra::lu{d}"))
            for d in (2, 3, 4)}


def substitute(src, table):
    for a, b in table:
        src = src.replace(a, b)
    return src


def geometry_program(main_src):
    """the composer the reference's derive macro builds: every module a shader derives from, registered under its
    #define_import_path (#[shader(derive(WgMinMax, WgSymmetricEigen2, WgRot2), ..)] eig3.rs:25, eig4.rs:25; derive(WgTrig)
    rot2.rs:5, svd2.rs:23; derive(WgQuat) svd3.rs:25)"""
    pr = W.Program()
    for rel in ("utils/trig.wgsl", "utils/min_max.wgsl", "geometry/rot2.wgsl", "geometry/quat.wgsl", "geometry/eig2.wgsl", "geometry/inv.wgsl"):
        pr.add_module(read("wgebra/src/" + rel))
    pr.set_main(main_src)
    return pr


# the one family without a test kernel in the reference (inv.rs has no tests): a wrapper of ours in the shape of the others
INV_KERNEL = """
#import wgebra::inv as Inv
@group(0) @binding(0)
var<storage, read_write> in: array<matDxD<f32>>;
@group(0) @binding(1)
var<storage, read_write> out: array<matDxD<f32>>;

@compute @workgroup_size(1, 1, 1)
fn test(@builtin(global_invocation_id) invocation_id: vec3<u32>) {
    let i = invocation_id.x;
    out[i] = Inv::invD(in[i]);
}
"""


def geometry_source(op, dim):
    g = "wgebra/src/geometry/"
    if op == "cholesky":      # cholesky.rs:72: substitute(format!("{}\n{}", S::src(), test_kernel)); S::src() is already substituted
        return substitute(substitute(read(g + "cholesky.wgsl"), CHOLESKY_SUBST[dim]) + "\n" + test_kernel_of(g + "cholesky.rs"), CHOLESKY_SUBST[dim])
    if op == "lu":            # lu.rs:118
        return substitute(substitute(read(g + "lu.wgsl"), LU_SUBST[dim]) + "\n" + test_kernel_of(g + "lu.rs"), LU_SUBST[dim])
    if op == "qr":            # qr2.rs:46, qr3.rs, qr4.rs
        return read(g + f"qr{dim}.wgsl") + "\n" + test_kernel_of(g + f"qr{dim}.rs")
    if op == "eig":           # eig2.rs:45, eig3.rs:48, eig4.rs
        return read(g + f"eig{dim}.wgsl") + "\n" + test_kernel_of(g + f"eig{dim}.rs")
    if op == "svd":           # svd2.rs:44, svd3.rs:46
        return read(g + f"svd{dim}.wgsl") + "\n" + test_kernel_of(g + f"svd{dim}.rs")
    if op == "inv":
        return INV_KERNEL.replace("D", str(dim))
    raise KeyError(op)


def make_geometry():
    from oracle import oracle as O
    out = {}
    for op, dim in C.geometry_cases():
        t = time.time()
        mats = C.geometry_inputs(op, dim)                         # [n, dim, dim]
        packed = O.geom_pack(mats)                                # WGSL storage layout of array<matDxD<f32>>
        n = packed.shape[0]
        src = geometry_source(op, dim)
        pr = geometry_program(src)
        if op == "eig" and dim == 2:
            # eig2.wgsl is also registered as the importable module wgebra::eig2 (for eig3 / eig4); as a main module it is the
            # same text plus the kernel
            pass
        words = C.GEOMETRY_OUT_WORDS[(op, dim)]
        res = np.zeros((n, words), np.float32)
        pr.dispatch("test", {(0, 0): u8(packed.reshape(-1)), (0, 1): u8(res.reshape(-1))}, (n, 1, 1))   # .dispatch(matrices.len()) cholesky.rs:119
        out[f"{op}{dim}"] = res
        print(f"geometry   {op}{dim:<28d} {time.time() - t:7.1f} s", flush=True)
    save("ref_wgsl_geometry.npz", out)


# ------------------------------------------------------------------------------------------------ prefix sum / radix sort
def run_prefix_sum(data):
    """WgPrefixSum::dispatch (wgrapier/src/dynamics/prefix_sum.rs:47-99) with PrefixSumWorkspace::reserve (:171-201):
    stage buffers of ceil(n / 256), ceil(that / 256), ... elements, the last one always of length 1"""
    pr = W.Program()
    pr.set_main(read("wgrapier/src/dynamics/prefix_sum.wgsl"))
    THREADS = 256
    stages = []
    stage_len = div_ceil(len(data), THREADS)
    while stage_len != 1:
        stages.append(np.zeros(stage_len, np.uint32))
        stage_len = div_ceil(stage_len, THREADS)
    stages.append(np.zeros(1, np.uint32))
    num_stages = len(stages)
    ngroups0 = len(stages[0])
    pr.dispatch("prefix_sum", {(0, 0): u8(data), (0, 1): u8(stages[0])}, (ngroups0, 1, 1))                    # :64-68
    for i in range(num_stages - 1):                                                                          # :70-78
        pr.dispatch("prefix_sum", {(0, 0): u8(stages[i]), (0, 1): u8(stages[i + 1])}, (len(stages[i + 1]), 1, 1))
    if num_stages > 2:                                                                                       # :80-90
        for i in reversed(range(num_stages - 2)):
            pr.dispatch("add_data_grp", {(0, 0): u8(stages[i]), (0, 1): u8(stages[i + 1])}, (len(stages[i + 1]), 1, 1))
    if num_stages > 1:                                                                                       # :92-96
        pr.dispatch("add_data_grp", {(0, 0): u8(data), (0, 1): u8(stages[0])}, (ngroups0, 1, 1))
    return data


def run_radix_sort(keys, values, n_sort, sorting_bits):
    """RadixSort::dispatch (wgparry/src/utils/radix_sort/mod.rs:204-322).  The scan shaders read `counts` / `reduced` past the
    number of workgroups without a bounds check (sort_scan_add.wgsl: every invocation loads 4 entries whatever num_wgs is) and
    only use what lies below it, so they depend on WebGPU's robust buffer access: Program(robust=True)."""
    g = "wgparry/src/utils/radix_sort/"
    BLOCK_SIZE = 1024

    def program(name):
        pr = W.Program(robust=True)
        pr.add_module(read(g + "sorting.wgsl"))
        pr.set_main(read(g + name))
        return pr
    init, count, reduce_, scan, scan_add, scatter = (program(n) for n in ("init_indirect_dispatches.wgsl", "sort_count.wgsl", "sort_reduce.wgsl",
                                                                          "sort_scan.wgsl", "sort_scan_add.wgsl", "sort_scatter.wgsl"))
    max_n = len(keys)
    n_sort_buf = np.array([n_sort], np.uint32)
    count_buf = np.zeros(div_ceil(max_n, BLOCK_SIZE) * 16, np.uint32)                   # :219-223
    reduced_buf = np.zeros(BLOCK_SIZE, np.uint32)                                       # :120-124
    num_wgs, num_reduce_wgs = np.ones(3, np.uint32), np.ones(3, np.uint32)
    init.dispatch("main", {(0, 0): u8(n_sort_buf), (0, 1): u8(num_wgs), (0, 2): u8(num_reduce_wgs)}, (1, 1, 1))   # :225-231
    out_keys, out_values = keys.copy(), values.copy()                                   # the test initialises the outputs with the inputs (:276-277)
    pong_keys, pong_values = np.zeros(max_n, np.uint32), np.zeros(max_n, np.uint32)
    cur_keys, cur_vals = keys.copy(), values.copy()
    user_keys, user_values = out_keys, out_values
    num_passes = div_ceil(sorting_bits, 4)
    if num_passes % 2 == 0:                                                             # :252-258
        out_keys, pong_keys = pong_keys, out_keys
        out_values, pong_values = pong_values, out_values
    for pass_id in range(num_passes):
        uniforms = np.array([pass_id * 4], np.uint32)                                   # :261-267
        count.dispatch("main", {(0, 0): u8(uniforms), (0, 1): u8(n_sort_buf), (0, 2): u8(cur_keys), (0, 3): u8(count_buf)}, tuple(num_wgs))
        reduce_.dispatch("main", {(0, 0): u8(n_sort_buf), (0, 1): u8(count_buf), (0, 2): u8(reduced_buf)}, tuple(num_reduce_wgs))
        scan.dispatch("main", {(0, 0): u8(n_sort_buf), (0, 1): u8(reduced_buf)}, (1, 1, 1))
        scan_add.dispatch("main", {(0, 0): u8(n_sort_buf), (0, 1): u8(reduced_buf), (0, 2): u8(count_buf)}, tuple(num_reduce_wgs))
        scatter.dispatch("main", {(0, 0): u8(uniforms), (0, 1): u8(n_sort_buf), (0, 2): u8(cur_keys), (0, 3): u8(cur_vals), (0, 4): u8(count_buf),
                                  (0, 5): u8(out_keys), (0, 6): u8(out_values)}, tuple(num_wgs))
        if pass_id == 0:                                                                # :312-320
            cur_keys, cur_vals = out_keys, out_values
            out_keys, out_values = pong_keys, pong_values
        else:
            cur_keys, out_keys = out_keys, cur_keys
            cur_vals, out_values = out_values, cur_vals
    oob = sum(p.oob_accesses for p in (init, count, reduce_, scan, scan_add, scatter))
    assert cur_keys is user_keys and cur_vals is user_values, "the last pass must land in the caller's buffers"
    return user_keys, user_values, oob


def make_scan_sort():
    out = {}
    for case in C.scan_cases():
        t = time.time()
        out["scan/" + case["name"]] = run_prefix_sum(C.scan_input(case).copy())
        print(f"scan       {case['name']:32s} {time.time() - t:7.1f} s", flush=True)
    for case in C.sort_cases():
        t = time.time()
        keys, values = C.sort_input(case)
        k, v, oob = run_radix_sort(keys, values, case["n_sort"], case["bits"])
        out["sort/" + case["name"] + "/keys"], out["sort/" + case["name"] + "/values"] = k, v
        print(f"sort       {case['name']:32s} {time.time() - t:7.1f} s   ({oob} out-of-bounds reads served by robust buffer access)", flush=True)
    save("ref_wgsl_scan_sort.npz", out)


# ------------------------------------------------------------------------------------------------ shape.wgsl by itself
def make_shape():
    """iv / im / it / with_vec4_elts of shape.wgsl, called directly, for the column-major build and for the ROW_MAJOR build
    (row_major_shader_defs(), shape.rs:11-15)"""
    out = {"views": np.array(C.SHAPE_VIEWS, np.uint32), "queries": np.array(C.SHAPE_QUERIES, np.uint32)}
    for tag, defs in (("col", ()), ("row", ("ROW_MAJOR",))):
        pr = W.Program()
        mod = pr.set_main(read("wgebra/src/linalg/shape.wgsl"), defs)
        ty = mod.resolve_type(("ty", "Shape", [], 0))
        names = [n for n, _ in ty.members]
        res = np.zeros((len(C.SHAPE_VIEWS), len(C.SHAPE_QUERIES), 3), np.uint32)
        v4 = np.zeros((len(C.SHAPE_VIEWS), 6), np.uint32)
        for a, view in enumerate(C.SHAPE_VIEWS):
            sv = W.StructVal(ty, {n: np.uint32(x) for n, x in zip(names, view)})
            r = pr.call_function("with_vec4_elts", [sv])
            v4[a] = [r.f[n] for n in names]
            for q, (i, j, t) in enumerate(C.SHAPE_QUERIES):
                i, j, t = np.uint32(i), np.uint32(j), np.uint32(t)
                res[a, q] = [pr.call_function("iv", [sv, i]), pr.call_function("im", [sv, i, j]), pr.call_function("it", [sv, i, j, t])]
        out[tag + "/index"], out[tag + "/vec4"] = res, v4
    save("ref_wgsl_shape.npz", out)
    print("shape      iv / im / it / with_vec4_elts, both builds", flush=True)


if __name__ == "__main__":
    CHECK = "check" in sys.argv[1:]
    which = [a for a in sys.argv[1:] if a != "check"] or ["shape", "linalg", "geometry", "scan_sort"]
    if "shape" in which:
        make_shape()
    if "linalg" in which:
        make_linalg()
    if "geometry" in which:
        make_geometry()
    if "scan_sort" in which:
        make_scan_sort()
