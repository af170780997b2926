"""The oracle against golden vectors computed FROM THE REFERENCE'S OWN SHADER SOURCE.

tests/golden/ref_wgsl_linalg.npz holds the output buffers of the unmodified WGSL files of /root/reference (gemm.wgsl,
gemv.wgsl, op_assign.wgsl, reduce.wgsl composed with shape.wgsl the way the Rust side composes them), executed by the WGSL
interpreter tests/golden/wgsl_interp.py with the reference's dispatch grids (tests/golden/make_reference_vectors.py).  The C
restatement in oracle/ must reproduce every one of those buffers BIT FOR BIT — padding, untouched regions and the garbage
rows the shaders write past ragged views included.  That pins the oracle's index arithmetic, loop structure, reduction trees,
initial values and edge behaviour to the reference's text; what stays implementation-defined in WGSL (FMA contraction, the
summation order inside mat * mat / mat * vec) is fixed to the same choice in both (header of wgsl_interp.py).

The interpreter itself is unit-tested below on shaders written for these tests (no reference text involved)."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)
import reference_cases as C  # noqa: E402
import wgsl_interp as W  # noqa: E402


@pytest.fixture(scope="module")
def linalg_vectors():
    return np.load(os.path.join(GOLD, "ref_wgsl_linalg.npz"))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def oracle_run(case):
    b = C.inputs(case)
    k = case["kind"]
    if k == "gemm":
        rc = O.gemm(C.GEMM_VARIANTS[case["variant"]], b["out"], O.Shape(*case["so"]), b["m1"], O.Shape(*case["s1"]), b["m2"], O.Shape(*case["s2"]))
        assert rc == O.ORC_OK
        return b["out"]
    if k == "gemv":
        rc, ran = O.gemv(C.GEMV_VARIANTS[case["variant"]], b["out"], O.Shape(*case["so"]), b["m"], O.Shape(*case["sm"]), b["v"], O.Shape(*case["sv"]))
        assert rc == O.ORC_OK and ran == C.GEMV_VARIANTS[case["variant"]]
        return b["out"]
    if k == "op_assign":
        assert O.op_assign(C.OP_ASSIGN[case["op"]], b["a"], O.Shape(*case["sa"]), b["b"], O.Shape(*case["sb"])) == O.ORC_OK
        return b["a"]
    return np.array([O.reduce(C.REDUCE[case["op"]], b["x"], O.Shape(*case["s"]))], np.float32)


@pytest.mark.parametrize("case", C.all_linalg_cases(), ids=lambda c: f"{c['kind']}:{c['name']}")
def test_oracle_reproduces_the_reference_shaders_bit_for_bit(linalg_vectors, case):
    want = linalg_vectors[case["name"]]
    got = oracle_run(case)
    assert got.shape == want.shape
    np.testing.assert_array_equal(bits(got), bits(want))


def test_every_stored_vector_has_a_case(linalg_vectors):
    assert sorted(linalg_vectors.files) == sorted(c["name"] for c in C.all_linalg_cases())
    geo = np.load(os.path.join(GOLD, "ref_wgsl_geometry.npz"))
    assert sorted(geo.files) == sorted(f"{op}{dim}" for op, dim in C.geometry_cases())
    ss = np.load(os.path.join(GOLD, "ref_wgsl_scan_sort.npz"))
    want = ["scan/" + c["name"] for c in C.scan_cases()] + [f"sort/{c['name']}/{w}" for c in C.sort_cases() for w in ("keys", "values")]
    assert sorted(ss.files) == sorted(want)


def test_shape_index_functions_match_shape_wgsl_in_both_builds():
    """shape.wgsl's iv / im / it / with_vec4_elts executed by the interpreter in the column-major build and in the ROW_MAJOR build
    (shape.rs:11-15): the oracle's index functions — what orc_gemm / orc_gemm_ord address memory with — and the element
    addressing of the host mirror's views agree with both tables."""
    import wgmath_b200 as w
    g = np.load(os.path.join(GOLD, "ref_wgsl_shape.npz"))
    assert g["views"].tolist() == [list(v) for v in C.SHAPE_VIEWS] and g["queries"].tolist() == [list(q) for q in C.SHAPE_QUERIES]
    for tag, rm in (("col", False), ("row", True)):
        for a, view in enumerate(C.SHAPE_VIEWS):
            s = O.Shape(*view)
            for q, (i, j, t) in enumerate(C.SHAPE_QUERIES):
                want = g[tag + "/index"][a, q]
                got = [O.shape_index(s, rm, fn, i, j, t)[0] for fn in (0, 1, 2)]
                assert got == want.tolist(), (tag, view, (i, j, t))
            assert list(O.shape_index(s, rm, 0, 0)[1]) == g[tag + "/vec4"][a].tolist()
            # host mirror: a view's sub-view constructors move `offset` by the same element strides
            vs = w.ViewShape((view[0], view[1], view[2]), view[3], view[4], view[5])
            tv = w.GpuTensorView(vs, None, "f32", 3, w.RowMajor if rm else w.ColumnMajor)
            if view[0] > 3 and view[1] > 4 and view[2] > 2:
                sub = tv.matrix(2).columns(4, 1).rows(3, 1)
                assert sub.shape().offset == int(g[tag + "/index"][a, 3][2])          # query (3, 4, 2)


# ------------------------------------------------------------------------------------------------ factorizations, scan, sort
@pytest.fixture(scope="module")
def geometry_vectors():
    return np.load(os.path.join(GOLD, "ref_wgsl_geometry.npz"))


@pytest.fixture(scope="module")
def scan_sort_vectors():
    return np.load(os.path.join(GOLD, "ref_wgsl_scan_sort.npz"))


@pytest.mark.parametrize("op,dim", C.geometry_cases(), ids=lambda x: str(x))
def test_geometry_oracle_reproduces_the_reference_shaders_bit_for_bit(geometry_vectors, op, dim):
    """cholesky.wgsl / lu.wgsl / qr*.wgsl / eig*.wgsl / svd*.wgsl / inv.wgsl composed with the test kernels embedded in the
    reference's Rust tests: every output word (factors, permutations, zero padding) equals geometry_oracle.c's."""
    want = geometry_vectors[f"{op}{dim}"]
    got = O.geom_batch(C.GEOMETRY_OPS[op], dim, O.geom_pack(C.geometry_inputs(op, dim)))
    assert got.shape == want.shape == (C.GEOMETRY_BATCH, C.GEOMETRY_OUT_WORDS[(op, dim)])
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("case", C.scan_cases(), ids=lambda c: c["name"])
def test_scan_oracle_reproduces_the_reference_shaders(scan_sort_vectors, case):
    data = C.scan_input(case).copy()
    assert O.prefix_sum(data) == O.ORC_OK
    np.testing.assert_array_equal(data, scan_sort_vectors["scan/" + case["name"]])


@pytest.mark.parametrize("case", C.sort_cases(), ids=lambda c: c["name"])
def test_sort_oracle_reproduces_the_reference_shaders(scan_sort_vectors, case):
    keys, values = C.sort_input(case)
    out_k, out_v = keys.copy(), values.copy()                     # as the reference's test initialises them (mod.rs:276-277)
    assert O.radix_sort(keys, values, case["n_sort"], case["bits"], out_k, out_v) == O.ORC_OK
    n = case["n_sort"]
    np.testing.assert_array_equal(out_k[:n], scan_sort_vectors[f"sort/{case['name']}/keys"][:n])
    np.testing.assert_array_equal(out_v[:n], scan_sort_vectors[f"sort/{case['name']}/values"][:n])


# ------------------------------------------------------------------------------------------------ the interpreter itself
def run(src, entry, bindings, grid, modules=(), defs=()):
    pr = W.Program()
    for m in modules:
        pr.add_module(m, defs)
    pr.set_main(src, defs)
    pr.dispatch(entry, bindings, grid)
    return pr


def test_interp_barriers_pointers_and_workgroup_memory():
    src = """
    @group(0) @binding(0) var<storage, read_write> data: array<f32>;
    @group(0) @binding(1) var<uniform> n: u32;
    const WG: u32 = 8;
    var<workgroup> sk: array<f32, WG>;
    fn red(i: u32, s: u32) { if i < s { sk[i] += sk[i + s]; } workgroupBarrier(); }
    fn bump(p: ptr<function, vec2<f32>>, k: u32) { (*p)[k] = (*p)[k] * 2.0; }
    fn total(i: u32) -> f32 { red(i, 4u); red(i, 2u); red(i, 1u); return sk[0]; }
    @compute @workgroup_size(WG, 1, 1)
    fn main(@builtin(local_invocation_id) lid: vec3<u32>, @builtin(workgroup_id) wid: vec3<u32>) {
      var acc = 0.0;
      for (var i = lid.x; i < n; i += WG) { acc += data[i + wid.x * n]; }
      sk[lid.x] = acc;
      workgroupBarrier();
      let t = total(lid.x);
      var q = vec2(1.0, 3.0);
      bump(&q, 1u);
      if lid.x == 0u { data[wid.x * n] = t + q.y; }
    }"""
    buf = np.arange(32, dtype=np.float32)
    run(src, "main", {(0, 0): buf.view(np.uint8), (0, 1): np.array([16], np.uint32).view(np.uint8)}, (2, 1, 1))
    assert buf[0] == sum(range(16)) + 6 and buf[16] == sum(range(16, 32)) + 6


def test_interp_storage_layout_of_vec3_mat3_and_structs():
    src = """
    struct S { a: u32, v: vec3<f32>, m: mat3x3<f32>, b: f32 }
    @group(0) @binding(0) var<storage, read_write> s: array<S>;
    @compute @workgroup_size(1)
    fn main(@builtin(global_invocation_id) id: vec3<u32>) {
      var x = s[id.x];
      x.m[1][2] = x.v.z + f32(x.a);
      x.b = x.m[2].y;
      s[id.x] = x;
    }"""
    pr = W.Program()
    mod = pr.set_main(src)
    ty = mod.resolve_type(("ty", "S", [], 0))
    assert W.layout(ty) == (96, 16)                       # a @0, v @16, m @32 (3 columns x 16), b @80, size rounded to 96
    assert W.member_offset(ty, "v")[0] == 16 and W.member_offset(ty, "m")[0] == 32 and W.member_offset(ty, "b")[0] == 80
    words = np.arange(48, dtype=np.float32)
    raw = words.view(np.uint8).copy()
    raw.view(np.uint32)[0], raw.view(np.uint32)[24] = 5, 7
    before = raw.copy()
    pr.dispatch("main", {(0, 0): raw}, (2, 1, 1))
    f, fb = raw.view(np.float32), before.view(np.float32)
    for base, a in ((0, 5), (24, 7)):
        assert f[base + 8 + 4 + 2] == fb[base + 4 + 2] + a          # m[1][2] = v.z + a
        assert f[base + 20] == fb[base + 8 + 8 + 1]                 # b = m[2].y


def test_interp_preprocessor_imports_and_redirect():
    lib = """
    #define_import_path demo::lib
    #ifdef TWICE
    fn f(x: f32) -> f32 { return x * 2.0; }
    #else
    fn f(x: f32) -> f32 { return x + 1.0; }
    #endif
    """
    src = """
    #import demo::lib as L
    @group(0) @binding(0) var<storage, read_write> d: array<f32>;
    fn plus(a: f32) -> f32 { return a + 100.0; }
    fn hook(a: f32) -> f32 { return a; }
    @compute @workgroup_size(4)
    fn main(@builtin(global_invocation_id) id: vec3<u32>) { d[id.x] = hook(L::f(d[id.x])); }
    """
    for defs, redirect, want in (((), False, [1, 2, 3, 4]), (("TWICE",), False, [0, 2, 4, 6]), (("TWICE",), True, [100, 102, 104, 106])):
        d = np.arange(4, dtype=np.float32)
        pr = W.Program()
        pr.add_module(lib, defs)
        pr.set_main(src, defs)
        if redirect:
            pr.redirect_function("hook", "plus")
        pr.dispatch("main", {(0, 0): d.view(np.uint8)}, (1, 1, 1))
        assert d.tolist() == want


def test_interp_arithmetic_model():
    src = """
    @group(0) @binding(0) var<storage, read_write> o: array<f32>;
    @group(0) @binding(1) var<storage, read_write> u: array<u32>;
    @group(0) @binding(2) var<storage, read_write> c: atomic<u32>;
    @compute @workgroup_size(3)
    fn main(@builtin(local_invocation_index) li: u32) {
      let a = vec4(o[0], o[1], o[2], o[3]);
      let m = mat4x4(a, a * 2.0, a + vec4(1.0), a - vec4(0.5));
      if li == 0u {
        o[4] = dot(a, a);
        let p = m * a;
        o[5] = p.x;
        o[6] = fma(o[0], o[1], o[2]);
        o[7] = (transpose(m) * a).y;
        u[0] = u[1] - 3u;                  // wraps
        u[2] = u[3] / 0u + (7u % 4u);      // x / 0 = x
        u[4] = select(1u, 2u, o[0] < o[1]) << 4u;
      }
      atomicAdd(&c, li + 1u);
    }"""
    o = np.array([0.1, 0.7, 1.3, 2.9, 0, 0, 0, 0], np.float32)
    u = np.array([0, 1, 0, 9, 0], np.uint32)
    c = np.zeros(1, np.uint32)
    run(src, "main", {(0, 0): o.view(np.uint8), (0, 1): u.view(np.uint8), (0, 2): c.view(np.uint8)}, (1, 1, 1))
    a = o[:4]
    f = np.float32
    assert o[4] == ((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]) + a[3] * a[3]                     # left to right, no contraction
    cols = [a, a * f(2), a + f(1), a - f(0.5)]
    assert o[5] == ((cols[0][0] * a[0] + cols[1][0] * a[1]) + cols[2][0] * a[2]) + cols[3][0] * a[3]
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.fmaf.restype = ctypes.c_float
    libm.fmaf.argtypes = [ctypes.c_float] * 3
    assert o[6] == f(libm.fmaf(float(a[0]), float(a[1]), float(a[2])))
    assert o[7] == ((cols[1][0] * a[0] + cols[1][1] * a[1]) + cols[1][2] * a[2]) + cols[1][3] * a[3]
    assert u.tolist() == [0xFFFFFFFE, 1, 12, 9, 32] and c[0] == 6


def test_interp_control_flow_swizzles_bitcasts_and_struct_pointers():
    src = """
    struct Pair { lo: vec2<f32>, n: u32 }
    @group(0) @binding(0) var<storage, read_write> o: array<f32>;
    @group(0) @binding(1) var<storage, read_write> u: array<u32>;
    fn grow(p: ptr<function, Pair>) { (*p).lo.y += 1.0; (*p).n++; }
    fn classify(k: u32) -> u32 {
      switch k { case 0u, 1u: { return 10u; } case 2u: { return 20u; } default: { return 30u; } }
    }
    fn fast_rsqrt(val: f32) -> f32 {           // the bit trick svd3.wgsl uses: i32 arithmetic shift, wrapping subtraction
      var i = bitcast<i32>(val);
      i = 0x5f375a82 - (i >> 1);
      return bitcast<f32>(i);
    }
    @compute @workgroup_size(1)
    fn main() {
      var i = 0u;
      var acc = 0u;
      loop {
        if i == 2u { i++; continue; }
        acc += i;
        continuing { i++; break if i >= 6u; }
      }
      u[0] = acc;                              // 0 + 1 + 4 + 5: at i = 2 the body bumps i and `continue` still runs `continuing`
      u[1] = classify(1u) + classify(2u) + classify(7u);
      var p = Pair(vec2(1.0, 2.0), 5u);
      grow(&p);
      let q = &p;
      (*q).lo.x = 9.0;
      o[0] = p.lo.x + p.lo.y;
      u[2] = p.n;
      let v = vec4(1.0, 2.0, 3.0, 4.0);
      let w = v.zyx;
      o[1] = w.x * 100.0 + w.y * 10.0 + w.z;
      var m = mat3x3<f32>();
      m[1] = vec3(1.0, 2.0, 3.0);
      m[2][0] = 7.0;
      o[2] = (m * vec3(0.0, 1.0, 1.0)).x;
      o[3] = select(vec2(1.0, 2.0), vec2(3.0, 4.0), vec2(true, false)).x + select(vec2(1.0, 2.0), vec2(3.0, 4.0), vec2(true, false)).y;
      o[4] = fast_rsqrt(4.0);
      u[3] = bitcast<u32>(-0.0);
      var k = -7i;
      u[4] = u32(k >> 1u) & 0xffu;             // arithmetic shift: -4 -> 0xfc
      u[5] = u32(k / 2i + 10i) * 10u + u32(k % 3i + 5i);   // truncating division and remainder: -3, -1
      while k < 0i { k += 3i; }
      u[6] = u32(k);
      o[5] = max(o[6] - o[6], 1.5);            // inf - inf = nan: min / max return the other operand
    }"""
    o = np.zeros(7, np.float32)
    o[6] = np.inf
    u = np.zeros(7, np.uint32)
    run(src, "main", {(0, 0): o.view(np.uint8), (0, 1): u.view(np.uint8)}, (1, 1, 1))
    assert u.tolist() == [10, 60, 6, 0x80000000, 0xFC, 74, 2]
    i = np.array([4.0], np.float32).view(np.int32)[0]
    want = np.array([0x5f375a82 - (int(i) >> 1)], np.int32).view(np.float32)[0]
    assert o[:6].tolist() == [12.0, 321.0, 8.0, 5.0, want, 1.5]


def test_interp_matrix_conventions_and_precedence():
    """WGSL matrices are column-major: matCxR has C columns of R rows, m[c][r]; M * v combines columns, v * M dots with columns."""
    src = """
    @group(0) @binding(0) var<storage, read_write> o: array<f32>;
    @group(0) @binding(1) var<storage, read_write> m_io: mat2x3<f32>;
    @compute @workgroup_size(1)
    fn main() {
      let m = mat2x3<f32>(1.0, 2.0, 3.0, 4.0, 5.0, 6.0);       // columns (1,2,3) and (4,5,6)
      let a = m * vec2(10.0, 100.0);                             // 10 * col0 + 100 * col1
      o[0] = a.x; o[1] = a.y; o[2] = a.z;
      let b = vec3(1.0, 10.0, 100.0) * m;                        // (dot(v, col0), dot(v, col1))
      o[3] = b.x; o[4] = b.y;
      let t = transpose(m);                                      // mat3x2: t[r][c] = m[c][r]
      o[5] = t[2][1];
      let p = t * m;                                             // (2 rows x 3 columns) * (3 rows x 2 columns) = mat2x2
      o[6] = p[1][0];                                            // column 1, row 0
      o[7] = 2.0 + 3.0 * 4.0 - 6.0 / 3.0;                        // 12
      o[8] = f32((7u & 3u) | (8u >> 2u)) + f32(1u << 3u);        // (3 | 2) + 8 = 11
      o[9] = f32(-3i * -3i % 5i);                                // 9 % 5 = 4
      o[10] = select(1.0, 2.0, 1.0 < 2.0 && !(3.0 <= 2.0) || false);
      m_io[1][2] = m_io[0][1] + 0.5;                             // storage layout: column stride 16 bytes for 3 rows
    }"""
    o = np.zeros(11, np.float32)
    mio = np.arange(8, dtype=np.float32)                          # col0 = (0,1,2), pad, col1 = (4,5,6), pad
    run(src, "main", {(0, 0): o.view(np.uint8), (0, 1): mio.view(np.uint8)}, (1, 1, 1))
    assert o[:6].tolist() == [410.0, 520.0, 630.0, 321.0, 654.0, 6.0]
    # t is 3 columns of 2 rows: t[c][r] = m[r][c]; (t * m)[1][0] = sum_k t[k][0] * m[1][k] = 1*4 + 2*5 + 3*6
    assert o[6] == 32.0
    assert o[7:].tolist() == [12.0, 11.0, 4.0, 2.0]
    assert mio.tolist() == [0, 1, 2, 3, 4, 5, 1.5, 7]


def test_interp_rejects_what_it_cannot_run_faithfully():
    oob = """
    @group(0) @binding(0) var<storage, read_write> d: array<f32>;
    @compute @workgroup_size(1) fn main() { d[4] = 1.0; }"""
    with pytest.raises(W.WgslError, match="out-of-bounds"):
        run(oob, "main", {(0, 0): np.zeros(4, np.float32).view(np.uint8)}, (1, 1, 1))
    divergent = """
    var<workgroup> x: u32;
    @compute @workgroup_size(2) fn main(@builtin(local_invocation_index) i: u32) { if i == 0u { workgroupBarrier(); } }"""
    with pytest.raises(W.WgslError, match="non-uniform"):
        run(divergent, "main", {}, (1, 1, 1))
    # ... unless robust buffer access is asked for: reads give zero, writes are dropped, and the accesses are counted
    pr = W.Program(robust=True)
    pr.set_main(oob.replace("d[4] = 1.0;", "d[4] = 1.0; d[0] = d[7] + 2.0;"))
    d = np.ones(4, np.float32)
    pr.dispatch("main", {(0, 0): d.view(np.uint8)}, (1, 1, 1))
    assert d.tolist() == [2.0, 1.0, 1.0, 1.0] and pr.oob_accesses == 2
    with pytest.raises(W.WgslError, match="different types"):
        run("@compute @workgroup_size(1) fn main() { var a = 1u; var b = 1i; let c = a + b; }", "main", {}, (1, 1, 1))
