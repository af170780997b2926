"""A small WGSL interpreter: executes the reference's compute shaders *from their own source text* on the CPU.

TEST INFRASTRUCTURE (fixture generation).  The reference is Rust + wgpu and cannot be built in this image (no Rust, no
WebGPU runtime), so `tests/golden/make_reference_vectors.py` feeds the unmodified `.wgsl` files of /root/reference — composed
the way the reference's Rust side composes them (naga_oil `#import` / `#ifdef`, `redirect_function`, textual substitution) —
to this interpreter and commits the outputs as golden vectors.  The shader text is read at generation time and never copied
into this repository; nothing here is imported by the product package.

Scope: the WGSL subset those shaders use — structs, functions, `ptr<function, T>` parameters, private / workgroup / storage /
uniform variables, `for` / `while` / `loop` / `if` / `switch`, scalar / vector / matrix arithmetic, swizzles, atomics, the
numeric built-ins, `workgroupBarrier` — plus the naga_oil directives `#define_import_path`, `#import A as B`, `#ifdef` /
`#ifndef` / `#else` / `#endif`.

Arithmetic model (what a WebGPU backend may legally do differs; this is ONE legal evaluation, the one oracle/*.c also fixes):
  * every f32 operation is an IEEE binary32 operation (numpy float32 scalars), evaluated in source order, never contracted
    into an FMA; `fma()` itself is a correctly rounded fused multiply-add (libm fmaf);
  * `dot(a, b)` = ((a0 b0 + a1 b1) + a2 b2) + a3 b3; `M * v` = ((c0 v0 + c1 v1) + c2 v2) + c3 v3 over the columns of M;
    `A * B` column j = `A * B[j]`; `v * M` component j = dot(v, M[j]); `length(v)` = sqrt(dot(v, v));
  * u32 / i32 arithmetic wraps.
Storage and uniform buffers are byte arrays with WGSL's memory layout (vec3 aligned to 16, matCx3 columns padded to 16, array
stride = size rounded up to alignment), so the bytes in and out are the bytes the reference's Rust side uploads and reads.

Invocations of a workgroup run as coroutines that advance from barrier to barrier; functions that (transitively) contain a
barrier are evaluated through generators, everything else through plain recursion.
"""
from __future__ import annotations

import ctypes
import math
import re

import numpy as np

F32, U32, I32, BOOL = np.float32, np.uint32, np.int32, np.bool_
_libm = ctypes.CDLL("libm.so.6")
_libm.fmaf.restype = ctypes.c_float
_libm.fmaf.argtypes = [ctypes.c_float] * 3


class WgslError(Exception):
    pass


# ------------------------------------------------------------------------------------------------ preprocessor (naga_oil subset)
def preprocess(src: str, defs=()):
    """Returns (import_path or None, {alias: module path}, text without directives); inactive #ifdef branches are blanked so
    that line numbers survive."""
    out, imports, path = [], {}, None
    stack = []  # (parent_active, this_branch_taken)
    active = True
    for line in src.split("\n"):
        s = line.strip()
        if s.startswith("#"):
            word = s.split()[0]
            if word in ("#ifdef", "#ifndef"):
                name = s.split()[1]
                cond = (name in defs) == (word == "#ifdef")
                stack.append((active, cond))
                active = active and cond
            elif word == "#else":
                parent, taken = stack[-1]
                active = parent and not taken
            elif word == "#endif":
                active = stack.pop()[0]
            elif not active:
                pass
            elif word == "#define_import_path":
                path = s.split()[1]
            elif word == "#import":
                m = re.match(r"#import\s+([\w:]+)(?:\s+as\s+(\w+))?\s*;?\s*$", s)
                if not m:
                    raise WgslError(f"unsupported import form: {s}")
                imports[m.group(2) or m.group(1).split("::")[-1]] = m.group(1)
            else:
                raise WgslError(f"unsupported directive: {s}")
            out.append("")
        else:
            out.append(line if active else "")
    if stack:
        raise WgslError("unterminated #ifdef")
    return path, imports, "\n".join(out)


# ------------------------------------------------------------------------------------------------ lexer
_TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<float>(?:0[xX][0-9a-fA-F]*\.?[0-9a-fA-F]*[pP][+-]?\d+[fh]?)
             |(?:\d+\.\d*(?:[eE][+-]?\d+)?[fh]?|\.\d+(?:[eE][+-]?\d+)?[fh]?|\d+[eE][+-]?\d+[fh]?|\d+[fh]))
  | (?P<int>0[xX][0-9a-fA-F]+[iu]?|\d+[iu]?)
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op><<=|>>=|\+\+|--|->|&&|\|\||==|!=|<=|>=|<<|>>|\+=|-=|\*=|/=|%=|&=|\|=|\^=|::|[-+*/%&|^~!<>=(){}\[\],;:.@])
""", re.X | re.S)


def tokenize(src):
    toks, pos, line = [], 0, 1
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise WgslError(f"line {line}: cannot tokenize {src[pos:pos + 20]!r}")
        kind = m.lastgroup
        text = m.group()
        if kind != "ws":
            toks.append((kind, text, line))
        line += text.count("\n")
        pos = m.end()
    toks.append(("eof", "", line))
    return toks


# ------------------------------------------------------------------------------------------------ types
class Ty:
    """kind: 'scalar' (name f32/u32/i32/bool), 'vec' (n, elem), 'mat' (cols, rows), 'array' (elem, n or None), 'struct'
    (name, members [(name, Ty)]), 'atomic' (elem), 'ptr' (elem)."""

    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)

    def __repr__(self):
        return f"Ty({self.kind}, {', '.join(f'{k}={v}' for k, v in self.__dict__.items() if k != 'kind')})"


_SCALARS = {"f32": F32, "u32": U32, "i32": I32, "bool": BOOL}
T_F32, T_U32, T_I32, T_BOOL = (Ty("scalar", name=n) for n in ("f32", "u32", "i32", "bool"))
_SC = {"f32": T_F32, "u32": T_U32, "i32": T_I32, "bool": T_BOOL}


def _round_up(a, x):
    return (x + a - 1) // a * a


def layout(t: Ty):
    """(size, align) in bytes, WGSL §14.4.1"""
    k = t.kind
    if k in ("scalar", "atomic"):
        return 4, 4
    if k == "vec":
        return {2: (8, 8), 3: (12, 16), 4: (16, 16)}[t.n]
    if k == "mat":
        cs, ca = layout(Ty("vec", n=t.rows, elem=T_F32))
        return t.cols * _round_up(ca, cs), ca
    if k == "array":
        es, ea = layout(t.elem)
        stride = _round_up(ea, es)
        return (stride * (t.n or 0)), ea
    if k == "struct":
        off, al = 0, 1
        for _, mt in t.members:
            ms, ma = layout(mt)
            off = _round_up(ma, off) + ms
            al = max(al, ma)
        return _round_up(al, off), al
    raise WgslError(f"no layout for {t}")


def member_offset(t: Ty, name):
    off = 0
    for mn, mt in t.members:
        ms, ma = layout(mt)
        off = _round_up(ma, off)
        if mn == name:
            return off, mt
        off += ms
    raise WgslError(f"no member {name} in struct {t.name}")


class StructVal:
    __slots__ = ("ty", "f")

    def __init__(self, ty, f):
        self.ty, self.f = ty, f

    def __repr__(self):
        return f"{self.ty.name}({self.f})"


def zero_value(t: Ty):
    k = t.kind
    if k == "scalar":
        return _SCALARS[t.name](0)
    if k == "atomic":
        return zero_value(t.elem)
    if k == "vec":
        return np.zeros(t.n, _SCALARS[t.elem.name])
    if k == "mat":
        return np.zeros((t.cols, t.rows), F32)
    if k == "array":
        return [zero_value(t.elem) for _ in range(t.n)]
    if k == "struct":
        return StructVal(t, {n: zero_value(mt) for n, mt in t.members})
    raise WgslError(f"no zero value for {t}")


def load(t: Ty, buf: np.ndarray, off: int):
    k = t.kind
    if k == "scalar":
        return buf[off:off + 4].view(_SCALARS[t.name] if t.name != "bool" else U32)[0]
    if k == "atomic":
        return load(t.elem, buf, off)
    if k == "vec":
        return buf[off:off + 4 * t.n].view(_SCALARS[t.elem.name]).copy()
    if k == "mat":
        cstride = _round_up(layout(Ty("vec", n=t.rows, elem=T_F32))[1], 4 * t.rows)
        return np.stack([buf[off + c * cstride:off + c * cstride + 4 * t.rows].view(F32) for c in range(t.cols)]).copy()
    if k == "array":
        es, ea = layout(t.elem)
        stride = _round_up(ea, es)
        return [load(t.elem, buf, off + i * stride) for i in range(t.n)]
    if k == "struct":
        return StructVal(t, {n: load(mt, buf, off + member_offset(t, n)[0]) for n, mt in t.members})
    raise WgslError(f"cannot load {t}")


def store(t: Ty, buf: np.ndarray, off: int, v):
    k = t.kind
    if k == "scalar":
        buf[off:off + 4].view(_SCALARS[t.name])[0] = v
    elif k == "atomic":
        store(t.elem, buf, off, v)
    elif k == "vec":
        buf[off:off + 4 * t.n].view(_SCALARS[t.elem.name])[:] = v
    elif k == "mat":
        cstride = _round_up(layout(Ty("vec", n=t.rows, elem=T_F32))[1], 4 * t.rows)
        for c in range(t.cols):
            buf[off + c * cstride:off + c * cstride + 4 * t.rows].view(F32)[:] = v[c]
    elif k == "array":
        es, ea = layout(t.elem)
        stride = _round_up(ea, es)
        for i in range(t.n):
            store(t.elem, buf, off + i * stride, v[i])
    elif k == "struct":
        for n, mt in t.members:
            store(mt, buf, off + member_offset(t, n)[0], v.f[n])
    else:
        raise WgslError(f"cannot store {t}")


# ------------------------------------------------------------------------------------------------ references (lvalues, pointers)
class Cell:
    """A function / private / workgroup variable."""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v

    def get(self):
        return self.v

    def set(self, v):
        self.v = v


class SubRef:
    """Component / element / member of another reference; values are immutable, a write rebuilds the parent."""
    __slots__ = ("p", "k")

    def __init__(self, p, k):
        self.p, self.k = p, k

    def get(self):
        v = self.p.get()
        return v.f[self.k] if isinstance(v, StructVal) else v[self.k]

    def set(self, x):
        v = self.p.get()
        if isinstance(v, StructVal):
            f = dict(v.f)
            f[self.k] = x
            self.p.set(StructVal(v.ty, f))
        elif isinstance(v, list):
            v = list(v)
            v[self.k] = x
            self.p.set(v)
        else:
            v = v.copy()
            v[self.k] = x
            self.p.set(v)


class _NullRef:
    """target of an out-of-bounds access under robust buffer access: reads as zero, ignores writes"""
    __slots__ = ("ty",)

    def __init__(self, ty):
        self.ty = ty

    def get(self):
        return zero_value(self.ty)

    def set(self, v):
        pass

    def index(self, i):
        return _NullRef(self.ty.elem if self.ty.kind == "array" else T_F32)

    def member(self, name):
        return _NullRef(member_offset(self.ty, name)[1] if self.ty.kind == "struct" else self.ty.elem)


class MemRef:
    """A typed location in a storage / uniform buffer."""
    __slots__ = ("buf", "off", "ty")
    robust = None            # None: out-of-bounds accesses raise; [count]: robust buffer access (set by Program.dispatch)

    def __init__(self, buf, off, ty):
        self.buf, self.off, self.ty = buf, off, ty

    def get(self):
        return load(self.ty, self.buf, self.off)

    def set(self, v):
        store(self.ty, self.buf, self.off, v)

    def index(self, i):
        t = self.ty
        i = int(i)
        if t.kind == "array":
            es, ea = layout(t.elem)
            stride = _round_up(ea, es)
            if t.n is None:
                n = (len(self.buf) - self.off) // stride
            else:
                n = t.n
            if not 0 <= i < n:
                # WebGPU makes out-of-bounds accesses safe but leaves their result to the implementation; a shader whose
                # OUTPUT depends on one has no defined result, so by default this is an error.  Program(robust=True) applies
                # one of the allowed behaviours instead (reads return zero, writes are dropped) and counts the accesses.
                if MemRef.robust is None:
                    raise WgslError(f"out-of-bounds storage access: index {i} of {n}")
                MemRef.robust[0] += 1
                return _NullRef(t.elem)
            return MemRef(self.buf, self.off + i * stride, t.elem)
        if t.kind == "vec":
            return MemRef(self.buf, self.off + 4 * i, t.elem)
        if t.kind == "mat":
            cstride = _round_up(layout(Ty("vec", n=t.rows, elem=T_F32))[1], 4 * t.rows)
            return MemRef(self.buf, self.off + i * cstride, Ty("vec", n=t.rows, elem=T_F32))
        raise WgslError(f"cannot index {t}")

    def member(self, name):
        t = self.ty
        if t.kind == "struct":
            off, mt = member_offset(t, name)
            return MemRef(self.buf, self.off + off, mt)
        if t.kind == "vec" and len(name) == 1:
            return self.index("xyzw".index(name) if name in "xyzw" else "rgba".index(name))
        raise WgslError(f"cannot take member {name} of {t}")


_REFS = (Cell, SubRef, MemRef, _NullRef)


# ------------------------------------------------------------------------------------------------ parser
_TYPE_GENERATORS = {"vec2", "vec3", "vec4", "array", "ptr", "atomic", "bitcast"} | {f"mat{c}x{r}" for c in (2, 3, 4) for r in (2, 3, 4)}
_SHORT = {}
for _n in (2, 3, 4):
    for _s, _t in (("f", "f32"), ("u", "u32"), ("i", "i32")):
        _SHORT[f"vec{_n}{_s}"] = ("vec", _n, _t)
for _c in (2, 3, 4):
    for _r in (2, 3, 4):
        _SHORT[f"mat{_c}x{_r}f"] = ("mat", _c, _r)
_ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>="}
_BINARY_LEVELS = [("||",), ("&&",), ("|",), ("^",), ("&",), ("==", "!="), ("<", ">", "<=", ">="), ("<<", ">>"), ("+", "-"), ("*", "/", "%")]


class Node:
    __slots__ = ("k", "a", "line", "barrier")

    def __init__(self, k, line, *a):
        self.k, self.a, self.line, self.barrier = k, a, line, False

    def __repr__(self):
        return f"{self.k}{self.a}"


class Parser:
    def __init__(self, src):
        self.t = tokenize(src)
        self.i = 0

    # -- token helpers
    def peek(self, o=0):
        return self.t[self.i + o]

    def at(self, text):
        return self.t[self.i][1] == text and self.t[self.i][0] in ("op", "id")

    def accept(self, text):
        if self.at(text):
            self.i += 1
            return True
        return False

    def expect(self, text):
        if not self.accept(text):
            k, tx, ln = self.peek()
            raise WgslError(f"line {ln}: expected {text!r}, found {tx!r}")

    def ident(self):
        k, tx, ln = self.peek()
        if k != "id":
            raise WgslError(f"line {ln}: expected identifier, found {tx!r}")
        self.i += 1
        return tx

    def path(self):
        name = self.ident()
        while self.at("::"):
            self.i += 1
            name += "::" + self.ident()
        return name

    def attributes(self):
        attrs = {}
        while self.accept("@"):
            name = self.ident()
            args = []
            if self.accept("("):
                while not self.at(")"):
                    args.append(self.expr())
                    if not self.accept(","):
                        break
                self.expect(")")
            attrs[name] = args
        return attrs

    # -- types: parsed to a syntax tuple, resolved by the module (struct names, aliases, constants in array sizes)
    def type_(self):
        ln = self.peek()[2]
        name = self.path()
        args = []
        if self.at("<") and (name in _TYPE_GENERATORS):
            self.i += 1
            while not self.at(">"):
                # array<T, N>: N is an expression; others are types or address-space / access keywords
                if name == "array" and args:
                    args.append(("expr", self.expr(no_gt=True)))
                else:
                    args.append(("type", self.type_()))
                if not self.accept(","):
                    break
            self._close_template()
        return ("ty", name, args, ln)

    def _close_template(self):
        # '>>' may close two template lists
        k, tx, ln = self.peek()
        if tx == ">":
            self.i += 1
        elif tx == ">>":
            self.t[self.i] = ("op", ">", ln)
        elif tx == ">=":
            self.t[self.i] = ("op", "=", ln)
        else:
            raise WgslError(f"line {ln}: expected '>' closing a template list, found {tx!r}")

    # -- module level
    def module(self):
        decls = []
        while self.peek()[0] != "eof":
            if self.accept(";"):
                continue
            attrs = self.attributes()
            k, tx, ln = self.peek()
            if tx == "struct":
                self.i += 1
                name = self.ident()
                self.expect("{")
                members = []
                while not self.at("}"):
                    self.attributes()
                    mn = self.ident()
                    self.expect(":")
                    members.append((mn, self.type_()))
                    if not self.accept(","):
                        break
                self.expect("}")
                decls.append(Node("struct", ln, name, members))
            elif tx == "var":
                self.i += 1
                space = "private"
                if self.accept("<"):
                    space = self.ident()
                    if self.accept(","):
                        self.ident()
                    self.expect(">")
                name = self.ident()
                ty = None
                if self.accept(":"):
                    ty = self.type_()
                init = self.expr() if self.accept("=") else None
                self.expect(";")
                decls.append(Node("gvar", ln, name, space, ty, init, attrs))
            elif tx in ("const", "override"):
                self.i += 1
                name = self.ident()
                ty = self.type_() if self.accept(":") else None
                self.expect("=")
                init = self.expr()
                self.expect(";")
                decls.append(Node("gconst", ln, name, ty, init))
            elif tx == "alias":
                self.i += 1
                name = self.ident()
                self.expect("=")
                ty = self.type_()
                self.expect(";")
                decls.append(Node("alias", ln, name, ty))
            elif tx == "fn":
                self.i += 1
                name = self.ident()
                self.expect("(")
                params = []
                while not self.at(")"):
                    pattrs = self.attributes()
                    pn = self.ident()
                    self.expect(":")
                    params.append((pn, self.type_(), pattrs))
                    if not self.accept(","):
                        break
                self.expect(")")
                ret = None
                if self.accept("->"):
                    self.attributes()
                    ret = self.type_()
                body = self.block()
                decls.append(Node("fn", ln, name, params, ret, body, attrs))
            elif tx in ("enable", "requires", "diagnostic"):
                while not self.accept(";"):
                    self.i += 1
            else:
                raise WgslError(f"line {ln}: unexpected {tx!r} at module level")
        return decls

    # -- statements
    def block(self):
        ln = self.peek()[2]
        self.expect("{")
        stmts = []
        while not self.at("}"):
            s = self.statement()
            if s is not None:
                stmts.append(s)
        self.expect("}")
        return Node("block", ln, stmts)

    def simple_statement(self):
        """let / var / const / assignment / increment / call — without the trailing ';' (also used in `for` headers)."""
        k, tx, ln = self.peek()
        if tx in ("let", "var", "const"):
            self.i += 1
            if tx == "var" and self.accept("<"):
                self.ident()
                self.expect(">")
            name = self.ident()
            ty = self.type_() if self.accept(":") else None
            init = self.expr() if self.accept("=") else None
            return Node("decl", ln, tx, name, ty, init)
        if tx == "_":
            self.i += 1
            self.expect("=")
            return Node("expr", ln, self.expr())
        lhs = self.unary()
        k2, tx2, _ = self.peek()
        if tx2 in _ASSIGN_OPS:
            self.i += 1
            return Node("assign", ln, tx2, lhs, self.expr())
        if tx2 in ("++", "--"):
            self.i += 1
            return Node("assign", ln, "+=" if tx2 == "++" else "-=", lhs, Node("num", ln, 1))
        if lhs.k != "call":
            raise WgslError(f"line {ln}: expression statement must be a call")
        return Node("expr", ln, lhs)

    def statement(self):
        k, tx, ln = self.peek()
        if tx == ";":
            self.i += 1
            return None
        if tx == "{":
            return self.block()
        if tx == "if":
            self.i += 1
            cond = self.expr()
            then = self.block()
            other = None
            if self.accept("else"):
                other = self.statement() if self.at("if") else self.block()
            return Node("if", ln, cond, then, other)
        if tx == "for":
            self.i += 1
            self.expect("(")
            init = None if self.at(";") else self.simple_statement()
            self.expect(";")
            cond = None if self.at(";") else self.expr()
            self.expect(";")
            upd = None if self.at(")") else self.simple_statement()
            self.expect(")")
            return Node("for", ln, init, cond, upd, self.block())
        if tx == "while":
            self.i += 1
            cond = self.expr()
            return Node("for", ln, None, cond, None, self.block())
        if tx == "loop":
            self.i += 1
            self.expect("{")
            stmts, cont = [], None
            while not self.at("}"):
                if self.accept("continuing"):
                    cont = self.block()
                else:
                    s = self.statement()
                    if s is not None:
                        stmts.append(s)
            self.expect("}")
            return Node("loop", ln, Node("block", ln, stmts), cont)
        if tx == "switch":
            self.i += 1
            sel = self.expr()
            self.expect("{")
            cases = []
            while not self.at("}"):
                if self.accept("default"):
                    self.accept(":")
                    cases.append((None, self.block()))
                else:
                    self.expect("case")
                    vals = []
                    while True:
                        vals.append(None if self.accept("default") else self.expr())
                        if not self.accept(","):
                            break
                    self.accept(":")
                    cases.append((vals, self.block()))
            self.expect("}")
            return Node("switch", ln, sel, cases)
        if tx in ("break", "continue", "discard"):
            self.i += 1
            if tx == "break" and self.accept("if"):       # `break if cond;` in a continuing block
                cond = self.expr()
                self.expect(";")
                return Node("if", ln, cond, Node("block", ln, [Node("break", ln)]), None)
            self.expect(";")
            return Node(tx, ln)
        if tx == "return":
            self.i += 1
            e = None if self.at(";") else self.expr()
            self.expect(";")
            return Node("return", ln, e)
        s = self.simple_statement()
        self.expect(";")
        return s

    # -- expressions
    def expr(self, level=0, no_gt=False):
        if level == len(_BINARY_LEVELS):
            return self.unary()
        ops = _BINARY_LEVELS[level]
        lhs = self.expr(level + 1, no_gt)
        while True:
            k, tx, ln = self.peek()
            if k == "op" and tx in ops and not (no_gt and tx in (">", ">>", ">=")):
                self.i += 1
                lhs = Node("bin", ln, tx, lhs, self.expr(level + 1, no_gt))
            else:
                return lhs

    def unary(self):
        k, tx, ln = self.peek()
        if k == "op" and tx in ("-", "!", "~", "*", "&"):
            self.i += 1
            return Node("un", ln, tx, self.unary())
        return self.postfix(self.primary())

    def postfix(self, e):
        while True:
            k, tx, ln = self.peek()
            if tx == "[":
                self.i += 1
                idx = self.expr()
                self.expect("]")
                e = Node("index", ln, e, idx)
            elif tx == ".":
                self.i += 1
                e = Node("member", ln, e, self.ident())
            else:
                return e

    def primary(self):
        k, tx, ln = self.peek()
        if k == "int":
            self.i += 1
            if tx[-1] == "u":
                return Node("num", ln, U32(int(tx[:-1], 0)))
            if tx[-1] == "i":
                return Node("num", ln, I32(int(tx[:-1], 0)))
            return Node("num", ln, int(tx, 0))
        if k == "float":
            self.i += 1
            if tx.lower().startswith("0x"):
                v = float.fromhex(tx.rstrip("fh"))
                return Node("num", ln, F32(v) if tx[-1] == "f" else v)
            if tx[-1] == "f":
                return Node("num", ln, F32(tx[:-1]))
            return Node("num", ln, float(tx.rstrip("h")))
        if tx == "(":
            self.i += 1
            e = self.expr()
            self.expect(")")
            return e
        if tx in ("true", "false"):
            self.i += 1
            return Node("num", ln, tx == "true")
        if k == "id":
            # a type constructor / generator with a template list, a call, or a name
            save = self.i
            name = self.path()
            if self.at("<") and name in _TYPE_GENERATORS:
                self.i = save
                ty = self.type_()
                self.expect("(")
                args = self.call_args()
                return Node("call", ln, ty, args)
            if self.accept("("):
                return Node("call", ln, name, self.call_args())
            return Node("name", ln, name)
        raise WgslError(f"line {ln}: unexpected {tx!r} in expression")

    def call_args(self):
        args = []
        while not self.at(")"):
            args.append(self.expr())
            if not self.accept(","):
                break
        self.expect(")")
        return args


# ------------------------------------------------------------------------------------------------ numeric built-ins
def _conc(x):
    """abstract numeric -> concrete default type"""
    if isinstance(x, bool):
        return BOOL(x)
    if isinstance(x, int):
        return I32(x)
    if isinstance(x, float):
        return F32(x)
    return x


def _is_float(x):
    return isinstance(x, (float, F32)) or (isinstance(x, np.ndarray) and x.dtype == F32)


def _dot(a, b):
    acc = a[0] * b[0]
    for i in range(1, len(a)):
        acc = acc + a[i] * b[i]
    return acc


def _mat_vec(m, v):
    acc = m[0] * v[0]
    for c in range(1, m.shape[0]):
        acc = acc + m[c] * v[c]
    return acc


def _mat_mat(a, b):
    return np.stack([_mat_vec(a, b[j]) for j in range(b.shape[0])])


def _vec_mat(v, m):
    return np.array([_dot(v, m[j]) for j in range(m.shape[0])], F32)


def _fma(a, b, c):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) or isinstance(c, np.ndarray):
        n = max(len(x) for x in (a, b, c) if isinstance(x, np.ndarray))
        pick = lambda x, i: x[i] if isinstance(x, np.ndarray) else x      # noqa: E731
        return np.array([_fma(pick(a, i), pick(b, i), pick(c, i)) for i in range(n)], F32)
    return F32(_libm.fmaf(float(a), float(b), float(c)))


def _map(fn):
    def g(*xs):
        if any(isinstance(x, np.ndarray) for x in xs):
            n = max(len(x) for x in xs if isinstance(x, np.ndarray))
            return np.array([fn(*[(x[i] if isinstance(x, np.ndarray) else x) for x in xs]) for i in range(n)])
        return fn(*xs)
    return g


def _f(x):
    return F32(x) if not isinstance(x, np.ndarray) else x.astype(F32)


def _sign(x):
    x = _conc(x)
    if _is_float(x):
        return F32(0.0) if x == 0 else (F32(1.0) if x > 0 else (F32(-1.0) if x < 0 else x))
    return type(x)(0 if x == 0 else (1 if x > 0 else -1))


def _min(a, b):
    """WGSL min / max: 'if one operand is a NaN, the other is returned' (the IEEE minNum / maxNum behaviour of C's fminf)"""
    a, b = _unify(a, b)
    if a != a:
        return b
    if b != b:
        return a
    return b if b < a else a


def _max(a, b):
    a, b = _unify(a, b)
    if a != a:
        return b
    if b != b:
        return a
    return b if a < b else a


def _unify(a, b):
    """abstract operands take the type of the concrete one"""
    ta = a.dtype.type if isinstance(a, (np.ndarray, np.generic)) else None
    tb = b.dtype.type if isinstance(b, (np.ndarray, np.generic)) else None
    if ta is None and tb is None:
        return a, b
    if ta is None:
        return tb(a), b
    if tb is None:
        return a, ta(b)
    return a, b


def _select(f, t, c):
    f, t = _unify(f, t)
    if isinstance(c, np.ndarray):
        return np.where(c, t, f)
    return t if c else f


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], F32)


def _libm1(name):
    fn = getattr(_libm, name)
    fn.restype = ctypes.c_float
    fn.argtypes = [ctypes.c_float]
    return _map(lambda x: F32(fn(float(_conc(x)))))


def _libm2(name):
    fn = getattr(_libm, name)
    fn.restype = ctypes.c_float
    fn.argtypes = [ctypes.c_float, ctypes.c_float]
    return _map(lambda x, y: F32(fn(float(_conc(x)), float(_conc(y)))))


def _sqrt(x):
    return np.sqrt(_f(_conc(x)))


def _det(m):
    n = m.shape[0]
    if n == 2:
        return m[0][0] * m[1][1] - m[0][1] * m[1][0]
    raise WgslError("determinant: only 2x2 is defined here (backends differ on the evaluation order of larger ones)")


_BUILTINS = {
    "abs": lambda x: np.abs(_conc(x)),
    "sqrt": _sqrt,
    "inverseSqrt": lambda x: F32(1.0) / _sqrt(x),
    "sign": _map(_sign),
    "min": _map(_min), "max": _map(_max),
    "clamp": _map(lambda x, lo, hi: _min(_max(x, lo), hi)),
    "select": _select,
    "fma": _fma,
    "dot": _dot,
    "cross": _cross,
    "length": lambda v: _sqrt(_dot(v, v)) if isinstance(v, np.ndarray) else np.abs(_f(v)),
    "normalize": lambda v: v / _sqrt(_dot(v, v)),
    "transpose": lambda m: np.ascontiguousarray(m.T),
    "determinant": _det,
    # transcendental functions: WGSL gives them an error bound, not a value; glibc's single-precision routines are used (the
    # ones the C oracle links), so that results through them can still be compared bit for bit
    "sin": _libm1("sinf"), "cos": _libm1("cosf"), "tan": _libm1("tanf"), "atan": _libm1("atanf"), "atan2": _libm2("atan2f"),
    "asin": _libm1("asinf"), "acos": _libm1("acosf"), "exp": _libm1("expf"), "exp2": _libm1("exp2f"), "log": _libm1("logf"),
    "tanh": _libm1("tanhf"),
    "floor": lambda x: np.floor(_f(_conc(x))), "ceil": lambda x: np.ceil(_f(_conc(x))), "trunc": lambda x: np.trunc(_f(_conc(x))),
    "round": lambda x: np.rint(_f(_conc(x))), "fract": lambda x: _f(x) - np.floor(_f(x)),
    "pow": _libm2("powf"),
    "all": lambda x: BOOL(np.all(x)), "any": lambda x: BOOL(np.any(x)),
    "countOneBits": _map(lambda x: type(x)(bin(int(x) & 0xFFFFFFFF).count("1"))),
    "firstLeadingBit": _map(lambda x: U32(0xFFFFFFFF) if int(x) == 0 else U32(int(x).bit_length() - 1)),
    "firstTrailingBit": _map(lambda x: U32(0xFFFFFFFF) if int(x) == 0 else U32((int(x) & -int(x)).bit_length() - 1)),
    "countLeadingZeros": _map(lambda x: U32(32 - int(x).bit_length())),
    "arrayLength": None,  # handled in the evaluator (needs the reference)
}


def _convert(name, x):
    """scalar conversion T(x) with WGSL semantics (float -> int truncates and saturates)"""
    if isinstance(x, np.ndarray):
        return np.array([_convert(name, e) for e in x], _SCALARS[name])
    if name == "f32":
        return F32(x)
    if name == "bool":
        return BOOL(x != 0)
    if _is_float(x):
        lo, hi = (0, 0xFFFFFFFF) if name == "u32" else (-0x80000000, 0x7FFFFFFF)
        fx = float(x)
        iv = 0 if math.isnan(fx) else (hi if fx >= hi else (lo if fx <= lo else int(fx)))
        return _SCALARS[name](iv)
    iv = int(x) & 0xFFFFFFFF
    if name == "i32" and iv >= 0x80000000:
        iv -= 1 << 32
    return _SCALARS[name](iv)


def _binary(op, a, b, line):
    if op == "&&":
        return BOOL(bool(a) and bool(b))
    if op == "||":
        return BOOL(bool(a) or bool(b))
    am, bm = isinstance(a, np.ndarray) and a.ndim == 2, isinstance(b, np.ndarray) and b.ndim == 2
    if op == "*" and (am or bm):
        if am and bm:
            return _mat_mat(a, b)
        if am and isinstance(b, np.ndarray):
            return _mat_vec(a, b)
        if bm and isinstance(a, np.ndarray):
            return _vec_mat(a, b)
    if op in ("<<", ">>"):
        sh = int(b) & 31 if not isinstance(b, np.ndarray) else (b.astype(np.int64) & 31)
        if isinstance(a, int) and not isinstance(a, bool):
            return a << sh if op == "<<" else a >> sh
        t = a.dtype.type
        if op == "<<":
            return t((int(a) << sh) & 0xFFFFFFFF) if t is U32 else _convert("i32", int(a) << sh)
        return t(int(a) >> sh)
    a, b = _unify(a, b)
    da, db = getattr(a, "dtype", None), getattr(b, "dtype", None)
    if da is not None and db is not None and da != db and BOOL not in (da.type, db.type):
        raise WgslError(f"line {line}: operands of {op} have different types ({da}, {db})")
    if op == "+":
        return a + b
    if op == "-":
        return a - b
    if op == "*":
        return a * b
    if op == "/":
        if _is_float(a) or _is_float(b):
            return a / b
        if isinstance(a, int) and isinstance(b, int):
            return int(a / b) if b else 0
        if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
            return (a // np.where(b == 0, 1, b)).astype(a.dtype if isinstance(a, np.ndarray) else b.dtype)
        if b == 0:
            return a                                           # WGSL: x / 0 = x for integers
        if a.dtype.type is I32:
            return I32(int(math.trunc(int(a) / int(b))))
        return a // b
    if op == "%":
        if _is_float(a) or _is_float(b):
            return np.fmod(a, b)
        if isinstance(a, int) and isinstance(b, int):
            return int(math.fmod(a, b)) if b else 0
        if not isinstance(a, np.ndarray) and not isinstance(b, np.ndarray) and b == 0:
            return type(a)(0)
        if isinstance(a, np.generic) and a.dtype.type is I32:
            return I32(int(math.fmod(int(a), int(b))))
        return a % b
    if op == "==":
        return a == b
    if op == "!=":
        return a != b
    if op == "<":
        return a < b
    if op == ">":
        return a > b
    if op == "<=":
        return a <= b
    if op == ">=":
        return a >= b
    if op == "&":
        return (a & b) if not isinstance(a, (bool, np.bool_)) else BOOL(bool(a) and bool(b))
    if op == "|":
        return (a | b) if not isinstance(a, (bool, np.bool_)) else BOOL(bool(a) or bool(b))
    if op == "^":
        return a ^ b
    raise WgslError(f"line {line}: unsupported operator {op}")


# ------------------------------------------------------------------------------------------------ modules and evaluation
class _Return(Exception):
    def __init__(self, v):
        self.v = v


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


_BARRIERS = {"workgroupBarrier", "storageBarrier", "textureBarrier"}
_SWZ = {c: i for s in ("xyzw", "rgba") for i, c in enumerate(s)}


class Module:
    def __init__(self, program, path, imports, decls):
        self.program, self.path, self.imports = program, path, imports
        self.structs, self.aliases, self.consts, self.fns, self.gvars = {}, {}, {}, {}, {}
        self.redirect = {}
        for d in decls:
            if d.k == "struct":
                self.structs[d.a[0]] = d
            elif d.k == "alias":
                self.aliases[d.a[0]] = d.a[1]
            elif d.k == "gconst":
                self.consts[d.a[0]] = d
            elif d.k == "fn":
                self.fns[d.a[0]] = d
            elif d.k == "gvar":
                self.gvars[d.a[0]] = d
        self._types = {}
        self._const_vals = {}

    # name resolution: 'Alias::item' goes to the imported module
    def split(self, name):
        if "::" in name:
            head, item = name.rsplit("::", 1)
            target = self.imports.get(head, head)
            if target not in self.program.modules:
                raise WgslError(f"module {target!r} (imported as {head!r}) is not loaded")
            return self.program.modules[target], item
        return self, name

    def resolve_type(self, syn, env=None) -> Ty:
        _, name, args, ln = syn
        mod, item = self.split(name)
        if mod is not self:
            return mod.resolve_type(("ty", item, args, ln))
        if name in _SC:
            return _SC[name]
        if name in _SHORT:
            s = _SHORT[name]
            return Ty("vec", n=s[1], elem=_SC[s[2]]) if s[0] == "vec" else Ty("mat", cols=s[1], rows=s[2])
        if name in ("vec2", "vec3", "vec4"):
            elem = self.resolve_type(args[0][1]) if args else None
            return Ty("vec", n=int(name[3]), elem=elem)
        if name.startswith("mat") and name in _TYPE_GENERATORS:
            return Ty("mat", cols=int(name[3]), rows=int(name[5]))
        if name == "array":
            if not args:
                return Ty("array", elem=None, n=None)
            elem = self.resolve_type(args[0][1])
            n = None
            if len(args) > 1:
                n = int(Interp(self.program, None).eval(args[1][1], _Env(self, None)))
            return Ty("array", elem=elem, n=n)
        if name == "atomic":
            return Ty("atomic", elem=self.resolve_type(args[0][1]))
        if name == "ptr":
            return Ty("ptr", elem=self.resolve_type(args[1][1]))
        if name in self.aliases:
            return self.resolve_type(self.aliases[name])
        if name in self.structs:
            if name not in self._types:
                d = self.structs[name]
                t = Ty("struct", name=name, members=[])
                self._types[name] = t
                t.members = [(mn, self.resolve_type(mt)) for mn, mt in d.a[1]]
            return self._types[name]
        raise WgslError(f"line {ln}: unknown type {name!r}")

    def function(self, name):
        mod, item = self.split(name)
        item = mod.redirect.get(item, item)
        return mod, mod.fns.get(item)


class Program:
    """A set of composed modules.  `add_module(src, defs)` registers an importable module (it must carry
    #define_import_path); `set_main(src, defs)` the module holding the entry points."""

    def __init__(self, robust=False):
        self.modules = {}
        self.main = None
        self.robust = robust
        self.oob_accesses = 0

    def _make(self, src, defs):
        path, imports, text = preprocess(src, defs)
        return Module(self, path, imports, Parser(text).module())

    def add_module(self, src, defs=()):
        m = self._make(src, defs)
        if m.path is None:
            raise WgslError("importable module without #define_import_path")
        self.modules[m.path] = m
        return m

    def set_main(self, src, defs=()):
        self.main = self._make(src, defs)
        self.modules["<main>"] = self.main
        return self.main

    def redirect_function(self, old, new, module=None):
        """naga_oil Redirector::redirect_function: every call of `old` calls `new` instead."""
        m = module or self.main
        if old not in m.fns or new not in m.fns:
            raise WgslError(f"redirect_function: {old!r} or {new!r} is not a function of the module")
        m.redirect[old] = new

    # -- static pass: which functions / statements contain a barrier (transitively)
    def _mark(self):
        tainted = set()
        changed = True

        def calls_tainted(node, mod):
            if isinstance(node, Node):
                if node.k == "call" and isinstance(node.a[0], str):
                    if node.a[0] in _BARRIERS:
                        return True
                    try:
                        m2, fn = mod.function(node.a[0])
                    except WgslError:
                        fn = None
                    if fn is not None and (id(fn)) in tainted:
                        return True
                return any(calls_tainted(x, mod) for x in node.a)
            if isinstance(node, (list, tuple)):
                return any(calls_tainted(x, mod) for x in node)
            return False

        while changed:
            changed = False
            for mod in self.modules.values():
                for fn in mod.fns.values():
                    if id(fn) not in tainted and calls_tainted(fn.a[3], mod):
                        tainted.add(id(fn))
                        changed = True

        def mark(node, mod):
            if isinstance(node, Node):
                hit = False
                for x in node.a:
                    hit |= mark(x, mod)
                if node.k == "call" and isinstance(node.a[0], str):
                    if node.a[0] in _BARRIERS:
                        hit = True
                    else:
                        try:
                            m2, fn = mod.function(node.a[0])
                        except WgslError:
                            fn = None
                        if fn is not None and id(fn) in tainted:
                            hit = True
                node.barrier = hit
                return hit
            if isinstance(node, (list, tuple)):
                hit = False
                for x in node:
                    hit |= mark(x, mod)
                return hit
            return False

        for mod in self.modules.values():
            for fn in mod.fns.values():
                fn.barrier = mark(fn.a[3], mod) or id(fn) in tainted

    # -- dispatch
    def dispatch(self, entry, bindings, grid):
        """Runs entry point `entry` over grid = (x, y, z) workgroups.  bindings: {(group, binding): numpy uint8 array} — storage
        buffers are modified in place."""
        self._mark()
        mod = self.main
        fn = mod.fns[entry]
        attrs = fn.a[4]
        if "compute" not in attrs:
            raise WgslError(f"{entry} is not a compute entry point")
        it = Interp(self, bindings)
        genv = _Env(mod, None)
        wg = [int(it.eval(e, genv)) for e in attrs["workgroup_size"]]
        wg += [1] * (3 - len(wg))
        grid = tuple(int(g) for g in grid) + (1,) * (3 - len(grid))
        if 0 in grid:
            return
        old = np.seterr(all="ignore")                            # wrapping integer arithmetic, inf / nan are WGSL behaviour
        MemRef.robust = [0] if self.robust else None
        try:
            for gz in range(grid[2]):
                for gy in range(grid[1]):
                    for gx in range(grid[0]):
                        it.run_workgroup(mod, fn, (gx, gy, gz), wg, grid)
        finally:
            np.seterr(**old)
            if self.robust:
                self.oob_accesses += MemRef.robust[0]
            MemRef.robust = None

    def call_function(self, name, args, bindings=None):
        """Calls a (barrier-free) module function directly with Python-side values: unit tests of library functions."""
        self._mark()
        it = Interp(self, bindings or {})
        mod, fn = self.main.function(name)
        old = np.seterr(all="ignore")
        try:
            return it.invoke(mod, fn, list(args), _Invocation({}))
        finally:
            np.seterr(**old)


class _Env:
    """lexical scopes of one function activation"""
    __slots__ = ("mod", "scopes", "inv")

    def __init__(self, mod, inv):
        self.mod, self.scopes, self.inv = mod, [{}], inv

    def push(self):
        self.scopes.append({})

    def pop(self):
        self.scopes.pop()

    def declare(self, name, v):
        self.scopes[-1][name] = v

    def lookup(self, name):
        for s in reversed(self.scopes):
            if name in s:
                return s[name]
        return None


class _Invocation:
    """per-invocation state: private variables; shared: workgroup variables"""

    def __init__(self, wg_vars):
        self.private = {}
        self.wg_vars = wg_vars


class Interp:
    def __init__(self, program, bindings):
        self.program, self.bindings = program, bindings
        self._global_refs = {}

    # ---- module-scope names
    def global_ref(self, mod, name, inv):
        d = mod.gvars.get(name)
        if d is None:
            return None
        _, space, ty, init, attrs = d.a
        key = (id(mod), name)
        if space in ("storage", "uniform"):
            if key not in self._global_refs:
                genv = _Env(mod, None)
                g = int(self.eval(attrs["group"][0], genv))
                b = int(self.eval(attrs["binding"][0], genv))
                if (g, b) not in self.bindings:
                    raise WgslError(f"no buffer bound at group {g} binding {b} ({name})")
                self._global_refs[key] = MemRef(self.bindings[(g, b)], 0, mod.resolve_type(ty))
            return self._global_refs[key]
        store_ = inv.wg_vars if space == "workgroup" else inv.private
        if key not in store_:
            genv = _Env(mod, inv)
            t = mod.resolve_type(ty) if ty is not None else None
            v = self.coerce(self.eval(init, genv), t) if init is not None else zero_value(t)
            store_[key] = Cell(v)
        return store_[key]

    def const_value(self, mod, name):
        if name not in mod._const_vals:
            d = mod.consts[name]
            genv = _Env(mod, None)
            v = self.eval(d.a[2], genv)
            if d.a[1] is not None:
                v = self.coerce(v, mod.resolve_type(d.a[1]))
            mod._const_vals[name] = v
        return mod._const_vals[name]

    @staticmethod
    def coerce(v, t):
        """abstract -> declared type"""
        if t is None:
            return _conc(v) if not isinstance(v, (list, StructVal, np.ndarray)) else v
        if t.kind == "scalar" and not isinstance(v, np.generic):
            return _SCALARS[t.name](v)
        if t.kind == "scalar" and isinstance(v, np.generic) and v.dtype.type is not _SCALARS[t.name] and t.name != "bool":
            raise WgslError(f"type mismatch: {v.dtype} value for {t.name}")
        if t.kind == "vec" and t.elem is not None and isinstance(v, np.ndarray) and v.dtype.type is not _SCALARS[t.elem.name]:
            if v.dtype.type is I32 and t.elem.name == "u32":      # vecN(0) built from abstract integers
                return v.astype(U32)
            raise WgslError(f"type mismatch: {v.dtype} vector for vec{t.n}<{t.elem.name}>")
        return v

    # ---- references
    def ref(self, e, env):
        k = e.k
        if k == "name":
            name = e.a[0]
            r = env.lookup(name) if "::" not in name else None
            if r is not None:
                if isinstance(r, _REFS):
                    return r
                if isinstance(r, _Ptr):                          # p[i] / p.x on a pointer (WGSL's dereference sugar)
                    return r.r
                raise WgslError(f"line {e.line}: {name} is not a variable (cannot be assigned or referenced)")
            mod, item = env.mod.split(name)
            g = self.global_ref(mod, item, env.inv)
            if g is None:
                raise WgslError(f"line {e.line}: unknown variable {name}")
            return g
        if k == "index":
            base = self.ref(e.a[0], env)
            i = int(self.eval(e.a[1], env))
            if isinstance(base, (MemRef, _NullRef)):
                return base.index(i)
            return SubRef(base, i)
        if k == "member":
            base = self.ref(e.a[0], env)
            name = e.a[1]
            if isinstance(base, (MemRef, _NullRef)):
                return base.member(name)
            v = base.get()
            if isinstance(v, StructVal):
                return SubRef(base, name)
            if len(name) == 1:
                return SubRef(base, _SWZ[name])
            raise WgslError(f"line {e.line}: cannot assign to swizzle .{name}")
        if k == "un" and e.a[0] == "*":
            p = self.eval(e.a[1], env)
            if not isinstance(p, _REFS):
                raise WgslError(f"line {e.line}: dereference of a non-pointer")
            return p
        raise WgslError(f"line {e.line}: not a reference expression: {e.k}")

    # ---- expressions
    def eval(self, e, env):
        k = e.k
        if k == "num":
            return e.a[0]
        if k == "name":
            name = e.a[0]
            if "::" not in name:
                r = env.lookup(name)
                if r is not None:
                    if isinstance(r, _Value):
                        return r.v
                    if isinstance(r, _Ptr):
                        return r.r
                    return r.get()
            mod, item = env.mod.split(name)
            if item in mod.consts:
                return self.const_value(mod, item)
            g = self.global_ref(mod, item, env.inv)
            if g is None:
                raise WgslError(f"line {e.line}: unknown identifier {name}")
            return g.get()
        if k == "bin":
            op = e.a[0]
            a = self.eval(e.a[1], env)
            if op == "&&" and not isinstance(a, np.ndarray):
                return BOOL(bool(a) and bool(self.eval(e.a[2], env)))
            if op == "||" and not isinstance(a, np.ndarray):
                return BOOL(bool(a) or bool(self.eval(e.a[2], env)))
            return _binary(op, a, self.eval(e.a[2], env), e.line)
        if k == "un":
            op = e.a[0]
            if op == "&":
                return self.ref(e.a[1], env)
            if op == "*":
                return self.ref(e, env).get()
            v = self.eval(e.a[1], env)
            if op == "-":
                return -v
            if op == "!":
                return np.logical_not(v) if isinstance(v, np.ndarray) else BOOL(not bool(v))
            return ~v
        if k == "index":
            base = e.a[0]
            if self._is_memory(base, env):
                return self.ref(e, env).get()
            v = self.eval(base, env)
            i = int(self.eval(e.a[1], env))
            if isinstance(v, _REFS):           # p[i]: dereference sugar
                v = v.get()
            n = len(v)
            if not 0 <= i < n:
                raise WgslError(f"line {e.line}: index {i} out of range {n}")
            return v[i]
        if k == "member":
            if self._is_memory(e.a[0], env):
                base = self.ref(e.a[0], env)
                if base.ty.kind == "struct" or (base.ty.kind == "vec" and len(e.a[1]) == 1):
                    return base.member(e.a[1]).get()
                v = base.get()
            else:
                v = self.eval(e.a[0], env)
            if isinstance(v, _REFS):
                v = v.get()
            name = e.a[1]
            if isinstance(v, StructVal):
                return v.f[name]
            if isinstance(v, np.ndarray) and v.ndim == 1:
                if len(name) == 1:
                    return v[_SWZ[name]]
                return np.array([v[_SWZ[c]] for c in name], v.dtype)
            raise WgslError(f"line {e.line}: no member {name} on {type(v).__name__}")
        if k == "call":
            return self.call(e, env)
        raise WgslError(f"line {e.line}: cannot evaluate {k}")

    def _is_memory(self, e, env):
        """does the expression name a location in a storage / uniform buffer (so that only the addressed bytes are read)?"""
        while e.k in ("index", "member"):
            e = e.a[0]
        if e.k != "name":
            return False
        name = e.a[0]
        if "::" not in name:
            r = env.lookup(name)
            if r is not None:
                return isinstance(r, MemRef) or (isinstance(r, _Ptr) and isinstance(r.r, MemRef))
        mod, item = env.mod.split(name)
        d = mod.gvars.get(item)
        return d is not None and d.a[1] in ("storage", "uniform")

    def construct(self, t: Ty, args, line):
        k = t.kind
        if k == "scalar":
            return _convert(t.name, args[0]) if args else zero_value(t)
        if k == "vec":
            comps = []
            for a in args:
                if isinstance(a, np.ndarray):
                    comps.extend(a)
                else:
                    comps.append(a)
            if t.elem is None:
                typed = [c for c in comps if isinstance(c, np.generic)]
                dt = typed[0].dtype.type if typed else (F32 if any(isinstance(c, float) for c in comps) else I32)
                if not comps:
                    dt = F32
            else:
                dt = _SCALARS[t.elem.name]
                if comps and any(isinstance(c, np.generic) and c.dtype.type is not dt for c in comps):
                    comps = [_convert(t.elem.name, c) for c in comps]
            if not comps:
                return np.zeros(t.n, dt)
            if len(comps) == 1:
                comps = comps * t.n
            if len(comps) != t.n:
                raise WgslError(f"line {line}: vec{t.n} constructor given {len(comps)} components")
            return np.array(comps, dt)
        if k == "mat":
            if not args:
                return np.zeros((t.cols, t.rows), F32)
            if len(args) == 1 and isinstance(args[0], np.ndarray) and args[0].ndim == 2:
                return args[0].astype(F32)
            if all(isinstance(a, np.ndarray) for a in args):
                if len(args) != t.cols or any(len(a) != t.rows for a in args):
                    raise WgslError(f"line {line}: mat{t.cols}x{t.rows} constructor: wrong column count / size")
                return np.stack([a.astype(F32) for a in args])
            if len(args) == t.cols * t.rows:
                return np.array([F32(a) for a in args], F32).reshape(t.cols, t.rows)
            raise WgslError(f"line {line}: unsupported matrix constructor form")
        if k == "array":
            if t.elem is not None:
                args = [self.coerce(a, t.elem) for a in args]
            else:
                args = [_conc(a) for a in args]
            if not args and t.n:
                return zero_value(t)
            return list(args)
        if k == "struct":
            if not args:
                return zero_value(t)
            if len(args) != len(t.members):
                raise WgslError(f"line {line}: {t.name} constructor given {len(args)} of {len(t.members)} members")
            return StructVal(t, {n: self.coerce(a, mt) for (n, mt), a in zip(t.members, args)})
        raise WgslError(f"line {line}: cannot construct {t}")

    def call(self, e, env):
        target, argn = e.a
        if not isinstance(target, str):                          # templated constructor: vec4<f32>(...), array<f32, 4>(...)
            if target[1] == "bitcast":
                to = env.mod.resolve_type(target[2][0][1])
                v = _conc(self.eval(argn[0], env))
                dt = _SCALARS[to.name if to.kind == "scalar" else to.elem.name]
                return np.asarray(v).view(dt)[()] if not isinstance(v, np.ndarray) else v.view(dt)
            t = env.mod.resolve_type(target)
            return self.construct(t, [self.eval(a, env) for a in argn], e.line)
        name = target
        if name in _BARRIERS:
            raise WgslError(f"line {e.line}: barrier reached through the non-generator path (internal error)")
        mod, fn = env.mod.function(name)
        if fn is not None:
            if fn.barrier:
                raise WgslError(f"line {e.line}: call of {name}, which contains a barrier, inside an expression")
            args = [self.eval(a, env) for a in argn]
            return self.invoke(mod, fn, args, env.inv)
        item = name.rsplit("::", 1)[-1]
        tmod = env.mod.split(name)[0]
        if item in tmod.structs or item in tmod.aliases or item in _SC or item in _SHORT or item in _TYPE_GENERATORS:
            t = tmod.resolve_type(("ty", item, [], e.line))
            return self.construct(t, [self.eval(a, env) for a in argn], e.line)
        if name.startswith("atomic"):
            r = self.eval(argn[0], env)
            old = r.get()
            if name == "atomicLoad":
                return old
            v = self.eval(argn[1], env)
            v = old.dtype.type(v) if not isinstance(v, np.generic) else v
            new = {"atomicStore": lambda: v, "atomicAdd": lambda: old + v, "atomicSub": lambda: old - v,
                   "atomicMax": lambda: max(old, v), "atomicMin": lambda: min(old, v), "atomicAnd": lambda: old & v,
                   "atomicOr": lambda: old | v, "atomicXor": lambda: old ^ v, "atomicExchange": lambda: v}[name]()
            r.set(new)
            return None if name == "atomicStore" else old
        if name == "arrayLength":
            r = self.eval(argn[0], env)
            es, ea = layout(r.ty.elem)
            return U32((len(r.buf) - r.off) // _round_up(ea, es))
        if name in _BUILTINS:
            return _BUILTINS[name](*[self.eval(a, env) for a in argn])
        raise WgslError(f"line {e.line}: unknown function {name}")

    def bind_params(self, mod, fn, args, inv):
        env = _Env(mod, inv)
        params = fn.a[1]
        if len(params) != len(args):
            raise WgslError(f"line {fn.line}: {fn.a[0]} takes {len(params)} arguments, {len(args)} given")
        for (pn, pt, _), a in zip(params, args):
            if isinstance(a, _REFS):
                env.declare(pn, _Ptr(a))
            else:
                env.declare(pn, _Value(self.coerce(a, mod.resolve_type(pt))))
        return env

    def invoke(self, mod, fn, args, inv):
        env = self.bind_params(mod, fn, args, inv)
        try:
            self.exec_block(fn.a[3], env)
        except _Return as r:
            return r.v
        return None

    # ---- statements (plain path)
    def exec_block(self, b, env):
        env.push()
        try:
            for s in b.a[0]:
                self.exec(s, env)
        finally:
            env.pop()

    def exec(self, s, env):
        k = s.k
        if k == "decl":
            self.declare(s, env, self.eval(s.a[3], env) if s.a[3] is not None else None)
        elif k == "assign":
            self.assign(s, env, self.eval(s.a[2], env))
        elif k == "expr":
            self.eval(s.a[0], env)
        elif k == "if":
            if bool(self.eval(s.a[0], env)):
                self.exec_block(s.a[1], env)
            elif s.a[2] is not None:
                self.exec(s.a[2], env) if s.a[2].k == "if" else self.exec_block(s.a[2], env)
        elif k == "for":
            init, cond, upd, body = s.a
            env.push()
            try:
                if init is not None:
                    self.exec(init, env)
                while cond is None or bool(self.eval(cond, env)):
                    try:
                        self.exec_block(body, env)
                    except _Break:
                        break
                    except _Continue:
                        pass
                    if upd is not None:
                        self.exec(upd, env)
            finally:
                env.pop()
        elif k == "loop":
            while True:
                try:
                    self.exec_block(s.a[0], env)
                except _Break:
                    break
                except _Continue:
                    pass
                if s.a[1] is not None:
                    try:
                        self.exec_block(s.a[1], env)
                    except _Break:
                        break
        elif k == "switch":
            body = self.switch_body(s, env)
            if body is not None:
                try:
                    self.exec_block(body, env)
                except _Break:
                    pass
        elif k == "block":
            self.exec_block(s, env)
        elif k == "return":
            raise _Return(self.eval(s.a[0], env) if s.a[0] is not None else None)
        elif k == "break":
            raise _Break()
        elif k == "continue":
            raise _Continue()
        else:
            raise WgslError(f"line {s.line}: unsupported statement {k}")

    def switch_body(self, s, env):
        sel = self.eval(s.a[0], env)
        default = None
        for vals, body in s.a[1]:
            if vals is None:
                default = body
                continue
            for v in vals:
                if v is None:
                    default = body
                elif int(self.eval(v, env)) == int(sel):
                    return body
        return default

    def declare(self, s, env, v):
        kind, name, ty, _ = s.a
        t = env.mod.resolve_type(ty) if ty is not None else None
        if v is None:
            v = zero_value(t)
        elif isinstance(v, _REFS):              # let p = &x;
            env.declare(name, _Ptr(v))
            return
        else:
            v = self.coerce(v, t)
        env.declare(name, Cell(v) if kind == "var" else _Value(v))

    def assign(self, s, env, rhs):
        op, lhs, _ = s.a
        r = self.ref(lhs, env)
        if op != "=":
            rhs = _binary(op[:-1], r.get(), rhs, s.line)
        else:
            cur = r.get() if not isinstance(r, (MemRef, _NullRef)) else None
            if isinstance(r, (MemRef, _NullRef)):
                pass
            elif isinstance(cur, np.generic) and not isinstance(rhs, (np.generic, np.ndarray)):
                rhs = cur.dtype.type(rhs)
        r.set(rhs)

    # ---- statements (generator path: only nodes whose subtree contains a barrier)
    def g_block(self, b, env):
        env.push()
        try:
            for s in b.a[0]:
                if s.barrier:
                    yield from self.g_exec(s, env)
                else:
                    self.exec(s, env)
        finally:
            env.pop()

    def g_call(self, e, env):
        """a call expression whose callee contains a barrier (or is one); generator returning the call's value"""
        name, argn = e.a
        if name in _BARRIERS:
            yield name
            return None
        mod, fn = env.mod.function(name)
        args = [self.eval(a, env) for a in argn]
        fenv = self.bind_params(mod, fn, args, env.inv)
        try:
            yield from self.g_block(fn.a[3], fenv)
        except _Return as r:
            return r.v
        return None

    def _tainted_call(self, e):
        return e is not None and e.k == "call" and e.barrier and isinstance(e.a[0], str)

    def g_exec(self, s, env):
        k = s.k
        if k == "expr" and self._tainted_call(s.a[0]):
            yield from self.g_call(s.a[0], env)
        elif k == "decl" and self._tainted_call(s.a[3]):
            v = yield from self.g_call(s.a[3], env)
            self.declare(s, env, v)
        elif k == "assign" and self._tainted_call(s.a[2]) and not s.a[1].barrier:
            v = yield from self.g_call(s.a[2], env)
            self.assign(s, env, v)
        elif k == "return" and self._tainted_call(s.a[0]):
            v = yield from self.g_call(s.a[0], env)
            raise _Return(v)
        elif k == "block":
            yield from self.g_block(s, env)
        elif k == "if":
            if s.a[0].barrier:
                raise WgslError(f"line {s.line}: barrier inside a condition")
            if bool(self.eval(s.a[0], env)):
                yield from self.g_block(s.a[1], env)
            elif s.a[2] is not None:
                if s.a[2].k == "if":
                    yield from self.g_exec(s.a[2], env) if s.a[2].barrier else self._once(s.a[2], env)
                else:
                    yield from self.g_block(s.a[2], env)
        elif k == "for":
            init, cond, upd, body = s.a
            if (init is not None and init.barrier) or (cond is not None and cond.barrier) or (upd is not None and upd.barrier):
                raise WgslError(f"line {s.line}: barrier inside a loop header")
            env.push()
            try:
                if init is not None:
                    self.exec(init, env)
                while cond is None or bool(self.eval(cond, env)):
                    try:
                        yield from self.g_block(body, env)
                    except _Break:
                        break
                    except _Continue:
                        pass
                    if upd is not None:
                        self.exec(upd, env)
            finally:
                env.pop()
        elif k == "loop":
            while True:
                try:
                    yield from self.g_block(s.a[0], env)
                except _Break:
                    break
                except _Continue:
                    pass
                if s.a[1] is not None:
                    try:
                        yield from self.g_block(s.a[1], env)
                    except _Break:
                        break
        elif k == "switch":
            body = self.switch_body(s, env)
            if body is not None:
                try:
                    yield from self.g_block(body, env)
                except _Break:
                    pass
        else:
            raise WgslError(f"line {s.line}: a barrier-containing call may only appear as a statement, an initialiser, "
                            f"the right-hand side of an assignment or a return value ({k})")

    def _once(self, s, env):
        self.exec(s, env)
        return
        yield

    # ---- one workgroup
    def run_workgroup(self, mod, fn, wid, wg, grid):
        wg_vars = {}
        gens = []
        for lz in range(wg[2]):
            for ly in range(wg[1]):
                for lx in range(wg[0]):
                    inv = _Invocation(wg_vars)
                    lid = (lx, ly, lz)
                    builtins = {
                        "workgroup_id": np.array(wid, U32),
                        "local_invocation_id": np.array(lid, U32),
                        "global_invocation_id": np.array([wid[i] * wg[i] + lid[i] for i in range(3)], U32),
                        "local_invocation_index": U32(lx + ly * wg[0] + lz * wg[0] * wg[1]),
                        "num_workgroups": np.array(grid, U32),
                    }
                    args = []
                    for pn, pt, pattrs in fn.a[1]:
                        b = pattrs["builtin"][0].a[0]
                        args.append(builtins[b])
                    env = self.bind_params(mod, fn, args, inv)
                    if fn.barrier:
                        gens.append(self._entry(fn, env))
                    else:
                        try:
                            self.exec_block(fn.a[3], env)
                        except _Return:
                            pass
        # advance every invocation from barrier to barrier
        live = gens
        while live:
            states = []
            nxt = []
            for g in live:
                try:
                    states.append(next(g))
                    nxt.append(g)
                except StopIteration:
                    states.append(None)
            if nxt and len(nxt) != len(live):
                raise WgslError("non-uniform control flow: some invocations finished while others wait at a barrier")
            live = nxt

    def _entry(self, fn, env):
        try:
            yield from self.g_block(fn.a[3], env)
        except _Return:
            pass


class _Value:
    """an immutable binding (let / const / by-value parameter)"""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v


class _Ptr:
    """a binding holding a pointer (ptr<function, T> parameter, or `let p = &x`)"""
    __slots__ = ("r",)

    def __init__(self, r):
        self.r = r
