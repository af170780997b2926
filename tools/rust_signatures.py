"""Public-surface check of the Rust shim crates without a Rust toolchain: extracts every `pub` item signature (fn / struct / enum /
trait / const / type / mod, and `impl ... for ...` headers) from the reference files the hot path cites and from their
counterparts under rust/, normalises whitespace, and diffs the two sets per file pair.  Differences that are deliberate are
listed with their reason in rust/SIGNATURES.notes (one `<signature substring> :: <reason>` per line); the report
rust/SIGNATURES.diff must contain no unannotated line.

    python tools/rust_signatures.py [--check]      # --check: exit 1 on an unannotated difference (used by tests/test_host.py)
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/crates"
PAIRS = [
    ("wgebra/src/linalg/gemm.rs", "rust/wgebra/src/linalg/gemm.rs"),
    ("wgebra/src/linalg/gemv.rs", "rust/wgebra/src/linalg/gemv.rs"),
    ("wgebra/src/linalg/op_assign.rs", "rust/wgebra/src/linalg/op_assign.rs"),
    ("wgebra/src/linalg/reduce.rs", "rust/wgebra/src/linalg/reduce.rs"),
    ("wgcore/src/shapes.rs", "rust/wgcore/src/shapes.rs"),
    ("wgcore/src/tensor.rs", "rust/wgcore/src/tensor.rs"),
    ("wgcore/src/kernel.rs", "rust/wgcore/src/kernel.rs"),
    ("wgcore/src/gpu.rs", "rust/wgcore/src/gpu.rs"),
    ("wgcore/src/timestamps.rs", "rust/wgcore/src/timestamps.rs"),
    ("wgcore/src/shader.rs", "rust/wgcore/src/shader.rs"),
]
ITEM = re.compile(r"\bpub(?:\([a-z]+\))?\s+(?:async\s+)?(?:unsafe\s+)?(fn|struct|enum|trait|const|type|mod)\b")
# methods of public traits carry no `pub`: files whose trait bodies are part of the surface
TRAIT_FILES = ("shader.rs", "kernel.rs")
TRAIT_FN = re.compile(r"(?<![\w])(?:async\s+)?fn\s+\w+|(?<![\w])const\s+[A-Z_]+\s*:")


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return "\n".join(re.sub(r"//.*$", "", l) for l in src.splitlines())


def take_signature(src: str, start: int) -> str:
    """Text from `start` to the first `{`, `;` (or `=` for consts / type aliases keep the right-hand side) outside brackets."""
    depth = 0
    k = start
    while k < len(src):
        c = src[k]
        if c in "([":
            depth += 1
        elif c in ")]":
            depth -= 1
        elif c == "<":
            depth += 1
        elif c == ">" and src[k - 1] != "-" and src[k - 1] != "=":
            depth -= 1
        elif c in "{;" and depth <= 0:
            break
        k += 1
    return src[start:k]


def normalise(sig: str) -> str:
    sig = re.sub(r"\s+", " ", sig).strip()
    if re.match(r"(pub\s+)?const\b", sig):
        sig = sig.split("=", 1)[0].strip()
    sig = re.sub(r"\s+where\b.*$", "", sig)
    sig = re.sub(r"\s*,\s*\)", ")", sig)                       # trailing commas of multi-line parameter lists
    sig = re.sub(r"\(\s+", "(", sig)
    sig = re.sub(r"\b(?:wgpu|crate::gpu|crate::shapes|crate::tensor|std::path|std::sync)::", "", sig)   # path spelling
    sig = re.sub(r"(?<=[(,] )mut (?=\w+:)|(?<=\()mut (?=\w+:)", "", sig)                               # `mut` on a parameter
    sig = re.sub(r"(?<=[(, ])_(?=[a-z]\w*:)", "", sig)                                                 # `_unused` parameter names
    return sig


def signatures(path: str):
    """(set of normalised signatures) of a Rust source file; test modules are skipped."""
    src = strip_comments(open(path).read())
    cut = src.find("#[cfg(test)]")
    if cut >= 0:
        src = src[:cut]
    out = set()
    for m in ITEM.finditer(src):
        out.add(normalise(take_signature(src, m.start())))
    if os.path.basename(path) in TRAIT_FILES:
        for tm in re.finditer(r"\bpub\s+trait\s+(\w+)[^{]*\{", src):
            depth, k = 1, tm.end()
            while k < len(src) and depth:
                depth += {"{": 1, "}": -1}.get(src[k], 0)
                k += 1
            body = src[tm.end():k - 1]
            # only the trait's own level: blank out nested blocks (default method bodies)
            flat, d = [], 0
            for c in body:
                if c == "{":
                    d += 1
                elif c == "}":
                    d -= 1
                    if d == 0:
                        flat.append(";")
                    continue
                if d == 0:
                    flat.append(c)
            flat = "".join(flat)
            for fm in TRAIT_FN.finditer(flat):
                out.add(f"trait {tm.group(1)} :: " + normalise(take_signature(flat, fm.start())))
    return out


def load_notes():
    notes = []
    p = os.path.join(ROOT, "rust", "SIGNATURES.notes")
    if os.path.exists(p):
        for l in open(p):
            l = l.rstrip("\n")
            if l.strip() and not l.startswith("#") and " :: " in l:
                k, why = l.split(" :: ", 1)
                notes.append((k.strip(), why.strip()))
    return notes


def main():
    notes = load_notes()
    unannotated = 0
    report = ["# tools/rust_signatures.py: public item signatures of the reference files (-) vs the shim files (+), per file pair.",
              "# A line is followed by `    = <reason>` when rust/SIGNATURES.notes explains it; a line without one is a defect.", ""]
    for ref_rel, shim_rel in PAIRS:
        ref_p, shim_p = os.path.join(REF, ref_rel), os.path.join(ROOT, shim_rel)
        if not os.path.exists(ref_p):
            report.append(f"## {ref_rel}: reference not available here (the check runs where /root/reference is mounted)")
            continue
        if not os.path.exists(shim_p):
            report.append(f"## {ref_rel} -> {shim_rel}: MISSING shim file")
            unannotated += 1
            continue
        a, b = signatures(ref_p), signatures(shim_p)
        only_ref, only_shim = sorted(a - b), sorted(b - a)
        report.append(f"## {ref_rel} -> {shim_rel}: {len(a & b)} identical, {len(only_ref)} only in the reference, {len(only_shim)} only in the shim")
        for tag, sigs in (("-", only_ref), ("+", only_shim)):
            for s in sigs:
                why = next((w for k, w in notes if k in s), None)
                report.append(f"{tag} {s}")
                if why:
                    report.append(f"    = {why}")
                else:
                    unannotated += 1
        report.append("")
    report.append(f"# unannotated differences: {unannotated}")
    text = "\n".join(report) + "\n"
    if "--check" in sys.argv:
        if unannotated:
            sys.stdout.write(text)
        return 1 if unannotated else 0
    open(os.path.join(ROOT, "rust", "SIGNATURES.diff"), "w").write(text)
    print(text)
    return 0


if __name__ == "__main__":
    sys.exit(main())
