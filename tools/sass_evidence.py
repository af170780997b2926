"""SASS evidence table of the tcgen05 GEMM kernels in the shipped library: mnemonic counts per instantiation.
    python tools/sass_evidence.py > profiles/r2_gemm_tc_sass_evidence.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "wgmath_b200", "libwgebra_b200.so")
COLS = [("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
        ("UTCBAR", r"\bUTCBAR"), ("STG", r"\bSTG\b"), ("mm.st", r"\bSTG\.E\.128\.STRONG\.SYS"), ("st.sys", r"\bST[G]?\.E[.\w]*\.STRONG\.SYS"),
        ("ld.sys", r"\bLD[G]?\.E[.\w]*\.STRONG\.SYS"), ("STS", r"\bSTS\b"), ("LDS", r"\bLDS\b"), ("SHFL", r"\bSHFL"), ("ACQBULK", r"\bACQBULK"),
        ("PREEXIT", r"\bPREEXIT")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return out


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    rows = []
    for f in funcs:
        name, body = f.split("\n", 1)
        if "gemm_tc_kernel" not in name:
            continue
        rows.append((name.strip(), [len(re.findall(rx, body)) for _, rx in COLS]))
    names = demangle([r[0] for r in rows])
    print("# cuobjdump -sass wgmath_b200/libwgebra_b200.so : mnemonic counts per gemm_tc_kernel instantiation (sm_100a)")
    print("# template args: <KIND (0 bf16, 1 tf32), A MN-major, B MN-major, BLOCK_N, PASSES, TOut, CTA group, FS (3xTF32 operand split in the kernel)>")
    print("# UTCHMMA = tcgen05.mma; UTMALDG = TMA load; UTMASTG = TMA bulk tensor STORE (the smem-staged epilogue, one per destination);")
    print("# LDTM / STTM = tcgen05.ld / st; UTCBAR = tcgen05.commit; ACQBULK / PREEXIT = griddepcontrol (PDL).")
    print("# mm.st = STG.E.128.STRONG.SYS: how `multimem.st.relaxed.sys.global.v4.f32` is encoded (a 128-bit system-scope store; that it is a")
    print("# multicast store is a property of the address it is given, the NVSwitch multicast mapping) — the multicast epilogue of the fused all-gather.")
    print("# st.sys / ld.sys = every system-scope store / load (the multicast stores plus the ready / done flags of the handshake); STS / LDS include")
    print("# the epilogue staging block and, in the FS kernels, the operand split done by the converter warps.")
    print(f"{'instantiation':46s}" + "".join(f"{c:>8s}" for c, _ in COLS))
    for (_, counts), dn in sorted(zip(rows, names), key=lambda x: x[1]):
        m = re.search(r"gemm_tc_kernel<(.*?)>\(", dn)
        label = "<" + (m.group(1) if m else dn[:40]).replace("(bool)", "").replace("__nv_bfloat16", "bf16") + ">"
        print(f"{label:46s}" + "".join(f"{c:8d}" for c in counts))
    print(f"# {len(rows)} instantiations")


if __name__ == "__main__":
    sys.exit(main())
