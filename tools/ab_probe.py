"""A/B timing of builds of libwgebra_b200.so on the same box: bf16 GEMMs through the C ABI with raw ctypes (only entry points every
build exports).  One library per process (two builds in one process share their template statics through STB_GNU_UNIQUE symbols);
the run script alternates processes.  Usage: python tools/ab_probe.py lib.so [n ...]"""
import ctypes
import sys

vp, u32, u64, ci = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int


class VS(ctypes.Structure):
    _fields_ = [("size", u32 * 3), ("stride", u32), ("stride_mat", u32), ("offset", u32)]


class Lib:
    def __init__(self, path):
        self.L = L = ctypes.CDLL(path)
        self.path = path
        L.wgb_ctx_create.argtypes = [ci, ctypes.POINTER(vp)]
        L.wgb_buffer_create.argtypes = [vp, ctypes.c_size_t, u32, ctypes.POINTER(vp)]
        L.wgb_fill_uniform.argtypes = [vp, vp, ctypes.POINTER(VS), ci, u64, u32, u32]
        L.wgb_pass_begin.argtypes = [vp, ctypes.c_char_p, vp, vp, ctypes.POINTER(vp)]
        L.wgb_pass_end.argtypes = [vp]
        L.wgb_gemm_ex.argtypes = [vp, ci, vp, ctypes.POINTER(VS), vp, ctypes.POINTER(VS), vp, ctypes.POINTER(VS), ci, ci, ci]
        L.wgb_event_create.argtypes = [vp, ctypes.POINTER(vp)]
        L.wgb_event_record.argtypes = [vp, vp]
        L.wgb_event_elapsed_ms.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_float)]
        L.wgb_ctx_sync.argtypes = [vp]
        L.wgb_last_error_string.restype = ctypes.c_char_p
        self.ctx = vp()
        self.ck(L.wgb_ctx_create(0, ctypes.byref(self.ctx)))
        self.e0, self.e1 = vp(), vp()
        self.ck(L.wgb_event_create(self.ctx, ctypes.byref(self.e0)))
        self.ck(L.wgb_event_create(self.ctx, ctypes.byref(self.e1)))
        self.sets = {}

    def ck(self, st):
        if st != 0:
            raise RuntimeError(f"{self.path}: status {st}: {self.L.wgb_last_error_string().decode()}")

    def operands(self, n, nsets=4):
        if n in self.sets:
            return self.sets[n]
        L = self.L
        out = []
        p = vp()
        self.ck(L.wgb_pass_begin(self.ctx, b"init", None, None, ctypes.byref(p)))
        s = VS((n, n, 1), n, n * n, 0)
        for _ in range(nsets):
            bufs = []
            for k in range(3):
                b = vp()
                self.ck(L.wgb_buffer_create(self.ctx, n * n * 2, 0x8C, ctypes.byref(b)))
                if k < 2:
                    self.ck(L.wgb_fill_uniform(p, b, ctypes.byref(s), 1, 0x5EED0001 + k, 0, 0))
                bufs.append(b)
            out.append(bufs)
        self.ck(L.wgb_pass_end(p))
        self.ck(L.wgb_ctx_sync(self.ctx))
        self.sets[n] = out
        return out

    def time(self, n, steps):
        L = self.L
        sets = self.operands(n)
        s = VS((n, n, 1), n, n * n, 0)
        p = vp()
        self.ck(L.wgb_pass_begin(self.ctx, b"t", None, None, ctypes.byref(p)))
        for i in range(3):
            a, b, c = sets[i % len(sets)]
            self.ck(L.wgb_gemm_ex(p, 0, c, ctypes.byref(s), a, ctypes.byref(s), b, ctypes.byref(s), 1, 1, 0))
        self.ck(L.wgb_event_record(self.e0, p))
        for i in range(steps):
            a, b, c = sets[i % len(sets)]
            self.ck(L.wgb_gemm_ex(p, 0, c, ctypes.byref(s), a, ctypes.byref(s), b, ctypes.byref(s), 1, 1, 0))
        self.ck(L.wgb_event_record(self.e1, p))
        self.ck(L.wgb_pass_end(p))
        ms = ctypes.c_float()
        self.ck(L.wgb_event_elapsed_ms(self.e0, self.e1, ctypes.byref(ms)))
        return ms.value / steps


def main():
    import time
    lib = Lib(sys.argv[1])
    sizes = [int(x) for x in sys.argv[2:]] or [4096, 8192]
    for n in sizes:
        steps = 20 if n <= 4096 else 8
        for rep in range(4):
            time.sleep(0.5)           # let the part fall back to its idle power state: every sample is a cold burst
            t = lib.time(n, steps)
            print(f"AB n={n} rep={rep} {lib.path.split('/')[-1]}: {t * 1e3:.2f} us  {2.0 * n ** 3 / t / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
