"""ctypes harness for the reference's own package C layer (TEST INFRASTRUCTURE ONLY).

oracle/Makefile compiles /root/reference/src/{cctx,dctx,raw-file,raw-file-in,raw-file-out,dictionaries,zstd-info,utils}.c UNMODIFIED
against the miniature R API of oracle/rstub/ into two shared objects:
  oracle/_ref/librlayer_cpu.so   linked with the reference's libzstd (checks the stub and the replayed tests without a GPU)
  oracle/_ref/librlayer_gpu.so   every libzstd / ZDICT call bound to libzstdlite_gpu.so (include/zstdlite_gpu_map.h)
This module builds R objects, calls the .Call entry points registered in src/init.c:46-82 and converts results back, the way
R/*.R does."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(os.path.dirname(_HERE), "oracle", "_ref")
RAWSXP, STRSXP, VECSXP, INTSXP, LGLSXP, REALSXP, NILSXP, EXTPTRSXP = 24, 16, 19, 13, 10, 14, 0, 22


class RError(RuntimeError):
    """error() raised inside the C layer"""


class RLayer:
    def __init__(self, flavour):
        path = os.path.join(_REF, f"librlayer_{flavour}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (built by `make -C oracle rlayer` where /root/reference is present)")
        L = C.CDLL(path, mode=os.RTLD_LOCAL)
        vp = C.c_void_p
        for name, res, args in (("rstub_nil", vp, []), ("rstub_raw", vp, [vp, C.c_size_t]), ("rstub_str", vp, [C.c_char_p]), ("rstub_int", vp, [C.c_int]),
                                ("rstub_lgl", vp, [C.c_int]), ("rstub_real", vp, [C.c_double]), ("rstub_list", vp, [C.c_int]),
                                ("rstub_list_set", None, [vp, C.c_int, C.c_char_p, vp]), ("rstub_type", C.c_int, [vp]), ("rstub_len", C.c_size_t, [vp]),
                                ("rstub_data", vp, [vp]), ("rstub_elt", vp, [vp, C.c_int]), ("rstub_name", C.c_char_p, [vp, C.c_int]),
                                ("rstub_class", C.c_char_p, [vp]), ("rstub_last_error", C.c_char_p, []), ("rstub_last_warning", C.c_char_p, []),
                                ("rstub_warning_count", C.c_int, []), ("rstub_printed", C.c_char_p, []), ("rstub_reset_messages", None, []),
                                ("rstub_finalize", None, [vp]), ("rstub_call", vp, [vp, C.c_int, C.POINTER(vp)])):
            fn = getattr(L, name); fn.restype = res; fn.argtypes = args
        self.L = L
        self.nil = L.rstub_nil()

    # ---- Python -> R
    def to_r(self, v):
        L = self.L
        if v is None:
            return self.nil
        if isinstance(v, _Ptr):
            return v.p
        if isinstance(v, bool):
            return L.rstub_lgl(1 if v else 0)
        if isinstance(v, int):
            return L.rstub_int(v)
        if isinstance(v, float):
            return L.rstub_real(v)
        if isinstance(v, str):
            return L.rstub_str(v.encode())
        if isinstance(v, (bytes, bytearray, memoryview)):
            b = bytes(v)
            return L.rstub_raw(b, len(b))
        if isinstance(v, dict):
            l = L.rstub_list(len(v))
            for i, (k, x) in enumerate(v.items()):
                L.rstub_list_set(l, i, k.encode(), self.to_r(x))
            return l
        if isinstance(v, (list, tuple)):
            l = L.rstub_list(len(v))
            for i, x in enumerate(v):
                L.rstub_list_set(l, i, b"", self.to_r(x))
            return l
        raise TypeError(type(v))

    # ---- R -> Python
    def from_r(self, s):
        L = self.L
        t = L.rstub_type(s)
        n = L.rstub_len(s)
        if t == NILSXP:
            return None
        if t == RAWSXP:
            return C.string_at(L.rstub_data(s), n)
        if t == STRSXP:
            return C.string_at(L.rstub_data(s)).decode()
        if t in (INTSXP, LGLSXP):
            a = (C.c_int * n).from_address(L.rstub_data(s))
            vals = [bool(x) for x in a] if t == LGLSXP else list(a)
            return vals[0] if n == 1 else vals
        if t == REALSXP:
            a = (C.c_double * n).from_address(L.rstub_data(s))
            return a[0] if n == 1 else list(a)
        if t == VECSXP:
            return {L.rstub_name(s, i).decode() or str(i): self.from_r(L.rstub_elt(s, i)) for i in range(n)}
        if t == EXTPTRSXP:
            return _Ptr(s, L.rstub_class(s).decode())
        raise TypeError(f"SEXP type {t}")

    def call(self, name, *args):
        """.Call(name, ...): returns the converted result; raises RError(message) when the C code calls error()."""
        L = self.L
        fn = C.cast(getattr(L, name), C.c_void_p)
        arr = (C.c_void_p * max(1, len(args)))(*[self.to_r(a) for a in args])
        L.rstub_reset_messages()
        r = L.rstub_call(fn, len(args), arr)
        if not r:
            raise RError(L.rstub_last_error().decode())
        return self.from_r(r)

    @property
    def warnings(self):
        return self.L.rstub_warning_count(), self.L.rstub_last_warning().decode()

    def finalize(self, ptr):
        """run the external pointer's finalizer, as R's collector would (src/cctx.c:191-207, src/dctx.c:40-56)"""
        self.L.rstub_finalize(ptr.p)

    # ---- the R wrappers (R/cctx.R:38-48,58-66; R/raw.R; R/dictionaries.R; R/info.R)
    def zstd_cctx(self, level=3, num_threads=1, include_checksum=False, dict=None):
        return self.call("init_cctx_", {"level": level, "num_threads": num_threads, "include_checksum": include_checksum, "dict": dict})

    def zstd_dctx(self, validate_checksum=True, dict=None):
        return self.call("init_dctx_", {"validate_checksum": validate_checksum, "dict": dict})

    def zstd_compress(self, src, file=None, cctx=None, use_file_streaming=False, **opts):
        return self.call("zstd_compress_", src, file, cctx, opts, use_file_streaming)

    def zstd_decompress(self, src, type="raw", dctx=None, use_file_streaming=False, **opts):
        return self.call("zstd_decompress_", src, type, dctx, opts, use_file_streaming)

    def zstd_train_dict_compress(self, samples, size=100000, optim=False, optim_shrink_allow=0):
        return self.call("zstd_train_dictionary_", list(samples), size, optim, optim_shrink_allow)

    def zstd_dict_id(self, d):
        return self.call("zstd_dict_id_", d)

    def zstd_info(self, src):
        return self.call("zstd_info_", src)


class _Ptr:
    def __init__(self, p, klass):
        self.p, self.klass = p, klass
