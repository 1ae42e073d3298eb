"""ctypes binding of oracle/libzl_oracle.so (oracle/zl_oracle.c).  TEST INFRA ONLY."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libzl_oracle.so")
_lib = None

ERR_NAMES = {1: "GENERIC", 10: "prefix_unknown", 14: "frameParameter_unsupported", 16: "windowTooLarge",
             20: "corruption_detected", 22: "checksum_wrong", 24: "literals_headerWrong",
             30: "dictionary_corrupted", 32: "dictionary_wrong", 44: "tableLog_tooLarge",
             48: "maxSymbolValue_tooSmall", 70: "dstSize_tooSmall", 72: "srcSize_wrong"}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH) or os.path.getmtime(_PATH) < os.path.getmtime(os.path.join(_HERE, "zl_oracle.c")):
            subprocess.check_call(["make", "-s", "-C", _HERE, "libzl_oracle.so"])
        L = C.CDLL(_PATH, mode=os.RTLD_LOCAL)
        vp, sz = C.c_void_p, C.c_size_t
        L.zlo_decompress.argtypes = [vp, sz, vp, sz, vp, sz, C.c_int]
        L.zlo_decompress.restype = sz
        L.zlo_find_frame_compressed_size.argtypes = [vp, sz]
        L.zlo_find_frame_compressed_size.restype = sz
        L.zlo_get_frame_content_size.argtypes = [vp, sz]
        L.zlo_get_frame_content_size.restype = C.c_uint64
        L.zlo_xxh64.argtypes = [vp, sz, C.c_uint64]
        L.zlo_xxh64.restype = C.c_uint64
        L.zlo_compress_bound.argtypes = [sz]
        L.zlo_compress_bound.restype = sz
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code):
        self.code = code
        super().__init__(ERR_NAMES.get(code, str(code)))


def is_error(r):
    return r > 2**64 - 121


def decompress(data, cap, dict=None, ignore_checksum=False):
    L = lib()
    data = bytes(data)
    dst = C.create_string_buffer(max(1, cap))
    d = bytes(dict) if dict else None
    r = L.zlo_decompress(dst, cap, data, len(data), d, len(d) if d else 0, 1 if ignore_checksum else 0)
    if is_error(r):
        raise OracleError(2**64 - r)
    return dst.raw[:r]


def xxh64(data, seed=0):
    data = bytes(data)
    return lib().zlo_xxh64(data, len(data), seed)


def frame_compressed_size(data):
    data = bytes(data)
    return lib().zlo_find_frame_compressed_size(data, len(data))


def frame_content_size(data):
    data = bytes(data)
    return lib().zlo_get_frame_content_size(data, len(data))


def compress_bound(n):
    return lib().zlo_compress_bound(n)
