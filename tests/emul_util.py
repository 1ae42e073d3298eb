"""ctypes access to the CPU emulation of the CUDA entropy kernels (tests/emul/, test infra only)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "emul", "libzl_emul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "emul", "emul_decode.cpp"), os.path.join(_HERE, "emul", "emul_encode.cpp")]
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unused-function", "-o", _SO] + srcs)
        L = C.CDLL(_SO)
        L.zl_emul_decompress_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint)]
        L.zl_emul_decompress_frame.restype = C.c_size_t
        L.zl_emul_compress_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint]
        L.zl_emul_compress_frame.restype = C.c_size_t
        L.zl_emul_compress_frame_dict.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_size_t]
        L.zl_emul_compress_frame_dict.restype = C.c_size_t
        L.zl_emul_sequences.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
        L.zl_emul_sequences.restype = C.c_size_t
        L.zl_emul_encode_sequences.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
        L.zl_emul_encode_sequences.restype = C.c_size_t
        _lib = L
    return _lib


def decompress_frame(c, cap):
    """returns bytes, or ('ERR', code)"""
    dst = C.create_string_buffer(cap + 8)
    n = C.c_uint(0)
    r = lib().zl_emul_decompress_frame(dst, cap, bytes(c), len(c), C.byref(n))
    if r > 2**63:
        return ("ERR", 2**64 - r)
    return dst.raw[:r]


def compress_frame(data, level=3, checksum=False, xxh32=0, dict=None):
    """CPU emulation of the CUDA compressor -> one frame (bytes), or ('ERR', code)."""
    data = bytes(data)
    cap = len(data) + (len(data) >> 7) + 1024
    dst = C.create_string_buffer(cap)
    if dict is not None:
        r = lib().zl_emul_compress_frame_dict(dst, cap, data, len(data), level, 1 if checksum else 0, xxh32, bytes(dict), len(dict))
    else:
        r = lib().zl_emul_compress_frame(dst, cap, data, len(data), level, 1 if checksum else 0, xxh32)
    if r > 2**63:
        return ("ERR", 2**64 - r)
    return dst.raw[:r]


def sequences(data, level=3):
    """the emulated match finder + greedy walk on ONE block -> numpy array [nseq, 3] of (litLength, matchLength, offset)"""
    import numpy as np
    data = bytes(data)
    out = np.zeros((len(data) // 3 + 16, 3), dtype=np.uint32)
    n = lib().zl_emul_sequences(data, len(data), level, out.ctypes.data, out.shape[0])
    return out[:n].copy()


def encode_sequences(data, seqs, level=3):
    """ONE block's frame from the given (litLength, matchLength, offset) rows through the product's entropy stage only"""
    import numpy as np
    data = bytes(data)
    seqs = np.ascontiguousarray(seqs, dtype=np.uint32)
    cap = len(data) + (len(data) >> 7) + 1024
    dst = C.create_string_buffer(cap)
    r = lib().zl_emul_encode_sequences(dst, cap, data, len(data), level, seqs.ctypes.data, seqs.shape[0])
    if r > 2**63:
        return ("ERR", 2**64 - r)
    return dst.raw[:r]
