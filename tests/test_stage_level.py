"""Stage-level differential tests against the reference: the match finder and the entropy stage are compared with
libzstd's SEPARATELY, through the reference's own experimental sequence API -- ZSTD_generateSequences (src/zstd/zstd.h:1584) and
ZSTD_compressSequences (src/zstd/zstd.h:1632) of oracle/_ref/libzstd_ref.so -- on one block at a time:

  * OUR sequences -> THEIR entropy stage: the frame decodes to the input, and our match finder + greedy walk leaves at most a few percent
    more to code than libzstd's fast / double-fast search at the same level (same entropy coder on both sides, so only the parse differs);
  * THEIR sequences -> OUR entropy stage: the frame decodes with libzstd, and is within 1.5 % of the frame libzstd's own entropy stage
    makes from the same sequences (same parse on both sides, so only the literal / sequence coding differs).

"Ours" is the product's ZL_HD source compiled for the CPU (tests/emul: the kernels' serial logic; -m gpu tests prove the CUDA kernels emit the
same bytes).  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from tests import emul_util


class _Seq(C.Structure):                                   # ZSTD_Sequence, src/zstd/zstd.h:1501-1530
    _fields_ = [("offset", C.c_uint), ("litLength", C.c_uint), ("matchLength", C.c_uint), ("rep", C.c_uint)]


@pytest.fixture(scope="module")
def L():
    from oracle import ref
    lib = ref.lib()
    lib.ZSTD_generateSequences.restype = C.c_size_t
    lib.ZSTD_generateSequences.argtypes = [C.c_void_p, C.POINTER(_Seq), C.c_size_t, C.c_void_p, C.c_size_t]
    lib.ZSTD_mergeBlockDelimiters.restype = C.c_size_t
    lib.ZSTD_mergeBlockDelimiters.argtypes = [C.POINTER(_Seq), C.c_size_t]
    lib.ZSTD_compressSequences.restype = C.c_size_t
    lib.ZSTD_compressSequences.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(_Seq), C.c_size_t, C.c_void_p, C.c_size_t]
    lib.ZSTD_sequenceBound.restype = C.c_size_t
    lib.ZSTD_sequenceBound.argtypes = [C.c_size_t]
    lib.ZSTD_createCCtx.restype = C.c_void_p
    lib.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.ZSTD_freeCCtx.argtypes = [C.c_void_p]
    return lib


def _their_sequences(L, data, level):
    """ZSTD_generateSequences at `level`, block delimiters merged away -> [n, 3] (litLength, matchLength, offset); last literals implicit"""
    c = L.ZSTD_createCCtx()
    L.ZSTD_CCtx_setParameter(c, 100, level)
    cap = L.ZSTD_sequenceBound(len(data))
    buf = (_Seq * cap)()
    n = L.ZSTD_generateSequences(c, buf, cap, data, len(data))
    assert not L.ZSTD_isError(n)
    n = L.ZSTD_mergeBlockDelimiters(buf, n)
    L.ZSTD_freeCCtx(c)
    rows = np.array([(buf[i].litLength, buf[i].matchLength, buf[i].offset) for i in range(n)], dtype=np.uint32).reshape(-1, 3)
    return rows[rows[:, 1] != 0]                            # (a trailing literals-only entry is implicit for both consumers)


def _their_entropy(L, data, seqs, level):
    """ZSTD_compressSequences over explicit sequences (no block delimiters: the leftover bytes are the last literals)"""
    c = L.ZSTD_createCCtx()
    L.ZSTD_CCtx_setParameter(c, 100, level)
    L.ZSTD_CCtx_setParameter(c, 1009, 1)                    # ZSTD_c_validateSequences (experimentalParam12)
    buf = (_Seq * max(1, len(seqs)))()
    for i, (ll, ml, off) in enumerate(seqs):
        buf[i].litLength, buf[i].matchLength, buf[i].offset, buf[i].rep = int(ll), int(ml), int(off), 0
    cap = len(data) + (len(data) >> 7) + 1024
    dst = C.create_string_buffer(cap)
    r = L.ZSTD_compressSequences(c, dst, cap, buf, len(seqs), data, len(data))
    L.ZSTD_freeCCtx(c)
    assert not L.ZSTD_isError(r), L.ZSTD_getErrorName(r)
    return dst.raw[:r]


CASES = [("text", 131072), ("text", 20000), ("rdf", 131072), ("rdf", 65536), ("lowent", 65536)]


@pytest.mark.parametrize("family,size", CASES)
@pytest.mark.parametrize("level", [1, 3])
def test_our_sequences_through_their_entropy_stage(L, family, size, level):
    from oracle import ref
    from zstdlite_b200 import corpus
    data = corpus.make(family, size, 2).tobytes()
    ours = emul_util.sequences(data, level)
    assert int(ours[:, 0].sum() + ours[:, 1].sum()) <= len(data) and (ours[:, 2] >= 1).all()
    frame = _their_entropy(L, data, ours, level)
    assert ref.decompress(frame) == data                    # the sequences describe the input exactly (validated by libzstd, too)
    theirs = _their_sequences(L, data, level)
    their_frame = _their_entropy(L, data, theirs, level)
    # same entropy coder on both sides: the parse alone decides the size.  Every position is a candidate in our finder, so the
    # greedy walk is at most 3 % behind the reference's search at the same level (measured -2 % .. +2 %)
    assert len(frame) <= 1.03 * len(their_frame) + 16, (len(frame), len(their_frame))
    if family != "lowent":
        assert int(ours[:, 1].sum()) >= 0.95 * int(theirs[:, 1].sum())          # bytes covered by matches


@pytest.mark.parametrize("family,size", CASES)
@pytest.mark.parametrize("level", [1, 3])
def test_their_sequences_through_our_entropy_stage(L, family, size, level):
    from oracle import ref
    from zstdlite_b200 import corpus
    data = corpus.make(family, size, 4).tobytes()
    theirs = _their_sequences(L, data, level)
    frame = emul_util.encode_sequences(data, theirs, level)
    assert not isinstance(frame, tuple), frame
    assert ref.decompress(frame) == data
    their_frame = _their_entropy(L, data, theirs, level)
    # same parse on both sides: literal (Huffman) and sequence (FSE) coding alone decide the size
    assert len(frame) <= 1.015 * len(their_frame) + 16, (len(frame), len(their_frame))
