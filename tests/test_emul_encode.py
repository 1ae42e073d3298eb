"""CPU tests of the compressor's entropy/format logic: the product's ZL_HD source (zl_enc_entropy.cuh, zl_enc_match.cuh)
compiled with g++ and driven serially (tests/emul/emul_encode.cpp).  Every frame must decode with the reference's
libzstd and the C restatement to the original bytes; ratio within 3% of libzstd at the same level (the north-star bar).
The -m gpu tests then require the CUDA kernels to reproduce these frames byte for byte."""
import numpy as np
import pytest


@pytest.mark.parametrize("family", ["text", "rdf", "lowent", "rand", "rle"])
def test_emulated_frames_decode_with_libzstd(ref, restate, family):
    from zstdlite_b200 import corpus
    from tests import emul_util
    for size in (0, 1, 6, 7, 8, 50, 255, 256, 1023, 1024, 5000, 16384, 65536, 131072, 131073, 300000):
        d = corpus.make(family, size, 3).tobytes()
        for lvl in (1, 2, 3):
            for ck in (False, True):
                c = emul_util.compress_frame(d, lvl, ck, restate.xxh64(d) & 0xFFFFFFFF)
                assert not isinstance(c, tuple), (family, size, lvl, c)
                assert len(c) <= ref.lib().ZSTD_compressBound(size)
                assert ref.decompress(c) == d, (family, size, lvl, ck)
                assert restate.decompress(c, size) == d
                assert emul_util.decompress_frame(c, size) == d           # and the emulated CUDA decoder


def test_emulated_ratio_within_3_percent(ref):
    from zstdlite_b200 import corpus
    from tests import emul_util
    for fam in ("text", "rdf", "lowent"):
        for fb in (65536, 131072):
            bufs = [corpus.make(fam, fb, 40 + i).tobytes() for i in range(6)]
            for lvl in (1, 3):
                ours = sum(len(emul_util.compress_frame(b, lvl)) for b in bufs)
                theirs = sum(len(ref.compress(b, lvl)) for b in bufs)
                assert ours <= theirs * 1.03, (fam, fb, lvl, ours, theirs)


def test_levels_4_and_5_within_3_percent_of_libzstd_at_those_levels(ref):
    """levels 4 and 5 are served by the level-3 engine (ZSTD_c_compressionLevel in zl_api_compress.cu): the claim behind that is this bar --
    its output within 3 % of libzstd's at the SAME level, per family, from 4 KB objects to a 2 MiB single frame, and with a dictionary.
    Level 6 is outside it (text of 2 MiB: 1.035x) and stays refused."""
    from zstdlite_b200 import corpus
    from tests import emul_util
    for lvl in (4, 5):
        for fam in ("text", "rdf", "lowent"):
            for fb, cnt in ((4000, 8), (131072, 4)):
                bufs = [corpus.make(fam, fb, 300 + i).tobytes() for i in range(cnt)]
                ours = sum(len(emul_util.compress_frame(b, lvl)) for b in bufs)
                theirs = sum(len(ref.compress(b, lvl)) for b in bufs)
                assert ours <= theirs * 1.03, (fam, fb, lvl, ours, theirs)
        big = corpus.make("text", 2 << 20, 300).tobytes()
        c = emul_util.compress_frame(big, lvl)
        assert ref.decompress(c) == big and len(c) <= len(ref.compress(big, lvl)) * 1.03, lvl
        objs = corpus.small_objects(4500)
        d = ref.train_dict(objs[:3000], 5000)
        test = objs[3000:4500:3]
        ours = sum(len(emul_util.compress_frame(o, lvl, dict=d)) for o in test)
        assert ours <= sum(len(ref.compress(o, lvl, dict=d)) for o in test) * 1.03, lvl
    big6 = len(emul_util.compress_frame(big, 6)) / len(ref.compress(big, 6))
    assert big6 > 1.03, big6            # (if this ever fails, level 6 can be served as well)


def test_odd_inputs(ref):
    """few symbols, skewed histograms (depth-limited Huffman), long runs, tiny alphabets, binary ramps"""
    from tests import emul_util
    rng = np.random.default_rng(77)
    cases = [bytes([7]) * 100000, bytes(range(256)) * 300, b"ab" * 40000, b"\x00" * 70000 + b"\x01" * 70000,
             rng.choice(np.arange(256, dtype=np.uint8), 90000, p=np.array([0.5 ** min(i + 1, 40) for i in range(256)]) / sum(0.5 ** min(i + 1, 40) for i in range(256))).tobytes(),
             rng.integers(0, 2, 50000, dtype=np.uint8).tobytes(), rng.integers(0, 256, 777, dtype=np.uint8).tobytes() * 90,
             np.repeat(rng.integers(0, 256, 3000, dtype=np.uint8), rng.integers(1, 90, 3000)).tobytes()]
    for d in cases:
        for lvl in (1, 3):
            c = emul_util.compress_frame(d, lvl, True, 0)
            assert ref.DCtx(validate_checksum=False).decompress(c) == d


def test_frame_header_layout_matches_reference(ref):
    """single-segment header bytes equal libzstd's for <= 128 KiB inputs (zstd.c:27089-27135)"""
    from tests import emul_util
    for size in (0, 1, 255, 256, 65535, 65536, 65791, 65792, 131072):
        d = bytes(size)
        ours, theirs = emul_util.compress_frame(d, 3), ref.compress(d, 3)
        hs = 5 + (1 if size < 256 else 2 if size < 65792 else 4)
        assert ours[:hs] == theirs[:hs], size


def test_emulated_dictionary_mode(ref):
    """configs[3] on the CPU emulation: frames made with a dictionary decode with libzstd + that dictionary, carry its
    ID, beat the no-dictionary size, and stay within 3% of libzstd's size with the same dictionary and level."""
    from zstdlite_b200 import corpus
    from tests import emul_util
    samples = corpus.small_objects(1500)
    trained = ref.train_dict(samples[:1000], 5000)
    raw = b"".join(samples[:20])
    work = samples[1000:1300]
    for dd in (trained, raw):
        did = ref.lib().ZDICT_getDictID(dd, len(dd))
        for lvl in (1, 3):
            rc, rd = ref.CCtx(level=lvl, dict=dd), ref.DCtx(dict=dd)
            ours = theirs = plain = 0
            for s in work:
                c = emul_util.compress_frame(s, lvl, dict=dd)
                assert rd.decompress(c, cap=len(s)) == s
                assert ref.lib().ZSTD_getDictID_fromFrame(c, len(c)) == did
                ours += len(c); theirs += len(rc.compress(s)); plain += len(emul_util.compress_frame(s, lvl))
            assert ours <= theirs * 1.03, (lvl, ours, theirs)
            assert ours < plain
    # multi-block input: only the first block may reference the dictionary; frames still decode
    d = corpus.make("text", 300000, 3).tobytes()
    tdict = ref.train_dict([corpus.make("text", 4000, 50 + i).tobytes() for i in range(200)], 20000)
    c = emul_util.compress_frame(d, 3, dict=tdict)
    assert ref.decompress(c, dict=tdict) == d
