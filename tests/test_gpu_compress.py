"""-m gpu parity tests for the CUDA compressor, through the C ABI (ZSTD_compress2 / zl_compress_batch / zl_compress_split).

Bar (BASELINE.json north_star): every GPU-compressed stream decodes with the reference's libzstd to byte-identical
input; compression ratio within 3% of the reference at the same level on the same framing.  In addition the CUDA
path must reproduce the CPU emulation of its own algorithm (tests/emul, same ZL_HD source) byte for byte."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def z():
    import torch
    assert torch.cuda.is_available()
    import zstdlite_b200 as zz
    return zz


def _gpu_compress_batch(z, bufs, level, checksum=False, device=True, caps=None, dict=None):
    import torch
    from tests.gpu_util import to_dev
    L = z._lib.lib()
    cctx = z.zstd_cctx(level=level, include_checksum=checksum, dict=dict)
    caps = caps or [L.ZSTD_compressBound(len(b)) for b in bufs]
    if device:
        offs = np.concatenate([[0], np.cumsum([len(b) for b in bufs])]).astype(np.int64)
        src = to_dev(np.frombuffer(b"".join(bufs), dtype=np.uint8))
        doffs = np.concatenate([[0], np.cumsum([(c + 15) // 16 * 16 for c in caps])]).astype(np.int64)
        dst = torch.zeros(int(doffs[-1]) + 64, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        res = z.compress_batch(cctx, [src.data_ptr() + int(o) for o in offs[:-1]], [len(b) for b in bufs],
                               [dst.data_ptr() + int(o) for o in doffs[:-1]], list(caps), device=True)
        host = dst.cpu().numpy()
        return res, [host[int(doffs[i]):int(doffs[i]) + res[i]].tobytes() if not z.is_error(res[i]) else None for i in range(len(bufs))]
    sb = [C.create_string_buffer(bytes(b), max(1, len(b))) for b in bufs]
    db = [C.create_string_buffer(max(1, c)) for c in caps]
    res = z.compress_batch(cctx, [C.addressof(b) for b in sb], [len(b) for b in bufs], [C.addressof(b) for b in db], list(caps), device=False)
    return res, [db[i].raw[:res[i]] if not z.is_error(res[i]) else None for i in range(len(bufs))]


@pytest.mark.parametrize("family", ["text", "rdf", "lowent", "rand", "rle"])
def test_round_trip_and_emulation_parity(z, ref, restate, family):
    from zstdlite_b200 import corpus
    from tests import emul_util
    sizes = (0, 1, 6, 7, 8, 9, 63, 64, 300, 4095, 4096, 65536, 131071, 131072, 131073, 400000)
    bufs = [corpus.make(family, s, 13).tobytes() for s in sizes]
    for lvl in (1, 2, 3):
        for ck in (False, True):
            res, outs = _gpu_compress_batch(z, bufs, lvl, ck)
            for d, r, c in zip(bufs, res, outs):
                assert not z.is_error(r), (len(d), lvl, z.error_name(r))
                assert len(c) <= ref.lib().ZSTD_compressBound(len(d))
                assert ref.decompress(c) == d, (family, len(d), lvl, ck)                  # libzstd accepts it
                assert restate.decompress(c, len(d)) == d
                want = emul_util.compress_frame(d, lvl, ck, restate.xxh64(d) & 0xFFFFFFFF)
                assert c == want, (family, len(d), lvl, ck, "CUDA output differs from the CPU emulation of the same algorithm")
    # our own decoder reads them back too
    from tests.gpu_util import gpu_decompress_batch
    res2, back = gpu_decompress_batch(outs, [len(b) for b in bufs])
    assert back == bufs


def test_one_shot_api_like_the_reference_tests(z, ref):
    """tests/testthat/test-compress-raw.R, test-cctx.R (determinism), test-checksums.R (+4 bytes)"""
    from zstdlite_b200 import corpus
    d = corpus.make("text", 300000, 2).tobytes() + corpus.make("rdf", 300000, 2).tobytes()
    a = z.zstd_compress(d)
    assert ref.decompress(a) == d and z.zstd_decompress(a) == d
    assert z.zstd_info(a)["uncompressed_size"] == len(d) and z.zstd_info(a)["compressed_size"] == len(a)
    cctx = z.zstd_cctx(level=3)
    assert z.zstd_compress(d, cctx=cctx) == a and z.zstd_compress(d, cctx=cctx) == a              # reused context, twice
    assert z.zstd_compress(d, num_threads=2) == a                                               # nbWorkers is a hint
    ck = z.zstd_compress(d, include_checksum=True)
    assert len(ck) == len(a) + 4 and ref.decompress(ck) == d
    bad = bytearray(ck); bad[-1] ^= 0xFF
    with pytest.raises(z.ZstdError, match="doesn't match checksum"):
        z.zstd_decompress(bytes(bad))
    assert z.zstd_compress("héllo wörld" * 50) and z.zstd_decompress(z.zstd_compress("héllo wörld" * 50), type="string") == "héllo wörld" * 50
    assert z.zstd_cctx(level=1, num_threads=3, include_checksum=True).settings() == {"level": 1, "num_threads": 3, "include_checksum": 1}
    for lvl in (-5, 0):                               # fast levels run the level-1 engine, 0 is the default: valid frames
        assert ref.decompress(z.zstd_compress(d[:50000], level=lvl)) == d[:50000]
    for lvl in (6, 9, 22):                            # levels >= 6: refused, or -- opted in -- the level-3 engine's bytes under the caller's label
        with pytest.raises(z.ZstdError, match="Bad compression level"):
            z.zstd_compress(d[:50000], level=lvl)
        assert z.zstd_compress(d[:50000], cctx=z.zstd_cctx(level=lvl, level_fallback=True)) == z.zstd_compress(d[:50000], level=3)


def test_destination_too_small(z):
    from zstdlite_b200 import corpus
    d = corpus.make("rand", 50000, 1).tobytes()
    res, outs = _gpu_compress_batch(z, [d, d], 3, caps=[1000, 60000])
    assert z.is_error(res[0]) and z.error_name(res[0]) == "Destination buffer is too small"
    assert not z.is_error(res[1])


def test_batch_ratio_vs_reference(z, ref):
    """configs[2] shape at reduced count: 128 KiB slabs, levels 1 and 3, ratio vs libzstd on the same slabs"""
    from zstdlite_b200 import corpus
    for fam in ("text", "rdf", "lowent"):
        bufs = [corpus.make(fam, 131072, 100 + i).tobytes() for i in range(24)]
        for lvl in (1, 3):
            res, outs = _gpu_compress_batch(z, bufs, lvl)
            ours = sum(len(c) for c in outs)
            theirs = sum(len(ref.compress(b, lvl)) for b in bufs)
            for b, c in zip(bufs, outs):
                assert ref.decompress(c) == b
            assert ours <= theirs * 1.03, (fam, lvl, ours, theirs)
    # host-pointer path gives the same bytes
    res_h, outs_h = _gpu_compress_batch(z, bufs[:6], 3, device=False)
    assert outs_h == outs[:6]


def test_levels_4_and_5_are_within_3_percent_of_libzstd_at_those_levels(z, ref):
    """levels 4 and 5 are served by the level-3 engine (every position inserted and verified, two tables): its output must stay within 3 %
    of what libzstd produces AT THE SAME LEVEL -- 128 KiB slabs per family, a 2 MiB single frame, small objects with a dictionary"""
    from zstdlite_b200 import corpus
    for lvl in (4, 5):
        for fam in ("text", "rdf", "lowent"):
            bufs = [corpus.make(fam, 131072, 300 + i).tobytes() for i in range(8)]
            res, outs = _gpu_compress_batch(z, bufs, lvl)
            for b, c in zip(bufs, outs):
                assert ref.decompress(c) == b
            ours, theirs = sum(len(c) for c in outs), sum(len(ref.compress(b, lvl)) for b in bufs)
            assert ours <= theirs * 1.03, (fam, lvl, ours, theirs)
        for fam in ("text", "rdf"):
            big = corpus.make(fam, 2 << 20, 300).tobytes()
            c = z.zstd_compress(big, level=lvl)
            assert ref.decompress(c) == big and len(c) <= len(ref.compress(big, lvl)) * 1.03, (fam, lvl)
        objs = corpus.small_objects(6000)
        d = ref.train_dict(objs[:3000], 5000)
        test = objs[3000:6000:3]
        res, outs = _gpu_compress_batch(z, test, lvl, dict=d)
        rd = ref.DCtx(dict=d)
        assert all(rd.decompress(c, cap=len(o)) == o for o, c in list(zip(test, outs))[::50])
        assert sum(len(c) for c in outs) <= sum(len(ref.compress(o, lvl, dict=d)) for o in test) * 1.03, lvl


def test_compress_split_is_a_standard_multi_frame_stream(z, ref):
    from zstdlite_b200 import corpus
    L = z._lib.lib()
    d = corpus.make("text", 1_000_000, 4).tobytes() + corpus.make("rand", 100_000, 4).tobytes()
    cctx = z.zstd_cctx(level=3, include_checksum=True)
    cap = len(d) + len(d) // 64 + 4096
    dst = C.create_string_buffer(cap)
    nf = (len(d) + 131071) // 131072
    fsz = (C.c_size_t * nf)()
    r = L.zl_compress_split(cctx._p, dst, cap, d, len(d), 131072, fsz, 0)
    assert not z.is_error(r), z.error_name(r)
    assert sum(fsz) == r
    blob = dst.raw[:r]
    assert ref.DCtx().decompress(blob, cap=len(d), all_frames=True) == d
    out = C.create_string_buffer(len(d))
    dctx = z.zstd_dctx()                                       # (kept alive: a temporary would be finalised before the call runs)
    rr = L.ZSTD_decompressDCtx(dctx._p, out, len(d), blob, len(blob))
    assert rr == len(d) and out.raw == d


def test_scattered_host_buffers_are_staged(z, ref):
    """thousands of separately allocated host objects (what a list of R raw vectors is): inputs are packed into pinned staging by
    the host, outputs gathered on the device and unpacked -- same frames as the device-pointer path, both directions"""
    import time
    from zstdlite_b200 import corpus
    from tests.gpu_util import gpu_decompress_batch
    objs = corpus.small_objects(3000) + [corpus.make("text", 40000, 5).tobytes(), b"", b"x" * 9]
    d = ref.train_dict(objs[:2000], 4096)
    res_h, outs_h = _gpu_compress_batch(z, objs, 3, checksum=True, device=False, dict=d)
    res_d, outs_d = _gpu_compress_batch(z, objs, 3, checksum=True, device=True, dict=d)
    assert outs_h == outs_d and not any(z.is_error(r) for r in res_h)
    rd = ref.DCtx(dict=d)
    for o, c in list(zip(objs, outs_h))[::37] + [(objs[-3], outs_h[-3]), (objs[-2], outs_h[-2]), (objs[-1], outs_h[-1])]:
        assert rd.decompress(c, cap=len(o)) == o
    t0 = time.time()
    res, outs = gpu_decompress_batch(outs_h, [len(o) for o in objs], dctx=z.zstd_dctx(dict=d), device=False)
    assert outs == objs and time.time() - t0 < 5.0
    # a destination that is too small for one object only fails that object
    caps = [len(o) for o in objs]; caps[5] = 3
    res2, outs2 = gpu_decompress_batch(outs_h, caps, dctx=z.zstd_dctx(dict=d), device=False)
    assert z.is_error(res2[5]) and outs2[4] == objs[4] and outs2[6] == objs[6]


def test_large_buffers_with_checksum(z, ref):
    """the trailer hash of a large buffer takes the warp-streamed XXH64 (16-byte aligned input) or the quad form (misaligned):
    libzstd verifies the checksum when it decodes"""
    import torch
    from zstdlite_b200 import corpus
    from tests.gpu_util import to_dev
    d = corpus.make("text", 2_500_001, 12).tobytes()
    c = z.zstd_compress(d, level=3, include_checksum=True)                       # host path: staged, aligned
    assert z.zstd_info(c)["has_checksum"] and ref.decompress(c) == d and z.zstd_decompress(c) == d
    L = z._lib.lib()
    cctx = z.zstd_cctx(level=1, include_checksum=True)
    cap = int(L.ZSTD_compressBound(len(d)))
    for shift in (0, 3):                                                        # device pointers: aligned and misaligned input
        src = to_dev(np.frombuffer(b"\x00" * shift + d, dtype=np.uint8))
        dst = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
        res = z.compress_batch(cctx, [src.data_ptr() + shift], [len(d)], [dst.data_ptr()], [cap], device=True)
        assert not z.is_error(res[0])
        out = dst[:res[0]].cpu().numpy().tobytes()
        assert ref.decompress(out) == d
        bad = bytearray(out); bad[-1] ^= 0x40
        with pytest.raises(z.ZstdError, match="checksum"):
            z.zstd_decompress(bytes(bad))


def test_compress_split_pipelined_host_path(z, ref):
    """host buffers of >= 2 chunks take the staged pipeline (copies of neighbouring chunks overlap the kernels): same bytes as the
    device-pointer path, a standard multi-frame stream for libzstd"""
    import torch
    from zstdlite_b200 import corpus
    L = z._lib.lib()
    fs = 16384
    pool, _ = corpus.mixed_frames(2048, fs, pool=32, rotate=True)
    data = np.tile(pool.reshape(-1), 9)[: 17000 * fs + 777]                    # 17,001 frames: two full chunks of 8,192 + a short one
    n = data.size
    nf = (n + fs - 1) // fs
    cap = nf * (int(L.ZSTD_compressBound(fs)) + 8)
    cctx = z.zstd_cctx(level=3, include_checksum=True)
    hsrc = torch.from_numpy(data.copy()).pin_memory()
    hdst = torch.zeros(cap, dtype=torch.uint8).pin_memory()
    fsz = (C.c_size_t * nf)()
    r = L.zl_compress_split(cctx._p, C.c_void_p(hdst.data_ptr()), cap, C.c_void_p(hsrc.data_ptr()), n, fs, fsz, 0)
    assert not z.is_error(r), z.error_name(r)
    assert sum(fsz) == r
    blob = hdst[:r].numpy().tobytes()
    dsrc = hsrc.cuda(); ddst = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    r2 = L.zl_compress_split(cctx._p, C.c_void_p(ddst.data_ptr()), cap, C.c_void_p(dsrc.data_ptr()), n, fs, None, 1)
    assert r2 == r and ddst[:r].cpu().numpy().tobytes() == blob
    assert ref.DCtx().decompress(blob, cap=n, all_frames=True) == data.tobytes()
    assert z.zstd_decompress(blob, all_frames=True) == data.tobytes()


def test_multi_frame_round_trip_through_the_package_api(z, ref):
    """split output and concatenated streams round-trip through zstd_compress / zstd_decompress /
    zstd_serialize / zstd_unserialize themselves (content sizes summed over all frames, one GPU batch)."""
    from zstdlite_b200 import corpus
    d = corpus.make("rdf", 700_000, 9).tobytes() + corpus.make("lowent", 50_001, 9).tobytes()
    blob = z.zstd_compress(d, level=3, frame_size=65536, include_checksum=True)
    nf = (len(d) + 65535) // 65536
    pos = cnt = 0
    while pos < len(blob):
        pos += ref.lib().ZSTD_findFrameCompressedSize(blob[pos:], len(blob) - pos); cnt += 1
    assert cnt == nf and pos == len(blob)
    assert ref.lib().ZSTD_findDecompressedSize(blob, len(blob)) == len(d) == z._lib.lib().ZSTD_findDecompressedSize(blob, len(blob))
    assert ref.DCtx().decompress(blob, cap=len(d), all_frames=True) == d
    assert z.zstd_decompress(blob, all_frames=True) == d
    assert z.zstd_decompress(blob) == d[:65536]                          # the reference's behaviour: first frame only
    # frames written by libzstd, a skippable frame in between
    skip = (0x184D2A50).to_bytes(4, "little") + (5).to_bytes(4, "little") + b"hello"
    cat = ref.compress(d[:100_000], 3) + skip + ref.compress(d[100_000:300_000], 1, include_checksum=True)
    assert z.zstd_decompress(cat, all_frames=True) == d[:300_000]
    obj = {"Integer": list(range(2000)), "Real": [k / 100 for k in range(2000)], "Factor": ["a", "b"] * 1000, "nested": {"x": b"\x00" * 70000}}
    assert z.zstd_unserialize(z.zstd_serialize(obj, level=3)) == obj
    assert z.zstd_unserialize(z.zstd_serialize(obj, level=1, frame_size=16384)) == obj
    import pickle
    assert pickle.loads(ref.decompress(z.zstd_serialize(obj))) == obj


def test_dictionary_compress(z, ref):
    """configs[3]: small objects with a trained dictionary (and a raw-content one).  Every GPU frame must decode with the
    reference's libzstd + the same dictionary, carry the dictionary ID, and the batch must stay within 3% of the size the
    reference produces with that dictionary at the same level (src/cctx.c:296-312 -> ZSTD_CCtx_loadDictionary)."""
    from zstdlite_b200 import corpus
    from tests import emul_util
    from tests.gpu_util import gpu_decompress_batch
    samples = corpus.small_objects(3000)
    trained = ref.train_dict(samples[:1000], 5000)
    big = ref.train_dict(samples[:1000], 112640)
    raw = b"".join(samples[:20])
    work = samples[1000:3000]
    for dd, name in ((trained, "trained 5000 B"), (big, "trained 112640 B"), (raw, "raw content")):
        did = z.zstd_dict_id(dd)
        for lvl in (1, 3):
            res, outs = _gpu_compress_batch(z, work, lvl, dict=dd)
            rd = ref.DCtx(dict=dd)
            for i, (s_, r, c) in enumerate(zip(work, res, outs)):
                assert not z.is_error(r), z.error_name(r)
                assert rd.decompress(c, cap=len(s_)) == s_
                assert z.zstd_info(c)["dict_id"] == did
                if i < 300:
                    assert c == emul_util.compress_frame(s_, lvl, dict=dd), "CUDA output differs from the CPU emulation (dictionary mode)"
            ours = sum(len(c) for c in outs)
            rc = ref.CCtx(level=lvl, dict=dd)
            theirs = sum(len(rc.compress(s_)) for s_ in work)
            nodict = sum(len(ref.compress(s_, lvl)) for s_ in work[:200]) * len(work) / 200
            print(f"dict {name} level {lvl}: ours {ours} libzstd {theirs} (no dict ~{nodict:.0f})")
            assert ours <= theirs * 1.03, (name, lvl, ours, theirs)
            # and our own decoder with the dictionary
            res2, back = gpu_decompress_batch(outs, [len(s_) for s_ in work], dctx=z.zstd_dctx(dict=dd))
            assert back == work
    # larger inputs: only the first block of a frame sees the dictionary; host-pointer one-shot path
    d = corpus.make("text", 300000, 3).tobytes()
    tdict = ref.train_dict([corpus.make("text", 4000, 50 + i).tobytes() for i in range(200)], 20000)
    c = z.zstd_compress(d, level=3, dict=tdict)
    assert ref.decompress(c, dict=tdict) == d and z.zstd_decompress(c, dict=tdict) == d
    assert len(c) <= len(z.zstd_compress(d, level=3)) * 1.01          # an unrelated dictionary must not hurt
    # a frame made with a dictionary does not decode without it
    with pytest.raises(z.ZstdError):
        z.zstd_decompress(z.zstd_compress(work[0], dict=trained))


def test_streaming_entry_points_interoperate_with_one_shot(z, ref):
    """tests/testthat/test-compress-raw.R:61-89: all four combinations of streaming / one-shot writer x reader must agree,
    and the reference's libzstd (one-shot and streaming) must read what the streaming writer produced."""
    import io
    from zstdlite_b200 import corpus
    for d in (b"", b"tiny", corpus.make("text", 131072, 9).tobytes(), corpus.make("rdf", 700001, 9).tobytes()):
        sink = io.BytesIO()
        z.zstd_compress_stream(d, sink.write, level=3, include_checksum=True)
        streamed = sink.getvalue()
        one = z.zstd_compress(d, level=3, include_checksum=True)
        assert streamed == one                                           # same engine, same frame
        assert z.zstd_info(streamed)["uncompressed_size"] == len(d)
        assert ref.decompress(streamed) == d
        for blob in (streamed, ref.compress(d, 3, True), ref.compress(d, 19)):
            assert z.zstd_decompress_stream(io.BytesIO(blob).read) == d
            assert z.zstd_decompress_stream(io.BytesIO(blob).read, out_chunk=1000) == d      # small output chunks
    # two frames + a skippable frame in one stream; truncated input is an error
    a, b = corpus.make("text", 50000, 1).tobytes(), corpus.make("lowent", 200000, 2).tobytes()
    skip = (0x184D2A53).to_bytes(4, "little") + (5).to_bytes(4, "little") + b"hello"
    blob = ref.compress(a, 3, True) + skip + ref.compress(b, 1)
    assert z.zstd_decompress_stream(io.BytesIO(blob).read) == a + b
    with pytest.raises(z.ZstdError):
        z.zstd_decompress_stream(io.BytesIO(blob[:-7]).read)
    # pledged size mismatch is reported like libzstd does (zstd.c:27190)
    L = z._lib.lib()
    c = z.zstd_cctx()
    L.ZSTD_CCtx_setPledgedSrcSize(c._p, 10)
    ib = C.create_string_buffer(b"abc", 3); ob = C.create_string_buffer(100)
    r = L.ZSTD_compressStream2(c._p, C.byref(z._lib.OutBuffer(C.cast(ob, C.c_void_p), 100, 0)), C.byref(z._lib.InBuffer(C.cast(ib, C.c_void_p), 3, 0)), 2)
    assert z.is_error(r) and z.error_name(r) == "Src size is incorrect"


@pytest.mark.parametrize("family,size,bar", [("rdf", 2 << 20, 1.03), ("text", 2 << 20, 1.03), ("text", 16 << 20, 1.03), ("text", (40 << 20) + 12345, 1.03)])
def test_large_single_frame_ratio_against_libzstd(z, ref, family, size, bar):
    """ZSTD_compress2 of ONE large buffer (what zstd_compress / zstd_serialize of a real object calls), level 3, against libzstd on the same
    buffer.  libzstd uses a 2 MB window there (zstd.c:29527); this compressor's blocks are searched independently with offsets <= 64 KiB, plus
    the far candidates of zl_enc_match.cuh (one table of earliest occurrences per 8 MiB region of the frame, offsets < 16 MiB).  Without them
    text-like payloads were 12 % larger than libzstd's ("window beyond one block"); the 40 MiB case covers several regions.
    The frame is the CPU emulation's byte for byte, declares a window of at most 16 MiB, and our own decoder reads it back."""
    from zstdlite_b200 import corpus
    from tests import emul_util
    d = corpus.make(family, size, 21).tobytes()
    ours = z.zstd_compress(d, level=3)
    assert ref.decompress(ours) == d                                       # valid, round-trips through libzstd
    assert ours == emul_util.compress_frame(d, 3), "CUDA output differs from the CPU emulation (far candidates)"
    assert z.zstd_decompress(ours) == d
    assert (ours[4] & 0x20) == 0 and 10 + (ours[5] >> 3) <= 24             # window descriptor: <= 2^24
    theirs = ref.compress(d, 3)
    ratio = len(ours) / len(theirs)
    assert ratio <= bar, (family, size, ratio)


def test_window_log_parameter(z, ref):
    """ZSTD_c_windowLog (src/zstd/zstd.h:347): 17 keeps every match inside its 128 KiB block (no far candidates, the fast setting);
    18..23 bound the far offsets; values under 17 cannot be honoured with 128 KiB blocks and are refused."""
    from zstdlite_b200 import corpus
    L = z._lib.lib()
    d = corpus.make("text", 6 << 20, 33).tobytes()
    sizes = {}
    for wl in (0, 17, 20, 27):
        cc = z.zstd_cctx(level=3)
        assert not z.is_error(L.ZSTD_CCtx_setParameter(cc._p, z._lib.ZSTD_c_windowLog, wl))
        c = z.zstd_compress(d, cctx=cc)
        assert ref.decompress(c) == d and z.zstd_decompress(c) == d
        wlog = 10 + (c[5] >> 3)
        assert wlog == {0: 23, 17: 17, 20: 20, 27: 23}[wl], (wl, wlog)          # 6 MiB needs 2^23 to be covered
        sizes[wl] = len(c)
    assert sizes[0] == sizes[27] < sizes[20] < sizes[17]
    cc = z.zstd_cctx(level=3)
    r = L.ZSTD_CCtx_setParameter(cc._p, z._lib.ZSTD_c_windowLog, 12)
    assert z.is_error(r) and z.error_name(r) == "Parameter is out of bound"


def test_many_small_frames_in_parts(z, ref):
    """A wave of >= 16,384 one-block frames is described and launched in four parts (zl_enc_wave), blocks of <= 2 KiB go to the warp-per-block
    match kernel and the others to the CTA one, in the same wave; with checksums on, XXH64 runs per part too.  Every frame must be what the
    same input gives on its own (the table size of a block depends on the block alone), decode with libzstd and carry the right checksum."""
    from zstdlite_b200 import corpus
    rng = np.random.default_rng(5)
    n = 20000
    sizes = rng.integers(1, 700, n)
    sizes[::97] = rng.integers(2049, 6000, len(sizes[::97]))          # some blocks above the small-block limit, spread over all parts
    sizes[5] = 0
    pool = corpus.make("text", 1 << 20, 3).tobytes() + corpus.make("rdf", 1 << 20, 4).tobytes()
    bufs = [pool[o:o + s] for o, s in zip(rng.integers(0, len(pool) - 6000, n), sizes)]
    res, outs = _gpu_compress_batch(z, bufs, 3, True)
    assert not any(z.is_error(int(r)) for r in res)
    cc = z.zstd_cctx(level=3, include_checksum=True)
    for i in list(range(0, n, 487)) + [5, 97, 194, n - 1]:
        assert ref.decompress(outs[i]) == bufs[i], i
        assert outs[i] == z.zstd_compress(bufs[i], cctx=cc), i
    from tests.gpu_util import gpu_decompress_batch
    res2, back = gpu_decompress_batch(outs, [len(b) for b in bufs])
    assert back == bufs
