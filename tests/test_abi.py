"""CPU tests: the C-ABI library loads without a GPU, exports every symbol include/zstdlite_gpu.h declares (canonical
name and zlg_ alias for the libzstd entry points), answers the host-only calls, and fails loudly -- never falls back
to a CPU path -- when a compute call is made without a CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "zstdlite_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b((?:ZSTD|ZDICT|zl)_[A-Za-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if not n.startswith("ZSTD_CONTENTSIZE")))


def test_every_declared_symbol_is_exported():
    from zstdlite_b200 import _lib
    L = _lib.lib()
    assert L._missing == []
    decl = _declared()
    assert len(decl) >= 35
    for name in decl:
        assert hasattr(L, name), name
        if name.startswith(("ZSTD_", "ZDICT_")):
            assert hasattr(L, "zlg_" + name), "missing alias zlg_" + name
    assert set(_lib.EXPORTED_SYMBOLS) == set(decl)


def test_host_only_entry_points():
    import zstdlite_b200 as z
    from zstdlite_b200 import _lib
    L = _lib.lib()
    assert z.zstd_version() == "1.5.6"
    assert b"no CPU fallback" in L.zl_backend_string()
    # ZSTD_compressBound: zstd.c:4548
    assert L.ZSTD_compressBound(0) == 64 and L.ZSTD_compressBound(65536) == 65824 and L.ZSTD_compressBound(131072) == 131584
    assert L.ZSTD_compressBound(1 << 30) == (1 << 30) + (1 << 22)
    # errors: codes and the strings the reference's tests grep for (tests/testthat/test-checksums.R:26)
    err = lambda code: (1 << 64) - code
    assert L.ZSTD_isError(err(22)) and not L.ZSTD_isError(12345) and not L.ZSTD_isError(err(121))
    assert L.ZSTD_getErrorName(err(22)) == b"Restored data doesn't match checksum"
    assert L.ZSTD_getErrorName(err(70)) == b"Destination buffer is too small"
    assert L.ZSTD_getErrorName(err(20)) == b"Data corruption detected"
    assert L.ZSTD_getErrorName(0) == b"No error detected"
    # frame introspection on the reference's known-answer file (zstd-info.c:60-87)
    c = open(os.path.join(ROOT, "tests", "golden", "data.json.zst"), "rb").read()
    assert z.zstd_info(c) == {"uncompressed_size": 172, "compressed_size": 126, "dict_id": 0, "has_checksum": True}
    assert L.ZSTD_getFrameContentSize(c, 4) == _lib.CONTENTSIZE_ERROR
    assert L.ZSTD_isError(L.ZSTD_findFrameCompressedSize(c, 100))
    assert L.ZSTD_getFrameContentSize(b"\x00\x01\x02\x03\x04\x05\x06\x07", 8) == _lib.CONTENTSIZE_ERROR
    skip = (0x184D2A5F).to_bytes(4, "little") + (3).to_bytes(4, "little") + b"abc"
    assert L.ZSTD_findFrameCompressedSize(skip, len(skip)) == 11
    d = open(os.path.join(ROOT, "tests", "golden", "sample_dict.raw"), "rb").read()
    assert z.zstd_dict_id(d) == 1895278874
    # ZSTD_findDecompressedSize (zstd.c:41244): sums frames, skips skippable frames, trailing garbage is an error
    assert L.ZSTD_findDecompressedSize(c, len(c)) == 172
    assert L.ZSTD_findDecompressedSize(c + skip + c, 2 * len(c) + len(skip)) == 344
    assert L.ZSTD_findDecompressedSize(c + b"\x00\x01", len(c) + 2) == _lib.CONTENTSIZE_ERROR
    assert L.ZSTD_findDecompressedSize(c[:100], 100) == _lib.CONTENTSIZE_ERROR
    assert L.ZSTD_findDecompressedSize(b"", 0) == 0


def test_context_parameters_round_trip():
    """src/cctx.c:257-290,343-369; src/dctx.c:156-173,228-245"""
    import zstdlite_b200 as z
    from zstdlite_b200 import _lib
    L = _lib.lib()
    c = z.zstd_cctx(level=2, num_threads=4, include_checksum=True)
    assert c.settings() == {"level": 2, "num_threads": 4, "include_checksum": 1}
    assert z.zstd_cctx(level=-99).settings()["level"] == -5                           # clamp, src/cctx.c:261-268
    # levels 4 and 5 run the level-3 engine (within 3 % of libzstd at those levels); 6..22 are not implemented: refused
    # (parameter_unsupported -> "Bad compression level"), never served under another label ...
    with pytest.raises(z.ZstdError, match="Bad compression level"):
        z.zstd_cctx(level=9)
    assert L.ZSTD_getErrorName(L.ZSTD_CCtx_setParameter(c._p, 100, 6)) == b"Unsupported parameter" and c.settings()["level"] == 2
    five = z.zstd_cctx(level=5)
    assert five.settings()["level"] == 5 and L.zl_cctx_engine_level(five._p) == 3
    # ... unless the caller opts into the level-3 engine for them; the label stays what was set, the engine is reported beside it
    f = z.zstd_cctx(level=99, level_fallback=True)
    fast = z.zstd_cctx(level=-3)
    assert f.settings()["level"] == 22 and L.zl_cctx_engine_level(f._p) == 3 and L.zl_cctx_engine_level(fast._p) == 1
    assert z.zstd_cctx().settings() == {"level": 3, "num_threads": 0, "include_checksum": 0}
    assert L.ZSTD_isError(L.ZSTD_CCtx_setParameter(c._p, 201, 7))                    # checksumFlag out of bounds
    assert L.ZSTD_isError(L.ZSTD_CCtx_setParameter(c._p, 12345, 1))                  # unknown parameter
    L.ZSTD_CCtx_reset(c._p, 3)
    assert c.settings() == {"level": 3, "num_threads": 0, "include_checksum": 0}
    with pytest.warns(UserWarning, match="Unknown option"):
        z.zstd_cctx(bogus=1)
    d = z.zstd_dctx(validate_checksum=False)
    assert d.settings() == {"validate_checksum": 1}                                  # reference quirk: reports forceIgnoreChecksum
    assert z.zstd_dctx().settings() == {"validate_checksum": 0}


def test_compute_calls_fail_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import zstdlite_b200 as z
    c = open(os.path.join(ROOT, "tests", "golden", "data.json.zst"), "rb").read()
    with pytest.raises(z.ZstdError):
        z.zstd_decompress(c)
    with pytest.raises(z.ZstdError):
        z.zstd_compress(b"no silent CPU path" * 100)
    with pytest.raises(z.ZstdError, match="Training error"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            z.zstd_train_dict_compress([bytes([i % 251] * 40) for i in range(100)], 1024)
