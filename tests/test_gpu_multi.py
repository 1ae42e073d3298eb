"""-m gpu tests that need TWO devices (skipped on a one-GPU box): one process driving several GPUs through the C ABI.

* a context belongs to the device that was current at its first use; per-device state (the encoder's constant tables, occupancy
  limits, function attributes) is set up for every device a process touches (compress on device 1 after device 0);
* batches of host buffers are spread over the GPUs by the library itself (contiguous frame ranges, no collective):
  zstd_dctx(num_gpus = N) / zstd_cctx(num_threads = N) -- results must be what one GPU gives."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import zstdlite_b200 as z
    from oracle import ref
    from zstdlite_b200 import corpus
    return torch, z, ref, corpus


def test_second_device_after_first(env):
    torch, z, ref, corpus = env
    raws = [corpus.make(f, 100000 + 7 * i, i).tobytes() for i, f in enumerate(("text", "rdf", "lowent", "rand"))]
    outs = {}
    for dev in (0, 1, 0):
        torch.cuda.set_device(dev)
        for lvl in (1, 3):
            cc, dc = z.zstd_cctx(level=lvl, include_checksum=True), z.zstd_dctx()
            for k, d in enumerate(raws):
                c = z.zstd_compress(d, cctx=cc)
                assert ref.decompress(c) == d, (dev, lvl, k)
                assert z.zstd_decompress(c, dctx=dc) == d
                assert outs.setdefault((lvl, k), c) == c                  # same bytes whichever device made them
    torch.cuda.set_device(0)


def test_context_keeps_its_device(env):
    torch, z, ref, corpus = env
    d = corpus.make("text", 300000, 9).tobytes()
    torch.cuda.set_device(1)
    cc, dc = z.zstd_cctx(level=3), z.zstd_dctx()
    c = z.zstd_compress(d, cctx=cc)
    torch.cuda.set_device(0)                                              # the caller moved on; the contexts did not
    assert z.zstd_compress(d, cctx=cc) == c and z.zstd_decompress(c, dctx=dc) == d
    assert torch.cuda.current_device() == 0


def test_batches_spread_over_two_gpus(env):
    torch, z, ref, corpus = env
    import ctypes as C
    n, fb = 2048, 65536
    data, _ = corpus.mixed_frames(n, fb, mix=(("text", 0.4), ("rdf", 0.4), ("lowent", 0.1), ("rand", 0.1)), pool=16, rotate=True)
    frames = [ref.compress(data[i].tobytes(), 3, include_checksum=(i % 2 == 0)) for i in range(n)]
    sbufs = [C.create_string_buffer(f, len(f)) for f in frames]

    def decode(dctx):
        dbufs = [C.create_string_buffer(fb) for _ in range(n)]
        res = z.decompress_batch(dctx, [C.addressof(b) for b in sbufs], [len(f) for f in frames], [C.addressof(b) for b in dbufs], [fb] * n, device=False)
        assert all(int(r) == fb for r in res)
        return b"".join(b.raw for b in dbufs)
    want = data.tobytes()
    assert decode(z.zstd_dctx()) == want
    assert decode(z.zstd_dctx(num_gpus=2)) == want
    # one damaged frame (it carries a checksum) in the second device's range: its error is reported in place, the others are unaffected
    bad = bytearray(frames[n - 6]); bad[len(bad) // 2] ^= 0x55; sbufs[n - 6] = C.create_string_buffer(bytes(bad), len(bad))
    dbufs = [C.create_string_buffer(fb) for _ in range(n)]
    res = z.decompress_batch(z.zstd_dctx(num_gpus=2), [C.addressof(b) for b in sbufs], [len(f) for f in frames], [C.addressof(b) for b in dbufs], [fb] * n, device=False)
    assert z.is_error(int(res[n - 6])) and all(int(r) == fb for i, r in enumerate(res) if i != n - 6)
    # compress: num_threads = 2 -> two GPUs; identical frames to one GPU's
    raws = [C.create_string_buffer(data[i].tobytes(), fb) for i in range(n)]
    cap = fb + 1024

    def encode(cctx):
        outs = [C.create_string_buffer(cap) for _ in range(n)]
        res = z.compress_batch(cctx, [C.addressof(b) for b in raws], [fb] * n, [C.addressof(b) for b in outs], [cap] * n, device=False)
        assert not any(z.is_error(int(r)) for r in res)
        return [o.raw[:int(r)] for o, r in zip(outs, res)]
    one, two = encode(z.zstd_cctx(level=3)), encode(z.zstd_cctx(level=3, num_threads=2))
    assert one == two
    for i in range(0, n, 97):
        assert ref.decompress(two[i]) == data[i].tobytes()
