"""-m gpu tests at BASELINE.json's full sizes, through size-independent properties (compress -> decompress round trips
compared on the device, sampled frames cross-decoded by the reference's libzstd, ratio against libzstd on the same slabs).

configs[2]: a 4 GiB raw buffer as 32,768 independent 128 KiB frames at levels 1 and 3.
configs[4]: one GPU's share of the 32 GiB mixed corpus when 8 GPUs take contiguous frame ranges (4 GiB, families
            text/rdf/lowent/rand/rle 30/30/15/15/10), compress (level 3, checksums on) then decompress.
The slabs are made on the device from a pool of distinct slabs per family, every copy rotated by a different number of
bytes, so that no two frames of the buffer are equal."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FB = 131072
MIX3 = (("text", 0.4), ("rdf", 0.4), ("lowent", 0.1), ("rand", 0.1))
MIX5 = (("text", 0.3), ("rdf", 0.3), ("lowent", 0.15), ("rand", 0.15), ("rle", 0.1))


@pytest.fixture(scope="module")
def z():
    import torch
    assert torch.cuda.is_available()
    import zstdlite_b200 as zz
    return zz


def _round_trip(z, ref, nslabs, mix, level, checksum):
    import torch
    from zstdlite_b200 import corpus
    src, pools, meta = corpus.device_mixed_slabs(nslabs, FB, mix)
    L = z._lib.lib()
    bound = int(L.ZSTD_compressBound(FB))
    slot = (bound + 255) // 256 * 256
    comp = torch.empty(nslabs * slot + 64, dtype=torch.uint8, device="cuda")
    cctx = z.zstd_cctx(level=level, include_checksum=checksum)
    res = z.compress_batch(cctx, [src.data_ptr() + i * FB for i in range(nslabs)], [FB] * nslabs,
                           [comp.data_ptr() + i * slot for i in range(nslabs)], [bound] * nslabs, device=True)
    sizes = np.array(res, dtype=np.uint64)
    assert not any(z.is_error(int(s)) for s in sizes), "compress errors"
    assert int(sizes.max()) <= bound
    # every frame through the GPU decoder, compared on the device
    back = torch.empty((nslabs, FB), dtype=torch.uint8, device="cuda")
    dres = z.decompress_batch(z.zstd_dctx(), [comp.data_ptr() + i * slot for i in range(nslabs)], [int(s) for s in sizes],
                              [back.data_ptr() + i * FB for i in range(nslabs)], [FB] * nslabs, device=True)
    assert all(int(r) == FB for r in dres), "decode errors"
    assert torch.equal(back, src), "round trip differs"
    del back
    # a sample of frames through the reference's libzstd (headers, checksums, content), and the ratio on the sample
    ours = theirs = 0
    for i in range(0, nslabs, max(1, nslabs // 96)):
        fam, row, shift = meta[i]
        want = np.roll(pools[fam][row], shift).tobytes()
        frame = comp[i * slot:i * slot + int(sizes[i])].cpu().numpy().tobytes()
        info = z.zstd_info(frame)
        assert info["uncompressed_size"] == FB and info["compressed_size"] == len(frame) and info["has_checksum"] == checksum
        assert ref.decompress(frame) == want, f"frame {i} ({fam}) does not decode with libzstd"
        ours += len(frame); theirs += len(ref.compress(want, level))
    assert ours <= theirs * 1.03, (ours, theirs)
    return float(nslabs) * FB / float(sizes.sum())


@pytest.mark.parametrize("level", [1, 3])
def test_config3_4GiB_as_128KiB_frames_round_trips(z, ref, level):
    ratio = _round_trip(z, ref, 32768, MIX3, level, checksum=False)
    assert 2.0 < ratio < 3.5


def test_config5_one_gpu_share_of_the_32GiB_corpus_round_trips(z, ref):
    ratio = _round_trip(z, ref, 32768, MIX5, 3, checksum=True)
    assert ratio > 2.0
