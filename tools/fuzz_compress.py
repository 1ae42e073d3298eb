"""Randomised round-trip stress of the CUDA compressor: random sizes (0 .. 400 KB), families, levels, checksum flags, mixed in
one batch; every frame must decode with the reference's libzstd to the input, equal the CPU emulation byte for byte, and decode
with the CUDA decoder.  Run on the GPU box, optionally under compute-sanitizer:  python tools/fuzz_compress.py 400"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(n):
    import torch
    import zstdlite_b200 as z
    from oracle import ref, restate
    from zstdlite_b200 import corpus
    from tests import emul_util
    from tests.gpu_util import to_dev, gpu_decompress_batch
    rng = np.random.default_rng(99)
    fams = ("text", "rdf", "lowent", "rand", "rle")
    sizes_pool = [0, 1, 7, 8, 100, 1000, 16383, 16384, 16385, 32768, 65535, 65536, 131071, 131072, 131073, 147456, 262144, 300000]
    bad = 0
    for lvl in (1, 2, 3):
        for ck in (False, True):
            bufs = []
            for i in range(n // 6):
                size = int(rng.choice(sizes_pool)) if rng.integers(0, 2) else int(rng.integers(0, 400000))
                d = corpus.make(fams[int(rng.integers(0, 5))], max(size, 1), int(rng.integers(0, 1000))).tobytes()[:size]
                if rng.integers(0, 8) == 0 and size > 64:      # splice: a repeat of an earlier part far behind, and a run
                    k = int(rng.integers(1, size // 2)); d = d[:size - k] + d[:k]
                bufs.append(d)
            L = z._lib.lib()
            cctx = z.zstd_cctx(level=lvl, include_checksum=ck)
            caps = [int(L.ZSTD_compressBound(len(b))) for b in bufs]
            offs = np.concatenate([[0], np.cumsum([len(b) for b in bufs])]).astype(np.int64)
            src = to_dev(np.frombuffer(b"".join(bufs), dtype=np.uint8))
            doffs = np.concatenate([[0], np.cumsum([(c + 15) // 16 * 16 for c in caps])]).astype(np.int64)
            dst = torch.zeros(int(doffs[-1]) + 64, dtype=torch.uint8, device="cuda")
            res = z.compress_batch(cctx, [src.data_ptr() + int(o) for o in offs[:-1]], [len(b) for b in bufs],
                                   [dst.data_ptr() + int(o) for o in doffs[:-1]], caps, device=True)
            host = dst.cpu().numpy()
            frames = []
            for i, b in enumerate(bufs):
                if z.is_error(res[i]):
                    print("compress error", lvl, ck, len(b), z.error_name(res[i])); bad += 1; frames.append(b""); continue
                c = host[int(doffs[i]):int(doffs[i]) + res[i]].tobytes()
                frames.append(c)
                if ref.decompress(c) != b:
                    print("libzstd round trip differs", lvl, ck, len(b)); bad += 1
                e = emul_util.compress_frame(b, lvl, ck, restate.xxh64(b) & 0xFFFFFFFF)
                if e != c:
                    print("differs from the CPU emulation", lvl, ck, len(b)); bad += 1
            r2, outs = gpu_decompress_batch(frames, [len(b) for b in bufs])
            for b, o, r in zip(bufs, outs, r2):
                if z.is_error(r) or o != b:
                    print("CUDA decoder round trip differs", lvl, ck, len(b)); bad += 1
    print(f"fuzz_compress: {6 * (n // 6)} frames, {bad} problems")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 300))
