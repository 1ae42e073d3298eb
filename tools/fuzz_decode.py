"""Fuzz the CUDA decoder with mutated frames (bit flips, byte splats, truncations, spliced sections) and compare the verdict
with the reference's libzstd.  Meant to be run under compute-sanitizer on the GPU box:
    compute-sanitizer --tool memcheck python tools/fuzz_decode.py 2000
Every mutated frame must give an error or exactly libzstd's output; the sanitizer must stay silent.
Known, deliberate difference: a damaged 4-stream Huffman section whose reader runs past the START of its stream is an error
here; libzstd's fast loop keeps reading the bytes in front of the stream and returns garbage (zstd.c:38772-38950)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(n):
    import zstdlite_b200 as z
    from oracle import ref
    from zstdlite_b200 import corpus
    from tests.gpu_util import gpu_decompress_batch
    rng = np.random.default_rng(2024)
    bases = []
    sizes = (("text", 20000, 3), ("rdf", 30000, 1), ("lowent", 9000, 3), ("rle", 50000, 3), ("text", 200000, 19), ("text", 300, 3))
    if "--large" in sys.argv:              # frames of >= 1 MiB take the block-parallel execute path (zl_dec_large.cuh)
        sizes = (("text", 1200000, 3), ("rdf", 1500000, 1), ("rle", 2000000, 3))
    for fam, size, lvl in sizes:
        d = corpus.make(fam, size, 5).tobytes()
        bases.append((d, ref.compress(d, lvl, True)))
        bases.append((d, ref.compress(d, lvl, False)))          # no checksum: corrupted payloads that still parse must decode identically
    frames, caps = [], []
    for i in range(n):
        d, c = bases[i % len(bases)]
        m = bytearray(c)
        kind = int(rng.integers(0, 5))
        if kind == 0:
            for _ in range(int(rng.integers(1, 4))):
                m[int(rng.integers(0, len(m)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            k = int(rng.integers(0, len(m))); m[k:k + int(rng.integers(1, 9))] = bytes(int(rng.integers(0, 256)) for _ in range(1))
        elif kind == 2:
            m = m[:int(rng.integers(1, len(m)))]
        elif kind == 3:
            a, b = sorted(int(x) for x in rng.integers(0, len(m), 2)); m = m[:a] + m[b:]
        else:
            k = int(rng.integers(5, min(len(m), 40))); m[k] = int(rng.integers(0, 256))          # header area
        frames.append(bytes(m)); caps.append(len(d) if rng.integers(0, 4) else int(rng.integers(0, len(d))))
    res, outs = gpu_decompress_batch(frames, caps)
    bad = stricter = 0
    for f, cap, r, o in zip(frames, caps, res, outs):
        try:
            want = ref.DCtx().decompress(f, cap=cap, all_frames=True)
        except ref.RefError:
            want = None
        if want is None:
            if not z.is_error(r):
                bad += 1                      # accepted something libzstd rejects
        elif z.is_error(r):
            stricter += 1                     # libzstd decodes it (to garbage: no checksum), we report corruption
        elif o != want:
            bad += 1                          # both accept, different bytes
    print(f"fuzz: {n} frames, {sum(1 for r in res if z.is_error(r))} rejected, {bad} disagreements with libzstd, "
          f"{stricter} rejected here but decoded (to garbage) by libzstd")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 1000))
