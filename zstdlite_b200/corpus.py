"""Deterministic synthetic corpora.  Pure numpy; used by tests and bench.

Families: text (Zipf words), rdf (columnar f64/i32/factor slab, as man/benchmark.R:14-21),
lowent (4-bit entropy bytes), rand (uniform bytes), rle (long runs).
"""
import numpy as np

FAMILIES = ("text", "rdf", "lowent", "rand", "rle")


def _text(rng, n):
    vocab_n = 4096
    lens = rng.integers(2, 10, size=vocab_n)
    letters = rng.integers(97, 123, size=int(lens.sum()), dtype=np.uint8)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    ranks = np.arange(1, vocab_n + 1, dtype=np.float64)
    p = 1.0 / ranks
    p /= p.sum()
    need = n // 5 + 64
    ids = rng.choice(vocab_n, size=need, p=p)
    wl = lens[ids] + 1
    total = int(wl.sum())
    while total < n:
        more = rng.choice(vocab_n, size=need, p=p)
        ids = np.concatenate([ids, more])
        wl = lens[ids] + 1
        total = int(wl.sum())
    out = np.full(total, 32, dtype=np.uint8)
    pos = np.concatenate([[0], np.cumsum(wl)[:-1]])
    # vectorised scatter of word letters
    maxlen = int(lens.max())
    for k in range(maxlen):
        m = lens[ids] > k
        out[pos[m] + k] = letters[starts[ids[m]] + k]
    return out[:n]


def _rdf(rng, n):
    # byte proportion 8:4:4 (f64 : i32 : i32 factor codes), column-major inside the slab
    rows = max(1, n // 16)
    real = (rng.integers(0, 20, size=rows).astype(np.float64) / 100.0).astype("<f8").view(np.uint8)
    integer = rng.integers(1, rows + 1, size=rows).astype("<i4").view(np.uint8)
    factor = rng.integers(1, 11, size=rows).astype("<i4").view(np.uint8)
    out = np.concatenate([real, integer, factor])
    if out.size < n:
        out = np.concatenate([out, np.zeros(n - out.size, dtype=np.uint8)])
    return out[:n]


def _lowent(rng, n):
    return rng.integers(0, 16, size=n, dtype=np.uint8)


def _rand(rng, n):
    return rng.integers(0, 256, size=n, dtype=np.uint8)


def _rle(rng, n):
    out = np.empty(n, dtype=np.uint8)
    pos = 0
    while pos < n:
        run = int(rng.geometric(1.0 / 4096.0))
        out[pos:pos + run] = rng.integers(0, 256)
        pos += run
    return out


_GEN = {"text": _text, "rdf": _rdf, "lowent": _lowent, "rand": _rand, "rle": _rle}
_SEED = {"text": 1, "rdf": 2, "lowent": 3, "rand": 4, "rle": 5}


def make(family, nbytes, index=0):
    """One buffer of `nbytes` bytes of the named family; `index` varies the stream."""
    rng = np.random.default_rng([_SEED[family], index])
    return _GEN[family](rng, int(nbytes))


def mixed_frames(nframes, frame_bytes, mix=(("text", 0.4), ("rdf", 0.4), ("lowent", 0.1), ("rand", 0.1)), pool=64, rotate=False):
    """Config-2/3 style corpus: returns (uint8 array [nframes, frame_bytes], family name per frame).

    To keep generation fast, each family is generated as `pool` distinct frames and tiled; every
    frame is still decoded/encoded independently so throughput is unaffected by the repetition.
    rotate=True rotates the k-th copy of a pool frame by 24*k bytes, so that all frames differ.
    """
    fams = []
    out = np.empty((nframes, frame_bytes), dtype=np.uint8)
    start = 0
    for k, (fam, frac) in enumerate(mix):
        cnt = nframes - start if k == len(mix) - 1 else int(round(nframes * frac))
        npool = min(pool, max(cnt, 1))
        base = make(fam, npool * frame_bytes, index=7).reshape(npool, frame_bytes)
        for j in range(cnt):
            out[start + j] = np.roll(base[j % npool], (j // npool) * 24 % frame_bytes) if rotate else base[j % npool]
        fams += [fam] * cnt
        start += cnt
    return out, fams


def small_objects(n, seed=4, lo=10, hi=50):
    """Config-4 style corpus: `n` R-serialized-like records of ~0.2-1 KB -- a named integer vector over a
    permutation of 50 country names (README.md:283-293) behind a short serialization header."""
    rng = np.random.default_rng(seed)
    names = [("country_%02d" % i).encode() for i in range(50)]
    out = []
    for _ in range(n):
        perm = rng.permutation(50)
        k = int(rng.integers(lo, hi))
        out.append(b"X\n\x00\x00\x00\x03" + b"".join(names[j] + int(rng.integers(0, 1000)).to_bytes(4, "little") for j in perm[:k]))
    return out


def r_data_frame(rows=1_000_000, seed=6):
    """Config-1 payload: a restatement of R's binary serialization (R_pstream_binary_format, version 3,
    src/serialize.c:53-62) of a data.frame with columns Integer (i32), Real (f64, 20 distinct values as in man/benchmark.R:19)
    and Factor (i32 codes 1..10 with `levels` / `class` attributes), followed by the names / class / compact row.names
    attributes.  The attribute bytes are < 1 KB of ~16 MB; the column layout is what the codec sees."""
    import struct
    rng = np.random.default_rng(seed)

    def i32(v):
        return struct.pack("<i", v)

    def charsxp(t):
        b = t.encode()
        return i32(0x00040009) + i32(len(b)) + b                 # CHARSXP, UTF-8 flag

    def strsxp(items):
        return i32(16) + i32(len(items)) + b"".join(charsxp(t) for t in items)

    def sym(name):
        return i32(1) + charsxp(name)                            # SYMSXP

    def attr(name, value, first=True):
        return i32(0x402) + sym(name) + value                    # LISTSXP with tag

    nil = i32(254)
    head = b"B\n" + i32(3) + i32(0x00040301) + i32(0x00030500) + i32(5) + b"UTF-8"
    integer = i32(13) + i32(rows) + rng.integers(1, rows + 1, size=rows).astype("<i4").tobytes()
    real = i32(14) + i32(rows) + (rng.integers(0, 20, size=rows).astype(np.float64) / 100.0).astype("<f8").tobytes()
    levels = ["Athens", "Barcelona", "Brussels", "Calais", "Cherbourg", "Cologne", "Copenhagen", "Geneva", "Gibraltar", "Hamburg"]
    factor = (i32(13 | 0x300) + i32(rows) + rng.integers(1, 11, size=rows).astype("<i4").tobytes()
              + attr("levels", strsxp(levels)) + attr("class", strsxp(["factor"])) + nil)
    body = (i32(19 | 0x300) + i32(3) + integer + real + factor
            + attr("names", strsxp(["Integer", "Real", "Factor"])) + attr("class", strsxp(["data.frame"]))
            + attr("row.names", i32(13) + i32(2) + i32(-2147483648) + i32(-rows)) + nil)
    return head + body


def device_mixed_slabs(nslabs, slab_bytes, mix, pool=48, index=11, device="cuda"):
    """Config-3/5 style corpus made ON THE DEVICE (torch is plumbing here): `pool` distinct slabs per family are generated on
    the host, uploaded, and every copy is rotated by a different number of bytes, so that no two slabs are equal.
    -> (uint8 tensor [nslabs, slab_bytes], {family: host array [pool, slab_bytes]}, per-slab (family, pool row, shift));
    slab i equals np.roll(pools[family][row], shift)."""
    import torch
    out = torch.empty((nslabs, slab_bytes), dtype=torch.uint8, device=device)
    pools, meta = {}, []
    start = 0
    col = torch.arange(slab_bytes, device=device, dtype=torch.int64)
    for k, (fam, frac) in enumerate(mix):
        cnt = nslabs - start if k == len(mix) - 1 else int(round(nslabs * frac))
        base = make(fam, pool * slab_bytes, index=index).reshape(pool, slab_bytes)
        pools[fam] = base
        dbase = torch.from_numpy(base.copy()).to(device)
        for c0 in range(0, cnt, 1024):
            c1 = min(cnt, c0 + 1024)
            j = torch.arange(c0, c1, device=device, dtype=torch.int64)
            row, shift = j % pool, (j // pool) * 13 % slab_bytes
            idx = (col[None, :] - shift[:, None]) % slab_bytes                     # np.roll(slab, shift)
            out[start + c0:start + c1] = torch.gather(dbase[row], 1, idx)
        meta += [(fam, j % pool, (j // pool) * 13 % slab_bytes) for j in range(cnt)]
        start += cnt
    return out, pools, meta
