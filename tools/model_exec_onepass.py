"""Lane-level model of the one-pass execute kernel (zl_dec_exec.cuh, zl_exec_block) -- development tool.

Restates the warp algorithm with 32 explicit lanes (ballot / shuffle / REDUX written out) and checks it against the
serial definition of ZSTD_execSequence on random sequence lists, including overlapping matches, offsets that fall inside
the current 64-byte step, empty literal runs / matches and every destination alignment.  No GPU, no oracle: it only
validates the index arithmetic before it is written in CUDA.
"""
import random
import sys

LIT = 1 << 31


def serial(prefix, seqs, lits):
    out = bytearray(prefix)
    lp = 0
    for ll, ml, off in seqs:
        out += lits[lp:lp + ll]
        lp += ll
        for _ in range(ml):
            out.append(out[-off])
    return bytes(out), lp


def popc(x):
    return bin(x & 0xFFFFFFFF).count("1")


def warp_batch(out, out_pos, out_addr, recs, lits, lit_pos):
    """one batch of <= 32 (ll, ml, off) records; `out` is a bytearray already holding [0, out_pos)."""
    n = len(recs)
    ll = [recs[i][0] if i < n else 0 for i in range(32)]
    ml = [recs[i][1] if i < n else 0 for i in range(32)]
    off = [recs[i][2] if i < n else 0 for i in range(32)]
    sl, so = [0] * 32, [0] * 32
    a = b = 0
    for i in range(32):
        a += ll[i]; b += ll[i] + ml[i]
        sl[i], so[i] = a, b
    totalL, totalO = sl[31], so[31]
    bL = sum((1 << i) for i in range(32) if ll[i])
    bM = sum((1 << i) for i in range(32) if ml[i])
    seg = [(0, 0)] * 64
    fL, fM = [0] * 32, [0] * 32
    for i in range(32):
        lt = (1 << i) - 1
        ordL = popc(bL & lt) + popc(bM & lt)
        ordM = ordL + (1 if ll[i] else 0)
        fL[i] = so[i] - ll[i] - ml[i]
        fM[i] = so[i] - ml[i]
        if ll[i]:
            seg[ordL] = (fL[i] | LIT, (lit_pos + sl[i] - ll[i] - fL[i]) & 0xFFFFFFFF)
        if ml[i]:
            seg[ordM] = (fM[i], off[i])
    mis = (out_addr + out_pos) & 31
    out.extend(b"\0" * totalO)
    cnt = 0
    a0 = 0
    steps = fix_iters = 0
    while a0 < totalO + mis:
        steps += 1
        r = a0 >> 5
        m = [0, 0]
        for k in range(2):
            for i in range(32):
                if ll[i] and ((fL[i] + mis) >> 5) == r + k:
                    m[k] |= 1 << ((fL[i] + mis) & 31)
                if ml[i] and ((fM[i] + mis) >> 5) == r + k:
                    m[k] |= 1 << ((fM[i] + mis) & 31)
        j0 = a0 - mis
        val = [[0, 0] for _ in range(32)]
        res = [[True, True] for _ in range(32)]
        srcj = [[0, 0] for _ in range(32)]
        act = [[False, False] for _ in range(32)]
        c_base = cnt
        for k in range(2):
            for lane in range(32):
                le = (2 << lane) - 1
                c = c_base + popc(m[k] & le) - 1
                j = j0 + 32 * k + lane
                ok = 0 <= j < totalO
                act[lane][k] = ok
                if not ok:
                    continue
                x, y = seg[c & 63]
                if x & LIT:
                    d = y - (1 << 32) if y >= (1 << 31) else y
                    val[lane][k] = lits[j + d]
                else:
                    S = x
                    assert S <= j, (S, j)
                    sj = j - y
                    if sj >= max(j0, 0):           # in-step source
                        if sj >= S:                 # inside its own match: periodic form
                            sj = S - y + (j - S) % y
                        if sj >= max(j0, 0):
                            res[lane][k] = False
                            srcj[lane][k] = sj
                            continue
                    val[lane][k] = out[out_pos + sj]
            c_base += popc(m[k])
        cnt = c_base
        # in-step resolution by shuffles
        while any(not res[l][k] for l in range(32) for k in range(2)):
            fix_iters += 1
            snap_val = [v[:] for v in val]
            snap_res = [v[:] for v in res]
            for lane in range(32):
                for k in range(2):
                    if act[lane][k] and not res[lane][k]:
                        rel = srcj[lane][k] - j0
                        sl_, sh_ = rel & 31, rel >> 5
                        assert 0 <= rel < 64 and (sh_, sl_) < (k, lane)
                        if snap_res[sl_][sh_]:
                            val[lane][k] = snap_val[sl_][sh_]
                            res[lane][k] = True
        for k in range(2):
            for lane in range(32):
                if act[lane][k]:
                    out[out_pos + j0 + 32 * k + lane] = val[lane][k]
        a0 += 64
    return totalL, totalO, steps, fix_iters


def run(seqs, lits, prefix, out_addr):
    out = bytearray(prefix)
    pos, lp = len(prefix), 0
    for b in range(0, len(seqs), 32):
        tl, to, _, _ = warp_batch(out, pos, out_addr, seqs[b:b + 32], lits, lp)
        pos += to; lp += tl
    return bytes(out), lp


def random_case(rng):
    prefix = bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 7, 40, 300])))
    nseq = rng.choice([1, 5, 31, 32, 33, 64, 100])
    seqs, pos = [], len(prefix)
    style = rng.choice(["mixed", "short", "rle", "periodic"])
    for _ in range(nseq):
        ll = rng.choice([0, 0, 1, 2, 3, 5, 9, 40, 100]) if style != "periodic" else rng.choice([0, 0, 0, 8])
        if pos + ll == 0:
            ll = 1
        pos += ll
        if style == "rle":
            off = rng.choice([1, 1, 2, 3]); ml = rng.choice([3, 30, 70, 200])
        elif style == "periodic":
            off = min(pos, 8 * rng.randrange(1, 6)); ml = rng.choice([8, 8, 16, 24])
        elif style == "short":
            off = rng.randrange(1, min(pos, 40) + 1); ml = rng.choice([3, 4, 5, 8])
        else:
            off = rng.randrange(1, pos + 1); ml = rng.choice([0, 3, 4, 7, 12, 33, 64, 65, 130])
        off = max(1, min(off, pos))
        seqs.append((ll, ml, off)); pos += ml
    lits = bytes(rng.randrange(256) for _ in range(sum(s[0] for s in seqs)))
    return prefix, seqs, lits


if __name__ == "__main__":
    rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    for t in range(n):
        prefix, seqs, lits = random_case(rng)
        want, _ = serial(prefix, seqs, lits)
        for addr in (0, 1, 13, 31):
            got, _ = run(seqs, lits, prefix, addr)
            assert got == want, (t, addr, seqs[:8])
    print("model ok:", n, "cases x 4 alignments")
