"""One-slice (whole-batch) decode launches for ncu: every call launches each decode kernel exactly once (development tool).
usage: prof_exec.py nframes frame_bytes mix calls   (run under: ncu -k regex:zl_k_execute -s 2 -c 1 ...)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zstdlite_b200 as z
from zstdlite_b200 import corpus
from oracle import ref

nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
fb = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
mixname = sys.argv[3] if len(sys.argv) > 3 else "mix"
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 4
mix = {"mix": (("text", 0.4), ("rdf", 0.4), ("lowent", 0.1), ("rand", 0.1)), "text": (("text", 1.0),), "rdf": (("rdf", 1.0),),
       "lowent": (("lowent", 1.0),), "rand": (("rand", 1.0),), "rle": (("rle", 1.0),)}[mixname]
data, fams = corpus.mixed_frames(nframes, fb, mix=mix, pool=64)
cache, frames = {}, []
for i in range(nframes):
    k = data[i].tobytes()
    if k not in cache: cache[k] = ref.compress(k, 3)
    frames.append(cache[k])
perm = np.random.default_rng(1).permutation(nframes)
frames = [frames[i] for i in perm]; data = data[perm]
sizes = [len(f) for f in frames]
offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
src = torch.from_numpy(np.frombuffer(b"".join(frames), dtype=np.uint8).copy()).cuda()
src = torch.cat([src, torch.zeros(64, dtype=torch.uint8, device="cuda")])
dst = torch.zeros(nframes * fb + 64, dtype=torch.uint8, device="cuda")
d = z.zstd_dctx()
d.set_profile(True)
plan = z.BatchPlan([src.data_ptr() + int(o) for o in offs[:-1]], sizes, [dst.data_ptr() + i * fb for i in range(nframes)], [fb] * nframes)
L = z._lib.lib()
acc = np.zeros(4)
for it in range(calls):
    res = plan.decompress(d)
    if it: acc += [L.zl_dctx_last_stage_ms(d._p, k) for k in range(4)]
ok = bool((dst[:nframes * fb].cpu().numpy().reshape(nframes, fb) == data).all())
print(f"{mixname} {nframes} x {fb}: stages ms (index+literals, sequences, execute, checksum):", " ".join(f"{v:.3f}" for v in acc / max(calls - 1, 1)), "bytes equal:", ok, flush=True)
