"""Quick device-resident timing probe of the decode pipeline (not the bench; used during development)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zstdlite_b200 as z
from zstdlite_b200 import corpus
from oracle import ref

nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
fb = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
mixname = sys.argv[4] if len(sys.argv) > 4 else "mix"
mix = {"mix": (("text", 0.4), ("rdf", 0.4), ("lowent", 0.1), ("rand", 0.1)), "text": (("text", 1.0),), "rdf": (("rdf", 1.0),),
       "lowent": (("lowent", 1.0),), "rand": (("rand", 1.0),), "rle": (("rle", 1.0),)}[mixname]
t0 = time.time()
data, fams = corpus.mixed_frames(nframes, fb, mix=mix, pool=32)
cache = {}
frames = []
for i in range(nframes):
    k = data[i].tobytes()
    if k not in cache: cache[k] = ref.compress(k, 3)
    frames.append(cache[k])
csz = sum(len(f) for f in frames)
print(f"corpus {nframes} x {fb}: ratio {nframes*fb/csz:.3f}, prep {time.time()-t0:.1f}s", flush=True)
order = os.environ.get("PROBE_ORDER", "grouped")          # grouped (by family) | shuffled | sorted (by compressed size)
if order != "grouped":
    perm = np.random.default_rng(1).permutation(nframes)
    if order == "sorted":
        perm = np.argsort([-len(f) for f in frames], kind="stable")
    frames = [frames[i] for i in perm]; data = data[perm]
sizes = [len(f) for f in frames]
offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
src = torch.from_numpy(np.frombuffer(b"".join(frames), dtype=np.uint8).copy()).cuda()
src = torch.cat([src, torch.zeros(64, dtype=torch.uint8, device="cuda")])
dst = torch.zeros(nframes * fb + 64, dtype=torch.uint8, device="cuda")
d = z.zstd_dctx()
plan = z.BatchPlan([src.data_ptr() + int(o) for o in offs[:-1]], sizes, [dst.data_ptr() + i * fb for i in range(nframes)], [fb] * nframes)
for it in range(iters):
    t = time.time()
    res = plan.decompress(d)
    wall = (time.time() - t) * 1e3
    ms = d.last_kernel_ms
    print(f"iter {it}: kernels {ms:.3f} ms -> {nframes*fb/ms/1e6:.1f} GB/s ; wall {wall:.2f} ms", flush=True)
bad = [i for i in range(nframes) if res[i] != fb]
print("bad results:", len(bad), bad[:5], [z.error_name(res[i]) for i in bad[:3]])
out = dst[:nframes * fb].cpu().numpy().reshape(nframes, fb)
print("bytes equal:", bool((out == data).all()))
# per-kernel times: the batch as one slice on one stream
d.set_profile(True)
L = z._lib.lib()
acc = np.zeros(4)
plan.decompress(d)
for it in range(iters):
    plan.decompress(d)
    acc += [L.zl_dctx_last_stage_ms(d._p, k) for k in range(4)]
print("stages ms (index+literals, sequences, execute, checksum):", " ".join(f"{v:.3f}" for v in acc / iters), f"sum {acc.sum() / iters:.3f}", flush=True)
