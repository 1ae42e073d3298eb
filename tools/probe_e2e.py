"""End-to-end (pinned host buffers through the C ABI) timing probe + raw PCIe copy bandwidth of the box (development tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zstdlite_b200 as z
from bench import make_corpus

n, fb = 16384, 65536
data, frames = make_corpus(n, fb)
sizes = [len(f) for f in frames]
offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
blob = np.frombuffer(b"".join(frames), dtype=np.uint8)
hsrc = torch.from_numpy(blob.copy()).pin_memory()
hdst = torch.zeros(n * fb, dtype=torch.uint8).pin_memory()
if len(sys.argv) > 1 and sys.argv[1] == "pcie":
    dsrc = torch.empty_like(hsrc, device="cuda"); ddst = torch.empty(n * fb, dtype=torch.uint8, device="cuda")
    for name, fn, nbytes in (("H2D", lambda: dsrc.copy_(hsrc, non_blocking=True), hsrc.numel()), ("D2H", lambda: hdst.copy_(ddst, non_blocking=True), hdst.numel())):
        fn(); torch.cuda.synchronize()
        t = time.time()
        for _ in range(5): fn()
        torch.cuda.synchronize()
        print(f"{name}: {nbytes * 5 / (time.time() - t) / 1e9:.1f} GB/s", flush=True)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize(); t = time.time()
    for _ in range(5):
        with torch.cuda.stream(s1): dsrc.copy_(hsrc, non_blocking=True)
        with torch.cuda.stream(s2): hdst.copy_(ddst, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.time() - t
    print(f"both directions at once: H2D {hsrc.numel() * 5 / dt / 1e9:.1f} + D2H {hdst.numel() * 5 / dt / 1e9:.1f} GB/s (wall of the pair)", flush=True)
d = z.zstd_dctx()
if os.environ.get("PROBE_NULL_STREAM"):
    d.set_stream(torch.cuda.current_stream().cuda_stream)
plan = z.BatchPlan([hsrc.data_ptr() + int(o) for o in offs[:-1]], sizes, [hdst.data_ptr() + i * fb for i in range(n)], [fb] * n)
for _ in range(2): plan.decompress(d, device=False)
assert (hdst.numpy().reshape(n, fb) == data).all()
ts = []
for _ in range(6):
    t = time.time(); plan.decompress(d, device=False); ts.append(time.time() - t)
print(f"e2e: best {min(ts) * 1e3:.2f} ms, mean {np.mean(ts) * 1e3:.2f} ms -> {n * fb / np.mean(ts) / 1e9:.1f} GB/s (best {n * fb / min(ts) / 1e9:.1f})", flush=True)
