"""End-to-end decode with pinned vs pageable host buffers (development probe). usage: probe_pageable.py [nframes]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zstdlite_b200 as z
sys.argv = [sys.argv[0]] + sys.argv[1:]
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
fb = 65536
data, frames = bench.make_corpus(n, fb)
sizes = [len(f) for f in frames]
offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
blob = np.frombuffer(b"".join(frames), dtype=np.uint8)
d = z.zstd_dctx()
def run(src_ptr, dst_ptr, label, check):
    plan = z.BatchPlan([src_ptr + int(o) for o in offs[:-1]], sizes, [dst_ptr + i * fb for i in range(n)], [fb] * n)
    plan.decompress(d, device=False)
    assert check(), label
    tt = []
    for _ in range(5):
        t = time.time(); plan.decompress(d, device=False); tt.append(time.time() - t)
    print(f"{label}: best {min(tt)*1e3:.1f} ms -> {n*fb/min(tt)/1e9:.1f} GB/s (median {np.median(tt)*1e3:.1f} ms)", flush=True)
hs = torch.from_numpy(blob.copy()).pin_memory(); hd = torch.zeros(n * fb, dtype=torch.uint8).pin_memory()
run(hs.data_ptr(), hd.data_ptr(), "pinned", lambda: bool((hd.numpy().reshape(n, fb) == data).all()))
ps = blob.copy(); pd = np.zeros(n * fb, dtype=np.uint8)
run(ps.ctypes.data, pd.ctypes.data, "pageable", lambda: bool((pd.reshape(n, fb) == data).all()))
# one call through the libzstd symbol on the concatenated stream (what src/raw-file.c:189 does)
import ctypes as C
L = z._lib.lib()
out = np.zeros(n * fb, dtype=np.uint8)
for _ in range(3):
    t = time.time(); r = L.ZSTD_decompressDCtx(d._p, C.c_void_p(out.ctypes.data), n * fb, C.c_void_p(ps.ctypes.data), len(blob)); dt = time.time() - t
assert int(r) == n * fb and bool((out.reshape(n, fb) == data).all())
print(f"ZSTD_decompressDCtx on the concatenated stream, pageable: {dt*1e3:.1f} ms -> {n*fb/dt/1e9:.1f} GB/s")
# one-shot ZSTD_compress2 of one buffer (the reference's zstd_compress path): pinned vs pageable input/output
from oracle import ref
m = 4096
raw = data[:m].reshape(-1).copy()
cap = int(L.ZSTD_compressBound(raw.size))
cc = z.zstd_cctx(level=3)
for label, src_t, dst_t in (("pinned", torch.from_numpy(raw.copy()).pin_memory(), torch.zeros(cap, dtype=torch.uint8).pin_memory()), ("pageable", torch.from_numpy(raw.copy()), torch.zeros(cap, dtype=torch.uint8))):
    tt = []
    for _ in range(4):
        t = time.time(); r = L.ZSTD_compress2(cc._p, C.c_void_p(dst_t.data_ptr()), cap, C.c_void_p(src_t.data_ptr()), raw.size); tt.append(time.time() - t)
    assert not z.is_error(r)
    back = ref.decompress(dst_t.numpy()[:int(r)].tobytes())
    assert back == raw.tobytes()
    print(f"ZSTD_compress2 {raw.size >> 20} MiB, {label}: best {min(tt[1:])*1e3:.1f} ms -> {raw.size/min(tt[1:])/1e9:.1f} GB/s, ratio {raw.size/int(r):.2f}")
