"""Development probe: zl_compress_split on one pinned host buffer (the compress end-to-end arm of bench.py), wall clock; ZL_ENC_TRACE=1 prints
the host timeline of its chunks.  usage: python tools/probe_compress_e2e.py [GiB] [level]"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zstdlite_b200 as z
from zstdlite_b200 import corpus
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
lvl = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fb = 131072
n = int(gib * (1 << 30)) // fb
data, _ = corpus.mixed_frames(n, fb, mix=(("text", 0.4), ("rdf", 0.4), ("lowent", 0.1), ("rand", 0.1)), pool=32)
hsrc = torch.from_numpy(data.reshape(-1)).pin_memory()
L = z._lib.lib()
hcap = n * (int(L.ZSTD_compressBound(fb)) + 8)
hdst = torch.zeros(hcap, dtype=torch.uint8).pin_memory()
cctx = z.zstd_cctx(level=lvl)
for it in range(3):
    t = time.time()
    r = L.zl_compress_split(cctx._p, C.c_void_p(hdst.data_ptr()), hcap, C.c_void_p(hsrc.data_ptr()), n * fb, fb, None, 0)
    dt = time.time() - t
    assert not z.is_error(r), z.error_name(r)
    print(f"level {lvl}: {n * fb / (1 << 30):.2f} GiB in {dt * 1e3:.1f} ms -> {n * fb / dt / 1e9:.2f} GB/s, ratio {n * fb / r:.3f}", flush=True)
