"""Development probe: latency of ONE package-level call on a small object (what a caller compressing objects one at a time pays)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zstdlite_b200 as z
from zstdlite_b200 import corpus
from oracle import ref
for size in (1000, 10000, 100000, 1000000):
    d = corpus.make("text", size, 3).tobytes()
    cctx, dctx = z.zstd_cctx(level=3), z.zstd_dctx()
    c = z.zstd_compress(d, cctx=cctx)
    for _ in range(20): z.zstd_compress(d, cctx=cctx); z.zstd_decompress(c, dctx=dctx)
    t = time.time()
    for _ in range(200): z.zstd_compress(d, cctx=cctx)
    tc = (time.time() - t) / 200
    t = time.time()
    for _ in range(200): z.zstd_decompress(c, dctx=dctx)
    td = (time.time() - t) / 200
    t = time.time()
    for _ in range(50): ref.compress(d, 3)
    rc = (time.time() - t) / 50
    t = time.time()
    for _ in range(50): ref.decompress(c)
    rd = (time.time() - t) / 50
    print(f"{size:>8} B: zstd_compress {tc*1e6:7.0f} us (libzstd {rc*1e6:6.0f}), zstd_decompress {td*1e6:7.0f} us (libzstd {rd*1e6:6.0f})", flush=True)
