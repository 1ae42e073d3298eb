"""Development probe: wall clock of the package-level one-shot calls (zstd_compress / zstd_decompress on Python bytes) for one large buffer."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zstdlite_b200 as z
from zstdlite_b200 import corpus
from oracle import ref
for mib in (16, 64):
    d = corpus.make("text", mib << 20, 21).tobytes()
    c = ref.compress(d, 3)
    dctx, cctx = z.zstd_dctx(), z.zstd_cctx(level=3)
    for _ in range(3):
        t = time.time(); out = z.zstd_decompress(c, dctx=dctx); dt = time.time() - t
    assert out == d
    for _ in range(3):
        t = time.time(); cc = z.zstd_compress(d, cctx=cctx); dt2 = time.time() - t
    assert ref.decompress(cc) == d
    t = time.time(); ref.decompress(c); tr = time.time() - t
    t = time.time(); ref.compress(d, 3); tc = time.time() - t
    print(f"{mib} MiB: zstd_decompress {dt*1e3:.1f} ms (kernels {dctx.last_kernel_ms:.1f}; libzstd {tr*1e3:.0f} ms), zstd_compress {dt2*1e3:.1f} ms (kernels {cctx.last_kernel_ms:.1f}; libzstd {tc*1e3:.0f} ms)", flush=True)
