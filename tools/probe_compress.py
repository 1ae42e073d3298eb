"""Device-resident timing probe of the compression pipeline with the per-kernel split (development tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zstdlite_b200 as z
from zstdlite_b200 import corpus

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
fb = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mixname = sys.argv[4] if len(sys.argv) > 4 else "mix"
levels = [int(x) for x in (sys.argv[5].split(",") if len(sys.argv) > 5 else ["1", "3"])]
mix = {"mix": (("text", 0.4), ("rdf", 0.4), ("lowent", 0.1), ("rand", 0.1)), "text": (("text", 1.0),), "rdf": (("rdf", 1.0),),
       "lowent": (("lowent", 1.0),), "rand": (("rand", 1.0),), "rle": (("rle", 1.0),)}[mixname]
data, fams = corpus.mixed_frames(n, fb, mix=mix, pool=32)
src = torch.from_numpy(data.reshape(-1)).cuda()
L = z._lib.lib()
bound = int(L.ZSTD_compressBound(fb)); slot = (bound + 255) // 256 * 256
dst = torch.zeros(n * slot + 64, dtype=torch.uint8, device="cuda")
names = ("match", "parse", "literals", "sequences", "plan+assemble")
for lvl in levels:
    c = z.zstd_cctx(level=lvl)
    plan = z.BatchPlan([src.data_ptr() + i * fb for i in range(n)], [fb] * n, [dst.data_ptr() + i * slot for i in range(n)], [bound] * n)
    for it in range(iters):
        res = plan.compress(c)
        ms = c.last_kernel_ms
        st = [L.zl_cctx_last_stage_ms(c._p, k) for k in range(5)]
        print(f"L{lvl} iter {it}: kernels {ms:.3f} ms -> {n*fb/ms/1e6:.1f} GB/s ; " + " ".join(f"{a}={b:.2f}" for a, b in zip(names, st)), flush=True)
    sizes = np.array(list(res), dtype=np.int64)
    print(f"L{lvl} ratio {n*fb/sizes.sum():.3f}")
