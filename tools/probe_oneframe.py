"""Development probe: ONE large frame compressed device-resident (the engine under ZSTD_compress2), stage times with and without far candidates.
usage: python tools/probe_oneframe.py [MiB] [family]   (ZL_ENC_NOFAR=1 in the environment switches the far tables off)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zstdlite_b200 as z
from zstdlite_b200 import corpus

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fam = sys.argv[2] if len(sys.argv) > 2 else "text"
n = mib << 20
raw = corpus.make(fam, n, 5)
dev = torch.device("cuda:0")
src = torch.from_numpy(raw).to(dev)
L = z._lib.lib()
bound = int(L.ZSTD_compressBound(n))
dst = torch.zeros(bound + 64, dtype=torch.uint8, device=dev)
for lvl in (1, 3):
    cctx = z.zstd_cctx(level=lvl)
    plan = z.BatchPlan([src.data_ptr()], [n], [dst.data_ptr()], [bound])
    for _ in range(2):
        res = plan.compress(cctx)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3):
        res = plan.compress(cctx)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    st = [L.zl_cctx_last_stage_ms(cctx._p, k) for k in range(5)]
    print(f"{fam} {mib} MiB level {lvl}: {ms:.2f} ms -> {n / ms / 1e6:.1f} GB/s, ratio {n / int(list(res)[0]):.3f}; kernels {cctx.last_kernel_ms:.2f} ms, stages "
          + ", ".join(f"{nm} {v:.2f}" for nm, v in zip(("match(+far build)", "parse", "literals", "sequences", "plan+assemble"), st)))
