"""Stage times of dictionary-mode batch compression / decompression of small objects (configs[3]; development tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zstdlite_b200 as z
from zstdlite_b200 import corpus
from oracle import ref
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
objs = corpus.small_objects(n)
d = ref.train_dict(objs[:10000], 5000)
sizes = [len(o) for o in objs]; total = sum(sizes)
offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
src = torch.from_numpy(np.frombuffer(b"".join(objs), dtype=np.uint8).copy()).cuda()
L = z._lib.lib()
caps = [int(L.ZSTD_compressBound(s)) for s in sizes]
coffs = np.concatenate([[0], np.cumsum([(c + 15) // 16 * 16 for c in caps])]).astype(np.int64)
cdst = torch.zeros(int(coffs[-1]) + 64, dtype=torch.uint8, device="cuda")
for use_dict in (True, False):
    cctx = z.zstd_cctx(level=3, dict=d if use_dict else None)
    plan = z.BatchPlan([src.data_ptr() + int(o) for o in offs[:-1]], sizes, [cdst.data_ptr() + int(o) for o in coffs[:-1]], caps)
    for it in range(3):
        res = plan.compress(cctx)
    st = [L.zl_cctx_last_stage_ms(cctx._p, k) for k in range(5)]
    csz = sum(int(r) for r in res)
    print(f"dict={use_dict}: kernels {cctx.last_kernel_ms:.3f} ms -> {total / cctx.last_kernel_ms / 1e6:.2f} GB/s, ratio {total / csz:.3f}; " +
          " ".join(f"{a}={b:.2f}" for a, b in zip(("match", "parse", "literals", "sequences", "plan+assemble"), st)), flush=True)
# ---- decompression of the dictionary-compressed batch: wall, kernels, stages
import time
cctx = z.zstd_cctx(level=3, dict=d)
res = z.BatchPlan([src.data_ptr() + int(o) for o in offs[:-1]], sizes, [cdst.data_ptr() + int(o) for o in coffs[:-1]], caps).compress(cctx)
csz = [int(r) for r in res]
ddst = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
dctx = z.zstd_dctx(dict=d)
dplan = z.BatchPlan([cdst.data_ptr() + int(o) for o in coffs[:-1]], csz, [ddst.data_ptr() + int(o) for o in offs[:-1]], sizes)
for prof in (False, True):
    dctx.set_profile(prof)
    for it in range(3):
        t = time.time(); dplan.decompress(dctx); wall = (time.time() - t) * 1e3
    st = [L.zl_dctx_last_stage_ms(dctx._p, k) for k in range(4)]
    print(f"decode profile={prof}: wall {wall:.3f} ms, kernels {dctx.last_kernel_ms:.3f} ms -> {total / wall / 1e6:.2f} GB/s; stages " + " ".join(f"{v:.3f}" for v in st), flush=True)
assert bytes(ddst[:total].cpu().numpy()) == b"".join(objs)
