"""Compress timing of one large buffer (config 1's shape) through the batch API, device-resident (development tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zstdlite_b200 as z
from zstdlite_b200 import corpus
from oracle import ref
payload = corpus.r_data_frame(1000000)
n = len(payload)
raw = torch.from_numpy(np.frombuffer(payload, dtype=np.uint8).copy()).cuda()
L = z._lib.lib()
cap = int(L.ZSTD_compressBound(n))
dst = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
c = z.zstd_cctx(level=3)
plan = z.BatchPlan([raw.data_ptr()], [n], [dst.data_ptr()], [cap])
for _ in range(3): res = plan.compress(c)
torch.cuda.synchronize(); t = time.time()
for _ in range(5): res = plan.compress(c)
torch.cuda.synchronize(); dt = (time.time() - t) / 5
st = [L.zl_cctx_last_stage_ms(c._p, k) for k in range(5)]
out = bytes(dst[:int(res[0])].cpu().numpy())
assert ref.decompress(out) == payload
print(f"16 MB buffer: {dt*1e3:.2f} ms = {n/dt/1e9:.2f} GB/s, size {len(out)}; stages " + " ".join(f"{v:.2f}" for v in st))
