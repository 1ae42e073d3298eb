"""Decode timing of a few LARGE single frames (config 1's shape: one multi-megabyte frame, 2 MB window, dependent blocks)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zstdlite_b200 as z
from oracle import ref
from zstdlite_b200 import corpus

def run(fam, mb, lvl=3, reps=3):
    d = corpus.make(fam, mb << 20, 21).tobytes()
    c = ref.compress(d, lvl)
    src = torch.from_numpy(np.frombuffer(c, dtype=np.uint8).copy()).cuda()
    dst = torch.zeros(len(d) + 64, dtype=torch.uint8, device="cuda")
    dctx = z.zstd_dctx(); dctx.set_profile(os.environ.get("PROBE_PROFILE", "1") == "1")
    plan = z.BatchPlan([src.data_ptr()], [len(c)], [dst.data_ptr()], [len(d)])
    res = plan.decompress(dctx)
    assert int(res[0]) == len(d), z.error_name(int(res[0]))
    assert bytes(dst[:len(d)].cpu().numpy()) == d
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(reps): plan.decompress(dctx)
    torch.cuda.synchronize(); dt = (time.time() - t0) / reps
    st = [z._lib.lib().zl_dctx_last_stage_ms(dctx._p, k) for k in range(4)]
    t1 = time.time(); ref.decompress(c); cpu = time.time() - t1
    print(f"{fam} {mb} MiB L{lvl}: ratio {len(d)/len(c):.2f}  GPU {dt*1e3:.1f} ms = {len(d)/dt/1e9:.2f} GB/s  stages {['%.1f' % s for s in st]}  | libzstd 1 thread {len(d)/cpu/1e9:.2f} GB/s")

if __name__ == "__main__":
    for fam, mb in (("rdf", 16), ("text", 16), ("text", 64)):
        run(fam, mb)
