"""One process, several GPUs: end-to-end decode (pinned and pageable host buffers) and compress of host batches with the library's own
fan-out (development probe).  usage: probe_multigpu.py [frames_per_gpu]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zstdlite_b200 as z
import bench
ng = torch.cuda.device_count()
per = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
fb = 65536
data, frames = bench.make_corpus(per, fb)
sizes1 = [len(f) for f in frames]
blob1 = np.frombuffer(b"".join(frames), dtype=np.uint8)
G = 1
while G <= ng:
    n = per * G
    blob = np.tile(blob1, G); sizes = sizes1 * G
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    for label, pin in (("pinned", True), ("pageable", False)):
        if pin:
            hs = torch.from_numpy(blob.copy()).pin_memory(); hd = torch.zeros(n * fb, dtype=torch.uint8).pin_memory(); sp, dp = hs.data_ptr(), hd.data_ptr()
        else:
            hs = blob.copy(); hd = np.zeros(n * fb, dtype=np.uint8); sp, dp = hs.ctypes.data, hd.ctypes.data
        d = z.zstd_dctx(num_gpus=G)
        plan = z.BatchPlan([sp + int(o) for o in offs[:-1]], sizes, [dp + i * fb for i in range(n)], [fb] * n)
        plan.decompress(d, device=False); plan.decompress(d, device=False)
        out = hd.numpy() if pin else hd
        assert (out.reshape(G, per, fb) == data[None]).all()
        tt = []
        for _ in range(5):
            t = time.time(); plan.decompress(d, device=False); tt.append(time.time() - t)
        print(f"decode e2e, {G} GPU(s), {label}: best {min(tt)*1e3:.1f} ms -> {n*fb/min(tt)/1e9:.1f} GB/s (median {np.median(tt)*1e3:.1f} ms)", flush=True)
        del plan, d, hs, hd
    G *= 2
