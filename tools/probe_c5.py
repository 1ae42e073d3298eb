"""configs[4]-style probe per family: 128 KiB slabs made on the device, compressed by the GPU compressor (checksums on or off),
decoded device-resident; prints the decode time per family.  usage: probe_c5.py [frames] [checksum 0|1] [families...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zstdlite_b200 as z
from zstdlite_b200 import corpus

m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cks = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
fams = sys.argv[3:] or ["text", "rdf", "lowent", "rand", "rle", "mix5"]
fb = 131072
MIX5 = (("text", 0.3), ("rdf", 0.3), ("lowent", 0.15), ("rand", 0.15), ("rle", 0.1))
bound = int(z._lib.lib().ZSTD_compressBound(fb)); slot = (bound + 255) // 256 * 256
stream = torch.cuda.current_stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for fam in fams:
    mix = MIX5 if fam == "mix5" else ((fam, 1.0),)
    cctx, dctx = z.zstd_cctx(level=3, include_checksum=cks), z.zstd_dctx()
    cctx.set_stream(stream.cuda_stream); dctx.set_stream(stream.cuda_stream)
    src, pools, meta = corpus.device_mixed_slabs(m, fb, mix, index=31, device="cuda")
    comp = torch.empty(m * slot + 64, dtype=torch.uint8, device="cuda")
    back = torch.empty((m, fb), dtype=torch.uint8, device="cuda")
    cplan = z.BatchPlan([src.data_ptr() + i * fb for i in range(m)], [fb] * m, [comp.data_ptr() + i * slot for i in range(m)], [bound] * m)
    sizes = [int(r) for r in cplan.compress(cctx)]
    dplan = z.BatchPlan([comp.data_ptr() + i * slot for i in range(m)], sizes, [back.data_ptr() + i * fb for i in range(m)], [fb] * m)
    dplan.decompress(dctx); torch.cuda.synchronize()
    best = 1e9
    for it in range(3):
        e0.record(stream); res = dplan.decompress(dctx); e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ok = all(int(r) == fb for r in res) and torch.equal(back, src)
    print(f"{fam:7s} checksum={int(cks)} frames={m} ratio={m*fb/sum(sizes):.3f} decode {best:.2f} ms = {m*fb/best/1e6:.1f} GB/s ok={ok}", flush=True)
    del src, comp, back, cplan, dplan
