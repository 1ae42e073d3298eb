"""GPU dictionary training probe: trains on the config-4 corpus, compares with the reference's ZDICT on the same samples (development tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zstdlite_b200 as z
from zstdlite_b200 import corpus
from oracle import ref

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
size = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
optim = len(sys.argv) > 3 and sys.argv[3] == "optim"
objs = corpus.small_objects(n + 5000)
train = objs[:n]
test = objs[n:n + 5000]
t = time.time(); d_ref = ref.train_dict(train, size); t_ref = time.time() - t
z.zstd_train_dict_compress(train, size, optim=optim)   # warm-up (contexts and arenas of the trainer are kept for the process)
t = time.time(); d_gpu = z.zstd_train_dict_compress(train, size, optim=optim); t_gpu = time.time() - t
print(f"samples {n} ({sum(map(len, train))} B), dict {size}: ZDICT {t_ref:.2f}s ({len(d_ref)} B), GPU {t_gpu:.3f}s ({len(d_gpu)} B), ids {z.zstd_dict_id(d_ref)} {z.zstd_dict_id(d_gpu)}")
# the k the reference's fastCOVER search picked, and the GPU trainer forced to the same k: equal IDs = equal content
import ctypes as C
class FCP(C.Structure):
    _fields_ = [("k", C.c_uint), ("d", C.c_uint), ("f", C.c_uint), ("steps", C.c_uint), ("nbThreads", C.c_uint), ("splitPoint", C.c_double),
                ("accel", C.c_uint), ("shrinkDict", C.c_uint), ("shrinkDictMaxRegression", C.c_uint),
                ("compressionLevel", C.c_int), ("notificationLevel", C.c_uint), ("dictID", C.c_uint)]
par = FCP(); par.d = 8; par.steps = 4; par.compressionLevel = 3
blob = b"".join(train); sizes = (C.c_size_t * len(train))(*[len(s) for s in train]); out = C.create_string_buffer(size)
RL = ref.lib()
RL.ZDICT_optimizeTrainFromBuffer_fastCover.restype = C.c_size_t
RL.ZDICT_optimizeTrainFromBuffer_fastCover.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]
r = RL.ZDICT_optimizeTrainFromBuffer_fastCover(out, size, blob, sizes, len(train), C.byref(par))
print("reference picked k =", par.k, "d =", par.d, "f =", par.f, "accel =", par.accel, "split =", par.splitPoint, "id", z.zstd_dict_id(out.raw[:r]))
from zstdlite_b200 import _lib
cp = _lib.CoverParams(); cp.k = par.k; cp.d = 8; cp.splitPoint = 0.75
o2 = C.create_string_buffer(size)
r2 = _lib.lib().ZDICT_optimizeTrainFromBuffer_cover(o2, size, blob, sizes, len(train), C.byref(cp))
d_same_k = o2.raw[:r2]
print("GPU trainer at that k: id", z.zstd_dict_id(d_same_k))
tot = sum(map(len, test))
for name, d in (("ZDICT", d_ref), ("GPU", d_gpu), ("GPU@k", d_same_k)):
    rc, rd = ref.CCtx(level=3, dict=d), ref.DCtx(dict=d)
    cs = [rc.compress(o) for o in test]
    assert all(rd.decompress(c, cap=len(o)) == o for c, o in zip(cs[:200], test[:200]))
    print(f"  {name:5s} dictionary, libzstd level 3 on {len(test)} held-out objects: ratio {tot / sum(map(len, cs)):.3f}")
# the four combinations of trainer and compressor (GPU compressor through the batch API, device buffers)
import numpy as np, torch
def gpu_total(d):
    cc = z.zstd_cctx(level=3, dict=d)
    src = torch.from_numpy(np.frombuffer(b"".join(test), dtype=np.uint8).copy()).cuda()
    offs = np.concatenate([[0], np.cumsum([len(o) for o in test])]).astype(np.int64)
    caps = [len(o) + 64 for o in test]; coffs = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
    dst = torch.zeros(int(coffs[-1]) + 64, dtype=torch.uint8, device="cuda")
    res = z.compress_batch(cc, [src.data_ptr() + int(o) for o in offs[:-1]], [len(o) for o in test], [dst.data_ptr() + int(o) for o in coffs[:-1]], caps)
    return sum(int(r) for r in res)
for name, d in (("ZDICT", d_ref), ("GPU", d_gpu), ("GPU@k", d_same_k)):
    print(f"  {name:5s} dictionary, GPU compressor level 3: ratio {tot / gpu_total(d):.3f}")
cs0 = [ref.compress(o, 3) for o in test]
print(f"  no dictionary: ratio {tot / sum(map(len, cs0)):.3f}")
