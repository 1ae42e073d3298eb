"""1e5 separately allocated host objects through the host-pointer batch calls (staged copies; development tool)."""
import sys, time; sys.path.insert(0, '.')
import ctypes as C
import zstdlite_b200 as z
from zstdlite_b200 import corpus
from oracle import ref
objs = corpus.small_objects(100000)
d = ref.train_dict(objs[:10000], 5000)
L = z._lib.lib()
cctx = z.zstd_cctx(level=3, dict=d); dctx = z.zstd_dctx(dict=d)
sb = [C.create_string_buffer(o, len(o)) for o in objs]
caps = [int(L.ZSTD_compressBound(len(o))) for o in objs]
db = [C.create_string_buffer(c) for c in caps]
args = ([C.addressof(b) for b in sb], [len(o) for o in objs], [C.addressof(b) for b in db], caps)
cplan = z.BatchPlan(*args)
for it in range(3):
    t = time.time(); res = list(cplan.compress(cctx, device=False)); tc = time.time() - t
frames = [db[i].raw[:res[i]] for i in range(len(objs))]
fb = [C.create_string_buffer(f, len(f)) for f in frames]
ob = [C.create_string_buffer(len(o)) for o in objs]
dargs = ([C.addressof(b) for b in fb], [len(f) for f in frames], [C.addressof(b) for b in ob], [len(o) for o in objs])
dplan = z.BatchPlan(*dargs)
for it in range(3):
    t = time.time(); r2 = dplan.decompress(dctx, device=False); td = time.time() - t
assert all(ob[i].raw == objs[i] for i in range(0, len(objs), 97))
tot = sum(len(o) for o in objs)
print(f"scattered host objects, 1e5 x ~420 B: compress {tc*1e3:.1f} ms ({tot/tc/1e9:.2f} GB/s), decompress {td*1e3:.1f} ms ({tot/td/1e9:.2f} GB/s) (prepared argument arrays: the C call alone)")
