"""One invocation of every kernel outside the two hot batches, for ncu (development tool): a 4 MiB frame with a checksum decoded by
the block-parallel path (L1-L4, large checksum), a multi-block frame compressed with a checksum (far table, XXH64), dictionary
training (T1-T5, statistics), dictionary-mode batch compress / decompress of small objects."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings
import numpy as np, torch
import zstdlite_b200 as z
from zstdlite_b200 import corpus
from oracle import ref
d = corpus.make("text", 4 << 20, 21).tobytes()
for _ in range(1):
    c = z.zstd_compress(d, level=3, include_checksum=True)
    assert z.zstd_decompress(ref.compress(d, 3, include_checksum=True)) == d
assert ref.decompress(c) == d
objs = corpus.small_objects(3004)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    dic = z.zstd_train_dict_compress(objs[:3000], 2000)
cc, dc = z.zstd_cctx(level=3, dict=dic), z.zstd_dctx(dict=dic)
for o in objs[3000:3004]:
    assert z.zstd_decompress(z.zstd_compress(o, cctx=cc), dctx=dc) == o
print("ok")
# many small frames in device memory: their descriptors are built by zl_k_build_descs
from tests.gpu_util import gpu_decompress_batch
many = corpus.small_objects(6000)
res, outs = gpu_decompress_batch([ref.compress(o, 3, dict=dic) for o in many], [len(o) for o in many], dctx=z.zstd_dctx(dict=dic))
assert outs == many
print("ok (device-built descriptors)")
