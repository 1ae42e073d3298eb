"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group.  Each rank takes its
contiguous frame range, runs a codec on it (here the reference's libzstd stands in for the CUDA kernels -- the product
has no CPU path), the ranks exchange byte counts with one all-gather, and rank 0 checks that the arenas laid end to end
at the exchanged offsets form one standard multi-frame stream that decodes to the input."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_properties():
    from zstdlite_b200 import shard
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 5, 8, 1000):
            sizes = rng.integers(1, 200000, n)
            parts = shard.partition(sizes, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert a <= b == c <= d
            if n >= 50 * world:
                loads = [int(sizes[a:b].sum()) for a, b in parts]
                assert max(loads) - min(loads) <= 2 * int(sizes.max())
    assert shard.exclusive_offsets([5, 0, 7]) == ([0, 5, 5], 12)
    offs, total = shard.frame_offsets([10, 20, 30])
    assert list(offs) == [0, 10, 30] and total == 60


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ref
        from zstdlite_b200 import corpus, shard
        # the same frame list on every rank (seeded), ragged sizes incl. empty frames
        rng = np.random.default_rng(11)
        fams = ["text", "rdf", "lowent", "rand", "rle"]
        frames = [corpus.make(fams[i % 5], int(rng.integers(0, 40000)), i).tobytes() for i in range(37)]
        a, b = shard.partition([len(f) for f in frames], world)[rank]
        # compress side: my arena, then where it goes
        arena = b"".join(ref.compress(f, 3, include_checksum=(i % 2 == 0)) for i, f in enumerate(frames[a:b], start=a))
        off, total, counts = shard.arena_offsets(len(arena))
        assert counts[rank] == len(arena) and total == sum(counts)
        # assemble on rank 0 (host-side gather; payload never enters a collective on the data path of the product)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((off, arena), gathered, dst=0)
        ok = True
        if rank == 0:
            stream = bytearray(total)
            for o, ar in gathered:
                stream[o:o + len(ar)] = ar
            want = b"".join(frames)
            ok = ref.DCtx().decompress(bytes(stream), cap=len(want), all_frames=True) == want
            # decode side: output offsets come from the headers, every rank could write its slice directly
            sizes = [len(f) for f in frames]
            offs, tot = shard.frame_offsets(sizes)
            ok = ok and tot == len(want) and all(want[int(o):int(o) + s] == f for o, s, f in zip(offs, sizes, frames))
        flag = torch.tensor([1 if ok else 0])
        dist.broadcast(flag, src=0)
        q.put((rank, bool(flag.item()), (a, b)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_shards_form_one_stream():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for _, ok, _ in res)
    (a0, b0), (a1, b1) = res[0][2], res[1][2]
    assert a0 == 0 and b0 == a1 and b1 == 37 and 0 < b0 < 37
