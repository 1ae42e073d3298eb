"""Multi-GPU sharding of a frame list: frames are independent, so every rank (one process per GPU) takes a
contiguous range of the frame index balanced by bytes and runs the codec on it alone.  There is no collective on the
data path; the only exchange is one all-gather of per-rank byte counts, from which every rank derives where its arena
sits in the concatenated stream (compress) -- decode offsets are known up front from the frame headers.

torch.distributed is plumbing here (NCCL on the GPU box, gloo in the CPU tests); nothing in this module touches payload
bytes."""
import numpy as np


def partition(sizes, world):
    """Contiguous ranges [(start, end)] of len `world` over the frame index, balanced by sum(sizes).
    Every frame belongs to exactly one range; ranges may be empty when there are fewer frames than ranks."""
    sizes = np.asarray(sizes, dtype=np.int64)
    n = int(sizes.size)
    if world <= 0:
        raise ValueError("world must be positive")
    cum = np.concatenate([[0], np.cumsum(sizes)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        # choose the boundary closest to the target
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, n)] - target):
            k -= 1
        k = min(max(k, cuts[-1]), n)
        cuts.append(k)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def exclusive_offsets(counts):
    """Host-side exclusive scan of per-rank byte counts -> (offsets, total)."""
    counts = [int(c) for c in counts]
    offs, acc = [], 0
    for c in counts:
        offs.append(acc)
        acc += c
    return offs, acc


def arena_offsets(local_bytes, group=None):
    """All-gather of the ranks' arena sizes (world integers) -> (offset of this rank's arena, total bytes, all counts).
    Works without an initialised process group (single rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, int(local_bytes), [int(local_bytes)]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor([int(local_bytes)], dtype=torch.int64, device=dev)
    allc = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allc, mine, group=group)
    counts = [int(t.item()) for t in allc]
    offs, total = exclusive_offsets(counts)
    return offs[rank], total, counts


def frame_offsets(content_sizes):
    """Decode side: output offset of every frame (exclusive scan of the content sizes in the headers)."""
    c = np.asarray(content_sizes, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(c)])[:-1], int(c.sum())
