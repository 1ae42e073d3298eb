"""Helpers for the -m gpu tests: device buffers via torch (plumbing only), calls through the C ABI."""
import numpy as np
import torch

import zstdlite_b200 as z


def to_dev(b):
    a = np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    t = torch.empty(a.size + 64, dtype=torch.uint8, device="cuda")
    if a.size:
        t[:a.size].copy_(torch.from_numpy(a.copy()))
    return t


def pack_frames(frames):
    """list of bytes -> (device tensor, offsets, sizes); frames are packed back to back (arbitrary alignment)."""
    sizes = [len(f) for f in frames]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    blob = np.frombuffer(b"".join(frames), dtype=np.uint8)
    return to_dev(blob), offs, sizes


def gpu_decompress_batch(frames, caps, dctx=None, device=True):
    """Decode a list of frames on the GPU; returns (list of result codes, list of output bytes)."""
    dctx = dctx or z.zstd_dctx()
    if device:
        src, offs, sizes = pack_frames(frames)
        dofs = np.concatenate([[0], np.cumsum([(c + 15) // 16 * 16 for c in caps])]).astype(np.int64)
        dst = torch.zeros(int(dofs[-1]) + 64, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        res = z.decompress_batch(dctx, [src.data_ptr() + int(o) for o in offs[:-1]], sizes,
                                 [dst.data_ptr() + int(o) for o in dofs[:-1]], list(caps), device=True)
        host = dst.cpu().numpy()
        outs = [host[int(dofs[i]):int(dofs[i]) + (res[i] if not z.is_error(res[i]) else 0)].tobytes() for i in range(len(frames))]
        return res, outs
    import ctypes as C
    sbufs = [C.create_string_buffer(bytes(f), max(1, len(f))) for f in frames]
    dbufs = [C.create_string_buffer(max(1, c)) for c in caps]
    res = z.decompress_batch(dctx, [C.addressof(b) for b in sbufs], [len(f) for f in frames],
                             [C.addressof(b) for b in dbufs], list(caps), device=False)
    outs = [dbufs[i].raw[:res[i]] if not z.is_error(res[i]) else b"" for i in range(len(frames))]
    return res, outs
