"""Multi-GPU sharding of independent files: units = files, as `sketch_files` treats
them (lib/src/lib.rs:34-47).  One process per GPU; no data-path collective; one gather of the
finished sketches to rank 0, returned in input order.

The host logic here is backend-agnostic (NCCL on GPUs, gloo in the CPU tests): `sketch_fn` does
the per-file work (the GPU sketcher in production).
"""
from typing import Callable, List, Sequence

import numpy as np


def lpt_assign(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of file indices to `world` ranks."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += int(sizes[i])
    return [sorted(x) for x in out]


def pack_sketch(hashes, counts, extras, final_size: int) -> np.ndarray:
    """Fixed-size record for the gather: [n, hash[final_size], count[final_size], extra[final_size]] as int64."""
    n = len(hashes)
    assert n <= final_size
    rec = np.zeros(1 + 3 * final_size, np.int64)
    rec[0] = n
    rec[1:1 + n] = np.asarray(hashes, np.uint64).view(np.int64)
    rec[1 + final_size:1 + final_size + n] = np.asarray(counts, np.int64)
    rec[1 + 2 * final_size:1 + 2 * final_size + n] = np.asarray(extras, np.int64)
    return rec


def unpack_sketch(rec: np.ndarray, final_size: int):
    n = int(rec[0])
    h = rec[1:1 + n].view(np.uint64).copy()
    c = rec[1 + final_size:1 + final_size + n].astype(np.uint32)
    x = rec[1 + 2 * final_size:1 + 2 * final_size + n].astype(np.uint32)
    return h, c, x


def sketch_files_sharded(paths: Sequence[str], sizes: Sequence[int], final_size: int,
                         sketch_fn: Callable[[str], tuple], dist, device="cpu"):
    """Every rank sketches its LPT share with `sketch_fn(path) -> (hashes, counts, extras)`; rank 0
    receives all sketches through ONE gather and returns them in input order (None elsewhere)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = lpt_assign(sizes, world)[rank]
    per_rank = max(len(x) for x in lpt_assign(sizes, world))
    buf = np.zeros((per_rank, 1 + 3 * final_size), np.int64)
    idx = np.full(per_rank, -1, np.int64)
    for slot, i in enumerate(mine):
        buf[slot] = pack_sketch(*sketch_fn(paths[i]), final_size)
        idx[slot] = i
    payload = torch.from_numpy(np.concatenate([idx[:, None], buf], axis=1)).to(device)
    gathered = [torch.empty_like(payload) for _ in range(world)] if rank == 0 else None
    dist.gather(payload, gathered, dst=0)
    if rank != 0:
        return None
    out = [None] * len(paths)
    for t in gathered:
        a = t.cpu().numpy()
        for row in a:
            if row[0] >= 0:
                out[int(row[0])] = unpack_sketch(row[1:], final_size)
    return out
