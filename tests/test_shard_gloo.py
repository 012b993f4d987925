"""N > 1 host logic on CPU: world_size-2 gloo run of the file sharding + single gather."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_sketch(path):
    """Deterministic stand-in for the GPU sketcher (no GPU in this test)."""
    seed = int(path.split("_")[-1])
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3, 10))
    h = np.sort(rng.integers(0, 2**63, size=n, dtype=np.uint64))
    h[-1] = np.uint64(2**64 - 1 - seed)          # exercise the top bit through the int64 view
    return h, rng.integers(1, 100, size=n).astype(np.uint32), rng.integers(0, 2, size=n).astype(np.uint32)


def _worker(rank, world, port, paths, sizes, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from finch_rs_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    out = shard.sketch_files_sharded(paths, sizes, 10, _fake_sketch, dist)
    if rank == 0:
        q.put([(o[0].tolist(), o[1].tolist(), o[2].tolist()) for o in out])
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_assign():
    sys.path.insert(0, ROOT)
    from finch_rs_b200 import shard
    a = shard.lpt_assign([5, 9, 1, 7, 3, 3], 2)
    assert sorted(a[0] + a[1]) == list(range(6))
    loads = [sum([5, 9, 1, 7, 3, 3][i] for i in x) for x in a]
    assert abs(loads[0] - loads[1]) <= 2   # LPT: within the smallest job of optimal here
    assert shard.lpt_assign([4, 4, 4], 8)[3:] == [[]] * 5


def test_sharded_gather_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    paths = [f"file_{i}" for i in range(7)]
    sizes = [50, 10, 40, 30, 20, 60, 5]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, paths, sizes, q)) for r in range(2)]
    [p.start() for p in procs]
    got = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for p, g in zip(paths, got):
        h, c, x = _fake_sketch(p)
        assert g == (h.tolist(), c.tolist(), x.tolist())
