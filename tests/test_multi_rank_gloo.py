"""CPU test of the multi-GPU reduction (minrms_b200/dist.py) with world_size 2 over gloo:
per-rank best tables come from the oracle on disjoint orientation ranges; the merged
table must equal the oracle's sequential search over the whole list."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_ctypes as oc
    from helpers import random_chain, related_chain
    from minrms_b200.dist import allreduce_best, partition
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    A = random_chain(rng, 24)
    B = related_chain(rng, A, 27)
    p = oc.make_params(n_min=4, m=24, nB=27, max_rmsd=6.0)
    og = oc.gen_orientations(A, B, 1, 3, 4)
    ws = oc.Workspace(A, B, p)
    a, b = partition(og.shape[0], rank, world)
    local = ws.search(og[a:b], first_id=a)
    merged = allreduce_best(local)
    if rank == 0:
        full = oc.Workspace(A, B, p).search(og)
        ok = (np.array_equal(merged["ssd"].view(np.uint32), full["ssd"].view(np.uint32))
              and np.array_equal(merged["o_id"], full["o_id"])
              and np.array_equal(merged["T"].view(np.uint32), full["T"].view(np.uint32)))
        for r in range(full["ssd"].shape[0]):
            if full["ssd"][r] < 1e26:
                ok = ok and np.array_equal(merged["pairs"][r][:4 + r], full["pairs"][r][:4 + r])
        open(tmp, "w").write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_two_rank_merge_equals_sequential(tmp_path):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, 29531 + os.getpid() % 500, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
