"""Multi-GPU reduction of per-size best tables: orientations are
partitioned across ranks (one process per GPU); after the local FirstPass the per-size
records are exchanged once and the minimum of the 64-bit keys (float_bits(ssd) << 32 |
orientation id) picks, per size, the lowest score with the lowest id winning ties -- the
reference's sequential rule (bin/minrms/search_or_dyn3d.cc:1825-1835).  The payload
(transform + alignment) of each size is the row of the rank that holds that key.

torch.distributed is plumbing only (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

UPPER_BOUND_BITS = 0x6AA56FA6  # float32 bits of 1e26f
KEY_NONE = (UPPER_BOUND_BITS << 32) | 0xFFFFFFFF


def keys_from_table(ssd, o_id):
    """keys of a host table (same encoding as mr_fetch_keys)."""
    ssd = np.ascontiguousarray(ssd, np.float32)
    bits = ssd.view(np.uint32).astype(np.int64)
    ids = np.asarray(o_id, np.int64)
    keys = (bits << 32) | (ids & 0xFFFFFFFF)
    keys[(ids < 0) | ~(ssd < np.float32(1.0e26))] = KEY_NONE
    return keys


def partition(count, rank, world):
    """contiguous [start, stop) slice of the shuffled orientation list for this rank
    (the reference API is already a range: SearchOr::Search(og, start, stop))."""
    per = (count + world - 1) // world
    start = min(count, rank * per)
    return start, min(count, start + per)


def allreduce_best(table, device=None, group=None):
    """table: dict(ssd[R] f32, T[R,12] f32, pairs[R,N,2] i16, o_id[R] i64) of this rank.
    Returns the merged table (identical on every rank).

    One collective: every rank contributes its R records {key, transform, alignment} as int32
    rows (all_gather, 166 KB per rank for config 2); each rank then takes, per size, the row of
    the rank holding the minimum key -- the keys are unique per orientation id, so the minimum
    identifies exactly one owner."""
    R = table["ssd"].shape[0]
    dev = device if device is not None else torch.device("cpu")
    world = dist.get_world_size(group)
    keys = keys_from_table(table["ssd"], table["o_id"])
    C = 2 + 12 + table["pairs"][0].size
    rows = np.empty((R, C), np.int32)
    rows[:, :2] = keys.view(np.int32).reshape(R, 2)
    rows[:, 2:14] = table["T"].reshape(R, 12).view(np.int32)
    rows[:, 14:] = table["pairs"].reshape(R, -1)
    mine = torch.from_numpy(rows).to(dev, non_blocking=True)
    allrows = torch.empty((world, R, C), dtype=torch.int32, device=dev)
    try:
        dist.all_gather_into_tensor(allrows, mine, group=group)
    except (RuntimeError, NotImplementedError):      # backend without the single-tensor form
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        allrows = torch.stack(parts)
    allkeys = allrows[:, :, :2].contiguous().view(torch.int64).squeeze(-1)   # [world, R]
    owner = allkeys.argmin(dim=0)                                       # first rank on (impossible) ties
    sel = allrows[owner, torch.arange(R, device=allrows.device)]        # [R, C]: the owners' rows
    pay = sel.cpu().numpy()
    red_np = np.ascontiguousarray(pay[:, :2]).view(np.int64).reshape(R)
    none = red_np == KEY_NONE
    ssd = (red_np >> 32).astype(np.uint32).view(np.float32).copy()
    o_id = (red_np & 0xFFFFFFFF).astype(np.int64)
    T = np.ascontiguousarray(pay[:, 2:14]).view(np.float32).copy()
    pairs = pay[:, 14:].astype(np.int16).reshape(table["pairs"].shape)
    ssd[none] = np.float32(1.0e26)
    o_id[none] = -1
    T[none] = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    pairs[none] = -1
    return dict(ssd=ssd, T=T, pairs=pairs, o_id=o_id)
