"""Multi-GPU reduction of per-size best tables: orientations are
partitioned across ranks (one process per GPU); after the local FirstPass one small
all-reduce(MIN) of the 64-bit keys (float_bits(ssd) << 32 | orientation id) picks, per
size, the lowest score with the lowest id winning ties -- the reference's sequential
rule (bin/minrms/search_or_dyn3d.cc:1825-1835).  The payload (transform + alignment) of
each size is then taken from the rank whose local key equals the reduced key.

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
    Returns the merged table (identical on every rank)."""
    R = table["ssd"].shape[0]
    dev = device if device is not None else torch.device("cpu")
    keys = torch.from_numpy(keys_from_table(table["ssd"], table["o_id"])).to(dev)
    red = keys.clone()
    dist.all_reduce(red, op=dist.ReduceOp.MIN, group=group)
    mine = (keys == red) & (red != KEY_NONE)
    # payload: rows this rank owns, zero elsewhere; the keys are unique per id so exactly
    # one rank owns each size -> a SUM all-reduce is a gather of the owners' rows.
    payload = np.concatenate([table["T"].reshape(R, -1).view(np.int32),
                              table["pairs"].reshape(R, -1).astype(np.int32)], axis=1)
    pt = torch.from_numpy(np.ascontiguousarray(payload)).to(dev)
    pt = pt * mine.to(pt.dtype).unsqueeze(1)
    dist.all_reduce(pt, op=dist.ReduceOp.SUM, group=group)
    red_np = red.cpu().numpy()
    pay = pt.cpu().numpy()
    none = red_np == KEY_NONE
    ssd = (red_np >> 32).astype(np.uint32).view(np.float32).copy()
    o_id = (red_np & 0xFFFFFFFF).astype(np.int64)
    T = pay[:, :12].astype(np.int32).view(np.float32).copy()
    pairs = pay[:, 12:].astype(np.int16).reshape(table["pairs"].shape)
    ssd[none] = np.float32(1.0e26)
    o_id[none] = -1
    T[none] = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    pairs[none] = -1
    return dict(ssd=ssd, T=T, pairs=pairs, o_id=o_id)
