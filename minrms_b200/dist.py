"""Multi-GPU reduction of per-size best tables: orientations are
partitioned across ranks (one process per GPU); after the local FirstPass the minimum of
the 64-bit keys (float_bits(ssd) << 32 | orientation id) picks, per size, the lowest score
with the lowest id winning ties -- the reference's sequential rule
(bin/minrms/search_or_dyn3d.cc:1825-1835).  The payload (transform + alignment) of each
size is the row of the rank that holds that key.

Two forms:
  reduce_best_device   the keys are reduced where they live: one NCCL all_reduce(MIN, int64)
                       on a tensor that aliases the context's device key buffer; every rank
                       then re-evaluates only the sizes it won and one all_gather moves the
                       fixed-stride rows (620 B per size for the test pair); the merged table
                       is copied to the host once.  Nothing is staged through host memory.
  allreduce_best       the same selection for tables that already sit in host memory (the
                       CPU tests with gloo; any table between FirstPass and Refine).

torch.distributed is plumbing only (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

UPPER_BOUND_BITS = 0x6AA56FA6  # float32 bits of 1e26f
KEY_NONE = (UPPER_BOUND_BITS << 32) | 0xFFFFFFFF
IDENTITY = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)


def keys_from_table(ssd, o_id):
    """keys of a host table (same encoding as mr_fetch_keys).  Valid between FirstPass and
    Refine only: rows rewritten by mr_refine carry o_id = -2 and no orientation id."""
    ssd = np.ascontiguousarray(ssd, np.float32)
    bits = ssd.view(np.uint32).astype(np.int64)
    ids = np.asarray(o_id, np.int64)
    if np.any((ids < 0) & (ssd < np.float32(1.0e26))):
        raise ValueError("keys_from_table: the table holds solutions without an orientation id "
                         "(rows refined by Refine()); reduce FirstPass tables, then refine")
    keys = (bits << 32) | (ids & 0xFFFFFFFF)
    keys[(ids < 0) | ~(ssd < np.float32(1.0e26))] = KEY_NONE
    return keys


def partition(count, rank, world):
    """contiguous [start, stop) slice of the shuffled orientation list for this rank
    (the reference API is already a range: SearchOr::Search(og, start, stop))."""
    per = (count + world - 1) // world
    start = min(count, rank * per)
    return start, min(count, start + per)


def rows_from_table(table):
    """int32 rows {key lo, key hi, 12 transform floats, N pairs as i | j << 16}: the layout
    mr_fetch_rows_device writes on the GPU"""
    R, N = table["ssd"].shape[0], table["pairs"].shape[1]
    keys = keys_from_table(table["ssd"], table["o_id"])
    rows = np.zeros((R, 2 + 12 + N), np.int32)
    rows[:, :2] = keys.view(np.int32).reshape(R, 2)
    rows[:, 2:14] = np.ascontiguousarray(table["T"], np.float32).reshape(R, 12).view(np.int32)
    rows[:, 14:] = np.ascontiguousarray(table["pairs"], np.int16).reshape(R, 2 * N).view(np.int32)
    none = keys == KEY_NONE
    rows[none, 2:] = 0
    return rows


def table_from_rows(rows, n_min=None):
    """host table of int32 rows; n_min (size of row 0) given: pair entries beyond a row's size
    are set to -1 like a table filled by mr_fetch_best"""
    rows = np.ascontiguousarray(rows, np.int32)
    R, N = rows.shape[0], rows.shape[1] - 14
    keys = np.ascontiguousarray(rows[:, :2]).view(np.int64).reshape(R)
    none = keys == KEY_NONE
    ssd = (keys >> 32).astype(np.uint32).view(np.float32).copy()
    o_id = (keys & 0xFFFFFFFF).astype(np.int64)
    T = np.ascontiguousarray(rows[:, 2:14]).view(np.float32).copy()
    pairs = np.ascontiguousarray(rows[:, 14:]).view(np.int16).reshape(R, N, 2).copy()
    # the reference's empty slots (search_or_dyn3d.cc:165-184)
    ssd[none] = np.float32(1.0e26)
    o_id[none] = -1
    T[none] = IDENTITY
    pairs[none] = -1
    if n_min is not None:
        beyond = np.arange(N)[None, :] >= (n_min + np.arange(R))[:, None]
        pairs[beyond] = -1
    return dict(ssd=ssd, T=T, pairs=pairs, o_id=o_id)


def _gather_select(rows, group=None):
    """rows: int32 tensor [R, C] of this rank (device or CPU).  One all_gather; per size the
    row of the rank holding the minimum key (keys are unique per orientation id, so the
    minimum has exactly one owner; sizes nobody solved keep KEY_NONE)."""
    world = dist.get_world_size(group)
    R, C = rows.shape
    allrows = torch.empty((world, R, C), dtype=torch.int32, device=rows.device)
    try:
        dist.all_gather_into_tensor(allrows, rows.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):      # backend without the single-tensor form
        parts = [torch.empty_like(rows) for _ in range(world)]
        dist.all_gather(parts, rows.contiguous(), group=group)
        allrows = torch.stack(parts)
    allkeys = allrows[:, :, :2].contiguous().view(torch.int64).squeeze(-1)   # [world, R]
    owner = allkeys.argmin(dim=0)
    return allrows[owner, torch.arange(R, device=rows.device)]               # [R, C]


def allreduce_best(table, device=None, group=None):
    """table: dict(ssd[R] f32, T[R,12] f32, pairs[R,N,2] i16, o_id[R] i64) of this rank, in
    host memory.  Returns the merged table (identical on every rank)."""
    dev = device if device is not None else torch.device("cpu")
    mine = torch.from_numpy(rows_from_table(table)).to(dev, non_blocking=True)
    return table_from_rows(_gather_select(mine, group).cpu().numpy())


class _DeviceArray:
    """zero-copy view of device memory owned by the C library (__cuda_array_interface__)"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def reduce_keys_device(search, device, group=None):
    """The cross-rank selection itself: one all_reduce(MIN) over the R 64-bit keys, in place in
    the device buffer of `search` (its own keys are snapshotted first, so that
    gather_payload_device knows which sizes it won)."""
    kptr, R = search.device_keys()                    # snapshot of my keys + address of the live ones
    keys = torch.as_tensor(_DeviceArray(kptr, (R,), "<i8"), device=device)   # keys < 2^63: signed order = unsigned order
    dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    torch.cuda.current_stream(device).synchronize()   # the library works on its own stream


def gather_payload_device(search, device, group=None):
    """After reduce_keys_device: every rank re-evaluates only the sizes its own orientations won
    (transform + alignment, as rows in device memory), one all_gather moves the rows and the
    merged table is copied to the host once.  Returns it (identical on every rank) and leaves
    it in `search`, ready for Refine()."""
    rptr, C = search.fetch_rows_device()
    rows = torch.as_tensor(_DeviceArray(rptr, (search.R, C), "<i4"), device=device)
    merged = table_from_rows(_gather_select(rows, group).cpu().numpy(), search.n_min)
    search.load_table(merged)
    return merged


def reduce_best_device(search, device, group=None):
    """Cross-rank reduction of the device-side table of `search` (a SearchOrCuda after
    search_resident / Search over this rank's share of the orientations): keys, then payload."""
    reduce_keys_device(search, device, group)
    return gather_payload_device(search, device, group)
