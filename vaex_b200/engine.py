"""Row-sharded multi-GPU driver: one process per GPU, each bins its row range into its own full-size grids, then ONE
NCCL all-reduce per grid over NVLink (sum for count/sum/moment grids, min/max for min/max grids).

This is the B200 replacement for the reference's only parallel strategy on this path — row-range data parallelism over
threads with private grids merged at the end (vaex/execution.py:432-455, src/agg_base.hpp:33-48, src/agg_count.cpp:24-41).
torch is used for plumbing only: `torch.distributed` (NCCL) and zero-copy tensor views of the device grids.
"""
import numpy as np

from . import _lib

_TORCH_DTYPE = {"float64": "float64", "float32": "float32", "int64": "int64", "int32": "int32", "uint64": "int64", "uint32": "int32"}


class _GridView:
    """Exposes an aggregator's device grid through __cuda_array_interface__ (no copy, no ownership)."""

    def __init__(self, agg, which=0):
        ptr, nbytes = agg.device_pointer(which)
        dt = agg.device_dtype if which == 0 else np.dtype("uint64")
        # torch has no uint64/uint32 reductions: view them as the signed type of equal width (add is two's complement;
        # min/max on unsigned grids are handled separately in all_reduce_agg)
        name = _TORCH_DTYPE[dt.name]
        self._agg = agg
        self.signed_view = name != dt.name
        self.__cuda_array_interface__ = {
            "shape": (nbytes // dt.itemsize,),
            "typestr": np.dtype(name).str,
            "data": (ptr, False),
            "version": 3,
            "strides": None,
        }


def grid_tensor(agg, which=0):
    """torch tensor aliasing the device grid of `agg` (flat, dim 0 of the N-d grid fastest)."""
    import torch
    view = _GridView(agg, which)
    t = torch.as_tensor(view, device=f"cuda:{agg._ctx.device}")
    t._b200_keepalive = view
    return t, view.signed_view


def slot_stream(ctx, slot=0):
    """The slot's cudaStream_t as a torch ExternalStream, so NCCL and timing events order after our kernels."""
    import torch
    return torch.cuda.ExternalStream(ctx.stream(slot), device=f"cuda:{ctx.device}")


def all_reduce_tensor(t, op, unsigned_as_signed=False, group=None):
    """All-reduce one flat grid tensor in place.  `op` is the aggregator op (AGG_*): sum-like grids add, min/max grids take
    the extremum.  Backend agnostic (NCCL on the GPUs; gloo in the CPU tests of the sharding logic)."""
    import torch
    import torch.distributed as dist
    if op in (_lib.AGG_MIN, _lib.AGG_MAX):
        rop = dist.ReduceOp.MAX if op == _lib.AGG_MAX else dist.ReduceOp.MIN
        if unsigned_as_signed:  # unsigned min/max through a signed view: flip the sign bit so the order is preserved
            bias = torch.iinfo(t.dtype).min
            t.add_(bias)
            dist.all_reduce(t, op=rop, group=group)
            t.sub_(bias)
        else:
            dist.all_reduce(t, op=rop, group=group)
    elif op in (_lib.AGG_FIRST, _lib.AGG_LAST):
        raise NotImplementedError("first/last grids are reduced with all_reduce_first (key/row state, not the value grid)")
    elif op == _lib.AGG_NUNIQUE:
        raise NotImplementedError("nunique grids cannot be merged (the reference cannot either: src/agg_nunique.cpp:43-46)")
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


_SIGN = -(1 << 63)


def all_reduce_first_tensors(key, row, value, order, masked, group=None):
    """first/last across row shards: every rank holds, per cell, the packed winner of ITS rows — `key` (the order
    key as u64 bits, smaller wins; LAST stores the complement), `row` (global row index, the tie-break), the winner's `value`
    and `order` bits (int64 tensors) and `masked` (1 = no row of this rank fell into the cell).  THREE all-reduces: MIN over the
    keys, MIN over the rows of the ranks that hold that key — together the global (key, row) winner, which exactly one rank owns
    (rows are globally unique) — and one SUM over the stacked [value bits, order bits, has-a-row flag] to which only the owner
    contributes its bits.  All tensors are updated in place; works on any backend (gloo in the CPU test)."""
    import torch
    import torch.distributed as dist
    imax = torch.iinfo(torch.int64).max
    has = masked == 0
    k = torch.where(has, key ^ _SIGN, torch.full_like(key, imax))  # u64 order through a signed view: flip the sign bit
    gk = k.clone()
    dist.all_reduce(gk, op=dist.ReduceOp.MIN, group=group)
    r = torch.where(has & (k == gk), row, torch.full_like(row, imax))
    gr = r.clone()
    dist.all_reduce(gr, op=dist.ReduceOp.MIN, group=group)
    owner = has & (k == gk) & (r == gr)
    packed = torch.stack([torch.where(owner, value, torch.zeros_like(value)), torch.where(owner, order, torch.zeros_like(order)), has.to(torch.int64)])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    any_has = packed[2] > 0
    key.copy_(torch.where(any_has, gk ^ _SIGN, key))
    row.copy_(torch.where(any_has, gr, row))
    value.copy_(torch.where(any_has, packed[0], value))
    order.copy_(torch.where(any_has, packed[1], order))
    masked.copy_((~any_has).to(masked.dtype))


def _raw_tensor(agg, which, np_dtype):
    """flat torch view of one of the aggregator's device buffers, as the SIGNED integer type of the same width"""
    import torch
    ptr, nbytes = agg.device_pointer(which)
    isz = np.dtype(np_dtype).itemsize

    class V:
        pass
    v = V()
    v.__cuda_array_interface__ = {"shape": (nbytes // isz,), "typestr": "<i%d" % isz if isz > 1 else "|i1", "data": (ptr, False), "version": 3, "strides": None}
    v._keep = agg
    t = torch.as_tensor(v, device=f"cuda:{agg._ctx.device}")
    t._b200_keepalive = v
    return t


def all_reduce_first(agg, slot=0, group=None):
    """In-place cross-rank reduction of a first/last aggregator (value grid, order grid, {key,row} state, cell_masked)."""
    import torch
    with torch.cuda.stream(slot_stream(agg._ctx, slot)):
        state = _raw_tensor(agg, 1, "int64").view(-1, 2)
        vraw = _raw_tensor(agg, 0, agg.device_dtype)
        oraw = _raw_tensor(agg, 2, agg._dtype2)
        masked = _raw_tensor(agg, 3, "int8")
        key, row = state[:, 0].contiguous(), state[:, 1].contiguous()
        value, order = vraw.to(torch.int64), oraw.to(torch.int64)
        all_reduce_first_tensors(key, row, value, order, masked, group)
        state[:, 0].copy_(key)
        state[:, 1].copy_(row)
        vraw.copy_(value.to(vraw.dtype))
        oraw.copy_(order.to(oraw.dtype))


def all_reduce_agg(agg, slot=0, group=None):
    """In-place all-reduce of one aggregator's device grid across the ranks of `group` (ordered after the slot's kernels)."""
    import torch
    if agg._op in (_lib.AGG_FIRST, _lib.AGG_LAST):  # superagg.AggFirst_* carry AGG_FIRST for both first and last
        return all_reduce_first(agg, slot, group)
    t, signed_view = grid_tensor(agg)
    with torch.cuda.stream(slot_stream(agg._ctx, slot)):
        all_reduce_tensor(t, agg._op, signed_view, group)


def all_reduce(aggs, slot=0, group=None):
    for a in aggs:
        all_reduce_agg(a, slot, group)


def shard_range(nrows, rank, world):
    """Contiguous row range of `rank` (rows/world each, remainder spread over the first ranks)."""
    base, rem = divmod(int(nrows), int(world))
    i1 = rank * base + min(rank, rem)
    return i1, i1 + base + (1 if rank < rem else 0)


def union_key_sets(local_set, group=None, make_set=None):
    """Groupby pass 1 across GPUs: every rank built an ordered set over its own row range; all-gather the
    unique keys (<= 8 MB for 1e6 int64 keys), and let every rank rebuild the SAME set by inserting the per-rank key arrays in
    rank order — so all ranks derive identical ordinals (first-seen order of the row-sharded frame) without further exchange.

    `local_set` is anything with the ordered_set protocol (`key_array()`, `nan_count`, `null_count`, `null_index`, `nan_index`
    and a constructor `type(local_set)(nmaps)` + `update(keys[, masks], start_index)`): the device set here, the oracle's
    restatement in the CPU test.  Returns the union set (a new object of the same type)."""
    import numpy as np
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    keys = np.asarray(local_set.key_array())
    special = (int(local_set.nan_count > 0), int(local_set.null_count > 0), int(local_set.nan_index), int(local_set.null_index))
    payload = (keys, special)
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, payload, group=group)
    else:
        gathered[0] = payload
    union = make_set() if make_set is not None else type(local_set)(getattr(local_set, "nmaps", 1))
    for k, (has_nan, has_null, nan_i, null_i) in gathered:
        if len(k) == 0:
            continue
        mask = None
        if has_null:
            mask = np.zeros(len(k), bool)
            mask[null_i] = True
        # NaN slots carry NaN in key_array(), so they re-enter as NaN keys; null slots need the mask
        union.update(k, masks=mask, start_index=-1)
    return union
