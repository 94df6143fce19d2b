"""One permanent over several GPUs, one process per GPU.

The Gray-code term space is cut into equal segments (``plan.plan``); rank g of
G walks the contiguous segment range ``plan.segment_range(nseg, g, G)`` with
the same kernels as the single-GPU call and leaves an UNSCALED double-double
partial (re_hi, re_lo, im_hi, im_lo) in its own HBM.  The only exchange step
of the path is ONE collective on those four doubles -- NCCL over
NVLink/NVSwitch when the process group is NCCL -- after which every rank holds
the permanent.  It is an all-GATHER followed by a local error-free sum
(``pq_perm_combine``), not an all-reduce: the rank partials can be orders of
magnitude larger than their sum, and a plain floating-point reduction of the
hi parts would discard the low-order bits the kernels preserved.  The payload
is 32 bytes per rank: the collective is latency-bound, there is nothing to
overlap, and the reference has no counterpart (its parallelism is one OpenMP
loop, src/permanent.cpp:152-155).
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._math.permanent import _check_shapes, _raise, _resolve_matrix, _resolve_mult


def _device_partial(a, r, c, part, nparts, device_index):
    """Enqueue this rank's partial on the current torch stream; returns the
    (4,) float64 CUDA tensor, or a complex for a reference early-out."""
    import torch

    lib = _lib.load()
    out = torch.zeros(4, dtype=torch.float64, device="cuda:%d" % device_index)
    stream = torch.cuda.current_stream(device_index).cuda_stream
    status = ctypes.c_int(0)
    triv = np.zeros(2)
    rc = lib.pq_perm_partial_c128(
        a.ctypes.data_as(_lib.c_double_p), a.shape[0], a.shape[1],
        r.ctypes.data_as(_lib.c_int32_p), c.ctypes.data_as(_lib.c_int32_p),
        part, nparts, device_index, ctypes.c_void_p(stream),
        ctypes.c_void_p(out.data_ptr()), ctypes.byref(status),
        triv.ctypes.data_as(_lib.c_double_p))
    _raise(rc)
    if status.value == 1:
        return complex(triv[0], triv[1])
    return out


def combine(quads):
    """Error-free sum of per-rank quadruples, shape (G, 4) -> (4,)."""
    lib = _lib.load()
    q = np.ascontiguousarray(np.asarray(quads, dtype=np.float64)).reshape(-1, 4)
    out = np.zeros(4)
    _lib.check(lib.pq_perm_combine(q.ctypes.data_as(_lib.c_double_p), q.shape[0],
                                   out.ctypes.data_as(_lib.c_double_p)))
    return out


def finish(partial4, sum_rows):
    """(hi+lo) * 2^-(sum_rows-1), src/permanent.cpp:259."""
    lib = _lib.load()
    p = np.ascontiguousarray(np.asarray(partial4, dtype=np.float64))
    out = np.zeros(2)
    _lib.check(lib.pq_perm_finish(p.ctypes.data_as(_lib.c_double_p), int(sum_rows),
                                  out.ctypes.data_as(_lib.c_double_p)))
    return complex(out[0], out[1])


def permanent_allgather(matrix, rows, cols, group=None, device_index=None):
    """``permanent(matrix, rows, cols)`` computed by every rank of ``group``
    together; all ranks return the same 0-d complex128 array.

    Must be called by all ranks with identical arguments."""
    import torch
    import torch.distributed as dist

    a = np.ascontiguousarray(_resolve_matrix(matrix), dtype=np.complex128)
    r = _resolve_mult(rows, "rows")
    c = _resolve_mult(cols, "cols")
    _check_shapes(a, r, c)
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    if device_index is None:
        device_index = torch.cuda.current_device() if torch.cuda.is_available() else 0
    part = _device_partial(a, r, c, rank, world, device_index)
    if isinstance(part, complex):
        return np.array(np.complex128(part))
    if world > 1:
        gathered = torch.empty(world * 4, dtype=part.dtype, device=part.device)
        dist.all_gather_into_tensor(gathered, part, group=group)
        total = combine(gathered.cpu().numpy())
    else:
        total = part.cpu().numpy()
    return np.array(np.complex128(finish(total, int(r.sum()))))


def _laplace_device_partial(a, r, c, part, nparts):
    """This rank's share of one permanent_laplace problem (pq_perm_laplace_partial_c128):
    complex128 array of length C, or 1 on the reference's early-out."""
    lib = _lib.load()
    out = np.zeros(2 * max(a.shape[1], 1))
    out_len = ctypes.c_int(0)
    rc = lib.pq_perm_laplace_partial_c128(
        a.ctypes.data_as(_lib.c_double_p), a.shape[0], a.shape[1],
        r.ctypes.data_as(_lib.c_int32_p), c.ctypes.data_as(_lib.c_int32_p),
        part, nparts, out.ctypes.data_as(_lib.c_double_p), ctypes.byref(out_len))
    _raise(rc)
    return out[: 2 * out_len.value].view(np.complex128).copy()


def permanent_laplace_allgather(matrix, rows, cols, group=None, device_index=None):
    """``permanent_laplace(matrix, rows, cols)`` of ONE large problem computed by
    all ranks of ``group`` together: rank g walks the
    contiguous share ``[nseg*g//G, nseg*(g+1)//G)`` of the problem's Gray-code
    segments, and the only exchange step is one all-gather of C complex numbers per
    rank, summed by every rank in rank order (identical result on all ranks).

    Batches of small problems (the sampler) shard by problem instead
    (:func:`generate_samples_sharded`).  Must be called by all ranks with
    identical arguments.  On a GPU box the library's device list is set to this
    rank's device (``pq_set_devices``), as for every Laplace entry."""
    import torch
    import torch.distributed as dist

    a = np.ascontiguousarray(_resolve_matrix(matrix), dtype=np.complex128)
    r = _resolve_mult(rows, "rows")
    c = _resolve_mult(cols, "cols")
    _check_shapes(a, r, c)
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    if torch.cuda.is_available():
        # the Laplace entries run on the library's first device: make it this rank's
        if device_index is None:
            device_index = torch.cuda.current_device()
        _lib.check(_lib.load().pq_set_devices((ctypes.c_int32 * 1)(device_index), 1))
    import os
    if (world > 1 and torch.cuda.is_available() and dist.get_backend(group) == "nccl"
            and not os.environ.get("PQ_LAPLACE_ALLGATHER_HOST")):
        # device end to end: the walk leaves its k complex partial sums in this rank's
        # HBM, NCCL gathers them there, the ranks add them there in rank order, and ONE
        # download brings the result home
        lib = _lib.load()
        local = torch.empty(2 * max(a.shape[1], 1), dtype=torch.float64,
                            device="cuda:%d" % device_index)  # every column is written
        triv = np.zeros(2)
        out_len = ctypes.c_int(0)
        rc = lib.pq_perm_laplace_partial_dev_c128(
            a.ctypes.data_as(_lib.c_double_p), a.shape[0], a.shape[1],
            r.ctypes.data_as(_lib.c_int32_p), c.ctypes.data_as(_lib.c_int32_p), rank, world,
            device_index, ctypes.c_void_p(local.data_ptr()),
            triv.ctypes.data_as(_lib.c_double_p), ctypes.byref(out_len))
        _raise(rc)
        if not np.isnan(triv[0]):  # the reference's early-out [1]: identical on all ranks
            return np.array([1.0 + 0.0j])
        local = local[: 2 * out_len.value]
        gathered = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(gathered, local.contiguous(), group=group)
        parts = gathered.view(world, -1)
        total = parts[0] + parts[1]
        for g in range(2, world):  # fixed order: every rank adds the same numbers the same way
            total = total + parts[g]
        return total.cpu().numpy().view(np.complex128).copy()
    mine = _laplace_device_partial(a, r, c, rank, world)
    if world == 1:
        return mine
    local = torch.from_numpy(np.ascontiguousarray(mine.view(np.float64)))
    if dist.get_backend(group) == "nccl":
        local = local.to("cuda:%d" % device_index)
    gathered = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, local, group=group)
    parts = gathered.cpu().numpy().reshape(world, -1)
    total = np.zeros(parts.shape[1])
    for g in range(world):  # fixed order: every rank adds the same numbers the same way
        total = total + parts[g]
    return total.view(np.complex128).copy()


def generate_samples_sharded(input, shots, interferometer, seed_sequence,
                             reject_condition=None, group=None, pmf_rows=None,
                             device_index=None, as_array=False):
    """The lock-step sampler with the SHOTS sharded over the ranks of ``group``.

    Shots are independent (shot ``idx`` owns ``default_rng(seed_sequence + idx)``),
    so rank g of G simply runs shots ``[shots*g//G, shots*(g+1)//G)`` on its own
    GPU; there is no exchange step on the data path, only one all-gather of the
    finished samples at the end.  Every rank returns the full list, identical to
    the single-GPU (and to the reference's) result.

    ``as_array`` returns the samples as one (shots, d) int32 array instead of the
    reference's list of tuples (10^4 tuples of 100 Python ints take 20 ms to build --
    a tenth of an 8-GPU run).

    ``device_index`` is the CUDA device this rank's shots run on; the default is
    the rank's current torch device (``torch.cuda.current_device()``), i.e. what
    ``torch.cuda.set_device(LOCAL_RANK)`` selected -- never silently device 0.

    ``reject_condition`` is evaluated by every rank for ALL shots in the
    reference's shot-major order (it may draw from a generator the ranks seeded
    identically), and each rank keeps its own rows.  ``pmf_rows`` is handed to
    :func:`piquasso_b200.sampling.generate_samples` (CPU tests inject the oracle)."""
    import torch.distributed as dist

    from .sampling import generate_samples

    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    n = int(np.sum(np.asarray(input, dtype=int)))
    begin, end = (shots * rank) // world, (shots * (rank + 1)) // world
    rejects = None
    if reject_condition is not None:
        table = [[bool(reject_condition()) for _ in range(n)] for _ in range(shots)]
        flat = iter([x for row in table[begin:end] for x in row])
        rejects = lambda: next(flat)  # noqa: E731
    devices = None
    if pmf_rows is None:
        if device_index is None:
            import torch
            device_index = torch.cuda.current_device() if torch.cuda.is_available() else 0
        devices = [int(device_index)]
    mine = generate_samples(input, end - begin, interferometer, seed_sequence + begin,
                            reject_condition=rejects, pmf_rows=pmf_rows, devices=devices,
                            as_array=True)
    if world > 1:
        # one all-gather of the finished samples as an int32 array (rank shares padded
        # to equal length) -- not pickled tuples
        import torch
        d = len(np.asarray(input))
        per = -(-shots // world)
        local = np.zeros((per, d), dtype=np.int32)
        local[: end - begin] = mine
        t = torch.from_numpy(local)
        if dist.get_backend(group) == "nccl":
            t = t.to("cuda:%d" % (devices[0] if devices else torch.cuda.current_device()))
        gathered = torch.empty((world * per, d), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(gathered, t, group=group)
        g = gathered.cpu().numpy().reshape(world, per, d)
        mine = np.concatenate([g[r, : (shots * (r + 1)) // world - (shots * r) // world]
                               for r in range(world)], axis=0)
    return mine if as_array else [tuple(row) for row in mine.tolist()]
