"""Multi-GPU ensembles: one process per GPU, paths sharded by contiguous global
path-id ranges, ONE all-reduce of the packed accumulators at the end.

Paths are independent (each euler() call of vignettes/Killing.Rmd:189 touches
nothing shared), and the fused generator is keyed by the *global* path id, so
the ensemble's histogram is bit-identical for any number of ranks and the power
sums agree up to summation order.  There is no data-path collective; the only
exchange is the (5*nx + nbins+3)-double accumulator buffer.
"""
from __future__ import annotations

import numpy as np


def shard_range(npaths: int, rank: int, world: int):
    """Contiguous path-id range [lo, hi) of `rank`: ceil(N/G) per rank."""
    per = -(-npaths // world)
    lo = min(npaths, rank * per)
    return lo, min(npaths, lo + per)


def pack_stats(power_sums, hist_counts):
    """(5,nx) float64 + (nbins+3,) uint64 -> one float64 vector (counts are exact below 2^53)."""
    parts = []
    if power_sums is not None:
        parts.append(np.asarray(power_sums, dtype=np.float64).ravel(order="F"))
    if hist_counts is not None:
        parts.append(np.asarray(hist_counts).astype(np.float64))
    return np.concatenate(parts) if parts else np.zeros(0)


def unpack_stats(buf, nx, nbins, have_ps=True):
    buf = np.asarray(buf, dtype=np.float64)
    ps = buf[:5 * nx].reshape(5, nx, order="F") if have_ps else None
    off = 5 * nx if have_ps else 0
    hc = np.rint(buf[off:off + nbins + 3]).astype(np.uint64) if nbins else None
    return ps, hc


def allreduce_sum(t):
    """In-place sum over ranks of a torch tensor (nccl for CUDA tensors, gloo for CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def run_sharded(engine, model, scheme, params, times, x0, *, nx, npaths_total, projection="identity", seed=0,
                center=None, hist=None, rank=0, world=1, stats_dev=None, want_time=False):
    """Advance this rank's shard of a fused-generator ensemble and all-reduce the accumulators.

    stats_dev: optional preallocated torch.cuda float64 tensor of 5*nx + nbins+3 (+1) elements whose
    tail is viewed as the uint64 histogram; allocated when None.  Returns (stats tensor on device after
    the all-reduce [power sums | counts as float64], shard size, kernel ms or None)."""
    import torch
    lo, hi = shard_range(npaths_total, rank, world)
    nb = hist[2] if hist is not None else 0
    dev = engine.torch_device() if hasattr(engine, "torch_device") else torch.device("cuda", engine.device)
    ps = torch.zeros(5 * nx, dtype=torch.float64, device=dev)
    hc = torch.zeros(nb + 3, dtype=torch.int64, device=dev) if nb else None
    ms = engine.run(model, scheme, params, times, x0, nx=nx, npaths=hi - lo, projection=projection, seed=seed,
                    path_offset=lo, power_sums=ps, center=center, hist_counts=hc, hist=hist, want_time=want_time)
    packed = torch.cat([ps, hc.to(torch.float64)]) if nb else ps
    allreduce_sum(packed)
    return packed, hi - lo, ms
