"""numpy reference of the counter-based increment generator the CUDA kernels
fuse into the step loop.  TEST INFRASTRUCTURE ONLY.

The reference draws increments with R's `rnorm` (R/sdetools.R:19, Mersenne-
Twister + inversion), a stream that cannot exist on the device; the product
replaces it with Philox4x32-10 (Salmon et al., SC'11; Random123 KATs in
SURVEY.md App. F).  This file states that generator exactly so the device
stream can be checked draw by draw:

  key      = (seed & 0xffffffff, seed >> 32)
  counter  = (path & 0xffffffff, path >> 32, block, channel)
  r0..r3   = philox4x32_10(counter, key)
  m1, m2   = top 52 bits of r1:r0 and of r3:r2
  u1       = 2 - (1 + m1 * 2^-52)                       in (0,1]
  v        =      1 + m2 * 2^-52                        in [1,2)   (angle fraction + 1)
  rad      = sqrt(-2 log u1)
  z[2*block]   = rad * cospi(2 v)         (Box-Muller pair, one block = two steps)
  z[2*block+1] = rad * sinpi(2 v)
  dB[i,k]  = sqrt(dt[i]) * z[i]   for channel k
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable arrays of uint32 values (held in uint64)."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniforms(path, block, channel, seed):
    path = np.asarray(path, dtype=np.uint64)
    r0, r1, r2, r3 = philox4x32_10(path & MASK, path >> np.uint64(32), block, channel,
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    m1 = ((r1 << np.uint64(32)) | r0) >> np.uint64(12)
    m2 = ((r3 << np.uint64(32)) | r2) >> np.uint64(12)
    u1 = 2.0 - (1.0 + m1.astype(np.float64) * 2.0 ** -52)
    u2 = m2.astype(np.float64) * 2.0 ** -52          # v - 1: cospi/sinpi have period 2
    return u1, u2


def normals(paths, nsteps, channel, seed):
    """z[(step, path)] for steps 0..nsteps-1 — float64 array (nsteps, npaths)."""
    paths = np.asarray(paths, dtype=np.uint64)
    nblk = (nsteps + 1) // 2
    blk = np.arange(nblk, dtype=np.uint64)[:, None]
    u1, u2 = uniforms(paths[None, :], blk, channel, seed)
    rad = np.sqrt(-2.0 * np.log(u1))
    # cospi/sinpi evaluated with exact quadrant reduction like the device does
    z0 = rad * _cospi(2.0 * u2)
    z1 = rad * _sinpi(2.0 * u2)
    z = np.empty((2 * nblk, paths.shape[0]))
    z[0::2] = z0
    z[1::2] = z1
    return z[:nsteps]


def _sinpi(x):
    x = np.mod(x, 2.0)
    return np.where(x <= 0.25, np.sin(np.pi * x),
           np.where(x <= 0.75, np.cos(np.pi * (x - 0.5)),
           np.where(x <= 1.25, -np.sin(np.pi * (x - 1.0)),
           np.where(x <= 1.75, -np.cos(np.pi * (x - 1.5)), np.sin(np.pi * (x - 2.0))))))


def _cospi(x):
    return _sinpi(x + 0.5)
