"""numpy reference of the counter-based increment generator the CUDA kernels
fuse into the step loop.  TEST INFRASTRUCTURE ONLY.

The reference draws increments with R's `rnorm` (R/sdetools.R:19): Mersenne-Twister
uniforms pushed through qnorm (inversion).  The device keeps the inversion but takes the
uniform from Philox4x32-10 (Salmon et al., SC'11; Random123 KATs in SURVEY.md App. F), so a
draw is a pure function of (seed, global path id, step, channel).  This file states that
generator exactly so the device stream can be checked draw by draw:

  key      = (seed & 0xffffffff, seed >> 32)
  counter  = (path & 0xffffffff, path >> 32, block, channel)      one block = two steps
  r0..r3   = philox4x32_10(counter, key)
  step 2*block uses (lo,hi) = (r0,r1); step 2*block+1 uses (r2,r3):
    m    = top 52 bits of hi:lo;  u = m * 2^-52  (u = 2^-52 if m == 0)
    z    = (-1)^(bit 11 of lo) * -qnorm(u/2)
  dB[i,k]  = sqrt(dt[i]) * z[i]   for channel k
The device evaluates -qnorm(u/2) from a table of piecewise degree-5 polynomials with
|error| < 2e-12 (tools/gen_icdf_table.py); here it is scipy's ndtri.
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


def bits_to_normal(lo, hi):
    from scipy.special import ndtri
    lo = np.asarray(lo, dtype=np.uint64)
    hi = np.asarray(hi, dtype=np.uint64)
    m = ((hi << np.uint64(32)) | lo) >> np.uint64(12)
    u = np.where(m == 0, 2.0 ** -52, m.astype(np.float64) * 2.0 ** -52)
    a = -ndtri(0.5 * u)
    return np.where((lo & np.uint64(0x800)) != 0, -a, a)


def normals(paths, nsteps, channel, seed):
    """z[(step, path)] for steps 0..nsteps-1 — float64 array (nsteps, npaths)."""
    paths = np.asarray(paths, dtype=np.uint64)
    nblk = (nsteps + 1) // 2
    blk = np.arange(nblk, dtype=np.uint64)[:, None]
    r0, r1, r2, r3 = philox4x32_10(paths[None, :] & MASK, paths[None, :] >> np.uint64(32), blk, channel,
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    z = np.empty((2 * nblk, paths.shape[0]))
    z[0::2] = bits_to_normal(r0, r1)
    z[1::2] = bits_to_normal(r2, r3)
    return z[:nsteps]
