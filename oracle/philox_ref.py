"""numpy reference of the counter-based increment generator the CUDA kernels
fuse into the step loop.  TEST INFRASTRUCTURE ONLY.

The reference draws increments with R's `rnorm` (R/sdetools.R:19): Mersenne-Twister
uniforms pushed through qnorm (inversion).  The device keeps the inversion but takes the
uniform from Philox4x32-10 (Salmon et al., SC'11; Random123 KATs in SURVEY.md App. F), so a
draw is a pure function of (seed, global path id, step, channel).  This file states that
generator exactly so the device stream can be checked draw by draw:

  key      = (seed & 0xffffffff, seed >> 32)
  counter  = (path & 0xffffffff, path >> 32, block, channel)      one block = FOUR steps
  r0..r3   = philox4x32_10(counter, key);  step 4*block+s uses word w = r_s ("stream v2"):
    s = (w | 1) - 2^31 (odd, never 0), v = |s|, m = v << 21
    if v < 2^19: m = ((v-1) << 21) + (e_s >> 10) + 1 with e = philox4x32_10((path, block, channel | 0x40000000), key)
    u    = m * 2^-52;   z = sign(s) * -qnorm(u/2)
  dB[i,k]  = sqrt(dt[i]) * z[i]   for channel k
(The linear model's matrix-valued g uses one block per (path, step, channel pair) with 52-bit uniforms;
only the C oracle restates that variant.)
The device evaluates -qnorm(u/2) from a table of piecewise cubics with |error| < 8e-10
(tools/gen_icdf_table.py); here it is scipy's ndtri.
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


def normals(paths, nsteps, channel, seed):
    """z[(step, path)] for steps 0..nsteps-1 — float64 array (nsteps, npaths)."""
    from scipy.special import ndtri
    paths = np.asarray(paths, dtype=np.uint64)
    nblk = (nsteps + 3) // 4
    blk = np.arange(nblk, dtype=np.uint64)[:, None]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    plo, phi = paths[None, :] & MASK, paths[None, :] >> np.uint64(32)
    r = philox4x32_10(plo, phi, blk, channel, k0, k1)
    e = philox4x32_10(plo, phi, blk, channel | 0x40000000, k0, k1)
    z = np.empty((4 * nblk, paths.shape[0]))
    for s_ in range(4):
        w = np.broadcast_to(r[s_], (nblk, paths.shape[0]))
        si = (w | np.uint64(1)).astype(np.int64) - np.int64(1 << 31)
        v = np.abs(si).astype(np.uint64)
        m = v << np.uint64(21)
        ext = ((v - np.uint64(1)) << np.uint64(21)) + (np.broadcast_to(e[s_], m.shape) >> np.uint64(10)) + np.uint64(1)
        m = np.where(v < np.uint64(1 << 19), ext, m)
        u = m.astype(np.float64) * 2.0 ** -52
        a = -ndtri(0.5 * u)
        z[s_::4] = np.where(si < 0, -a, a)
    return z[:nsteps]
