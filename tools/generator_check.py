#!/usr/bin/env python
"""The fused normal generator on its own, at the ensemble size of config C5.

OU with lambda = 0, sigma = 1 from x0 = 0 is X_T = sum_i sqrt(dt_i) z_i.  On a grid of four intervals with ONE interval of length 1
and three of length 1e-24 the terminal value is a single draw (the others move it by 1e-12): interval k of a block of four steps
uses word k of the path's Philox block, so four grids isolate the four words.  For every word: N/4 paths through `k_terminal`,
power sums about 0 and a 512-bin histogram on [-6, 6] in the kernel, compared with the exact N(0,1) law: mean, variance, fourth
moment (standard errors 1/sqrt(n), sqrt(2/n), sqrt(96/n)), chi-square over the bins with expectation > 50, and the two tails
beyond |z| = 6 against their Poisson expectation.

    python tools/generator_check.py [--draws 1e10] [--seed 0x5DE7000A] > profiles/r2_generator_check.json
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def check(eng, draws, seed, nb=512):
    from scipy import stats
    from sdetools_b200 import distributed as sdist
    eps = 1e-24
    per_word = []
    total_ps, total_hc, total_n = np.zeros(5), np.zeros(nb + 3), 0
    n_each = int(draws) // 4
    edges = np.linspace(-6.0, 6.0, nb + 1)
    p = np.diff(stats.norm.cdf(edges))
    ptail = stats.norm.sf(6.0)

    def verdict(ps, hc, n):
        cnt, s1, s2, s4 = ps[0], ps[1], ps[2], ps[4]
        keep = p * n > 50
        obs, exp = hc[1:nb + 1][keep].astype(float), p[keep] * n
        chi2 = float(((obs - exp) ** 2 / exp).sum())
        lo, hi = float(hc[0]), float(hc[nb + 1])
        return {"draws": int(cnt), "mean_err_in_se": float(s1 / cnt * np.sqrt(cnt)), "var_err_in_se": float((s2 / cnt - 1.0) / np.sqrt(2.0 / cnt)),
                "m4_err_in_se": float((s4 / cnt - 3.0) / np.sqrt(96.0 / cnt)), "chi2": chi2, "chi2_dof": int(keep.sum()) - 1,
                "chi2_p": float(stats.chi2.sf(chi2, int(keep.sum()) - 1)), "below_minus6": lo, "above_6": hi,
                "tail_expected_each": float(ptail * n), "tail_poisson_p": float(stats.poisson.sf(lo + hi - 1, 2 * ptail * n) if lo + hi > 2 * ptail * n
                                                                                else stats.poisson.cdf(lo + hi, 2 * ptail * n)),
                "nan": float(hc[nb + 2])}

    for k in range(4):
        dt = np.full(4, eps)
        dt[k] = 1.0
        t = np.concatenate(([0.0], np.cumsum(dt)))
        packed, _, _ = sdist.run_sharded(eng, "ou", "euler", [0.0, 0.0, 1.0], t, [0.0], nx=1, npaths_total=n_each, seed=seed,
                                         center=[0.0], hist=(-6.0, 6.0, nb), rank=0, world=1)
        ps, hc = sdist.unpack_stats(packed.cpu().numpy(), 1, nb)
        ps, hc = ps[:, 0].astype(float), hc.astype(float)
        per_word.append({"word": k, **verdict(ps, hc, n_each)})
        total_ps += ps
        total_hc += hc
        total_n += n_each
    return {"config": "fused generator alone: OU lambda = 0, one dominant interval per grid, k_terminal, 512 bins on [-6, 6]",
            "seed": seed, "all_words": verdict(total_ps, total_hc, total_n), "per_word": per_word}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--draws", type=float, default=1e10)
    ap.add_argument("--seed", type=lambda v: int(v, 0), default=0x5DE7000A)
    a = ap.parse_args()
    import sdetools_b200 as sde
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = check(sde.default_engine(0), a.draws, a.seed)
    os.write(real_stdout, (json.dumps(out, indent=1) + "\n").encode())
