#!/usr/bin/env python
"""BASELINE.json config 5 at its full size: OU euler(), 10^10 paths x 10^3 steps, sharded over the GPUs of one
node, one all-reduce of the accumulators, checked against the exact law of the Euler recursion
and against dOU with the O(dt) bias.  Secondary measurement (not the driver's bench contract).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 \
        tools/run_c5.py [--paths 1e10] > gpurun_out/c5.json
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--paths", type=float, default=1e10)
    ap.add_argument("--nsteps", type=int, default=1000)
    ap.add_argument("--seed", type=lambda v: int(v, 0), default=0x5DE70005)
    ap.add_argument("--abi-devices", type=int, default=0,
                    help="run in ONE process through the C ABI's multi-GPU context (sdegpu_create_multi) on this many devices "
                         "instead of one torch.distributed rank per GPU")
    args = ap.parse_args()
    if args.abi_devices:
        return main_abi(args)
    import torch
    import torch.distributed as dist
    import sdetools_b200 as sde
    from sdetools_b200 import distributed as sdist
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    real_stdout = os.dup(1)
    os.dup2(2, 1)                                            # library chatter goes to stderr
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = sde.default_engine(local)
    lam, xi, sig, x0, n = 1.3, 0.7, 0.9, 2.0, args.nsteps
    h = 1.0 / n
    t = np.arange(n + 1) * h
    N = int(args.paths)
    a = 1 - lam * h                                          # exact law of the recursion: Gaussian
    mean = xi + (x0 - xi) * a ** n
    var = sig * sig * h * (1 - a ** (2 * n)) / (1 - a * a)
    sd = np.sqrt(var)
    nb = 512
    hist = (mean - 6 * sd, mean + 6 * sd, nb)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    packed, mine, _ = sdist.run_sharded(eng, "ou", "euler", [lam, xi, sig], t, [x0], nx=1, npaths_total=N, seed=args.seed,
                                        center=[mean], hist=hist, rank=rank, world=world)
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=packed.device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        buf = packed.cpu().numpy()
        ps, hc = sdist.unpack_stats(buf, 1, nb)
        cnt, s1, s2 = ps[0, 0], ps[1, 0], ps[2, 0]
        m_hat, v_hat = mean + s1 / cnt, (s2 - s1 * s1 / cnt) / (cnt - 1)
        from scipy import stats
        edges = np.linspace(hist[0], hist[1], nb + 1)
        p = np.diff(stats.norm.cdf(edges, mean, sd))
        keep = p * cnt > 50
        obs, exp = hc[1:nb + 1][keep].astype(float), p[keep] * cnt
        chi2 = float(((obs - exp) ** 2 / exp).sum())
        exact_mean, exact_sd = xi + (x0 - xi) * np.exp(-lam), sig * np.sqrt((1 - np.exp(-2 * lam)) / 2 / lam)
        out = {"config": "C5 OU euler", "paths": N, "nsteps": n, "n_gpus": world, "seed": args.seed, "ms": float(ms[0]),
               "path_steps_per_s": N * n / float(ms[0]) * 1e3, "counted": int(cnt), "hist_total": int(hc.sum()),
               "mean_err_in_se": float((m_hat - mean) / (sd / np.sqrt(cnt))),
               "var_err_in_se": float((v_hat - var) / (var * np.sqrt(2.0 / cnt))),
               "chi2": chi2, "chi2_dof": int(keep.sum()) - 1, "chi2_p": float(stats.chi2.sf(chi2, int(keep.sum()) - 1)),
               "bias_vs_dOU_mean": float(m_hat - exact_mean), "bias_vs_dOU_sd": float(np.sqrt(v_hat) - exact_sd),
               "expected_bias_mean": float(mean - exact_mean), "expected_bias_sd": float(sd - exact_sd)}
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main_abi(args):
    """The same run through sdegpu_run on a multi-device context: host pointers, path ranges per device, one all-reduce."""
    import time
    from scipy import stats
    from sdetools_b200.engine import Engine
    eng = Engine(list(range(args.abi_devices)))
    lam, xi, sig, x0, n = 1.3, 0.7, 0.9, 2.0, args.nsteps
    h = 1.0 / n
    t = np.arange(n + 1) * h
    N = int(args.paths)
    a = 1 - lam * h
    mean = xi + (x0 - xi) * a ** n
    var = sig * sig * h * (1 - a ** (2 * n)) / (1 - a * a)
    sd = np.sqrt(var)
    nb = 512
    hist = (mean - 6 * sd, mean + 6 * sd, nb)
    ps, hc = np.zeros(5), np.zeros(nb + 3, dtype=np.uint64)
    eng.run("ou", "euler", [lam, xi, sig], t, [x0], nx=1, npaths=1 << 20, seed=args.seed, power_sums=ps, center=[mean], hist_counts=hc, hist=hist)
    t0 = time.perf_counter()
    ms = eng.run("ou", "euler", [lam, xi, sig], t, [x0], nx=1, npaths=N, seed=args.seed, power_sums=ps, center=[mean], hist_counts=hc,
                 hist=hist, want_time=True)
    wall = time.perf_counter() - t0
    cnt, s1, s2 = ps[0], ps[1], ps[2]
    m_hat, v_hat = mean + s1 / cnt, (s2 - s1 * s1 / cnt) / (cnt - 1)
    edges = np.linspace(hist[0], hist[1], nb + 1)
    p = np.diff(stats.norm.cdf(edges, mean, sd))
    keep = p * cnt > 50
    obs, exp = hc[1:nb + 1][keep].astype(float), p[keep] * cnt
    chi2 = float(((obs - exp) ** 2 / exp).sum())
    print(json.dumps({"config": "C5 OU euler through sdegpu_create_multi (one process, C ABI)", "paths": N, "nsteps": n, "n_gpus": args.abi_devices,
                      "seed": args.seed, "reduction": eng.last_reduction, "wall_s": wall, "slowest_device_kernel_ms": ms,
                      "path_steps_per_s_wall": N * n / wall, "counted": int(cnt), "hist_total": int(hc.sum()),
                      "mean_err_in_se": float((m_hat - mean) / (sd / np.sqrt(cnt))),
                      "var_err_in_se": float((v_hat - var) / (var * np.sqrt(2.0 / cnt))),
                      "chi2": chi2, "chi2_dof": int(keep.sum()) - 1, "chi2_p": float(stats.chi2.sf(chi2, int(keep.sum()) - 1))}))


if __name__ == "__main__":
    main()
