#!/usr/bin/env python
"""Secondary measurements (not the driver's bench contract): the other BASELINE.json configs
on one GPU, device-resident buffers, CUDA-event timing inside libsdegpu (kernels only).

    python tools/bench_modes.py [--quick] > gpurun_out/modes.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--c3-paths", type=int, default=0, help="override the C3 path count (tuning runs)")
    ap.add_argument("--c3b-paths", type=int, default=0, help="override the supplied-B path count (tuning runs)")
    args = ap.parse_args()
    import torch
    import sdetools_b200 as sde
    eng = sde.default_engine(0)
    dev = torch.device("cuda", 0)
    hbm = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
    fp64 = eng.measure_fp64_peak()
    out = []

    def best_ms(fn, reps=3):
        fn()
        return min(fn() for _ in range(reps))

    def emit(name, **kw):
        kw["name"] = name
        print(json.dumps(kw), flush=True)

    want = lambda n: not args.only or n in args.only.split(",")

    # C1: the reference's own CPU-sized case, end to end with HOST buffers (copies inside the timed region)
    if want("c1"):
        import time
        P, nt = 1000, 1001
        t = np.arange(nt) * 0.01
        rng = np.random.default_rng(1)
        Bh = np.asfortranarray(np.cumsum(np.concatenate([np.zeros((1, 1, P)), rng.standard_normal((nt - 1, 1, P)) * 0.1]), axis=0))
        Xh = np.empty((nt, 1, P), order="F")
        eng.run("ou", "euler", [1.0, 0.0, 1.0], t, [0.0], nx=1, npaths=P, B=Bh, X=Xh)
        best = 1e30
        for _ in range(5):
            t0 = time.perf_counter()
            eng.run("ou", "euler", [1.0, 0.0, 1.0], t, [0.0], nx=1, npaths=P, B=Bh, X=Xh)
            best = min(best, time.perf_counter() - t0)
        emit("c1_ou_euler_suppliedB_host_buffers", paths=P, nt=nt, wall_ms=best * 1e3, path_steps_per_s=P * (nt - 1) / best,
             h2d_bytes=Bh.nbytes, d2h_bytes=Xh.nbytes)

    # C3: van der Pol, full trajectory, fused generator (HBM-write bound; 16 B per path-step)
    if want("c3"):
        # two full waves of the trajectory kernel (one 512-thread CTA per SM): equal-length paths cannot fill a
        # partial wave, so sizes that are not a multiple of 148*512 lose the remainder (2^17 paths: 86 %)
        P = (1 << 15) if args.quick else (args.c3_paths or 2 * eng.info["sm_count"] * 512)
        nt = 10001
        t = np.arange(nt) * 0.01
        X = torch.empty(P * 2 * nt, dtype=torch.float64, device=dev)
        ms = best_ms(lambda: eng.run("vdp", "euler", [1.0, 0.5], t, [0.0, 0.0], nx=2, npaths=P, seed=3, X=X, want_time=True))
        steps = P * (nt - 1)
        emit("c3_vdp_traj_philox", paths=P, nt=nt, ms=ms, path_steps_per_s=steps / ms * 1e3, gbs=16.0 * steps / ms / 1e6,
             hbm_frac=16.0 * steps / ms / 1e6 / hbm, bytes_per_step=16)
        chk = X.view(P, 2, nt)[:4, :, :3].cpu().numpy().tolist()
        emit("c3_check", first=chk[0])
        # scalar models through the same kernel (8 B per path-step; 768 threads per SM there)
        if not args.quick:
            P1 = 2 * eng.info["sm_count"] * 768
            X1 = X[: P1 * nt]
            t1 = np.arange(nt) * 1e-3
            for name, model, params, scheme, proj in (("traj_ou_euler", "ou", [1.3, 0.7, 0.9], "euler", "identity"),
                                                      ("traj_cir_heun", "cir", [1.0, 1.0, 1.0], "heun", "abs")):
                ms = best_ms(lambda: eng.run(model, scheme, params, t1, [1.0], nx=1, npaths=P1, projection=proj, seed=3, X=X1,
                                             want_time=True))
                emit(name, paths=P1, nt=nt, ms=ms, path_steps_per_s=P1 * (nt - 1) / ms * 1e3, gbs=8.0 * P1 * (nt - 1) / ms / 1e6,
                     hbm_frac=8.0 * P1 * (nt - 1) / ms / 1e6 / hbm, bytes_per_step=8)
        # supplied-B variant: 8 B read + 16 B written per path-step
        P2 = (P // 2) if args.quick else (args.c3b_paths or eng.info["sm_count"] * 512)
        B = torch.empty(P2 * nt, dtype=torch.float64, device=dev)
        eng.rvbm(t, P2, B, seed=9)
        X2 = X[: P2 * 2 * nt]
        for arith in ("strict", "fast"):
            ms = best_ms(lambda: eng.run("vdp", "euler", [1.0, 0.5], t, [0.0, 0.0], nx=2, npaths=P2, B=B, arith=arith, X=X2,
                                         want_time=True))
            steps = P2 * (nt - 1)
            emit(f"c3_vdp_traj_suppliedB_{arith}", paths=P2, nt=nt, ms=ms, path_steps_per_s=steps / ms * 1e3,
                 gbs=24.0 * steps / ms / 1e6, hbm_frac=24.0 * steps / ms / 1e6 / hbm, bytes_per_step=24)
        del X, X2, B
        torch.cuda.empty_cache()

    # C5 per-GPU share: OU euler, terminal moments + 512-bin histogram (FP64/issue bound; 71 flops per path-step)
    if want("c5"):
        P = (1 << 22) if args.quick else 50_000_000
        n = 1000
        t = np.arange(n + 1) * 1e-3
        ps = torch.zeros(5, dtype=torch.float64, device=dev)
        hc = torch.zeros(515, dtype=torch.int64, device=dev)
        ms = best_ms(lambda: eng.run("ou", "euler", [1.3, 0.7, 0.9], t, [2.0], nx=1, npaths=P, seed=0x5DE70005, power_sums=ps,
                                     center=[1.05], hist_counts=hc, hist=(-2.2, 4.3, 512), want_time=True))
        steps = P * n
        rate = steps / ms * 1e3
        # executed-instruction fractions from the SASS count of this kernel's step loop (tools/sass_count.py --loop, see
        # profiles/r2_sass_k_terminal_ou_euler.json); the survey's 71-flop convention reads > 1 and is kept only under that name
        try:
            sass = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                                               "r2_sass_k_terminal_ou_euler.json")))
        except Exception:
            sass = {}
        nins, n64 = sass.get("instructions_per_path_step"), sass.get("fp64_per_path_step")
        issue_peak = eng.info["sm_count"] * 4 * eng.info["clock_khz"] * 1e3
        emit("c5_ou_euler_terminal", paths=P, nsteps=n, ms=ms, path_steps_per_s=rate, fp64_peak=fp64,
             instructions_per_path_step=nins, fp64_instructions_per_path_step=n64,
             frac_issue=nins * rate / 32.0 / issue_peak if nins else None, frac_fp64_pipe=2.0 * n64 * rate / 1e12 / fp64 if n64 else None,
             frac_convention=71.0 * rate / 1e12 / fp64)

    # integrals: stochint l/r/c on stored paths (16 B read, 24 B written per element)
    if want("integrals"):
        n, nc = 1001, (1 << 16 if args.quick else 1 << 19)
        f = torch.randn(nc * n, dtype=torch.float64, device=dev)
        g = torch.randn(nc * n, dtype=torch.float64, device=dev).cumsum(0) * 0.01
        l, r, c = (torch.empty_like(f) for _ in range(3))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.use_torch_stream()

        def timed(fn):
            fn()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                ev0.record()
                fn()
                ev1.record()
                torch.cuda.synchronize()
                best = min(best, ev0.elapsed_time(ev1))
            return best
        ms = timed(lambda: eng.stochint(f, g, n, nc, l, r, c))
        emit("stochint_lrc", n=n, ncols=nc, ms=ms, elems_per_s=n * nc / ms * 1e3, gbs=40.0 * n * nc / ms / 1e6,
             hbm_frac=40.0 * n * nc / ms / 1e6 / hbm)
        ms = timed(lambda: eng.itointegral(f, g, n, nc, l))
        emit("itointegral", n=n, ncols=nc, ms=ms, elems_per_s=n * nc / ms * 1e3, gbs=24.0 * n * nc / ms / 1e6,
             hbm_frac=24.0 * n * nc / ms / 1e6 / hbm)
        ms = timed(lambda: eng.qv(g, n, nc, l))
        emit("quadratic_variation", n=n, ncols=nc, ms=ms, elems_per_s=n * nc / ms * 1e3, gbs=16.0 * n * nc / ms / 1e6,
             hbm_frac=16.0 * n * nc / ms / 1e6 / hbm)
        eng.use_own_stream()
        del f, g, l, r, c
        torch.cuda.empty_cache()

    # integrals accumulated in the step loop (k_functionals): four 80-bit accumulators per component, nothing stored but 4 doubles
    if want("functionals"):
        P, n = ((1 << 18) if args.quick else 1 << 22), 1000
        t = np.arange(n + 1) * 1e-3
        PI = torch.empty(P * 4, dtype=torch.float64, device=dev)
        XT = torch.empty(P, dtype=torch.float64, device=dev)
        ms = best_ms(lambda: eng.run("ou", "euler", [1.3, 0.7, 0.9], t, [2.0], nx=1, npaths=P, seed=5, XT=XT, path_integrals=PI,
                                     want_time=True))
        emit("functionals_ou_euler_generator", paths=P, nsteps=n, ms=ms, path_steps_per_s=P * n / ms * 1e3)
        PI2 = torch.empty(P * 8, dtype=torch.float64, device=dev)
        XT2 = torch.empty(P * 2, dtype=torch.float64, device=dev)
        ms = best_ms(lambda: eng.run("vdp", "heun", [1.0, 0.5], t, [0.1, -0.2], nx=2, npaths=P, seed=5, XT=XT2, path_integrals=PI2,
                                     want_time=True))
        emit("functionals_vdp_heun_generator", paths=P, nsteps=n, ms=ms, path_steps_per_s=P * n / ms * 1e3)
        del PI, XT, PI2, XT2
        torch.cuda.empty_cache()

    # rvBM ensemble: 8 B written per element (sector stores from registers) and its functionals (nothing written)
    if want("rvbm"):
        nt, nc = 1001, (1 << 16 if args.quick else 1 << 21)
        t = np.arange(nt) * 0.01
        Bd = torch.empty(nc * nt, dtype=torch.float64, device=dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.use_torch_stream()

        def timed_bm(fn):
            fn()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                ev0.record()
                fn()
                ev1.record()
                torch.cuda.synchronize()
                best = min(best, ev0.elapsed_time(ev1))
            return best
        ms = timed_bm(lambda: eng.rvbm(t, nc, Bd, seed=9))
        emit("rvbm", nt=nt, ncols=nc, ms=ms, elems_per_s=nt * nc / ms * 1e3, gbs=8.0 * nt * nc / ms / 1e6, hbm_frac=8.0 * nt * nc / ms / 1e6 / hbm)
        vmax = torch.empty(nc, dtype=torch.float64, device=dev)
        ms = timed_bm(lambda: eng.bm_functionals(t, nc, seed=9, vmax=vmax))
        emit("rvbm_functionals_max", nt=nt, ncols=nc, ms=ms, elems_per_s=nt * nc / ms * 1e3)
        eng.use_own_stream()
        del Bd, vmax
        torch.cuda.empty_cache()

    # next-1: killing / stopping (vignettes/Killing.Rmd:152-207 scaled up): OU with exp killing rate, lifetime only
    if want("killed"):
        import sdetools_b200.catalogue as cat
        P = (1 << 20) if args.quick else 1 << 24
        n = 1000
        t = np.arange(n + 1) * 0.01
        tau = torch.empty(P, dtype=torch.float64, device=dev)
        S = torch.empty(P, dtype=torch.float64, device=dev)
        XT = torch.empty(P, dtype=torch.float64, device=dev)
        for name, kw in (("killed_ou_exp_rate", dict(kill=cat.kill_exp(0.5, 1.0))),
                         ("stopped_ou_outside", dict(stop=cat.stop_outside(-1.0, 1.5))),
                         ("killed_and_stopped", dict(kill=cat.kill_quad(1.0, 0.0), stop=cat.stop_above(2.0)))):
            ms = best_ms(lambda: eng.run("ou", "euler", [1.0, 0.0, 1.0], t, [0.0], nx=1, npaths=P, seed=12, XT=XT, tau=tau,
                                         S_tau=S if "kill" in kw else None, want_time=True, **kw))
            alive = float(tau.sum()) / (P * t[-1])          # mean lifetime / horizon = fraction of path-steps actually taken
            emit(name, paths=P, nsteps=n, ms=ms, grid_path_steps_per_s=P * n / ms * 1e3, mean_lifetime_frac=alive,
                 taken_path_steps_per_s=alive * P * n / ms * 1e3)
        del tau, S, XT
        torch.cuda.empty_cache()

    # next-2: exact transition-law samplers, device-resident output (draws = transitions)
    if want("laws"):
        n = (1 << 20) if args.quick else 1 << 24
        out = torch.empty(n, dtype=torch.float64, device=dev)
        x0 = torch.full((1,), 1.0, dtype=torch.float64, device=dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.use_torch_stream()

        def timed_law(fn):
            fn()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                ev0.record()
                fn()
                ev1.record()
                torch.cuda.synchronize()
                best = min(best, ev0.elapsed_time(ev1))
            return best
        for name, law, params, t, steps in (("rOU_100_transitions", "ou", (1.3, 0.7, 0.9), 0.01, 100),
                                             ("rCIR_df4_t0.5", "cir", (1.0, 1.0, 1.0), 0.5, 10),
                                             ("rCIR_df4_dt1e-3_ptrs", "cir", (1.0, 1.0, 1.0), 1e-3, 10),
                                             ("rCIR_df0.18", "cir", (0.5, 0.2, 1.5), 1.0, 10)):
            ms = timed_law(lambda: eng.rlaw(law, params, t, x0, n, out, nsteps=steps, seed=1))
            emit(name, draws=n, transitions=steps, ms=ms, transitions_per_s=n * steps / ms * 1e3)
        eng.use_own_stream()
        del out
        torch.cuda.empty_cache()

    # linear SDE at small and medium d (register kernel for d, nB <= 4, generic kernel above)
    if want("linear_small"):
        for d, nB, P, n in ((2, 2, 1 << 20, 1000), (4, 2, 1 << 19, 1000), (8, 8, 1 << 18, 1000), (16, 16, 1 << 17, 500)):
            if args.quick:
                P >>= 3
            rng = np.random.default_rng(d)
            R = rng.standard_normal((d, d))
            A = -0.5 * np.eye(d) + 0.5 * (R - R.T) / np.sqrt(d)
            G = rng.standard_normal((d, nB)) * 0.5
            params = np.concatenate([A.ravel(order="F"), G.ravel(order="F")])
            XT = torch.empty(P * d, dtype=torch.float64, device=dev)
            t = np.arange(n + 1) / n
            for scheme in ("euler", "heun"):
                ms = best_ms(lambda: eng.run("linear", scheme, params, t, np.ones(d), nx=d, nB=nB, npaths=P, seed=1, XT=XT,
                                             want_time=True), reps=2)
                emit(f"linear_d{d}_nB{nB}_{scheme}", paths=P, nsteps=n, ms=ms, path_steps_per_s=P * n / ms * 1e3)
            del XT
        torch.cuda.empty_cache()

    # DMMA below and off the d = 256 headline: d = 64, 100 (padded to 128), 128
    if want("linear_mid"):
        for d, P, n in ((64, 8 * eng.info["sm_count"] * 64, 500), (100, 4 * eng.info["sm_count"] * 64, 500),
                        (128, 4 * eng.info["sm_count"] * 64, 500)):
            rng = np.random.default_rng(d)
            R = rng.standard_normal((d, d))
            A = -0.5 * np.eye(d) + (R - R.T) / np.sqrt(d)
            params = np.concatenate([A.ravel(order="F"), np.eye(d).ravel(order="F")])
            x0 = np.zeros(d)
            x0[0] = 1.0
            XT = torch.empty(P * d, dtype=torch.float64, device=dev)
            ms = best_ms(lambda: eng.run("linear", "euler", params, np.arange(n + 1) * (1.0 / n), x0, nx=d, nB=d, npaths=P, seed=4,
                                         XT=XT, want_time=True), reps=2)
            fl = 2.0 * d * d + 4 * d
            emit(f"linear_dmma_d{d}", paths=P, nsteps=n, ms=ms, path_steps_per_s=P * n / ms * 1e3, tflops=fl * P * n / ms / 1e9,
                 frac=fl * P * n / ms / 1e9 / fp64)
            # dense (lower-triangular) G: the increment term is a second GEMM; 2d^2 + d(d+1) + 2d flops per path-step
            Gd = np.tril(rng.standard_normal((d, d))) / np.sqrt(d)
            pd = np.concatenate([A.ravel(order="F"), Gd.ravel(order="F")])
            ms = best_ms(lambda: eng.run("linear", "euler", pd, np.arange(n + 1) * (1.0 / n), x0, nx=d, nB=d, npaths=P, seed=4,
                                         XT=XT, want_time=True), reps=2)
            fld = 4.0 * d * d + 2 * d
            emit(f"linear_dmma_dense_d{d}", paths=P, nsteps=n, ms=ms, path_steps_per_s=P * n / ms * 1e3,
                 tflops=fld * P * n / ms / 1e9, frac=fld * P * n / ms / 1e9 / fp64)
            del XT
        torch.cuda.empty_cache()

    # the general tensor kernel (k_linear_tc): heun at d = 256, supplied B at d = 256, small systems on DMMA tiles
    if want("linear_tc"):
        d = 256
        rng = np.random.default_rng(256)
        R = rng.standard_normal((d, d))
        A = -0.5 * np.eye(d) + (R - R.T) / np.sqrt(d)
        params = np.concatenate([A.ravel(order="F"), np.eye(d).ravel(order="F")])
        x0 = np.zeros(d)
        x0[0] = 1.0
        P, n = 2 * eng.info["sm_count"] * 32, 500
        XT = torch.empty(P * d, dtype=torch.float64, device=dev)
        ms = best_ms(lambda: eng.run("linear", "heun", params, np.arange(n + 1) * (1.0 / n), x0, nx=d, nB=d, npaths=P, seed=4,
                                     XT=XT, want_time=True), reps=2)
        fl = 2 * (2.0 * d * d) + 8 * d
        emit("c4_linear_d256_heun_tc", paths=P, nsteps=n, ms=ms, path_steps_per_s=P * n / ms * 1e3, tflops=fl * P * n / ms / 1e9,
             frac=fl * P * n / ms / 1e9 / fp64)
        Pb, nb = eng.info["sm_count"] * 64, 51
        tb = np.arange(nb) * 0.02
        B = torch.empty(Pb * d * nb, dtype=torch.float64, device=dev)
        eng.rvbm(tb, Pb * d, B, seed=2)
        for scheme in ("euler", "heun"):
            ms = best_ms(lambda: eng.run("linear", scheme, params, tb, x0, nx=d, nB=d, npaths=Pb, B=B, arith="fast", XT=XT[: Pb * d],
                                         want_time=True), reps=2)
            fls = (2.0 * d * d + 4 * d) * (2 if scheme == "heun" else 1)
            emit(f"linear_d256_suppliedB_{scheme}_tc", paths=Pb, nsteps=nb - 1, ms=ms, path_steps_per_s=Pb * (nb - 1) / ms * 1e3,
                 tflops=fls * Pb * (nb - 1) / ms / 1e9, b_gbs=8.0 * d * Pb * (nb - 1) / ms / 1e6)
            ms = best_ms(lambda: eng.run("linear", scheme, params, tb, x0, nx=d, nB=d, npaths=Pb, B=B, arith="strict", XT=XT[: Pb * d],
                                         want_time=True), reps=1)
            emit(f"linear_d256_suppliedB_{scheme}_strict_cuda_cores", paths=Pb, nsteps=nb - 1, ms=ms,
                 path_steps_per_s=Pb * (nb - 1) / ms * 1e3)
        del B, XT
        torch.cuda.empty_cache()
        # 256 < d <= 512: two row blocks per warp
        d5 = 512
        R5 = np.random.default_rng(512).standard_normal((d5, d5))
        A5 = -0.5 * np.eye(d5) + (R5 - R5.T) / np.sqrt(d5)
        p5 = np.concatenate([A5.ravel(order="F"), np.eye(d5).ravel(order="F")])
        x5 = np.zeros(d5)
        x5[0] = 1.0
        P5, n5 = 2 * eng.info["sm_count"] * 32, 200
        XT5 = torch.empty(P5 * d5, dtype=torch.float64, device=dev)
        for scheme in ("euler", "heun"):
            row = {}
            for knob, label in (("1", "tc"), ("0", "cuda_cores")):
                os.environ["SDEGPU_LINEAR_TC"] = knob
                PPk = P5 if knob == "1" else P5 // 8
                ms = best_ms(lambda: eng.run("linear", scheme, p5, np.arange(n5 + 1) * (1.0 / n5), x5, nx=d5, nB=d5, npaths=PPk, seed=4,
                                             XT=XT5[: PPk * d5], want_time=True), reps=1)
                row[label] = PPk * n5 / ms * 1e3
            os.environ.pop("SDEGPU_LINEAR_TC", None)
            fl5 = (2.0 * d5 * d5 + 4 * d5) * (2 if scheme == "heun" else 1)
            emit(f"linear_d512_{scheme}", paths=P5, nsteps=n5, path_steps_per_s_tc=row["tc"], path_steps_per_s_cuda_cores=row["cuda_cores"],
                 tflops_tc=fl5 * row["tc"] / 1e12, frac=fl5 * row["tc"] / 1e12 / fp64)
        del XT5
        torch.cuda.empty_cache()
        for dd, PP, nn in ((8, 1 << 20, 500), (16, 1 << 19, 500), (32, 1 << 18, 500)):
            Rr = np.random.default_rng(dd).standard_normal((dd, dd))
            Aa = -0.5 * np.eye(dd) + 0.5 * (Rr - Rr.T) / np.sqrt(dd)
            pp = np.concatenate([Aa.ravel(order="F"), (0.5 * np.eye(dd)).ravel(order="F")])
            XTs = torch.empty(PP * dd, dtype=torch.float64, device=dev)
            for scheme in ("euler", "heun"):
                row = {}
                for knob, label in (("1", "tc"), ("0", "cuda_cores")):
                    os.environ["SDEGPU_LINEAR_TC"] = knob
                    ms = best_ms(lambda: eng.run("linear", scheme, pp, np.arange(nn + 1) / nn, np.ones(dd), nx=dd, nB=dd, npaths=PP, seed=1,
                                                 XT=XTs, want_time=True), reps=2)
                    row[label] = PP * nn / ms * 1e3
                os.environ.pop("SDEGPU_LINEAR_TC", None)
                emit(f"linear_d{dd}_diag_{scheme}", paths=PP, nsteps=nn, path_steps_per_s_tc=row["tc"],
                     path_steps_per_s_cuda_cores=row["cuda_cores"], normals_per_s_tc=row["tc"] * dd)
            del XTs
        torch.cuda.empty_cache()

    # C4: linear SDE d = 256 (A x as a GEMM over paths); 2d^2+4d flops per path-step
    if want("c4"):
        d = 256
        P = (1 << 10) if args.quick else 2 * eng.info["sm_count"] * 64      # two full waves of 64-path CTAs
        n = 100 if args.quick else 1000
        rng = np.random.default_rng(256)
        R = rng.standard_normal((d, d))
        A = -0.5 * np.eye(d) + (R - R.T) / np.sqrt(d)
        G = np.eye(d)
        params = np.concatenate([A.ravel(order="F"), G.ravel(order="F")])
        x0 = np.zeros(d)
        x0[0] = 1.0
        XT = torch.empty(P * d, dtype=torch.float64, device=dev)
        ms = best_ms(lambda: eng.run("linear", "euler", params, np.arange(n + 1) * (1.0 / n), x0, nx=d, nB=d, npaths=P, seed=4,
                                     XT=XT, want_time=True), reps=2)
        steps = P * n
        fl = 2.0 * d * d + 4 * d
        emit("c4_linear_d256", paths=P, nsteps=n, ms=ms, path_steps_per_s=steps / ms * 1e3, tflops=fl * steps / ms / 1e9,
             fp64_peak=fp64, frac=fl * steps / ms / 1e9 / fp64)
        xt = XT.view(P, d)
        emit("c4_check", mean0=float(xt[:, 0].mean()), var_mean=float(xt.var(dim=0).mean()), var_expected=1 - np.exp(-1.0))
        if not args.quick:
            # BASELINE.json config 4 at its full size: 10^6 paths, checked against the exact law of the Euler recursion
            # m <- (I + A h) m, S <- (I + A h) S (I + A h)' + h I and against dLinSDE's S_t = (1 - e^-t) I, E X = e^{At} x0
            P6 = 1_000_000
            del XT
            XT = torch.empty(P6 * d, dtype=torch.float64, device=dev)
            ms = best_ms(lambda: eng.run("linear", "euler", params, np.arange(n + 1) * (1.0 / n), x0, nx=d, nB=d, npaths=P6, seed=4,
                                         XT=XT, want_time=True), reps=1)
            xt = XT.view(P6, d)
            h = 1.0 / n
            M = np.eye(d) + A * h
            m, S = x0.copy(), np.zeros((d, d))
            for _ in range(n):
                m = M @ m
                S = M @ S @ M.T + h * np.eye(d)
            C = torch.cov(xt.T).cpu().numpy()
            mean = xt.mean(dim=0).cpu().numpy()
            se_m = np.sqrt(np.diag(S) / P6)
            law = sde.dLinSDE(A, G, 1.0, x0=x0)
            emit("c4_full_size", paths=P6, nsteps=n, ms=ms, path_steps_per_s=P6 * n / ms * 1e3, tflops=fl * P6 * n / ms / 1e9,
                 frac=fl * P6 * n / ms / 1e9 / fp64,
                 max_mean_err_in_se=float(np.max(np.abs(mean - m) / se_m)),
                 max_cov_err_in_se=float(np.max(np.abs(C - S)) / (np.max(np.diag(S)) * np.sqrt(2.0 / P6))),
                 discrete_vs_dLinSDE_cov=float(np.max(np.abs(S - law["St"]))), mc_vs_dLinSDE_cov=float(np.max(np.abs(C - law["St"]))),
                 mc_vs_dLinSDE_mean=float(np.max(np.abs(mean - law["EX"]))))


if __name__ == "__main__":
    main()
