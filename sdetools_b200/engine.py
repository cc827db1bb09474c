"""Engine: one libsdegpu context on one GPU, with host (numpy) and device (torch
CUDA tensor) entry points.  PyTorch is plumbing only here: device memory, the
current stream and torch.distributed; all compute is in libsdegpu.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ARITH, MODEL, PROJ, SCHEME, Output, Problem, check


def _is_torch(x) -> bool:
    return x is not None and type(x).__module__.startswith("torch")


def _ptr(a):
    """Raw address of a numpy array or torch tensor (None -> NULL)."""
    if a is None:
        return None
    if _is_torch(a):
        return a.data_ptr()
    return a.ctypes.data


class Engine:
    def __init__(self, device=0):
        """device: one CUDA ordinal, or a sequence of ordinals for a multi-GPU context of this process
        (sdegpu_create_multi: the ensemble is split into contiguous path ranges, accumulators all-reduced)."""
        self.lib = _lib.load()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*[int(d) for d in device])
            check(self.lib.sdegpu_create_multi(ids, len(device), C.byref(h)))
            self.devices = [int(d) for d in device]
        else:
            check(self.lib.sdegpu_create(int(device), C.byref(h)))
            self.devices = [int(device)]
        self._h = h
        self.device = self.devices[0]
        self._intr_keep = None
        info = _lib.DevInfo()
        check(self.lib.sdegpu_device_info(self._h, C.byref(info)))
        self.info = {"name": info.name.decode(), "sm_count": info.sm_count, "cc": (info.cc_major, info.cc_minor),
                     "clock_khz": info.clock_khz, "global_mem_bytes": info.global_mem_bytes}

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sdegpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- streams ------------------------------------------------------------
    def use_torch_stream(self):
        """Launch on torch's current stream so torch.cuda.Event timing sees the kernels."""
        import torch
        check(self.lib.sdegpu_set_stream(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def use_own_stream(self):
        check(self.lib.sdegpu_reset_stream(self._h))

    def synchronize(self):
        check(self.lib.sdegpu_synchronize(self._h))

    # -- chunking and interruption of host-pointer runs -------------------------
    def set_interrupt(self, fn=None):
        """fn() -> truthy stops the running sdegpu_run between two chunks (SdeGpuError code -5); None removes it."""
        if fn is None:
            self._intr_keep = None
            check(self.lib.sdegpu_set_interrupt(self._h, None, None))
            return
        cb = _lib.INTERRUPT_FN(lambda _user: 1 if fn() else 0)
        self._intr_keep = cb                                   # ctypes callbacks must outlive their registration
        check(self.lib.sdegpu_set_interrupt(self._h, C.cast(cb, C.c_void_p), None))

    def set_chunk_paths(self, paths: int = 0):
        check(self.lib.sdegpu_set_chunk_paths(self._h, int(paths)))

    @property
    def last_reduction(self) -> str:
        return self.lib.sdegpu_last_reduction(self._h).decode()

    # -- ensembles ------------------------------------------------------------
    def run(self, model, scheme, params, times, x0, *, nx, nB=1, npaths, B=None, projection="identity",
            arith=None, seed=0, path_offset=0, X=None, XT=None, power_sums=None, center=None,
            hist_counts=None, hist=None, accumulate=False, want_time=False, kill=None, stop=None, S0=1.0,
            Stau=None, tau=None, S_tau=None, path_integrals=None, exact_grid=False, S_path=None, model_integrals=None,
            drift_forcing=None, diffusion_scale=None):
        """Thin, allocation-free call of sdegpu_run.  B / per-path x0 / outputs are either all numpy
        (host pointers) or all torch CUDA tensors (device pointers: work is only enqueued, on torch's current stream).
        Layouts are the reference's: B[(p*nB+k)*nt+i], X[(p*nx+j)*nt+i], XT[p*nx+j]."""
        bufs = [b for b in (B, X, XT, power_sums, hist_counts, Stau, tau, S_tau, path_integrals, S_path, model_integrals) if b is not None]
        on_dev = any(_is_torch(b) for b in bufs)
        if on_dev and not all(_is_torch(b) and b.is_cuda for b in bufs):
            raise TypeError("mixing host and device buffers in one call is not supported")
        for b in bufs:
            if _is_torch(b):
                if not b.is_contiguous():
                    raise ValueError("device buffers must be contiguous")
            elif not b.flags["C_CONTIGUOUS"] and not b.flags["F_CONTIGUOUS"]:
                raise ValueError("host buffers must be contiguous")
            if b is not hist_counts and str(b.dtype).replace("torch.", "") != "float64":
                raise TypeError(f"buffers must be float64, got {b.dtype}")
        params = np.ascontiguousarray(np.ravel(params), dtype=np.float64)
        times = np.ascontiguousarray(times, dtype=np.float64)
        per_path_x0 = _is_torch(x0) or (np.size(x0) != nx)
        if per_path_x0:
            if _is_torch(x0) != on_dev and bufs:
                raise TypeError("per-path x0 must live where the other buffers live")
            x0_keep = x0 if _is_torch(x0) else np.ascontiguousarray(x0, dtype=np.float64)
            on_dev = on_dev or _is_torch(x0)
        else:
            x0_keep = np.ascontiguousarray(np.ravel(x0), dtype=np.float64)
        if arith is None:
            arith = "strict" if B is not None else "fast"
        pr = Problem()
        pr.struct_size = C.sizeof(Problem)
        pr.model, pr.scheme = MODEL[model], SCHEME[scheme]
        pr.arith, pr.projection = ARITH[arith], PROJ[projection]
        pr.nx, pr.nB, pr.nparams = int(nx), int(nB), int(params.size)
        pr.params, pr.times, pr.nt = _ptr(params), _ptr(times), int(times.size)
        pr.x0, pr.x0_stride = _ptr(x0_keep), (int(nx) if per_path_x0 else 0)
        pr.B = _ptr(B)
        pr.npaths, pr.path_offset, pr.seed = int(npaths), int(path_offset), int(seed) & 0xFFFFFFFFFFFFFFFF
        if on_dev:
            self._follow_torch_stream()
        pr.flags = (_lib.DEVICE_PTRS if on_dev else 0) | (_lib.ACCUMULATE if accumulate else 0) | (_lib.EXACT_GRID if exact_grid else 0)
        if kill is not None:                 # catalogue.KillRate
            pr.kill_kind, pr.kill_comp, pr.kill_a, pr.kill_b = _lib.KILL[kill.sdegpu_kill], kill.component, kill.a, kill.b
        if stop is not None:                 # catalogue.StopRule
            pr.stop_kind, pr.stop_comp, pr.stop_lo, pr.stop_hi = _lib.STOP[stop.sdegpu_stop], stop.component, stop.lo, stop.hi
        pr.S0 = float(S0)
        pr.Stau = _ptr(Stau)
        forcing_keep = scale_keep = None
        if drift_forcing is not None:                        # c[j*nt + i]: column-major nt x nx
            forcing_keep = np.ascontiguousarray(np.asarray(drift_forcing, dtype=np.float64).reshape(times.size, int(nx)).T)
            pr.drift_forcing = _ptr(forcing_keep)
        if diffusion_scale is not None:
            scale_keep = np.ascontiguousarray(np.asarray(diffusion_scale, dtype=np.float64).reshape(times.size))
            pr.diffusion_scale = _ptr(scale_keep)
        out = Output()
        out.struct_size = C.sizeof(Output)
        out.X, out.XT, out.power_sums, out.hist_counts = _ptr(X), _ptr(XT), _ptr(power_sums), _ptr(hist_counts)
        out.tau, out.S_tau, out.path_integrals, out.S = _ptr(tau), _ptr(S_tau), _ptr(path_integrals), _ptr(S_path)
        out.model_integrals = _ptr(model_integrals)
        center_keep = None
        if center is not None:
            center_keep = np.ascontiguousarray(np.ravel(center), dtype=np.float64)
            out.center = _ptr(center_keep)
        if hist is not None:
            lo, hi, nbins = hist[:3]
            out.hist_lo, out.hist_hi, out.hist_nbins = float(lo), float(hi), int(nbins)
            out.hist_component = int(hist[3]) if len(hist) > 3 else 0
        ms = C.c_float(0.0)
        if want_time:
            out.elapsed_ms = C.addressof(ms)
        check(self.lib.sdegpu_run(self._h, C.byref(pr), C.byref(out)))
        return ms.value if want_time else None

    # -- integrals --------------------------------------------------------------
    def _flags(self, *bufs):
        dev = [_is_torch(b) for b in bufs if b is not None]
        if any(dev) and not all(dev):
            raise TypeError("mixing host and device buffers in one call is not supported")
        if any(dev):
            self._follow_torch_stream()
        return _lib.DEVICE_PTRS if any(dev) else 0

    def _follow_torch_stream(self):
        """Calls on torch tensors are only enqueued; they go to torch's current stream so that they are ordered
        with the torch operations that produce and consume those tensors."""
        self.use_torch_stream()

    def stochint(self, f, g, n, ncols, l=None, r=None, c=None):
        check(self.lib.sdegpu_stochint(self._h, _ptr(f), _ptr(g), n, ncols, _ptr(l), _ptr(r), _ptr(c), self._flags(f, g, l, r, c)))

    def itointegral(self, G, B, n, ncols, out):
        check(self.lib.sdegpu_itointegral(self._h, _ptr(G), _ptr(B), n, ncols, _ptr(out), self._flags(G, B, out)))

    def qv(self, X, n, ncols, out):
        check(self.lib.sdegpu_qv(self._h, _ptr(X), n, ncols, _ptr(out), self._flags(X, out)))

    def crossvar(self, X, Y, n, ncols, out):
        check(self.lib.sdegpu_crossvar(self._h, _ptr(X), _ptr(Y), n, ncols, _ptr(out), self._flags(X, Y, out)))

    def rvbm(self, times, ncols, out, sigma=1.0, B0=0.0, u=0.0, seed=0, path_offset=0):
        times = np.ascontiguousarray(times, dtype=np.float64)
        check(self.lib.sdegpu_rvbm(self._h, _ptr(times), int(times.size), int(ncols), float(sigma), float(B0), float(u),
                                   int(seed) & 0xFFFFFFFFFFFFFFFF, int(path_offset), _ptr(out), self._flags(out)))

    def bm_functionals(self, times, ncols, sigma=1.0, B0=0.0, u=0.0, level=0.0, seed=0, path_offset=0, vmax=None, vmin=None,
                       tmax=None, tmin=None, thit=None):
        times = np.ascontiguousarray(times, dtype=np.float64)
        check(self.lib.sdegpu_bm_functionals(self._h, _ptr(times), int(times.size), int(ncols), float(sigma), float(B0), float(u),
                                             float(level), int(seed) & 0xFFFFFFFFFFFFFFFF, int(path_offset), _ptr(vmax),
                                             _ptr(vmin), _ptr(tmax), _ptr(tmin), _ptr(thit),
                                             self._flags(vmax, vmin, tmax, tmin, thit)))

    def rlaw(self, law, params, t, x0, n, out, nsteps=1, path=None, stratonovich=False, seed=0, path_offset=0):
        """Exact-law sampler (sdegpu_rlaw): law "ou" | "cir", params = (lambda, xi, sigma|gamma); x0 scalar or one per draw."""
        p = (C.c_double * 3)(*[float(v) for v in params])
        if _is_torch(x0):
            stride = 0 if x0.numel() == 1 else 1
        else:
            x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1)
            stride = 0 if x0.size == 1 else 1
            if stride and x0.size != n:
                raise ValueError("x0 must be a scalar or have one entry per draw")
        check(self.lib.sdegpu_rlaw(self._h, {"ou": 0, "cir": 1}[law], p, float(t), int(bool(stratonovich)), _ptr(x0), stride, int(n),
                                   int(nsteps), int(seed) & 0xFFFFFFFFFFFFFFFF, int(path_offset), _ptr(out), _ptr(path),
                                   self._flags(x0, out, path)))

    def law_eval(self, law, what, params, t, x0, x, out, stratonovich=False):
        """sdegpu_law_eval: what = "d" density | "p" distribution function of law "ou" | "cir" at x given x0."""
        p = (C.c_double * 3)(*[float(v) for v in params])
        n = int(x.numel()) if _is_torch(x) else int(x.size)
        if _is_torch(x0):
            stride = 0 if x0.numel() == 1 else 1
        else:
            x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1)
            stride = 0 if x0.size == 1 else 1
            if stride and x0.size != n:
                raise ValueError("x0 must be a scalar or have one entry per point")
        check(self.lib.sdegpu_law_eval(self._h, {"ou": 0, "cir": 1}[law], {"d": 0, "p": 1}[what], p, float(t),
                                       int(bool(stratonovich)), _ptr(x0), stride, _ptr(x), n, _ptr(out), self._flags(x0, x, out)))

    # -- linear-SDE laws ----------------------------------------------------------
    def lyap(self, A, Q):
        """sdegpu_lyap: X with A X + X A' + Q = 0 (column-major host matrices)."""
        A = np.asfortranarray(A, dtype=np.float64)
        Q = np.asfortranarray(Q, dtype=np.float64)
        X = np.empty_like(A, order="F")
        check(self.lib.sdegpu_lyap(self._h, _ptr(A), _ptr(Q), A.shape[0], _ptr(X)))
        return X

    def dlinsde(self, A, G, t, x0=None, u=None, S0=None):
        """sdegpu_dlinsde: (eAt | EX, St) of dX = (A X + u) dt + G dB over time t."""
        A = np.asfortranarray(A, dtype=np.float64)
        d = A.shape[0]
        G = np.asfortranarray(np.asarray(G, dtype=np.float64).reshape(d, -1))
        x0k = None if x0 is None else np.ascontiguousarray(np.ravel(x0), dtype=np.float64)
        uk, ucols = None, 0
        if u is not None:
            uk = np.asfortranarray(np.asarray(u, dtype=np.float64).reshape(d, -1))
            ucols = uk.shape[1]
        S0k = None if S0 is None else np.asfortranarray(S0, dtype=np.float64)
        St = np.empty((d, d), order="F")
        first = np.empty(d) if x0 is not None else np.empty((d, d), order="F")
        check(self.lib.sdegpu_dlinsde(self._h, _ptr(A), _ptr(G), d, G.shape[1], float(t), _ptr(x0k), _ptr(uk), ucols, _ptr(S0k),
                                      None if x0 is not None else _ptr(first), _ptr(first) if x0 is not None else None, _ptr(St)))
        return first, St

    # -- introspection ------------------------------------------------------------
    def philox4x32_10(self, ctr, key):
        a = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        check(self.lib.sdegpu_philox4x32_10(self._h, a, k, o))
        return tuple(o)

    def normals(self, seed, path_offset, npaths, nsteps, channel=0):
        z = np.empty((nsteps, npaths))
        check(self.lib.sdegpu_normals(self._h, int(seed), int(path_offset), int(npaths), int(channel), int(nsteps), _ptr(z)))
        return z

    def measure_fp64_peak(self) -> float:
        v = C.c_double(0.0)
        check(self.lib.sdegpu_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def measure_hbm_peak(self) -> float:
        v = C.c_double(0.0)
        check(self.lib.sdegpu_measure_hbm_peak(self._h, C.byref(v)))
        return v.value


_engines = {}


def default_engine(device=None) -> Engine:
    """Process-wide engine per device (device=None: LOCAL_RANK or 0)."""
    import os
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if isinstance(device, (list, tuple)):
        device = tuple(int(d) for d in device)                 # a multi-GPU context of this process
        if len(device) == 1:
            device = device[0]
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]
