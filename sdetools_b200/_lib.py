"""ctypes binding of libsdegpu.so (include/sdegpu.h).  There is no CPU fallback:
if the library or a B200 is missing, calls raise."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libsdegpu.so")

ABI_VERSION = 5
DEVICE_PTRS = 1
ACCUMULATE = 2
EXACT_GRID = 4

MODEL = {"ou": 0, "cir": 1, "vdp": 2, "logistic": 3, "gbm": 4, "doublewell": 5, "linear": 6, "dwsqrt": 7, "loglogistic": 8}
SCHEME = {"euler": 0, "heun": 1}
PROJ = {"identity": 0, "abs": 1, "relu": 2}
ARITH = {"strict": 0, "fast": 1}
KILL = {"none": 0, "const": 1, "exp": 2, "quad": 3}
STOP = {"none": 0, "below": 1, "above": 2, "outside": 3}

_dp = C.POINTER(C.c_double)


class Problem(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("model", C.c_int32), ("scheme", C.c_int32), ("arith", C.c_int32),
                ("projection", C.c_int32), ("nx", C.c_int32), ("nB", C.c_int32), ("nparams", C.c_int32),
                ("params", C.c_void_p), ("nt", C.c_int32), ("x0_stride", C.c_int32), ("times", C.c_void_p),
                ("x0", C.c_void_p), ("B", C.c_void_p), ("npaths", C.c_uint64), ("path_offset", C.c_uint64),
                ("seed", C.c_uint64), ("flags", C.c_uint32), ("reserved", C.c_uint32),
                ("kill_kind", C.c_int32), ("kill_comp", C.c_int32), ("kill_a", C.c_double), ("kill_b", C.c_double),
                ("stop_kind", C.c_int32), ("stop_comp", C.c_int32), ("stop_lo", C.c_double), ("stop_hi", C.c_double),
                ("S0", C.c_double), ("Stau", C.c_void_p), ("drift_forcing", C.c_void_p), ("diffusion_scale", C.c_void_p)]


class Output(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("hist_component", C.c_int32), ("X", C.c_void_p), ("XT", C.c_void_p),
                ("power_sums", C.c_void_p), ("center", C.c_void_p), ("hist_counts", C.c_void_p),
                ("hist_lo", C.c_double), ("hist_hi", C.c_double), ("hist_nbins", C.c_int32), ("reserved", C.c_int32),
                ("elapsed_ms", C.c_void_p), ("tau", C.c_void_p), ("S_tau", C.c_void_p), ("path_integrals", C.c_void_p),
                ("S", C.c_void_p), ("model_integrals", C.c_void_p)]


class DevInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("clock_khz", C.c_int32), ("global_mem_bytes", C.c_uint64)]


INTERRUPT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)          # sdegpu_interrupt_fn
ERR_INTERRUPT = -5


class SdeGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libsdegpu error {code}: {message}")
        self.code = code


# every symbol include/sdegpu.h declares: (restype, argtypes)
_u64, _u32, _i32, _vp = C.c_uint64, C.c_uint32, C.c_int32, C.c_void_p
SYMBOLS = {
    "sdegpu_abi_version": (C.c_int, []),
    "sdegpu_last_error": (C.c_char_p, []),
    "sdegpu_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "sdegpu_create_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "sdegpu_device_count": (C.c_int, [_vp]),
    "sdegpu_last_reduction": (C.c_char_p, [_vp]),
    "sdegpu_set_interrupt": (C.c_int, [_vp, _vp, _vp]),
    "sdegpu_set_chunk_paths": (C.c_int, [_vp, _u64]),
    "sdegpu_destroy": (C.c_int, [_vp]),
    "sdegpu_device_info": (C.c_int, [_vp, C.POINTER(DevInfo)]),
    "sdegpu_set_stream": (C.c_int, [_vp, _vp]),
    "sdegpu_reset_stream": (C.c_int, [_vp]),
    "sdegpu_synchronize": (C.c_int, [_vp]),
    "sdegpu_run": (C.c_int, [_vp, C.POINTER(Problem), C.POINTER(Output)]),
    "sdegpu_stochint": (C.c_int, [_vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _u32]),
    "sdegpu_itointegral": (C.c_int, [_vp, _vp, _vp, _u64, _u64, _vp, _u32]),
    "sdegpu_qv": (C.c_int, [_vp, _vp, _u64, _u64, _vp, _u32]),
    "sdegpu_crossvar": (C.c_int, [_vp, _vp, _vp, _u64, _u64, _vp, _u32]),
    "sdegpu_rvbm": (C.c_int, [_vp, _vp, _i32, _u64, C.c_double, C.c_double, C.c_double, _u64, _u64, _vp, _u32]),
    "sdegpu_bm_functionals": (C.c_int, [_vp, _vp, _i32, _u64, C.c_double, C.c_double, C.c_double, C.c_double, _u64, _u64,
                                        _vp, _vp, _vp, _vp, _vp, _u32]),
    "sdegpu_rlaw": (C.c_int, [_vp, _i32, _dp, C.c_double, _i32, _vp, C.c_int64, _u64, _i32, _u64, _u64, _vp, _vp, _u32]),
    "sdegpu_law_eval": (C.c_int, [_vp, _i32, _i32, _dp, C.c_double, _i32, _vp, C.c_int64, _vp, _u64, _vp, _u32]),
    "sdegpu_lyap": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "sdegpu_dlinsde": (C.c_int, [_vp, _vp, _vp, _i32, _i32, C.c_double, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "sdegpu_philox4x32_10": (C.c_int, [_vp, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]),
    "sdegpu_normals": (C.c_int, [_vp, _u64, _u64, _u64, _u32, _i32, _vp]),
    "sdegpu_measure_fp64_peak": (C.c_int, [_vp, _dp]),
    "sdegpu_measure_hbm_peak": (C.c_int, [_vp, _dp]),
}

_lib = None


def load():
    """Load libsdegpu.so; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: build it with `python -m sdetools_b200.build` "
                          "(nvcc, sm_100a). sdetools_b200 has no CPU fallback.")
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)            # AttributeError if the ABI lost a symbol
        fn.restype, fn.argtypes = res, args
    if lib.sdegpu_abi_version() != ABI_VERSION:
        raise ImportError(f"libsdegpu.so ABI {lib.sdegpu_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise SdeGpuError(rc, load().sdegpu_last_error().decode("utf-8", "replace"))
