"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads, exports every
symbol include/sdegpu.h declares, the ctypes structs match the header, the host mirror keeps
the reference's signatures, and nothing computes without a GPU (no fallback)."""
import ctypes as C
import inspect
import os
import re
import subprocess

import numpy as np
import pytest

import sdetools_b200 as sde
from sdetools_b200 import _lib, build, catalogue

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sdegpu.h")


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    decl = set(re.findall(r"SDEGPU_API\s+[\w\s\*]+?\b(sdegpu_\w+)\s*\(", open(HEADER).read()))
    assert decl == set(_lib.SYMBOLS), (decl ^ set(_lib.SYMBOLS))
    exported = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True).stdout
    for name in decl:
        assert re.search(rf"\bT {name}\b", exported), f"{name} not exported"
    assert lib.sdegpu_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sdegpu.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(sdegpu_problem),sizeof(sdegpu_output),sizeof(sdegpu_devinfo),offsetof(sdegpu_problem,B),"
                   "offsetof(sdegpu_problem,seed),offsetof(sdegpu_output,hist_lo),offsetof(sdegpu_output,elapsed_ms));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    P, O = _lib.Problem, _lib.Output
    assert got == [C.sizeof(P), C.sizeof(O), C.sizeof(_lib.DevInfo), P.B.offset, P.seed.offset, O.hist_lo.offset,
                   O.elapsed_ms.offset]


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "h.c"
    src.write_text('#include "sdegpu.h"\nint main(void){return SDEGPU_ABI_VERSION-1;}\n')
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(src), "-o", str(tmp_path / "h.o")], check=True)


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_reference_signatures_are_kept():
    ref = ["f", "g", "times", "x0", "B", "p", "h", "r", "S0", "Stau"]       # R/sdetools.R:375, :240
    for fn in (sde.euler, sde.heun):
        ps = inspect.signature(fn).parameters
        assert list(ps)[:10] == ref
        assert all(p.kind is inspect.Parameter.KEYWORD_ONLY for p in list(ps.values())[10:])
    assert list(inspect.signature(sde.stochint).parameters)[:3] == ["f", "g", "rule"]      # :105
    assert list(inspect.signature(sde.itointegral).parameters)[:2] == ["G", "B"]           # :155
    assert list(inspect.signature(sde.QuadraticVariation).parameters)[:1] == ["X"]         # :134
    assert list(inspect.signature(sde.CrossVariation).parameters)[:2] == ["X", "Y"]        # :184
    assert list(inspect.signature(sde.rBM).parameters)[:4] == ["times", "sigma", "B0", "u"]          # :16
    assert list(inspect.signature(sde.rvBM).parameters)[:5] == ["times", "n", "sigma", "B0", "u"]    # :40


def test_catalogue_closures_and_dispatch_errors():
    f, g = catalogue.cir(1.0, 2.0, 0.5)
    assert f(1.5) == pytest.approx(0.5) and g(0.0, 4.0) == pytest.approx(1.0)      # f(x) and g(t,x) forms
    assert f.sdegpu_model == "cir" and g.sdegpu_params.tolist() == [1.0, 2.0, 0.5]
    fv, gv = catalogue.vdp(1.0, 0.5)
    assert fv(np.array([1.0, 2.0])).tolist() == [2.0, -1.0] and gv(np.zeros(2)).tolist() == [0.0, 0.5]
    assert catalogue.projection_id(abs) == "abs" and catalogue.projection_id(None) == "identity"
    with pytest.raises(TypeError):
        catalogue.projection_id(lambda x: x + 1)
    with pytest.raises(TypeError):                     # arbitrary closures must fail loudly, not fall back
        sde.euler(lambda x: -x, lambda x: 1.0, [0, 1, 2], 0.0)
    with pytest.raises(TypeError):
        sde.euler(f, catalogue.ou()[1], [0, 1, 2], 0.0)


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.sdegpu_create(0, C.byref(h)) == -4          # SDEGPU_ERR_NODEVICE
    assert b"no CUDA device" in lib.sdegpu_last_error()
    with pytest.raises(_lib.SdeGpuError):
        sde.euler(*catalogue.ou(), [0.0, 0.1, 0.2], 0.0)
    with pytest.raises(_lib.SdeGpuError):
        sde.itointegral(np.ones(4), np.arange(4.0))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sdetools_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle/|sde_oracle", txt, flags=re.M), fn


def test_float80_emulation_matches_x87(tmp_path):
    """The device's integer emulation of R's 80-bit cumsum accumulator, compiled for the host and
    compared against the real FPU on 10^6 adversarial partial sums."""
    exe = tmp_path / "f80"
    subprocess.run(["/usr/bin/g++", "-O2", "-o", str(exe), os.path.join(ROOT, "tests", "native", "float80_check.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "mismatches=0" in r.stdout, r.stdout + r.stderr


def test_icdf_table_matches_qnorm():
    """The quantile table the fused generator evaluates on the device (csrc/icdf_table.inc), evaluated
    here exactly as philox.cuh does, against -qnorm(u/2)."""
    from scipy.special import ndtri
    txt = open(os.path.join(ROOT, "sdetools_b200", "csrc", "icdf_table.inc")).read()
    nseg = int(re.search(r"#define ICDF_NSEG (\d+)", txt).group(1))
    shift = int(re.search(r"#define ICDF_SHIFT (\d+)", txt).group(1))
    base = int(re.search(r"#define ICDF_BASE (\d+)", txt).group(1))
    body = txt[txt.index("] = {") + 5:txt.index("};")]
    words = np.array([int(v.strip().rstrip("ul"), 16) for v in body.split(",") if v.strip()], dtype=np.uint64)
    planes = words.view(np.float64).reshape(2, nseg, 2)          # plane A = {c0,c1}, plane B = {c2,c3}
    c0, c1, c2, c3 = planes[0, :, 0], planes[0, :, 1], planes[1, :, 0], planes[1, :, 1]
    rng = np.random.default_rng(0)
    m = np.concatenate([rng.integers(0, 2 ** 52, 400000), 2 ** rng.integers(0, 52, 2000) + rng.integers(0, 3, 2000),
                        np.arange(0, 4), 2 ** 52 - 1 - np.arange(0, 4)]).astype(np.uint64)
    u = m.astype(np.float64) * 2.0 ** -52                       # what y0 - 1.0 gives on the device
    bits = u.view(np.uint64)
    uh = (bits >> np.uint64(32)).astype(np.int64)
    t = np.maximum((uh >> shift) - base, 0)
    dh = (uh & ((1 << shift) - 1)) | 0x3FF00000                  # low mantissa of u re-based at 1
    d = ((dh.astype(np.uint64) << np.uint64(32)) | (bits & np.uint64(0xFFFFFFFF))).view(np.float64)
    r = d - (1.0 + 2.0 ** -(21 - shift))
    a = ((c3[t] * r + c2[t]) * r + c1[t]) * r + c0[t]
    want = -ndtri(0.5 * np.where(m == 0, 2.0 ** -52, u))
    assert t.max() < nseg and np.max(np.abs(r)) <= 2.0 ** -(21 - shift)
    assert np.max(np.abs(a - want)) < 1e-9
    assert np.all(a > -1e-9)          # next to u = 1 the true value (~1e-16) is below the table's error


def test_icdf32_table_matches_qnorm():
    """The fast quantile table of stream v2 (csrc/icdf32_table.inc), evaluated here exactly as NormalGen::quad does —
    s = (w|1) - 2^31 in FP64, segment from the high word of s, cubic in the raw |s| — against -qnorm(u/2), u = |s| 2^-31."""
    from scipy.special import ndtri
    txt = open(os.path.join(ROOT, "sdetools_b200", "csrc", "icdf32_table.inc")).read()
    nseg = int(re.search(r"#define ICDF32_NSEG (\d+)", txt).group(1))
    oct0 = int(re.search(r"#define ICDF32_OCT0 (\d+)", txt).group(1))
    body = txt[txt.index("] = {") + 5:txt.index("};")]
    words = np.array([int(v.strip().rstrip("ul"), 16) for v in body.split(",") if v.strip()], dtype=np.uint64)
    planes = words.view(np.float64).reshape(2, nseg, 2)
    c0, c1, c2, c3 = planes[0, :, 0], planes[0, :, 1], planes[1, :, 0], planes[1, :, 1]
    rng = np.random.default_rng(1)
    w = np.concatenate([rng.integers(0, 2 ** 32, 400000), [0, 1, 2 ** 32 - 1, 2 ** 31 + 2 ** 19, 2 ** 31 - 2 ** 19 - 2,
                                                            2 ** 31 + 2 ** 19 + 1, 2 ** 31 - 2 ** 19 - 1]]).astype(np.uint64)
    magic = ((np.uint64(0x43300000) << np.uint64(32)) | (w | np.uint64(1))).view(np.float64)
    s = magic - 4503601774854144.0                                # exact: an odd integer in (-2^31, 2^31)
    assert np.array_equal(s, (w | np.uint64(1)).astype(np.float64) - 2.0 ** 31)
    sh = (s.view(np.uint64) >> np.uint64(32)).astype(np.int64)
    xm = sh & 0x7FFF8000
    fast = xm >= 0x41200000                                       # |s| >= 2^19: the fast path's domain
    assert 0.9995 < fast.mean() < 1.0
    t = (xm[fast] >> 15) - ((1023 + oct0) << 5)
    assert t.min() >= 0 and t.max() < nseg
    v = np.abs(s[fast])
    a = ((c3[t] * v + c2[t]) * v + c1[t]) * v + c0[t]
    want = -ndtri(0.5 * v * 2.0 ** -31)
    assert np.max(np.abs(a - want)) < 1e-9


def test_r_shim_compiles_against_mock_r_headers():
    """R is not installed: the .Call shim (r/src/sdegpu_shim.c) is type-checked against a mock of R's C API."""
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "mock_r"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r", "src", "sdegpu_shim.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(ROOT, "r", "src", "sdegpu_shim.c")).read()
    used = set(re.findall(r"\b(sdegpu_[a-z0-9_]+)\s*\(", src))
    assert used <= set(_lib.SYMBOLS), used - set(_lib.SYMBOLS)
