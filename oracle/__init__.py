"""CPU oracle for the SDEtools sample-path hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference's algorithm
(`/root/reference/R/sdetools.R`, interpreted R, no native code).  It exists to
check the CUDA product in `sdetools_b200/`; nothing in the product path may
import it.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs call into it.

PARITY: the reference's own tests (`tests/testthat/test_solvers.R:3-18`)
assert only `dim(sol$X)` for `euler`; they hold no numeric golden vector for
`euler`/`heun`/`stochint`/`itointegral`, and GNU R is not installed in this image.
The restatement is therefore pinned to the reference's SOURCE TEXT executed by
`oracle/minir` (an evaluator for the R subset the reference is written in; it
parses the unmodified `R/sdetools.R`): `tests/golden/reference_minir/` holds its
outputs, `tests/test_reference_fixtures.py` and `tests/test_minir.py` compare this
oracle (and the CUDA path) with them bit for bit.  Further pins: (i) the six
`dLinSDE` known answers of `test_solvers.R:20-36`, (ii) the machine-precision
identities the vignettes assert (`vignettes/ItoIntegrals.Rmd:82-115,143-153`),
(iii) the analytic laws `dOU`/`dCIR`/`lyap` in the same file (`tests/test_oracle.py`).
Not pinned: GNU R's own primitives (libm, BLAS summation order) — the recipe for a
box with R is `tests/golden/make_reference_fixtures.R`.
"""
