"""CPU oracle for the SDEtools sample-path hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference's algorithm
(`/root/reference/R/sdetools.R`, interpreted R, no native code).  It exists to
check the CUDA product in `sdetools_b200/`; nothing in the product path may
import it.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs call into it.

PARITY UNPINNED: the reference's own tests (`tests/testthat/test_solvers.R:3-18`)
assert only `dim(sol$X)` for `euler`; they hold no numeric golden vector for
`euler`/`heun`/`stochint`/`itointegral`, and R is not installed in this image
so the reference itself cannot be run to generate one.  The restatement is
pinned instead by (i) the six `dLinSDE` known answers of
`test_solvers.R:20-36`, (ii) the machine-precision identities the vignettes
assert (`vignettes/ItoIntegrals.Rmd:82-115,143-153`), and (iii) the analytic
laws `dOU`/`dCIR`/`lyap` in the same file.  See `tests/test_oracle.py`.
"""
