#!/usr/bin/env Rscript
## make_reference_fixtures.R — pins parity to the REFERENCE ITSELF.
##
## R is not installed in the build image, so the committed fixtures under tests/golden/ are outputs of the
## line-by-line restatement (oracle/sde_oracle.py), not of R.  This script closes that gap on any box that has R:
##
##     Rscript tests/golden/make_reference_fixtures.R /path/to/SDEtools        # the reference checkout
##     python -m pytest tests/test_reference_fixtures.py                      # oracle + (with a B200) the CUDA path
##
## It sources the UNMODIFIED reference file R/sdetools.R, runs euler()/heun() (R/sdetools.R:375-470, :240-338),
## stochint/itointegral/QuadraticVariation/CrossVariation (:105-113, :155-158, :134-137, :184-187) and rBM (:16-22,
## with rnorm replaced by a function that returns mean + sd * z for recorded normals z) on the inputs that
## tests/golden/make_golden.py wrote to tests/golden/reference_R_inputs/ as C99 hexadecimal doubles, and writes
## every result to tests/golden/reference_R/ in the same format.  Nothing is rounded on the way: sprintf("%a")
## prints the exact bits, as.numeric("0x1.8p+1") reads them back.
##
## The drift/diffusion closures below are the catalogue's R closures (r/R/sdegpu.R), written out so that this
## script depends on nothing but base R and the reference file.

args <- commandArgs(trailingOnly = TRUE)
ref <- if (length(args) >= 1) args[1] else Sys.getenv("SDETOOLS_REF", "/root/reference")
here <- if (length(args) >= 2) args[2] else {
  f <- sub("^--file=", "", grep("^--file=", commandArgs(FALSE), value = TRUE))
  if (length(f)) dirname(normalizePath(f)) else "tests/golden"
}
indir <- file.path(here, "reference_R_inputs")
outdir <- file.path(here, if (length(args) >= 3) args[3] else "reference_R")
dir.create(outdir, showWarnings = FALSE)

## recorded normals for rBM: the reference calls rnorm(n, mean, sd) (R/sdetools.R:19); sourced into the global
## environment, rBM finds this definition before stats::rnorm
.z <- NULL; .zpos <- 0                  # successive rnorm() calls (rvBM: one per channel) walk through .z
rnorm <- function(n, mean = 0, sd = 1) { i <- .zpos + seq_len(n); .zpos <<- .zpos + n; mean + sd * .z[i] }

source(file.path(ref, "R", "sdetools.R"))

read.hex <- function(name) {
  l <- readLines(file.path(indir, paste0(name, ".txt")))
  d <- as.integer(strsplit(l[1], " ")[[1]])
  v <- as.numeric(l[-1])
  stopifnot(length(v) == prod(d))
  if (length(d) > 1) array(v, d) else v
}
write.hex <- function(name, x) {
  d <- if (is.null(dim(x))) length(x) else dim(x)
  writeLines(c(paste(d, collapse = " "), sprintf("%a", as.numeric(x))), file.path(outdir, paste0(name, ".txt")))
}

times <- read.hex("times")
B <- read.hex("B")                       # nt x 1 x P
nt <- length(times); P <- dim(B)[3]

cases <- list(
  list(f = function(x) 1.3 * (0.7 - x), g = function(x) 0.9, x0 = 2.0, p = function(x) x),                        # 0 ou
  list(f = function(x) 1.0 * (1.0 - x), g = function(x) 1.0 * sqrt(abs(x)), x0 = 1.0, p = abs),                   # 1 cir, abs
  list(f = function(x) 0.5 * (0.2 - x), g = function(x) 1.5 * sqrt(abs(x)), x0 = 0.05, p = function(x) pmax(x, 0)), # 2 cir, relu
  list(f = function(x) c(x[2], x[2] * (1.0 - x[1] * x[1]) - x[1]), g = function(x) c(0, 0.5), x0 = c(0.1, -0.2), p = function(x) x),  # 3 vdp
  list(f = function(x) x * (1 - x) - 1.0 * (x * x), g = function(x) 0.5 * x, x0 = 0.3, p = function(x) x),          # 4 logistic
  list(f = function(x) 0.4 * x, g = function(x) 1.0 * x, x0 = 1.0, p = abs),                                       # 5 gbm, abs
  list(f = function(x) x - x * x * x, g = function(x) 0.5, x0 = 0.2, p = function(x) x),                           # 6 doublewell
  list(f = function(x) x - x * x * x, g = function(x) sqrt(1 + x * x), x0 = 0.2, p = function(x) x),               # 7 dwsqrt
  list(f = function(y) (1 - exp(y)) - 0.5 * (0.25 * 0.25), g = function(y) 0.25, x0 = -0.5, p = function(x) x))    # 8 loglogistic

for (k in seq_along(cases)) {
  cs <- cases[[k]]
  nx <- length(cs$x0)
  for (scheme in c("euler", "heun")) {
    fn <- get(scheme)
    X <- array(NA_real_, c(nt, nx, P))
    for (p in 1:P) X[, , p] <- fn(cs$f, cs$g, times, cs$x0, B = B[, 1, p], p = cs$p)$X
    write.hex(sprintf("X_%s_%d", scheme, k - 1), X)
  }
}

## linear system, two noise channels (matrix-valued g, %*% on the BLAS path)
A <- read.hex("lin_A"); G <- read.hex("lin_G"); x0 <- read.hex("lin_x0"); B2 <- read.hex("lin_B")   # B2: nt x nB x P
for (scheme in c("euler", "heun")) {
  fn <- get(scheme)
  X <- array(NA_real_, c(nt, length(x0), P))
  for (p in 1:P) X[, , p] <- fn(function(x) A %*% x, function(x) G, times, x0, B = B2[, , p])$X
  write.hex(sprintf("lin_X_%s", scheme), X)
}

## mortality and stopping (R/sdetools.R:397-416,455-469; vignettes/Killing.Rmd:152-190): OU with
## r = 0.5 exp(2x), and a two-sided stopping rule; Stau supplied per call.  The reference truncates its return at
## tau; results are stored on the full grid with NA after tau (X, S) plus tau itself.
Stau <- read.hex("Stau")
kill <- function(x) 0.5 * exp(2 * x)
stop2 <- function(t, x) (x - (-0.3)) * (1.2 - x)
ou <- list(f = function(x) 1.0 * (0.0 - x), g = function(x) 1.0)
for (scheme in c("euler", "heun")) {
  fn <- get(scheme)
  for (what in c("kill", "stop", "both")) {
    X <- array(NA_real_, c(nt, 1, P)); S <- array(NA_real_, c(nt, P)); tau <- numeric(P)
    for (p in 1:P) {
      sol <- fn(ou$f, ou$g, times, 0.5, B = B[, 1, p],
                h = if (what == "kill") NULL else stop2,
                r = if (what == "stop") function(x) 0 else kill, Stau = Stau[p])
      n <- length(sol$times)
      X[1:n, 1, p] <- sol$X; S[1:n, p] <- sol$S; tau[p] <- sol$tau
    }
    write.hex(sprintf("kill_X_%s_%s", scheme, what), X)
    write.hex(sprintf("kill_S_%s_%s", scheme, what), S)
    write.hex(sprintf("kill_tau_%s_%s", scheme, what), tau)
  }
}

## time-dependent closures (R/sdetools.R:381-394): f(t,x) = f0(x) + c(t), g(t,x) = s(t) g0(x) on the CIR model, and the
## reference's own example f(t,x) = A %*% x + F*sin(2*t) (:211,:363) on the linear inputs
cfun <- function(t) 0.3 * sin(2 * t); sfun <- function(t) 1 + 0.5 * cos(3 * t)
Fin <- x0 * 0 + c(0.5, -1, 0.25)[seq_along(x0)]
for (scheme in c("euler", "heun")) {
  fn <- get(scheme)
  X <- array(NA_real_, c(nt, 1, P))
  for (p in 1:P) X[, 1, p] <- fn(function(t, x) 1.0 * (1.0 - x) + cfun(t), function(t, x) sfun(t) * (1.0 * sqrt(abs(x))), times, 1.0,
                                 B = B[, 1, p], p = abs)$X
  write.hex(sprintf("td_X_%s", scheme), X)
  XL <- array(NA_real_, c(nt, length(x0), P))
  for (p in 1:P) XL[, , p] <- fn(function(t, x) A %*% x + Fin * sin(2 * t), function(x) G, times, x0, B = B2[, , p])$X
  write.hex(sprintf("td_lin_X_%s", scheme), XL)
}

## the linear laws on the same system: lyap (:593-600) and dLinSDE (:508-572) with no input, a constant input
## (zero-order hold) and an input affine in time (first-order hold)
if (requireNamespace("Matrix", quietly = TRUE)) {
  GG <- G %*% t(G)
  write.hex("law_lyap", lyap(A, GG))
  S0 <- diag(seq_along(x0) / 10)
  u1 <- Fin; u2 <- cbind(Fin, 2 * Fin)
  l0 <- dLinSDE(A, G, 0.7); write.hex("law_eAt", as.matrix(l0$eAt)); write.hex("law_St", as.matrix(l0$St))
  write.hex("law_St_S0", as.matrix(dLinSDE(A, G, 0.7, S0 = S0)$St))
  write.hex("law_EX", as.numeric(dLinSDE(A, G, 0.7, x0 = x0)$EX))
  write.hex("law_EX_zoh", as.numeric(dLinSDE(A, G, 0.7, x0 = x0, u = u1)$EX))
  write.hex("law_EX_foh", as.numeric(dLinSDE(A, G, 0.7, x0 = x0, u = u2)$EX))
}

## integrals: columns of int_f / int_g (nc x n in the file, one vector per row)
fcol <- read.hex("int_f"); gcol <- read.hex("int_g")
nc <- dim(fcol)[1]; n <- dim(fcol)[2]
res <- list(l = array(0, c(nc, n)), r = array(0, c(nc, n)), c = array(0, c(nc, n)), ito = array(0, c(nc, n)),
            qv = array(0, c(nc, n)), cv = array(0, c(nc, n)))
for (i in 1:nc) {
  f <- fcol[i, ]; g <- gcol[i, ]
  si <- stochint(f, g, rule = c("l", "r", "c"))
  res$l[i, ] <- si$l; res$r[i, ] <- si$r; res$c[i, ] <- si$c
  res$ito[i, ] <- itointegral(f, g)
  res$qv[i, ] <- QuadraticVariation(g)
  res$cv[i, ] <- CrossVariation(f, g)
}
write.hex("int_stochint_l", res$l); write.hex("int_stochint_r", res$r); write.hex("int_stochint_c", res$c)
write.hex("int_ito", res$ito); write.hex("int_qv", res$qv); write.hex("int_cv", res$cv)

## rBM with recorded normals: rBM(times, sigma, B0, u), first increment spans [0, times[1]] (R/sdetools.R:18)
zz <- read.hex("rbm_z")                  # nt x P
tb <- read.hex("rbm_times")
BM <- array(0, dim(zz))
for (p in 1:dim(zz)[2]) { .z <- zz[, p]; .zpos <- 0; BM[, p] <- rBM(tb, sigma = 0.7, B0 = 0.25, u = -0.1) }
write.hex("rbm_B", BM)

## rvBM (R/sdetools.R:40-46: one rBM per channel through sapply, vector sigma / B0 / u) and rBrownianBridge (:62-73)
.z <- as.numeric(zz); .zpos <- 0         # the columns of rbm_z, consumed channel by channel
write.hex("rvbm_B", rvBM(tb, n = dim(zz)[2], sigma = c(0.7, 1.3, 0.2), B0 = c(0.25, -1, 0), u = c(-0.1, 0, 0.4)))
.z <- zz[, 2]; .zpos <- 0
write.hex("bridge_B", rBrownianBridge(tb, B0T = c(1, -2), sigma = 0.8))

cat("wrote", length(list.files(outdir)), "files to", outdir, "\n")
