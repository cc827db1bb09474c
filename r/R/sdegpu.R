## sdegpu.R — R-side binding of libsdegpu.so for the SDEtools sample-path API.
## A maintainer adds this file to R/, src/sdegpu_shim.c to src/, and
##   useDynLib(SDEtools, .registration = TRUE)
## to NAMESPACE.  The existing euler()/heun() bodies (R/sdetools.R:375-470, :240-338) are kept
## verbatim under the names euler.R()/heun.R(): they remain the path for arbitrary closures and the
## oracle the GPU path is tested against.  Signatures below keep every reference argument in place;
## new arguments are appended with defaults, so positional calls keep their meaning.

.sdegpu_models <- c(ou = 0L, cir = 1L, vdp = 2L, logistic = 3L, gbm = 4L, doublewell = 5L, linear = 6L, dwsqrt = 7L, loglogistic = 8L)

.catalogue <- function(model, params, f, g, nx = 1L, nB = 1L) {
  tag <- function(fn) { attr(fn, "sdegpu_model") <- model; attr(fn, "sdegpu_params") <- as.numeric(params)
                        attr(fn, "sdegpu_nx") <- nx; attr(fn, "sdegpu_nB") <- nB; fn }
  list(f = tag(f), g = tag(g))
}

## Catalogue constructors: ordinary closures (usable with the unmodified reference loop) that also carry
## the identity the device needs.  Operation order is the one the kernels reproduce in STRICT arithmetic.
sde.ou <- function(lambda = 1, xi = 0, sigma = 1)
  .catalogue("ou", c(lambda, xi, sigma), function(x) lambda * (xi - x), function(x) sigma)
sde.cir <- function(lambda = 1, xi = 1, gamma = 1)
  .catalogue("cir", c(lambda, xi, gamma), function(x) lambda * (xi - x), function(x) gamma * sqrt(abs(x)))
sde.vdp <- function(mu = 1, sigma = 0.5)
  .catalogue("vdp", c(mu, sigma), function(x) c(x[2], x[2] * (mu - x[1] * x[1]) - x[1]), function(x) c(0, sigma), nx = 2L)
sde.logistic <- function(sigma = 0.5, harvest = 0)
  .catalogue("logistic", c(sigma, harvest), function(x) x * (1 - x) - harvest * (x * x), function(x) sigma * x)
sde.gbm <- function(mu = 0, sigma = 1)
  .catalogue("gbm", c(mu, sigma), function(x) mu * x, function(x) sigma * x)
sde.doublewell <- function(sigma = 0.5)
  .catalogue("doublewell", sigma, function(x) x - x * x * x, function(x) sigma)
sde.dwsqrt <- function()                                      # vignettes/ItoIntegrals.Rmd:66-69
  .catalogue("dwsqrt", numeric(0), function(x) x - x * x * x, function(x) sqrt(1 + x * x))
sde.loglogistic <- function(sigma = 0.25)                     # vignettes/ItosFormula.Rmd:113-116
  .catalogue("loglogistic", sigma, function(y) (1 - exp(y)) - 0.5 * (sigma * sigma), function(y) sigma)
sde.linear <- function(A, G) {
  A <- as.matrix(A); G <- matrix(G, nrow = nrow(A))
  .catalogue("linear", c(as.numeric(A), as.numeric(G)), function(x) A %*% x, function(x) G, nx = nrow(A), nB = ncol(G))
}

## Killing rates r(x) and stopping functions h(t,x) (R/sdetools.R:397-416,455-457; Killing.Rmd:152-190):
## ordinary closures again, tagged so the device can evaluate them.
.tag_rule <- function(fn, what, kind, ...) { attr(fn, what) <- c(list(kind = kind), list(...)); fn }
sde.kill.const <- function(a) .tag_rule(function(x) a, "sdegpu_kill", 1L, a = a, b = 0, comp = 0L)
sde.kill.exp <- function(a, b, comp = 1L) .tag_rule(function(x) a * exp(b * x[comp]), "sdegpu_kill", 2L, a = a, b = b, comp = comp - 1L)
sde.kill.quad <- function(a, b = 0, comp = 1L) .tag_rule(function(x) a * ((x[comp] - b) * (x[comp] - b)), "sdegpu_kill", 3L, a = a, b = b, comp = comp - 1L)
sde.stop.below <- function(lo, comp = 1L) .tag_rule(function(t, x) x[comp] - lo, "sdegpu_stop", 1L, lo = lo, hi = 0, comp = comp - 1L)
sde.stop.above <- function(hi, comp = 1L) .tag_rule(function(t, x) hi - x[comp], "sdegpu_stop", 2L, lo = 0, hi = hi, comp = comp - 1L)
sde.stop.outside <- function(lo, hi, comp = 1L) .tag_rule(function(t, x) (x[comp] - lo) * (hi - x[comp]), "sdegpu_stop", 3L, lo = lo, hi = hi, comp = comp - 1L)

.projection_id <- function(p) {
  if (is.null(p)) return(0L)
  if (identical(p, abs)) return(1L)
  if (is.character(p)) return(match(p, c("identity", "abs", "relu")) - 1L)
  if (isTRUE(all.equal(body(p), quote(x))) ) return(0L)
  NA_integer_
}

.sdegpu_dispatch <- function(scheme, fallback, f, g, times, x0, B, p, h, r, S0, Stau,
                             npaths, seed, output, moments, center, hist, arith, integrals = FALSE) {
  proj <- .projection_id(p)
  on.device <- !is.null(attr(f, "sdegpu_model")) && identical(attr(f, "sdegpu_model"), attr(g, "sdegpu_model")) &&
    identical(attr(f, "sdegpu_params"), attr(g, "sdegpu_params")) && !is.na(proj) &&
    (is.null(h) || !is.null(attr(h, "sdegpu_stop"))) && (is.null(r) || !is.null(attr(r, "sdegpu_kill")))
  if (!on.device) return(fallback(f, g, times, x0, B, p = if (is.null(p)) function(x) x else p, h = h, r = r, S0 = S0, Stau = Stau))
  force(Stau)                       # the reference consumes one runif() per call (lazy default forced at :457/:327)
  nx <- attr(f, "sdegpu_nx"); nB <- attr(f, "sdegpu_nB"); nt <- length(times)
  if (!is.null(B)) {
    if (!is.array(B) || length(dim(B)) < 3) B <- array(matrix(B, nrow = nt, ncol = nB), c(nt, nB, 1L))   # :435
    npaths <- dim(B)[3]
  } else if (is.null(npaths)) npaths <- if (is.matrix(x0)) ncol(x0) else 1L
  if (integrals && output == "path") output <- "terminal"    # integrals accumulate in the step loop, no path is stored
  opts <- list(npaths = npaths, nx = nx, nB = nB, projection = proj, output = match(output, c("none", "terminal", "path")) - 1L,
               moments = as.numeric(moments), center = center, hist = hist, seed = seed, integrals = as.numeric(integrals))
  if (!is.null(arith)) opts$arith <- match(arith, c("strict", "fast")) - 1L
  if (!is.null(r)) { k <- attr(r, "sdegpu_kill"); opts <- c(opts, list(kill_kind = k$kind, kill_comp = k$comp, kill_a = k$a, kill_b = k$b)) }
  if (!is.null(h)) { k <- attr(h, "sdegpu_stop"); opts <- c(opts, list(stop_kind = k$kind, stop_comp = k$comp, stop_lo = k$lo, stop_hi = k$hi)) }
  if (!is.null(r) || !is.null(h)) opts <- c(opts, list(S0 = S0, Stau = rep_len(as.numeric(Stau), npaths)))
  res <- .Call(C_sdegpu_run, scheme, .sdegpu_models[[attr(f, "sdegpu_model")]], attr(f, "sdegpu_params"),
               as.numeric(times), as.numeric(x0), B, opts[!vapply(opts, is.null, TRUE)])
  X <- res$X
  if (!is.null(X)) {
    if (npaths == 1L) { X <- matrix(X, nrow = nt, ncol = nx); colnames(X) <- names(x0) }       # list(times=, X=)  :464
    else dimnames(X) <- list(NULL, names(x0), NULL)
  }
  if (!is.null(r) && npaths == 1L) {                          # the reference truncates at tau with mortality (:466-468)
    k <- which(times == res$tau)[1]
    return(list(times = times[1:k], X = if (nx == 1L) X[1:k, 1] else X[1:k, , drop = FALSE], S = res$S, tau = res$tau))
  }
  if (!is.null(res$integrals)) dimnames(res$integrals) <- list(c("QV", "ito", "strat", "time"), names(x0), NULL)
  extra <- res[c("XT", "power_sums", "hist", "seed", "tau", "S", "integrals")]
  c(list(times = times, X = X), extra[!vapply(extra, is.null, TRUE)])
}

euler <- function(f, g, times, x0, B = NULL, p = function(x) x, h = NULL, r = NULL, S0 = 1, Stau = runif(1),
                  npaths = NULL, seed = NULL, output = "path", moments = FALSE, center = NULL, hist = NULL, arith = NULL,
                  integrals = FALSE)
  .sdegpu_dispatch(0L, euler.R, f, g, times, x0, B, p, h, r, S0, Stau, npaths, seed, output, moments, center, hist, arith, integrals)

heun <- function(f, g, times, x0, B = NULL, p = function(x) x, h = NULL, r = NULL, S0 = 1, Stau = runif(1),
                 npaths = NULL, seed = NULL, output = "path", moments = FALSE, center = NULL, hist = NULL, arith = NULL,
                 integrals = FALSE)
  .sdegpu_dispatch(1L, heun.R, f, g, times, x0, B, p, h, r, S0, Stau, npaths, seed, output, moments, center, hist, arith, integrals)

## Integrals: same results bit for bit (the device emulates cumsum's 80-bit accumulator); matrices are
## integrated column by column in one launch.
stochint <- function(f, g, rule = "l") {
  want <- c("l", "r", "c") %in% rule
  res <- .Call(C_sdegpu_stochint, as.numeric(f), as.numeric(g), as.integer(want))
  names(res) <- c("l", "r", "c")
  if (length(rule) == 1) return(res[[rule]])
  data.frame(res[c("l", "r", "c")[want]])[, rule]                                                  # :112
}
itointegral <- function(G, B) .Call(C_sdegpu_itointegral, G + 0, B + 0)
QuadraticVariation <- function(X) .Call(C_sdegpu_qv, X + 0)
CrossVariation <- function(X, Y) .Call(C_sdegpu_crossvar, X + 0, Y + 0)

rvBM.gpu <- function(times, n = 1, sigma = 1, B0 = 0, u = 0, seed = NULL)
  .Call(C_sdegpu_rvbm, as.numeric(times), as.integer(n), sigma, B0, u, if (is.null(seed)) list() else list(seed = seed))

## max / min / which.max / which.min / first passage of every column of rvBM.gpu(times, n, ...) under the same
## seed, without building the matrix (vignettes/BrownianMotion.Rmd:101-144: apply(rvBM(0:Nt,Np),2,max)).
rvBM.functionals.gpu <- function(times, n = 1, sigma = 1, B0 = 0, u = 0, level = NULL, seed = NULL)
  .Call(C_sdegpu_bm_functionals, as.numeric(times), as.integer(n), sigma, B0, u,
        c(if (is.null(level)) list() else list(level = level), if (is.null(seed)) list() else list(seed = seed)))

## Exact transition-law samplers on the device.  Like rvBM.gpu they keep their own names: the draws come
## from Philox, not from R's stream, so they are the same law as rOU/rCIR (R/sdetools.R:784-791, :695-702)
## but not the same numbers.  steps > 1: X_{k+1} ~ law(. | X_k, t), the vignette's recursive sampling.
rOU.gpu <- function(n, x0, lambda, xi, sigma, t, steps = 1, seed = NULL)
  .Call(C_sdegpu_rlaw, 0L, as.integer(n), as.numeric(x0), c(lambda, xi, sigma), t,
        c(list(steps = steps), if (is.null(seed)) list() else list(seed = seed)))

rCIR.gpu <- function(n, x0, lambda, xi, gamma, t, Stratonovich = FALSE, steps = 1, seed = NULL)
  .Call(C_sdegpu_rlaw, 1L, as.integer(n), as.numeric(x0), c(lambda, xi, gamma), t,
        c(list(steps = steps, stratonovich = as.numeric(Stratonovich)), if (is.null(seed)) list() else list(seed = seed)))

## The laws themselves on the device (lower tail, natural scale); the reference's dOU/pOU/dCIR/pCIR stay as they are.
pOU.gpu <- function(x, x0, lambda, xi, sigma, t) .Call(C_sdegpu_law_eval, 0L, 1L, as.numeric(x), as.numeric(x0), c(lambda, xi, sigma), t, 0L)
dOU.gpu <- function(x, x0, lambda, xi, sigma, t) .Call(C_sdegpu_law_eval, 0L, 0L, as.numeric(x), as.numeric(x0), c(lambda, xi, sigma), t, 0L)
pCIR.gpu <- function(x, x0, lambda, xi, gamma, t, Stratonovich = FALSE)
  .Call(C_sdegpu_law_eval, 1L, 1L, as.numeric(x), as.numeric(x0), c(lambda, xi, gamma), t, as.integer(Stratonovich))
dCIR.gpu <- function(x, x0, lambda, xi, gamma, t, Stratonovich = FALSE)
  .Call(C_sdegpu_law_eval, 1L, 0L, as.numeric(x), as.numeric(x0), c(lambda, xi, gamma), t, as.integer(Stratonovich))
