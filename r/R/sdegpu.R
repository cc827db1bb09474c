## sdegpu.R — R-side binding of libsdegpu.so for the SDEtools sample-path API.
## A maintainer adds this file to R/, src/sdegpu_shim.c to src/, and
##   useDynLib(SDEtools, .registration = TRUE)
## to NAMESPACE.  The existing euler()/heun() bodies (R/sdetools.R:375-470, :240-338) are kept
## verbatim under the names euler.R()/heun.R(): they remain the path for arbitrary closures and the
## oracle the GPU path is tested against (likewise stochint.R, itointegral.R, QuadraticVariation.R, CrossVariation.R,
## lyap.R, dLinSDE.R).  Signatures below keep every reference argument in place;
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

## Time-dependent closures (R/sdetools.R:381-394 accept f(t,x), g(t,x)): f(t,x) = f0(x) + c(t) and g(t,x) = s(t) * g0(x)
## with catalogue f0, g0.  The reference evaluates them at grid times only, so the wrapper tabulates c and s at `times`
## (with R's own arithmetic, hence the same doubles the closure would produce) and the device adds / scales per step.
## The reference's example  f <- function(t,x) A %*% x + F * sin(2*t)  (:211,:363) is sde.forced(sde.linear(A,G)$f, function(t) F*sin(2*t)).
sde.forced <- function(f, c) {
  f0 <- f; out <- function(t, x) f0(x) + c(t)
  attributes(out) <- attributes(f0); attr(out, "sdegpu_forcing") <- c; out
}
sde.scaled <- function(g, s) {
  g0 <- g; out <- function(t, x) s(t) * g0(x)
  attributes(out) <- attributes(g0); attr(out, "sdegpu_scale") <- s; out
}

.sdegpu_dispatch <- function(scheme, fallback, f, g, times, x0, B, p, h, r, S0, Stau, Stau.missing,
                             npaths, seed, output, moments, center, hist, arith, integrals = FALSE,
                             model_integrals = FALSE, devices = NULL) {
  proj <- .projection_id(p)
  on.device <- !is.null(attr(f, "sdegpu_model")) && identical(attr(f, "sdegpu_model"), attr(g, "sdegpu_model")) &&
    identical(attr(f, "sdegpu_params"), attr(g, "sdegpu_params")) && !is.na(proj) &&
    (is.null(h) || !is.null(attr(h, "sdegpu_stop"))) && (is.null(r) || !is.null(attr(r, "sdegpu_kill")))
  if (!on.device) return(fallback(f, g, times, x0, B, p = if (is.null(p)) function(x) x else p, h = h, r = r, S0 = S0, Stau = Stau))
  force(Stau)                       # the reference consumes one runif() per call (lazy default forced at :457/:327)
  nx <- attr(f, "sdegpu_nx"); nB <- attr(f, "sdegpu_nB"); nt <- length(times)
  if (!is.null(B)) {
    if (!is.array(B) || length(dim(B)) < 3) B <- array(matrix(B, nrow = nt, ncol = nB), c(nt, nB, 1L))   # :435
    if (dim(B)[1] != nt || dim(B)[2] != nB) stop("B must be length(times) x nB (x npaths)")
    if (!is.null(npaths) && npaths != dim(B)[3]) stop("npaths disagrees with the third dimension of B")
    npaths <- dim(B)[3]
    storage.mode(B) <- "double"
  } else if (is.null(npaths)) npaths <- if (is.matrix(x0)) ncol(x0) else 1L
  if (length(x0) != nx && length(x0) != nx * npaths) stop("x0 must have length nx, or be nx x npaths")
  if ((integrals || model_integrals) && output == "path") output <- "terminal"    # accumulated in the step loop, no path is stored
  killed <- !is.null(r) || !is.null(h)
  opts <- list(npaths = npaths, nx = nx, nB = nB, projection = proj, output = match(output, c("none", "terminal", "path")) - 1L,
               moments = as.numeric(moments), center = if (is.null(center)) NULL else as.numeric(center),
               hist = if (is.null(hist)) NULL else as.numeric(hist), seed = seed, integrals = as.numeric(integrals),
               model_integrals = as.numeric(model_integrals), devices = devices)
  if (!is.null(arith)) opts$arith <- match(arith, c("strict", "fast")) - 1L
  if (!is.null(cf <- attr(f, "sdegpu_forcing")))        # c(times[i]) as the reference's loop would evaluate it: nt x nx
    opts$forcing <- as.numeric(t(vapply(times, function(ti) rep_len(as.numeric(cf(ti)), nx), numeric(nx))))
  if (!is.null(sf <- attr(g, "sdegpu_scale"))) opts$gscale <- vapply(times, function(ti) as.numeric(sf(ti))[1], 0)
  if (!is.null(r)) { k <- attr(r, "sdegpu_kill"); opts <- c(opts, list(kill_kind = k$kind, kill_comp = k$comp, kill_a = k$a, kill_b = k$b)) }
  if (!is.null(h)) { k <- attr(h, "sdegpu_stop"); opts <- c(opts, list(stop_kind = k$kind, stop_comp = k$comp, stop_lo = k$lo, stop_hi = k$hi)) }
  if (killed) {
    opts$S0 <- S0
    opts$Spath <- as.numeric(!is.null(r) && output == "path")
    ## The reference draws one runif(1) per euler()/heun() call, i.e. one threshold per path of an sapply ensemble
    ## (Killing.Rmd:183-189).  With the default Stau the device draws one Philox uniform per path; a user-supplied
    ## Stau is one number for one path or one per path, never recycled across an ensemble.
    if (!Stau.missing || npaths == 1L) {
      if (length(Stau) != npaths) stop("Stau must have one threshold per path (length npaths)")
      opts$Stau <- as.numeric(Stau)
    }
  }
  res <- .Call(C_sdegpu_run, scheme, .sdegpu_models[[attr(f, "sdegpu_model")]], attr(f, "sdegpu_params"),
               as.numeric(times), as.numeric(x0), B, opts[!vapply(opts, is.null, TRUE)])
  X <- res$X
  if (!is.null(X)) {
    if (npaths == 1L) { X <- matrix(X, nrow = nt, ncol = nx); colnames(X) <- names(x0) }       # list(times=, X=)  :464
    else dimnames(X) <- list(NULL, names(x0), NULL)
  }
  if (!is.null(r) && npaths == 1L && !is.null(X)) {           # the reference truncates at tau with mortality (:466-468)
    k <- which(times == res$tau)[1]
    return(list(times = times[1:k], X = if (nx == 1L) X[1:k, 1] else X[1:k, , drop = FALSE], S = res$Spath[1:k, 1], tau = res$tau))
  }
  if (!is.null(res$integrals)) dimnames(res$integrals) <- list(c("QV", "ito", "strat", "time"), names(x0), NULL)
  if (!is.null(res$model_integrals)) dimnames(res$model_integrals) <- list(c("drift", "noise"), names(x0), NULL)
  extra <- res[c("XT", "power_sums", "hist", "seed", "tau", "S", "Spath", "integrals", "model_integrals")]
  c(list(times = times, X = X), extra[!vapply(extra, is.null, TRUE)])
}

euler <- function(f, g, times, x0, B = NULL, p = function(x) x, h = NULL, r = NULL, S0 = 1, Stau = runif(1),
                  npaths = NULL, seed = NULL, output = "path", moments = FALSE, center = NULL, hist = NULL, arith = NULL,
                  integrals = FALSE, model_integrals = FALSE, devices = NULL)
  .sdegpu_dispatch(0L, euler.R, f, g, times, x0, B, p, h, r, S0, Stau, missing(Stau), npaths, seed, output, moments, center, hist,
                   arith, integrals, model_integrals, devices)

heun <- function(f, g, times, x0, B = NULL, p = function(x) x, h = NULL, r = NULL, S0 = 1, Stau = runif(1),
                 npaths = NULL, seed = NULL, output = "path", moments = FALSE, center = NULL, hist = NULL, arith = NULL,
                 integrals = FALSE, model_integrals = FALSE, devices = NULL)
  .sdegpu_dispatch(1L, heun.R, f, g, times, x0, B, p, h, r, S0, Stau, missing(Stau), npaths, seed, output, moments, center, hist,
                   arith, integrals, model_integrals, devices)

## Integrals: same results bit for bit (the device emulates cumsum's 80-bit accumulator); matrices are integrated column
## by column in one launch.  A single vector — the reference's own use, ~1e2..2e4 points — is one dependent chain that R's
## vectorised cumsum finishes before a host-device round trip would: below .sdegpu_integral_min elements the original
## definitions (kept as stochint.R, itointegral.R, ... : R/sdetools.R:105-113,134-137,155-158,184-187) run unchanged.
.sdegpu_integral_min <- 2^20
stochint <- function(f, g, rule = "l") {
  if (length(f) < .sdegpu_integral_min) return(stochint.R(f, g, rule))
  want <- c("l", "r", "c") %in% rule
  res <- .Call(C_sdegpu_stochint, f + 0, g + 0, as.integer(want))
  names(res) <- c("l", "r", "c")
  if (length(rule) == 1) return(res[[rule]])
  data.frame(res[c("l", "r", "c")[want]])[, rule]                                                  # :112
}
itointegral <- function(G, B) if (length(G) < .sdegpu_integral_min) itointegral.R(G, B) else .Call(C_sdegpu_itointegral, G + 0, B + 0)
QuadraticVariation <- function(X) if (length(X) < .sdegpu_integral_min) QuadraticVariation.R(X) else .Call(C_sdegpu_qv, X + 0)
CrossVariation <- function(X, Y) if (length(X) < .sdegpu_integral_min) CrossVariation.R(X, Y) else .Call(C_sdegpu_crossvar, X + 0, Y + 0)

## lyap() and dLinSDE() beyond the reach of the reference's O(d^6) Kronecker solve (R/sdetools.R:593-600, :508-572): same
## arguments, same return lists.  Small systems keep the reference's code (lyap.R, dLinSDE.R).
.sdegpu_lyap_min <- 25L
lyap <- function(A, Q) {
  A <- as.matrix(A)
  if (nrow(A) < .sdegpu_lyap_min) return(lyap.R(A, Q))
  .Call(C_sdegpu_lyap, A + 0, as.matrix(Q) + 0)
}
dLinSDE <- function(A, G, t, x0 = NULL, u = NULL, S0 = 0 * A) {
  A <- as.matrix(A); G <- as.matrix(G)
  if (nrow(A) < .sdegpu_lyap_min) return(dLinSDE.R(A, G, t, x0, u, S0))
  .Call(C_sdegpu_dlinsde, A + 0, G + 0, as.numeric(t), if (is.null(x0)) NULL else as.numeric(x0),
        if (is.null(u)) NULL else as.numeric(u), as.matrix(S0) + 0)
}

sdegpu.info <- function() .Call(C_sdegpu_info)

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
