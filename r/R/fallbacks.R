## The reference implementations this package falls back to (and is tested against where R exists): the functions of
## SDEtools itself.  As a stand-alone package `sdegpu` masks euler()/heun()/... after library(SDEtools); a maintainer who
## merges r/R/sdegpu.R into SDEtools instead renames the original bodies to these names (INTEGRATION.md section 1).
euler.R <- function(...) SDEtools::euler(...)
heun.R <- function(...) SDEtools::heun(...)
stochint.R <- function(...) SDEtools::stochint(...)
itointegral.R <- function(...) SDEtools::itointegral(...)
QuadraticVariation.R <- function(...) SDEtools::QuadraticVariation(...)
CrossVariation.R <- function(...) SDEtools::CrossVariation(...)
lyap.R <- function(...) SDEtools::lyap(...)
dLinSDE.R <- function(...) SDEtools::dLinSDE(...)
