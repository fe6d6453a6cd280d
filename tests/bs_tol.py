"""Tolerance model of the batched Black-Scholes outputs (K1) against the oracle, used by the GPU tests and bench.py's
parity gate.  Every output is a sum of products of a normal CDF / density with S or PV, so its rounding error scales with
the MAGNITUDE OF ITS TERMS (the condition number of the subtraction), not with the result:

  call = Phi1 S - Phi2 PV            scale = eps (|Phi1 S| + |Phi2 PV| + (S + PV)/4)     [ + : absolute error of Phi itself]
  delta = Phi1 + g / (S sst) ...     the adjoint through the placeholders is the RESIDUAL g = S phi(d1) - PV phi(d2), 0 in
                                     exact arithmetic (example/black_scholes.cpp:16-39 differentiates through it anyway):
                                     rounding noise of size eps S phi(d1), amplified by 1/sst on its way to S, sigma and r.
  vega  = PV phi(d2) sqrt(tau) ...   phi(d) = exp(-d^2/2) amplifies the rounding of d (itself (|log S/K| + |drift|)/sst ulps
                                     after the division by a small sst) by |d|: for |d| ~ 5 and sst ~ 0.02 two correct
                                     implementations differ by ~1e3 ulp of vega -- conditioning, not inaccuracy.

`scales()` returns eps x those magnitudes per output; an implementation passes when |gpu - oracle| <= C x scale with a
small constant C (see test_black_scholes_ulp_report_full_size, which also prints the measured maxima).  This replaces round
1's flat 1e-13 max(S, K) floor (about 2000 ulp of an at-the-money price)."""
import numpy as np
from scipy.special import erfc

EPS = np.finfo(np.float64).eps
NAMES = ["call", "put", "dC/dS", "dC/dsig", "dC/dr", "dP/dS", "dP/dsig", "dP/dr"]


def scales(inp):
    S, K, sig, tau, r = (np.asarray(inp[k], dtype=np.float64) for k in ("S", "K", "sigma", "tau", "r"))
    sq = np.sqrt(tau)
    sst = sig * sq
    c0 = np.log(S / K)
    drift = (r + 0.5 * sig * sig) * tau
    d1 = (c0 + drift) / sst
    d2 = d1 - sst
    Phi = lambda d: 0.5 * erfc(-d / np.sqrt(2.0))
    phi1 = np.exp(-0.5 * d1 * d1) / np.sqrt(2 * np.pi)
    PV = K * np.exp(-r * tau)
    # rounding of d itself, in units of eps: log and drift are each rounded, then divided by the (small) sst
    dd = (np.abs(c0) + np.abs(drift)) / sst + np.abs(d1) + np.abs(d2) + 1.0
    # phi(d) = exp(-d^2/2)/sqrt(2 pi) amplifies it: relative error |d| dd.  S phi(d1) == PV phi(d2) (the BS identity)
    amp = 1.0 + (np.abs(d1) + np.abs(d2)) * dd
    A = S * phi1
    noise = A * amp / sst                               # the residual S phi(d1) - PV phi(d2) on its way out, per unit eps
    price_abs = (S + PV) / 4                            # absolute error of Phi itself (0.5 (erf + 1) on the oracle side)
    # prices: the error of d shared by d1 and d2 cancels (d call / d d1 = 0); what is left is the rounding of d2 = d1 - sst
    call = Phi(d1) * S + Phi(d2) * PV + price_abs + A * (np.abs(d2) + 1.0)
    put = Phi(-d1) * S + Phi(-d2) * PV + price_abs + A * (np.abs(d2) + 1.0)
    delta_c = Phi(d1) + 0.25 + phi1 * dd + noise / S
    delta_p = Phi(-d1) + 0.25 + phi1 * dd + noise / S
    vega = A * sq * amp + noise * sq * (1.0 + np.abs(d1)) + noise * tau * sig
    rho_c = Phi(d2) * tau * PV + tau * PV / 4 + A * tau * dd + noise * tau
    rho_p = Phi(-d2) * tau * PV + tau * PV / 4 + A * tau * dd + noise * tau
    return [np.maximum(EPS * v, 1e-300) for v in (call, put, delta_c, vega, rho_c, delta_p, vega, rho_p)]


def report(inp, got, ref):
    """max |got - ref| / scale per output, and the max error of the two prices in ulps of the price for near-the-money options."""
    sc = scales(inp)
    ratios = [float(np.max(np.abs(g - o) / s)) for g, o, s in zip(got, ref, sc)]
    S, K = np.asarray(inp["S"]), np.asarray(inp["K"])
    atm = np.abs(S / K - 1.0) < 0.05
    atm_ulps = [float(np.max(np.abs(got[k][atm] - ref[k][atm]) / np.spacing(np.abs(ref[k][atm])))) if atm.any() else 0.0 for k in (0, 1)]
    return ratios, atm_ulps
