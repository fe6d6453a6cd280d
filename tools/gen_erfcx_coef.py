"""Generates the erfcx polynomial of include/fastad_b200/fastmath.cuh (mpmath, 40 digits)."""
import mpmath as mp
import mpmath as mp, numpy as np
mp.mp.dps = 40
def erfcx(w):
    w = mp.mpf(w)
    if w < 25:
        return mp.exp(w*w)*mp.erfc(w)
    # asymptotic for large w
    s = mp.mpf(1); term = mp.mpf(1); 
    for k in range(1, 30):
        term *= -(2*k-1)/(2*w*w); s += term
    return s/(w*mp.sqrt(mp.pi))
def cheb_fit(f, deg):
    # chebyshev interpolation on [-1,1]
    N = deg+1
    xs = [mp.cos(mp.pi*(k+mp.mpf(1)/2)/N) for k in range(N)]
    fs = [f(x) for x in xs]
    c = []
    for j in range(N):
        s = mp.mpf(0)
        for k in range(N):
            s += fs[k]*mp.cos(mp.pi*j*(k+mp.mpf(1)/2)/N)
        c.append(2*s/N)
    c[0] /= 2
    return c
def cheb_eval(c, x):
    b1=b2=mp.mpf(0)
    for cj in reversed(c[1:]):
        b1, b2 = 2*x*b1 - b2 + cj, b1
    return x*b1 - b2 + c[0]

cpar = 2.0
def g(x):
    t = (x+1)/2
    if t < mp.mpf('1e-30'): return 1/(cpar*mp.sqrt(mp.pi))
    w = cpar*(1-t)/t
    return erfcx(w)/t
deg = 27
c = cheb_fit(g, deg)
T = [[mp.mpf(0)]*(deg+1) for _ in range(deg+1)]
T[0][0] = mp.mpf(1); T[1][1] = mp.mpf(1)
for n in range(2, deg+1):
    for k in range(deg+1):
        T[n][k] = (2*T[n-1][k-1] if k>0 else 0) - T[n-2][k]
mono = [sum(c[n]*T[n][k] for n in range(deg+1)) for k in range(deg+1)]
print("// erfcx(w) * (w+2)/2 as a degree-27 polynomial in x = 2t-1, t = 2/(w+2); Chebyshev interpolant")
print("// computed with mpmath at 40 digits (generator: tools/gen_erfcx_coef.py)")
for k,m in enumerate(mono):
    print("    %s,%s" % (mp.nstr(m, 20), "" if k%1 else ""))
