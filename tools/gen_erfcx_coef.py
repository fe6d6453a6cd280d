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

# ---- piecewise table (fm::kErfcx32): g(t) = erfcx(2(1-t)/t)/t on 32 equal sub-intervals of (0,1], degree 7
# in the local variable x = 64 t - (2 j + 1); printed coefficient-major: entry k*32 + j.
def g_t(t):
    if t < mp.mpf('1e-30'): return 1/(cpar*mp.sqrt(mp.pi))
    return erfcx(cpar*(1-t)/t)/t
def cheb_fit_ab(f, deg, a, b):
    N = deg+1
    xs = [mp.cos(mp.pi*(k+mp.mpf(1)/2)/N) for k in range(N)]
    fs = [f((a+b)/2 + (b-a)/2*x) for x in xs]
    c = [2*sum(fs[k]*mp.cos(mp.pi*j*(k+mp.mpf(1)/2)/N) for k in range(N))/N for j in range(N)]
    c[0] /= 2
    return c
K8, D8 = 32, 7
Tm = [[mp.mpf(0)]*(D8+1) for _ in range(D8+1)]
Tm[0][0] = mp.mpf(1); Tm[1][1] = mp.mpf(1)
for n in range(2, D8+1):
    for k in range(D8+1):
        Tm[n][k] = (2*Tm[n-1][k-1] if k>0 else 0) - Tm[n-2][k]
table = []
for j in range(K8):
    cj = cheb_fit_ab(g_t, D8, mp.mpf(j)/K8, mp.mpf(j+1)/K8)
    table.append([sum(cj[n]*Tm[n][k] for n in range(D8+1)) for k in range(D8+1)])
print("// kErfcx32: 32 intervals x 8 coefficients, entry k*32 + j")
for k in range(D8+1):
    print("    " + ", ".join(mp.nstr(table[j][k], 20) for j in range(K8)) + ", \\")
# max relative error of the monomial form evaluated in 40-digit arithmetic
worst = 0
for j in range(K8):
    for s in range(201):
        x = mp.mpf(-1) + mp.mpf(2)*s/200
        t = (x + 2*j + 1)/(2*K8)
        p = sum(table[j][k]*x**k for k in range(D8+1))
        worst = max(worst, abs(p/g_t(t) - 1))
print("// max relative truncation error:", mp.nstr(worst, 3))

# ---- log table (fm::kLogT): z in [0.75, 1.5) = 2^-k x; i = ((bits(x) - bits(0.75) + 2^44) >> 45) & 255 picks one of
# 129 centres c_i (c_64 = 1 exactly; spacing 2^-8 below 1, 2^-7 above); entry i = {invc, logc_hi, logc_lo} with
# invc = double(1 / c_i) and logc = -log(invc) for THAT double, split so that logc_hi has 32 trailing zero bits
# (k ln2_hi + logc_hi is then exact).  |r| = |z invc - 1| <= 2^-8 (+ rounding of invc).
import struct
def dbits(x): return struct.unpack("<Q", struct.pack("<d", x))[0]
def fbits(u): return struct.unpack("<d", struct.pack("<Q", u))[0]
rows = []
for i in range(129):
    c = mp.mpf(1) - (64 - i) * mp.mpf(2) ** -8 if i <= 64 else mp.mpf(1) + (i - 64) * mp.mpf(2) ** -7
    invc = float(1 / c)
    logc = -mp.log(mp.mpf(invc))
    hi = fbits(dbits(float(logc)) & ~((1 << 26) - 1)) if logc != 0 else 0.0       # 27 significant bits kept... exactness margin
    lo = float(logc - mp.mpf(hi))
    rows.append((invc, hi, lo))
print("// kLogT: 129 x {invc, logc_hi, logc_lo}")
for i, (a, b, c) in enumerate(rows):
    print("    %.17g, %.17g, %.17g, \\" % (a, b, c))
