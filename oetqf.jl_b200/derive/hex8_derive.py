"""Derivation of the closed-form stress of a uniformly strained cuboid in an elastic half-space
(the quantity `stress_vol_hex8!` of GeoGreensFunctions.jl returns; Barbot et al. 2017), from its definition.

This script is the derivation AND the code generator: it writes
    oetqf.jl_b200/csrc/hex8_gen.cuh         (CUDA device code, included by hex8_dev.cuh)
    oracle/hex8_gen.inc                     (the same expressions as plain C for the CPU checker)
Run from the repo root:  python oetqf.jl_b200/derive/hex8_derive.py

Method ("closed-form derivation plan"):
  u_i(x) = sum_k m_jk * surface integral over the two faces normal to k of G_ij(x, xi) (outward sign),
  with G the Mindlin tensor in Okada's (1992) form.  Every entry of G is a derivative (in the components of
  the source-receiver vector) of three scalar potentials
        P0 = 1/R,   P1 = R,   P3 = R - R3 ln(R + R3)        (P3 is harmonic; lap P1 = 2 P0; lap P0 = 0)
  so a face integral either cancels an in-plane derivative or needs a single antiderivative S_c[P] (or, for P0
  only, a double antiderivative D_ab[P0]); normal-normal second derivatives are removed with the Laplacian
  identities.  With R_i = x_i - xi_i (and R3 = -x3 - xi3 for the image terms) each integration contributes
  [F]_corner with the sign s1*s2*s3, so
        grad u = sum over the 8 corners  s1 s2 s3 * (explicit functions of the corner vector)
  The functions are produced here by symbolic differentiation (all transcendental pieces are symbols with
  hand-coded derivative rules, so expressions stay rational) and common-subexpression elimination.
Every antiderivative and derivative rule is verified numerically below before code is emitted; the emitted
code is validated against an independent quadrature evaluation of the definition by tests/test_oracle_hex8.py.
"""
import itertools
import os
import sys

import mpmath as mp
import sympy as sp

R1, R2, R3, R = sp.symbols("R1 R2 R3 R", real=True)
w1, w2, w3 = sp.symbols("w1 w2 w3", real=True)          # w_c = R + R_c
q1, q2, q3 = sp.symbols("q1 q2 q3", real=True)          # q_c = R^2 - R_c^2
L1, L2, L3 = sp.symbols("L1 L2 L3", real=True)          # ln(R + R_c)
A1, A2, A3 = sp.symbols("A1 A2 A3", real=True)          # atan(R_a R_b / (R_c R))
Ba, Bb = sp.symbols("Ba Bb", real=True)                 # atan(R1/R2), atan(R2/R1)
x3, al = sp.symbols("x3 al", real=True)
iR = sp.Symbol("iR", real=True)                         # 1/R
iw1, iw2, iw3 = sp.symbols("iw1 iw2 iw3", real=True)    # 1/w_c
iq1, iq2, iq3 = sp.symbols("iq1 iq2 iq3", real=True)    # 1/q_c
RS, WS, QS, LS, AS = [R1, R2, R3], [w1, w2, w3], [q1, q2, q3], [L1, L2, L3], [A1, A2, A3]
IWS, IQS = [iw1, iw2, iw3], [iq1, iq2, iq3]
# reciprocals are symbols of their own (with derivative rules), so every generated expression is a
# polynomial: the kernels evaluate 7 reciprocals per corner and then only multiply-add.


def others(c):
    return [i for i in range(3) if i != c]


def rule(sym, i):
    """d sym / d R_i for the non-coordinate symbols"""
    Ri = RS[i]
    if sym == R:
        return Ri * iR
    if sym == iR:
        return -Ri * iR ** 3
    for c in range(3):
        if sym == WS[c]:
            return Ri * iR + (1 if i == c else 0)
        if sym == IWS[c]:
            return -(Ri * iR + (1 if i == c else 0)) * IWS[c] ** 2
        if sym == QS[c]:
            return 0 if i == c else 2 * Ri
        if sym == IQS[c]:
            return 0 if i == c else -2 * Ri * IQS[c] ** 2
        if sym == LS[c]:
            return iR if i == c else Ri * iR * IWS[c]
        if sym == AS[c]:
            a, b = others(c)
            if i == c:
                return -RS[a] * RS[b] * (R ** 2 + RS[c] ** 2) * iR * IQS[a] * IQS[b]
            o = b if i == a else a          # the third axis
            return RS[c] * RS[o] * iR * IQS[o]
    if sym == Ba:
        return {0: R2 * iq3, 1: -R1 * iq3, 2: 0}[i]
    if sym == Bb:
        return {0: -R2 * iq3, 1: R1 * iq3, 2: 0}[i]
    raise KeyError(sym)


AUX = [R, iR] + WS + IWS + QS + IQS + LS + AS + [Ba, Bb]


def Dr(expr, i):
    """total derivative with respect to the coordinate R_i"""
    out = sp.diff(expr, RS[i])
    for s in AUX:
        d = sp.diff(expr, s)
        if d != 0:
            out += d * rule(s, i)
    return out


def Dn(expr, beta):
    for i, n in enumerate(beta):
        for _ in range(n):
            expr = Dr(expr, i)
    return expr


# ---- potentials and their antiderivatives ---------------------------------------------------------
P = {"P0": iR, "P1": R, "P3": R - R3 * L3}
LAPL = {"P0": None, "P1": ("P0", 2), "P3": None}       # lap P1 = 2 P0, others harmonic


def S(c, pot):
    """single antiderivative of the potential along R_c"""
    if pot == "P0":
        return LS[c]
    if pot == "P1":
        return (RS[c] * R + QS[c] * LS[c]) / 2
    if pot == "P2":        # ln(R + R3)
        if c == 2:
            return R3 * L3 - R
        if c == 0:
            return R1 * L3 + R3 * L1 - R1 + R2 * (Ba - A2)
        return R2 * L3 + R3 * L2 - R2 + R1 * (Bb - A1)
    if pot == "P3":
        if c == 2:
            return sp.Rational(3, 4) * R3 * R + (q3 / 4 - R3 ** 2 / 2) * L3
        return S(c, "P1") - R3 * S(c, "P2")
    raise KeyError(pot)


def D0(c):
    """double antiderivative of P0 over the two axes other than c"""
    a, b = others(c)
    return RS[a] * LS[b] + RS[b] * LS[a] - RS[c] * AS[c]


def face_integral(pot, beta, k):
    """antiderivative over the two in-face axes of  d^beta pot  (face normal k), as an expression"""
    beta = list(beta)
    a, b = others(k)
    need = []
    for ax in (a, b):
        if beta[ax] > 0:
            beta[ax] -= 1
        else:
            need.append(ax)
    if len(need) == 0:
        return Dn(P[pot], beta)
    if len(need) == 1:
        return Dn(S(need[0], pot), beta)
    # both in-face axes still need integrating
    if pot == "P0":
        return Dn(D0(k), beta)
    assert beta[k] >= 2, (pot, beta, k)
    # d_k^2 pot = c*P0 - d_a^2 pot - d_b^2 pot
    rest = list(beta)
    rest[k] -= 2
    out = 0
    if LAPL[pot] is not None:
        p2, cst = LAPL[pot]
        out += cst * face_integral(p2, rest, k)
    for ax in (a, b):
        bb = list(rest)
        bb[ax] += 2
        out -= face_integral(pot, bb, k)
    return out


def e(i):
    v = [0, 0, 0]
    v[i] = 1
    return v


def add(*bs):
    return [sum(t) for t in zip(*bs)]


def green_terms(i, j, image):
    """8*pi*mu * G_ij as a list of (coefficient, potential, beta)"""
    dij = 1 if i == j else 0
    t = []
    if not image:
        if dij:
            t.append((2, "P0", [0, 0, 0]))
        t.append((-al, "P1", add(e(i), e(j))))
        return t
    # -uA(R) + uB(R) + x3 uC(R)
    if dij:
        t.append((2, "P0", [0, 0, 0]))
    t.append((al - 2, "P1", add(e(i), e(j))))
    sj = -1 if j == 2 else 1
    t.append((2 * (1 - al) / al * sj, "P3", add(e(i), e(j))))
    si = -1 if i == 2 else 1                                   # (1 - 2 delta_i3)
    if j == 2:
        t.append((-2 * si * (2 - al) * x3, "P0", e(i)))
    if i == 2:
        t.append((+2 * si * (2 - al) * x3, "P0", e(j)))
    t.append((2 * si * al * x3 ** 2, "P0", add(e(i), e(j))))
    t.append((2 * si * al * x3, "P1", add(e(i), e(j), e(2))))
    if i == 2:
        t.append((-2 * si * al * x3, "P0", e(j)))
    if j == 2:
        t.append((-2 * si * al * x3, "P0", e(i)))
    return t


def build(image):
    """F[i][l][j][k] = d/dx_l of the face-k integral of 8*pi*mu*G_ij, per corner"""
    F = {}
    for i, j, k in itertools.product(range(3), repeat=3):
        E = 0
        for coef, pot, beta in green_terms(i, j, image):
            E += coef * face_integral(pot, beta, k)
        for l in range(3):
            if image and l == 2:
                d = -Dr(E, 2) + sp.diff(E, x3)
            else:
                d = Dr(E, l)
            F[(i, l, j, k)] = d
    return F


PAIRS = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]      # xx, xy, xz, yy, yz, zz


def strain_kernels(F):
    """Q[(il),(jk)]: strain component (il) per unit moment component (jk) (symmetrised both ways)"""
    Q = []
    for (i, l) in PAIRS:
        for (j, k) in PAIRS:
            v = (F[(i, l, j, k)] + F[(l, i, j, k)]) / 2
            if j != k:
                v += (F[(i, l, k, j)] + F[(l, i, k, j)]) / 2
            Q.append(v)
    return Q


# ---- numerical verification of every rule ---------------------------------------------------------
def numeric_env(r1, r2, r3):
    r = mp.sqrt(r1 * r1 + r2 * r2 + r3 * r3)
    rs = [r1, r2, r3]
    env = {R1: r1, R2: r2, R3: r3, R: r, iR: 1 / r}
    for c in range(3):
        a, b = others(c)
        env[WS[c]] = r + rs[c]
        env[IWS[c]] = 1 / (r + rs[c])
        env[QS[c]] = r * r - rs[c] * rs[c]
        env[IQS[c]] = 1 / (r * r - rs[c] * rs[c])
        env[LS[c]] = mp.log(r + rs[c])
        env[AS[c]] = mp.atan(rs[a] * rs[b] / (rs[c] * r))
    env[Ba] = mp.atan(r1 / r2)
    env[Bb] = mp.atan(r2 / r1)
    return env


def evalf(expr, r1, r2, r3, extra=None):
    env = numeric_env(mp.mpf(r1), mp.mpf(r2), mp.mpf(r3))
    if extra:
        env.update(extra)
    f = sp.lambdify(list(env.keys()), expr, "mpmath")
    return f(*env.values())


def verify():
    mp.mp.dps = 40
    pts = [(0.7, -1.3, 0.9), (-0.4, 0.8, 1.7), (1.9, 0.6, -0.5), (-1.1, -0.7, -2.2)]
    h = mp.mpf(10) ** -15

    def fd(expr, i, p):
        pp, pm = list(p), list(p)
        pp[i] += h
        pm[i] -= h
        return (evalf(expr, *pp) - evalf(expr, *pm)) / (2 * h)

    worst = 0
    # derivative rules
    for s in AUX:
        for i in range(3):
            for p in pts:
                worst = max(worst, abs(fd(s, i, list(map(mp.mpf, p))) - evalf(rule(s, i), *p)))
    # single antiderivatives: d/dR_c S_c[P] = P
    truth = {"P0": iR, "P1": R, "P2": L3, "P3": R - R3 * L3}
    for pot in truth:
        for c in range(3):
            for p in pts:
                worst = max(worst, abs(evalf(Dr(S(c, pot), c) - truth[pot], *p)))
    # double antiderivative of P0: mixed derivative
    for c in range(3):
        a, b = others(c)
        for p in pts:
            worst = max(worst, abs(evalf(Dr(Dr(D0(c), a), b) - iR, *p)))
    # Laplacians
    for pot, want in (("P0", 0), ("P1", 2 * iR), ("P3", 0)):
        lap = sum(Dr(Dr(P[pot], i), i) for i in range(3))
        for p in pts:
            worst = max(worst, abs(evalf(lap - want, *p)))
    print("max rule/antiderivative residual:", mp.nstr(worst, 5))
    assert worst < mp.mpf(10) ** -12


# ---- code generation -------------------------------------------------------------------------------
def count_ops(exprs):
    return sum(sp.count_ops(e_) for e_ in exprs)


from sympy.printing.c import C99CodePrinter


class MulPrinter(C99CodePrinter):
    def _print_Pow(self, expr):
        b, ex = expr.base, expr.exp
        if ex.is_Integer and 1 < int(ex) <= 6:
            return "(" + "*".join([self._print(b)] * int(ex)) + ")"
        return super()._print_Pow(expr)


_printer = MulPrinter()


def emit(name, exprs, inputs, real_t="double"):
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("t"), optimizations="basic")
    lines = []
    for s, ex in repl:
        lines.append(f"    const {real_t} {s} = {_printer.doprint(ex)};")
    for n, ex in enumerate(red):
        lines.append(f"    q[{n}] += sgn * ({_printer.doprint(ex)});")
    body = "\n".join(lines)
    nops = count_ops([ex for _, ex in repl] + list(red))
    return body, nops


def main():
    verify()
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(os.path.dirname(here))   # <repo>/oetqf.jl_b200/derive -> <repo>
    out = {}
    for image in (False, True):
        F = build(image)
        Q = [sp.together(sp.expand(v)) if False else v for v in strain_kernels(F)]
        body, nops = emit("img" if image else "real", Q, None)
        out[image] = (body, nops)
        print("image" if image else "real", "ops after CSE:", nops, file=sys.stderr)
    header = ("// GENERATED by oetqf.jl_b200/derive/hex8_derive.py -- do not edit.\n"
              "// q[36] += sgn * Q[(il),(jk)] for one corner; (il),(jk) in the order xx,xy,xz,yy,yz,zz.\n"
              "// Inputs: R1,R2,R3 corner vector, R its norm, w_c = R+R_c, q_c = R^2-R_c^2, iR/iw_c/iq_c their reciprocals, L_c = ln(w_c),\n"
              "// A_c = atan(R_a R_b/(R_c R)), Ba = atan(R1/R2), Bb = atan(R2/R1), x3 receiver depth (<= 0), al = alpha.\n")
    for path, qual in ((os.path.join(root, "oracle", "hex8_gen.inc"), "static inline"),
                       (os.path.join(root, "oetqf.jl_b200", "csrc", "hex8_gen.cuh"), "__device__ __forceinline__")):
        with open(path, "w") as fh:
            fh.write(header)
            fh.write(f"{qual} void hex8_corner_real(double R1, double R2, double R3, double R, double w1, double w2, "
                     "double w3, double q1, double q2, double q3, double iR, double iw1, double iw2, double iw3, "
                     "double iq1, double iq2, double iq3, double L1, double L2, double L3, double A1, "
                     "double A2, double A3, double al, double sgn, double* q)\n{\n")
            fh.write(out[False][0])
            fh.write("\n}\n\n")
            fh.write(f"{qual} void hex8_corner_image(double R1, double R2, double R3, double R, double w1, double w2, "
                     "double w3, double q1, double q2, double q3, double iR, double iw1, double iw2, double iw3, "
                     "double iq1, double iq2, double iq3, double L1, double L2, double L3, double A1, "
                     "double A2, double A3, double Ba, double Bb, double x3, double al, double sgn, double* q)\n{\n")
            fh.write(out[True][0])
            fh.write("\n}\n")
        print("wrote", path)


if __name__ == "__main__":
    main()
