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
x3, al, ial = sp.symbols("x3 al ial", real=True)          # ial = 1/al
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


def TD(pot, gamma):
    """d^gamma of the TRIPLE antiderivative T[pot] (gamma_c = 0 means "still integrated along c")."""
    gamma = list(gamma)
    missing = [c for c in range(3) if gamma[c] == 0]
    if len(missing) == 0:
        return Dn(P[pot], [g - 1 for g in gamma])
    if len(missing) == 1:
        c = missing[0]
        return Dn(S(c, pot), [g - 1 if i != c else 0 for i, g in enumerate(gamma)])
    assert len(missing) == 2, (pot, gamma)
    k = [c for c in range(3) if gamma[c] > 0][0]
    a, b = others(k)
    if pot == "P0":
        beta = [0, 0, 0]
        beta[k] = gamma[k] - 1
        return Dn(D0(k), beta)
    # d_k^2 T[pot] = c T[P0] - d_a^2 T[pot] - d_b^2 T[pot]   (lap P1 = 2 P0; P3 harmonic)
    assert gamma[k] >= 3, (pot, gamma)
    rest = [0, 0, 0]
    rest[k] = gamma[k] - 2
    out = 0
    if LAPL[pot] is not None:
        p2, cst = LAPL[pot]
        out += cst * TD(p2, rest)
    for ax in (a, b):
        g2 = list(rest)
        g2[ax] = 2
        out -= TD(pot, g2)
    return out


def e(i):
    v = [0, 0, 0]
    v[i] = 1
    return v


def add(*bs):
    return [sum(t) for t in zip(*bs)]


def green_terms(i, j, image):
    """8*pi*mu * G_ij as a list of (coefficient, potential, beta): coefficient * d^beta potential"""
    dij = 1 if i == j else 0
    t = []
    if not image:
        if dij:
            t.append((sp.Integer(2), "P0", [0, 0, 0]))
        t.append((-al, "P1", add(e(i), e(j))))
        return t
    # -uA(R) + uB(R) + x3 uC(R)
    if dij:
        t.append((sp.Integer(2), "P0", [0, 0, 0]))
    t.append((al - 2, "P1", add(e(i), e(j))))
    sj = -1 if j == 2 else 1
    t.append((2 * (1 - al) * ial * sj, "P3", add(e(i), e(j))))
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


def F_combo(i, l, j, k, image):
    """d/dx_l of the face-k integral of 8*pi*mu*G_ij as a linear combination {(pot, gamma): coefficient} of
    derivatives of triple antiderivatives, evaluated per corner"""
    combo = {}

    def put(key, c):
        combo[key] = combo.get(key, 0) + c

    for coef, pot, beta in green_terms(i, j, image):
        gam = add(beta, e(k))
        if not image or l != 2:
            put((pot, tuple(add(gam, e(l)))), coef)
        else:
            # image source: R3 = -x3 - xi3, so d/dx3 = -d/dR3 + explicit x3-derivative of the coefficient
            put((pot, tuple(add(gam, e(2)))), -coef)
            dc = sp.diff(coef, x3)
            if dc != 0:
                put((pot, tuple(gam)), dc)
    return combo


PAIRS = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]      # xx, xy, xz, yy, yz, zz


def strain_combos(image):
    """Q[(il),(jk)] as linear combinations of basis sums (symmetrised in (il) and in (jk))"""
    out = []
    for (i, l) in PAIRS:
        for (j, k) in PAIRS:
            tot = {}
            parts = [(F_combo(i, l, j, k, image), sp.Rational(1, 2)), (F_combo(l, i, j, k, image), sp.Rational(1, 2))]
            if j != k:
                parts += [(F_combo(i, l, k, j, image), sp.Rational(1, 2)), (F_combo(l, i, k, j, image), sp.Rational(1, 2))]
            for combo, w in parts:
                for key, c in combo.items():
                    tot[key] = tot.get(key, 0) + w * c
            out.append({k_: sp.expand(v) for k_, v in tot.items() if sp.expand(v) != 0})
    return out


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


def cse_block(assigns, prefix):
    """assigns: list of (target string, expr, row mask) -> C lines with common subexpressions hoisted; each
    accumulation is guarded by HEX8_NEED(mask) (rows of Q the function feeds) so that a caller needing only some
    strain components compiles the rest away"""
    repl, red = sp.cse([ex for _, ex, _ in assigns], symbols=sp.numbered_symbols(prefix), optimizations="basic")
    lines = [f"        const double {s_} = {_printer.doprint(ex)};" for s_, ex in repl]
    # The running sums live in shared memory behind volatile accesses.  Accumulating one function at a time
    # (load, add, store, load, ...) exposes the full shared-memory latency once per function; batches of
    # HEX8_BATCH functions issue their loads back to back and store after the arithmetic.
    batch = int(os.environ.get("HEX8_BATCH", "6"))
    todo = list(zip(assigns, red))
    for b0 in range(0, len(todo), batch):
        part = todo[b0:b0 + batch]
        if batch == 1:
            (tgt, _, mask), ex = part[0]
            lines.append(f"        if (HEX8_NEED({mask})) {tgt} += sgn * ({_printer.doprint(ex)});")
            continue
        lines.append("        {")
        for n, ((tgt, _, mask), ex) in enumerate(part):
            lines.append(f"        double hA{n} = 0.0; if (HEX8_NEED({mask})) hA{n} = {tgt};")
        for n, ((tgt, _, mask), ex) in enumerate(part):
            lines.append(f"        if (HEX8_NEED({mask})) {tgt} = hA{n} + sgn * ({_printer.doprint(ex)});")
        lines.append("        }")
    nops = count_ops([ex for _, ex in repl] + list(red))
    return lines, nops


def row_masks(combos):
    """bit a of the mask of a basis function is set if it contributes to strain row a (xx,xy,xz,yy,yz,zz)"""
    m = {}
    for n, c in enumerate(combos):
        for key in c:
            m[key] = m.get(key, 0) | (1 << (n // 6))
    return m


def emit_basis(keys, index, prefix, masks):
    """code accumulating sgn * basis function into ACC(index[key]), grouped by (potential, order) so that each
    group is a short block with its own common subexpressions (short live ranges, no spills)"""
    groups = {}
    for key in keys:
        groups.setdefault((key[0], sum(key[1])), []).append(key)
    lines, total = [], 0
    # large groups are cut into blocks of at most MAXF functions: a block is register-allocated on its own
    # (HEX8_GROUP_BARRIER), at the price of recomputing a few shared powers
    MAXF = int(os.environ.get('HEX8_MAXF', '99'))   # measured: splitting costs more (recomputed powers) than it saves
    blocks = []
    for gname, gkeys in sorted(groups.items()):
        for i0 in range(0, len(gkeys), MAXF):
            blocks.append((gname, gkeys[i0:i0 + MAXF]))
    for gi, (gname, gkeys) in enumerate(blocks):
        assigns = [(f"ACC({index[key]})", TD(*key), masks[key]) for key in gkeys]
        blk, nops = cse_block(assigns, f"{prefix}{gi}_")
        lines.append(f"    {{   /* {gname[0]}, derivative order {gname[1]}: {len(gkeys)} functions */")
        lines.append("        HEX8_GROUP_BARRIER")
        lines += blk
        lines.append("    }")
        total += nops
    return "\n".join(lines), total


def emit_combine(combos_real, combos_img, idx_r, idx_i):
    lines = []
    exprs = []
    accr = {k_: sp.Symbol(f"ACCR{n}") for k_, n in idx_r.items()}
    acci = {k_: sp.Symbol(f"ACCI{n}") for k_, n in idx_i.items()}
    for cr, ci in zip(combos_real, combos_img):
        ex = sum(c * accr[k_] for k_, c in cr.items()) + sum(c * acci[k_] for k_, c in ci.items())
        exprs.append(ex)
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("c"), optimizations="basic")
    import re
    def fix(txt):
        txt = re.sub(r"ACCR(\d+)", r"ACCR(\1)", txt)
        return re.sub(r"ACCI(\d+)", r"ACCI(\1)", txt)
    for s_, ex in repl:
        lines.append(f"    const double {s_} = {fix(_printer.doprint(ex))};")
    for n, ex in enumerate(red):
        lines.append(f"    Q[{n}] = {fix(_printer.doprint(ex))};")
    return "\n".join(lines), count_ops([ex for _, ex in repl] + list(red))


def main():
    verify()
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(os.path.dirname(here))   # <repo>/oetqf.jl_b200/derive -> <repo>
    cr, ci = strain_combos(False), strain_combos(True)
    keys_r = sorted({k_ for c in cr for k_ in c})
    keys_i = sorted({k_ for c in ci for k_ in c})
    idx_r = {k_: n for n, k_ in enumerate(keys_r)}
    idx_i = {k_: n for n, k_ in enumerate(keys_i)}
    body_r, ops_r = emit_basis(keys_r, idx_r, "r", row_masks(cr))
    body_i, ops_i = emit_basis(keys_i, idx_i, "m", row_masks(ci))
    body_c, ops_c = emit_combine(cr, ci, idx_r, idx_i)
    print(f"real basis: {len(keys_r)} functions, {ops_r} ops/corner; image basis: {len(keys_i)} functions, "
          f"{ops_i} ops/corner; combination: {ops_c} ops/pair", file=sys.stderr)
    header = ("// GENERATED by oetqf.jl_b200/derive/hex8_derive.py -- do not edit.\n"
              "// The strain of a uniformly strained cuboid is a linear combination (coefficients in alpha and the\n"
              "// receiver depth x3 only) of corner sums of BASIS functions = derivatives of the triple antiderivatives\n"
              "// of P0 = 1/R, P1 = R, P3 = R - R3 ln(R+R3).  hex8_basis_* accumulate sgn * basis into ACC(b) for one\n"
              "// corner; hex8_combine turns the 8-corner sums into Q[(il),(jk)] (times 8*pi*mu), pairs xx,xy,xz,yy,yz,zz.\n"
              "// The includer defines ACC(b) (accumulator b of the current function), ACCR(b)/ACCI(b) (real/image sums) and\n"
              "// HEX8_NEED(mask) (non-zero if any strain row xx,xy,xz,yy,yz,zz = bit 0..5 of mask is wanted) and\n"
              "// HEX8_GROUP_BARRIER (expanded at the start of every group; may be empty, or an optimisation barrier\n"
              "// on the inputs that stops the compiler from sharing temporaries across groups).\n"
              "// Inputs: R1,R2,R3 corner vector, R its norm, w_c = R+R_c, q_c = R^2-R_c^2, iR/iw_c/iq_c reciprocals,\n"
              "// L_c = ln(w_c), A_c = atan(R_a R_b/(R_c R)), Ba = atan(R1/R2), Bb = atan(R2/R1).\n"
              f"#define HEX8_NB_REAL {len(keys_r)}\n#define HEX8_NB_IMAGE {len(keys_i)}\n")
    sig_common = ("double R1, double R2, double R3, double R, double w1, double w2, double w3, double q1, double q2, "
                  "double q3, double iR, double iw1, double iw2, double iw3, double iq1, double iq2, double iq3, "
                  "double L1, double L2, double L3, double A1, double A2, double A3")
    for path, qual in ((os.path.join(root, "oracle", "hex8_gen.inc"), "static inline"),
                       (os.path.join(root, "oetqf.jl_b200", "csrc", "hex8_gen.cuh"), "__device__ __forceinline__")):
        with open(path, "w") as fh:
            fh.write(header)
            fh.write(f"#define HEX8_BASIS_REAL_BODY \\\n" + " \\\n".join(body_r.split("\n")) + "\n\n")
            fh.write(f"#define HEX8_BASIS_IMAGE_BODY \\\n" + " \\\n".join(body_i.split("\n")) + "\n\n")
            fh.write(f"#define HEX8_COMBINE_BODY \\\n" + " \\\n".join(body_c.split("\n")) + "\n")
        print("wrote", path)


if __name__ == "__main__":
    main()
