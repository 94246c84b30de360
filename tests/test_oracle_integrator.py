"""CPU checks of the oracle integrators (oracle/integrator.py) on problems with known solutions.  OrdinaryDiffEq
is absent (parity with its exact step sequence is unpinned); what can be pinned on CPU is the order of accuracy,
and -- for the Adams pair -- that the divided-difference form the oracle follows (Hairer-Norsett-Wanner III.5)
and the Lagrange form the device kernel uses (csrc/solve.cu: abm_coef_kernel) are the same formula."""
import numpy as np

from oracle import integrator as I


def _f(u):
    return np.array([u[1], -u[0], -0.5 * u[2]])


U0 = np.array([1.0, 0.0, 2.0])
EXACT = np.array([np.cos(10.0), -np.sin(10.0), 2 * np.exp(-5.0)])


def test_tsit5_converges_at_order_five():
    errs, steps = [], []
    for tol in (1e-6, 1e-8, 1e-10):
        ts, us, st = I.tsit5(_f, U0, 0.0, 10.0, reltol=tol, abstol=tol * 1e-2, dt0=1e-3)
        assert ts[-1] == 10.0 and st["nreject"] <= 2
        errs.append(np.abs(us[-1] - EXACT).max())
        steps.append(st["naccept"])
    for k in (0, 1):
        order = np.log(errs[k] / errs[k + 1]) / np.log(steps[k + 1] / steps[k])
        assert 4.5 < order < 6.5, (errs, steps)


def test_vcabm5_converges_at_order_five_with_two_evaluations_per_step():
    errs, steps = [], []
    for tol in (1e-6, 1e-8, 1e-10):
        ts, us, st = I.vcabm5(_f, U0, 0.0, 10.0, reltol=tol, abstol=tol * 1e-2, dt0=1e-3)
        assert ts[-1] == 10.0
        # 1 initial evaluation, 6 per starting (Tsit5) step, 2 per Adams step
        assert st["nrhs"] == 1 + 6 * 4 + 2 * (st["naccept"] + st["nreject"] - 4)
        errs.append(np.abs(us[-1] - EXACT).max())
        steps.append(st["naccept"])
    assert errs[-1] < 1e-7
    for k in (0, 1):
        order = np.log(errs[k] / errs[k + 1]) / np.log(steps[k + 1] / steps[k])
        assert 4.3 < order < 6.5, (errs, steps)


def _lagrange_weights(dt, hdt):
    """the device kernel's formulation: integrals over [0, dt] of the Lagrange basis through the nodes
    dt (predicted derivative), 0, -h1, -h1-h2, ... by 3-point Gauss-Legendre (exact to degree 5)"""
    tau = [dt, 0.0]
    for h in hdt:
        tau.append(tau[-1] - h)
    gx, gw = (-0.7745966692414834, 0.0, 0.7745966692414834), (5 / 9, 8 / 9, 5 / 9)

    def w(lo, hi, m):
        den = np.prod([tau[m] - tau[k] for k in range(lo, hi + 1) if k != m])
        acc = 0.0
        for x, ww in zip(gx, gw):
            xq = 0.5 * dt * (1 + x)
            acc += ww * np.prod([xq - tau[k] for k in range(lo, hi + 1) if k != m])
        return 0.5 * dt * acc / den

    wp = [w(1, 4, m) for m in range(1, 5)]
    wc = [w(0, 4, m) for m in range(5)]
    w6 = [w(0, 5, m) for m in range(6)]
    return wp, wc, [w6[j] - (wc[j] if j < 5 else 0.0) for j in range(6)]


def test_adams_lagrange_form_equals_divided_difference_form():
    def f(u):
        return np.array([u[1], -u[0] * (1 + 0.3 * u[2]), -0.5 * u[2] + 0.1 * u[0]])

    ts, us, _ = I.vcabm5(f, U0, 0.0, 10.0, reltol=1e-8, abstol=1e-10, dt0=1e-3)
    hf, hd, worst, checked = [f(us[0])], [], 0.0, 0
    for n in range(1, len(ts)):
        dt = ts[n] - ts[n - 1]
        if len(hd) >= 4:
            wp, wc, _ = _lagrange_weights(dt, hd[:4])
            p = us[n - 1] + sum(wp[m] * hf[m] for m in range(4))
            un = us[n - 1] + wc[0] * f(p) + sum(wc[m + 1] * hf[m] for m in range(4))
            worst = max(worst, np.abs(un - us[n]).max())
            checked += 1
        hf = [f(us[n])] + hf[:5]
        hd = [dt] + hd[:4]
    assert checked > 100 and worst < 5e-14, (checked, worst)


def test_adams_weights_reduce_to_the_classical_constant_step_coefficients():
    wp, wc, we = _lagrange_weights(1.0, [1.0] * 4)
    np.testing.assert_allclose(wp, np.array([55, -59, 37, -9]) / 24, rtol=1e-13)               # AB4
    np.testing.assert_allclose(wc, np.array([251, 646, -264, 106, -19]) / 720, rtol=1e-13)     # AM, 4-step
    am5 = np.array([475, 1427, -798, 482, -173, 27]) / 1440                                    # AM, 5-step
    np.testing.assert_allclose(we, am5 - np.append(wc, 0.0), rtol=1e-12, atol=1e-15)
