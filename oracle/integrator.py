"""oracle/integrator.py -- CPU restatement of the adaptive Tsit5 stepping the reference obtains from
OrdinaryDiffEq (`solve(prob, Tsit5(); reltol, abstol, dt, dtmax)`, test/tests.jl:11, src/io.jl:128-130).
TEST INFRASTRUCTURE ONLY.

OrdinaryDiffEq is a third-party dependency absent from /root/reference (Project.toml compat "6"); this
restates its published algorithm: the Tsitouras 5(4) tableau, the scaled RMS error norm over all state
scalars, and the PI controller with beta1 = 7/50, beta2 = 2/25, gamma = 0.9, qmin = 0.2, qmax = 10,
qoldinit = 1e-4.  Parity with OrdinaryDiffEq's exact step sequence is UNPINNED (no Julia here); what the
tests pin is CPU-oracle == GPU stepping, plus order-of-accuracy checks on problems with known solutions.
"""
from __future__ import annotations

import numpy as np

C = np.array([0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0])
A = [
    [],
    [0.161],
    [-0.008480655492356989, 0.335480655492357],
    [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
    [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
    [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
    [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774],
]
BTILDE = np.array([-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                   -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 0.015151515151515152])


def tsit5(f, u0: np.ndarray, t0: float, tstop: float, reltol=1e-3, abstol=1e-6, dt0=0.0, dtmax=0.0,
          maxiters=100000, fixed=False):
    """Integrate u' = f(u) (autonomous, flat float64 vector).  Returns (ts, us, stats) with the state after
    every accepted step (index 0 = initial condition)."""
    u = np.array(u0, dtype=np.float64)
    t = float(t0)
    dtmax = dtmax if dtmax > 0 else (tstop - t0)
    dt = dt0 if dt0 > 0 else 1e-6 * (tstop - t0)
    dt = min(dt, dtmax)
    if t + dt > tstop:
        dt = tstop - t
    qold = 1e-4
    beta1, beta2, gamma, qmin, qmax = 7 / 50, 2 / 25, 0.9, 0.2, 10.0
    k = [None] * 7
    k[0] = f(u)
    ts, us = [t], [u.copy()]
    naccept = nreject = 0
    done = False
    it = 0
    while not done and it < maxiters:
        it += 1
        for s in range(1, 7):
            acc = np.zeros_like(u)
            for j in range(s):
                acc = acc + A[s][j] * k[j]          # same accumulation order as the device stage kernel
            y = u + dt * acc
            k[s] = f(y)
        unew = y
        err = np.zeros_like(u)
        for j in range(7):
            err = err + BTILDE[j] * k[j]
        err = err * dt
        sk = abstol + reltol * np.maximum(np.abs(u), np.abs(unew))
        eest = float(np.sqrt(np.sum((err / sk) ** 2) / u.size))
        if fixed:
            accept = True
        else:
            q11 = eest ** beta1
            q = q11 / qold ** beta2
            q = max(1 / qmax, min(1 / qmin, q / gamma))
            accept = eest <= 1.0
        if accept:
            naccept += 1
            t += dt
            u = unew
            k[0] = k[6]
            if fixed:
                if t + dt > tstop:
                    dt = tstop - t
                if t >= tstop - 4e-16 * abs(tstop) or dt <= 0:
                    done = True
            else:
                qold = max(eest, 1e-4)
                dtn = min(dt / q, dtmax)
                if t >= tstop - 4e-16 * abs(tstop):
                    done = True
                    t = tstop
                elif t + dtn > tstop:
                    dtn = tstop - t
                dt = dtn
            ts.append(t)
            us.append(u.copy())
        else:
            nreject += 1
            dt = dt / min(1 / qmin, q11 / gamma)
    return np.array(ts), us, dict(naccept=naccept, nreject=nreject)


def _pi_controller(eest, qold, dt, beta1=7 / 50, beta2=2 / 25, gamma=0.9, qmin=0.2, qmax=10.0):
    q11 = eest ** beta1
    q = q11 / qold ** beta2
    q = max(1 / qmax, min(1 / qmin, q / gamma))
    return q11, q


def _tsit5_step(f, u, k1, dt):
    """one Tsit5 step from (u, k1 = f(u)): returns (unew, err, f(unew)); 6 evaluations"""
    k = [k1] + [None] * 6
    for s in range(1, 7):
        acc = np.zeros_like(u)
        for j in range(s):
            acc = acc + A[s][j] * k[j]
        y = u + dt * acc
        k[s] = f(y)
    err = np.zeros_like(u)
    for j in range(7):
        err = err + BTILDE[j] * k[j]
    return y, err * dt, k[6]


def _adams_step(f, u, hist_f, hist_dt, dt):
    """one PECE step of the order-5 variable-coefficient Adams pair on the grid dts = (dt, h_{n-1}, h_{n-2}, ...).
    hist_f = [f(t_n), f(t_{n-1}), ...] (>= 5 samples), hist_dt = [h_{n-1}, h_{n-2}, ...] (>= 4 accepted steps).
    Returns (unew, err, f(unew)); 2 evaluations."""
    dts = [dt] + list(hist_dt[:4])

    # phi_j(m): modified divided differences, rebuilt from the raw samples through
    # phi_{j+1}(m) = phi_j(m) - phi*_j(m-1), phi*_j(m-1) = beta_j(m-1) phi_j(m-1),
    # beta_j(m-1) = prod_{i<j-1} (t_m - t_{m-1-i}) / (t_{m-1} - t_{m-2-i})
    def phis(level, depth):
        out = [hist_f[level]]
        if depth > 1:
            prev = phis(level + 1, depth - 1)
            grid = hist_dt[level:]
            beta, xi, xi0 = 1.0, grid[0], 0.0
            for j in range(1, depth):
                if j > 1:
                    xi0 += grid[j - 1]
                    beta = beta * xi / xi0
                    xi += grid[j - 1]
                out.append(out[j - 1] - beta * prev[j - 1])
        return out

    phi_n = phis(0, 5)
    beta = [1.0]
    xi, xi0 = dts[0], 0.0
    for i in range(1, 5):
        xi0 += dts[i]
        beta.append(beta[i - 1] * xi / xi0)
        xi += dts[i]
    phistar = [beta[i] * phi_n[i] for i in range(5)]
    # g_j(n): c_{1,q} = 1/q, c_{2,q} = 1/(q(q+1)), c_{j,q} = c_{j-1,q} - c_{j-1,q+1} dt / (t_{n+1} - t_{n+2-j})
    kk = 6
    c = np.zeros((kk + 1, kk + 2))
    g = np.zeros(kk + 1)
    xi = dts[0]
    for i in range(1, kk + 1):
        if i > 2:
            xi += dts[i - 2]
        for q in range(1, kk - (i - 1) + 1):
            if i == 1:
                c[i, q] = 1.0 / q
            elif i == 2:
                c[i, q] = 1.0 / (q * (q + 1))
            else:
                c[i, q] = c[i - 1, q] - dt / xi * c[i - 1, q + 1]
        g[i] = c[i, 1] * dt
    p = u.copy()
    for j in range(1, 5):
        p = p + g[j] * phistar[j - 1]
    fp = f(p)
    phi_np1 = [fp]
    for j in range(1, 6):
        phi_np1.append(phi_np1[j - 1] - phistar[j - 1])
    unew = p + g[5] * phi_np1[4]
    err = (g[6] - g[5]) * phi_np1[5]
    return unew, err, f(unew)


def _eest(err, u, unew, reltol, abstol):
    sk = abstol + reltol * np.maximum(np.abs(u), np.abs(unew))
    return float(np.sqrt(np.sum((err / sk) ** 2) / u.size))


def vcabm5(f, u0: np.ndarray, t0: float, tstop: float, reltol=1e-3, abstol=1e-6, dt0=0.0, dtmax=0.0,
           maxiters=100000):
    """Variable-coefficient Adams-Bashforth-Moulton PECE of order 5 -- the class of method the reference's
    example asks OrdinaryDiffEq for (`VCABM5()`, examples/otf-with-mantle.jl:160-162).

    Restates the published algorithm (Hairer, Norsett & Wanner, Solving ODEs I, III.5, which OrdinaryDiffEq
    cites for its VCABM family): modified divided differences phi_j(n), phi*_j(n) = beta_j(n) phi_j(n) and the
    coefficients g_j(n) from the c_{j,q} recurrence on the grid of past steps;
        predictor  p      = u_n + sum_{j=1..4} g_j phi*_j(n)              (Adams-Bashforth, order 4)
        corrector  u_{n+1} = p + g_5 phi_5(n+1)   with phi(n+1) from f(p)   (Adams-Moulton, order 5)
        estimate   (g_6 - g_5) phi_6(n+1)                                 (difference to the order-6 formula)
    then f(u_{n+1}) is evaluated for the next step (PECE, 2 evaluations per step).  The first four steps are
    adaptive Tsit5 steps that build the history.  Step-size control: the same PI controller as `tsit5` above
    (order-5 constants).  Parity with OrdinaryDiffEq's exact step sequence is UNPINNED; the tests pin
    CPU-oracle == GPU stepping (the GPU uses the equivalent Lagrange form of the same formulas) and the
    order of accuracy on problems with known solutions.
    """
    u = np.array(u0, dtype=np.float64)
    t = float(t0)
    dtmax = dtmax if dtmax > 0 else (tstop - t0)
    dt = dt0 if dt0 > 0 else 1e-6 * (tstop - t0)
    dt = min(dt, dtmax)
    if t + dt > tstop:
        dt = tstop - t
    qold = 1e-4
    hist_f = [f(u)]            # f(t_n), f(t_{n-1}), ... newest first
    hist_dt: list = []         # accepted step sizes, newest first
    ts, us = [t], [u.copy()]
    naccept = nreject = nrhs = 0
    done = False
    it = 0
    while not done and it < maxiters:
        it += 1
        if len(hist_dt) < 4:
            unew, err, fnew = _tsit5_step(f, u, hist_f[0], dt)      # starting procedure
            nrhs += 6
        else:
            unew, err, fnew = _adams_step(f, u, hist_f, hist_dt, dt)
            nrhs += 2
        eest = _eest(err, u, unew, reltol, abstol)
        q11, q = _pi_controller(eest, qold, dt)
        if eest <= 1.0:
            naccept += 1
            t += dt
            u = unew
            hist_f = [fnew] + hist_f[:5]
            hist_dt = [dt] + hist_dt[:4]
            qold = max(eest, 1e-4)
            dtn = min(dt / q, dtmax)
            if t >= tstop - 4e-16 * abs(tstop):
                done = True
                t = tstop
            elif t + dtn > tstop:
                dtn = tstop - t
            dt = dtn
            ts.append(t)
            us.append(u.copy())
        else:
            nreject += 1
            dt = dt / min(1 / 0.2, q11 / 0.9)
    return np.array(ts), us, dict(naccept=naccept, nreject=nreject, nrhs=nrhs + 1)


def vcabm5_on_grid(f, u0: np.ndarray, ts, reltol=1e-3, abstol=1e-6):
    """The same stepping as `vcabm5`, but along a GIVEN grid of accepted times (e.g. the one a device run chose):
    returns (us, eests), the state after every step and the scaled error estimate the controller would have seen.

    Why it exists: where the solution is almost steady the error estimate is pure round-off (1e-10 ... 1e-7), and
    the controller turns O(1) relative noise in it into per-cent differences of the next step, so two correct
    implementations need not pick the same grid.  Following the device's grid compares the formulas step by step
    (states), and the device's controller wherever the estimate is above the noise (tests/test_gpu_solve.py)."""
    u = np.array(u0, dtype=np.float64)
    hist_f, hist_dt = [f(u)], []
    us, eests = [u.copy()], []
    for n in range(1, len(ts)):
        dt = float(ts[n] - ts[n - 1])
        if len(hist_dt) < 4:
            unew, err, fnew = _tsit5_step(f, u, hist_f[0], dt)
        else:
            unew, err, fnew = _adams_step(f, u, hist_f, hist_dt, dt)
        eests.append(_eest(err, u, unew, reltol, abstol))
        u = unew
        hist_f = [fnew] + hist_f[:5]
        hist_dt = [dt] + hist_dt[:4]
        us.append(u.copy())
    return us, np.array(eests)


def pi_next_dt(eest, qold, dt):
    """step the PI controller proposes after an accepted step (before dtmax / tstop clipping)"""
    _, q = _pi_controller(eest, qold, dt)
    return dt / q
