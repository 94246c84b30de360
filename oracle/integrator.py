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
