"""Debug aid: device VCABM5 vs the oracle on the decay problem (prints the two accepted-time sequences)."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import oetqf_b200 as oq
from oracle import integrator, ref

oq.init(0)
nx, nxi = 4, 3
rng = np.random.default_rng(3)
a, b, L, sig = (rng.uniform(0.5, 1.5, (nx, nxi)) for _ in range(4))
v, th, dl = (rng.uniform(0.5, 1.5, (nx, nxi)) for _ in range(3))
st = np.zeros((nx, nxi, nxi), order="F")
pf_p = oq.RateStateQuasiDynamicProperty(a, b, L, sig, 0.7, 0.3, 0.6, 0.9)
pf_o = ref.FaultProp(a, b, L, sig, 0.7, 0.3, 0.6, 0.9)
u0 = oq.ArrayPartition(v, th, dl)
shapes = [x.shape for x in u0.x]
pack = lambda parts: np.concatenate([np.asarray(p).reshape(-1, order="F") for p in parts])
def unpack(u):
    out, off = [], 0
    for s in shapes:
        n = int(np.prod(s)); out.append(u[off:off + n].reshape(s, order="F")); off += n
    return out
def f(u):
    vv, tt, _ = unpack(u)
    return pack(ref.rhs_fault(pf_o, st, vv, tt, form="toeplitz"))
ts, us, stats = integrator.vcabm5(f, pack(u0.x), 0.0, 2.0, reltol=1e-8, abstol=1e-10, dt0=1e-3)
for form in ("dense", "fft"):
    prob = oq.assemble(st, pf_p, u0, (0.0, 2.0), gf11_form=form)
    sol = oq.solve(prob, oq.VCABM5(), reltol=1e-8, abstol=1e-10, dt=1e-3)
    print(form, sol.retcode, sol.stats, stats, len(sol.t), len(ts))
    m = min(len(ts), len(sol.t))
    for k in range(min(m, 14)):
        print(k, ts[k], sol.t[k], np.max(np.abs(pack(sol.u[k].x) - us[k]) / np.abs(us[k])))
    # fixed-step run: isolates the formulas from the controller
    solf = oq.solve(prob, oq.VCABM5(), dt=0.01, adaptive=False)
    print("fixed", solf.retcode, solf.stats, len(solf.t))
