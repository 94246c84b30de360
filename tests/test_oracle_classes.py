"""The translation classes the class-table assembly and the class-form operands rest on, checked on the CPU ORACLE's
own matrices: every (receiver, source) pair must carry the 6x6 (1x6) block of the pair that stands for its class.
The oracle evaluates every pair on its own coordinates (GF.jl:206-225, :262-290 restated in oracle/greens.c), so this
pins the invariance itself -- not the product's use of it."""
import numpy as np

import workloads as W
from oracle import ref


def _class_keys(oq, ma, mf, n_recv):
    from oetqf_b200 import gf
    ne = len(ma)
    recv = np.repeat(np.arange(n_recv, dtype=np.int32), ne)
    src = np.tile(np.arange(ne, dtype=np.int32), n_recv)
    counts, rjx, rix, rjyz, riyz = gf.hex8_pair_classes(ma, mf, 0, n_recv, recv, src)
    return counts, recv, src, rjx, rix, rjyz, riyz


def test_mantle_mantle_blocks_are_constant_on_translation_classes(oq):
    bs = W.BoxSpec(-16e3, -6e3, -8e3, 32e3, 12e3, -20e3, 6, 3, 3, tuple(np.cumprod(np.ones(3) * 1.3)))
    ma = oq.gen_mesh("BEMHex8Mesh", *bs.args())
    mao = ref.hex8_box(*bs.args())
    ne = len(ma)
    g = ref.gf_mantle_mantle(mao, W.LAM, W.MU)                        # [6 ne, 6 ne], row k*ne + j, column p*ne + i
    blocks = g.reshape(6, ne, 6, ne).transpose(1, 3, 0, 2)            # [receiver, source, k, p]
    counts, recv, src, rjx, rix, rjyz, riyz = _class_keys(oq, ma, None, ne)
    assert counts[0] * counts[1] * 4 <= ne * ne                        # the mesh has classes worth using
    scale = np.max(np.abs(g))
    worst = 0.0
    first = {}
    for r, s, key in zip(recv, src, zip(rjx, rix, rjyz, riyz)):
        ref_pair = first.setdefault(key, (r, s))
        worst = max(worst, float(np.max(np.abs(blocks[r, s] - blocks[ref_pair]))))
    assert len(first) == counts[0] * counts[1]
    assert worst <= 1e-10 * scale, worst / scale


def test_mantle_fault_blocks_are_constant_on_translation_classes(oq):
    fs = W.FaultSpec(24e3, 6e3, 1e3, 1e3)
    bs = W.BoxSpec(-12e3, -6e3, -6e3, 24e3, 12e3, -20e3, 6, 3, 3, tuple(np.cumprod(np.ones(3) * 1.3)))
    mf = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    ma = oq.gen_mesh("BEMHex8Mesh", *bs.args())
    mfo, mao = ref.fault_mesh(fs.x, fs.xi, fs.dx, fs.dxi, fs.dip), ref.hex8_box(*bs.args())
    nf, ne = mf.nx * mf.nxi, len(ma)
    g = ref.gf_mantle_fault(mao, mfo, W.LAM, W.MU)                    # [nf, 6 ne], column p*ne + i
    blocks = g.reshape(nf, 6, ne).transpose(0, 2, 1)                  # [receiver, source, p]
    counts, recv, src, rjx, rix, rjyz, riyz = _class_keys(oq, ma, mf, nf)
    assert counts[0] * counts[1] < nf * ne                          # fewer classes than pairs
    scale = np.max(np.abs(g))
    worst = 0.0
    first = {}
    for r, s, key in zip(recv, src, zip(rjx, rix, rjyz, riyz)):
        ref_pair = first.setdefault(key, (r, s))
        worst = max(worst, float(np.max(np.abs(blocks[r, s] - blocks[ref_pair]))))
    assert worst <= 1e-10 * scale, worst / scale
