"""Shared builders for the parity tests: the same NumPy inputs go to the oracle and to the product."""
import numpy as np

import workloads as W
from oracle import ref


def meshes(oq, fs: W.FaultSpec, bs: W.BoxSpec = None):
    mf_o = ref.fault_mesh(fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    mf_p = oq.gen_mesh("RectOkada", fs.x, fs.xi, fs.dx, fs.dxi, fs.dip)
    if bs is None:
        return mf_o, mf_p
    ma_o = ref.hex8_box(*bs.args())
    ma_p = oq.gen_mesh("BEMHex8Mesh", *bs.args())
    return mf_o, mf_p, ma_o, ma_p


def scaled_err(got, want, axis=None):
    """max |got-want| / max(|want| over the row/array): the scale-aware criterion of SURVEY.md §7 for
    cancellation-dominated entries (relative to the largest entry of the same row)."""
    got, want = np.asarray(got), np.asarray(want)
    scale = np.max(np.abs(want), axis=axis, keepdims=axis is not None)
    # rows/columns that vanish by symmetry hold only round-off of O(1) corner terms: floor at 1e-3 of the
    # largest entry of the whole array
    scale = np.maximum(scale, 1e-3 * np.max(np.abs(want)))
    scale = np.where(scale == 0, 1.0, scale)
    return float(np.max(np.abs(got - want) / scale))


def rel_err(got, want):
    got, want = np.asarray(got), np.asarray(want)
    den = np.where(want == 0, 1.0, np.abs(want))
    return float(np.max(np.abs(got - want) / den))
