"""Crowd initialisation on the host (simulations.py:104-160) with the reference's exact RNG consumption.

The reference tests every trial position against the WHOLE grid: ``sqrt((X_opt-x)**2 + (Y_opt-y)**2) < r_in``
(simulations.py:132,135).  Only nodes within r_in = 0.2 m can satisfy that, so the same numpy expression is
evaluated on a small window of the node coordinates -- identical mask bits, identical accept/reject
decisions, identical use of the legacy global ``np.random`` stream (two ``uniform(...,1)`` draws per trial,
then one ``normal(1.34,0.26,loc_N)`` per box), but O(window) instead of O(Nx*Ny) per trial.
"""
from __future__ import annotations

import numpy as np


def _window(coords, c, r):
    step = coords[1] - coords[0]
    lo = int(np.floor((c - r) / step)) - 2
    hi = int(np.ceil((c + r) / step)) + 3
    return max(lo, 0), min(hi, len(coords))


def place_box(box, X1, Y1, place_ped, r_in=0.2, rng=np.random):
    """``place_box_py`` below, run by the C helper ``oc_place_box`` of liboc_b200.so (same arithmetic, same MT19937
    stream, ~100x faster: the generator state is handed to C and written back).  Falls back to the numpy version for a
    mask that is not a C-contiguous float64 array."""
    from . import _lib
    import ctypes as C
    loc_N = int(box[4] * box[2] * box[3])                                   # simulations.py:122
    st = rng.get_state()
    ok = (st[0] == "MT19937" and isinstance(place_ped, np.ndarray) and place_ped.dtype == np.float64
          and place_ped.flags.c_contiguous and place_ped.shape == (len(Y1), len(X1)))
    if not ok:
        return place_box_py(box, X1, Y1, place_ped, r_in, rng)
    key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
    pos = C.c_int(int(st[2]))
    xs, ys = np.empty(loc_N), np.empty(loc_N)
    b5 = np.ascontiguousarray(np.asarray(box[:5], dtype=np.float64))
    Xc, Yc = np.ascontiguousarray(X1, dtype=np.float64), np.ascontiguousarray(Y1, dtype=np.float64)
    rc = _lib.load().oc_place_box(_lib._hp(b5), _lib._hp(Xc), len(Xc), _lib._hp(Yc), len(Yc), _lib._hp(place_ped),
                                  float(r_in), key.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(pos), _lib._hp(xs),
                                  _lib._hp(ys), loc_N)
    if rc < 0:
        _lib.check(int(rc))
    rng.set_state((st[0], key, pos.value, st[3], st[4]))
    v_des = rng.normal(1.34, 0.26, size=loc_N)                           # :140
    return xs, ys, v_des


def place_box_py(box, X1, Y1, place_ped, r_in=0.2, rng=np.random):
    """Rejection-sample ``int(rho*w*h)`` positions inside ``box`` = [cx, cy, w, h, rho, ...]; returns xs, ys, v_des.

    ``place_ped`` (Ny,Nx) is the occupancy mask shared by all boxes (simulations.py:110) and is updated in place.
    ``rng``: the legacy generator to draw from -- the global ``np.random`` module like the reference, or a
    ``np.random.RandomState(seed)`` (same stream as ``np.random.seed(seed)``) for ensemble members."""
    loc_N = int(box[4] * box[2] * box[3])                                   # simulations.py:122
    xs, ys = np.empty(loc_N), np.empty(loc_N)
    placed = 0
    while placed < loc_N:
        x_in = rng.uniform(box[0] - box[2] / 2, box[0] + box[2] / 2, 1)   # :130
        y_in = rng.uniform(box[1] - box[3] / 2, box[1] + box[3] / 2, 1)   # :131
        j0, j1 = _window(X1, x_in[0], r_in)
        i0, i1 = _window(Y1, y_in[0], r_in)
        near = np.sqrt((X1[None, j0:j1] - x_in) ** 2 + (Y1[i0:i1, None] - y_in) ** 2) < r_in
        sub = place_ped[i0:i1, j0:j1]
        if 1 in sub[near]:                                                      # :132
            continue
        sub[near] = 1                                                           # :135
        xs[placed], ys[placed] = x_in[0], y_in[0]
        placed += 1
    v_des = rng.normal(1.34, 0.26, size=loc_N)                           # :140
    return xs, ys, v_des
