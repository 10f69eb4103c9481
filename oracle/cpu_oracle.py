"""ctypes bindings for the CPU oracle library (oracle/_build/liboc_oracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboc_oracle.so")
_lib = None

dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oc_oracle_hjb.c", "oc_oracle_gcfm.c")]
    srcs.append(os.path.join(_HERE, "..", "optimal_crowds_b200", "csrc", "oc_math.h"))
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oco_wall_argmin.restype = C.c_long
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(dp)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class HjbStats(C.Structure):
    _fields_ = [("nfev", C.c_int), ("n_accepted", C.c_int), ("n_rejected", C.c_int), ("status", C.c_int),
                ("n_out", C.c_int), ("h0", C.c_double)]


class GcfmParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("dt", "dt2", "half_noise", "relaxation", "v_max", "cutoff", "a_min", "tau_a", "b_min", "b_max",
                 "eta", "eta_walls", "cos_fov", "one_minus_cos_fov", "dx", "dy", "room_length", "room_height")] + \
               [("Ny", C.c_int), ("Nx", C.c_int)]


class Key(C.Structure):
    _fields_ = [("V", dp), ("vx_opt", dp), ("vy_opt", dp), ("nt_opt", C.c_int), ("n_slices", C.c_int),
                ("doors", dp), ("n_doors", C.c_int)]


def gcfm_params(cfg: dict, room_length: float, room_height: float, Ny: int, Nx: int) -> GcfmParams:
    """Same Python expressions as the reference evaluates (simulations.py:81-97,303,314; pedestrians.py:74,262)."""
    p = GcfmParams()
    dt = cfg["dt"]
    p.dt, p.dt2 = dt, dt ** 2
    p.half_noise = cfg["hjb_params"]["sigma"] / 2
    p.relaxation, p.v_max, p.cutoff = cfg["relaxation"], cfg["v_max"], cfg["repulsion_cutoff"]
    p.a_min, p.tau_a, p.b_min, p.b_max = cfg["b_min"], cfg["tau_a"], cfg["b_min"], cfg["b_max"]
    p.eta, p.eta_walls = cfg["eta"], cfg["eta_walls"]
    c = float(np.cos(0.7 * np.pi))
    p.cos_fov, p.one_minus_cos_fov = c, 1 - c
    p.dx = p.dy = cfg["grid_step"]
    p.room_length, p.room_height, p.Ny, p.Nx = room_length, room_height, Ny, Nx
    return p


def hjb_rhs(phi, V, m, dx, dy, sigma, mu, g):
    V = _d(V); Ny, Nx = V.shape
    phi = _d(phi); out = np.empty(Ny * Nx)
    m = None if m is None else _d(m)
    lib().oco_hjb_rhs(_p(phi), _p(V), _p(m), Ny, Nx, C.c_double(dx), C.c_double(dy), C.c_double(sigma),
                      C.c_double(mu), C.c_double(g), _p(out))
    return out


def hjb_solve(V, m, T, nt, dx=0.05, dy=0.05, sigma=0.2, mu=5.0, g=-0.005, rtol=1e-3, atol=1e-6, trace_cap=100000,
              forced_h=None, forced_err=None):
    """Returns (phi_cols (nt, Ny*Nx) with row k = sol.y[:,k], stats dict, trace_h, trace_err)."""
    V = _d(V); Ny, Nx = V.shape
    m = None if m is None else _d(m)
    t_eval = np.linspace(T, 0, nt)  # optimals.py:194
    phi = np.empty((nt, Ny * Nx))
    st = HjbStats()
    th, te = np.empty(trace_cap), np.empty(trace_cap)
    ntr = C.c_int()
    lib().oco_hjb_solve(_p(V), _p(m), Ny, Nx, C.c_double(dx), C.c_double(dy), C.c_double(sigma), C.c_double(mu),
                        C.c_double(g), C.c_double(T), _p(t_eval), nt, C.c_double(rtol), C.c_double(atol), _p(phi),
                        C.byref(st), _p(th), _p(te), trace_cap, C.byref(ntr),
                        None if forced_h is None else _p(_d(forced_h)),
                        None if forced_err is None else _p(_d(forced_err)), 0 if forced_h is None else len(forced_h))
    n = min(ntr.value, trace_cap)
    stats = {f: getattr(st, f) for f, _ in HjbStats._fields_}
    return phi, stats, th[:n].copy(), te[:n].copy()


def fill_field(phi_cols, Ny, Nx, dx=0.05, dy=0.05, mu=5.0, lim=10e-3):
    nt = phi_cols.shape[0]
    vx = np.empty((nt - 1, Ny - 2, Nx - 2)); vy = np.empty_like(vx)
    lib().oco_fill_field(_p(_d(phi_cols)), nt, Ny, Nx, C.c_double(dx), C.c_double(dy), C.c_double(mu),
                         C.c_double(lim), _p(vx), _p(vy))
    return vx, vy


def vels(phi, Ny, Nx, dx=0.05, dy=0.05, mu=5.0, lim=10e-3):
    vx = np.empty((Ny - 2, Nx - 2)); vy = np.empty_like(vx)
    lib().oco_vels(_p(_d(phi)), Ny, Nx, C.c_double(dx), C.c_double(dy), C.c_double(mu), C.c_double(lim), _p(vx), _p(vy))
    return vx, vy


def math_fn(name, *args):
    args = [_d(a) for a in args]
    out = np.empty_like(args[0])
    getattr(lib(), "oco_math_" + name)(*[_p(a) for a in args], _p(out), args[0].size)
    return out


class KeyData:
    """Keeps numpy buffers alive behind an oco_key struct."""

    def __init__(self, V, vx_opt, vy_opt, nt_opt, doors):
        self.V = _d(V); self.vx = _d(vx_opt); self.vy = _d(vy_opt)
        self.doors = _d(np.asarray(doors, dtype=np.float64).reshape(-1, 4))
        self.nt_opt = int(nt_opt)

    def struct(self):
        return Key(_p(self.V), _p(self.vx), _p(self.vy), self.nt_opt, self.vx.shape[0], _p(self.doors),
                   self.doors.shape[0])


def choose_velocity(params, key: KeyData, x, y, t):
    ox, oy = C.c_double(), C.c_double()
    ks = key.struct()
    bad = lib().oco_choose_velocity(C.byref(params), C.byref(ks), C.c_double(x), C.c_double(y), int(t),
                                    C.byref(ox), C.byref(oy))
    return ox.value, oy.value, bad


def pair_force(params, pi, vi, vdes, pj, vj):
    fx, fy = C.c_double(), C.c_double()
    lib().oco_pair_force(C.byref(params), *[C.c_double(float(v)) for v in (pi[0], pi[1], vi[0], vi[1], vdes, pj[0],
                                                                         pj[1], vj[0], vj[1])], C.byref(fx), C.byref(fy))
    return fx.value, fy.value


def wall_force(params, X, Y, V, p, v, vdes):
    fx, fy, ind = C.c_double(), C.c_double(), C.c_long()
    X, Y, V = _d(X), _d(Y), _d(V)
    lib().oco_wall_force(C.byref(params), _p(X), _p(Y), _p(V), *[C.c_double(float(q)) for q in
                                                                  (p[0], p[1], v[0], v[1], vdes)],
                         C.byref(fx), C.byref(fy), C.byref(ind))
    return fx.value, fy.value, ind.value


def gcfm_step(params, state, v_des, key_id, keys, X, Y, perm, noise, simu_step):
    """state: dict of x,y,vx,vy,time (float64 (N,)) and status (uint8 (N,)) -- updated IN PLACE.
    Returns (exit_log array, bad flag, wall_ind)."""
    N = state["x"].shape[0]
    karr = (Key * len(keys))(*[k.struct() for k in keys])
    perm = np.ascontiguousarray(perm, dtype=np.int32)
    key_id = np.ascontiguousarray(key_id, dtype=np.int32)
    noise = _d(noise); v_des = _d(v_des); X = _d(X); Y = _d(Y)
    exit_log = np.empty(max(N, 1), dtype=np.int32)
    n_exit = C.c_int()
    wall_ind = np.full(N, -1, dtype=np.int64)
    for k in ("x", "y", "vx", "vy", "time"):
        assert state[k].dtype == np.float64 and state[k].flags.c_contiguous
    assert state["status"].dtype == np.uint8
    bad = lib().oco_gcfm_step(C.byref(params), N, _p(state["x"]), _p(state["y"]), _p(state["vx"]), _p(state["vy"]),
                              _p(state["time"]), state["status"].ctypes.data_as(C.POINTER(C.c_uint8)), _p(v_des),
                              key_id.ctypes.data_as(C.POINTER(C.c_int)), karr, _p(X), _p(Y),
                              perm.ctypes.data_as(C.POINTER(C.c_int)), _p(noise), int(simu_step),
                              exit_log.ctypes.data_as(C.POINTER(C.c_int)), C.byref(n_exit),
                              wall_ind.ctypes.data_as(C.POINTER(C.c_long)))
    return exit_log[: n_exit.value].copy(), bad, wall_ind


def density(X, Y, Vglobal, x, y, status, sigma):
    X, Y, Vg = _d(X), _d(Y), _d(Vglobal)
    Ny, Nx = Vg.shape
    d = np.empty((Ny, Nx))
    Cn = float(np.sqrt(4 * np.pi ** 2 * sigma ** 2))  # simulations.py:482
    status = np.ascontiguousarray(status, dtype=np.uint8)
    lib().oco_density(_p(X), _p(Y), _p(Vg), Ny, Nx, len(x), _p(_d(x)), _p(_d(y)),
                      status.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_double(sigma), C.c_double(Cn), _p(d))
    return d


def create_potential(X, Y, walls, holes, cyls, targets):
    X, Y = _d(X), _d(Y)
    Ny, Nx = len(Y), len(X)
    V = np.empty((Ny, Nx))
    w, h, c, t = (_d(np.asarray(a, dtype=np.float64).reshape(-1, k)) for a, k in
                  ((walls, 4), (holes, 4), (cyls, 3), (targets, 4)))
    lib().oco_create_potential(_p(X), _p(Y), Ny, Nx, _p(w), len(w), _p(h), len(h), _p(c), len(c), _p(t), len(t), _p(V))
    return V
