"""Oracle O1 harness: run the UNMODIFIED reference (/root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``optimal_crowds_b200/`` imports this module; it exists to
(a) generate the golden vectors committed under ``tests/golden/`` (see ``oracle/make_goldens.py``) and
(b) let CPU tests that run *in this container* compare against the live reference when it is present.
``/root/reference`` does not exist on the GPU box, so nothing marked ``gpu`` may import this.

The reference imports ``matplotlib`` and ``seaborn`` at module import (simulations.py:8-9,13,
optimals.py:7), neither of which is installed here, so no-op stubs from ``oracle/stubs`` are put on
``sys.path``.  The reference opens CWD-relative paths ``optimal_crowds/config.json`` and
``rooms/<room>.json`` (simulations.py:42,47; optimals.py:36,41), so a scratch directory holding a
symlink ``optimal_crowds -> /root/reference`` and a ``rooms/`` directory is created and made the CWD.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import shutil
import sys
import tempfile

import numpy as np

REFERENCE_DIR = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "simulations.py"))


class RefEnv:
    """Context manager: scratch CWD + stub plotting modules + imported reference modules."""

    def __init__(self, rooms: dict | None = None, rooms_dir: str | None = None, quiet: bool = True):
        self.rooms = rooms or {}
        self.rooms_dir = rooms_dir or os.path.join(_REPO, "rooms")
        self.quiet = quiet

    def __enter__(self):
        if not reference_available():
            raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
        self.scratch = tempfile.mkdtemp(prefix="oc_ref_")
        os.symlink(REFERENCE_DIR, os.path.join(self.scratch, "optimal_crowds"))
        os.makedirs(os.path.join(self.scratch, "rooms"))
        for fn in os.listdir(self.rooms_dir):
            if fn.endswith(".json"):
                shutil.copy(os.path.join(self.rooms_dir, fn), os.path.join(self.scratch, "rooms", fn))
        for name, spec in self.rooms.items():
            with open(os.path.join(self.scratch, "rooms", name + ".json"), "w") as f:
                json.dump(spec, f)
        self._cwd = os.getcwd()
        os.chdir(self.scratch)
        self._path = list(sys.path)
        sys.path.insert(0, os.path.join(_HERE, "stubs"))
        sys.path.insert(0, self.scratch)
        for m in [m for m in sys.modules if m == "optimal_crowds" or m.startswith("optimal_crowds.")]:
            del sys.modules[m]
        from optimal_crowds import simulations, optimals, pedestrians  # noqa: the reference itself
        self.simulations, self.optimals, self.pedestrians = simulations, optimals, pedestrians
        return self

    def __exit__(self, *exc):
        os.chdir(self._cwd)
        sys.path[:] = self._path
        for m in [m for m in sys.modules if m == "optimal_crowds" or m.startswith("optimal_crowds.")]:
            del sys.modules[m]
        shutil.rmtree(self.scratch, ignore_errors=True)
        return False

    @contextlib.contextmanager
    def silence(self):
        if self.quiet:
            with contextlib.redirect_stdout(io.StringIO()):
                yield
        else:
            yield


def traced_solve_ivp(trace: dict):
    """A ``solve_ivp`` wrapper that swaps 'RK45' for a subclass recording every step attempt.

    The subclass only observes (h_abs before/after, accept/reject, error norm); arithmetic is scipy's.
    """
    from scipy.integrate import solve_ivp as real_solve_ivp
    from scipy.integrate._ivp import rk as _rk

    class TracingRK45(_rk.RK45):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            trace.setdefault("h0", []).append(self.h_abs)

        def _estimate_error_norm(self, K, h, scale):
            v = super()._estimate_error_norm(K, h, scale)
            trace.setdefault("attempt_h", []).append(float(h))
            trace.setdefault("attempt_err", []).append(float(v))
            trace.setdefault("attempt_t", []).append(float(self.t))
            return v

    def wrapper(fun, t_span, y0, method="RK45", **kw):
        assert method == "RK45"
        sol = real_solve_ivp(fun, t_span, y0, method=TracingRK45, **kw)
        trace.setdefault("nfev", []).append(int(sol.nfev))
        trace.setdefault("status", []).append(int(sol.status))
        trace.setdefault("n_t", []).append(len(sol.t))
        trace["last_sol_y"] = sol.y
        return sol

    return wrapper


def agent_state(simu):
    """Snapshot (x, y, vx, vy, status, time) of every agent of a reference ``simulation``."""
    N = simu.N
    st = np.empty((N, 6))
    for i, a in enumerate(simu.agents):
        p, v = a.position(), a.velocity()
        st[i] = (p[0], p[1], v[0], v[1], 1.0 if a.status else 0.0, a.time)
    return st


def run_hjb(room: str, T: float, m=None, rooms: dict | None = None, t: float = 0.0):
    """Build the reference simulation for ``room`` (seed 0) and solve the HJB for every key.

    Returns {key: dict(V, vx_opt, vy_opt, nt_opt, trace, phi_slices)}.
    """
    out = {}
    with RefEnv(rooms) as env:
        np.random.seed(0)
        with env.silence():
            simu = env.simulations.simulation(room, T)
        for key, opt in simu.targets.items():
            trace = {}
            env.optimals.solve_ivp = traced_solve_ivp(trace)
            mm = np.zeros((simu.Ny, simu.Nx)) if m is None else m
            with env.silence():
                opt.compute_optimal_velocity(t, mm)
            out[key] = dict(V=opt.V.copy(), vx_opt=opt.vx_opt.copy(), vy_opt=opt.vy_opt.copy(),
                            nt_opt=opt.nt_opt, trace=trace, Nx=simu.Nx, Ny=simu.Ny,
                            sol_y=trace.pop("last_sol_y"))
    return out
