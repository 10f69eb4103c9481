"""Pedestrian agent -- host-side mirror of the reference's ``pedestrians.ped`` (pedestrians.py:14-356).

In the reference every agent is a Python object that owns its trajectory lists and evaluates the GCFM
forces itself with numpy.  Here the crowd lives on the GPU as struct-of-arrays (x, y, vx, vy, time,
status, v_des, key) owned by ``simulations.simulation``; a ``ped`` is a thin view of one row of that
state plus the per-step history the simulation records, so that every attribute and method the
reference exposes (``status, target, traj, vels, time, initial_position, v_des, position(), velocity(),
evac_time(), distance(), check_status(), evolve(), agents_repulsion(), wall_repulsion()``) keeps working.
The two force methods run the same device functions the sweep kernel uses (oc_pair_force /
oc_wall_force), one probe at a time -- they exist for API parity and tests, the hot path is
``simulation.step``.
"""
from __future__ import annotations

import numpy as np


class ped:
    def __init__(self, X, Y, step, V, target, all_targets, possible_targets, x, y, vx, vy, room_length, room_height,
                 v_des, a_min, tau_a, b_min, b_max, eta, eta_walls):
        # same signature as pedestrians.py:15
        self.initial_position = np.array((x, y), dtype=float)
        self.initial_velocity = np.array((vx, vy), dtype=float)
        self.v_des = v_des
        self.a_min = b_min  # the reference ignores its a_min argument (pedestrians.py:74)
        self.tau_a, self.b_min, self.b_max, self.eta, self.eta_walls = tau_a, b_min, b_max, eta, eta_walls
        self.target = target
        self.possible_targets = possible_targets
        self.all_targets = all_targets
        self.room_length, self.room_height = room_length, room_height
        self.X, self.Y, self.V, self.step = X, Y, V, step
        # standalone storage (used when the agent is not bound to a simulation)
        self._status = True
        self._time = 0
        self._traj = [self.initial_position]
        self._vels = [self.initial_velocity]
        self._sim = None
        self._idx = -1

    # -- binding to the simulation's SoA state -------------------------------------------------------
    def _bind(self, sim, idx):
        self._sim, self._idx = sim, idx

    @property
    def status(self):
        if self._sim is not None:
            return bool(self._sim._h_status[self._idx])
        return self._status

    @status.setter
    def status(self, v):
        if self._sim is not None:
            self._sim._set_status(self._idx, bool(v))
        else:
            self._status = bool(v)

    @property
    def time(self):
        if self._sim is not None:
            return self._sim._agent_time(self._idx)
        return self._time

    @property
    def traj(self):
        if self._sim is not None:
            return self._sim._agent_track(self._idx, 0)
        return self._traj

    @property
    def vels(self):
        if self._sim is not None:
            return self._sim._agent_track(self._idx, 1)
        return self._vels

    # -- reference API -----------------------------------------------------------------------------
    def position(self):
        """pedestrians.py:140-150"""
        if self._sim is not None:
            return self._sim._agent_now(self._idx)[0]
        return np.array(self._traj[-1], dtype=float)

    def velocity(self):
        """pedestrians.py:152-162"""
        if self._sim is not None:
            return self._sim._agent_now(self._idx)[1]
        return np.array(self._vels[-1], dtype=float)

    def evolve(self, x, y, vx, vy, dt):
        """pedestrians.py:166-191 (standalone agents only; bound agents are advanced by simulation.step)."""
        if self._sim is not None:
            raise RuntimeError("agents of a simulation are advanced on the GPU by simulation.step()")
        self._traj.append(np.array((x, y), dtype=float))
        self._vels.append(np.array((vx, vy), dtype=float))
        self._time += dt

    def check_status(self):
        """pedestrians.py:121-136: strict containment in any of the agent's doors."""
        x, y = self.position()
        for name in self.possible_targets:
            door = self.all_targets[name]
            if abs(x - door[0]) < door[2] * 0.5 and abs(y - door[1]) < door[3] * 0.5:
                self.status = False

    def evac_time(self):
        """pedestrians.py:195-212"""
        if self.status:
            raise ValueError('This pedestrian has not exited the room yet!')
        return self.time

    def distance(self, pos):
        """pedestrians.py:336-356"""
        x, y = self.position()
        return np.sqrt((pos[0] - x) ** 2 + (pos[1] - y) ** 2)

    # -- GCFM forces through the CUDA library --------------------------------------------------------
    def _ctx_prm(self):
        from . import _lib
        if self._sim is not None:
            return self._sim._ctx, self._sim._gcfm_prm
        if getattr(self, "_own_ctx", None) is None:
            import json
            import os
            with open(os.path.join(os.path.dirname(__file__), "config.json")) as f:
                cfg = json.load(f)
            cfg.update(b_min=self.b_min, b_max=self.b_max, tau_a=self.tau_a, eta=self.eta, eta_walls=self.eta_walls,
                       grid_step=self.step)
            self._own_ctx = _lib.Context(self.room_length, self.room_height, self.step)
            self._own_prm = _lib.gcfm_params(cfg, self.room_length, self.room_height, self._own_ctx.Ny,
                                             self._own_ctx.Nx)
        return self._own_ctx, self._own_prm

    def agents_repulsion(self, pos_j, vel_j):
        """pedestrians.py:216-280, evaluated by the device function of the sweep kernel."""
        ctx, prm = self._ctx_prm()
        dev = lambda a: ctx.to_device(np.asarray(a, dtype=np.float64).reshape(1, -1))
        f = ctx.pair_force(prm, dev(self.position()), dev(self.velocity()), dev([self.v_des]).reshape(1), dev(pos_j),
                           dev(vel_j))
        return f.cpu().numpy()[0]

    def wall_repulsion(self, X, Y, V):
        """pedestrians.py:282-334 (X, Y must be the simulation grid, as in every call the reference makes)."""
        ctx, prm = self._ctx_prm()
        p, v = self.position(), self.velocity()
        one = lambda a: ctx.to_device(np.array([a], dtype=np.float64))
        Vd = V if hasattr(V, "is_cuda") else ctx.to_device(np.asarray(V, dtype=np.float64))
        fx, fy, _ = ctx.wall_force(prm, Vd, one(p[0]), one(p[1]), one(v[0]), one(v[1]), one(self.v_des))
        return np.array((fx.item(), fy.item()), dtype=float)
