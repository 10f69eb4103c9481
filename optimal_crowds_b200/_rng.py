"""Per-step randomness of the GCFM sweep, drawn from numpy's legacy GLOBAL generator exactly as the reference does
(simulations.py:271 ``np.random.choice(np.arange(N), N, replace=False)`` then one ``np.random.normal(size=2)`` per
agent inside at its turn, :303), with look-ahead so that the host draws while the GPU runs the previous step.

Large crowds draw through ``oc_rng_step_draw_ckpt`` (csrc/oc_rng.h: numpy's legacy MT19937 algorithms restated in C and
pinned against numpy bit for bit, the Box-Muller transform evaluated by several host threads); the generator state
travels through ``get_state`` / ``set_state``, so ``np.random`` continues exactly where the reference would be.

Why look-ahead needs care: the number of normal pairs of step k+1 is the number of agents still inside after
step k, which is only known when step k has finished.  While the GPU runs step k the host therefore draws the
next permutation and an UPPER BOUND of pairs (nobody leaves) from a COPY of the generator state, and keeps snapshots
of the state as it is after n_upper - q pairs for q = 0 .. N_CKPT-1.  When the exit count q is known, the first
n_upper - q pairs of the block are exactly the reference's draws and the generator is set to snapshot q.  The global
``np.random`` state is left untouched between steps, so any other user of ``np.random`` sees the reference's stream;
if somebody else consumed random numbers in between, or more than N_CKPT - 1 agents left in one step, the look-ahead
is discarded and the step is drawn afresh.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def _same_state(a, b) -> bool:
    return a[0] == b[0] and a[2] == b[2] and a[3] == b[3] and a[4] == b[4] and np.array_equal(a[1], b[1])


_U32P = C.POINTER(C.c_uint32)


def _c_draw(state, N, n_pairs, n_ckpt=0):
    """(perm, Z, state_after, checkpoints) from a legacy MT19937 state tuple, through the C restatement"""
    from . import _lib
    lib = _lib.load()
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos, has, cached = C.c_int(int(state[2])), C.c_int(int(state[3])), C.c_double(float(state[4]))
    perm = np.empty(N, dtype=np.int32)
    Z = np.empty((n_pairs, 2))
    ck = None
    if n_ckpt:
        ck = (np.empty((n_ckpt, 624), dtype=np.uint32), np.zeros(n_ckpt, dtype=np.int32), np.zeros(n_ckpt, dtype=np.int32),
              np.zeros(n_ckpt))
    _lib.check(lib.oc_rng_step_draw_ckpt(key.ctypes.data_as(_U32P), C.byref(pos), C.byref(has), C.byref(cached), int(N),
                                         int(n_pairs), perm.ctypes.data_as(_lib.ip), _lib._hp(Z), int(n_ckpt),
                                         ck[0].ctypes.data_as(_U32P) if ck else None,
                                         ck[1].ctypes.data_as(_lib.ip) if ck else None,
                                         ck[2].ctypes.data_as(_lib.ip) if ck else None, _lib._hp(ck[3]) if ck else None))
    return perm, Z, ("MT19937", key, pos.value, has.value, cached.value), ck


class StepRandomness:
    N_CKPT = 64     # the look-ahead survives up to N_CKPT - 1 exits in one step
    MIN_N = 2048    # below this numpy's own draws cost less than moving the generator state in and out

    _pool = None   # one helper thread for all simulations of the process (the C draw releases the GIL)

    def __init__(self, lookahead: bool = True, rng=np.random):
        # rng: the global ``np.random`` module (reference behaviour) or a ``np.random.RandomState`` of an ensemble member
        self.rng = rng
        self.enabled = lookahead
        self._spec = None
        self.hits = self.misses = 0

    def draw(self, N: int, n_active: int):
        """(perm, noise) of the step that starts now; the generator ends where the reference's would."""
        sp, self._spec = self._spec, None
        if sp is not None and "future" in sp:     # the look-ahead is drawn by the helper thread: collect it
            sp["perm"], sp["Z"], _, sp["ck"] = sp.pop("future").result()
        if N < self.MIN_N:
            perm = self.rng.choice(np.arange(N), N, replace=False)
            noise = self.rng.normal(size=(n_active, 2)) if n_active else np.zeros((0, 2))
            return perm, noise
        st = self.rng.get_state()
        if st[0] != "MT19937":  # not a legacy generator state: numpy's own draws
            perm = self.rng.choice(np.arange(N), N, replace=False)
            return perm, (self.rng.normal(size=(n_active, 2)) if n_active else np.zeros((0, 2)))
        if sp is not None and sp["N"] == N and 0 <= sp["n_spec"] - n_active < sp["n_ckpt"] and _same_state(st, sp["state0"]):
            self.hits += 1
            q = sp["n_spec"] - n_active
            ck = sp["ck"]
            self.rng.set_state(("MT19937", ck[0][q], int(ck[1][q]), int(ck[2][q]), float(ck[3][q])))
            return sp["perm"], sp["Z"][:n_active]
        self.misses += 1
        perm, Z, after, _ = _c_draw(st, N, n_active)
        self.rng.set_state(after)
        return perm, Z

    def lookahead(self, N: int, n_upper: int):
        """Pre-draw the next step (call while the GPU is busy).  Leaves the generator state unchanged."""
        if not self.enabled or N < self.MIN_N:
            return
        state0 = self.rng.get_state()
        if state0[0] != "MT19937":
            return
        n_ckpt = min(self.N_CKPT, n_upper + 1)
        if StepRandomness._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            StepRandomness._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="oc-rng")
        # drawn from a COPY of the state on the helper thread, concurrently with the rest of the step's host work and the
        # GPU sweep (ctypes releases the GIL for the C call)
        fut = StepRandomness._pool.submit(_c_draw, state0, N, n_upper, n_ckpt)
        self._spec = dict(N=N, n_spec=n_upper, state0=state0, future=fut, n_ckpt=n_ckpt)
