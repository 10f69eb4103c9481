"""Per-step randomness of the GCFM sweep, drawn from numpy's legacy GLOBAL generator exactly as the reference does
(simulations.py:271 ``np.random.choice(np.arange(N), N, replace=False)`` then one ``np.random.normal(size=2)`` per
agent inside at its turn, :303), with look-ahead so that the host draws while the GPU runs the previous step.

Why look-ahead needs care: the number of normal pairs of step k+1 is the number of agents still inside after
step k, which is only known when step k has finished.  While the GPU runs step k the host therefore draws the
next permutation and an UPPER BOUND of pairs (nobody leaves), recording the generator state at checkpoints near
the end of the block.  When the exit count is known, the first n pairs of the block are exactly the reference's
draws, and the generator is moved to the state right after them (nearest checkpoint + a short re-draw).  The
global ``np.random`` state is left untouched between steps, so any other user of ``np.random`` sees the
reference's stream; if somebody else consumed random numbers in between, the look-ahead is discarded.
"""
from __future__ import annotations

import numpy as np


def _same_state(a, b) -> bool:
    return a[0] == b[0] and a[2] == b[2] and a[3] == b[3] and a[4] == b[4] and np.array_equal(a[1], b[1])


class StepRandomness:
    # checkpoints (pairs before the end of the speculative block): get_state/set_state cost ~75 us each, a pair
    # ~50 ns, so a handful of checkpoints and a short re-draw beat many checkpoints
    CKPT_BEFORE_END = (4096, 1024, 256, 64)
    MIN_N = 4096  # below this the draws take less time than the bookkeeping (get_state / set_state): no look-ahead

    def __init__(self, lookahead: bool = True, rng=np.random):
        # rng: the global ``np.random`` module (reference behaviour) or a ``np.random.RandomState`` of an ensemble member
        self.rng = rng
        self.enabled = lookahead
        self._spec = None
        self.hits = self.misses = 0

    def draw(self, N: int, n_active: int):
        """(perm, noise) of the step that starts now; the global generator ends where the reference's would."""
        sp, self._spec = self._spec, None
        if sp is not None and sp["N"] == N and n_active <= sp["n_spec"] and _same_state(self.rng.get_state(), sp["state0"]):
            self.hits += 1
            count, state = max((c for c in sp["ckpt"] if c[0] <= n_active), key=lambda c: c[0])
            self.rng.set_state(state)
            if n_active > count:
                self.rng.normal(size=(n_active - count, 2))  # re-draws pairs count..n_active (already in Z)
            return sp["perm"], sp["Z"][:n_active]
        self.misses += 1
        perm = self.rng.choice(np.arange(N), N, replace=False)
        noise = self.rng.normal(size=(n_active, 2)) if n_active else np.zeros((0, 2))
        return perm, noise

    def lookahead(self, N: int, n_upper: int):
        """Pre-draw the next step (call while the GPU is busy).  Leaves the global generator state unchanged."""
        if not self.enabled or N < self.MIN_N:
            return
        state0 = self.rng.get_state()
        perm = self.rng.choice(np.arange(N), N, replace=False)
        ckpt = [(0, self.rng.get_state())]
        parts = []
        done = 0
        # small crowds: fewer checkpoints (each costs a get_state), the re-draw after the nearest one stays short
        backs = self.CKPT_BEFORE_END if n_upper >= 16 * self.CKPT_BEFORE_END[0] else (256, 32)
        for back in backs + (0,):
            stop = n_upper - back
            if stop > done:
                parts.append(self.rng.normal(size=(stop - done, 2)))
                done = stop
                ckpt.append((done, self.rng.get_state()))
        Z = np.concatenate(parts) if parts else np.zeros((0, 2))
        self.rng.set_state(state0)
        self._spec = dict(N=N, n_spec=n_upper, state0=state0, perm=perm, Z=Z, ckpt=ckpt)
