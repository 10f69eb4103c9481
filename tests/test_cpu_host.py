"""CPU tests of the host-side logic: library exports, grid rule, crowd placement RNG parity, sampler index
semantics, synthetic rooms."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from conftest import REPO, golden, room_grid


def test_library_loads_and_exports_every_declared_symbol():
    from optimal_crowds_b200 import _lib
    from optimal_crowds_b200.build import build_lib
    build_lib()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    hdr = open(os.path.join(REPO, "include", "optimal_crowds.h")).read()
    declared = set(re.findall(r"\b(oc_[a-z_0-9]+)\s*\(", hdr))
    assert {"oc_hjb_solve", "oc_gcfm_step", "oc_rasterise", "oc_density", "oc_ctx_create"} <= declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in optimal_crowds.h but not exported"
    assert set(_lib.EXPORTS) <= declared
    lib.oc_abi_version.restype = ctypes.c_int
    assert lib.oc_abi_version() == 2


def test_no_cpu_fallback_context_raises_without_gpu():
    import torch
    from optimal_crowds_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.Context(10.0, 6.0, 0.05)


def test_grid_rule_quirk():
    """simulations.py:63-64: 10//0.05 == 199.0 -> 200 nodes; nodes_to_length inverts the rule exactly."""
    from optimal_crowds_b200 import _lib, synthetic
    assert _lib.grid_shape(10.0, 6.0, 0.05) == (120, 200)
    for n in (512, 2048, 4096, 16384):
        L = synthetic.nodes_to_length(n)
        assert _lib.grid_shape(L, L, 0.05) == (n, n)
    assert abs(synthetic.nodes_to_length(16384) - 819.175) < 1e-9


@pytest.mark.parametrize("rname", ["room_test", "exit_opposite", "dense", "small"])
def test_crowd_placement_consumes_rng_like_reference(rname):
    """same seed -> bit-identical initial positions and v_des as the reference's full-grid rejection sampler."""
    from optimal_crowds_b200 import _crowd
    u = golden("units")
    room = json.loads(str(u[f"rast_{rname}_room"]))
    L, H, Ny, Nx, X, Y = room_grid(room)
    np.random.seed(2)
    place = np.zeros((Ny, Nx))
    xs, ys, vd = [], [], []
    for box in room["initial_boxes"].values():
        x, y, v = _crowd.place_box(box, X, Y, place)
        xs.append(x); ys.append(y); vd.append(v)
    xy = np.column_stack([np.concatenate(xs), np.concatenate(ys)])
    assert np.array_equal(xy, u[f"dens_{rname}_xy"])
    assert np.array_equal(np.concatenate(vd), u[f"init_{rname}_vdes"])


def test_sampler_axis_nodes_match_reference_index_rule():
    from optimal_crowds_b200.optimals import optimals
    L, dx, Nx = 10.0, 0.05, 200
    f = optimals._axis_nodes
    assert f(0.02, L, dx, Nx) == [0]
    assert f(0.05, L, dx, Nx) == [1]                 # x > dx is strict (optimals.py:236)
    assert f(0.050001, L, dx, Nx) == [1, 2]
    assert f(0.15000000000000002, L, dx, Nx) == [3, 4]
    assert f(9.95, L, dx, Nx) == [Nx - 3]            # x >= L - dx  (optimals.py:238-239)
    assert f(9.96, L, dx, Nx) == [Nx - 3]
    assert f(9.9, L, dx, Nx) == [int(9.9 // dx), int(9.9 // dx) + 1]   # out of range for the (Nx-2) field: IndexError later


def test_synthetic_rooms_follow_schema_and_counts():
    from optimal_crowds_b200 import synthetic
    for room, n_agents in ((synthetic.slalom_room(16384, 2048, 12500), 12500), (synthetic.metro_room(4096, 10000), 10000),
                           (synthetic.ensemble_room(512, 1000), 1000)):
        assert set(room) == {"room_length", "room_height", "initial_boxes", "targets", "walls", "holes", "cylinders"}
        n = sum(int(b[4] * b[2] * b[3]) for b in room["initial_boxes"].values())
        assert n == n_agents
        for b in room["initial_boxes"].values():
            assert all(t in room["targets"] for t in b[5:])
    r = synthetic.slalom_room(16384, 2048, 12500)
    assert len({" or ".join(b[5:]) for b in r["initial_boxes"].values()}) == 1   # one HJB key
    m = synthetic.metro_room()
    assert len({" or ".join(b[5:]) for b in m["initial_boxes"].values()}) == 4   # four HJB keys


def test_step_randomness_lookahead_is_stream_exact():
    """the look-ahead draws give the same permutations, the same normal pairs and leave the global generator in the
    same state as the reference's plain sequence of calls, whatever the exit pattern and even if other code uses
    np.random between steps."""
    from optimal_crowds_b200._rng import StepRandomness
    N = 9000
    rng_exits = np.random.RandomState(99)
    exits = [0, 0, 3, 0, 1500, 7, 0, 0, 1, 400, 0, 1000, 90, 63, 0]   # incl. more exits in one step than snapshots are kept
    # reference sequence
    np.random.seed(42)
    ref, n = [], N
    for k, e in enumerate(exits):
        perm = np.random.choice(np.arange(N), N, replace=False)
        z = np.random.normal(size=(n, 2)) if n else np.zeros((0, 2))
        if k == 6:
            extra_ref = np.random.uniform()      # foreign consumer between steps
        ref.append((perm, z))
        n -= e
    end_ref = np.random.get_state()
    # look-ahead sequence
    np.random.seed(42)
    sr = StepRandomness(lookahead=True)
    sr.MIN_N = 0
    n = N
    for k, e in enumerate(exits):
        perm, z = sr.draw(N, n)
        assert np.array_equal(perm, ref[k][0]) and np.array_equal(z, ref[k][1]), k
        sr.lookahead(N, n)                       # "while the GPU runs"; must not move the global stream
        if k == 6:
            assert np.random.uniform() == extra_ref
        n -= e
    end = np.random.get_state()
    assert end[2] == end_ref[2] and np.array_equal(end[1], end_ref[1]) and end[3:] == end_ref[3:]
    # misses: the first step, the step after the foreign draw, and the steps after >= N_CKPT (64) exits (4 of them)
    assert sr.hits == len(exits) - 6 and sr.misses == 6


def test_member_rng_is_the_seeded_global_stream():
    """ensemble members draw from np.random.RandomState(seed): the very stream np.random.seed(seed) gives the
    reference (crowd placement: simulations.py:130-140; per-step draws: :271,303)"""
    from optimal_crowds_b200 import _crowd, _rng
    X, Y = np.linspace(0, 10, 200), np.linspace(0, 6, 120)
    box = [2.0, 3.0, 2.0, 3.0, 2.0, "door"]
    np.random.seed(41)
    a = _crowd.place_box(box, X, Y, np.zeros((120, 200)))
    sr = _rng.StepRandomness(lookahead=False)
    pa, za = sr.draw(12, 7)
    rs = np.random.RandomState(41)
    b = _crowd.place_box(box, X, Y, np.zeros((120, 200)), rng=rs)
    pb, zb = _rng.StepRandomness(lookahead=False, rng=rs).draw(12, 7)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    assert np.array_equal(pa, pb) and np.array_equal(za, zb)
    # and it equals the reference's literal calls
    np.random.seed(41)
    _crowd.place_box(box, X, Y, np.zeros((120, 200)))
    assert np.array_equal(np.random.choice(np.arange(12), 12, replace=False), pa)
    assert np.array_equal(np.array([np.random.normal(size=2) for _ in range(7)]), za)


def test_history_frame_is_lazy():
    from optimal_crowds_b200.simulations import _Frame
    calls = []
    f = _Frame([["p0", "v0", "door", 1.3], ["p1", "v1", "door", 1.2]], lambda: calls.append(1) or np.ones((3, 4)))
    assert len(f) == 3 and f[0][2] == "door" and f[1][3] == 1.2 and not calls
    assert [a[0] for a in f[:-1]] == ["p0", "p1"] and not calls        # agents only: nothing computed
    assert f[-1].shape == (3, 4) and calls == [1]
    assert f[2] is f[-1] and calls == [1]                              # computed once
    g = _Frame([], lambda: np.zeros((2, 2)))
    assert [type(x) for x in g] == [np.ndarray]                        # iteration materialises the density
    rows = []
    h = _Frame(lambda: rows.append(1) or [["p", "v", "door", 1.0]], lambda: calls.append(2) or np.zeros(1))
    assert not rows                                                    # the agents' rows are lazy too (device record)
    assert len(h) == 2 and rows == [1] and h[0][0] == "p" and rows == [1] and calls == [1]


def test_c_crowd_placement_is_bit_identical_to_the_numpy_restatement():
    """oc_place_box (C, liboc_b200.so) == _crowd.place_box_py (numpy, itself pinned to the reference by the golden
    crowds): same positions, same occupancy mask, same MT19937 state afterwards -- for a RandomState and for the global
    np.random module (simulations.py:122-140)."""
    from optimal_crowds_b200 import _crowd, _lib, synthetic
    for room, seed in ((synthetic.ensemble_room(512, 400), 3), (synthetic.metro_room(1024, 800), 11)):
        Ny, Nx = _lib.grid_shape(room["room_length"], room["room_height"], 0.05)
        X1, Y1 = np.linspace(0, room["room_length"], Nx), np.linspace(0, room["room_height"], Ny)
        got = []
        for fn in (_crowd.place_box_py, _crowd.place_box):
            pp, rs = np.zeros((Ny, Nx)), np.random.RandomState(seed)
            res = [fn(b, X1, Y1, pp, 0.2, rs) for b in room["initial_boxes"].values()]
            got.append((res, pp, rs.get_state()))
        for ra, rb in zip(got[0][0], got[1][0]):
            for a, b in zip(ra, rb):
                assert np.array_equal(a, b)
        assert np.array_equal(got[0][1], got[1][1])
        sa, sb = got[0][2], got[1][2]
        assert np.array_equal(sa[1], sb[1]) and sa[2:] == sb[2:]
    room = synthetic.ensemble_room(512, 200)
    Ny, Nx = _lib.grid_shape(room["room_length"], room["room_height"], 0.05)
    X1, Y1 = np.linspace(0, room["room_length"], Nx), np.linspace(0, room["room_height"], Ny)
    box = list(room["initial_boxes"].values())[0]
    st0 = np.random.get_state()
    try:
        np.random.seed(9); a = _crowd.place_box(box, X1, Y1, np.zeros((Ny, Nx)), 0.2, np.random); s1 = np.random.get_state()
        np.random.seed(9); b = _crowd.place_box_py(box, X1, Y1, np.zeros((Ny, Nx)), 0.2, np.random); s2 = np.random.get_state()
    finally:
        np.random.set_state(st0)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and np.array_equal(s1[1], s2[1]) and s1[2:] == s2[2:]


def test_c_legacy_rng_equals_numpy_draws_and_state():
    """oc_rng_step_draw_ckpt (csrc/oc_rng.h: MT19937, masked-rejection intervals, Fisher-Yates from the top, polar
    Box-Muller with its cached value; transform threaded for large counts) == RandomState.choice(arange(N), N, False) +
    normal(size=(n, 2)), bit for bit, incl. the generator state afterwards and the look-ahead snapshots"""
    import ctypes as C
    from optimal_crowds_b200 import _lib
    lib = _lib.load()
    U = C.POINTER(C.c_uint32)
    for seed in range(24):
        rs = np.random.RandomState(seed)
        if seed % 3 == 0:
            rs.normal(size=3)                       # leaves a cached gaussian in the state
        N = int(np.random.RandomState(seed + 99).randint(1, 30000 if seed % 6 == 0 else 2500))
        na = int(N * 0.8)
        st = rs.get_state()
        key = st[1].copy(); pos = C.c_int(st[2]); hg = C.c_int(st[3]); cg = C.c_double(st[4])
        perm = np.empty(N, dtype=np.int32); noise = np.empty((na, 2))
        nck = min(40, na + 1)
        ck = np.zeros((nck, 624), dtype=np.uint32); cp = np.zeros(nck, dtype=np.int32)
        ch = np.zeros(nck, dtype=np.int32); cc = np.zeros(nck)
        rc = lib.oc_rng_step_draw_ckpt(key.ctypes.data_as(U), C.byref(pos), C.byref(hg), C.byref(cg), N, na,
                                       perm.ctypes.data_as(_lib.ip), _lib._hp(noise), nck, ck.ctypes.data_as(U),
                                       cp.ctypes.data_as(_lib.ip), ch.ctypes.data_as(_lib.ip), _lib._hp(cc))
        ref = np.random.RandomState(); ref.set_state(st)
        p2 = ref.choice(np.arange(N), N, replace=False); n2 = ref.normal(size=(na, 2)); s2 = ref.get_state()
        assert rc == 0 and np.array_equal(perm, p2) and np.array_equal(noise, n2)
        assert np.array_equal(key, s2[1]) and (pos.value, hg.value, cg.value) == (s2[2], s2[3], s2[4])
        for q in (0, 1, nck - 1):                   # snapshot q = the state after na - q pairs
            r3 = np.random.RandomState(); r3.set_state(st)
            r3.choice(np.arange(N), N, replace=False); r3.normal(size=(na - q, 2)); s3 = r3.get_state()
            assert np.array_equal(ck[q], s3[1]) and (int(cp[q]), int(ch[q]), float(cc[q])) == (s3[2], s3[3], s3[4])


def test_lazy_agents_container_behaves_like_the_object_array():
    """simulation.agents builds its ped views on access (simulations.py:142-151 builds them eagerly): int / negative /
    slice / fancy / boolean indexing, iteration, len, np.asarray, identity of repeated accesses, IndexError"""
    from optimal_crowds_b200.simulations import _Agents
    made = []

    def make(i):
        made.append(i)
        return ("ped", i)
    a = _Agents(7, make)
    assert len(a) == 7 and a.shape == (7,) and a.size == 7 and made == []
    assert a[2] == ("ped", 2) and a[-1] == ("ped", 6) and a[2] is a[2] and made == [2, 6]
    assert list(a[1:4]) == [("ped", 1), ("ped", 2), ("ped", 3)] and isinstance(a[1:4], np.ndarray)
    assert list(a[np.array([5, 0])]) == [("ped", 5), ("ped", 0)]
    assert list(a[np.arange(7) % 3 == 0]) == [("ped", 0), ("ped", 3), ("ped", 6)]
    assert [p[1] for p in a] == list(range(7)) and sorted(made) == list(range(7))      # every agent built exactly once
    arr = np.asarray(a)
    assert arr.dtype == object and arr.shape == (7,) and arr[4] is a[4]
    with pytest.raises(IndexError):
        a[7]
    assert len(_Agents(0, make)) == 0 and list(_Agents(0, make)) == []


@pytest.mark.parametrize("T,recompute,draw", [(2.0, False, False), (3.1, True, False), (1.0, True, True), (0.05, False, False)])
def test_run_blocks_follow_the_reference_loop(T, recompute, draw, monkeypatch):
    """run() hands blocks of iterations to advance() (simulations.py:427-441 runs them one by one): the blocks must end
    exactly where the reference loop re-solves (every recompute_frequency steps) or draws (every 10th step) and at the
    end of the horizon, with the time accumulated by repeated `time += dt` as in the reference"""
    import sys
    import types
    from optimal_crowds_b200 import simulations
    plt = types.SimpleNamespace(show=lambda: None)
    monkeypatch.setitem(sys.modules, "matplotlib", types.SimpleNamespace(pyplot=plt))
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    s = object.__new__(simulations.simulation)
    s.inside, s.time, s.T, s.dt, s.simu_step = 5, 0.0, T, 0.02, 0
    s.recompute, s.recompute_step = recompute, 50
    events = []
    s._solve_all = lambda: events.append(("solve", s.simu_step))
    s.draw = lambda mode: events.append(("draw", s.simu_step))

    def advance(n, verbose=False):
        assert n >= 1
        events.append(("block", s.simu_step, n))
        for _ in range(n):
            s.time += s.dt
            s.simu_step += 1
    s.advance = advance
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        s.run(draw=draw)
    # the reference loop
    ref, t, k = [("solve", 0)], 0.0, 0
    while t < T:
        if recompute and k % 50 == 0 and k > 0:
            ref.append(("solve", k))
        t += 0.02
        k += 1
        if draw and k % 10 == 0:
            ref.append(("draw", k))
    assert s.simu_step == k and s.time == t
    assert [e for e in events if e[0] != "block"] == ref
    blocks = [e for e in events if e[0] == "block"]
    assert sum(b[2] for b in blocks) == k and all(b[2] <= 50 for b in blocks if recompute)


def test_c_legacy_rng_is_exact_under_concurrent_draws():
    """four threads draw large steps at the same time: one of them gets the transform pool (oc_rng.h TransformPool, parked
    helper threads), the others find it busy and transform inline -- every stream still equals numpy's"""
    import threading
    from optimal_crowds_b200 import _rng
    N, out = 20000, {}

    def work(seed):
        st = np.random.RandomState(seed).get_state()
        res = []
        for rep in range(3):
            perm, Z, st, _ = _rng._c_draw(st, N, N - 7 * rep, 8)
            res.append((perm, Z))
        out[seed] = (res, st)
    ths = [threading.Thread(target=work, args=(s,)) for s in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for seed in range(4):
        ref = np.random.RandomState(seed)
        for rep, (perm, Z) in enumerate(out[seed][0]):
            assert np.array_equal(perm, ref.choice(np.arange(N), N, replace=False))
            assert np.array_equal(Z, ref.normal(size=(N - 7 * rep, 2)))
        s2, s1 = ref.get_state(), out[seed][1]
        assert np.array_equal(s1[1], s2[1]) and tuple(s1[2:]) == tuple(s2[2:])
