"""GPU parity of the ensemble path (BASELINE configs[4]): every member must reproduce, bit for bit, the stand-alone
run of the same room with ``np.random.seed(seed)`` -- batched HJB solve, interleaved GCFM steps and per-member
RandomState change nothing."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu


@pytest.fixture()
def in_repo_cwd(monkeypatch):
    monkeypatch.chdir(REPO)


def _alone(room, T, recompute, seed):
    from optimal_crowds_b200 import simulations
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        simu = simulations.simulation(room, T, recompute=recompute, record=False, field_storage="phi")
        simu.run()
    simu._sync_host()
    return simu


@pytest.mark.parametrize("room,T,recompute", [("room_test", 3.0, False), ("exit_opposite", 2.5, True)])
def test_members_equal_standalone_runs(room, T, recompute, in_repo_cwd):
    from optimal_crowds_b200 import ensemble
    seeds = [3, 11, 12]
    ens = ensemble.ensemble(room, T, seeds, recompute=recompute)
    res = ens.run()
    assert sorted(res) == [0, 1, 2]
    for i, seed in enumerate(seeds):
        one = _alone(room, T, recompute, seed)
        r = res[i]
        assert r["steps"] == one.simu_step and r["inside"] == one.inside and r["evac_time"] == one.time
        assert np.array_equal(r["exit_order"], np.array(one._exit_order, dtype=np.int64))
        assert np.array_equal(r["final"], one._h_now)          # positions and velocities, bit for bit
        assert np.array_equal(r["times"], one._h_timev)
        for k, o in one.targets.items():
            assert r["hjb"][k]["nfev"] == o.last_stats["nfev"]
    assert ens.stats["agent_steps"] > 0 and ens.stats["cell_updates"] > 0
    # different seeds give different crowds
    assert not np.array_equal(res[0]["final"], res[1]["final"])


def test_sharding_and_waves_do_not_change_results(in_repo_cwd):
    """two ranks' shards (run one after the other here) and one-member waves == the single-batch run"""
    from optimal_crowds_b200 import ensemble, synthetic
    room = synthetic.ensemble_room(n=96, agents=12)
    room["initial_boxes"]["box"][0] = 1.6                 # keep the 2 x 2 m box inside the 4.775 m room
    room["initial_boxes"]["box"][2:4] = [2.0, 2.0]
    room["initial_boxes"]["box"][4] = 12.5 / 4.0      # int(rho*w*h) = 12 agents, far below the packing limit
    seeds = list(range(5))
    whole = ensemble.ensemble(room, 1.0, seeds).run()
    parts = {}
    for rank in range(2):
        e = ensemble.ensemble(room, 1.0, seeds, rank=rank, world=2, max_wave=1)
        parts.update(e.run(gather=False))
        assert e.stats["waves"] == len(e.mine)
    assert sorted(parts) == sorted(whole) == list(range(5))
    for i in whole:
        assert np.array_equal(parts[i]["final"], whole[i]["final"])
        assert np.array_equal(parts[i]["exit_order"], whole[i]["exit_order"])
        assert whole[i]["N"] == 12


def test_wave_planned_chunking_matches_standalone_with_the_same_chunk_rows(in_repo_cwd):
    """chunk_rows="wave": the chunk height is planned for the wave (fewer halo rows recomputed); a member then equals the
    stand-alone run that is given the same chunk_rows"""
    from optimal_crowds_b200 import ensemble, simulations
    seeds = [5, 6, 7, 8]
    ens = ensemble.ensemble("room_test", 2.0, seeds, chunk_rows="wave")
    res = ens.run()
    assert ens.chunk_rows_used > 0
    np.random.seed(seeds[2])
    with contextlib.redirect_stdout(io.StringIO()):
        one = simulations.simulation("room_test", 2.0, record=False, chunk_rows=ens.chunk_rows_used)
        one.run()
    assert np.array_equal(res[2]["final"], one._h_now) and np.array_equal(res[2]["times"], one._h_timev)
    assert res[2]["hjb"]["door_1"]["nfev"] == one.targets["door_1"].last_stats["nfev"]


def test_batched_and_member_by_member_steps_agree(in_repo_cwd):
    """oc_gcfm_step_multi_* (one launch per kernel for the wave, randomness drawn in C from the members' MT19937 states)
    == stepping every member with oc_gcfm_step and numpy draws, bit for bit, incl. a periodic re-solve"""
    from optimal_crowds_b200 import ensemble
    seeds = [21, 22, 23, 24, 25]
    a = ensemble.ensemble("exit_opposite", 2.5, seeds, recompute=True, batched_steps=True).run()
    b = ensemble.ensemble("exit_opposite", 2.5, seeds, recompute=True, batched_steps=False).run()
    for i in range(len(seeds)):
        assert np.array_equal(a[i]["final"], b[i]["final"]) and np.array_equal(a[i]["times"], b[i]["times"])
        assert np.array_equal(a[i]["exit_order"], b[i]["exit_order"]) and a[i]["steps"] == b[i]["steps"]
