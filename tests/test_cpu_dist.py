"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: band partition, unique-id hand-off, halo exchange
semantics and the decomposition-independent error-norm order -- with the CPU oracle standing in for the kernels."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import REPO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    import torch
    import torch.distributed as dist
    from optimal_crowds_b200 import dist as ocd
    from oracle import cpu_oracle as co
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # 1. communicator bootstrap: every rank ends up with rank 0's 128 bytes
        uid = ocd.share_unique_id(make_id=lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        # 2. band partition
        Ny, Nx = 64, 50
        own0, own1 = ocd.band_of(rank, Ny, world)
        assert own1 - own0 == Ny // world and own0 == rank * (Ny // world)
        # 3. halo exchange + stencil on the band == the same rows of the full-grid stencil (bit for bit)
        rng = np.random.RandomState(5)
        phi = np.exp(rng.normal(size=(Ny, Nx)))
        V = np.zeros((Ny, Nx)); V[0] = V[-1] = -100; V[:, 0] = V[:, -1] = -100; V[30:34, 10:40] = -100; V[20:26, -1] = 1
        m = rng.uniform(0, 1, (Ny, Nx))
        full = co.hjb_rhs(phi.ravel(), V, m, 0.05, 0.05, 0.2, 5.0, -0.005).reshape(Ny, Nx)
        band = phi[own0:own1].copy()
        up = torch.zeros(Nx, dtype=torch.float64); down = torch.zeros(Nx, dtype=torch.float64)
        reqs = []
        if rank > 0:
            reqs += [dist.isend(torch.from_numpy(band[0].copy()), rank - 1), dist.irecv(up, rank - 1)]
        if rank + 1 < world:
            reqs += [dist.isend(torch.from_numpy(band[-1].copy()), rank + 1), dist.irecv(down, rank + 1)]
        for r in reqs:
            r.wait()
        lo = 1 if rank > 0 else 0
        hi = 1 if rank + 1 < world else 0
        ext = np.vstack(([up.numpy()] if lo else []) + [band] + ([down.numpy()] if hi else []))
        sl = slice(own0 - lo, own1 + hi)
        out = co.hjb_rhs(ext.ravel(), V[sl], m[sl], 0.05, 0.05, 0.2, 5.0, -0.005).reshape(ext.shape)
        mine = out[lo:lo + (own1 - own0)]   # halo rows were evaluated against a fake boundary: dropped
        assert np.array_equal(mine, full[own0:own1])
        # 4. error norm: per-16-row block sums, gathered and added in global block order == undecomposed order
        blocks = np.array([np.sum(mine[i:i + 16] ** 2) for i in range(0, own1 - own0, 16)])
        gathered = [torch.zeros(len(blocks), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(blocks))
        total = 0.0
        for g in gathered:
            for v in g.numpy():
                total += v
        ref = 0.0
        for i in range(0, Ny, 16):
            ref += np.sum(full[i:i + 16] ** 2)
        assert total == ref
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_band_logic_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_band_partition_rules():
    from optimal_crowds_b200 import dist as ocd
    assert ocd.band_of(3, 16384, 8) == (6144, 8192)
    assert ocd.band_rows(2048, 1) == 2048
    with pytest.raises(ValueError):
        ocd.band_rows(100, 3)
    with pytest.raises(ValueError):
        ocd.band_rows(120, 4)   # 30 rows: not a multiple of 16


def _ens_worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from optimal_crowds_b200 import ensemble
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        mine = ensemble.shard(7, rank, world)
        local = {i: dict(seed=100 + i, evac_time=float(i) * 0.5, exit_order=np.arange(i)) for i in mine}
        merged = ensemble.gather_results(local)
        assert list(merged) == list(range(7))
        assert all(merged[i]["seed"] == 100 + i and len(merged[i]["exit_order"]) == i for i in merged)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_ensemble_shard_and_gather_gloo():
    """ensembles shard by member (no data-path collective); the only exchange is the final gather of results"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ens_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_ensemble_planning():
    from optimal_crowds_b200 import ensemble
    assert ensemble.shard(10, 1, 4) == [1, 5, 9]
    assert ensemble.shard(3, 3, 4) == []
    with pytest.raises(ValueError):
        ensemble.shard(3, 4, 4)
    per = ensemble.field_bytes_per_member(512, 512, 20.0, 0.02)
    assert 2.0e9 < per < 2.3e9                       # 1000 phi samples of a 512^2 room
    waves = ensemble.plan_waves(list(range(128)), per, int(140e9))
    assert sum(len(w) for w in waves) == 128 and max(len(w) for w in waves) <= 140e9 // per
    assert ensemble.plan_waves([1, 2, 3], per, 10) == [[1], [2], [3]]     # never less than one member per wave


class _FakeCtx:
    """stands in for _lib.Context in the peer-memory hand-shake (no GPU here)"""

    def __init__(self, rank, fail_export_on=None, fail_import_on=None):
        self.rank, self.Ny = rank, 64
        self.fail_export_on, self.fail_import_on = fail_export_on, fail_import_on
        self.imported, self.disabled = None, False

    def p2p_export(self, rows):
        if self.rank == self.fail_export_on:
            raise RuntimeError("no IPC here")
        return bytes([self.rank]) * 64

    def p2p_import(self, handles):
        if self.rank == self.fail_import_on:
            raise RuntimeError("cannot map peer")
        self.imported = list(handles)

    def p2p_disable(self):
        self.disabled = True


def _p2p_worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from optimal_crowds_b200 import dist as ocd
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        a = _FakeCtx(rank)
        assert ocd.enable_peer_memory(a) is True and not a.disabled
        assert a.imported == [bytes([0]) * 64, bytes([1]) * 64]          # every rank's handle, in rank order
        b = _FakeCtx(rank, fail_export_on=1)
        assert ocd.enable_peer_memory(b) is False and b.disabled         # all or nothing: both ranks fall back to NCCL
        c = _FakeCtx(rank, fail_import_on=0)
        assert ocd.enable_peer_memory(c) is False and c.disabled
        os.environ["OC_P2P"] = "0"
        d = _FakeCtx(rank)
        assert ocd.enable_peer_memory(d) is False and d.imported is None
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        os.environ.pop("OC_P2P", None)
        dist.destroy_process_group()


def test_peer_memory_handshake_is_all_or_nothing_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_p2p_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
