"""Host-side plumbing of the row-decomposed HJB solve: band partition and communicator bootstrap.

One process per GPU (torchrun).  `torch.distributed` (NCCL on the GPU box, gloo in the CPU tests) is used
only to hand the 128-byte NCCL unique id from rank 0 to the other ranks; the halo exchange and the error-norm
all-gather run inside liboc_b200.so on its own communicator (oc_hjb_dist.cu)."""
from __future__ import annotations

import ctypes as C

ROW_ALIGN = 16  # bands are multiples of the 16-row tiles so that the error-norm order is decomposition independent


def band_rows(Ny: int, world: int) -> int:
    """rows per band for an (Ny, Nx) grid split over `world` ranks (equal bands, multiples of ROW_ALIGN)"""
    if Ny % world:
        raise ValueError(f"grid height {Ny} is not divisible by {world} bands")
    rows = Ny // world
    if world > 1 and rows % ROW_ALIGN:
        raise ValueError(f"band height {rows} is not a multiple of {ROW_ALIGN}")
    return rows


def band_of(rank: int, Ny: int, world: int):
    rows = band_rows(Ny, world)
    return rank * rows, (rank + 1) * rows


def nccl_unique_id() -> bytes:
    from . import _lib
    buf = C.create_string_buffer(128)
    _lib.check(_lib.load().oc_dist_unique_id(buf))
    return buf.raw


def share_unique_id(make_id=nccl_unique_id) -> bytes:
    """rank 0 creates the id, every rank returns the same 128 bytes (torch.distributed must be initialised)"""
    import torch.distributed as dist
    box = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def share_ipc_handles(mine: bytes):
    """every rank contributes its 64-byte CUDA IPC handle and gets the handles of all ranks, in rank order"""
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, bytes(mine))
    return out


def enable_peer_memory(ctx, all_agree=None) -> bool:
    """NVLink peer-memory mode of the row-band solve (halo exchange and error-norm all-gather inside the step launch).
    Collective: every rank calls it.  All or nothing: if one rank cannot export or map the band arrays (or OC_P2P=0),
    every rank stays on the NCCL exchange and False is returned."""
    import os
    import torch.distributed as dist
    world = dist.get_world_size()
    want = os.environ.get("OC_P2P", "1") != "0" and 1 < world <= 8
    mine = b""
    if want:
        try:
            mine = ctx.p2p_export(band_rows(ctx.Ny, world))
        except Exception:
            mine = b""
    handles = share_ipc_handles(mine)
    ok = all(len(h) == 64 for h in handles)
    if ok:
        try:
            ctx.p2p_import(handles)
        except Exception:
            ok = False
    votes = [None] * world
    dist.all_gather_object(votes, bool(ok))
    ok = all(votes) if all_agree is None else all_agree(votes)
    if not ok:
        ctx.p2p_disable()
    return ok


def init_comm(ctx, make_id=nccl_unique_id):
    """communicator only (key-sharded runs: no row bands, no peer-memory mapping)"""
    import torch.distributed as dist
    ctx.dist_init(share_unique_id(make_id), dist.get_rank(), dist.get_world_size())
    ctx.peer_memory = False


def init_context(ctx, make_id=nccl_unique_id, peer_memory=True):
    """create the library's communicator for `ctx` on every rank of the default process group (and, on one NVSwitch box,
    map the ranks' band arrays into each other for the peer-memory exchange)"""
    import torch.distributed as dist
    ctx.dist_init(share_unique_id(make_id), dist.get_rank(), dist.get_world_size())
    ctx.peer_memory = enable_peer_memory(ctx) if peer_memory else False
    return band_of(dist.get_rank(), ctx.Ny, dist.get_world_size())
