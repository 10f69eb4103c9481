"""Synthetic rooms of the shapes BASELINE.json names (the reference's preset rooms are GitHub attachments
that are not in its tree -- SURVEY.md section 0 #1 -- so stand-ins with the documented schema are generated).

All rooms use the reference's JSON schema (README.md:16-21): room_length, room_height, initial_boxes,
targets, walls, holes, cylinders.  ``nodes_to_length(n)`` gives the room extent for which the reference's
float floor-division grid rule (simulations.py:63-64) yields exactly n nodes.
"""
from __future__ import annotations

GRID_STEP = 0.05


def nodes_to_length(n: int, step: float = GRID_STEP) -> float:
    return (n - 1) * step + step / 2


def slalom_room(nx: int = 16384, ny: int = 2048, agents: int = 12500, pitch: float = 8.0, radius: float = 0.5,
                door_pitch: float = 64.0, door_size: float = 2.0, box: float = 5.0) -> dict:
    """configs[3]: cylinder field.  Cylinders of radius `radius` on a `pitch` lattice, square target doors on a
    `door_pitch` lattice (so every agent is near a target), one initial box per lattice cell, all boxes
    sharing ONE target set (=> one HJB key, as a single crowd heading for the nearest exit)."""
    L, H = nodes_to_length(nx), nodes_to_length(ny)
    cyl, targets, boxes = {}, {}, {}
    kx, ky = int(L // pitch), int(H // pitch)
    for iy in range(1, ky + 1):
        for ix in range(1, kx + 1):
            cx, cy = ix * pitch, iy * pitch
            if cx + radius < L and cy + radius < H:
                cyl[f"c_{iy}_{ix}"] = [cx, cy, radius]
    dx0, dy0 = door_pitch / 2, min(door_pitch / 2, H / 2)
    iy = 0
    while dy0 + iy * door_pitch < H - 1:
        ix = 0
        while dx0 + ix * door_pitch < L - 1:
            # doors sit in the middle of a lattice cell (offset pitch/2 from the cylinders)
            tx = round((dx0 + ix * door_pitch) / pitch) * pitch + pitch / 2
            ty = round((dy0 + iy * door_pitch) / pitch) * pitch + pitch / 2
            if tx + door_size < L and ty + door_size < H:
                targets[f"door_{iy}_{ix}"] = [tx, ty, door_size, door_size]
            ix += 1
        iy += 1
    names = list(targets)
    door_centres = {(t[0], t[1]) for t in targets.values()}
    cells = [(ix * pitch + pitch / 2, iy * pitch + pitch / 2) for iy in range(ky) for ix in range(kx)
             if (ix * pitch + pitch / 2, iy * pitch + pitch / 2) not in door_centres
             and ix * pitch + pitch / 2 + box / 2 < L - 0.5 and iy * pitch + pitch / 2 + box / 2 < H - 0.5]
    per = max(1, -(-agents // max(len(cells), 1)))
    left = agents
    for q, (cx, cy) in enumerate(cells):
        if left <= 0:
            break
        n = min(per, left)
        rho = (n + 0.5) / (box * box)  # int(rho*w*h) == n (simulations.py:122)
        boxes[f"box_{q}"] = [cx, cy, box, box, rho] + names
        left -= n
    return {"room_length": L, "room_height": H, "initial_boxes": boxes, "targets": targets, "walls": {},
            "holes": {}, "cylinders": cyl}


def metro_room(n: int = 4096, agents: int = 10000) -> dict:
    """configs[2]: station hall with a dividing wall pierced by 8 holes, 4 doors, 4 boxes with 4 distinct
    target sets (=> 4 independent HJB keys)."""
    L = H = nodes_to_length(n)
    walls = {"divider": [L / 2, H / 2, 1.0, H]}
    holes = {f"gate_{k}": [L / 2, (k + 0.5) * H / 8, 1.5, 3.0] for k in range(8)}
    targets = {"west": [0.0, H / 2, 1.0, 6.0], "east": [L, H / 2, 1.0, 6.0], "south": [L / 2 + L / 4, 0.0, 6.0, 1.0],
               "north": [L / 4, H, 6.0, 1.0]}
    per = agents // 4
    side = 60.0
    rho = (per + 0.5) / (side * side)
    boxes = {"b_sw": [L / 4, H / 4, side, side, rho, "west"],
             "b_nw": [L / 4, 3 * H / 4, side, side, rho, "north", "west"],
             "b_se": [3 * L / 4, H / 4, side, side, rho, "south", "east"],
             "b_ne": [3 * L / 4, 3 * H / 4, side, side, rho, "east"]}
    cyl = {f"pillar_{k}": [L / 2 + (k - 3.5) * 20.0, H / 2 + 30.0, 0.6] for k in range(8)}
    return {"room_length": L, "room_height": H, "initial_boxes": boxes, "targets": targets, "walls": walls,
            "holes": holes, "cylinders": cyl}


def ensemble_room(n: int = 512, agents: int = 1000) -> dict:
    """configs[4]: one member of the ensemble -- 512^2 grid, one 20 x 20 m box at 2.5 ped/m^2, one door,
    3 cylinders."""
    L = H = nodes_to_length(n)
    return {"room_length": L, "room_height": H,
            "initial_boxes": {"box": [L / 2 - 1.5, H / 2, 20.0, 20.0, (agents + 0.5) / 400.0, "door"]},
            "targets": {"door": [L, H / 2, 0.6, 2.0]}, "walls": {}, "holes": {},
            "cylinders": {"c1": [L - 2.0, H / 2 + 2.0, 0.3], "c2": [L - 2.0, H / 2 - 2.0, 0.3],
                          "c3": [L - 3.5, H / 2, 0.3]}}


def parity_room(nx: int = 2048, ny: int = 512, bands: int = 2, per_box: int = 20) -> dict:
    """Reduced room for the multi-GPU parity check of bench.py: one target set, and per row band a door with a small
    box of agents right next to it (so that agents leave within ~40 steps, in every band), a wall that straddles every
    internal band edge and a pillar -- everything a row-decomposed run can get wrong sits on or near a band edge."""
    L, H = nodes_to_length(nx), nodes_to_length(ny)
    bh = H / bands
    targets, boxes, walls, cyl = {}, {}, {}, {}
    for b in range(bands):
        yc = (b + 0.5) * bh
        for q, xc in enumerate((L * 0.25, L * 0.7)):
            targets[f"door_{b}_{q}"] = [xc, yc, 1.2, 1.2]
            cyl[f"pillar_{b}_{q}"] = [xc + 2.5, yc + 0.4, 0.3]
        if b > 0:
            walls[f"edge_{b}"] = [L / 2, b * bh, L / 3, 0.6]          # straddles the edge between bands b-1 and b
            targets[f"edge_door_{b}"] = [L * 0.9, b * bh, 1.0, 1.0]   # a door cut by the band edge
    names = list(targets)
    for b in range(bands):
        yc = (b + 0.5) * bh
        for q, xc in enumerate((L * 0.25, L * 0.7)):
            # 3 x 3 m at ~2.2 ped/m^2 (the reference's placement rule excludes ~0.4 m around an agent: it jams near 4)
            boxes[f"box_{b}_{q}"] = [xc - 1.9, yc, 3.0, 3.0, (per_box + 0.5) / 9.0] + names   # int(rho*w*h) == per_box
        if b > 0:
            boxes[f"edge_box_{b}"] = [L * 0.9 - 1.9, b * bh, 3.0, 3.0, (per_box + 0.5) / 9.0] + names
    return {"room_length": L, "room_height": H, "initial_boxes": boxes, "targets": targets, "walls": walls,
            "holes": {}, "cylinders": cyl}
