"""Seeded synthetic workloads for benchmarks and profiling runs: candidate
views for the BASELINE configurations and random-init networks with the reference's architecture.
No scene data or checkpoints exist offline, so every measured run uses these generators."""
from collections import OrderedDict

import numpy as np
import torch

from .run_nerf import get_pose


def look_at_uv(position, target):
    """Pitch u / yaw v that point a camera at ``target`` (planner convention,
    plannerServer_Object/core/interface.py:71-73)."""
    dx, dy, dz = target[0] - position[0], target[1] - position[1], target[2] - position[2]
    return np.arctan2(-dy, np.sqrt(dx ** 2 + dz ** 2)), np.arctan2(dx, dz)


def circle_views(V=32, r=3.0, y=1.0, center=(0.0, 1.0, 0.0)):
    """Config 1 (cabin-scale): V poses on a circle of radius r at height y, looking at ``center``."""
    out = []
    for k in range(V):
        a = 2 * np.pi * k / V
        pos = (center[0] + r * np.cos(a), y, center[2] + r * np.sin(a))
        u, v = look_at_uv(pos, center)
        out.append([pos[0], pos[1], pos[2], u, v])
    return np.array(out)


def hemisphere_views(V=512, center=(0.0, 1.0, 0.0), rmin=3.0, rmax=4.5, seed=0):
    """Config 2 (Object scene): Fibonacci upper-hemisphere positions, radius in [rmin, rmax], look-at
    plus the planner's +-5 deg pitch / +-10 deg yaw fan (plannerServer_Object/core/vpp.py:154-155)."""
    rng = np.random.RandomState(seed)
    golden = np.pi * (3.0 - np.sqrt(5.0))
    du = np.linspace(-5 / 180 * np.pi, 5 / 180 * np.pi, 3)
    dv = np.linspace(-10 / 180 * np.pi, 10 / 180 * np.pi, 5)
    out = []
    for k in range(V):
        h = (k + 0.5) / V
        rad = np.sqrt(max(0.0, 1 - h * h))
        r = rmin + (rmax - rmin) * rng.rand()
        pos = (center[0] + r * rad * np.cos(golden * k), center[1] + r * h, center[2] + r * rad * np.sin(golden * k))
        u, v = look_at_uv(pos, center)
        u = float(np.clip(u + du[rng.randint(3)], -np.pi / 2, np.pi / 2))
        v = v + dv[rng.randint(5)]
        out.append([pos[0], pos[1], pos[2], u, v])
    return np.array(out)


def room_views(V=4096, seed=0, aoi=((-1.5, 1.0), (0.7, 1.9), (0.0, 2.5)), dirs_per_loc=4):
    """Config 3 (Room scene): positions uniform in the area of interest
    (plannerServer_Room/core/tsdf_gpu.py:25), ``dirs_per_loc`` directions each from the Room
    direction fan (plannerServer_Room/core/vpp.py:151-154)."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(V // dirs_per_loc):
        x, y, z = rng.uniform(*aoi[0]), rng.uniform(*aoi[1]), rng.uniform(*aoi[2])
        u_c = np.arctan2(-(1.0 - y), 0.6)
        ufan = np.linspace(u_c - np.pi / 6, u_c + np.pi / 6, 5)
        vfan = np.linspace(0, 2 * np.pi, 12)
        for _d in range(dirs_per_loc):
            out.append([x, y, z, float(np.clip(ufan[rng.randint(5)], -np.pi / 2, np.pi / 2)), float(vfan[rng.randint(12)])])
    return np.array(out)


def views_to_poses(views):
    return [get_pose(list(v[0:3]), v[3], v[4]) for v in views]


def init_nerf_state_dict(seed=None, D=8, W=256, input_ch=63, input_ch_views=27, skips=(4,)):
    """Random-init state_dict with the reference NeRF's keys and shapes (run_nerf_helpers.py:79-106):
    nn.Linear default init, layers created in the reference's order, so one torch.manual_seed
    reproduces the weights the reference's constructor would draw."""
    if seed is not None:
        torch.manual_seed(seed)
    lin = torch.nn.Linear
    pts = [lin(input_ch, W)] + [lin(W + input_ch, W) if i in skips else lin(W, W) for i in range(D - 1)]
    named = [(f"pts_linears.{i}", l) for i, l in enumerate(pts)]
    named += [("views_linears.0", lin(input_ch_views + W, W // 2)), ("feature_linear", lin(W, W)),
              ("alpha_linear", lin(W, 1)), ("rgb_linear", lin(W // 2, 3)), ("uncertainty_linear", lin(W // 2, 1))]
    sd = OrderedDict()
    for name, l in named:
        sd[name + ".weight"] = l.weight.detach().clone()
        sd[name + ".bias"] = l.bias.detach().clone()
    return sd


def room_occupancy_bitfield(G=128, aabb_min=(-5.0, 0.01, -5.0), aabb_max=(5.0, 5.01, 6.0)):
    """Config 3 occupancy grid: a synthetic room shell (walls, floor and ceiling of a 5 x 2.5 x 5 m box, one voxel
    thick) plus a 1 m cube, rasterised at G^3 over the scene box.  uint8 bitfield, cell (x*G+y)*G+z -> bit
    (cell & 7) of byte cell >> 3."""
    amin, amax = np.asarray(aabb_min, dtype=np.float64), np.asarray(aabb_max, dtype=np.float64)
    c = (np.arange(G) + 0.5) / G
    X, Y, Z = np.meshgrid(amin[0] + c * (amax[0] - amin[0]), amin[1] + c * (amax[1] - amin[1]),
                          amin[2] + c * (amax[2] - amin[2]), indexing="ij")
    vox = (amax - amin) / G
    lo, hi = np.array([-2.5, 0.01, -2.0]), np.array([2.5, 2.51, 3.0])
    inside = (X > lo[0]) & (X < hi[0]) & (Y > lo[1]) & (Y < hi[1]) & (Z > lo[2]) & (Z < hi[2])
    core = (X > lo[0] + vox[0]) & (X < hi[0] - vox[0]) & (Y > lo[1] + vox[1]) & (Y < hi[1] - vox[1]) & \
           (Z > lo[2] + vox[2]) & (Z < hi[2] - vox[2])
    occ = (inside & ~core) | ((np.abs(X) < 0.5) & (Y < 1.0) & (np.abs(Z - 0.5) < 0.5))
    return np.packbits(occ.reshape(-1), bitorder="little")


def state_volume(dims, seed=0):
    """Synthetic planner state map for benchmarks: State_vol_origin (0 unknown, 1 empty, 2 occupied) and Nray_vol in the
    layout evpp_b200.TSDF.set_state takes.  An observed ball around the scene centre, an occupied box shell, a floor and
    sparse specks; voxel [0,0,0] is empty as in the reference (tsdf_gpu.py:779)."""
    rng = np.random.RandomState(seed)
    X, Y, Z = [int(v) for v in dims]
    gx, gy, gz = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
    state = np.zeros((X, Y, Z), dtype=np.float32)
    c = np.array([X / 2, Y / 3, Z / 2])
    state[np.sqrt((gx - c[0]) ** 2 + (gy - c[1]) ** 2 + (gz - c[2]) ** 2) < 0.38 * min(X, Z)] = 1
    box = (np.abs(gx - c[0]) < 10) & (gy < 25) & (np.abs(gz - c[2]) < 15)
    inner = (np.abs(gx - c[0]) < 9) & (gy < 24) & (np.abs(gz - c[2]) < 14)
    state[box & ~inner] = 2
    state[(gy <= 1) & (state == 1)] = 2
    state[rng.rand(X, Y, Z) < 0.002] = 2
    state[0, 0, 0] = 1
    nray = np.zeros_like(state)
    nray[state == 2] = rng.randint(0, 40, size=int((state == 2).sum())).astype(np.float32)
    return state, nray
