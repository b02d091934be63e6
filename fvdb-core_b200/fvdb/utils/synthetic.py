"""Synthetic sparse-grid generators for the BASELINE.json configurations.

Pure torch, device-agnostic (they run on the GPU for the bench and on the CPU for the oracle tests), and
deterministic for a given seed.  Each returns int32 voxel coordinates ``[n, 3]`` (unique rows).
"""

from __future__ import annotations

import math

import torch


def _unique_rows(ijk: torch.Tensor) -> torch.Tensor:
    return torch.unique(ijk.to(torch.int32), dim=0)


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def sphere_shell(target: int = 100_000, domain: int = 256, seed: int = 0, device="cpu", tol: float = 0.02) -> torch.Tensor:
    """C1: random surface points on a sphere shell in a ``domain^3`` box, voxelised with ``round`` (fVDB voxelises
    points with round, ops/BuildGridFromPoints.cu:89).  Radius 90 (scaled with the domain), Gaussian radial jitter
    sigma = 0.5 voxel; the number of sampled points is bisected until ``target`` +- tol unique voxels remain."""
    centre, radius = domain / 2.0, 90.0 * domain / 256.0
    lo, hi = target, target * 16
    best = None
    for _ in range(24):
        m = (lo + hi) // 2
        g = _gen(seed, device)
        d = torch.randn((m, 3), generator=g, device=device, dtype=torch.float32)
        d = d / d.norm(dim=1, keepdim=True).clamp_min(1e-12)
        r = radius + 0.5 * torch.randn((m, 1), generator=g, device=device, dtype=torch.float32)
        ijk = torch.round(centre + d * r).to(torch.int32)
        keep = ((ijk >= 0) & (ijk < domain)).all(dim=1)
        best = _unique_rows(ijk[keep])
        n = best.shape[0]
        if abs(n - target) <= tol * target:
            break
        if n < target:
            lo = m + 1
        else:
            hi = m - 1
    return best


def _rect(u: torch.Tensor, v: torch.Tensor, axis: int, w: int) -> torch.Tensor:
    uu, vv = torch.meshgrid(u, v, indexing="ij")
    cols = [None, None, None]
    cols[axis] = torch.full_like(uu, w)
    a, b = [d for d in range(3) if d != axis]
    cols[a], cols[b] = uu, vv
    return torch.stack([c.reshape(-1) for c in cols], dim=1)


def indoor_room(target: int = 200_000, seed: int = 0, device="cpu", dims=(400, 300, 150), patch: int = 16) -> torch.Tensor:
    """C2/C3: an 8 m x 6 m x 3 m room at 2 cm (400 x 300 x 150 lattice): floor, ceiling and four walls, one voxel
    thick, plus the shells of 10-20 random axis-aligned boxes.  Surfaces are cut into ``patch^2`` tiles; whole
    tiles are kept in random order until ``target`` voxels are reached (scanner-like partial coverage, contiguous
    surfaces, so 3^3 neighbourhoods look like real indoor scans rather than salt-and-pepper noise)."""
    g = _gen(seed, "cpu")
    X, Y, Z = dims
    ar = lambda n: torch.arange(n, dtype=torch.int32)  # noqa: E731
    surfaces = [_rect(ar(X), ar(Y), 2, 0), _rect(ar(X), ar(Y), 2, Z - 1), _rect(ar(X), ar(Z), 1, 0), _rect(ar(X), ar(Z), 1, Y - 1),
                _rect(ar(Y), ar(Z), 0, 0), _rect(ar(Y), ar(Z), 0, X - 1)]
    n_boxes = int(torch.randint(10, 21, (1,), generator=g))
    for _ in range(n_boxes):
        size = [int(torch.randint(lo, hi, (1,), generator=g)) for lo, hi in ((20, 90), (20, 90), (15, 70))]
        org = [int(torch.randint(1, dims[d] - size[d] - 1, (1,), generator=g)) for d in range(3)]
        org[2] = 1  # furniture stands on the floor
        rng = [torch.arange(org[d], org[d] + size[d], dtype=torch.int32) for d in range(3)]
        surfaces += [_rect(rng[0], rng[1], 2, org[2] + size[2] - 1), _rect(rng[0], rng[2], 1, org[1]), _rect(rng[0], rng[2], 1, org[1] + size[1] - 1),
                     _rect(rng[1], rng[2], 0, org[0]), _rect(rng[1], rng[2], 0, org[0] + size[0] - 1)]
    pts, tile_ids, base = [], [], 0
    for s, p in enumerate(surfaces):
        # tile id from the two in-plane coordinates (the constant axis has zero extent)
        ext = (p.max(dim=0).values - p.min(dim=0).values)
        a, b = [d for d in range(3) if ext[d] > 0][:2] if int((ext > 0).sum()) >= 2 else (0, 1)
        tu, tv = (p[:, a] - p[:, a].min()) // patch, (p[:, b] - p[:, b].min()) // patch
        nv = int(tv.max()) + 1
        tid = base + tu.to(torch.int64) * nv + tv.to(torch.int64)
        base = int(tid.max()) + 1
        pts.append(p)
        tile_ids.append(tid)
    pts, tile_ids = torch.cat(pts), torch.cat(tile_ids)
    order = torch.randperm(base, generator=g)
    rank = torch.empty(base, dtype=torch.int64)
    rank[order] = torch.arange(base)
    counts = torch.bincount(tile_ids, minlength=base)[order]
    keep_tiles = int(torch.searchsorted(torch.cumsum(counts, 0), torch.tensor(int(target * 1.03)))) + 1
    ijk = _unique_rows(pts[rank[tile_ids] < keep_tiles])
    if ijk.shape[0] > target * 1.05:  # overlapping box shells can overshoot; trim deterministically
        ijk = ijk[: int(target * 1.05)]
    return ijk.to(device)


def random_occupancy(bbox: int = 292, pct: float = 20.0, seed: int = 42, device="cpu") -> torch.Tensor:
    """C5: ``randperm(bbox^3)[: bbox^3 * pct / 100]`` decoded x-major (recipe of the reference's
    src/benchmarks/convolution/benchmark_sparse_conv_comparison.py:87-99)."""
    total = bbox**3
    g = _gen(seed, device)
    idx = torch.randperm(total, generator=g, device=device)[: int(total * pct / 100.0)]
    return torch.stack([idx // (bbox * bbox), (idx // bbox) % bbox, idx % bbox], dim=1).to(torch.int32)


def lidar_sweep(target: int = 1_000_000, seed: int = 0, device="cpu", voxel: float = 0.1, max_range: float = 80.0, max_sweeps: int = 256) -> torch.Tensor:
    """C4: accumulated 64-beam spinning-LiDAR sweeps (elevation -25..+3 deg, 0.1 deg azimuth) against a ground
    plane and 30-60 random boxes, voxelised at ``voxel`` m.  The sensor drives along a random heading (1.5 m
    per revolution, with pose jitter); sweeps are accumulated until ``target`` unique voxels are reached.  Scene
    parameters come from a CPU generator (deterministic per seed); the ray casting is vectorised over
    rays x boxes on ``device``."""
    g = _gen(seed, "cpu")
    n_boxes = int(torch.randint(30, 61, (1,), generator=g))
    centres = (torch.rand((n_boxes, 2), generator=g) - 0.5) * 2 * (max_range * 0.8)
    half = torch.rand((n_boxes, 3), generator=g) * torch.tensor([4.0, 4.0, 3.0]) + torch.tensor([1.0, 1.0, 1.0])
    lo = torch.cat([centres - half[:, :2], torch.zeros(n_boxes, 1)], dim=1).to(device)          # boxes sit on the ground
    hi = torch.cat([centres + half[:, :2], 2 * half[:, 2:3]], dim=1).to(device)
    elev = torch.deg2rad(torch.linspace(-25.0, 3.0, 64))
    azim = torch.deg2rad(torch.arange(0.0, 360.0, 0.1))
    el, az = torch.meshgrid(elev, azim, indexing="ij")
    dirs = torch.stack([torch.cos(el) * torch.cos(az), torch.cos(el) * torch.sin(az), torch.sin(el)], dim=-1).reshape(-1, 3).to(device)
    heading = float(torch.rand(1, generator=g)) * 2 * math.pi
    start = (torch.rand(2, generator=g) - 0.5) * 20.0
    jitter = (torch.rand((max_sweeps, 3), generator=g) - 0.5)                                   # xy jitter (m) and yaw
    sensor_h = 1.8
    # voxel key: 3 x 21-bit biased coordinates in one int64, so accumulation is a 1-D unique
    bias = 1 << 20
    keys = torch.empty(0, dtype=torch.int64, device=device)
    for sweep in range(max_sweeps):
        along = 1.5 * sweep
        pose = torch.tensor([float(start[0]) + along * math.cos(heading) + float(jitter[sweep, 0]),
                             float(start[1]) + along * math.sin(heading) + float(jitter[sweep, 1]), sensor_h], device=device)
        yaw = heading + 0.2 * float(jitter[sweep, 2])
        rot = torch.tensor([[math.cos(yaw), -math.sin(yaw), 0.0], [math.sin(yaw), math.cos(yaw), 0.0], [0.0, 0.0, 1.0]], device=device)
        d = dirs @ rot.T
        t_hit = torch.where(d[:, 2] < -1e-6, -sensor_h / d[:, 2].clamp(max=-1e-6), torch.full_like(d[:, 2], float("inf")))  # ground z = 0
        inv = 1.0 / torch.where(d.abs() < 1e-9, torch.full_like(d, 1e-9), d)
        for b0 in range(0, n_boxes, 16):  # slab test, 16 boxes at a time: [rays, 16, 3]
            t0 = (lo[b0:b0 + 16] - pose)[None] * inv[:, None]
            t1 = (hi[b0:b0 + 16] - pose)[None] * inv[:, None]
            tmin = torch.minimum(t0, t1).max(dim=2).values
            tmax = torch.maximum(t0, t1).min(dim=2).values
            tbox = torch.where((tmax >= tmin) & (tmin > 0), tmin, torch.full_like(tmin, float("inf"))).min(dim=1).values
            t_hit = torch.minimum(t_hit, tbox)
        ok = t_hit <= max_range
        pts = pose + d[ok] * t_hit[ok, None]
        ijk = torch.floor(pts / voxel).to(torch.int64) + bias
        keys = torch.unique(torch.cat([keys, (ijk[:, 0] << 42) | (ijk[:, 1] << 21) | ijk[:, 2]]))
        if keys.numel() >= target:
            break
    if keys.numel() > target * 1.05:
        keep = torch.randperm(keys.numel(), generator=g)[:target].to(device)
        keys = keys[keep.sort().values]
    return torch.stack([(keys >> 42) - bias, ((keys >> 21) & ((1 << 21) - 1)) - bias, (keys & ((1 << 21) - 1)) - bias], dim=1).to(torch.int32).contiguous()
