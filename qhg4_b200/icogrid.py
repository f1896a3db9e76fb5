"""Synthetic grids, environments and start populations for the per-step agent update.

Host-side test/bench inputs only (numpy); nothing here runs on the hot path.

The icosahedral grid follows the node-numbering SCHEME of the reference's EQsahedron
(`icosa/EQConnectivity.cpp:344-400`, `icosa/EQsahedron.h:44-46`): the 12 icosahedron
vertices come first, then ``S`` interior nodes for each of the 30 edges, then ``S(S-1)/2``
interior nodes for each of the 20 faces, ``10(S+1)^2 + 2`` nodes in total, cell index ==
node ID, and each cell's neighbour list holds the linked node IDs in ascending order
padded with -1 (`tools_io/GridFactory.cpp:164-183,1465-1498`).  The base icosahedron's
vertex/edge/face tables are our own, so individual IDs are a relabelling of the
reference's; the layout class (scattered seam nodes first, then 20 compact triangles) is
the same, which is what matters for sharding.
"""
from __future__ import annotations

import numpy as np

MAX_NEIGH = 6  # core/SCell.h:4


def _base_icosahedron():
    phi = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array(
        [(-1, phi, 0), (1, phi, 0), (-1, -phi, 0), (1, -phi, 0),
         (0, -1, phi), (0, 1, phi), (0, -1, -phi), (0, 1, -phi),
         (phi, 0, -1), (phi, 0, 1), (-phi, 0, -1), (-phi, 0, 1)], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array(
        [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11),
         (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9),
         (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)], dtype=np.int64)
    edges = sorted({tuple(sorted((int(a), int(b)))) for tri in f for a, b in
                    ((tri[0], tri[1]), (tri[1], tri[2]), (tri[2], tri[0]))})
    return v, f, edges


def num_cells(subdiv: int) -> int:
    """Number of nodes for ``subdiv`` interior nodes per edge (icosa/EQsahedron.h:44)."""
    return 10 * (subdiv + 1) ** 2 + 2


def make_ico_grid(subdiv: int):
    """Return ``(nbr, xyz)``: ``nbr`` int32 [nCells,6] ascending, -1 padded; ``xyz`` float64 [nCells,3]."""
    S = int(subdiv)
    n = S + 1  # segments per edge
    verts, faces, edges = _base_icosahedron()
    eidx = {e: k for k, e in enumerate(edges)}
    ncell = num_cells(S)
    xyz = np.zeros((ncell, 3))
    xyz[:12] = verts
    pairs = []
    # local lattice of one face: points (a,b) with a,b>=0, a+b<=n ; p = v0 + a/n (v1-v0) + b/n (v2-v0)
    aa, bb = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    inside = (aa + bb) <= n
    loc = -np.ones((n + 1, n + 1), dtype=np.int64)
    for fi, (v0, v1, v2) in enumerate(faces):
        gid = -np.ones((n + 1, n + 1), dtype=np.int64)
        gid[0, 0], gid[n, 0], gid[0, n] = v0, v1, v2

        def edge_ids(va, vb):
            lo, hi = (va, vb) if va < vb else (vb, va)
            base = 12 + eidx[(lo, hi)] * S
            ids = base + np.arange(S)
            return ids if va == lo else ids[::-1]

        t = np.arange(1, n)
        gid[t, 0] = edge_ids(v0, v1)          # b == 0
        gid[0, t] = edge_ids(v0, v2)          # a == 0
        gid[n - t, t] = edge_ids(v1, v2)      # a + b == n
        interior = (aa >= 1) & (bb >= 1) & (aa + bb <= n - 1)
        fbase = 12 + 30 * S + fi * (S * (S - 1) // 2)
        gid[interior] = fbase + np.arange(int(interior.sum()))
        # positions (overwriting shared nodes with identical values)
        w1 = aa[inside] / n
        w2 = bb[inside] / n
        p = (1 - w1 - w2)[:, None] * verts[v0] + w1[:, None] * verts[v1] + w2[:, None] * verts[v2]
        p /= np.linalg.norm(p, axis=1, keepdims=True)
        xyz[gid[inside]] = p
        # lattice links
        m = (aa + 1 + bb) <= n
        pairs.append(np.stack([gid[:-1, :][m[:-1, :]], gid[1:, :][m[:-1, :]]], 1))          # (a,b)-(a+1,b)
        pairs.append(np.stack([gid[:, :-1][m[:, :-1]], gid[:, 1:][m[:, :-1]]], 1))          # (a,b)-(a,b+1)
        pairs.append(np.stack([gid[1:, :-1][m[:-1, :-1]], gid[:-1, 1:][m[:-1, :-1]]], 1))   # (a+1,b)-(a,b+1)
    pr = np.concatenate(pairs)
    pr = np.concatenate([pr, pr[:, ::-1]])
    key = np.unique(pr[:, 0] * ncell + pr[:, 1])
    src = key // ncell
    dst = key % ncell
    deg = np.bincount(src, minlength=ncell)
    assert deg.max() <= MAX_NEIGH and deg.min() >= 5, (deg.min(), deg.max())
    start = np.concatenate([[0], np.cumsum(deg)])[:-1]
    col = np.arange(len(src)) - start[src]
    nbr = -np.ones((ncell, MAX_NEIGH), dtype=np.int32)
    nbr[src, col] = dst  # key-sorted => ascending within each row
    return nbr, xyz


def make_torus_grid(nx: int, ny: int):
    """Six-neighbour torus (the survey's probe grid); neighbours ascending."""
    x, y = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    idx = lambda i, j: ((i % nx) * ny + (j % ny))
    nb = np.stack([idx(x + 1, y), idx(x - 1, y), idx(x, y + 1), idx(x, y - 1), idx(x + 1, y - 1), idx(x - 1, y + 1)], -1)
    nb = np.sort(nb.reshape(nx * ny, 6), axis=1).astype(np.int32)
    return nb


def synthetic_altitude(xyz: np.ndarray, seed: int = 1) -> np.ndarray:
    """Smooth field in about [-500, 3500] m with ~30 % of the cells below sea level."""
    rng = np.random.default_rng(seed)
    f = np.zeros(len(xyz))
    for k in range(1, 5):
        for _ in range(3):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            f += rng.normal() / k * np.cos(k * np.pi * (xyz @ d) + rng.uniform(0, 2 * np.pi))
    q30 = np.quantile(f, 0.30)
    f = f - q30
    alt = np.where(f < 0, 500.0 * f / max(-f.min(), 1e-9), 3500.0 * f / max(f.max(), 1e-9))
    return alt.astype(np.float64)


def synthetic_population(n_agents: int, altitude: np.ndarray, seed: int = 1, t0: float = 0.0,
                         max_age: float = 60.0, cells: np.ndarray | None = None, fertile: bool = False):
    """Agents uniform over land cells, ages U(0,max_age), gender Bernoulli(0.5).

    Returns a dict of SoA numpy arrays in the reference's field order
    (core/SPopulation.h:43-50 + populations/tut_EnvironAltPop.h:16-21).
    """
    rng = np.random.default_rng(seed)
    land = np.flatnonzero(altitude > 0) if cells is None else np.asarray(cells)
    per_cell = rng.multinomial(n_agents, np.full(len(land), 1.0 / len(land)))  # uniform over the cells, already binned
    cell = np.repeat(land.astype(np.int32), per_cell)
    age = (rng.random(n_agents, dtype=np.float32) * np.float32(max_age)).astype(np.float32)
    birth = (np.float32(t0) - age).astype(np.float32)
    gender = rng.integers(0, 2, size=n_agents).astype(np.uint8)
    life = np.ones(n_agents, dtype=np.uint32)
    if fertile:  # the state Fertility::execute (actions/Fertility.cpp:49-74, tutorial parameters) leaves behind
        life[(age > 15) & ((gender == 1) | (age < 50))] = 5
    return dict(
        cell=cell,
        id=np.arange(n_agents, dtype=np.int64),
        birth=birth,
        gender=gender,
        age=age,
        last_birth=np.full(n_agents, -10.0 if fertile else -1.0, dtype=np.float32),
        life=life,
    )


def synthetic_climate(xyz: np.ndarray, altitude: np.ndarray, seed: int = 1) -> dict:
    """Smooth synthetic fields for the arrays NPPCapacity reads (actions/NPPCapacity.cpp:138-217): latitude and
    longitude in degrees, annual mean temperature (deg C), annual rainfall (mm), base NPP (kgC/m2/y), river water and
    a coastal flag (land cell with a sea neighbour is approximated by low altitude)."""
    rng = np.random.default_rng(seed)
    lat = np.degrees(np.arcsin(np.clip(xyz[:, 2], -1, 1)))
    lon = np.degrees(np.arctan2(xyz[:, 1], xyz[:, 0]))
    d = rng.normal(size=3); d /= np.linalg.norm(d)
    wave = np.cos(3 * np.pi * (xyz @ d))
    temp = 28.0 - 0.45 * np.abs(lat) - 0.0065 * np.maximum(altitude, 0) + 2.0 * wave
    rain = np.maximum(0.0, 1800.0 * np.cos(np.radians(lat)) ** 2 + 500.0 * wave)
    npp = np.maximum(0.0, 0.9 * np.cos(np.radians(lat)) ** 1.5 + 0.25 * wave) * (altitude > 0)
    npp[(lon > 115) & (lon < 150) & (lat > -12) & (lat < 1)] *= 0.0   # the Oceania box falls back to the Miami model
    water = (rng.random(len(xyz)) < 0.05) * rng.random(len(xyz)) * (altitude > 0)
    coastal = ((altitude > 0) & (altitude < 60)).astype(np.float64)
    return {"Latitude": lat, "Longitude": lon, "AnnualMeanTemp": temp, "AnnualRainfall": rain, "BaseNPP": npp,
            "Water": water, "Coastal": coastal}
