"""Density maps -> pharmacophore model state (host side, runs once per pocket).

What `PharmacophoreModel.create` needs (src/pmnet/pharmacophore_model.py:108-149) from the reference's
`DensityMapGraph` (src/pmnet/utils/density_map.py): every hotspot map is split into 26-connected components of at
least 8 voxels, each becomes a node (score-weighted centre in Angstrom as fp32, radius from the voxel count), nodes
get a complete edge table including self loops (mean = centre distance, std = sqrt(r1^2 + r2^2)), and nodes are
grouped into clusters by type and distance. The result is returned directly as the state dict that
`PharmacophoreModel.__setstate__` loads (same layout the reference pickles).

Node numbering follows the reference's discovery order: components are seeded by `set.pop()` on the set of mask
voxels built in `np.where` order (density_map.py:92-95), which is deterministic in CPython, and grown breadth first
with the same neighbour order, so centres (fp64 weighted averages) come out bit-identical.
"""

from __future__ import annotations

import itertools
import math

import numpy as np

from .constants import INTERACTION_LIST, INTERACTION_TO_PHARMACOPHORE

OVERLAP_DISTANCE = 1.5  # density_map.py:12
CLUSTER_DISTANCE = 3.0  # density_map.py:13
MIN_VOXELS = 8  # density_map.py:60-61
_NEIGHBOURS = [d for d in itertools.product((-1, 0, 1), repeat=3) if d != (0, 0, 0)]

# (cluster name, major NCI type prefixes, minor NCI type prefix) - density_map.py:118-133
_GROUPED = (
    ("Cation", ("SaltBridge_pneg", "PiCation_pring"), "HBond"),
    ("Anion", ("SaltBridge_lneg",), "HBond"),
    ("Aromatic", ("PiStacking", "PiCation_lring"), "Hydrophobic"),
)
# (cluster name, NCI type prefix) - density_map.py:158-162
_SINGLE = (("HBond", "HBond"), ("Hydrophobic", "Hydrophobic"), ("Halogen", "XBond"))


def connected_components(mask):
    """Yield (voxels [n,3] int, values [n] float) of the 26-connected components of `mask > 0`.

    mask: the dense [D,H,W] map, or its non-zero voxels as a pair (coords int [n,3] in C order, values [n]) - what
    `PharmacoNet.create_models` brings back from the device instead of the dense maps."""
    if isinstance(mask, tuple):
        coords, vals = mask
        xs, ys, zs = coords[:, 0], coords[:, 1], coords[:, 2]
        mask = {(int(x), int(y), int(z)): v for x, y, z, v in zip(xs, ys, zs, vals)}
    else:
        xs, ys, zs = np.where(mask > 0.0)
    todo = {(int(x), int(y), int(z)) for x, y, z in zip(xs, ys, zs)}
    while todo:
        seed = todo.pop()
        comp = [seed]
        head = 0
        while head < len(comp):
            x, y, z = comp[head]
            head += 1
            for dx, dy, dz in _NEIGHBOURS:
                q = (x + dx, y + dy, z + dz)
                if q in todo:
                    todo.remove(q)
                    comp.append(q)
        vox = np.array(comp)
        yield vox, np.array([float(mask[p]) for p in comp])


def _position(coords, center, resolution: float, size: int) -> tuple[float, float, float]:
    half = resolution * (size - 1) / 2
    return tuple(float((c - half) + g * resolution) for c, g in zip(center, coords))


def build_model_state(pdbblock: str, center, hotspot_infos: list[dict], resolution: float = 0.5, size: int = 64) -> dict:
    assert len(center) == 3
    if not isinstance(center, tuple):
        center = tuple(np.asarray(center).tolist())

    # ---- nodes (density_map.py:50-73, 206-231)
    ntype: list[str] = []
    hotspot_pos: list[tuple] = []
    score: list[float] = []
    centers: list[np.ndarray] = []
    radius: list[float] = []
    for info in hotspot_infos:
        hp = tuple(np.asarray(info["hotspot_position"]).tolist())
        for vox, val in connected_components(info["point_map_sparse"] if "point_map_sparse" in info else info["point_map"]):
            if len(vox) < MIN_VOXELS:
                continue
            c = np.average(vox, axis=0, weights=val)
            ntype.append(info["nci_type"])
            hotspot_pos.append(hp)
            score.append(float(info["hotspot_score"]))
            centers.append(np.array(_position(c, center, resolution, size), dtype=np.float32))
            radius.append((vox.shape[0] / (4 * math.pi / 3)) ** (1 / 3) * resolution)
    n = len(ntype)

    # ---- complete edge table incl. self loops, in creation order (density_map.py:66-72, 253-278)
    edges = []
    edge_of = {}
    mean = np.zeros((n, n))
    for j in range(n):
        for i in range(j + 1):  # the new node links to every older node, last to itself
            a, b = i, j
            m = np.linalg.norm(centers[a] - centers[b]).item()
            ta, tb = ntype[a], ntype[b]
            edges.append(
                dict(
                    index=len(edges),
                    node_indices=(a, b),
                    edge_type=(min(ta, tb), max(ta, tb)),
                    distance_mean=m,
                    distance_std=math.sqrt(radius[a] ** 2 + radius[b] ** 2),
                )
            )
            edge_of[(a, b)] = edge_of[(b, a)] = edges[-1]["index"]
            mean[a, b] = mean[b, a] = m

    # neighbour / overlap lists in the order the reference fills them (add_neighbors, density_map.py:239-250)
    neigh: list[dict[int, int]] = [dict() for _ in range(n)]
    overlapped: list[list[int]] = [[] for _ in range(n)]
    for j in range(n):
        for i in range(j + 1):
            neigh[j][i] = edge_of[(i, j)]
            neigh[i][j] = edge_of[(i, j)]
            if mean[i, j] < OVERLAP_DISTANCE:
                overlapped[j].append(i)
                overlapped[i].append(j)

    # ---- clusters (density_map.py:112-181)
    cluster_lists: dict[str, list[set[int]]] = {k: [] for k in ("Cation", "Anion", "HBond", "Aromatic", "Hydrophobic", "Halogen")}
    used: set[int] = set()
    for i in range(n):
        if i in used:
            continue
        for name, major, minor in _GROUPED:
            if ntype[i].startswith(major):
                # built with the same sequence of set operations as the reference, so that iteration order (and with
                # it the fp32 summation order of the cluster centre) is the same
                members = {i}
                members.update(o for o in overlapped[i] if ntype[o].startswith(major))
                members.update(
                    k for k in range(n)
                    if ntype[k].startswith(minor) and any(mean[k, m_] < CLUSTER_DISTANCE for m_ in members)
                )  # fmt: skip
                used |= members
                cluster_lists[name].append(members)
                break
    for i in range(n):
        if i in used:
            continue
        for name, prefix in _SINGLE:
            if ntype[i].startswith(prefix):
                members = {k for k in range(n) if ntype[k].startswith(prefix) and mean[i, k] < CLUSTER_DISTANCE}
                members.add(i)
                used |= members
                cluster_lists[name].append(members)
                break

    def cluster_kwargs(name: str, members: set[int]) -> dict:
        # DensityMapNodeCluster (density_map.py:184-203): fp32 centre mean, size = max(|c_i - c| + 2 r_i)
        idx = list(members)
        pos = np.array([centers[i] for i in idx])
        rad = np.array([radius[i] * 2 for i in idx])
        c = np.mean(pos, axis=0)
        dist = np.linalg.norm(pos - c.reshape(1, 3), axis=-1) + rad
        return dict(
            cluster_type=name,
            node_indices=tuple({i for i in idx}),
            node_types=tuple({INTERACTION_TO_PHARMACOPHORE[ntype[i]] for i in idx}),
            center=tuple(c.tolist()),
            size=np.max(dist).item(),
        )

    nodes = [
        dict(
            index=i,
            type=INTERACTION_TO_PHARMACOPHORE[ntype[i]],
            interaction_type=ntype[i],
            hotspot_position=hotspot_pos[i],
            score=score[i],
            center=tuple(centers[i].tolist()),
            radius=radius[i],
            neighbor_edge_dict=neigh[i],
            overlapped_nodes=overlapped[i],
        )
        for i in range(n)
    ]
    node_dict = {t: [i for i in range(n) if ntype[i] == t] for t in INTERACTION_LIST}
    return dict(
        pdbblock=pdbblock,
        nodes=nodes,
        edges=edges,
        node_cluster_dict={k: [cluster_kwargs(k, m) for m in v] for k, v in cluster_lists.items()},
        node_dict=node_dict,
    )
