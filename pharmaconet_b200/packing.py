"""Packed (flat-array) forms of the two operands of the scoring kernel.

`PackedModel`   - the pharmacophore model tables the kernel pins in shared memory: node types, the complete
                  (mu, sigma) edge table incl. self loops (density_map.py:66-72), clusters, and the
                  model-cluster pair distance / size-sum tables of the cluster prefilter (graph_match.py:258-268).
`LigandBatch`   - a CSR library of ligands: per-ligand topology (node type masks, clusters in matcher priority
                  order) and node coordinates laid out [node][xyz][conformer] so that one warp (lane = conformer)
                  reads 128 B coalesced rows.

Both are host numpy containers; `.to(device)` gives torch tensors (device memory is torch's, plumbing only).
The layouts are documented in DESIGN.md and declared for C in include/pmnet_b200.h.
"""

from __future__ import annotations

import math
import os
from collections.abc import Iterable, Sequence
from dataclasses import dataclass

import numpy as np

from .constants import TYPE_INDEX
from .ligand import LigandTopology, TypedLigand, build_topology, node_positions

MAX_LIGAND_NODES = 255  # node ids are uint8
CONF_ALIGN = 4  # conformer stride is padded to 4 floats = 16 B


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class PackedModel:
    node_type: np.ndarray  # uint8 [Nm], index into PHARMACOPHORE_TYPES
    edge_mu: np.ndarray  # float32 [Nm, Nm], symmetric, diagonal = self loops
    edge_sigma: np.ndarray  # float32 [Nm, Nm]
    cluster_mask: np.ndarray  # uint8 [Km], bit t set iff PHARMACOPHORE_TYPES[t] in cluster.node_types
    cluster_node_off: np.ndarray  # int32 [Km+1]
    cluster_nodes: np.ndarray  # uint8 [sum], ascending node index inside each cluster
    cluster_dist: np.ndarray  # float32 [Km, Km]  ||centre_k - centre_l|| (fp64 sqrt rounded once)
    cluster_size_sum: np.ndarray  # float32 [Km, Km]  size_k + size_l

    @property
    def num_nodes(self) -> int:
        return int(self.node_type.shape[0])

    @property
    def num_clusters(self) -> int:
        return int(self.cluster_mask.shape[0])

    @property
    def edge_rsigma(self) -> np.ndarray:
        """1/sigma rounded once in fp32: the numba-compiled reference multiplies by this reciprocal
        (fastmath `arcp`; verified on the JIT's assembly, DESIGN.md section 'numerics')."""
        return (np.float32(1.0) / self.edge_sigma).astype(np.float32)

    @classmethod
    def from_model(cls, model) -> "PackedModel":
        """Pack any object with the reference's PharmacophoreModel attributes (pharmacophore_model.py:51-58):
        nodes[i].type / .index / .neighbor_edge_dict, node_clusters[k].nodes / .node_types / .center / .size."""
        nodes = list(model.nodes)
        nm = len(nodes)
        if nm > 255:
            raise ValueError("models with more than 255 nodes are not supported")
        node_type = np.array([TYPE_INDEX[n.type] for n in nodes], dtype=np.uint8)
        mu = np.zeros((nm, nm), dtype=np.float32)
        sg = np.ones((nm, nm), dtype=np.float32)
        for a in nodes:
            assert nodes[a.index] is a
            for b, e in a.neighbor_edge_dict.items():
                mu[a.index, b.index] = np.float32(e.distance_mean)
                sg[a.index, b.index] = np.float32(e.distance_std)
        clusters = list(model.node_clusters)
        km = len(clusters)
        cmask = np.zeros(km, dtype=np.uint8)
        off = [0]
        cn: list[int] = []
        for k, c in enumerate(clusters):
            for t in c.node_types:
                cmask[k] |= 1 << TYPE_INDEX[t]
            idx = sorted(int(n.index) for n in c.nodes)
            cn.extend(idx)
            off.append(len(cn))
        cdist = np.zeros((km, km), dtype=np.float32)
        csize = np.zeros((km, km), dtype=np.float32)
        for k, ck in enumerate(clusters):
            x1, y1, z1 = (float(v) for v in ck.center)
            for l, cl in enumerate(clusters):
                x2, y2, z2 = (float(v) for v in cl.center)
                # graph_match.py:258-260: math.sqrt on Python floats (fp64), later cast to fp32 by numpy
                cdist[k, l] = np.float32(math.sqrt((x1 - x2) ** 2 + (y1 - y2) ** 2 + (z1 - z2) ** 2))
                csize[k, l] = np.float32(float(ck.size) + float(cl.size))
        return cls(
            node_type=node_type,
            edge_mu=mu,
            edge_sigma=sg,
            cluster_mask=cmask,
            cluster_node_off=np.asarray(off, dtype=np.int32),
            cluster_nodes=np.asarray(cn, dtype=np.uint8),
            cluster_dist=cdist,
            cluster_size_sum=csize,
        )

    def arrays(self) -> dict[str, np.ndarray]:
        return dict(
            node_type=self.node_type,
            edge_mu=self.edge_mu,
            edge_sigma=self.edge_sigma,
            cluster_mask=self.cluster_mask,
            cluster_node_off=self.cluster_node_off,
            cluster_nodes=self.cluster_nodes,
            cluster_dist=self.cluster_dist,
            cluster_size_sum=self.cluster_size_sum,
        )

    @classmethod
    def from_arrays(cls, d) -> "PackedModel":
        return cls(**{k: np.ascontiguousarray(d[k]) for k in cls.__dataclass_fields__})


@dataclass
class LigandBatch:
    """CSR ligand library. Index arrays are int32 except coordinate offsets (int64, in floats).

    coords: for ligand i, node n, axis a, conformer c:
        coords[coord_off[i] + (n*3 + a)*conf_stride[i] + c], conf_stride = round_up(n_conf, 4)
    clusters of ligand i: lig_cluster_off[i] .. lig_cluster_off[i+1], already in matcher priority order;
    nodes of cluster q: cluster_nodes[cluster_node_off[q] .. cluster_node_off[q+1]] (ligand-local node ids,
    high-priority node first).
    """

    lig_node_off: np.ndarray  # int32 [n+1]
    lig_cluster_off: np.ndarray  # int32 [n+1]
    cluster_node_off: np.ndarray  # int32 [total clusters + 1]
    cluster_nodes: np.ndarray  # uint8
    node_type_mask: np.ndarray  # uint8 [total nodes]
    n_conf: np.ndarray  # int32 [n]
    coord_off: np.ndarray  # int64 [n+1]
    coords: np.ndarray  # float32 [coord_off[-1]]

    @property
    def num_ligands(self) -> int:
        return int(self.n_conf.shape[0])

    @property
    def num_conformers_total(self) -> int:
        return int(self.n_conf.sum())

    @property
    def max_conformers(self) -> int:
        return int(self.n_conf.max()) if self.num_ligands else 0

    def algorithmic_bytes(self) -> int:
        """Bytes one scoring pass must read/write: every packed input array once + one fp32 score per ligand
        (SURVEY.md section 8d)."""
        n = sum(a.nbytes for a in self.arrays().values())
        return int(n + 4 * self.num_ligands)

    def arrays(self) -> dict[str, np.ndarray]:
        return dict(
            lig_node_off=self.lig_node_off,
            lig_cluster_off=self.lig_cluster_off,
            cluster_node_off=self.cluster_node_off,
            cluster_nodes=self.cluster_nodes,
            node_type_mask=self.node_type_mask,
            n_conf=self.n_conf,
            coord_off=self.coord_off,
            coords=self.coords,
        )

    @classmethod
    def from_arrays(cls, d) -> "LigandBatch":
        return cls(**{k: np.ascontiguousarray(d[k]) for k in cls.__dataclass_fields__})

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_topologies(
        cls, tops: Sequence[LigandTopology], positions: Iterable[np.ndarray]
    ) -> "LigandBatch":
        """positions[i]: float32 [Nn_i, C_i, 3] node coordinates of ligand i."""
        node_off = [0]
        clu_off = [0]
        cn_off = [0]
        cn: list[int] = []
        masks: list[np.ndarray] = []
        nconf: list[int] = []
        coord_off = [0]
        chunks: list[np.ndarray] = []
        for top, pos in zip(tops, positions, strict=True):
            nn = top.num_nodes
            if nn > MAX_LIGAND_NODES:
                raise ValueError(f"ligand with {nn} pharmacophore nodes exceeds {MAX_LIGAND_NODES}")
            pos = np.asarray(pos, dtype=np.float32)
            assert pos.shape[0] == nn and pos.shape[2] == 3
            c = int(pos.shape[1])
            stride = _round_up(c, CONF_ALIGN)
            buf = np.zeros((nn, 3, stride), dtype=np.float32)
            buf[:, :, :c] = pos.transpose(0, 2, 1)
            chunks.append(buf.reshape(-1))
            coord_off.append(coord_off[-1] + buf.size)
            nconf.append(c)
            node_off.append(node_off[-1] + nn)
            masks.append(top.node_type_mask)
            for nodes in top.ordered_cluster_nodes():
                cn.extend(nodes)
                cn_off.append(len(cn))
            clu_off.append(len(cn_off) - 1)
        return cls(
            lig_node_off=np.asarray(node_off, dtype=np.int32),
            lig_cluster_off=np.asarray(clu_off, dtype=np.int32),
            cluster_node_off=np.asarray(cn_off, dtype=np.int32),
            cluster_nodes=np.asarray(cn, dtype=np.uint8),
            node_type_mask=np.concatenate(masks).astype(np.uint8) if masks else np.zeros(0, np.uint8),
            n_conf=np.asarray(nconf, dtype=np.int32),
            coord_off=np.asarray(coord_off, dtype=np.int64),
            coords=np.concatenate(chunks) if chunks else np.zeros(0, np.float32),
        )

    @classmethod
    def from_typed(cls, ligands: Sequence[TypedLigand]) -> "LigandBatch":
        tops = [build_topology(l) for l in ligands]
        return cls.from_topologies(tops, (node_positions(t, l.atom_positions) for t, l in zip(tops, ligands)))

    @classmethod
    def from_reference_graphs(cls, graphs: Sequence) -> "LigandBatch":
        """Pack the reference's own `LigandGraph` objects (ligand.py:110-259), e.g. `Ligand(...).graph`, so that a
        caller who still types ligands with OpenBabel can score them here. Only attribute reads."""
        from .constants import CLUSTER_PRIORITY

        tops, poss = [], []
        for g in graphs:
            nodes = list(g.nodes)
            mask = np.zeros(len(nodes), dtype=np.uint8)
            for n in nodes:
                for t in n.types:
                    mask[n.index] |= 1 << TYPE_INDEX[t]
            clusters = list(g.node_clusters)

            def prio(ci, clusters=clusters):
                c = clusters[ci]
                grp, rank = CLUSTER_PRIORITY[c.type]
                cnodes = c.nodes
                return (grp, -len(cnodes), rank, min(cnodes[0].atom_indices))

            order = sorted(range(len(clusters)), key=prio)
            from .ligand import LigandCluster

            lcs = []
            for c in clusters:
                lc = LigandCluster(kind=c.type)
                lc.low = [n.index for n in c.nodes]
                lcs.append(lc)
            tops.append(
                LigandTopology(
                    num_atoms=0,
                    node_type_mask=mask,
                    node_center_atoms=[],
                    node_atoms=[],
                    clusters=lcs,
                    cluster_order=order,
                )
            )
            poss.append(np.stack([np.asarray(n.positions, dtype=np.float32) for n in nodes]) if nodes else np.zeros((0, g.num_conformers, 3), np.float32))
        return cls.from_topologies(tops, poss)

    # ------------------------------------------------------------------ views
    def select(self, idx: Sequence[int]) -> "LigandBatch":
        """Sub-library with the given ligands (host copy; used by tests and overflow re-runs)."""
        idx = [int(i) for i in idx]
        node_off = [0]
        clu_off = [0]
        cn_off = [0]
        cn, masks, chunks, nconf, coord_off = [], [], [], [], [0]
        for i in idx:
            a, b = int(self.lig_node_off[i]), int(self.lig_node_off[i + 1])
            masks.append(self.node_type_mask[a:b])
            node_off.append(node_off[-1] + b - a)
            qa, qb = int(self.lig_cluster_off[i]), int(self.lig_cluster_off[i + 1])
            for q in range(qa, qb):
                s, e = int(self.cluster_node_off[q]), int(self.cluster_node_off[q + 1])
                cn.extend(self.cluster_nodes[s:e].tolist())
                cn_off.append(len(cn))
            clu_off.append(len(cn_off) - 1)
            s, e = int(self.coord_off[i]), int(self.coord_off[i + 1])
            chunks.append(self.coords[s:e])
            coord_off.append(coord_off[-1] + e - s)
            nconf.append(int(self.n_conf[i]))
        return LigandBatch(
            lig_node_off=np.asarray(node_off, dtype=np.int32),
            lig_cluster_off=np.asarray(clu_off, dtype=np.int32),
            cluster_node_off=np.asarray(cn_off, dtype=np.int32),
            cluster_nodes=np.asarray(cn, dtype=np.uint8),
            node_type_mask=np.concatenate(masks).astype(np.uint8) if masks else np.zeros(0, np.uint8),
            n_conf=np.asarray(nconf, dtype=np.int32),
            coord_off=np.asarray(coord_off, dtype=np.int64),
            coords=np.concatenate(chunks) if chunks else np.zeros(0, np.float32),
        )


def save_library(path, batch: LigandBatch, names: Sequence[str] | None = None) -> None:
    """Packed library on disk: the LigandBatch arrays + optional ligand names (one per ligand).
    `path` ending in .npz -> one file; anything else -> a DIRECTORY with one .npy per array (+ names.txt), which
    `load_library` memory-maps: a library larger than host memory is then streamed block by block from the page cache
    / the disk by `Screener.screen_host` (its blocks are slices of these arrays)."""
    path = os.fspath(path)
    if path.endswith(".npz"):
        extra = {} if names is None else {"names": np.asarray(list(names))}
        np.savez(path, **batch.arrays(), **extra)
        return
    os.makedirs(path, exist_ok=True)
    for k, v in batch.arrays().items():
        np.save(os.path.join(path, k + ".npy"), np.ascontiguousarray(v))
    if names is not None:
        with open(os.path.join(path, "names.txt"), "w") as f:
            f.write("\n".join(str(n).replace("\n", " ") for n in names))


def is_library_dir(path) -> bool:
    return os.path.isdir(path) and os.path.isfile(os.path.join(path, "coords.npy"))


def load_library(path, mmap: bool = True) -> tuple[LigandBatch, list[str]]:
    """Load a packed library written by `save_library`. A directory library is memory-mapped (copy-on-write: nothing is
    read until a block is sliced, nothing is ever written back) unless mmap=False."""
    path = os.fspath(path)
    if is_library_dir(path):
        mode = "c" if mmap else None
        arrays = {k: np.load(os.path.join(path, k + ".npy"), mmap_mode=mode) for k in LigandBatch.__dataclass_fields__}
        batch = LigandBatch.from_arrays(arrays)
        nf = os.path.join(path, "names.txt")
        if os.path.isfile(nf):
            with open(nf) as f:
                names = f.read().split("\n")
        else:
            names = [f"ligand_{i}" for i in range(batch.num_ligands)]
        return batch, names
    z = np.load(path, allow_pickle=False)
    batch = LigandBatch.from_arrays({k: z[k] for k in LigandBatch.__dataclass_fields__})
    names = [str(n) for n in z["names"]] if "names" in z.files else [f"ligand_{i}" for i in range(batch.num_ligands)]
    return batch, names
