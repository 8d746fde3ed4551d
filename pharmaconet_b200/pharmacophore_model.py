"""`PharmacophoreModel` with the reference's public surface (src/pmnet/pharmacophore_model.py:51-204) on top of
the B200 scoring path.

Same artefacts: `.pm` (pickle of a plain dict) and `.json` files written by either implementation load in the
other (state layout: SURVEY.md appendix F / pharmacophore_model.py:178-204). Same attributes: `pdbblock`,
`nodes`, `edges`, `node_dict`, `node_cluster_dict`, `node_clusters`. Scoring differs in one way only: it runs
on the GPU through the C-ABI (`scoring.score_batch`), one ligand or a whole library per call; there is no CPU
path here.
"""

from __future__ import annotations

import json
import os
import pickle
from collections.abc import Iterable, Sequence
from pathlib import Path

import numpy as np

from .constants import INTERACTION_TO_PHARMACOPHORE
from .ligand import TypedLigand
from .packing import LigandBatch, PackedModel


class ModelNode:
    """One pharmacophore point of the protein pocket (pharmacophore_model.py:249-320)."""

    __slots__ = (
        "graph", "index", "type", "interaction_type", "hotspot_position", "score", "center", "radius",
        "_neighbor_edge_dict", "_overlapped_nodes", "neighbor_edge_dict", "overlapped_nodes",
    )  # fmt: skip

    def __init__(self, graph, index, type, interaction_type, hotspot_position, score, center, radius,
                 neighbor_edge_dict, overlapped_nodes):  # fmt: skip
        self.graph = graph
        self.index = int(index)
        self.type = type
        self.interaction_type = interaction_type
        self.hotspot_position = hotspot_position
        self.score = score
        self.center = center
        self.radius = radius
        self._neighbor_edge_dict = neighbor_edge_dict
        self._overlapped_nodes = overlapped_nodes
        self.neighbor_edge_dict = {}
        self.overlapped_nodes = []

    def setup(self):
        # JSON turns the int keys into str (pharmacophore_model.py:277-284)
        g = self.graph
        self.neighbor_edge_dict = {g.nodes[int(n)]: g.edges[int(e)] for n, e in self._neighbor_edge_dict.items()}
        self.overlapped_nodes = [g.nodes[int(n)] for n in self._overlapped_nodes]

    def __hash__(self):
        return self.index

    def __eq__(self, other):
        return self is other

    def __repr__(self):
        return f"ModelNode({self.index})[{self.type}]"

    def get_kwargs(self):
        return dict(
            index=self.index,
            type=self.type,
            interaction_type=self.interaction_type,
            hotspot_position=self.hotspot_position,
            score=self.score,
            center=self.center,
            radius=self.radius,
            neighbor_edge_dict=self._neighbor_edge_dict,
            overlapped_nodes=self._overlapped_nodes,
        )


class ModelEdge:
    """Distance statistics between two model nodes, self loops included (pharmacophore_model.py:323-365)."""

    __slots__ = ("graph", "index", "nodes", "node_indices", "type", "distance_mean", "distance_std")

    def __init__(self, graph, index, node_indices, edge_type, distance_mean, distance_std):
        self.graph = graph
        self.index = int(index)
        self.node_indices = node_indices
        self.nodes = (graph.nodes[node_indices[0]], graph.nodes[node_indices[1]])
        self.type = edge_type
        self.distance_mean = distance_mean
        self.distance_std = distance_std

    def __hash__(self):
        return self.index

    def __eq__(self, other):
        return self is other

    def get_kwargs(self):
        return dict(
            index=self.index,
            node_indices=self.node_indices,
            edge_type=self.type,
            distance_mean=self.distance_mean,
            distance_std=self.distance_std,
        )


class ModelNodeCluster:
    """Group of model nodes matched as one unit (pharmacophore_model.py:207-246)."""

    __slots__ = ("type", "nodes", "node_indices", "node_types", "center", "size")

    def __init__(self, graph, cluster_type, node_indices: Iterable[int], node_types: Iterable[str], center, size):
        self.type = cluster_type
        self.node_indices = {int(i) for i in node_indices}
        self.nodes = {graph.nodes[i] for i in self.node_indices}
        self.node_types = set(node_types)
        self.center = center
        self.size = size

    def __repr__(self):
        return f"ModelCluster({self.type})[{sorted(self.node_indices)}]"

    def get_kwargs(self):
        return dict(
            cluster_type=self.type,
            node_indices=tuple(self.node_indices),
            node_types=tuple(self.node_types),
            center=self.center,
            size=self.size,
        )


class PharmacophoreModel:
    def __init__(self):
        self.pdbblock: str | None = None
        self.nodes: list[ModelNode] = []
        self.edges: list[ModelEdge] = []
        self.node_dict: dict[str, list[ModelNode]] = {}
        self.node_cluster_dict: dict[str, list[ModelNodeCluster]] = {}
        self.node_clusters: list[ModelNodeCluster] = []
        self._packed: PackedModel | None = None
        self._device_models: dict = {}

    # ------------------------------------------------------------------ persistence (pharmacophore_model.py:151-204)
    def __getstate__(self):
        return dict(
            pdbblock=self.pdbblock,
            nodes=[n.get_kwargs() for n in self.nodes],
            edges=[e.get_kwargs() for e in self.edges],
            node_cluster_dict={t: [c.get_kwargs() for c in cl] for t, cl in self.node_cluster_dict.items()},
            node_dict={t: [n.index for n in nodes] for t, nodes in self.node_dict.items()},
        )

    def __setstate__(self, state):
        self.pdbblock = state.get("pdbblock")
        self.nodes = [ModelNode(self, **kw) for kw in state["nodes"]]
        self.edges = [ModelEdge(self, **kw) for kw in state["edges"]]
        for n in self.nodes:
            n.setup()
        self.node_dict = {t: [self.nodes[i] for i in idx] for t, idx in state["node_dict"].items()}
        self.node_cluster_dict = {
            t: [ModelNodeCluster(self, **kw) for kw in cl] for t, cl in state["node_cluster_dict"].items()
        }
        self.node_clusters = [c for cl in self.node_cluster_dict.values() for c in cl]
        self._packed = None
        self._device_models = {}

    def save(self, save_path: str | Path):
        ext = os.path.splitext(save_path)[-1]
        state = self.__getstate__()
        if ext == ".pm":
            with open(save_path, "wb") as f:
                pickle.dump(state, f)
        elif ext == ".json":
            with open(save_path, "w") as f:
                json.dump(state, f, indent=2)
        else:
            raise NotImplementedError(f"unsupported model extension {ext!r} (use .pm or .json)")

    @classmethod
    def load(cls, save_path: str | Path) -> "PharmacophoreModel":
        ext = os.path.splitext(save_path)[-1]
        if ext == ".pm":
            with open(save_path, "rb") as f:
                state = pickle.load(f)  # noqa: S301 - same trust model as the reference's .pm files
        elif ext == ".json":
            with open(save_path) as f:
                state = json.load(f)
        else:
            raise NotImplementedError(f"unsupported model extension {ext!r} (use .pm or .json)")
        model = cls()
        model.__setstate__(state)
        return model

    @classmethod
    def create(cls, pdbblock: str, center, hotspot_infos: list[dict], resolution: float = 0.5, size: int = 64):
        """Density maps -> model graph (pharmacophore_model.py:108-149, utils/density_map.py)."""
        from .density_map import build_model_state

        model = cls()
        model.__setstate__(build_model_state(pdbblock, center, hotspot_infos, resolution, size))
        return model

    # ------------------------------------------------------------------ packed / device forms
    @property
    def packed(self) -> PackedModel:
        if self._packed is None:
            self._packed = PackedModel.from_model(self)
        return self._packed

    def device_model(self, device="cuda"):
        import torch

        from .scoring import DeviceModel

        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        dm = self._device_models.get(dev)
        if dm is None:
            dm = self._device_models[dev] = DeviceModel(self.packed, dev)
        return dm

    # ------------------------------------------------------------------ scoring (pharmacophore_model.py:60-106)
    def scoring_batch(self, ligands, weights: dict[str, float] | None = None, device="cuda") -> np.ndarray:
        """Scores of a whole library: `ligands` is a LigandBatch or a sequence of TypedLigand. float32 [n]."""
        from .scoring import score_library

        batch = ligands if isinstance(ligands, LigandBatch) else LigandBatch.from_typed(list(ligands))
        return score_library(self.device_model(device), batch, weights)["scores"]

    def _scoring(self, ligand, weights: dict[str, float] | None = None) -> float:
        """One ligand: a TypedLigand, or any object with the reference's `Ligand.graph` (a LigandGraph)."""
        if isinstance(ligand, TypedLigand):
            batch = LigandBatch.from_typed([ligand])
        elif hasattr(ligand, "graph"):
            batch = LigandBatch.from_reference_graphs([ligand.graph])
        else:
            raise TypeError("ligand must be a TypedLigand or expose a reference-style .graph")
        from . import _abi
        from .scoring import score_library

        out = score_library(self.device_model("cuda"), batch, weights)
        if int(out["status"][0]) >= _abi.LIG_OVERFLOW:
            # the reference would compute a real score here: do not hand back a silent 0
            raise RuntimeError(
                "ligand could not be scored on the device (more than 128 conformers, or a pair table beyond the largest "
                "scratch configuration)"
            )
        return float(out["scores"][0])

    def scoring_pbmol(self, ligand_pbmol, atom_positions, conformer_axis: int | None = None, weights=None) -> float:
        from .ligand_typing import typed_ligand_from_pbmol

        return self._scoring(typed_ligand_from_pbmol(ligand_pbmol, atom_positions, conformer_axis), weights)

    def scoring_file(self, ligand_file: str | Path, weights=None, num_conformers: int | None = None) -> float:
        from .ligand_typing import typed_ligand_from_file

        return self._scoring(typed_ligand_from_file(ligand_file, num_conformers), weights)

    def scoring_smiles(self, ligand_smiles: str, num_conformers: int, weights=None) -> float:
        from .ligand_typing import typed_ligand_from_smiles

        return self._scoring(typed_ligand_from_smiles(ligand_smiles, num_conformers), weights)


def model_node_type(interaction_type: str) -> str:
    return INTERACTION_TO_PHARMACOPHORE[interaction_type]


__all__ = ["PharmacophoreModel", "ModelNode", "ModelEdge", "ModelNodeCluster"]
