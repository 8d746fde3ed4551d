"""`PharmacophoreModel.create` of this package against `.pm` files written by the reference's own create + save
(tests/golden/model_*.pm, oracle/make_golden.py) from the same seeded hotspot maps."""

import os
import pickle

import numpy as np
import pytest
from golden_util import GOLDEN

from pharmaconet_b200 import synthetic
from pharmaconet_b200.packing import PackedModel
from pharmaconet_b200.pharmacophore_model import PharmacophoreModel

MODELS = {
    "syn0": dict(seed=0),
    "loose": dict(seed=5, n_hotspots=24, r_lo=2.0, r_hi=3.5),
    "sparse": dict(seed=7, n_hotspots=8, type_probs=[0.5, 0, 0, 0, 0, 0.25, 0.25, 0, 0, 0]),
    "xbond": dict(seed=11, n_hotspots=4, type_probs=[0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0]),
}


@pytest.mark.parametrize("name", list(MODELS))
def test_create_matches_reference_state(name):
    mine = PharmacophoreModel.create("", (0.0, 0.0, 0.0), synthetic.make_hotspot_infos(**MODELS[name]))
    with open(os.path.join(GOLDEN, f"model_{name}.pm"), "rb") as f:
        ref = pickle.load(f)
    st = mine.__getstate__()
    assert len(st["nodes"]) == len(ref["nodes"]) and len(st["edges"]) == len(ref["edges"])
    for a, b in zip(st["nodes"], ref["nodes"]):
        for k in ("index", "type", "interaction_type", "score", "center", "radius", "overlapped_nodes"):
            assert a[k] == b[k], (k, a[k], b[k])
        assert tuple(a["hotspot_position"]) == tuple(b["hotspot_position"])
        assert {int(k): v for k, v in a["neighbor_edge_dict"].items()} == {int(k): v for k, v in b["neighbor_edge_dict"].items()}
    for a, b in zip(st["edges"], ref["edges"]):
        assert a["index"] == b["index"] and tuple(a["node_indices"]) == tuple(b["node_indices"])
        assert tuple(a["edge_type"]) == tuple(b["edge_type"])
        assert a["distance_mean"] == b["distance_mean"] and a["distance_std"] == b["distance_std"]  # bit-exact fp64
    assert list(st["node_cluster_dict"]) == list(ref["node_cluster_dict"])
    for typ in ref["node_cluster_dict"]:
        assert len(st["node_cluster_dict"][typ]) == len(ref["node_cluster_dict"][typ]), typ
        for a, b in zip(st["node_cluster_dict"][typ], ref["node_cluster_dict"][typ]):
            assert set(a["node_indices"]) == set(b["node_indices"]) and set(a["node_types"]) == set(b["node_types"])
            assert tuple(a["center"]) == tuple(b["center"]) and a["size"] == b["size"]
    assert st["node_dict"] == ref["node_dict"]
    # and therefore the tables the kernel consumes are identical
    pr = PackedModel.from_model(PharmacophoreModel.load(os.path.join(GOLDEN, f"model_{name}.pm")))
    pm = PackedModel.from_model(mine)
    for k, v in pr.arrays().items():
        assert np.array_equal(v, pm.arrays()[k]), k


def test_small_components_are_dropped():
    m = np.zeros((64, 64, 64))
    m[10:12, 10:12, 10] = 0.9  # 4 voxels < 8
    m[30:33, 30:33, 30:33] = 0.8  # 27 voxels
    info = dict(nci_type="Hydrophobic", hotspot_position=np.zeros(3), hotspot_score=0.5, point_map=m)
    model = PharmacophoreModel.create("", (0.0, 0.0, 0.0), [info])
    assert len(model.nodes) == 1 and len(model.node_clusters) == 1
    assert abs(model.nodes[0].radius - (27 / (4 * np.pi / 3)) ** (1 / 3) * 0.5) < 1e-12


@pytest.mark.parametrize("name", ["syn0", "loose"])
def test_sparse_maps_give_the_same_model_state(name):
    """`point_map_sparse` (non-zero voxels in C order, what create_models copies back from the device) == dense maps."""
    infos = synthetic.make_hotspot_infos(**MODELS[name])
    dense = PharmacophoreModel.create("", (1.0, -2.0, 0.5), infos).__getstate__()
    sparse_infos = []
    for info in infos:
        m = info["point_map"]
        coords = np.argwhere(m > 0).astype(np.int32)
        d = {k: v for k, v in info.items() if k != "point_map"}
        d["point_map_sparse"] = (coords, m[coords[:, 0], coords[:, 1], coords[:, 2]])
        sparse_infos.append(d)
    sparse = PharmacophoreModel.create("", (1.0, -2.0, 0.5), sparse_infos).__getstate__()
    assert pickle.dumps(dense) == pickle.dumps(sparse)
