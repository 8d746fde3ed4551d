"""PharmacophoreModel load/save/pickle round trips on .pm files written by the reference's own .save."""

import json
import os
import pickle

import numpy as np
import pytest
from golden_util import GOLDEN, load_case

from pharmaconet_b200.packing import PackedModel
from pharmaconet_b200.pharmacophore_model import PharmacophoreModel


@pytest.mark.parametrize("name,case", [("syn0", "syn0_c8"), ("loose", "loose_c8"), ("sparse", "sparse_c16"), ("xbond", "xbond_c4")])
def test_load_reference_pm_and_pack(name, case):
    m = PharmacophoreModel.load(os.path.join(GOLDEN, f"model_{name}.pm"))
    packed = PackedModel.from_model(m)
    gold = load_case(case)["model"]  # packed from the reference's own in-memory object
    for k, v in gold.arrays().items():
        assert np.array_equal(v, packed.arrays()[k]), k


def test_attributes_match_reference_surface():
    m = PharmacophoreModel.load(os.path.join(GOLDEN, "model_syn0.pm"))
    assert len(m.nodes) == 35 and len(m.node_clusters) == 26
    assert len(m.edges) == 35 * 36 // 2  # complete, self loops included
    assert list(m.node_cluster_dict) == ["Cation", "Anion", "HBond", "Aromatic", "Hydrophobic", "Halogen"]
    n = m.nodes[3]
    assert n.neighbor_edge_dict[n].distance_mean == 0.0
    assert sum(len(v) for v in m.node_dict.values()) == 35


def test_save_load_roundtrip_pm_json_pickle(tmp_path):
    m = PharmacophoreModel.load(os.path.join(GOLDEN, "model_syn0.pm"))
    ref = PackedModel.from_model(m)
    m.save(tmp_path / "a.pm")
    m.save(tmp_path / "a.json")
    clones = [
        PharmacophoreModel.load(tmp_path / "a.pm"),
        PharmacophoreModel.load(tmp_path / "a.json"),
        pickle.loads(pickle.dumps(m)),
    ]
    for c in clones:
        p = PackedModel.from_model(c)
        for k, v in ref.arrays().items():
            assert np.array_equal(v, p.arrays()[k]), k
    # the state written is the reference's layout (appendix F): same top-level keys
    state = json.load(open(tmp_path / "a.json"))
    assert set(state) == {"pdbblock", "nodes", "edges", "node_cluster_dict", "node_dict"}
    with open(os.path.join(GOLDEN, "model_syn0.pm"), "rb") as f:
        orig = pickle.load(f)
    assert pickle.load(open(tmp_path / "a.pm", "rb")).keys() == orig.keys()
    with pytest.raises(NotImplementedError):
        m.save(tmp_path / "a.txt")
