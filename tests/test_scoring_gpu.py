"""Parity of the CUDA scoring kernel (through the C-ABI) with the reference golden vectors and the CPU oracle."""

import numpy as np
import pytest
import torch
from golden_util import CASES, FALLBACK_CASES, load_case, load_fallback, rel_err, weights_dict

import oracle as orc
from pharmaconet_b200 import _abi, scoring, synthetic
from pharmaconet_b200.packing import LigandBatch

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # BASELINE.json north_star: scores within 1e-5 relative of the reference


def _run(model, batch, weights, **kw):
    dm = scoring.DeviceModel(model, "cuda:0")
    return scoring.score_library(dm, batch, weights_dict(weights) if weights is not None else None, with_stats=True, **kw)


@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_reference_golden(name):
    c = load_case(name)
    out = _run(c["model"], c["batch"], c["weights"])
    assert np.all(out["status"] <= _abi.LIG_EMPTY)
    assert rel_err(out["scores"], c["ref"]).max() <= REL_TOL
    # exact zeros where the reference returns 0 (no candidates)
    assert np.array_equal(out["scores"] == 0.0, c["ref"] == 0.0)


@pytest.mark.parametrize("name", FALLBACK_CASES)
def test_kernel_matches_reference_numpy_fallback(name):
    """second reference scorer (match_utils.py, fp32 throughout; graph_match.py:12-15) on the same ligands"""
    c = load_case(name)
    fb = load_fallback(name)
    out = _run(c["model"], c["batch"], c["weights"])
    assert rel_err(out["scores"], fb).max() <= REL_TOL
    assert np.array_equal(out["scores"] == 0.0, fb == 0.0)


@pytest.mark.parametrize("name", CASES)
def test_kernel_tree_shape_identical_to_oracle(name):
    # all discrete decisions are meant to be bit-identical: same number of tree nodes and leaves per ligand
    c = load_case(name)
    out = _run(c["model"], c["batch"], c["weights"])
    o = orc.score(c["model"], c["batch"], c["weights"])
    assert np.array_equal(out["status"], o["status"])
    assert np.array_equal(out["stats"][:, 0].astype(np.uint64), o["stats"][:, 0])
    assert np.array_equal(out["stats"][:, 1].astype(np.uint64), o["stats"][:, 1])
    assert np.array_equal(out["stats"][:, 3].astype(np.uint64), o["stats"][:, 3])


@pytest.mark.parametrize("nconf,n,seed", [(32, 4096, 101), (8, 2048, 102), (13, 1024, 103)])
def test_kernel_matches_oracle_fresh_inputs(nconf, n, seed):
    c = load_case("syn0_c32")
    batch = LigandBatch.from_typed(synthetic.make_ligands(n, nconf, seed=seed))
    out = _run(c["model"], batch, None)
    o = orc.score(c["model"], batch, None)
    assert np.array_equal(out["status"], o["status"])
    assert rel_err(out["scores"], o["scores"]).max() <= REL_TOL
    assert np.array_equal(out["stats"][:, 0].astype(np.uint64), o["stats"][:, 0])


def test_per_conformer_scores_match_oracle():
    c = load_case("syn0_c8")
    dm = scoring.DeviceModel(c["model"], "cuda:0")
    out = scoring.score_batch(dm, scoring.DeviceLigandBatch.from_host(c["batch"], "cuda:0"), with_conf=True)
    conf = out["conf"].cpu().numpy()[:, :8]
    o = orc.score(c["model"], c["batch"], c["weights"], with_conf=True)
    assert np.abs(conf - o["conf"][:, :8]).max() <= REL_TOL * max(1.0, np.abs(o["conf"]).max())


def test_ligand_order_invariance_bit_exact():
    c = load_case("syn0_c32")
    b = c["batch"]
    perm = np.random.default_rng(0).permutation(b.num_ligands)
    a = _run(c["model"], b, None)["scores"]
    p = _run(c["model"], b.select(perm), None)["scores"]
    assert np.array_equal(a[perm], p)


def test_identical_conformers_equal_single_conformer():
    c = load_case("syn0_c1")
    ligs = synthetic.make_ligands(**c["gen_kwargs"])
    for lig in ligs:
        lig.atom_positions = np.repeat(lig.atom_positions, 32, axis=1)
    rep = LigandBatch.from_typed(ligs)
    a = _run(c["model"], c["batch"], None)["scores"]
    r = _run(c["model"], rep, None)["scores"]
    assert rel_err(r, a).max() <= 1e-6


def test_overflow_is_reported_and_rerun():
    c = load_case("syn0_c5_big")
    dm = scoring.DeviceModel(c["model"], "cuda:0")
    db = scoring.DeviceLigandBatch.from_host(c["batch"], "cuda:0")
    tiny = scoring.ScoreConfig(warps_per_block=4, blocks=8, scratch_rows=64)
    st = scoring.score_batch(dm, db, config=tiny)["status"].cpu().numpy()
    assert (st == _abi.LIG_OVERFLOW).sum() > 0
    out = scoring.score_library(dm, c["batch"], config=tiny)
    assert np.all(out["status"] == _abi.LIG_OK)
    assert rel_err(out["scores"], c["ref"]).max() <= REL_TOL


def test_launch_geometry_does_not_change_results():
    c = load_case("syn0_c8")
    a = _run(c["model"], c["batch"], None)["scores"]
    for cfg in (scoring.ScoreConfig(1, 3, 4096), scoring.ScoreConfig(8, 296, 2048), scoring.ScoreConfig(4, 1, 8192)):
        b = _run(c["model"], c["batch"], None, config=cfg)["scores"]
        assert np.array_equal(a, b)


def test_empty_batch_and_bad_args():
    c = load_case("syn0_c8")
    dm = scoring.DeviceModel(c["model"], "cuda:0")
    out = scoring.score_library(dm, c["batch"].select([]))
    assert out["scores"].shape == (0,)
    with pytest.raises(RuntimeError):
        scoring.DeviceModel(c["model"], "cpu")


def test_topk_matches_sort():
    g = torch.Generator(device="cpu").manual_seed(0)
    s = torch.rand(100003, generator=g)
    s[::7] = s[3]  # ties
    sd = s.cuda()
    ks, ki = scoring.topk(sd, 1000, id_base=5000)
    order = np.lexsort((np.arange(s.numel()), -s.numpy()))[:1000]
    assert np.array_equal(ki.cpu().numpy(), order + 5000)
    assert np.array_equal(ks.cpu().numpy(), s.numpy()[order])
    ks2, ki2 = scoring.topk(sd[:10], 16)
    assert np.all(ki2.cpu().numpy()[10:] == -1) and np.all(np.isinf(ks2.cpu().numpy()[10:]))


def test_streamed_host_screening_equals_single_launch():
    from pharmaconet_b200 import screening

    c = load_case("syn0_c8")
    batch = LigandBatch.from_typed(synthetic.make_ligands(700, 8, seed=77) + synthetic.make_ligands(300, 5, seed=78))
    whole = _run(c["model"], batch, None)["scores"]
    scr = screening.Screener(c["model"], "cuda:0", k=50, block_ligands=128)
    for lib in (batch, screening.pin_library(batch)):
        res = scr.screen_host(lib)
        assert np.array_equal(res.scores, whole)  # block views with un-rebased offsets give bit-identical scores
        order = np.lexsort((np.arange(1000), -whole.astype(np.float64)))[:50]
        assert np.array_equal(res.topk_ids.cpu().numpy(), order)
        assert res.n_ligands == 1000 and res.n_conformers == 700 * 8 + 300 * 5
    # two-rank sharding by hand: union of both ranks' candidates equals the global head
    parts = [scr.screen_host(batch, rank=r, world=2, gather=False) for r in range(2)]
    ms, mi = screening.merge_topk(
        torch.cat([p.topk_scores for p in parts]), torch.cat([p.topk_ids for p in parts]), 50
    )
    assert np.array_equal(mi.cpu().numpy(), order)
    # device-resident path
    dres = scr.screen_device(scoring.DeviceLigandBatch.from_host(batch, "cuda:0"))
    assert np.array_equal(dres.topk_ids.cpu().numpy(), order)


@pytest.mark.parametrize("n_slots", [2, 4])
def test_streamed_screening_any_number_of_staging_slots(n_slots):
    """Blocks alternate between two compute streams whatever the number of staging slots; a slot is reused only after
    the kernels that read it have finished."""
    from pharmaconet_b200 import screening

    c = load_case("syn0_c8")
    batch = LigandBatch.from_typed(synthetic.make_ligands(900, 8, seed=79))
    whole = _run(c["model"], batch, None)["scores"]
    scr = screening.Screener(c["model"], "cuda:0", k=30, block_ligands=64, n_slots=n_slots)
    for _ in range(2):
        res = scr.screen_host(screening.pin_library(batch))
        assert np.array_equal(res.scores, whole)


def test_streamed_screening_reruns_overflow():
    from pharmaconet_b200 import screening

    c = load_case("syn0_c5_big")
    scr = screening.Screener(c["model"], "cuda:0", k=8, block_ligands=7, config=scoring.ScoreConfig(4, 8, 64))
    res = scr.screen_host(c["batch"])
    assert res.n_overflow > 0
    assert rel_err(res.scores, c["ref"]).max() <= REL_TOL
    order = np.lexsort((np.arange(len(c["ref"])), -res.scores.astype(np.float64)))[:8]
    assert np.array_equal(res.topk_ids.cpu().numpy(), order)


def test_screening_cli_on_packed_library(tmp_path):
    import os
    import subprocess
    import sys

    from golden_util import GOLDEN

    from pharmaconet_b200.packing import save_library

    c = load_case("syn0_c8")
    names = [f"lig_{i:04d}.sdf" for i in range(c["batch"].num_ligands)]
    save_library(tmp_path / "lib.npz", c["batch"], names)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "out.csv"
    subprocess.run(
        [sys.executable, os.path.join(root, "screening.py"), "-p", os.path.join(GOLDEN, "model_syn0.pm"),
         "-d", str(tmp_path / "lib.npz"), "-o", str(out)],
        check=True, cwd=root, timeout=300,
    )  # fmt: skip
    lines = out.read_text().splitlines()
    assert lines[0] == "path,score" and len(lines) == 1 + len(names)
    got = {ln.split(",")[0]: float(ln.split(",")[1]) for ln in lines[1:]}
    vals = [float(ln.split(",")[1]) for ln in lines[1:]]
    assert vals == sorted(vals, reverse=True)  # screening.py:70 sorts by score, descending
    ref = dict(zip(names, c["ref"]))
    assert max(abs(got[n] - ref[n]) / max(abs(ref[n]), 1e-12) for n in names) <= REL_TOL


def test_screening_over_a_memory_mapped_library_directory(tmp_path):
    """A library saved as a directory of .npy files is memory-mapped by load_library: Screener.screen_host streams its
    blocks straight from the mapped arrays (pageable copies), and the CLI accepts the directory - same scores."""
    import os
    import subprocess
    import sys

    from golden_util import GOLDEN

    from pharmaconet_b200 import screening
    from pharmaconet_b200.packing import load_library, save_library

    c = load_case("syn0_c8")
    names = [f"lig_{i:04d}.sdf" for i in range(c["batch"].num_ligands)]
    save_library(tmp_path / "lib_dir", c["batch"], names)
    lib, got_names = load_library(tmp_path / "lib_dir")
    assert got_names == names
    whole = _run(c["model"], c["batch"], None)["scores"]
    res = screening.Screener(c["model"], "cuda:0", k=16, block_ligands=64).screen_host(lib)
    assert np.array_equal(res.scores, whole)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "out.csv"
    subprocess.run(
        [sys.executable, os.path.join(root, "screening.py"), "-p", os.path.join(GOLDEN, "model_syn0.pm"),
         "-d", str(tmp_path / "lib_dir"), "-o", str(out)],
        check=True, cwd=root, timeout=300,
    )  # fmt: skip
    lines = out.read_text().splitlines()
    got = {ln.split(",")[0]: float(ln.split(",")[1]) for ln in lines[1:]}
    ref = dict(zip(names, c["ref"]))
    assert max(abs(got[n] - ref[n]) / max(abs(ref[n]), 1e-12) for n in names) <= REL_TOL


def test_more_than_32_conformers_mixed_batch():
    # 2 and 4 conformers per lane, mixed with short ligands in the same launch; per-conformer maxima too
    c = load_case("syn0_c100")
    ligs = synthetic.make_ligands(40, 70, seed=31) + synthetic.make_ligands(40, 7, seed=32) + synthetic.make_ligands(8, 128, seed=33)
    batch = LigandBatch.from_typed(ligs)
    out = _run(c["model"], batch, None)
    o = orc.score(c["model"], batch, None, with_conf=True)
    assert np.array_equal(out["status"], o["status"]) and np.all(out["status"] == _abi.LIG_OK)
    assert rel_err(out["scores"], o["scores"]).max() <= REL_TOL
    assert np.array_equal(out["stats"][:, 0].astype(np.uint64), o["stats"][:, 0])
    dm = scoring.DeviceModel(c["model"], "cuda:0")
    conf = scoring.score_batch(dm, scoring.DeviceLigandBatch.from_host(batch, "cuda:0"), with_conf=True)["conf"]
    assert conf.shape == (88, 128)
    assert np.abs(conf.cpu().numpy() - o["conf"]).max() <= REL_TOL * np.abs(o["conf"]).max()
    too_many = LigandBatch.from_typed(synthetic.make_ligands(2, 130, seed=34))
    with pytest.raises(ValueError):
        scoring.score_batch(dm, scoring.DeviceLigandBatch.from_host(too_many, "cuda:0"))


def test_ligand_without_pharmacophores_scores_zero():
    from pharmaconet_b200.ligand import TypedLigand

    c = load_case("syn0_c8")
    ligs = synthetic.make_ligands(3, 8, seed=40)
    empty = TypedLigand([6, 8], [[1], [0]], [], np.zeros((2, 8, 3), dtype=np.float32))
    batch = LigandBatch.from_typed([ligs[0], empty, ligs[1], empty, ligs[2]])
    out = _run(c["model"], batch, None)
    o = orc.score(c["model"], batch, None)
    assert out["scores"][1] == 0.0 and out["scores"][3] == 0.0  # graph_match.py:95-96
    assert np.array_equal(out["status"], o["status"]) and out["status"][1] == _abi.LIG_EMPTY
    assert rel_err(out["scores"][[0, 2, 4]], o["scores"][[0, 2, 4]]).max() <= REL_TOL


def test_screen_models_equals_one_screener_per_model():
    """Many models against one resident library (BASELINE configs[4]) == one Screener per model."""
    from pharmaconet_b200 import screening

    lib = LigandBatch.from_typed(synthetic.make_ligands(3000, 16, seed=77))
    db = scoring.DeviceLigandBatch.from_host(lib, "cuda:0")
    models = [load_case(n)["model"] for n in ("syn0_c8", "sparse_c16", "xbond_c4")]
    res = screening.screen_models(models, db, host_lib=lib, k=50, keep_scores=True)
    assert len(res) == 3
    for m, r in zip(models, res):
        one = screening.Screener(m, "cuda:0", k=50).screen_device(db)
        assert torch.equal(r.topk_ids, one.topk_ids)
        assert torch.equal(r.topk_scores, one.topk_scores)
        assert torch.equal(r.scores, one.scores)


def test_streamed_screening_ramp_up_spans():
    """First block cut into growing spans (1/8, 1/8, 1/4, 1/2): same scores, ids and top-k as one launch."""
    from pharmaconet_b200 import screening

    c = load_case("syn0_c8")
    batch = LigandBatch.from_typed(synthetic.make_ligands(9000, 4, seed=91))
    whole = _run(c["model"], batch, None)["scores"]
    scr = screening.Screener(c["model"], "cuda:0", k=64, block_ligands=8192)
    res = scr.screen_host(screening.pin_library(batch))
    assert np.array_equal(res.ids, np.arange(9000))
    assert np.array_equal(res.scores, whole)
    order = np.lexsort((np.arange(9000), -whole.astype(np.float64)))[:64]
    assert np.array_equal(res.topk_ids.cpu().numpy(), order)
    # four ramp spans + the second block, each: cost kernel, the scoring call's kernels (specialised, three task
    # launches - the first one also serves the deferred ligands -, heavy-ligand finish), id fill, top-k
    assert res.launches == 5 * (1 + scoring.launches_per_call() + 2) == 5 * 8


def test_cost_order_is_a_stable_permutation_and_does_not_change_scores():
    c = load_case("syn0_c8")
    dm = scoring.DeviceModel(c["model"], "cuda:0")
    lib = LigandBatch.from_typed(synthetic.make_ligands(5000, 8, seed=5))
    db = scoring.DeviceLigandBatch.from_host(lib, "cuda:0")
    base = scoring.score_batch(dm, db, with_stats=True)
    order = scoring.cost_order(dm, db)
    o = order.cpu().numpy()
    assert np.array_equal(np.sort(o), np.arange(5000))
    # key = number of (level, model cluster) entries = what the oracle reports as `entries`; descending, stable
    ent = orc.score(c["model"], lib)["stats"][:, 2].astype(np.int64)
    assert np.array_equal(o, np.lexsort((np.arange(5000), -ent)))
    db.set_order(order)
    ordered = scoring.score_batch(dm, db, with_stats=True)
    for k in ("scores", "status", "stats"):
        assert torch.equal(base[k], ordered[k])
    # a chunk view with un-rebased offsets gets the same relative order
    from pharmaconet_b200 import screening

    scr = screening.Screener(c["model"], "cuda:0", k=32, block_ligands=1024)
    assert np.array_equal(scr.screen_host(lib).scores, base["scores"].cpu().numpy())


def test_screening_cli_on_sdf_directory(tmp_path):
    """BASELINE configs[0] plumbing: a directory of multi-conformer .sdf files through screening.py (built-in typing,
    OpenBabel is absent) equals typing + scoring the same files through the API and the CPU oracle."""
    import os
    import subprocess
    import sys

    from golden_util import GOLDEN
    from sdf_util import molblock, ring

    from pharmaconet_b200.ligand_typing import typed_ligand_from_file

    rng = np.random.default_rng(3)
    lib = tmp_path / "lib"
    lib.mkdir()

    def write(name, atoms, bonds, nconf=3):
        text = ""
        for _ in range(nconf):
            jit = rng.normal(0, 0.15, (len(atoms), 3))
            text += molblock([(s, x + j[0], y + j[1], z + j[2]) for (s, x, y, z), j in zip(atoms, jit)], bonds)
        (lib / name).write_text(text)

    # phenol-like ring with a hydroxyl, a pyridine with an amine tail, an acid chain
    a, b = ring(6, "CCCCCC", [2, 1, 2, 1, 2, 1], [("O", 2.8, 0.0, 0.0), ("H", 3.4, 0.7, 0.0)], [(1, 7, 1), (7, 8, 1)])
    write("phenol.sdf", a, b)
    a, b = ring(6, "NCCCCC", [2, 1, 2, 1, 2, 1],
                [("C", -2.8, 0.0, 0.3), ("N", -3.9, 0.9, 0.5), ("C", -5.2, 0.3, 0.2), ("C", -3.9, 2.3, 0.9), ("C", -4.0, 0.0, 2.0)],
                [(4, 7, 1), (7, 8, 1), (8, 9, 1), (8, 10, 1), (8, 11, 1)])  # fmt: skip
    write("pyridine_amine.sdf", a, b, nconf=5)
    write("acid.sdf", [("C", 0, 0, 0), ("C", 1.5, 0, 0), ("C", 2.3, 1.2, 0), ("O", 3.5, 1.2, 0.4), ("O", 1.7, 2.3, -0.3),
                       ("H", 2.3, 3.0, -0.2), ("Cl", -1.7, 0.3, 0.2)],
          [(1, 2, 1), (2, 3, 1), (3, 4, 2), (3, 5, 1), (5, 6, 1), (1, 7, 1)], nconf=4)  # fmt: skip
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "out.csv"
    subprocess.run(
        [sys.executable, os.path.join(root, "screening.py"), "-p", os.path.join(GOLDEN, "model_syn0.pm"),
         "-d", str(lib), "-o", str(out)],
        check=True, cwd=root, timeout=300,
    )  # fmt: skip
    lines = out.read_text().splitlines()
    assert lines[0] == "path,score" and len(lines) == 4
    got = {os.path.basename(ln.split(",")[0]): float(ln.split(",")[1]) for ln in lines[1:]}
    files = sorted(os.listdir(lib))
    ligs = [typed_ligand_from_file(str(lib / f), perception="builtin") for f in files]
    batch = LigandBatch.from_typed(ligs)
    assert list(batch.n_conf) == [4, 3, 5]
    ref = orc.score(load_case("syn0_c8")["model"], batch)["scores"]
    assert any(r > 0 for r in ref)
    for f, r in zip(files, ref):
        assert abs(got[f] - r) <= REL_TOL * max(abs(r), 1e-12), (f, got[f], r)


def test_large_model_tables_in_global_memory():
    """A model whose edge / cluster tables (182 KB) do not fit in shared memory next to the DFS stacks: the kernel
    reads them from global memory instead (103 nodes, 45 overlapping clusters, ~90 k tree nodes per ligand)."""
    from pharmaconet_b200.packing import PackedModel
    from pharmaconet_b200.pharmacophore_model import PharmacophoreModel

    m = PharmacophoreModel.create("", (0.0, 0.0, 0.0), synthetic.make_hotspot_infos(seed=21, n_hotspots=130))
    pm = PackedModel.from_model(m)
    assert pm.num_nodes**2 * 16 + pm.num_clusters**2 * 8 > 100 * 1024
    batch = LigandBatch.from_typed(synthetic.make_ligands(64, 8, seed=3))
    out = _run(pm, batch, None)
    ref = orc.score(pm, batch)
    assert np.array_equal(out["status"], ref["status"])
    assert rel_err(out["scores"], ref["scores"]).max() <= REL_TOL
    assert np.array_equal(out["stats"][:, 0], ref["stats"][:, 0].astype(np.uint32))  # tree nodes
    assert np.array_equal(out["stats"][:, 1], ref["stats"][:, 1].astype(np.uint32))  # leaves
    # same answer through the screening driver (per-model order, two streams)
    from pharmaconet_b200 import screening

    db = scoring.DeviceLigandBatch.from_host(batch, "cuda:0")
    res = screening.screen_models([pm, load_case("syn0_c8")["model"]], db, host_lib=batch, k=16, keep_scores=True)
    assert rel_err(res[0].scores.cpu().numpy(), ref["scores"]).max() <= REL_TOL


def test_specialised_and_general_kernel_agree_bit_for_bit():
    """Default configuration = specialised kernel (+ general kernel for what it defers); an explicit launch shape =
    general kernel alone. Same fp32 operations in the same order: identical scores, statuses and tree shapes."""
    c = load_case("syn0_c32")
    batch = LigandBatch.from_typed(synthetic.make_ligands(3000, 32, seed=211) + synthetic.make_ligands(200, 7, seed=212, frag_range=(9, 15)))
    a = _run(c["model"], batch, None)
    b = _run(c["model"], batch, None, config=scoring.ScoreConfig(32, 148, 8192))
    assert np.array_equal(a["status"], b["status"]) and np.all(a["status"] <= _abi.LIG_EMPTY)
    assert np.array_equal(a["scores"], b["scores"])
    assert np.array_equal(a["stats"][:, [0, 1, 3]], b["stats"][:, [0, 1, 3]])
    o = orc.score(c["model"], batch, None)
    assert rel_err(a["scores"], o["scores"]).max() <= REL_TOL
    assert np.array_equal(a["stats"][:, 0].astype(np.uint64), o["stats"][:, 0])


def test_no_ligand_is_left_deferred():
    """Ligands beyond the specialised kernel's caps (many clusters, wide model) are finished by the general kernel in the
    same call: PMNET_LIG_DEFERRED never reaches the caller."""
    for name in ("syn0_c4_deep", "syn0_c5_big", "loose_c8"):
        c = load_case(name)
        dm = scoring.DeviceModel(c["model"], "cuda:0")
        out = scoring.score_batch(dm, scoring.DeviceLigandBatch.from_host(c["batch"], "cuda:0"))
        st = out["status"].cpu().numpy()
        assert not np.any(st == _abi.LIG_DEFERRED)


def test_screen_device_reruns_overflow_in_place():
    from pharmaconet_b200 import screening

    c = load_case("syn0_c5_big")
    scr = screening.Screener(c["model"], "cuda:0", k=8, config=scoring.ScoreConfig(4, 8, 64))
    res = scr.screen_device(scoring.DeviceLigandBatch.from_host(c["batch"], "cuda:0"))
    assert res.n_overflow > 0
    assert rel_err(res.scores.cpu().numpy(), c["ref"]).max() <= REL_TOL
    order = np.lexsort((np.arange(len(c["ref"])), -res.scores.cpu().numpy().astype(np.float64)))[:8]
    assert np.array_equal(res.topk_ids.cpu().numpy(), order)


def test_rescore_status_only_touches_matching_ligands():
    c = load_case("syn0_c8")
    dm = scoring.DeviceModel(c["model"], "cuda:0")
    db = scoring.DeviceLigandBatch.from_host(c["batch"], "cuda:0")
    out = scoring.score_batch(dm, db)
    ref = out["scores"].clone()
    # pretend three ligands overflowed and wipe their scores: only they are recomputed
    idx = torch.tensor([3, 17, 90], device="cuda:0")
    out["status"][idx] = _abi.LIG_OVERFLOW
    out["scores"][:] = -1.0
    n = scoring.rescore_overflowed_device(dm, db, out)
    assert n == 3
    s = out["scores"].cpu().numpy()
    assert np.array_equal(s[idx.cpu().numpy()], ref[idx].cpu().numpy())
    keep = np.ones(len(s), bool)
    keep[idx.cpu().numpy()] = False
    assert np.all(s[keep] == -1.0)


def test_streamed_screening_mixed_conformer_counts():
    """A library whose blocks differ in their largest conformer count (8 vs 40): every block uses the library-wide
    maximum, so one kernel instantiation and one workspace layout serve the whole screen."""
    from pharmaconet_b200 import screening

    c = load_case("syn0_c32")
    ligs = synthetic.make_ligands(300, 8, seed=301) + synthetic.make_ligands(60, 40, seed=302) + synthetic.make_ligands(200, 8, seed=303)
    batch = LigandBatch.from_typed(ligs)
    whole = _run(c["model"], batch, None)["scores"]
    scr = screening.Screener(c["model"], "cuda:0", k=20, block_ligands=128, ramp=False)
    res = scr.screen_host(batch)
    assert np.array_equal(res.scores, whole)
    o = orc.score(c["model"], batch, None)
    assert rel_err(res.scores, o["scores"]).max() <= REL_TOL


def _score_with_budget(model, batch, budget, config=None, with_conf=True):
    """One pmnet_score_batch call with an explicit workspace so that the number of ligands handed to the task-parallel
    walk (workspace header word 1) can be read back."""
    from dataclasses import replace

    dm = scoring.DeviceModel(model, "cuda:0")
    db = scoring.DeviceLigandBatch.from_host(batch, "cuda:0")
    cfg = replace(config or scoring.ScoreConfig(), heavy_budget=budget)
    ws = torch.zeros(scoring.workspace_bytes(dm, cfg, db.max_conformers), dtype=torch.uint8, device="cuda:0")
    out = scoring.score_batch(dm, db, None, cfg, with_stats=True, with_conf=with_conf, workspace=ws)
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    hdr = ws[:256].view(torch.int32).cpu().numpy()
    # header words: [1] ligands over the budget, [16] tasks donated, [24] tasks taken, [9] replay failures (never)
    res["n_heavy"], res["n_tasks"], res["n_bad"] = int(hdr[1]), int(hdr[16]), int(hdr[9])
    assert int(hdr[24]) == min(int(hdr[16]), 1 << 18) and int(hdr[10]) == 0  # every task taken, no walker left
    return res


@pytest.mark.parametrize("name", ["syn0_c32", "syn0_c8", "syn0_c4_deep", "syn0_c5_big", "loose_c8"])
@pytest.mark.parametrize("general_only", [False, True])
def test_task_parallel_walk_is_identical_to_single_warp_walk(name, general_only):
    """PmScoreConfig.heavy_budget: a ligand whose tree outgrows the budget is abandoned and walked by the task kernel,
    whose walkers give unvisited candidates away as tasks every `budget` nodes. A tiny budget sends a large share of
    the golden ligands down that path and splits them many times: scores, per-conformer scores, statuses and tree
    shapes must equal the un-split walk bit for bit (and the oracle's tree shapes)."""
    c = load_case(name)
    cfg = scoring.ScoreConfig(16, 148, 8192) if general_only else None
    base = _score_with_budget(c["model"], c["batch"], -1, cfg)
    assert base["n_heavy"] == 0
    split = _score_with_budget(c["model"], c["batch"], 40, cfg)
    assert split["n_heavy"] > 0 and split["n_tasks"] > 0 and split["n_bad"] == 0
    assert np.array_equal(split["status"], base["status"]) and not np.any(split["status"] == _abi.LIG_HEAVY)
    assert np.array_equal(split["scores"], base["scores"])
    assert np.array_equal(split["conf"], base["conf"])
    assert np.array_equal(split["stats"], base["stats"])
    if not general_only:
        o = orc.score(c["model"], c["batch"], None)
        ok = split["status"] == _abi.LIG_OK  # (an overflowed ligand has no tree here: score_batch alone does not re-run)
        assert ok.sum() > 0
        assert np.array_equal(split["stats"][ok, 0].astype(np.uint32), o["stats"][ok, 0].astype(np.uint32))
        assert np.array_equal(split["stats"][ok, 1].astype(np.uint32), o["stats"][ok, 1].astype(np.uint32))


@pytest.mark.parametrize("nconf", [40, 100])
def test_task_parallel_walk_with_more_than_32_conformers(nconf):
    """2 or 4 conformers per lane: the accumulator of a heavy ligand holds up to 128 per-conformer maxima."""
    c = load_case("syn0_c32")
    batch = LigandBatch.from_typed(synthetic.make_ligands(600, nconf, seed=431) + synthetic.make_ligands(200, 7, seed=432))
    base = _score_with_budget(c["model"], batch, -1)
    split = _score_with_budget(c["model"], batch, 40)
    assert split["n_heavy"] > 100 and split["n_tasks"] > 0 and split["n_bad"] == 0
    for k in ("status", "scores", "conf", "stats"):
        assert np.array_equal(split[k], base[k]), k
    o = orc.score(c["model"], batch, None)
    assert rel_err(split["scores"], o["scores"]).max() <= REL_TOL
    assert np.array_equal(split["stats"][:, 0].astype(np.uint32), o["stats"][:, 0].astype(np.uint32))


def test_task_parallel_walk_full_queues():
    """A budget of 8 nodes on 6500 ligands: nearly every ligand goes to the task kernel and is cut into many tasks
    (paths through None children included), taken by idle warps of the same launch as they appear."""
    c = load_case("syn0_c32")
    batch = LigandBatch.from_typed(synthetic.make_ligands(6000, 32, seed=411) + synthetic.make_ligands(500, 5, seed=412, frag_range=(2, 6)))
    base = _score_with_budget(c["model"], batch, -1)
    split = _score_with_budget(c["model"], batch, 8)
    assert split["n_heavy"] > 4096 and split["n_tasks"] > 50000 and split["n_bad"] == 0
    for k in ("status", "scores", "conf", "stats"):
        assert np.array_equal(split[k], base[k]), k


def test_task_parallel_walk_full_heavy_list():
    """More ligands over the budget than the list of heavy ligands holds (65 536): the rest are walked to the end by
    the warp that has them."""
    c = load_case("syn0_c8")
    base_lib = LigandBatch.from_typed(synthetic.make_ligands(3000, 8, seed=421))
    batch = base_lib.select(np.tile(np.arange(3000), 30))
    base = _score_with_budget(c["model"], batch, -1, with_conf=False)
    split = _score_with_budget(c["model"], batch, 8, with_conf=False)
    assert split["n_heavy"] > 65536 and split["n_bad"] == 0
    for k in ("status", "scores", "stats"):
        assert np.array_equal(split[k], base[k]), k


def test_default_budget_leaves_ordinary_libraries_alone():
    c = load_case("syn0_c32")
    out = _score_with_budget(c["model"], c["batch"], 0)
    assert out["n_heavy"] == 0
    assert rel_err(out["scores"], c["ref"]).max() <= REL_TOL
