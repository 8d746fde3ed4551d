"""Host-side screening logic without a GPU: block sharding, top-k merge, and the world_size-2 gather over gloo."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pharmaconet_b200 import screening


def test_shard_blocks_partition():
    for n, world, blk in [(10, 1, 3), (1000, 4, 64), (65536 * 3 + 5, 8, 65536), (5, 8, 2), (0, 2, 4)]:
        seen = []
        for r in range(world):
            for a, b in screening.shard_blocks(n, r, world, blk):
                assert 0 <= a < b <= n and a % blk == 0 and (a // blk) % world == r
                seen.extend(range(a, b))
        assert sorted(seen) == list(range(n))


def _naive_topk(scores, ids, k):
    order = sorted(range(len(scores)), key=lambda i: (-scores[i], ids[i]))[:k]
    return [scores[i] for i in order], [ids[i] for i in order]


def test_merge_topk_ties_and_padding():
    g = torch.Generator().manual_seed(0)
    s = torch.randint(0, 20, (500,), generator=g).float()
    ids = torch.randperm(500, generator=g)
    ks, ki = screening.merge_topk(s, ids, 64)
    es, ei = _naive_topk(s.tolist(), ids.tolist(), 64)
    assert ks.tolist() == es and ki.tolist() == ei
    # padding entries (-1 ids) never win, short inputs are padded
    s2 = torch.tensor([5.0, 9.0, 7.0])
    i2 = torch.tensor([3, -1, 1])
    ks, ki = screening.merge_topk(s2, i2, 4)
    assert ki.tolist() == [1, 3, -1, -1] and ks[:2].tolist() == [7.0, 5.0] and torch.isinf(ks[2:]).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, k, blk, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scores = torch.from_numpy(np.random.default_rng(7).integers(0, 50, n).astype(np.float32))  # same on all ranks
    mine_s, mine_i = [], []
    for a, b in screening.shard_blocks(n, rank, world, blk):
        mine_s.append(scores[a:b])
        mine_i.append(torch.arange(a, b))
    ls, li = screening.merge_topk(torch.cat(mine_s), torch.cat(mine_i), k)
    gs, gi = screening.gather_topk(ls, li, k)
    torch.save((gs, gi), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_world2_gather_equals_global_topk(tmp_path):
    n, k, blk, world = 1000, 37, 64, 2
    mp.spawn(_worker, args=(world, _free_port(), n, k, blk, str(tmp_path)), nprocs=world, join=True)
    scores = np.random.default_rng(7).integers(0, 50, n).astype(np.float32)
    es, ei = _naive_topk(scores.tolist(), list(range(n)), k)
    for r in range(world):
        gs, gi = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert gs.tolist() == es and gi.tolist() == ei


def test_write_csv_matches_reference_format(tmp_path):
    p = tmp_path / "out.csv"
    screening.write_csv(str(p), ["a.sdf", "b.sdf", "c.sdf"], [1.5, 3.0, 1.5])
    assert p.read_text().splitlines() == ["path,score", "b.sdf,3.0", "a.sdf,1.5", "c.sdf,1.5"]


def test_numa_binding_is_a_noop_without_nvml():
    """affinity.bind_to_gpu never raises and leaves the affinity alone when NVML / the GPU is not there (this container)."""
    import os

    from pharmaconet_b200.affinity import bind_to_gpu

    before = os.sched_getaffinity(0)
    cpus = bind_to_gpu(0)
    if cpus is None:
        assert os.sched_getaffinity(0) == before
    else:  # a GPU box: bound to a non-empty subset of what was allowed
        assert set(cpus) <= before and len(cpus) > 0
        os.sched_setaffinity(0, before)


def _pipeline_worker(rank, world, port, out_dir):
    from pharmaconet_b200 import pipeline
    from pharmaconet_b200.packing import PackedModel
    from pharmaconet_b200.pharmacophore_model import PharmacophoreModel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    names = ["syn0", "loose", "sparse", "xbond", "syn0"]  # 5 "pockets": rank r builds pockets r, r + 2, ...
    local = {
        p: PackedModel.from_model(PharmacophoreModel.load(os.path.join(golden, f"model_{names[p]}.pm")))
        for p in pipeline.pockets_of_rank(len(names), rank, world)
    }
    models = pipeline.exchange_models(local, len(names))
    torch.save([m.arrays() for m in models], os.path.join(out_dir, f"m{rank}.pt"))
    dist.destroy_process_group()


def test_pipeline_model_exchange_world2(tmp_path):
    """configs[4]: the packed pharmacophore models built on different ranks reach every rank, in pocket order."""
    from pharmaconet_b200 import pipeline
    from pharmaconet_b200.packing import PackedModel
    from pharmaconet_b200.pharmacophore_model import PharmacophoreModel

    assert pipeline.pockets_of_rank(5, 0, 2) == [0, 2, 4] and pipeline.pockets_of_rank(5, 1, 2) == [1, 3]
    port = _free_port()
    mp.spawn(_pipeline_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(tmp_path, "m0.pt"), weights_only=False)
    b = torch.load(os.path.join(tmp_path, "m1.pt"), weights_only=False)
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    names = ["syn0", "loose", "sparse", "xbond", "syn0"]
    for p, name in enumerate(names):
        ref = PackedModel.from_model(PharmacophoreModel.load(os.path.join(golden, f"model_{name}.pm"))).arrays()
        for k, v in ref.items():
            assert np.array_equal(a[p][k], v) and np.array_equal(b[p][k], v), (p, k)
    # single process: no collective, same result
    local = {p: PackedModel.from_arrays(a[p]) for p in range(5)}
    solo = pipeline.exchange_models(local, 5)
    assert all(np.array_equal(solo[p].arrays()["edge_mu"], a[p]["edge_mu"]) for p in range(5))
