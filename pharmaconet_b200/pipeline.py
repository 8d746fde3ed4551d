"""Many pockets x one ligand library on several GPUs (BASELINE configs[4]): the reference runs `modeling.py` once per
pocket (modeling.py:60-108) and `screening.py` once per model (screening.py:46-75); here

  1. pocket p belongs to rank p mod world: its CNN forward (batched in chunks), mask head, density maps and graph
     construction run there (`PharmacoNet.create_models`);
  2. ONE all-gather of the packed pharmacophore models (a few KB each) gives every rank every model, in pocket order;
  3. every rank screens ITS ligand shard against ALL models (`screening.screen_models`: the shard stays resident in
     HBM, one scoring launch + top-k per model, two streams);
  4. ONE all-gather of the per-model top-k (k x 12 B per model and rank) and the same merge on every rank.

Nothing else crosses the fabric: no data-path collective. torch.distributed is the plumbing.
"""

from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np
import torch

from .packing import PackedModel


def pockets_of_rank(n_pockets: int, rank: int, world: int) -> list[int]:
    return list(range(rank, n_pockets, world))


def exchange_models(local: dict[int, PackedModel], n_pockets: int, device=None) -> list[PackedModel | None]:
    """All-gather the packed models: `local` = {pocket index: PackedModel} built on this rank. Returns the models of ALL
    pockets in pocket order (None for a pocket whose model is empty) - identical on every rank. One collective: the
    arrays of every local model are serialised into one uint8 tensor, sizes first."""
    import io

    import torch.distributed as dist

    def pack(models: dict[int, PackedModel]) -> bytes:
        buf = io.BytesIO()
        np.savez(buf, idx=np.asarray(sorted(models), dtype=np.int64),
                 **{f"{p}_{k}": v for p, m in models.items() for k, v in m.arrays().items()})  # fmt: skip
        return buf.getvalue()

    def unpack(blob: bytes) -> dict[int, PackedModel]:
        z = np.load(io.BytesIO(blob), allow_pickle=False)
        return {
            int(p): PackedModel.from_arrays({k: z[f"{int(p)}_{k}"] for k in PackedModel.__dataclass_fields__})
            for p in z["idx"]
        }

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        merged = dict(local)
    else:
        world = dist.get_world_size()
        dev = torch.device(device) if device is not None else torch.device("cpu")
        blob = np.frombuffer(pack(local), dtype=np.uint8)
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        mine = torch.tensor([blob.size], dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, mine)
        cap = int(sizes.max().item())
        send = torch.zeros(cap, dtype=torch.uint8, device=dev)
        send[: blob.size] = torch.from_numpy(blob.copy()).to(dev)
        recv = torch.empty(world * cap, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(recv, send)
        recv = recv.cpu().numpy()
        merged = {}
        for r in range(world):
            merged.update(unpack(recv[r * cap : r * cap + int(sizes[r].item())].tobytes()))
    return [merged.get(p) if (p in merged and merged[p].num_nodes > 0) else None for p in range(n_pockets)]


@dataclass
class PipelineResult:
    models: list  # PackedModel or None per pocket (every rank holds all of them)
    topk_scores: list  # per pocket: tensor [k] (None for an empty model)
    topk_ids: list  # per pocket: tensor [k] int64 global ligand ids
    n_overflow: int
    seconds: dict = field(default_factory=dict)  # stage -> wall seconds on this rank (device-synchronised)


def model_and_screen(net, protein_data_list, shard, id_base: int = 0, k: int = 1000, rank: int = 0, world: int = 1,
                     centers=None, chunk: int = 8, weights=None) -> PipelineResult:
    """protein_data_list: ALL pockets (every rank passes the same list; a rank only touches its own);
    shard: this rank's device-resident ligand library (scoring.DeviceLigandBatch); id_base: global id of its first
    ligand. Returns per-pocket top-k merged over all ranks."""
    from . import screening

    dev = shard.device
    n = len(protein_data_list)
    mine = pockets_of_rank(n, rank, world)
    t0 = time.perf_counter()
    built = net.create_models(
        [protein_data_list[p] for p in mine], centers=[centers[p] for p in mine] if centers else None, chunk=chunk
    )
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    models = exchange_models({p: m.packed for p, m in zip(mine, built)}, n, device=dev)
    t2 = time.perf_counter()
    live = [p for p in range(n) if models[p] is not None]
    res = screening.screen_models([models[p] for p in live], shard, weights=weights, k=k, id_base=id_base, gather=world > 1)
    torch.cuda.synchronize(dev)
    t3 = time.perf_counter()
    ks, ki = [None] * n, [None] * n
    for p, r in zip(live, res):
        ks[p], ki[p] = r.topk_scores, r.topk_ids
    return PipelineResult(
        models, ks, ki, sum(r.n_overflow for r in res),
        dict(modeling=t1 - t0, exchange=t2 - t1, screening=t3 - t2, pockets_here=len(mine), models_screened=len(live)),
    )  # fmt: skip
