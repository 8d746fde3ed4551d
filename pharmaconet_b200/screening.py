"""Library screening driver: the data-parallel loop of the reference's `screening.py:46-75`
(`Pool.map(model.scoring_file, files)` -> sort by score) on GPUs.

* one process per GPU; the library is cut into blocks of `block_ligands` ligands and block b belongs to rank
  b mod world (interleaving evens out the DFS-cost variance between regions of a library);
* host-resident libraries stream through three device staging slots (each with its own scratch; blocks alternate
  between two compute streams): the copies
  of the next blocks (pinned host -> HBM, on a copy stream) overlap the scoring kernels of the earlier ones, and the next
  block's warps back-fill the SMs while the previous block's longest ligands finish (measured on one B200, 1 M ligands:
  61.4 M conformers/s end to end with 2 slots of 131 072 ligands, 64.9 M with 3 slots of 262 144, 69.5 M with two
  compute streams instead of one per slot);
* every rank keeps the k best (score, ligand id) of its shard; the only collective is one all-gather of those
  k pairs per rank, followed by the same merge on every rank (descending score, ties by ascending id - what
  sorting the reference's full result list gives for its head).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _abi
from .packing import LigandBatch, PackedModel
from .scoring import (
    DeviceLigandBatch, DeviceModel, ScoreConfig, cost_order, launches_per_call, order_workspace_bytes as _lib_order_bytes,
    rescore_overflowed, rescore_overflowed_device, score_batch, topk, warn_unscored, workspace_bytes,
)  # fmt: skip


def shard_blocks(n_ligands: int, rank: int, world: int, block_ligands: int) -> list[tuple[int, int]]:
    """[begin, end) ligand ranges owned by `rank`: block b -> rank b mod world."""
    n_blocks = -(-n_ligands // block_ligands)
    return [
        (b * block_ligands, min(n_ligands, (b + 1) * block_ligands)) for b in range(rank, n_blocks, world)
    ]


def merge_topk(scores: torch.Tensor, ids: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
    """k best of candidate (score, id) pairs: descending score, ties by ascending id; padding ids (-1) last."""
    valid = ids >= 0
    s = torch.where(valid, scores, torch.full_like(scores, float("-inf")))
    big = torch.iinfo(torch.int64).max
    key_id = torch.where(valid, ids, torch.full_like(ids, big))
    o1 = torch.sort(key_id, stable=True).indices
    o2 = torch.sort(s[o1], descending=True, stable=True).indices
    order = o1[o2][:k]
    out_s, out_i = s[order], ids[order]
    if out_s.numel() < k:
        pad = k - out_s.numel()
        out_s = torch.cat([out_s, torch.full((pad,), float("-inf"), dtype=s.dtype, device=s.device)])
        out_i = torch.cat([out_i, torch.full((pad,), -1, dtype=ids.dtype, device=ids.device)])
    return out_s, out_i


def gather_topk(scores: torch.Tensor, ids: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
    """All-gather every rank's top-k and merge (identical result on every rank). No-op without a process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return merge_topk(scores, ids, k)
    world = dist.get_world_size()
    gs = torch.empty(world * k, dtype=scores.dtype, device=scores.device)
    gi = torch.empty(world * k, dtype=ids.dtype, device=ids.device)
    dist.all_gather_into_tensor(gs, scores.contiguous())
    dist.all_gather_into_tensor(gi, ids.contiguous())
    return merge_topk(gs, gi, k)


def pin_library(lib: LigandBatch) -> LigandBatch:
    """Copy a host library into page-locked memory (numpy views of pinned torch tensors) for async H2D."""
    arrays = {}
    keep = []
    for k, v in lib.arrays().items():
        t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
        keep.append(t)
        arrays[k] = t.numpy()
    out = LigandBatch.from_arrays(arrays)
    out._pinned = keep  # keep the owning tensors alive
    return out


@dataclass
class ScreenResult:
    topk_scores: torch.Tensor  # [k] fp32, device, descending
    topk_ids: torch.Tensor  # [k] int64 global ligand ids (-1 padding)
    scores: np.ndarray | torch.Tensor | None  # this rank's scores in processing order
    ids: np.ndarray | None  # global ligand id of each entry of `scores` (None = identity)
    n_ligands: int
    n_conformers: int
    n_overflow: int
    launches: int  # kernels of this package launched


class _Slot:
    """Device staging buffers for one in-flight block."""

    def __init__(self, caps: dict[str, int], dtypes: dict[str, torch.dtype], device):
        self.t = {k: torch.empty(max(1, caps[k]), dtype=dtypes[k], device=device) for k in caps}
        self.ready = torch.cuda.Event()  # H2D finished
        self.free = torch.cuda.Event()  # kernel finished, slot reusable
        self.workspace: torch.Tensor | None = None
        self.order: torch.Tensor | None = None  # longest-first processing order of the block in flight
        self.order_ws: torch.Tensor | None = None


class Screener:
    def __init__(
        self,
        model,
        device=None,
        weights: dict[str, float] | None = None,
        k: int = 1000,
        config: ScoreConfig | None = None,
        block_ligands: int = 262144,
        n_slots: int = 3,
        ramp: bool = True,
        stream_config: ScoreConfig | None = None,
        lpt: bool = True,
    ):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        packed = model if isinstance(model, PackedModel) else model.packed
        self.model = DeviceModel(packed, self.device)
        self.weights = weights
        self.k = int(k)
        self.config = config or ScoreConfig()
        self.block_ligands = int(block_ligands)
        self.n_slots = max(2, int(n_slots))  # device staging buffers (each with its own scratch)
        self.ramp = bool(ramp)
        self.ramp_cuts = (0.125, 0.25, 0.5)  # cumulative fractions of the first block where it is cut
        # hand ligands to the warps longest first (scoring.cost_order): removes the end-of-launch tail
        self.lpt = bool(lpt)
        # launch shape of the streamed path (None = the library default)
        self.stream_config = stream_config
        self._copy_stream = torch.cuda.Stream(self.device)
        # Blocks alternate between TWO compute streams (each slot has its own scratch), so that the next block's warps
        # back-fill the SMs while the previous block's longest ligands are still finishing (the tail of a persistent
        # grid). Not one stream per slot: the short kernels that end a scoring call (general kernel, task rounds) then
        # wait behind the persistent kernels other streams have queued - measured: a slot was released two blocks late,
        # its next copy could not start and the GPU idled 41 ms of a 494 ms pass. With two streams the kernel queued
        # behind a call's short kernels is always the next block of the SAME stream.
        self._compute_streams = [torch.cuda.Stream(self.device) for _ in range(2)]
        self._slots: list[_Slot] | None = None
        self._slot_caps: dict[str, int] | None = None
        # optional CUDA-event pairs around every scoring launch (bench.py: kernel duration for the roofline)
        self.record_kernel_events = False
        self.kernel_events: list[tuple[torch.cuda.Event, torch.cuda.Event]] = []

    def _timed_score(self, batch, config=None, **kw):
        config = config or self.config
        if not self.record_kernel_events:
            return score_batch(self.model, batch, self.weights, config, **kw)
        stream = kw.get("stream") or torch.cuda.current_stream(self.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = score_batch(self.model, batch, self.weights, config, **kw)
        e1.record(stream)
        self.kernel_events.append((e0, e1))
        return out

    # ------------------------------------------------------------------ device-resident shard: one launch
    def screen_device(self, batch: DeviceLigandBatch, id_base: int = 0, gather: bool = True) -> ScreenResult:
        # the scoring kernels + id fill + top-k write-out (the sort passes are CUB's)
        launches = launches_per_call(self.config, batch.max_conformers) + 2
        if self.lpt and batch.order is None:
            # once per resident library and model: the order only depends on topologies and the model's cluster types
            batch.set_order(cost_order(self.model, batch))
            launches += 1
        out = self._timed_score(batch)
        # ligands whose pair table overflowed the default scratch are re-run in place with the roomier configurations
        # (one device->host read of the overflow count; nothing else leaves the device)
        n_over = rescore_overflowed_device(self.model, batch, out, self.weights)
        if n_over:
            launches += 1
            warn_unscored(out["status"], "screen_device")
        ks, ki = topk(out["scores"], self.k, id_base)
        if gather:
            ks, ki = gather_topk(ks, ki, self.k)
        return ScreenResult(ks, ki, out["scores"], None, batch.n_ligands, batch.n_conformers_total, n_over, launches)

    # ------------------------------------------------------------------ host-resident library: streamed
    def _block_slices(self, lib: LigandBatch, a: int, b: int):
        n0, n1 = int(lib.lig_node_off[a]), int(lib.lig_node_off[b])
        q0, q1 = int(lib.lig_cluster_off[a]), int(lib.lig_cluster_off[b])
        c0, c1 = int(lib.cluster_node_off[q0]), int(lib.cluster_node_off[q1])
        x0, x1 = int(lib.coord_off[a]), int(lib.coord_off[b])
        sl = dict(
            lig_node_off=lib.lig_node_off[a : b + 1],
            lig_cluster_off=lib.lig_cluster_off[a : b + 1],
            cluster_node_off=lib.cluster_node_off[q0 : q1 + 1],
            cluster_nodes=lib.cluster_nodes[c0:c1],
            node_type_mask=lib.node_type_mask[n0:n1],
            n_conf=lib.n_conf[a:b],
            coord_off=lib.coord_off[a : b + 1],
            coords=lib.coords[x0:x1],
        )
        bases = dict(coord_base=x0, node_base=n0, cluster_base=q0, cnode_base=c0)
        return sl, bases

    def _ensure_slots(self, lib: LigandBatch, blocks):
        caps = {k: 0 for k in _abi.BATCH_FIELDS}
        for a, b in blocks:
            sl, _ = self._block_slices(lib, a, b)
            for k, v in sl.items():
                caps[k] = max(caps[k], v.shape[0])
        if self._slots is not None and all(self._slot_caps[k] >= caps[k] for k in caps):
            return
        dtypes = {k: torch.from_numpy(v[:0]).dtype for k, v in lib.arrays().items()}
        self._slots = [_Slot(caps, dtypes, self.device) for _ in range(self.n_slots)]
        self._slot_caps = caps

    def screen_host(
        self, lib: LigandBatch, rank: int = 0, world: int = 1, gather: bool = True, return_scores: bool = True
    ) -> ScreenResult:
        """Score this rank's blocks of a host library (pinned memory makes the copies asynchronous)."""
        dev = self.device
        blocks = shard_blocks(lib.num_ligands, rank, world, self.block_ligands)
        n_mine = sum(b - a for a, b in blocks)
        # ramp-up: nothing overlaps the copy of the very first span, so the first block is cut into growing pieces
        # (1/8, 1/8, 1/4, 1/2) - the kernel starts after an eighth of a block has arrived
        if self.ramp and blocks and blocks[0][1] - blocks[0][0] >= 8192:
            a0, b0 = blocks[0]
            n0 = b0 - a0
            cuts = [a0] + [a0 + int(n0 * f) for f in self.ramp_cuts] + [b0]
            blocks = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)] + blocks[1:]
        scores = torch.empty(n_mine, dtype=torch.float32, device=dev)
        status = torch.empty(n_mine, dtype=torch.int32, device=dev)
        cand_s, cand_i = [], []
        launches = 0
        n_conf = 0
        if blocks:
            self._ensure_slots(lib, blocks)
        main = torch.cuda.current_stream(dev)
        start = torch.cuda.Event()
        start.record(main)
        scfg = self.stream_config or self.config
        need = workspace_bytes(self.model, scfg, max(1, lib.max_conformers))
        pos = 0
        spans = []
        for it, (a, b) in enumerate(blocks):
            slot = self._slots[it % self.n_slots]
            if slot.workspace is None or slot.workspace.numel() < need:
                slot.workspace = torch.empty(need, dtype=torch.uint8, device=dev)
            sl, bases = self._block_slices(lib, a, b)
            with torch.cuda.stream(self._copy_stream):
                if it < self.n_slots:
                    self._copy_stream.wait_event(start)
                else:
                    self._copy_stream.wait_event(slot.free)
                views = {}
                for k, v in sl.items():
                    dst = slot.t[k][: v.shape[0]]
                    dst.copy_(torch.from_numpy(v), non_blocking=True)
                    views[k] = dst
                slot.ready.record(self._copy_stream)
            nb = b - a
            nc = int(lib.n_conf[a:b].sum())
            n_conf += nc
            # the library-wide maximum for every block: one kernel instantiation and one workspace layout per screen
            db = DeviceLigandBatch(views, nb, nc, bases, max_conformers=max(1, lib.max_conformers))
            cstream = self._compute_streams[it % 2]
            if it < 2:
                cstream.wait_event(start)
            cstream.wait_event(slot.ready)
            if self.lpt:
                if slot.order is None or slot.order.numel() < nb:
                    slot.order = torch.empty(max(nb, self.block_ligands), dtype=torch.int32, device=dev)
                    slot.order_ws = torch.empty(
                        int(_lib_order_bytes(max(nb, self.block_ligands))), dtype=torch.uint8, device=dev
                    )
                cost_order(self.model, db, stream=cstream, out=slot.order, workspace=slot.order_ws)
                db.set_order(slot.order)
            self._timed_score(
                db, config=scfg, out_scores=scores[pos : pos + nb], out_status=status[pos : pos + nb],
                stream=cstream, workspace=slot.workspace,
            )  # fmt: skip
            slot.free.record(cstream)
            spans.append((pos, nb, a))
            launches += launches_per_call(scfg, max(1, lib.max_conformers)) + int(self.lpt)
            pos += nb
        for slot in (self._slots or [])[: len(blocks)]:
            main.wait_event(slot.free)
        for p0, nb, a in spans:
            ks, ki = topk(scores[p0 : p0 + nb], self.k, a)
            cand_s.append(ks)
            cand_i.append(ki)
            launches += 2
        ids = np.concatenate([np.arange(a, b, dtype=np.int64) for a, b in blocks]) if blocks else np.zeros(0, np.int64)
        # ligands whose pair table overflowed the per-warp scratch: re-run them with the roomy configuration
        st = status.cpu().numpy()
        over = np.nonzero(st == _abi.LIG_OVERFLOW)[0]
        if len(over):
            sub = lib.select(ids[over])
            o2 = rescore_overflowed(self.model, sub, self.weights)
            warn_unscored(o2["status"], "screen_host")
            scores[torch.from_numpy(over).to(dev)] = o2["scores"]
            launches += 1
            over_ids = torch.from_numpy(ids[over]).to(dev)
            # drop the placeholder entries of these ligands from the per-block candidates
            cand_i = [torch.where(torch.isin(ci, over_ids), torch.full_like(ci, -1), ci) for ci in cand_i]
            cand_s.append(o2["scores"])
            cand_i.append(over_ids)
        if cand_s:
            ks, ki = merge_topk(torch.cat(cand_s), torch.cat(cand_i), self.k)
        else:
            ks = torch.full((self.k,), float("-inf"), dtype=torch.float32, device=dev)
            ki = torch.full((self.k,), -1, dtype=torch.int64, device=dev)
        if gather:
            ks, ki = gather_topk(ks, ki, self.k)
        host_scores = scores.cpu().numpy() if return_scores else None
        return ScreenResult(ks, ki, host_scores, ids, n_mine, n_conf, int(len(over)), launches)


def screen_models(
    models,
    batch: DeviceLigandBatch,
    host_lib: LigandBatch | None = None,
    weights: dict[str, float] | None = None,
    k: int = 1000,
    config: ScoreConfig | None = None,
    id_base: int = 0,
    gather: bool = True,
    keep_scores: bool = False,
) -> list[ScreenResult]:
    """Screen ONE device-resident library shard against MANY pharmacophore models (BASELINE configs[4]: the models
    of a batch of pockets against the same library; the reference runs `screening.py` once per model).

    The library stays in HBM; every model is one scoring launch plus a top-k, alternating between two streams with
    their own scratch so that the long-ligand tail of one model's persistent grid overlaps the start of the next.
    There is no host synchronisation until all models are enqueued. Ligands whose pair table overflowed the per-warp
    scratch are re-run in place on the resident shard with the roomier configurations (`host_lib` is accepted for
    backward compatibility and not needed any more); `n_overflow` reports how many.
    """
    dev = batch.device
    cfg = config or ScoreConfig()
    dms = [DeviceModel(m if isinstance(m, PackedModel) else m.packed, dev) for m in models]
    if not dms:
        return []
    main = torch.cuda.current_stream(dev)
    start = torch.cuda.Event()
    start.record(main)
    streams = [torch.cuda.Stream(dev) for _ in range(min(2, len(dms)))]
    spaces = []
    for i in range(len(streams)):
        need = max(workspace_bytes(dm, cfg, batch.max_conformers) for dm in dms[i :: len(streams)])
        spaces.append(torch.empty(need, dtype=torch.uint8, device=dev))
    outs = []
    prev_order = batch.order
    orders = []  # kept alive until the launches that read them have run
    for i, dm in enumerate(dms):
        st = streams[i % len(streams)]
        if i < len(streams):
            st.wait_event(start)
        with torch.cuda.stream(st):  # outputs are allocated on the stream that writes them
            # longest ligands first; the order depends on the model's cluster types, so it is per model. The batch
            # struct is copied into the launch, so re-pointing it between launches is safe.
            orders.append(cost_order(dm, batch, stream=st))
            batch.set_order(orders[-1])
            o = score_batch(dm, batch, weights, cfg, stream=st, workspace=spaces[i % len(streams)])
            ks, ki = topk(o["scores"], k, id_base, stream=st)
        for t in (o["scores"], o["status"], ks, ki):
            t.record_stream(main)  # read on the caller's stream below
        outs.append((o, ks, ki))
    batch.set_order(prev_order)
    for st in streams:
        main.wait_stream(st)
    results = []
    for dm, (o, ks, ki) in zip(dms, outs):
        # (first host sync of the call) overflowed ligands are re-run in place on the resident shard
        n_over = rescore_overflowed_device(dm, batch, o, weights)
        launches = launches_per_call(None, batch.max_conformers) + 2
        if n_over:
            warn_unscored(o["status"], "screen_models")
            ks, ki = topk(o["scores"], k, id_base)
            launches += 3
        if gather:
            ks, ki = gather_topk(ks, ki, k)
        results.append(
            ScreenResult(ks, ki, o["scores"] if keep_scores else None, None, batch.n_ligands, batch.n_conformers_total,
                         n_over, launches)
        )  # fmt: skip
    return results


def write_csv(path: str, names, scores) -> None:
    """The reference's output format (screening.py:70-75): `path,score`, descending by score."""
    order = np.lexsort((np.arange(len(scores)), -np.asarray(scores, dtype=np.float64)))
    with open(path, "w") as w:
        w.write("path,score\n")
        for i in order:
            w.write(f"{names[i]},{float(scores[i])}\n")
