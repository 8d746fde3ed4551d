"""Host side of the batched scoring path: device-resident model / library containers and the C-ABI call.

`score_batch` is the batched equivalent of calling the reference's `PharmacophoreModel._scoring`
(src/pmnet/pharmacophore_model.py:101-106) once per ligand. torch is used for device memory and streams only.
"""

from __future__ import annotations

import ctypes as C
import warnings
from dataclasses import dataclass, replace

import numpy as np
import torch

from . import _abi, _lib
from .constants import weights_vector
from .packing import LigandBatch, PackedModel


def _require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("pharmaconet_b200 scores on CUDA devices only (no CPU fallback)")
    return device


class DeviceModel:
    """PackedModel resident in HBM + the PmModel struct pointing at it."""

    def __init__(self, model: PackedModel, device="cuda"):
        self.device = _require_cuda(device)
        self.host = model
        self.tensors = {k: torch.from_numpy(np.ascontiguousarray(v)).to(self.device) for k, v in model.arrays().items()}
        self.struct = _abi.model_struct(
            model.num_nodes,
            model.num_clusters,
            int(model.cluster_node_off[-1]),
            {k: t.data_ptr() for k, t in self.tensors.items()},
        )

    @property
    def num_nodes(self) -> int:
        return self.host.num_nodes

    @property
    def num_clusters(self) -> int:
        return self.host.num_clusters


class DeviceLigandBatch:
    """LigandBatch resident in HBM (or a view into pre-allocated staging tensors) + the PmLigandBatch struct."""

    def __init__(
        self,
        tensors: dict[str, torch.Tensor],
        n_ligands: int,
        n_conformers_total: int,
        bases: dict[str, int] | None = None,
        max_conformers: int | None = None,
    ):
        """bases: for a block of a larger library whose CSR offsets were not re-based (see include/pmnet_b200.h).
        max_conformers: largest n_conf in the batch (read back from the device when not given)."""
        self.tensors = tensors
        self.n_ligands = int(n_ligands)
        self.n_conformers_total = int(n_conformers_total)
        if max_conformers is None:
            max_conformers = int(tensors["n_conf"][: self.n_ligands].max().item()) if self.n_ligands else 1
        self.max_conformers = int(max_conformers)
        self.device = tensors["coords"].device
        self.struct = _abi.batch_struct(
            self.n_ligands, {k: tensors[k].data_ptr() for k in _abi.BATCH_FIELDS}, bases
        )
        self.order: torch.Tensor | None = None

    @classmethod
    def from_host(cls, batch: LigandBatch, device="cuda", non_blocking: bool = False) -> "DeviceLigandBatch":
        device = _require_cuda(device)
        t = {
            k: torch.from_numpy(np.ascontiguousarray(v)).to(device, non_blocking=non_blocking)
            for k, v in batch.arrays().items()
        }
        return cls(t, batch.num_ligands, batch.num_conformers_total, max_conformers=max(1, batch.max_conformers))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors.values())

    def set_order(self, order: torch.Tensor | None) -> None:
        """Processing order of the persistent grid (int32 [n_ligands] on the device) or None for index order."""
        self.order = order
        self.struct.order = order.data_ptr() if order is not None else None


@dataclass
class ScoreConfig:
    warps_per_block: int = 0  # 0 = library default
    blocks: int = 0
    scratch_rows: int = 0
    rescore_status: int = 0  # != 0: only ligands whose status holds this code are scored (in-place re-runs)
    heavy_budget: int = 0  # tree nodes after which a ligand's tree is split over many warps (0 default, < 0 never)

    def struct(self, max_conformers: int = 32) -> _abi.PmScoreConfig:
        return _abi.PmScoreConfig(
            self.warps_per_block, self.blocks, self.scratch_rows, int(max_conformers), int(self.rescore_status),
            int(self.heavy_budget),
        )


def launches_per_call(config: "ScoreConfig | None" = None, max_conformers: int = 32) -> int:
    """Kernels one pmnet_score_batch call enqueues (csrc/scoring.cu): the specialised kernel (default launch shape and
    <= 32 conformers only; models whose tables do not fit in shared memory skip it), then either the general kernel
    (heavy_budget < 0 or a status-restricted re-run) or three launches of the task kernel -
    the first serves the ligand queue like the general kernel and then the tasks its walkers donate, the others return
    at once when nothing is left - plus the kernel that finishes the heavy ligands."""
    cfg = config or ScoreConfig()
    custom = cfg.warps_per_block > 0 or cfg.blocks > 0 or cfg.scratch_rows > 0 or cfg.rescore_status != 0
    n = int(not custom and max_conformers <= 32)
    heavy = cfg.heavy_budget >= 0 and cfg.rescore_status == 0
    return n + (4 if heavy else 1)


def conf_stride(max_conformers: int) -> int:
    """Row length of the per-conformer output: 32 conformers per lane word, 1 / 2 / 4 words."""
    return 32 if max_conformers <= 32 else (64 if max_conformers <= 64 else 128)


def big_config(model) -> ScoreConfig:
    """A roomy configuration for ligands whose pair table overflowed the default per-warp scratch: rows for the
    worst case of this model (T = min(20 Km, 1024) entries, T^2/2 pairs), on fewer warps."""
    t = min(20 * model.num_clusters, 1024)
    rows = max(65536, t * t // 2 + t)  # >= 65536 rows also selects the 255-node distance table
    return ScoreConfig(warps_per_block=4, blocks=16, scratch_rows=rows)


def mid_config(model) -> ScoreConfig:
    """Four times the default pair-table scratch on half the warps: the first retry for ligands that overflowed the
    default scratch (wide models push many ligands past 8192 pair rows; `big_config` alone would serialise them on
    64 warps)."""
    return ScoreConfig(warps_per_block=16, blocks=0, scratch_rows=32768)


def warn_unscored(status: torch.Tensor, what: str) -> int:
    """Ligands that no configuration could score keep score 0; say so instead of returning a silent 0."""
    n_bad = int((status >= _abi.LIG_OVERFLOW).sum().item())
    if n_bad:
        warnings.warn(
            f"{what}: {n_bad} ligand(s) could not be scored (status OVERFLOW / UNSUPPORTED after the last retry: more "
            f"than {_abi.MAX_CONFORMERS} conformers, or a pair table beyond the largest scratch); their score is 0",
            RuntimeWarning,
            stacklevel=3,
        )
    return n_bad


def rescore_overflowed_device(
    model: "DeviceModel", batch: "DeviceLigandBatch", out: dict, weights=None, stream: torch.cuda.Stream | None = None
) -> int:
    """Re-run, IN PLACE on the device-resident `batch`, the ligands that a previous `score_batch` left with status
    OVERFLOW: `mid_config` first, `big_config` for what still does not fit (C-ABI PmScoreConfig.rescore_status: the
    kernel's queue skips every other ligand, so no host copy of the library and no compaction is needed).
    `out` = the dict returned by score_batch; its tensors are updated. Returns the number of ligands re-run."""
    n_over = int((out["status"] == _abi.LIG_OVERFLOW).sum().item())
    if n_over == 0:
        return 0
    for cfg in (mid_config(model), big_config(model)):
        score_batch(
            model, batch, weights, replace(cfg, rescore_status=_abi.LIG_OVERFLOW), out_scores=out["scores"],
            out_status=out["status"], out_stats=out.get("stats"), out_conf=out.get("conf"), stream=stream,
        )  # fmt: skip
        if not bool((out["status"] == _abi.LIG_OVERFLOW).any().item()):
            break
    return n_over


def rescore_overflowed(model: "DeviceModel", sub: LigandBatch, weights=None, with_stats: bool = False) -> dict:
    """Score a HOST batch of ligands whose pair table overflowed the default per-warp scratch (streamed screening:
    their block has left the device by the time the status is read). Returns device tensors like score_batch."""
    dev_sub = DeviceLigandBatch.from_host(sub, model.device)
    out = score_batch(model, dev_sub, weights, mid_config(model), with_stats=with_stats)
    rescore_overflowed_device(model, dev_sub, out, weights)
    return out


_workspaces: dict[tuple, torch.Tensor] = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(),)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def release_workspaces() -> None:
    _workspaces.clear()


def workspace_bytes(model: "DeviceModel", config: "ScoreConfig | None" = None, max_conformers: int = 32) -> int:
    cfg = (config or ScoreConfig()).struct(max_conformers)
    with torch.cuda.device(model.device):
        return int(_lib.lib().pmnet_score_workspace_bytes(model.num_nodes, model.num_clusters, C.byref(cfg)))


def score_batch(
    model: DeviceModel,
    batch: DeviceLigandBatch,
    weights: dict[str, float] | None = None,
    config: ScoreConfig | None = None,
    out_scores: torch.Tensor | None = None,
    out_status: torch.Tensor | None = None,
    with_stats: bool = False,
    with_conf: bool = False,
    stream: torch.cuda.Stream | None = None,
    workspace: torch.Tensor | None = None,
    out_stats: torch.Tensor | None = None,
    out_conf: torch.Tensor | None = None,
):
    """Enqueue one scoring launch on `stream` (default: torch's current stream). Returns a dict of device tensors:
    scores f32[n], status i32[n] (+ stats u32[n,4] -> {tree nodes, leaves, rows, pair entries}; conf f32[n,32])."""
    L = _lib.lib()
    dev = model.device
    n = batch.n_ligands
    if batch.max_conformers > _abi.MAX_CONFORMERS:
        raise ValueError(f"ligands with more than {_abi.MAX_CONFORMERS} conformers are not supported")
    cfg = (config or ScoreConfig()).struct(batch.max_conformers)
    with torch.cuda.device(dev):
        need = L.pmnet_score_workspace_bytes(model.num_nodes, model.num_clusters, C.byref(cfg))
        if workspace is None:
            ws = _workspace(dev, need)
        else:
            ws = workspace
            if ws.numel() < need:
                raise RuntimeError(f"workspace of {ws.numel()} bytes is too small ({need} needed)")
        scores = out_scores if out_scores is not None else torch.empty(n, dtype=torch.float32, device=dev)
        status = out_status if out_status is not None else torch.empty(n, dtype=torch.int32, device=dev)
        with_stats = with_stats or out_stats is not None
        with_conf = with_conf or out_conf is not None
        stats = out_stats if out_stats is not None else (torch.zeros((n, 4), dtype=torch.int32, device=dev) if with_stats else None)
        conf = out_conf
        if conf is None and with_conf:
            conf = torch.zeros((n, conf_stride(batch.max_conformers)), dtype=torch.float32, device=dev)
        w = (C.c_float * 7)(*weights_vector(weights))
        s = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = L.pmnet_score_batch(
            C.byref(model.struct), C.byref(batch.struct), w, scores.data_ptr(),
            conf.data_ptr() if with_conf else None, status.data_ptr(), stats.data_ptr() if with_stats else None,
            ws.data_ptr(), ws.numel(), C.byref(cfg), C.c_void_p(s.cuda_stream),
        )  # fmt: skip
        _lib.check(rc, "pmnet_score_batch")
    out = dict(scores=scores, status=status)
    if with_stats:
        out["stats"] = stats
    if with_conf:
        out["conf"] = conf
    return out


def score_library(
    model: DeviceModel,
    host_batch: LigandBatch,
    weights: dict[str, float] | None = None,
    config: ScoreConfig | None = None,
    with_stats: bool = False,
) -> dict[str, np.ndarray]:
    """Score a host-resident library chunk: H2D, one launch, in-place re-run of the overflowed ligands with the roomier
    configurations, D2H. Returns numpy arrays (scores f32, status i32[, stats])."""
    dev_batch = DeviceLigandBatch.from_host(host_batch, model.device)
    out = score_batch(model, dev_batch, weights, config, with_stats=with_stats)
    rescore_overflowed_device(model, dev_batch, out, weights)
    warn_unscored(out["status"], "score_library")
    scores = out["scores"].cpu().numpy()
    status = out["status"].cpu().numpy()
    stats = out["stats"].cpu().numpy().view(np.uint32) if with_stats else None
    res = dict(scores=scores, status=status)
    if with_stats:
        res["stats"] = stats
    return res


def order_workspace_bytes(n_ligands: int) -> int:
    return int(_lib.lib().pmnet_order_workspace_bytes(int(n_ligands)))


def cost_order(
    model: DeviceModel,
    batch: DeviceLigandBatch,
    stream: torch.cuda.Stream | None = None,
    out: torch.Tensor | None = None,
    workspace: torch.Tensor | None = None,
) -> torch.Tensor:
    """Longest-first processing order of `batch` for `model` (int32 [n_ligands]; C-ABI pmnet_cost_order): ligands
    sorted by decreasing pair-table size. With it the end-of-launch tail of the persistent grid all but disappears
    (a 131 072-ligand launch takes 1.23x its ideal time in index order, 1.004x in this order).
    Attach it with `batch.set_order(order)`."""
    L = _lib.lib()
    dev = batch.device
    n = batch.n_ligands
    with torch.cuda.device(dev):
        order = out if out is not None else torch.empty(max(1, n), dtype=torch.int32, device=dev)
        need = L.pmnet_order_workspace_bytes(n)
        ws = workspace if workspace is not None else torch.empty(need, dtype=torch.uint8, device=dev)
        if ws.numel() < need:
            raise RuntimeError(f"order workspace of {ws.numel()} bytes is too small ({need} needed)")
        s = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = L.pmnet_cost_order(
            C.byref(model.struct), C.byref(batch.struct), order.data_ptr(), ws.data_ptr(), ws.numel(),
            C.c_void_p(s.cuda_stream),
        )  # fmt: skip
        _lib.check(rc, "pmnet_cost_order")
        if stream is not None and workspace is None:
            ws.record_stream(stream)
    return order[:n]


def topk(scores: torch.Tensor, k: int, id_base: int = 0, stream: torch.cuda.Stream | None = None):
    """k best (score, id_base + index) of a device score vector, descending, ties by ascending id."""
    L = _lib.lib()
    dev = scores.device
    _require_cuda(dev)
    n = scores.numel()
    with torch.cuda.device(dev):
        need = L.pmnet_topk_workspace_bytes(n, k)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        out_s = torch.empty(k, dtype=torch.float32, device=dev)
        out_i = torch.empty(k, dtype=torch.int64, device=dev)
        s = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = L.pmnet_topk(
            scores.data_ptr(), n, int(id_base), int(k), out_s.data_ptr(), out_i.data_ptr(), ws.data_ptr(), need,
            C.c_void_p(s.cuda_stream),
        )  # fmt: skip
        _lib.check(rc, "pmnet_topk")
        # `ws` goes back to torch's stream-ordered caching allocator when it is dropped: it cannot be handed to
        # work on this stream before the kernels enqueued above have run. A foreign stream needs recording.
        if stream is not None:
            ws.record_stream(stream)
    return out_s, out_i
