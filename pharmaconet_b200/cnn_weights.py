"""Seeded synthetic weights in the reference's state-dict layout (no trained `model.tar` exists offline:
src/pmnet/utils/download_weight.py needs the network). Every tensor is drawn from its own generator seeded by
(seed, key), so the reference modules (golden-vector script) and this package's model get identical values without
sharing any construction code. Scales keep activations O(1) through the ~40 layers (fan-in scaling, positive
BatchNorm variances); `initialize_weights()` of the reference is NOT used (it zeroes the Swin post-norms,
SURVEY appendix C-6). Data creation only."""

from __future__ import annotations

import math
import zlib

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1_000_003 + zlib.crc32(key.encode())) % (2**63 - 1))
    return g


def synth_tensor(key: str, shape, seed: int) -> torch.Tensor:
    g = _gen(seed, key)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_var":
        return torch.rand(shape, generator=g) + 0.5
    if leaf == "running_mean":
        return torch.randn(shape, generator=g) * 0.1
    if leaf == "logit_scale":
        return torch.full(shape, math.log(10.0)) + torch.randn(shape, generator=g) * 0.1
    if leaf in ("q_bias", "v_bias", "bias"):
        return torch.randn(shape, generator=g) * 0.1
    if leaf == "weight":
        if len(shape) == 1:  # LayerNorm / BatchNorm gain
            return torch.rand(shape, generator=g) + 0.5
        if "interaction_embedding" in key:
            return torch.rand(shape, generator=g) * 2.0 - 1.0
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
    raise KeyError(f"no synthetic rule for state-dict entry {key!r}")


def synth_state_dict(manifest: dict[str, list[int]], buffers: dict[str, torch.Tensor], seed: int = 0):
    """manifest: key -> shape for every learnable / running-stat entry; buffers: the constructor-computed Swin
    buffers (relative_coords_table, relative_position_index, attn_mask), stored once per distinct shape family."""
    sd = {}
    for key, shape in manifest.items():
        leaf = key.rsplit(".", 1)[-1]
        if leaf in ("relative_coords_table", "relative_position_index", "attn_mask"):
            sd[key] = buffers[buffer_name(key, shape)].clone()
        else:
            sd[key] = synth_tensor(key, shape, seed)
    return sd


def buffer_name(key: str, shape) -> str:
    leaf = key.rsplit(".", 1)[-1]
    return leaf + "_" + "x".join(str(int(s)) for s in shape)


def synth_checkpoint(manifest, buffers, seed: int = 0) -> dict:
    """A stand-in for the reference's model.tar (module.py:82-93): synthetic weights + seeded score distributions."""
    import numpy as np

    from .constants import INTERACTION_LIST

    rng = np.random.default_rng(seed + 17)
    dists = {typ: {"focus": np.sort(rng.uniform(0.2, 1.0, size=2000)).tolist()} for typ in INTERACTION_LIST}
    return {"config": {"MODEL": {}}, "model": synth_state_dict(manifest, buffers, seed), "score_distributions": dists}
