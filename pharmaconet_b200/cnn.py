"""The reference's `PharmacoNetModel` (src/pmnet/network/detector.py:12-91) for inference on B200.

Same four entry points with the same tensor signatures - `forward_feature`, `forward_cavity_extraction`,
`forward_token_prediction`, `forward_segmentation` - built from a reference-format state dict (the `model` entry of
the reference's `model.tar`, module.py:82-85). The convolution stack (FPN decoder, cavity head, mask head: ~90 % of
the FLOPs) runs on the tcgen05 kernel of csrc/conv3d.cu with BatchNorm(eval) folded in; the glue between the
convolutions is csrc/pointwise.cu; the Swin backbone and the token MLPs are plain fp32 GEMMs through torch.
Activations between convolutions are bf16 in the 8-channel-chunk layout; accumulation is fp32.
"""

from __future__ import annotations

import ctypes as C
from collections.abc import Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib, conv, gemm
from .swin import SwinV2Backbone

_FEATURE_CHANNELS = (33, 96, 192, 384, 768)  # builder.py:27
_NUM_CONVS = (1, 2, 2, 2, 2)


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _ConvBN:
    """Folded BaseConv3d (nn/layers.py:4-46): conv weight + per-channel scale / bias."""

    def __init__(self, sd, prefix: str):
        w = sd[prefix + "_conv.weight"]
        self.kernel = w.shape[2]
        self.cin = w.shape[1]
        if prefix + "_norm.weight" in sd:
            self.scale, self.bias = conv.fold_bn(
                sd.get(prefix + "_conv.bias"), sd[prefix + "_norm.weight"], sd[prefix + "_norm.bias"],
                sd[prefix + "_norm.running_mean"], sd[prefix + "_norm.running_var"],
            )  # fmt: skip
        else:
            self.scale = torch.ones(w.shape[0], dtype=torch.float32, device=w.device)
            self.bias = sd[prefix + "_conv.bias"].float().contiguous()
        self.weight = w
        if self.kernel == 3 and self.cin == 96 and w.shape[0] == 96:
            self.packed = conv.pack_weights_k3(w)
            w_hi, w_lo = conv.split_bf16(w)  # two-term split for the split-precision mode
            self.packed_lo = conv.pack_weights_k3(w_lo)
            assert torch.equal(self.packed, conv.pack_weights_k3(w_hi))
        elif self.kernel == 1:
            self.w_t = w.reshape(w.shape[0], self.cin).t().contiguous().float()  # [C_in][C_out]


class Features(tuple):
    """multi-scale features, top-down, as NCDHW fp32 tensors (the reference's return type); `.c8` keeps the bf16
    chunked activations (conv.Act) the kernels consume so that later stages do not convert again."""

    c8: list


class PharmacoNetModel:
    def __init__(self, state_dict: dict[str, torch.Tensor], device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pharmaconet_b200.cnn runs on CUDA devices only (no CPU fallback)")
        sd = {k: v.to(self.device) for k, v in state_dict.items()}
        self.sd = sd
        self.backbone = SwinV2Backbone(sd, "embedding.backbone.")
        self.num_interactions = sd["token_head.interaction_embedding.weight"].shape[0]
        dec = "embedding.decoder."
        self.fpn_lateral = [_ConvBN(sd, f"{dec}lateral_conv_list.{l}.") if l < 4 else None for l in range(5)]
        self.fpn_convs = [[_ConvBN(sd, f"{dec}fpn_convs_list.{l}.{i}.") for i in range(_NUM_CONVS[l])] for l in range(5)]
        self.cavity = {
            name: (_ConvBN(sd, f"cavity_head.{name}.0."), _ConvBN(sd, f"cavity_head.{name}.1."))
            for name in ("short_head", "long_head")
        }
        mdec = "mask_head.decoder."
        self.mask_lateral = [_ConvBN(sd, f"{mdec}lateral_conv_list.{l}.") if l < 4 else None for l in range(5)]
        self.mask_convs = [[_ConvBN(sd, f"{mdec}fpn_convs_list.{l}.{i}.") for i in range(_NUM_CONVS[l])] for l in range(5)]
        self.mask_logit_w = sd["mask_head.conv_logits.weight"].reshape(96).float().contiguous()
        self.mask_logit_b = float(sd["mask_head.conv_logits.bias"].item())
        self._L = _lib.lib()
        self._wops: dict = {}
        # "bf16": every convolution is one tcgen05 pass on bf16 operands (fastest; ~2^-8 relative error per layer, a
        # few hundred of 262144 mask voxels differ from the fp32 reference). "bf16x3": activations and weights of the
        # convolution stack travel as two-term bf16 splits and every convolution is three passes (~2^-16 relative
        # error per product): the precision mode for outputs that feed a threshold (module.py:232-233, 288).
        self._precision = "bf16"
        self.precision = "bf16"

    @property
    def split(self) -> bool:
        return self.precision == "bf16x3"

    @property
    def precision(self) -> str:
        return self._precision

    @precision.setter
    def precision(self, value: str) -> None:
        if value not in ("bf16", "bf16x3"):
            raise ValueError("precision must be 'bf16' or 'bf16x3'")
        self._precision = value
        self.backbone.precision = value  # the backbone's linear layers follow (csrc/gemm.cu)

    def _linear(self, x: torch.Tensor, key, weight: torch.Tensor, bias: torch.Tensor | None = None, act: int = 0):
        """fp32 [M, K] x weight[N, K]^T (+ bias, activation) on the tcgen05 GEMM in the model's precision -> fp32 [M, N].
        The weight operand is split / converted once and cached under `key`."""
        ck = (key, self.split)
        w = self._wops.get(ck)
        if w is None:
            w = self._wops[ck] = gemm.Operand.from_float(weight, self.split)
        n = weight.shape[0]
        if n % 32:  # e.g. the 1-wide score layer: pad the outputs to the GEMM's granularity and slice
            npad = (n + 31) // 32 * 32
            ck = (key, self.split, "pad")
            wp = self._wops.get(ck)
            if wp is None:
                wpad = torch.zeros((npad, weight.shape[1]), dtype=torch.float32, device=weight.device)
                wpad[:n] = weight
                bpad = torch.zeros(npad, dtype=torch.float32, device=weight.device)
                if bias is not None:
                    bpad[:n] = bias
                wp = self._wops[ck] = (gemm.Operand.from_float(wpad, self.split), bpad)
            y, _ = gemm.linear(gemm.Operand.from_float(x, self.split), wp[0], wp[1], act)
            return y[:, :n]
        y, _ = gemm.linear(gemm.Operand.from_float(x, self.split), w, bias.float().contiguous() if bias is not None else None, act)
        return y

    # ------------------------------------------------------------------ kernels
    def _k3(self, x: "conv.Act", layer: _ConvBN, head=None, store_out=True):
        hw, hb = head if head is not None else (None, 0.0)
        if x.lo is not None:
            return conv.conv3d_k3_c96_x3(x, layer.packed, layer.packed_lo, layer.scale, layer.bias, True, hw, hb, store_out)
        y, h = conv.conv3d_k3_c96(x.hi, layer.packed, layer.scale, layer.bias, True, hw, hb, store_out)
        return (conv.Act(y) if y is not None else None), h

    def _lateral(self, x, x_is_c8: bool, layer: _ConvBN, up: "conv.Act | None", affine=True, split=None) -> "conv.Act":
        split = self.split if split is None else split
        x_lo = None
        if x_is_c8:
            x, x_lo = x.hi, x.lo
            B, _, D, H, W, _ = x.shape
        else:
            B, _, D, H, W = x.shape
            x = x.contiguous().float()
            if D * H * W <= 4096 and affine:
                # 8^3 / 16^3 levels with 384 / 192 input channels: < 0.2 GFLOP, but a 75-150 KB weight tile per CTA for
                # the fused CUDA-core kernel - here the 1x1 conv is the tcgen05 GEMM over the voxels
                xt = x.reshape(B, layer.cin, -1).transpose(1, 2).reshape(-1, layer.cin)  # [B * V, C_in]
                y = self._linear(xt, ("lat", id(layer)), layer.w_t.t()).view(B, D, H, W, 96).permute(0, 4, 1, 2, 3)
                y = torch.relu(y * layer.scale.view(1, -1, 1, 1, 1) + layer.bias.view(1, -1, 1, 1, 1))
                if up is not None:
                    y = y + F.interpolate(up.float_ncdhw(), scale_factor=2, mode="nearest")
                return conv.to_act(y, split)
        out = torch.empty((B, 12, D, H, W, 8), dtype=torch.bfloat16, device=self.device)
        out_lo = torch.empty_like(out) if split else None
        rc = self._L.pmnet_lateral_c96_split(
            x.data_ptr(), x_lo.data_ptr() if x_lo is not None else None, int(x_is_c8), layer.cin, layer.w_t.data_ptr(),
            layer.scale.data_ptr() if affine else None, layer.bias.data_ptr() if affine else None, int(affine),
            up.hi.data_ptr() if up is not None else None, up.lo.data_ptr() if (up is not None and up.lo is not None) else None,
            out.data_ptr(), out_lo.data_ptr() if split else None, B, D, H, W, _stream(self.device),
        )  # fmt: skip
        _lib.check(rc, "pmnet_lateral_c96_split")
        return conv.Act(out, out_lo)

    def _combine(self, s: "conv.Act", u, pvec, pvox, layer: _ConvBN | None, up: "conv.Act | None") -> "conv.Act":
        nbox = u.shape[0]
        _, D, H, W, _ = s.hi.shape
        split = s.lo is not None
        out = torch.empty((nbox, 12, D, H, W, 8), dtype=torch.bfloat16, device=self.device)
        out_lo = torch.empty_like(out) if split else None
        rc = self._L.pmnet_box_combine_c96_split(
            s.hi.data_ptr(), s.lo.data_ptr() if split else None, u.data_ptr(), pvec.data_ptr(), pvox.data_ptr(),
            layer.scale.data_ptr() if layer is not None else None, layer.bias.data_ptr() if layer is not None else None,
            int(layer is not None), up.hi.data_ptr() if up is not None else None,
            up.lo.data_ptr() if (up is not None and up.lo is not None) else None, out.data_ptr(),
            out_lo.data_ptr() if split else None, nbox, D, H, W, _stream(self.device),
        )  # fmt: skip
        _lib.check(rc, "pmnet_box_combine_c96_split")
        return conv.Act(out, out_lo)

    # ------------------------------------------------------------------ detector.py:36-43
    @torch.no_grad()
    def forward_feature(self, in_image: torch.Tensor, nchw: bool = True) -> Features:
        """nchw=False skips the fp32 NCDHW copies of the five maps (100 MB per pocket at 64^3) and returns the bf16
        chunked tensors themselves - enough for the other three entry points of this class."""
        image = in_image.to(self.device, torch.float32).contiguous()
        bottom_up = [image, *self.backbone.forward(image)]
        # FPNDecoder.forward (decoders/fpn_decoder.py:86-115), top (4^3) to bottom (64^3)
        top = self.fpn_convs[4][0]
        # 768 -> 96, k = 3 at 4^3 (0.25 GFLOP): im2col + one fp32 GEMM (a cuDNN fp32 convolution would silently run on
        # TF32 tensor cores and put 1e-3 of relative error into every level below)
        x4 = bottom_up[4]
        B4, C4, D4 = x4.shape[0], x4.shape[1], x4.shape[2]
        cols = F.pad(x4, (1, 1, 1, 1, 1, 1)).unfold(2, 3, 1).unfold(3, 3, 1).unfold(4, 3, 1)  # [B, C, D, H, W, 3, 3, 3]
        cols = cols.permute(0, 2, 3, 4, 1, 5, 6, 7).reshape(B4 * D4**3, C4 * 27)
        y = self._linear(cols, "fpn_top", top.weight.reshape(96, -1)).view(B4, D4, D4, D4, 96).permute(0, 4, 1, 2, 3)
        y = torch.relu(y * top.scale.view(1, -1, 1, 1, 1) + top.bias.view(1, -1, 1, 1, 1))
        fpn = self._k3(conv.to_act(y, self.split), self.fpn_convs[4][1])[0]
        outs = [fpn]
        for level in (3, 2, 1, 0):
            fpn = self._lateral(bottom_up[level], False, self.fpn_lateral[level], fpn)
            for layer in self.fpn_convs[level]:
                fpn = self._k3(fpn, layer)[0]
            outs.append(fpn)
        if not nchw:
            feats = Features(outs)
            feats.c8 = outs
            return feats
        full = []
        for o in outs:
            t = o.float_ncdhw()
            t._pm_c8 = o  # the bf16 chunked twin travels with the tensor: later stages skip the conversion
            full.append(t)
        feats = Features(full)
        feats.c8 = outs
        return feats

    def _as_c8(self, t) -> "conv.Act":
        """NCDHW feature tensor -> conv.Act (free when the tensor came out of forward_feature)."""
        if isinstance(t, conv.Act):
            return t
        twin = getattr(t, "_pm_c8", None)
        if twin is not None:
            return twin
        if t.dim() == 6:
            return conv.Act(t)
        return conv.to_act(t.to(self.device), self.split)

    # ------------------------------------------------------------------ detector.py:45-53, cavity_head.py:45-60
    @torch.no_grad()
    def forward_cavity_extraction(self, features) -> tuple[torch.Tensor, torch.Tensor]:
        x = self._as_c8(features)
        outs = []
        for name in ("short_head", "long_head"):
            k3, k1 = self.cavity[name]
            hw = k1.weight.reshape(96).float().contiguous()
            _, logits = self._k3(x, k3, head=(hw, float(k1.bias.item())), store_out=False)
            outs.append(logits.unsqueeze(1))
        return outs[0], outs[1]

    # ------------------------------------------------------------------ detector.py:55-71, token_head.py:50-86
    @torch.no_grad()
    def forward_token_prediction(self, features, tokens_list: Sequence[torch.Tensor]):
        sd = self.sd
        x = self._as_c8(features)
        toks = [t.to(self.device, torch.long) for t in tokens_list]
        counts = [int(t.shape[0]) for t in toks]
        if sum(counts) == 0:
            return (
                [torch.empty((0,), dtype=torch.float32, device=self.device) for _ in toks],
                [torch.empty((0, 192), dtype=torch.float32, device=self.device) for _ in toks],
            )
        # all pockets' tokens through the MLPs at once (the reference loops over images, token_head.py:62-66)
        allt = torch.cat(toks, 0)
        bidx = torch.repeat_interleave(torch.arange(len(toks), device=self.device), torch.tensor(counts, device=self.device))
        voxel = x.hi[bidx, :, allt[:, 0], allt[:, 1], allt[:, 2], :].reshape(-1, 96).float()
        if x.lo is not None:
            voxel = voxel + x.lo[bidx, :, allt[:, 0], allt[:, 1], allt[:, 2], :].reshape(-1, 96).float()
        h0 = torch.cat([voxel, sd["token_head.interaction_embedding.weight"][allt[:, 3]]], dim=1)
        def lin(x, name, act=0):
            return self._linear(x, name, sd[name + ".weight"], sd[name + ".bias"], act)

        skip = lin(h0, "token_head.skip") if "token_head.skip.weight" in sd else h0
        h = h0
        for i in (0, 2, 4):
            h = lin(h, f"token_head.feature_mlp.{i}", gemm.ACT_SILU)
        tf = skip + h
        s = tf
        for i in (0, 2):
            s = lin(s, f"token_head.score_mlp.{i}", gemm.ACT_RELU)
        s = lin(s, "token_head.score_mlp.4").squeeze(-1)
        return list(torch.split(s, counts)), list(torch.split(tf, counts))

    # ------------------------------------------------------------------ detector.py:73-91, mask_head.py:38-196
    def _shared_laterals(self, features, b: int):
        """conv1x1 of the pocket's feature maps for the mask-head decoder: linear, shared by every box of pocket b."""
        cache = getattr(features, "_mask_lateral_cache", None)
        if cache is None:
            cache = {}
            try:
                features._mask_lateral_cache = cache
            except AttributeError:
                pass
        if b not in cache:
            per_level = []
            for level in range(5):  # bottom-up level = 4 - top-down index
                f = self._as_c8(features[4 - level])[b : b + 1]
                per_level.append(
                    f[0] if level == 4 else self._lateral(f, True, self.mask_lateral[level], None, affine=False, split=f.lo is not None)[0]
                )
            cache[b] = per_level
        return cache[b]

    @torch.no_grad()
    def forward_segmentation(
        self, multi_scale_features, box_tokens_list, box_token_features_list, return_aux=False, group_size=None
    ):
        """mask_head.py:38-80. The reference adds every box's point feature at the token voxels of ALL boxes passed
        in the same call (mask_head.py:190-194) and its caller passes groups of 4 (module.py:261-272). `group_size`
        reproduces that grouping inside ONE call over all hotspots of a pocket (boxes i..i+group_size-1 form a
        group), which lets the convolutions run at full batch; None = the whole list is one group, exactly like the
        reference called with that list."""
        if return_aux:
            raise NotImplementedError("auxiliary multi-scale masks are a training-time output of the reference")
        sd = self.sd
        out_masks = []
        for b, (tokens, tfeat) in enumerate(zip(box_tokens_list, box_token_features_list)):
            tokens = tokens.to(self.device, torch.long)
            nbox = tokens.shape[0]
            size = self._as_c8(multi_scale_features[4]).shape[2]
            if nbox == 0:
                out_masks.append(torch.empty((0, size, size, size), dtype=torch.float32, device=self.device))
                continue
            gs = nbox if group_size is None else int(group_size)
            if gs > 4:
                raise NotImplementedError("groups of more than 4 boxes are not supported by the combine kernel")
            tfeat = tfeat.to(self.device, torch.float32)
            shared = self._shared_laterals(multi_scale_features, b)
            # token voxels of every box's group, per level, built on the host in one go (tokens are a few dozen
            # rows): pvox_levels[level][j, k] = flat voxel of the k-th member of box j's group, -1 when absent
            tok_np = tokens.cpu().numpy()
            member = (np.arange(nbox) // gs) * gs
            src = member[:, None] + np.arange(4)[None, :]
            ok = (np.arange(4)[None, :] < gs) & (src < nbox)
            src = np.where(ok, src, 0)
            pv = np.empty((5, nbox, 4), dtype=np.int32)
            for level in range(5):
                D = shared[level].shape[1]
                div = size // D
                vox = ((tok_np[:, 0] // div) * D + tok_np[:, 1] // div) * D + tok_np[:, 2] // div
                pv[level] = np.where(ok, vox[src], -1)
            pvox_levels = torch.from_numpy(pv).to(self.device)
            pieces = []
            for lo in range(0, nbox, 48):  # bound the per-call activation memory (48 boxes x 50 MB at 64^3)
                hi = min(nbox, lo + 48)
                fpn = None
                for level in (4, 3, 2, 1, 0):
                    s = shared[level]
                    pvox = pvox_levels[level, lo:hi].contiguous()
                    bgn, ptn = f"mask_head.background_mlp_list.{level}", f"mask_head.point_mlp_list.{level}"
                    bg = self._linear(tfeat[lo:hi], bgn, sd[bgn + ".weight"], sd[bgn + ".bias"])
                    pt = self._linear(tfeat[lo:hi], ptn, sd[ptn + ".weight"], sd[ptn + ".bias"])
                    lat = self.mask_lateral[level]
                    if lat is not None:  # push the per-box vectors through the (linear) 1x1 conv
                        wl = lat.weight.reshape(96, 96).float()
                        bg = self._linear(bg, ("mask_lat", level), wl)
                        pt = self._linear(pt, ("mask_lat", level), wl)
                    fpn = self._combine(s, bg.contiguous(), pt.contiguous(), pvox, lat, fpn)
                    convs = self.mask_convs[level]
                    for i, layer in enumerate(convs):
                        if level == 0 and i == len(convs) - 1:
                            _, logits = self._k3(fpn, layer, head=(self.mask_logit_w, self.mask_logit_b), store_out=False)
                        else:
                            fpn = self._k3(fpn, layer)[0]
                pieces.append(logits)
            out_masks.append(pieces[0] if len(pieces) == 1 else torch.cat(pieces, 0))
        return out_masks, None


def density_post(logits, tokens, protein_mask, cavity_mask, threshold: float = 0.5):
    """module.py:277-288 on the device: sigmoid, box/protein/cavity mask, 5^3 Gaussian, mask, threshold."""
    import math

    dev = logits.device
    n, S = logits.shape[0], logits.shape[-1]
    out = torch.empty((n, S, S, S), dtype=torch.float32, device=dev)
    if n == 0:
        return out
    w = [math.exp(-0.5 * (d / 0.5) ** 2) for d in (2, 1, 0)]
    tot = 2 * w[0] + 2 * w[1] + w[2]
    taps = (C.c_float * 3)(*[v / tot for v in w])
    lg = logits.contiguous().float()
    tk = tokens.to(dev, torch.int32).contiguous()
    pm = protein_mask.to(dev).reshape(-1).to(torch.uint8).contiguous()
    cm = cavity_mask.to(dev).reshape(-1).to(torch.uint8).contiguous()
    rc = _lib.lib().pmnet_density_post(
        lg.data_ptr(), tk.data_ptr(), pm.data_ptr(), cm.data_ptr(), taps, C.c_float(threshold), out.data_ptr(), n, S,
        _stream(dev),
    )  # fmt: skip
    _lib.check(rc, "pmnet_density_post")
    return out
