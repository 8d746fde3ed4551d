"""3-D Swin Transformer V2 backbone of the reference network, written functionally on top of a state dict.

Reference: src/pmnet/network/backbones/swinv2.py (WindowAttention :20-163, SwinTransformerBlock :166-311,
PatchMerging :314-363, PatchEmbed :450-500, SwinTransformerV2.forward :626-644) with the fixed configuration of
builder.py:15-24 (patch 2, dim 96, depths (2,6,2,2), heads (3,6,12,24), window 4). It is ~10 % of the forward FLOPs
(SURVEY appendix B). On the device every linear layer (qkv, proj, fc1 + GELU, fc2, PatchMerging reduction, PatchEmbed
as a GEMM over 2^3 patches) is the tcgen05 GEMM of csrc/gemm.cu and the rest of a block is two kernels of
csrc/swin_ops.cu (window attention; LayerNorm + residual, which also emits the next GEMM's operand). Reference quirks
kept: the cyclic shift rolls only the first two spatial axes (swinv2.py:277,296) while the stored attention mask covers
three; the continuous position bias is 16*sigmoid(cpb_mlp(table)); logit scale clamped at log(100); res-post-norm.

precision: "bf16x3" (default) - GEMM operands as two-term bf16 splits, three tensor-core passes, fp32 attention;
           "bf16"   - one pass on bf16 operands, bf16 qkv / attention on mma.sync tensor cores (fastest);
           "fp32"   - the linear layers through torch (library GEMMs): the CHECKER of the two modes above, also what the
                      op-by-op statement (`fused = False`, CPU) uses.
"""

from __future__ import annotations

import ctypes as C_
import math

import torch
import torch.nn.functional as F

DEPTHS = (2, 6, 2, 2)
HEADS = (3, 6, 12, 24)
WINDOW = 4
EMBED = 96
PATCH = 2


def _window_partition(x: torch.Tensor, ws: int) -> torch.Tensor:
    B, D, H, W, C = x.shape
    x = x.view(B, D // ws, ws, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).reshape(-1, ws * ws * ws, C)


def _window_reverse(win: torch.Tensor, ws: int, B: int, D: int, H: int, W: int) -> torch.Tensor:
    x = win.view(B, D // ws, H // ws, W // ws, ws, ws, ws, -1)
    return x.permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(B, D, H, W, -1)


class SwinV2Backbone:
    def __init__(self, sd: dict[str, torch.Tensor], prefix: str = "embedding.backbone.", image_size: int = 64):
        self.sd = sd
        self.p = prefix
        self.res0 = image_size // PATCH
        self._bias_cache: dict[str, torch.Tensor] = {}
        self.precision = "bf16x3"  # see the module docstring
        self._wop: dict = {}
        self._w16: dict[str, torch.Tensor] = {}
        # fused = True (default on CUDA): the non-GEMM part of every block runs in two kernels of csrc/swin_ops.cu
        # (window attention incl. shift / partition / bias / mask / softmax, and LayerNorm + residual);
        # fused = False is the op-by-op statement of the reference used for CPU checks
        self.fused = True
        self._scale_cache: dict[str, torch.Tensor] = {}

    def _lin(self, x: torch.Tensor, wname: str, bias: torch.Tensor | None = None) -> torch.Tensor:
        if self.precision == "bf16":
            w = self._w16.get(wname)
            if w is None:
                w = self._w16[wname] = self._g(wname).to(torch.bfloat16)
            y = F.linear(x.to(torch.bfloat16), w).float()
            return y if bias is None else y + bias
        return F.linear(x, self._g(wname), bias)

    def _g(self, name: str) -> torch.Tensor:
        if name.endswith("#2d"):  # a convolution weight viewed as the [C_out, C_in * k^3] matrix of its GEMM form
            t = self._scale_cache.get(name)
            if t is None:
                w = self.sd[self.p + name[:-3]]
                t = self._scale_cache[name] = w.reshape(w.shape[0], -1).contiguous()
            return t
        return self.sd[self.p + name]

    def _rel_bias(self, blk: str, heads: int, n: int) -> torch.Tensor:
        """16 * sigmoid(cpb_mlp(relative_coords_table))[relative_position_index] -> [heads, n, n]; input independent."""
        if blk not in self._bias_cache:
            # input independent (a 343 x 3 table through a 3 -> 512 -> heads MLP): evaluated ONCE on the host when the
            # block is first used, then kept on the device - no library GEMM on the forward path
            dev = self._g(blk + "attn.relative_coords_table").device
            c = lambda name: self._g(blk + name).detach().float().cpu()  # noqa: E731
            h = F.relu(F.linear(c("attn.relative_coords_table"), c("attn.cpb_mlp.0.weight"), c("attn.cpb_mlp.0.bias")))
            t = F.linear(h, c("attn.cpb_mlp.2.weight")).view(-1, heads)
            idx = self._g(blk + "attn.relative_position_index").view(-1).cpu()
            bias = t[idx].view(n, n, heads).permute(2, 0, 1).contiguous()
            self._bias_cache[blk] = (16.0 * torch.sigmoid(bias)).to(dev)
        return self._bias_cache[blk]

    def _attention(self, xw: torch.Tensor, blk: str, heads: int, mask: torch.Tensor | None) -> torch.Tensor:
        Bw, N, C = xw.shape
        qb, vb = self._g(blk + "attn.q_bias"), self._g(blk + "attn.v_bias")
        qkv = self._lin(xw, blk + "attn.qkv.weight", torch.cat((qb, torch.zeros_like(vb), vb)))
        qkv = qkv.reshape(Bw, N, 3, heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
        scale = torch.clamp(self._g(blk + "attn.logit_scale"), max=math.log(1.0 / 0.01)).exp()
        attn = attn * scale + self._rel_bias(blk, heads, N).unsqueeze(0)
        if mask is not None:
            nW = mask.shape[0]
            attn = (attn.view(Bw // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
        attn = torch.softmax(attn, dim=-1)
        out = (attn @ v).transpose(1, 2).reshape(Bw, N, C)
        return self._lin(out, blk + "attn.proj.weight", self._g(blk + "attn.proj.bias"))

    def _block(self, x: torch.Tensor, blk: str, res: int, heads: int, shift: int) -> torch.Tensor:
        if self.fused and x.is_cuda:
            return self._block_fused(x, blk, res, heads, shift)
        B, L, C = x.shape
        ws = WINDOW
        if res <= ws:  # swinv2.py:206-209
            ws, shift = res, 0
        h = x.view(B, res, res, res, C)
        if shift > 0:
            h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
        mask = self.sd.get(self.p + blk + "attn_mask") if shift > 0 else None
        aw = self._attention(_window_partition(h, ws), blk, heads, mask)
        h = _window_reverse(aw, ws, B, res, res, res)
        if shift > 0:
            h = torch.roll(h, shifts=(shift, shift), dims=(1, 2))
        h = h.reshape(B, L, C)
        x = x + F.layer_norm(h, (C,), self._g(blk + "norm1.weight"), self._g(blk + "norm1.bias"))
        m = self._lin(x, blk + "mlp.fc1.weight", self._g(blk + "mlp.fc1.bias"))
        m = self._lin(F.gelu(m), blk + "mlp.fc2.weight", self._g(blk + "mlp.fc2.bias"))
        return x + F.layer_norm(m, (C,), self._g(blk + "norm2.weight"), self._g(blk + "norm2.bias"))

    # ------------------------------------------------------------------ fused path (csrc/swin_ops.cu)
    def _gemm(self, x: torch.Tensor, wname: str, bias: torch.Tensor | None, keep16: bool = False) -> torch.Tensor:
        """Plain library GEMM; bf16 operands when precision == 'bf16' (output stays bf16 if keep16)."""
        if self.precision == "bf16":
            w = self._w16.get(wname)
            if w is None:
                w = self._w16[wname] = self._g(wname).to(torch.bfloat16)
            b16 = None
            if bias is not None:
                b16 = self._w16.get(wname + "#b")
                if b16 is None:
                    b16 = self._w16[wname + "#b"] = bias.to(torch.bfloat16)
            y = F.linear(x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16), w, b16)
            return y if keep16 else y.float()
        return F.linear(x, self._g(wname), bias)

    def _ln_res(self, shortcut: torch.Tensor | None, h: torch.Tensor, wname: str, bname: str, out=None) -> torch.Tensor:
        from . import _lib

        rows, C = h.numel() // h.shape[-1], h.shape[-1]
        h = h.contiguous()
        y = out if out is not None else torch.empty(h.shape, dtype=torch.float32, device=h.device)
        rc = _lib.lib().pmnet_ln_residual(
            shortcut.data_ptr() if shortcut is not None else None, h.data_ptr(), int(h.dtype == torch.bfloat16),
            self._g(wname).data_ptr(), self._g(bname).data_ptr(), y.data_ptr(), rows, C, C_.c_float(1e-5),
            C_.c_void_p(torch.cuda.current_stream(h.device).cuda_stream),
        )  # fmt: skip
        _lib.check(rc, "pmnet_ln_residual")
        return y

    def _block_fused(self, x: torch.Tensor, blk: str, res: int, heads: int, shift: int) -> torch.Tensor:
        from . import _lib

        B, L, C = x.shape
        if res <= WINDOW:
            shift = 0
        qb, vb = self._g(blk + "attn.q_bias"), self._g(blk + "attn.v_bias")
        bias = self._scale_cache.get(blk + "qkvb")
        if bias is None:
            bias = self._scale_cache[blk + "qkvb"] = torch.cat((qb, torch.zeros_like(vb), vb))
            self._scale_cache[blk + "scale"] = (
                torch.clamp(self._g(blk + "attn.logit_scale"), max=math.log(1.0 / 0.01)).exp().reshape(-1).contiguous()
            )
        qkv = self._gemm(x, blk + "attn.qkv.weight", bias, keep16=True).contiguous()
        attn = torch.empty((B, L, C), dtype=qkv.dtype, device=x.device)
        mask = self.sd.get(self.p + blk + "attn_mask") if shift > 0 else None
        rc = _lib.lib().pmnet_window_attention(
            qkv.data_ptr(), attn.data_ptr(), self._scale_cache[blk + "scale"].data_ptr(),
            self._rel_bias(blk, heads, 64).data_ptr(), mask.data_ptr() if mask is not None else None,
            B, res, shift, heads, int(qkv.dtype == torch.bfloat16),
            C_.c_void_p(torch.cuda.current_stream(x.device).cuda_stream),
        )  # fmt: skip
        _lib.check(rc, "pmnet_window_attention")
        proj = self._gemm(attn, blk + "attn.proj.weight", self._g(blk + "attn.proj.bias"), keep16=True)
        x = self._ln_res(x, proj, blk + "norm1.weight", blk + "norm1.bias")
        m = self._gemm(x, blk + "mlp.fc1.weight", self._g(blk + "mlp.fc1.bias"), keep16=True)
        m = self._gemm(F.gelu(m), blk + "mlp.fc2.weight", self._g(blk + "mlp.fc2.bias"), keep16=True)
        return self._ln_res(x, m, blk + "norm2.weight", blk + "norm2.bias", out=x)

    # ------------------------------------------------------------------ tcgen05 path (csrc/gemm.cu + csrc/swin_ops.cu)
    @property
    def _split(self) -> bool:
        return self.precision == "bf16x3"

    def _w(self, name: str):
        """weight [N, K] as a GEMM operand of the current precision (cached)"""
        from . import gemm

        key = (name, self._split)
        op = self._wop.get(key)
        if op is None:
            op = self._wop[key] = gemm.Operand.from_float(self._g(name), self._split)
        return op

    def _ln_op(self, shortcut, h: torch.Tensor, wname: str, bname: str, out=None, want_op: bool = True):
        """y = shortcut + LayerNorm(h) (fp32) and, for the next linear layer, y as a GEMM operand."""
        from . import _lib, gemm

        C = h.shape[-1]
        rows = h.numel() // C
        h = h.contiguous()
        y = out if out is not None else torch.empty(h.shape, dtype=torch.float32, device=h.device)
        hi = torch.empty(h.shape, dtype=torch.bfloat16, device=h.device) if want_op else None
        lo = torch.empty(h.shape, dtype=torch.bfloat16, device=h.device) if (want_op and self._split) else None
        rc = _lib.lib().pmnet_ln_residual_split(
            shortcut.data_ptr() if shortcut is not None else None, h.data_ptr(), int(h.dtype == torch.bfloat16),
            self._g(wname).data_ptr(), self._g(bname).data_ptr(), y.data_ptr(),
            hi.data_ptr() if hi is not None else None, lo.data_ptr() if lo is not None else None, rows, C,
            C_.c_float(1e-5), C_.c_void_p(torch.cuda.current_stream(h.device).cuda_stream),
        )  # fmt: skip
        _lib.check(rc, "pmnet_ln_residual_split")
        return y, (gemm.Operand(hi, lo) if want_op else None)

    def _block_tc(self, x: torch.Tensor, xop, blk: str, res: int, heads: int, shift: int):
        from . import _lib, gemm

        B, L, C = x.shape
        if res <= WINDOW:
            shift = 0
        split = self._split
        bias = self._scale_cache.get(blk + "qkvb")
        if bias is None:
            qb, vb = self._g(blk + "attn.q_bias"), self._g(blk + "attn.v_bias")
            bias = self._scale_cache[blk + "qkvb"] = torch.cat((qb, torch.zeros_like(vb), vb)).float().contiguous()
            self._scale_cache[blk + "scale"] = (
                torch.clamp(self._g(blk + "attn.logit_scale"), max=math.log(1.0 / 0.01)).exp().reshape(-1).contiguous()
            )
        mask = self.sd.get(self.p + blk + "attn_mask") if shift > 0 else None
        stream = C_.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        scale, rel = self._scale_cache[blk + "scale"], self._rel_bias(blk, heads, 64)
        if split:
            qkv, _ = gemm.linear(xop, self._w(blk + "attn.qkv.weight"), bias)  # fp32 [B, L, 3C]
            a_hi = torch.empty((B, L, C), dtype=torch.bfloat16, device=x.device)
            a_lo = torch.empty_like(a_hi)
            rc = _lib.lib().pmnet_window_attention_split(
                qkv.data_ptr(), a_hi.data_ptr(), a_lo.data_ptr(), scale.data_ptr(), rel.data_ptr(),
                mask.data_ptr() if mask is not None else None, B, res, shift, heads, stream,
            )  # fmt: skip
            _lib.check(rc, "pmnet_window_attention_split")
            aop = gemm.Operand(a_hi, a_lo)
        else:
            _, qop = gemm.linear(xop, self._w(blk + "attn.qkv.weight"), bias, want_f32=False, want_operand=True)
            a_hi = torch.empty((B, L, C), dtype=torch.bfloat16, device=x.device)
            rc = _lib.lib().pmnet_window_attention(
                qop.hi.data_ptr(), a_hi.data_ptr(), scale.data_ptr(), rel.data_ptr(),
                mask.data_ptr() if mask is not None else None, B, res, shift, heads, 1, stream,
            )  # fmt: skip
            _lib.check(rc, "pmnet_window_attention")
            aop = gemm.Operand(a_hi)
        proj, pop = gemm.linear(aop, self._w(blk + "attn.proj.weight"), self._g(blk + "attn.proj.bias"),
                                want_f32=split, want_operand=not split)
        x, xop = self._ln_op(x, proj if split else pop.hi, blk + "norm1.weight", blk + "norm1.bias")
        _, hop = gemm.linear(xop, self._w(blk + "mlp.fc1.weight"), self._g(blk + "mlp.fc1.bias"), gemm.ACT_GELU,
                             want_f32=False, want_operand=True)
        m, mop = gemm.linear(hop, self._w(blk + "mlp.fc2.weight"), self._g(blk + "mlp.fc2.bias"),
                             want_f32=split, want_operand=not split)
        return self._ln_op(x, m if split else mop.hi, blk + "norm2.weight", blk + "norm2.bias", out=x)

    def _merge_tc(self, xop, pre: str, res: int, B: int):
        """PatchMerging (swinv2.py:346-363): the 2^3 neighbours are gathered on the bf16 operand(s) (pure data
        movement), then reduction GEMM + LayerNorm."""
        from . import gemm

        def gather(t):
            C = t.shape[-1]
            t = t.view(B, res, res, res, C)
            parts = [t[:, i::2, j::2, k::2, :] for k in (0, 1) for j in (0, 1) for i in (0, 1)]
            return torch.cat(parts, -1).reshape(B, -1, 8 * C).contiguous()

        gop = gemm.Operand(gather(xop.hi), gather(xop.lo) if xop.lo is not None else None)
        y, yop = gemm.linear(gop, self._w(pre + "reduction.weight"), None, want_f32=self._split, want_operand=not self._split)
        return self._ln_op(None, y if self._split else yop.hi, pre + "norm.weight", pre + "norm.bias")

    @torch.no_grad()
    def forward_tokens(self, image: torch.Tensor) -> list[torch.Tensor]:
        """image fp32 [B, 33, 64, 64, 64] -> the four stage outputs token-major, fp32 [B, res^3, dim]."""
        from . import gemm

        assert image.is_cuda and self.precision in ("bf16", "bf16x3")
        B, Cin, D, H, W = image.shape
        patches = image.view(B, Cin, D // PATCH, PATCH, H // PATCH, PATCH, W // PATCH, PATCH)
        patches = patches.permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B, -1, Cin * PATCH**3)
        pop = gemm.Operand.from_float(patches, self._split)
        y, yop = gemm.linear(pop, self._w("patch_embed.proj.weight#2d"), self._g("patch_embed.proj.bias"),
                             want_f32=self._split, want_operand=not self._split)
        x, xop = self._ln_op(None, y if self._split else yop.hi, "patch_embed.norm.weight", "patch_embed.norm.bias")
        outs = []
        res = self.res0
        for li, (depth, heads) in enumerate(zip(DEPTHS, HEADS)):
            for bi in range(depth):
                x, xop = self._block_tc(x, xop, f"layers.{li}.blocks.{bi}.", res, heads, 0 if bi % 2 == 0 else WINDOW // 2)
            outs.append(self._ln_op(None, x, f"norm{li}.weight", f"norm{li}.bias", want_op=False)[0])
            if li < len(DEPTHS) - 1:
                x, xop = self._merge_tc(xop, f"layers.{li}.downsample.", res, B)
                res //= 2
        return outs

    def _merge(self, x: torch.Tensor, pre: str, res: int) -> torch.Tensor:
        B, L, C = x.shape
        x = x.view(B, res, res, res, C)
        parts = [x[:, i::2, j::2, k::2, :] for k in (0, 1) for j in (0, 1) for i in (0, 1)]  # swinv2.py:346-354 order
        x = torch.cat(parts, -1).reshape(B, -1, 8 * C)
        if self.fused and x.is_cuda:
            return self._ln_res(None, self._gemm(x, pre + "reduction.weight", None, keep16=True), pre + "norm.weight", pre + "norm.bias")
        x = self._lin(x, pre + "reduction.weight")
        return F.layer_norm(x, (2 * C,), self._g(pre + "norm.weight"), self._g(pre + "norm.bias"))

    @torch.no_grad()
    def forward(self, image: torch.Tensor) -> list[torch.Tensor]:
        """image fp32 [B, 33, 64, 64, 64] -> [B,96,32^3], [B,192,16^3], [B,384,8^3], [B,768,4^3] (NCDHW fp32)."""
        if self.fused and image.is_cuda and self.precision != "fp32":
            outs, res = [], self.res0
            for o in self.forward_tokens(image):
                outs.append(o.view(o.shape[0], res, res, res, o.shape[-1]).permute(0, 4, 1, 2, 3).contiguous())
                res //= 2
            return outs
        # PatchEmbed (swinv2.py:484-500): the k = 2, stride = 2 convolution is a GEMM over non-overlapping 2^3 patches
        # (K = 33 * 8 = 264). Written as one: cuDNN would run an fp32 convolution on TF32 tensor cores by default.
        B, Cin, D, H, W = image.shape
        w = self._g("patch_embed.proj.weight")
        C = w.shape[0]
        patches = image.view(B, Cin, D // PATCH, PATCH, H // PATCH, PATCH, W // PATCH, PATCH)
        patches = patches.permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B, -1, Cin * PATCH**3)
        x = self._lin(patches, "patch_embed.proj.weight#2d", self._g("patch_embed.proj.bias"))
        if self.fused and x.is_cuda:
            x = self._ln_res(None, x.contiguous(), "patch_embed.norm.weight", "patch_embed.norm.bias")
        else:
            x = F.layer_norm(x, (C,), self._g("patch_embed.norm.weight"), self._g("patch_embed.norm.bias"))
        outs = []
        res, dim = self.res0, EMBED
        for li, (depth, heads) in enumerate(zip(DEPTHS, HEADS)):
            for bi in range(depth):
                x = self._block(x, f"layers.{li}.blocks.{bi}.", res, heads, 0 if bi % 2 == 0 else WINDOW // 2)
            if self.fused and x.is_cuda:
                o = self._ln_res(None, x, f"norm{li}.weight", f"norm{li}.bias")
            else:
                o = F.layer_norm(x, (dim,), self._g(f"norm{li}.weight"), self._g(f"norm{li}.bias"))
            outs.append(o.view(B, res, res, res, dim).permute(0, 4, 1, 2, 3).contiguous())
            if li < len(DEPTHS) - 1:
                x = self._merge(x, f"layers.{li}.downsample.", res)
                res //= 2
                dim *= 2
        return outs
