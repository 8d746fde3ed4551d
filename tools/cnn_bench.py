"""Developer probe / CNN measurement (BASELINE configs[3]): batch of synthetic 64^3 pockets through
forward_feature + token prediction + cavity extraction, and the mask head per hotspot group, with algorithmic
FLOPs from SURVEY appendix B (218.8 + 261.0 + 0.07 GFLOP per pocket; 693.2 GFLOP per group of 4 hotspots).
Also times a torch/cuDNN statement of the same conv stack (what the reference's nn.Modules dispatch to) on the GPU."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pharmaconet_b200 import cnn, cnn_weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--chunk", type=int, default=8, help="pockets per forward call")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--groups", type=int, default=8, help="hotspot groups (of 4) for the mask-head timing")
ap.add_argument("--torch-baseline", action="store_true")
ap.add_argument("--reference", action="store_true", help="also time the unmodified reference modules (oracle/_ref) on this GPU")
a = ap.parse_args()
G = os.path.join(ROOT, "tests", "golden")
man = json.load(open(os.path.join(G, "cnn_manifest.json")))
buf = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "cnn_buffers.npz")).items()}
sd = cnn_weights.synth_state_dict(man, buf, 0)
model = cnn.PharmacoNetModel(sd, "cuda:0")
g = torch.Generator().manual_seed(0)
images = torch.rand((a.chunk, 33, 64, 64, 64), generator=g).cuda()
tokens = torch.cat([torch.randint(0, 64, (200, 3), generator=g), torch.randint(0, 10, (200, 1), generator=g)], 1).long().cuda()


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def stage_times():
    out = {}
    out["backbone"] = timed(lambda: model.backbone.forward(images), a.iters)
    feats = model.forward_feature(images)
    out["forward_feature"] = timed(lambda: model.forward_feature(images, nchw=False), a.iters)
    out["cavity"] = timed(lambda: model.forward_cavity_extraction(feats[-1]), a.iters)
    out["token"] = timed(lambda: model.forward_token_prediction(feats[-1], [tokens] * a.chunk), a.iters)
    return out, feats


for prec in ("bf16x3", "bf16"):
    model.precision = prec
    t, feats = stage_times()
    n_chunks = a.batch // a.chunk
    per_pocket = (t["forward_feature"] + t["cavity"] + t["token"]) / a.chunk
    flop = 479.9e9
    print(f"[{prec}] per chunk of {a.chunk}: " + ", ".join(f"{k} {v:.2f} ms" for k, v in t.items()))
    print(f"[{prec}] pocket forward (feature+cavity+token): {per_pocket:.3f} ms/pocket, "
          f"{flop/per_pocket/1e9:.1f} TFLOP/s algorithmic; batch {a.batch}: {per_pocket*a.batch:.1f} ms")
    conv_ms = t["forward_feature"] - t["backbone"]
    print(f"[{prec}] FPN decoder share: {conv_ms/a.chunk:.3f} ms/pocket ({170.0e9/(conv_ms/a.chunk)/1e9:.1f} TFLOP/s), "
          f"cavity {t['cavity']/a.chunk:.3f} ms/pocket ({261.0e9/(t['cavity']/a.chunk)/1e9:.1f} TFLOP/s)")

_, tfeat = model.forward_token_prediction(feats[-1], [tokens] * a.chunk)
one = tuple(f[:1] for f in feats)
for f_src, f_dst in zip(feats, one):
    f_dst._pm_c8 = f_src._pm_c8[:1]


def seg():
    for gi in range(a.groups):
        sl = slice(4 * gi, 4 * gi + 4)
        model.forward_segmentation(one, [tokens[sl]], [tfeat[0][sl]])


ms = timed(seg, a.iters) / a.groups
print(f"mask head, one call per group of 4 (the reference's call pattern): {ms:.3f} ms per group, "
      f"{693.2e9/ms/1e9:.1f} TFLOP/s algorithmic ({ms/4:.3f} ms per hotspot)")
nb = 4 * a.groups
ms = timed(lambda: model.forward_segmentation(one, [tokens[:nb]], [tfeat[0][:nb]], group_size=4), a.iters)
print(f"mask head, {nb} hotspots of a pocket in one call (groups of 4 kept): {ms:.3f} ms, "
      f"{173.3e9*nb/ms/1e9:.1f} TFLOP/s algorithmic ({ms/nb:.3f} ms per hotspot)")

if a.torch_baseline:
    # the same conv stack as torch ops (cuDNN): 64^3 k3 conv + BN + ReLU, fp32 with TF32 (torch default for convs)
    x = torch.randn((a.chunk, 96, 64, 64, 64), device="cuda")
    w = torch.randn((96, 96, 3, 3, 3), device="cuda") * 0.03
    s, b_ = torch.rand(96, device="cuda") + 0.5, torch.randn(96, device="cuda")
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    f32 = timed(lambda: torch.relu(F.batch_norm(F.conv3d(x, w, padding=1), b_, s, s, b_, False)), a.iters)
    xb, wb = x.bfloat16().contiguous(memory_format=torch.channels_last_3d), w.bfloat16().contiguous(memory_format=torch.channels_last_3d)
    bf = timed(lambda: torch.relu(F.batch_norm(F.conv3d(xb, wb, padding=1), b_.bfloat16(), s.bfloat16(), s.bfloat16(), b_.bfloat16(), False)), a.iters)
    flop = 2.0 * a.chunk * 64**3 * 96 * 96 * 27
    print(f"torch conv3d+BN+ReLU 96->96 @64^3 x{a.chunk}: fp32/TF32 NCDHW {f32:.2f} ms ({flop/f32/1e9:.0f} TFLOP/s), "
          f"bf16 channels_last_3d {bf:.2f} ms ({flop/bf/1e9:.0f} TFLOP/s)")

if a.reference:
    # the reference's own nn.Modules through torch / cuDNN on the same GPU (the bar to beat)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness

    ref_harness.import_reference()
    from pmnet.network import build_model

    ref = build_model({}).eval()
    ref.load_state_dict(sd, strict=True)
    ref = ref.cuda()
    with torch.no_grad():
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            rf = ref.forward_feature(images)
            t_feat = timed(lambda: ref.forward_feature(images), a.iters)
            t_cav = timed(lambda: ref.forward_cavity_extraction(rf[-1]), a.iters)
            t_tok = timed(lambda: ref.forward_token_prediction(rf[-1], [tokens] * a.chunk), a.iters)
            per = (t_feat + t_cav + t_tok) / a.chunk
            print(f"[reference modules on this GPU, tf32={tf32}] per chunk of {a.chunk}: forward_feature {t_feat:.2f} ms, cavity "
                  f"{t_cav:.2f} ms, token {t_tok:.2f} ms -> {per:.3f} ms/pocket ({479.9e9/per/1e9:.1f} TFLOP/s algorithmic)")
