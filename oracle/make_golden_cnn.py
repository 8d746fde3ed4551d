"""TEST INFRASTRUCTURE - golden vectors of the CNN forward from the REAL reference modules (build container only).

    python oracle/make_golden_cnn.py        # writes tests/golden/cnn_manifest.json, cnn_buffers.npz, cnn_golden.npz

The reference network (src/pmnet/network/builder.py:12-54) is built unmodified, loaded (strict) with the seeded
synthetic state dict of pharmaconet_b200.cnn_weights (no trained weights exist offline) and run on CPU in fp32 on a
seeded 33 x 64^3 grid and 200 seeded tokens. Stored: strided samples of every stage, the full cavity logits (fp16) and
their sign bits, token scores / features, the segmentation logits of one group of 4 hotspots and the reference's own
post-processing of them (module.py:277-288 re-executed with the reference's GaussianSmoothing and get_box_area).
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402

from pharmaconet_b200 import cnn_weights  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SEED = 0
NEAR = 5e-3  # |logit| band recorded next to the mask bits


def main():
    ref_harness.import_reference()
    from pmnet.data import token_inference
    from pmnet.network import build_model
    from pmnet.utils.smoothing import GaussianSmoothing

    torch.manual_seed(0)
    model = build_model({}).eval()
    ref_sd = model.state_dict()
    manifest = {k: list(v.shape) for k, v in ref_sd.items()}
    buffers = {}
    for k, v in ref_sd.items():
        leaf = k.rsplit(".", 1)[-1]
        if leaf in ("relative_coords_table", "relative_position_index", "attn_mask"):
            name = cnn_weights.buffer_name(k, v.shape)
            if name in buffers:
                assert torch.equal(buffers[name], v), f"buffer {k} differs from its shape family"
            buffers[name] = v.clone()
    with open(os.path.join(GOLDEN, "cnn_manifest.json"), "w") as f:
        json.dump(manifest, f)
    np.savez_compressed(os.path.join(GOLDEN, "cnn_buffers.npz"), **{k: v.numpy() for k, v in buffers.items()})
    sd = cnn_weights.synth_state_dict(manifest, buffers, SEED)
    model.load_state_dict(sd, strict=True)

    g = torch.Generator().manual_seed(SEED)
    image = torch.rand((1, 33, 64, 64, 64), generator=g)
    tokens = torch.cat(
        [torch.randint(0, 64, (200, 3), generator=g), torch.randint(0, 10, (200, 1), generator=g)], dim=1
    ).long()
    out = {"tokens": tokens.numpy()}
    with torch.no_grad():
        back = model.embedding.backbone(image)
        for i, t in enumerate(back):
            st = max(1, t.shape[-1] // 4)
            out[f"backbone{i}"] = t[0, :, ::st, ::st, ::st].numpy()  # 4^3 voxels per channel
        feats = model.forward_feature(image)
        for i, t in enumerate(feats):
            s = max(1, t.shape[-1] // 8)
            out[f"feat{i}"] = t[0, :, ::s, ::s, ::s].numpy().astype(np.float16)
            out[f"feat{i}_absmean"] = np.float64(t.abs().mean().item())
        narrow, wide = model.forward_cavity_extraction(feats[-1])
        for name, t in (("narrow", narrow), ("wide", wide)):
            out[f"cavity_{name}_f16_s4"] = t[0, 0, ::4, ::4, ::4].numpy().astype(np.float16)
            out[f"cavity_{name}_bits"] = np.packbits((t[0, 0] > 0).numpy())
            # voxels whose fp32 logit lies within 5e-3 of the threshold: where a result that differs from this CPU run by
            # fp32-level rounding (summation order) may legitimately land on the other side
            near = torch.nonzero(t[0, 0].reshape(-1).abs() < NEAR).reshape(-1)
            out[f"cavity_{name}_near_idx"] = near.numpy().astype(np.int32)
            out[f"cavity_{name}_near_val"] = t[0, 0].reshape(-1)[near].numpy()
        scores, tfeat = model.forward_token_prediction(feats[-1], [tokens])
        out["token_scores"] = scores[0].numpy()
        out["token_features"] = tfeat[0].numpy()
        hot = tokens[:4]
        seg = model.forward_segmentation(feats, [hot], [tfeat[0][:4]])[0][0]  # [4, 64, 64, 64] logits
        out["seg_f16_s4"] = seg[:, ::4, ::4, ::4].numpy().astype(np.float16)
        out["seg_bits"] = np.packbits((seg > 0).numpy())
        near = torch.nonzero(seg.reshape(-1).abs() < NEAR).reshape(-1)
        out["seg_near_idx"] = near.numpy().astype(np.int32)
        out["seg_near_val"] = seg.reshape(-1)[near].numpy()
        out["seg_rms"] = np.float64(seg.pow(2).mean().sqrt().item())
        # reference post-processing (module.py:277-288) on seeded logits / masks (regenerated from the seeds in the test)
        gm = torch.Generator().manual_seed(SEED + 1)
        post_logits = torch.randn((4, 64, 64, 64), generator=gm) * 3.0 + 1.0
        protein_mask = torch.rand((64, 64, 64), generator=gm) < 0.8
        cavity_narrow = (torch.rand((1, 64, 64, 64), generator=gm) < 0.8)
        dm = post_logits.sigmoid()
        box_area = torch.from_numpy(token_inference.get_box_area(hot.numpy())).bool()
        unavailable = ~(box_area & protein_mask & cavity_narrow)
        smoothing = GaussianSmoothing(kernel_size=5, sigma=0.5)
        dm.masked_fill_(unavailable, 0.0)
        dm = smoothing(dm)
        dm.masked_fill_(unavailable, 0.0)
        dm[dm < 0.5] = 0.0
        nz = torch.nonzero(dm.reshape(-1)).reshape(-1)
        out["post_nonzero_index"] = nz.numpy()
        out["post_nonzero_value"] = dm.reshape(-1)[nz].numpy()
        print("density maps: nonzero voxels", int(nz.numel()), "available voxels", int((~unavailable).sum()))
    np.savez_compressed(os.path.join(GOLDEN, "cnn_golden.npz"), **out)
    for k, v in out.items():
        print(k, getattr(v, "shape", None), getattr(v, "dtype", None))
    print("cavity narrow positive fraction", float((narrow > 0).float().mean()), "wide", float((wide > 0).float().mean()))
    print("token score range", float(scores[0].min()), float(scores[0].max()))


def main_pipeline():
    """The reference's PharmacoNet.create_density_maps (module.py:215-309) + PharmacophoreModel.create, unmodified,
    on a synthetic checkpoint and seeded protein data; tokens are drawn inside the reference's own cavity so that the
    filter keeps a useful number of them."""
    import pickle
    import tempfile

    ref_harness.import_reference()
    import pmnet.module as ref_module
    from pmnet.network import build_model
    from pmnet.pharmacophore_model import PharmacophoreModel

    manifest = json.load(open(os.path.join(GOLDEN, "cnn_manifest.json")))
    buffers = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, "cnn_buffers.npz")).items()}
    ckpt = cnn_weights.synth_checkpoint(manifest, buffers, SEED)
    ref_module.OmegaConf.create = lambda cfg: type("Cfg", (), {"MODEL": {}})()  # the stubbed OmegaConf
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "model.tar")
        torch.save(ckpt, path)
        net = ref_module.PharmacoNet("cpu", verbose=False, molvoxel_library="numpy", weight_path=path)
    g = torch.Generator().manual_seed(SEED)
    image = torch.rand((33, 64, 64, 64), generator=g)
    gm = torch.Generator().manual_seed(SEED + 2)
    mask = torch.rand((64, 64, 64), generator=gm) < 0.8
    with torch.no_grad():
        feats = net.model.forward_feature(image.unsqueeze(0))
        narrow, wide = net.model.forward_cavity_extraction(feats[-1])
    # tokens: 120 inside the narrow cavity (margin > 0.3 in logit so that bf16 noise does not move them out),
    # 40 inside the wide cavity, 40 anywhere
    def pick(logit, n, margin):
        idx = torch.nonzero(logit[0, 0] > margin)
        sel = idx[torch.randperm(idx.shape[0], generator=gm)[:n]]
        return sel
    short_types = torch.tensor([0, 5, 6, 9])
    long_types = torch.tensor([1, 2, 3, 4, 7, 8])
    t_short = pick(narrow, 120, 0.3)
    t_long = pick(wide, 40, 0.3)
    t_any = torch.randint(0, 64, (40, 3), generator=gm)
    tokens = torch.cat(
        [
            torch.cat([t_short, short_types[torch.randint(0, 4, (t_short.shape[0], 1), generator=gm)]], 1),
            torch.cat([t_long, long_types[torch.randint(0, 6, (t_long.shape[0], 1), generator=gm)]], 1),
            torch.cat([t_any, torch.randint(0, 10, (40, 1), generator=gm)], 1),
        ]
    ).long()
    token_pos = (tokens[:, :3].float() - 31.5) * 0.5
    infos = net.create_density_maps((image, mask, token_pos, tokens))
    model = PharmacophoreModel.create("", (0.0, 0.0, 0.0), infos)
    # which tokens were selected (the reference does not return indices: recover them from the positions)
    sel = []
    for info in infos:
        d = (token_pos - torch.as_tensor(info["hotspot_position"]).float()).abs().sum(1)
        cand = torch.nonzero(d < 1e-6).reshape(-1).tolist()
        sel.append([c for c in cand if ref_module.C.INTERACTION_LIST[int(tokens[c, 3])] == info["nci_type"]][0])
    out = dict(
        tokens=tokens.numpy(),
        selected_with_nonempty_map=np.asarray(sel, dtype=np.int64),
        rel_scores=np.asarray([i["hotspot_score"] for i in infos], dtype=np.float64),
        map_nonzero=np.asarray([int((i["point_map"] > 0).sum()) for i in infos], dtype=np.int64),
        map_sum=np.asarray([float(i["point_map"].sum()) for i in infos], dtype=np.float64),
        model_nodes=np.asarray(len(model.nodes)),
        model_clusters=np.asarray(len(model.node_clusters)),
    )
    np.savez_compressed(os.path.join(GOLDEN, "cnn_pipeline_golden.npz"), **out)
    print("pipeline: tokens", tokens.shape[0], "hotspots with maps", len(infos), "model nodes", len(model.nodes),
          "clusters", len(model.node_clusters))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pipeline":
        main_pipeline()
    else:
        main()
        main_pipeline()
