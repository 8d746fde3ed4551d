"""`PharmacoNet` with the reference's surface (src/pmnet/module.py:49-322) on top of the B200 CNN path.

Kept: constructor arguments, thresholds (module.py:30-43), `run`, `feature_extraction`, `run_extraction`,
`create_density_maps`, `get_center`, `device` / `to` / `cuda` / `cpu`, and `get_pmnet_dev` (api/__init__.py:12-32).
Different on purpose:
* token filtering (module.py:155-173, 235-253) is vectorised on the device - one searchsorted per interaction type
  instead of one `.item()` round trip per token; the selected indices are the same integers for the same scores;
* box areas, masking, Gaussian smoothing and thresholding (module.py:277-288) are one kernel (`cnn.density_post`);
* protein parsing / voxelisation needs OpenBabel, biopython and molvoxel: `run` and
  `feature_extraction` import the reference's `pmnet.data.parser.ProteinParser` when it is installed and raise a
  clear ImportError otherwise; `run_extraction` / `create_density_maps` / `create_model` take the parsed tuple
  `(image[33,64,64,64], mask[64,64,64], token_pos[N,3], tokens[N,4])` directly.
"""

from __future__ import annotations

import logging
from pathlib import Path
from typing import Any

import numpy as np
import torch

from . import cnn
from .constants import INTERACTION_LIST, INTERACTION_TO_HOTSPOT, INTERACTION_TO_PHARMACOPHORE
from .pharmacophore_model import PharmacophoreModel

DEFAULT_FOCUS_THRESHOLD = 0.5
DEFAULT_BOX_THRESHOLD = 0.5
DEFAULT_SCORE_THRESHOLD = {
    "PiStacking_P": 0.7,
    "PiStacking_T": 0.7,
    "SaltBridge_lneg": 0.7,
    "SaltBridge_pneg": 0.7,
    "PiCation_lring": 0.7,
    "PiCation_pring": 0.7,
    "XBond": 0.85,
    "HBond_ldon": 0.85,
    "HBond_pdon": 0.85,
    "Hydrophobic": 0.85,
}
LONG_INTERACTION = (1, 2, 3, 4, 7, 8)  # data/constant.py:43-50
SEGMENTATION_GROUP = 4  # module.py:261-264: `self.device == "cpu"` is never true, the group size is always 4

HotspotInfo = dict[str, Any]


class PharmacoNet:
    def __init__(
        self,
        device: str | torch.device = "cuda",
        score_threshold: float | dict[str, float] | None = DEFAULT_SCORE_THRESHOLD,
        verbose: bool = True,
        molvoxel_library: str = "numba",
        weight_path: str | Path | None = None,
        checkpoint: dict | None = None,
        precision: str = "bf16x3",
    ):
        """checkpoint: an already loaded `model.tar` dict (`model`, `score_distributions`[, `config`]); otherwise
        `weight_path` is read with torch.load. There is no download here (no network in this deployment)."""
        assert molvoxel_library in ["numpy", "numba"]
        self.molvoxel_library = molvoxel_library
        self._parser = None
        if checkpoint is None:
            if weight_path is None:
                weight_path = Path(__file__).parent / "weights" / "model.tar"
            weight_path = Path(weight_path)
            if not weight_path.exists():
                raise FileNotFoundError(
                    f"{weight_path} not found: pass weight_path= (the reference's model.tar) or checkpoint="
                )
            checkpoint = torch.load(weight_path, map_location="cpu")
        self.model = cnn.PharmacoNetModel(checkpoint["model"], device)
        # "bf16x3" (default): split-precision convolutions, integer outputs (cavity masks, selected hotspots, map
        # supports) follow the fp32 reference; "bf16": single-pass bf16 operands, ~3x faster convolution stack
        self.precision = precision
        self.score_distributions = {
            typ: np.sort(np.asarray(dist["focus"] if isinstance(dist, dict) else dist, dtype=np.float64))
            for typ, dist in checkpoint["score_distributions"].items()
        }
        self._dist_dev = {typ: torch.from_numpy(d).to(self.device) for typ, d in self.score_distributions.items()}
        self.focus_threshold: float = DEFAULT_FOCUS_THRESHOLD
        self.box_threshold: float = DEFAULT_BOX_THRESHOLD
        if isinstance(score_threshold, dict):
            self.score_threshold = score_threshold
        elif isinstance(score_threshold, float):
            self.score_threshold = dict.fromkeys(INTERACTION_LIST, score_threshold)
        else:
            self.score_threshold = DEFAULT_SCORE_THRESHOLD
        self.logger = logging.getLogger("PharmacoNet") if verbose else None

    @property
    def precision(self) -> str:
        return self.model.precision

    @precision.setter
    def precision(self, value: str) -> None:
        if value not in ("bf16", "bf16x3"):
            raise ValueError("precision must be 'bf16' or 'bf16x3'")
        self.model.precision = value

    # ------------------------------------------------------------------ device handling (module.py:311-322)
    @property
    def device(self) -> torch.device:
        return self.model.device

    def to(self, device):
        if torch.device(device) != self.device:
            precision = self.model.precision
            self.model = cnn.PharmacoNetModel(self.model.sd, device)
            self.model.precision = precision
            self._dist_dev = {t: d.to(self.device) for t, d in self._dist_dev.items()}

    def cuda(self):
        self.to("cuda")

    def cpu(self):
        raise RuntimeError("pharmaconet_b200 runs on CUDA devices only (no CPU fallback)")

    def print_log(self, level, log):
        if self.logger is not None:
            getattr(self.logger, level)(log)

    # ------------------------------------------------------------------ parsing (out of scope, delegated)
    @property
    def parser(self):
        if self._parser is None:
            try:
                from pmnet.data.parser import ProteinParser  # the reference package, if installed
            except Exception as e:  # noqa: BLE001
                raise ImportError(
                    "protein parsing / voxelisation needs the reference's pmnet.data (OpenBabel, biopython, molvoxel); "
                    "pass the parsed protein_data tuple to run_extraction / create_density_maps / create_model instead"
                ) from e
            self._parser = ProteinParser(molvoxel_library=self.molvoxel_library)
        return self._parser

    def get_center(self, ref_ligand_path=None, center=None) -> tuple[float, float, float]:
        if center is not None:
            assert len(center) == 3
            x, y, z = center
            return float(x), float(y), float(z)
        from openbabel import pybel  # module.py:205-213

        ext = Path(ref_ligand_path).suffix
        assert ext in [".sdf", ".pdb", ".mol2"]
        mol = next(pybel.readfile(ext[1:], str(ref_ligand_path)))
        x, y, z = np.mean([a.coords for a in mol.atoms], axis=0, dtype=np.float32).tolist()
        return float(x), float(y), float(z)

    def run(self, protein_pdb_path, ref_ligand_path=None, center=None) -> PharmacophoreModel:
        assert (ref_ligand_path is not None) or (center is not None)
        center = self.get_center(ref_ligand_path, center)
        protein_data = self.parser.parse(protein_pdb_path, center=center)
        with open(protein_pdb_path) as f:
            pdbblock = "\n".join(f.readlines())
        return self.create_model(protein_data, pdbblock, center)

    def feature_extraction(self, protein_pdb_path, ref_ligand_path=None, center=None):
        return self.run_extraction(self.parser.parse(protein_pdb_path, ref_ligand_path, center))

    def create_model(self, protein_data, pdbblock: str = "", center=(0.0, 0.0, 0.0)) -> PharmacophoreModel:
        return PharmacophoreModel.create(pdbblock, tuple(float(c) for c in center), self.create_density_maps(protein_data))

    # ------------------------------------------------------------------ shared front part
    @torch.no_grad()
    def _features_and_hotspots(self, protein_data, nchw: bool = True):
        return self._features_and_hotspots_batch([protein_data], nchw)[0]

    @torch.no_grad()
    def _features_and_hotspots_batch(self, protein_data_list, nchw: bool = True):
        """module.py:141-170 / 216-258 for several pockets at once: one batched forward_feature, token head and
        cavity head, then the per-pocket token filter."""
        dev = self.device
        images = torch.stack([pd[0].to(dev, torch.float32) for pd in protein_data_list])
        masks = [pd[1].to(dev, torch.bool) for pd in protein_data_list]
        token_pos = [pd[2].to(dev, torch.float32) for pd in protein_data_list]
        tokens = [pd[3].to(dev, torch.long) for pd in protein_data_list]
        feats = self.model.forward_feature(images, nchw=nchw)
        scores, tfeat = self.model.forward_token_prediction(feats[-1], tokens)
        narrow_all, wide_all = self.model.forward_cavity_extraction(feats[-1])
        out = []
        for b in range(len(protein_data_list)):
            abs_scores = scores[b].sigmoid()
            narrow = narrow_all[b].sigmoid() > self.focus_threshold  # [1, D, H, W]
            wide = wide_all[b].sigmoid() > self.focus_threshold
            keep, rel = self.select_hotspots(tokens[b], abs_scores, narrow, wide)
            idx = torch.nonzero(keep).reshape(-1)
            fb = cnn.Features(f[b : b + 1] for f in feats)
            for src, dst in zip(feats, fb):
                twin = getattr(src, "_pm_c8", None)
                if twin is not None:
                    dst._pm_c8 = twin[b : b + 1]
            out.append(
                dict(
                    feats=fb, mask=masks[b], narrow=narrow, wide=wide, hotspots=tokens[b][idx], positions=token_pos[b][idx],
                    features=tfeat[b][idx], rel_scores=rel[idx].cpu().tolist(),
                )
            )  # fmt: skip
        return out

    def select_hotspots(self, tokens, abs_scores, cavity_narrow, cavity_wide):
        """module.py:235-253 for all tokens at once. relative score = fraction of the type's training-score
        distribution below the token's score; kept iff it reaches the type's threshold and the token voxel lies in
        the cavity (wide cavity for long-range interaction types). Returns (keep mask, relative scores fp64)."""
        n = tokens.shape[0]
        rel = torch.zeros(n, dtype=torch.float64, device=tokens.device)
        thr = torch.zeros(n, dtype=torch.float64, device=tokens.device)
        typ = tokens[:, 3]
        for t, name in enumerate(INTERACTION_LIST):
            sel = typ == t
            if not bool(sel.any()):
                continue
            dist = self._dist_dev[name]
            cnt = torch.searchsorted(dist, abs_scores[sel].double(), right=False)  # #{distribution < score}
            rel[sel] = cnt.double() / dist.numel()
            thr[sel] = float(self.score_threshold[name])
        long = torch.zeros(n, dtype=torch.bool, device=tokens.device)
        for t in LONG_INTERACTION:
            long |= typ == t
        x, y, z = tokens[:, 0], tokens[:, 1], tokens[:, 2]
        in_cavity = torch.where(long, cavity_wide[0, x, y, z], cavity_narrow[0, x, y, z])
        return (rel >= thr) & in_cavity, rel

    # ------------------------------------------------------------------ module.py:137-188
    @torch.no_grad()
    def run_extraction(self, protein_data):
        r = self._features_and_hotspots(protein_data)
        infos = []
        for hotspot, score, pos, feat in zip(r["hotspots"], r["rel_scores"], r["positions"], r["features"], strict=True):
            name = INTERACTION_LIST[int(hotspot[3])]
            infos.append(
                dict(
                    nci_type=name,
                    hotspot_type=INTERACTION_TO_HOTSPOT[name],
                    hotspot_feature=feat,
                    hotspot_position=tuple(pos.tolist()),
                    hotspot_score=float(score),
                    point_type=INTERACTION_TO_PHARMACOPHORE[name],
                )
            )
        return tuple(r["feats"]), infos

    # ------------------------------------------------------------------ module.py:215-309
    @torch.no_grad()
    def create_density_maps(self, protein_data) -> list[HotspotInfo]:
        return self.create_density_maps_batch([protein_data])[0]

    @torch.no_grad()
    def create_models(self, protein_data_list, centers=None, pdbblocks=None, chunk: int = 8) -> list[PharmacophoreModel]:
        """Several pockets -> several PharmacophoreModels (BASELINE configs[4] shape): the CNN runs on chunks of
        `chunk` pockets, the graph construction per pocket on the host."""
        n = len(protein_data_list)
        centers = centers or [(0.0, 0.0, 0.0)] * n
        pdbblocks = pdbblocks or [""] * n
        models = []
        for lo in range(0, n, chunk):
            # only the non-zero voxels of the maps come back from the device (KBs instead of 1 MB per hotspot)
            infos = [
                self._density_maps_of(r, sparse=True)
                for r in self._features_and_hotspots_batch(protein_data_list[lo : lo + chunk], nchw=False)
            ]
            for k, info in enumerate(infos):
                models.append(PharmacophoreModel.create(pdbblocks[lo + k], tuple(float(c) for c in centers[lo + k]), info))
        return models

    @torch.no_grad()
    def create_density_maps_batch(self, protein_data_list) -> list[list[HotspotInfo]]:
        self.print_log("debug", f"Protein-based Pharmacophore Modeling... (device: {self.device})")
        # the maps stay in the kernels' layout
        return [self._density_maps_of(r) for r in self._features_and_hotspots_batch(protein_data_list, nchw=False)]

    def _density_maps_of(self, r, sparse: bool = False) -> list[HotspotInfo]:
        """sparse: give `point_map_sparse` = (voxel coordinates int32 [n,3] in C order, values fp32 [n]) instead of
        the dense `point_map` - the form `PharmacophoreModel.create` consumes without a 64^3 scan per hotspot."""
        hotspots, feats = r["hotspots"], r["feats"]
        logits = []
        if hotspots.shape[0] > 0:
            # all hotspots in one call; the reference's groups of 4 (module.py:261-272) are reproduced inside
            logits.append(
                self.model.forward_segmentation(feats, [hotspots], [r["features"]], group_size=SEGMENTATION_GROUP)[0][0]
            )
        infos: list[HotspotInfo] = []
        if logits:
            maps = cnn.density_post(torch.cat(logits, 0), hotspots, r["mask"], r["narrow"][0], self.box_threshold)
            alive = (maps.reshape(maps.shape[0], -1) >= 1e-6).any(dim=1).cpu().tolist()  # module.py:292-293
            if sparse:
                nz = torch.nonzero(maps > 0)  # [n, 4] = (hotspot, x, y, z), lexicographic = np.where order per map
                vals = maps[nz[:, 0], nz[:, 1], nz[:, 2], nz[:, 3]].cpu().numpy()
                counts = torch.bincount(nz[:, 0], minlength=maps.shape[0]).cpu().numpy()
                coords = nz[:, 1:].to(torch.int32).cpu().numpy()
                ends = np.cumsum(counts)
            else:
                maps = maps.cpu().numpy()
            positions = r["positions"].cpu().numpy()
            types = hotspots[:, 3].cpu().tolist()
            for k, score in enumerate(r["rel_scores"]):
                if not alive[k]:
                    continue
                name = INTERACTION_LIST[int(types[k])]
                info = dict(
                    nci_type=name,
                    hotspot_type=INTERACTION_TO_HOTSPOT[name],
                    hotspot_position=positions[k],
                    hotspot_score=score,
                    point_type=INTERACTION_TO_PHARMACOPHORE[name],
                )
                if sparse:
                    lo = int(ends[k] - counts[k])
                    info["point_map_sparse"] = (coords[lo : int(ends[k])], vals[lo : int(ends[k])])
                else:
                    info["point_map"] = maps[k]
                infos.append(info)
        self.print_log("debug", f"Protein-based Pharmacophore Modeling finish (Total {len(infos)} protein hotspots are detected)")
        return infos


def get_pmnet_dev(
    device: str | torch.device = "cuda",
    score_threshold: float = 0.5,
    molvoxel_library: str = "numpy",
    compile: bool = False,  # noqa: A002 - the reference's argument name
    **kwargs,
) -> PharmacoNet:
    """api/__init__.py:12-32. `compile` is accepted for signature compatibility and ignored: the hot ops are
    hand-written kernels, there is no tracing compiler on this path."""
    return PharmacoNet(device, score_threshold, False, molvoxel_library, **kwargs)
