"""`PharmacoNet.select_hotspots` (vectorised) against a literal restatement of the reference's per-token loop
(src/pmnet/module.py:235-253) on random scores - an integer output, so the comparison is exact. Runs on CPU tensors:
only the selection logic is exercised, no kernel."""

import numpy as np
import torch

from pharmaconet_b200.constants import INTERACTION_LIST
from pharmaconet_b200.module import DEFAULT_SCORE_THRESHOLD, LONG_INTERACTION, PharmacoNet


def _loop(tokens, scores, narrow, wide, dists, thr):
    keep, rel = [], []
    for i in range(tokens.shape[0]):
        x, y, z, typ = tokens[i].tolist()
        a = scores[i].item()
        r = float((dists[INTERACTION_LIST[int(typ)]] < a).mean())
        rel.append(r)
        if r < thr[INTERACTION_LIST[int(typ)]]:
            keep.append(False)
            continue
        cav = wide if typ in LONG_INTERACTION else narrow
        keep.append(bool(cav[0, x, y, z]))
    return np.array(keep), np.array(rel)


def test_vectorised_filter_equals_reference_loop():
    g = torch.Generator().manual_seed(3)
    rng = np.random.default_rng(3)
    n = 2000
    tokens = torch.cat([torch.randint(0, 16, (n, 3), generator=g), torch.randint(0, 10, (n, 1), generator=g)], 1).long()
    scores = torch.rand(n, generator=g)
    dists = {t: np.sort(rng.uniform(0, 1, size=997)) for t in INTERACTION_LIST}
    # scores that sit exactly on distribution values and on the thresholds' quantiles
    for i in range(0, 200):
        scores[i] = float(dists[INTERACTION_LIST[int(tokens[i, 3])]][rng.integers(0, 997)])
    narrow = torch.rand((1, 16, 16, 16), generator=g) < 0.5
    wide = torch.rand((1, 16, 16, 16), generator=g) < 0.7
    net = PharmacoNet.__new__(PharmacoNet)  # selection logic only: no weights, no device
    net._dist_dev = {t: torch.from_numpy(d) for t, d in dists.items()}
    net.score_threshold = DEFAULT_SCORE_THRESHOLD
    keep, rel = net.select_hotspots(tokens, scores, narrow, wide)
    ref_keep, ref_rel = _loop(tokens, scores, narrow, wide, dists, DEFAULT_SCORE_THRESHOLD)
    assert np.array_equal(keep.numpy(), ref_keep)
    assert np.array_equal(rel.numpy(), ref_rel)  # count / N in fp64 on both sides
    assert 0 < keep.sum() < n
