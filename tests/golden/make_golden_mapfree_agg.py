#!/usr/bin/env python
"""Golden vectors for the map-free correlation-volume aggregator, produced by the UNMODIFIED
reference module (mapfree_6dreg/lib/models/regression/aggregator.py:6-116, loaded from its own file; its only import,
PreActBlock, is not used by the shipped recipe) and checked against the oracle restatement.
Runs only in the build container:   python tests/golden/make_golden_mapfree_agg.py"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import far_oracle as O  # noqa: E402


def load_reference_aggregator():
    root = "/root/reference/mapfree_6dreg"
    for name in ("lib", "lib.models", "lib.models.regression", "lib.models.regression.encoder"):
        sys.modules.setdefault(name, types.ModuleType(name))
    pre = types.ModuleType("lib.models.regression.encoder.preact")
    pre.PreActBlock = object
    sys.modules["lib.models.regression.encoder.preact"] = pre
    spec = importlib.util.spec_from_file_location("ref_aggregator", root + "/lib/models/regression/aggregator.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    # config/regression/mapfree/rot6d_trans_with_loftr.yaml:8-11 over config/default.py:17-27
    cfg = types.SimpleNamespace(POSITION_ENCODER=True, POSITION_ENCODER_IM1=None, MAX_SCORE_CHANNEL=True,
                                CV_OUTLAYERS=0, CV_HALF_CHANNELS=False, UPSAMPLE_POS_ENC=0, DUSTBIN=False,
                                NORMALISE_DOT=False)
    return m.CorrelationVolumeWarping(cfg, 128)


def main():
    agg = load_reference_aggregator()
    g = torch.Generator().manual_seed(2024)
    B, D, H, W = 2, 128, 23, 17                      # small stand-in for the 90 x 68 grid (6120 tokens) of the recipe
    v0 = torch.randn(B, D, H, W, generator=g) * 0.35
    v1 = torch.randn(B, D, H, W, generator=g) * 0.35
    with torch.no_grad():
        ref = agg(v0, v1)
        ora = O.mapfree_correlation_aggregator(v0, v1)
    print("aggregator max|diff| oracle vs reference:", (ref - ora).abs().max().item(), tuple(ref.shape))
    assert ref.shape == (B, 2 * D + 3, H, W) and (ref - ora).abs().max() < 1e-6
    out = os.path.join(ROOT, "tests", "golden", "mapfree_agg.npz")
    np.savez_compressed(out, seed=np.int64(2024), shape=np.array([B, D, H, W]), warped=ref[:, D:2 * D, ::3, ::2].numpy(),
                        pos=ref[:, 2 * D:2 * D + 2].numpy(), max_score=ref[:, 2 * D + 2].numpy())
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
