#!/usr/bin/env python
"""Golden vectors for the prior-guided RANSAC scoring step (SURVEY.md 8f rank 2 / kernel K9), produced by the UNMODIFIED
reference `RANSAC.verify` / `RANSAC.get_prior_estimate` (mp3d_loftr/third_party/prior_ransac/ransac.py:203-231,256-292)
imported through the shims of oracle/ref_import.py, and checked against the oracle restatement on the way.
Runs only in the build container (needs /root/reference):   python tests/golden/make_golden_ransac.py
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import far_oracle as O  # noqa: E402
from oracle import ref_import as RI  # noqa: E402
from far_b200 import synth  # noqa: E402


def load_reference_ransac():
    RI.install_shims()
    kg = sys.modules["kornia.geometry"]
    for n in ("find_homography_dlt_iterated", "find_homography_lines_dlt", "find_homography_lines_dlt_iterated"):
        if not hasattr(kg, n):
            setattr(kg, n, (lambda *a, **k: None))
    kg.symmetrical_epipolar_distance = RI._symmetrical_epipolar_distance
    # kornia 0.7.1 essential_from_Rt(R1, t1, R2, t2): R = R2 R1^T, t = t2 - R t1, E = [t]_x R  (restated)
    epi = sys.modules["kornia.geometry.epipolar"]
    epi.essential_from_Rt = lambda R1, t1, R2, t2: O.essential_from_prior_rt(
        torch.cat([R2 @ R1.transpose(-2, -1), t2 - (R2 @ R1.transpose(-2, -1)) @ t1], -1))
    kg.epipolar = epi
    kh = sys.modules["kornia.geometry.homography"]
    for n in ("oneway_transfer_error", "sample_is_valid_for_homography", "line_segment_transfer_error_one_way"):
        if not hasattr(kh, n):
            setattr(kh, n, (lambda *a, **k: None))
    pr = "/root/reference/mp3d_loftr/third_party/prior_ransac"
    sys.path.insert(0, pr)
    for m in ("ransac", "utils", "essential", "cv_geometry"):
        sys.modules.pop(m, None)
    return importlib.import_module("ransac")


def main():
    ransac = load_reference_ransac()
    g = torch.Generator().manual_seed(77)
    N, H, NPCL = 600, 256, 300
    p1, p2, w, Rg, tg = synth.two_view_geometry(1, N, seed=41, noise=2e-4, outlier_frac=0.3)
    kp1, kp2 = p1[0].float(), p2[0].float()
    # hypotheses: 8-point models of random minimal samples (the 'fundamental' branch of the reference, ransac.py:140-145)
    idx = torch.stack([torch.randperm(N, generator=g)[:8] for _ in range(H)])
    models = O.run_8point(kp1[idx], kp2[idx], torch.ones(H, 8))
    # prior: the true pose perturbed by ~3 degrees, unit translation (setup_prior normalises it, ransac.py:180)
    ang = 0.05
    Rp = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]], dtype=torch.float32) @ Rg[0].float()
    tp = tg[0].float() + 0.05 * torch.randn(3, generator=g)
    prior_rt = torch.cat([Rp, (tp / tp.norm())[:, None]], 1)
    pcl = (torch.rand(NPCL, 3, generator=g) * 6 - 3).float()
    inl_th = 3e-7 * 1e3   # the recipe's 3e-7 is tuned for real matches; the synthetic noise level needs a looser one
    prior_params = {'rotation_pcl_error': True, 'rotation_error': False, 'K1': torch.eye(3), 'K2': torch.eye(3),
                    'RT': prior_rt.clone(), 'pcl': pcl.clone(), 'lambda': 0.3, 'biased_sampling': 'biased'}
    model = ransac.RANSAC(model_type='essential_cv2', max_iter=1, inl_th=inl_th, prior_params=prior_params, max_lo_iters=0,
                          batch_size=H, use_noexp_prior_scoring=True, use_linear_bias_sampling=True, bias_sigma_sq=0.1)
    with torch.no_grad():
        good = model.remove_bad_models(models)
        assert good.shape[0] == int(O.ransac_good_models(models).sum())
        err = model.get_prior_estimate(models)
        prior_ref = -err ** 2 / model.prior_lambda
        m_best, inl, score_best, inl_t, inl_u = model.verify(kp1, kp2, models, inl_th, prior_ref)
        # bias weights of the sampling stage (ransac.py:358-367)
        Fp = ransac.fundamental_from_RT(prior_params['RT'], prior_params['K1'], prior_params['K2'])
        bias_ref = torch.exp(-RI._symmetrical_epipolar_distance(kp1[None], kp2[None], Fp[None]) / 0.1).squeeze()
    # oracle restatement vs the reference
    prior_o = O.ransac_prior_estimate(models, prior_rt, pcl, 0.3)
    best_o, score_o, masks_o = O.ransac_verify(kp1, kp2, models, inl_th, prior_o)
    bias_o = O.ransac_bias_weight(kp1, kp2, prior_rt, 0.1)
    print("prior estimate max|diff|", (prior_o - prior_ref).abs().max().item())
    print("bias weight   max|diff|", (bias_o - bias_ref).abs().max().item())
    assert (prior_o - prior_ref).abs().max() < 1e-5
    assert (bias_o - bias_ref).abs().max() < 1e-5
    assert torch.equal(models[best_o], m_best), "best model differs"
    assert torch.equal(masks_o[0], inl) and torch.equal(masks_o[1], inl_t) and torch.equal(masks_o[2], inl_u)
    assert abs(float(score_o[best_o]) - score_best) < 1e-3
    print(f"best model {best_o}: score {score_best:.3f}, inliers {int(inl.sum())}/{int(inl_t.sum())}/{int(inl_u.sum())} of {N}")
    out = os.path.join(ROOT, "tests", "golden", "ransac.npz")
    np.savez_compressed(out, kp1=kp1.numpy(), kp2=kp2.numpy(), models=models.numpy(), prior_rt=prior_rt.numpy(),
                        pcl=pcl.numpy(), inl_th=np.float32(inl_th), prior_ref=prior_ref.numpy(), bias_ref=bias_ref.numpy(),
                        best=np.int64(best_o), score_best=np.float32(score_best), inl=inl.numpy(), inl_t=inl_t.numpy(),
                        inl_u=inl_u.numpy(), good=O.ransac_good_models(models).numpy())
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
