"""Prior-guided RANSAC round on the GPU: the deterministic scoring step against the fixture the
UNMODIFIED reference produced (tests/golden/make_golden_ransac.py: RANSAC.verify / get_prior_estimate /
remove_bad_models of mp3d_loftr/third_party/prior_ransac/ransac.py), candidate selection, and the whole stochastic
round through size-independent properties (it must recover the true pose of synthetic two-view geometry with 30 %
outliers; ragged pairs, an unsolvable pair)."""
import os

import numpy as np
import pytest
import torch

from oracle import far_oracle as O
from far_b200 import ops, synth
from far_b200.ransac import prior_ransac_round, bias_weights, normalise_prior

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(t):
    return t.to(DEV)


def _offsets(counts):
    off = torch.zeros(len(counts) + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(torch.tensor(counts), 0)
    return off


def test_prior_ransac_score_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ransac.npz"))
    kp1, kp2 = torch.from_numpy(g["kp1"]), torch.from_numpy(g["kp2"])
    models, prior_rt, pcl = torch.from_numpy(g["models"]), torch.from_numpy(g["prior_rt"]), torch.from_numpy(g["pcl"])
    inl_th = float(g["inl_th"])
    N, H = kp1.shape[0], models.shape[0]
    K = torch.eye(3)[None]                                  # normalised coordinates in, identity intrinsics
    off = _offsets([N])
    args = (cu(kp1), cu(kp2), cu(off), cu(K), cu(K), cu(models[None]))
    s_with, best, best_E, c3, mask = ops.prior_ransac_score(*args, cu(prior_rt[None]), cu(pcl), 0.3, inl_th)
    s_wo, best0, _, _, _ = ops.prior_ransac_score(*args, None, None, 0.3, inl_th)
    good = torch.from_numpy(g["good"])
    assert torch.equal(torch.isfinite(s_with[0]).cpu(), good), "remove_bad_models mask"
    # inlier counts: fp32 Sampson errors evaluated in a different order may flip correspondences sitting on the
    # threshold, nothing else
    err = O.sampson_epipolar_distance(kp1[None].expand(H, -1, 2), kp2[None].expand(H, -1, 2), models)
    cnt_ref = (err <= inl_th).sum(1).float()
    border = ((err - inl_th).abs() <= 1e-3 * inl_th).sum(1).float()
    assert ((s_wo[0].cpu() - cnt_ref).abs() <= border)[good].all()
    # without a prior the winner, masks and counters are those of the reference's verify() with a zero prior score
    # (the oracle's verify is pinned bit-for-bit to the reference in tests/test_oracle_golden.py)
    bo, so, mo = O.ransac_verify(kp1, kp2, models[good], inl_th, torch.zeros(int(good.sum())))
    bo = int(torch.nonzero(good)[bo])
    _, best0, bestE0, c30, mask0 = ops.prior_ransac_score(*args, None, None, 0.3, inl_th)
    assert int(best0[0]) == bo and torch.equal(bestE0[0].cpu(), models[bo])
    assert torch.equal(mask0.cpu().bool(), mo[0]) and c30[0].tolist() == [int(m.sum()) for m in mo]
    # prior term = scores with prior - scores without; the kernel scores both signs of T (see far_oracle)
    prior = (s_with - s_wo)[0].cpu()
    prior_o = O.ransac_prior_estimate(models, prior_rt, pcl, 0.3, both_signs=True)
    prior_ref = torch.from_numpy(g["prior_ref"])           # the reference's value: +T of LAPACK's SVD only
    assert ((prior[good] - prior_o[good]).abs() <= 2e-3 + 2e-4 * prior_o[good].abs()).all(), (prior[good] - prior_o[good]).abs().max()
    assert (prior[good] >= prior_ref[good] - 2e-3).all() and int(((prior - prior_ref).abs()[good] < 1e-3).sum()) > int(good.sum()) // 2
    # with the prior: winner / masks / counters of verify() fed with that prior score
    bw_, sw_, mw_ = O.ransac_verify(kp1, kp2, models[good], inl_th, prior_o[good])
    bw_ = int(torch.nonzero(good)[bw_])
    assert int(best[0]) == bw_ and torch.equal(best_E[0].cpu(), models[bw_])
    assert torch.equal(mask.cpu().bool(), mw_[0]) and c3[0].tolist() == [int(m.sum()) for m in mw_]
    # bias weights of the sampling stage (torch glue in far_b200/ransac.py) against the reference
    bw = bias_weights(cu(kp1), cu(kp2), torch.zeros(N, dtype=torch.int64, device=DEV), cu(normalise_prior(prior_rt[None])), 0.1)
    assert (bw.cpu() - torch.from_numpy(g["bias_ref"])).abs().max() < 1e-5


def _essential(R, t):
    tx = torch.zeros(R.shape[0], 3, 3)
    tx[:, 0, 1], tx[:, 0, 2], tx[:, 1, 0], tx[:, 1, 2], tx[:, 2, 0], tx[:, 2, 1] = -t[:, 2], t[:, 1], t[:, 2], -t[:, 0], -t[:, 1], t[:, 0]
    return tx @ R


def test_pose_from_essential_recovers_pose_and_honours_mask():
    P, N = 4, 200
    p1, p2, w, R, t = synth.two_view_geometry(P, N, seed=12, noise=1e-5, outlier_frac=0.0)
    K = torch.eye(3)[None].repeat(P, 1, 1)
    E = _essential(R, t)
    off = _offsets([N] * P)
    mask = torch.ones(P * N, dtype=torch.uint8)
    mask[::3] = 0
    E[3] = 0.0                                              # "no valid model" -> identity pose
    Rt, npos = ops.pose_from_essential(cu(p1.reshape(-1, 2)), cu(p2.reshape(-1, 2)), cu(mask), cu(off), cu(K), cu(K), cu(E))
    Rt, npos = Rt.cpu(), npos.cpu()
    for b in range(3):
        assert (Rt[b, :, :3] - R[b]).abs().max() < 1e-4
        assert (Rt[b, :, 3] - t[b]).abs().max() < 1e-4      # unit translation, sign fixed by cheirality
        assert int(npos[b]) == int(mask[b * N:(b + 1) * N].sum())
    assert torch.equal(Rt[3], torch.eye(3, 4))


def test_prior_ransac_round_recovers_true_pose_ragged_batch():
    sizes = [900, 5, 400, 1300]                             # pair 1 cannot be solved (< 8 matches) -> identity
    f, c = 517.97, torch.tensor([320.0, 240.0])
    K = torch.tensor([[f, 0, c[0]], [0, f, c[1]], [0, 0, 1.0]])[None].repeat(4, 1, 1)
    mk0, mk1, bids, Rs, ts = [], [], [], [], []
    for b, n in enumerate(sizes):
        p1, p2, w, R, t = synth.two_view_geometry(1, max(n, 8), seed=50 + b, noise=2e-4, outlier_frac=0.3)
        mk0.append((p1[0] * f + c)[:n]); mk1.append((p2[0] * f + c)[:n])
        bids.append(torch.full((n,), b, dtype=torch.int64)); Rs.append(R[0]); ts.append(t[0])
    data = {"mkpts0_f": cu(torch.cat(mk0)), "mkpts1_f": cu(torch.cat(mk1)), "m_bids": cu(torch.cat(bids))}
    # prior: truth rotated by ~3 degrees about z, translation perturbed
    ang = 0.05
    Rz = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]], dtype=torch.float32)
    prior = torch.stack([torch.cat([Rz @ Rs[b], (ts[b] + 0.05)[:, None]], 1) for b in range(4)])
    gen = torch.Generator(device=DEV).manual_seed(1)
    Rt = prior_ransac_round(data, cu(K), cu(K), cu(prior), batch_size=1024, inl_th=3e-7 * 1e3, generator=gen).cpu()
    for b in (0, 2, 3):
        cosang = ((Rt[b, :, :3].T @ Rs[b]).trace() - 1) / 2
        assert torch.rad2deg(torch.arccos(cosang.clamp(-1, 1))) < 3.0, f"pair {b}: rotation error"   # best MINIMAL-sample model (max_lo_iters = 0)
        assert torch.rad2deg(torch.arccos((Rt[b, :, 3] @ ts[b]).clamp(-1, 1))) < 30.0, f"pair {b}: translation direction"   # weakly constrained by a minimal sample; the prior is 5 deg off too
        n_in = int(data["num_correspondences_after_ransac"][b])
        assert 0.6 * sizes[b] <= n_in <= 0.75 * sizes[b], (b, n_in)       # 70 % inliers by construction
        assert int(data["inliers_best_tight"][b]) <= n_in and int(data["inliers_best_ultra_tight"][b]) <= int(data["inliers_best_tight"][b])
    assert torch.equal(Rt[1], torch.eye(3, 4)) and int(data["num_correspondences_after_ransac"][1]) == 0
    assert int(data["ransac_inlier_mask"].sum()) == int(data["num_correspondences_after_ransac"].sum())
    assert data["num_correspondences_before_ransac"].tolist() == sizes


def test_pipeline_with_prior_ransac_round_runs():
    from far_b200.loftr import LoFTR, far_eval_cfg
    from far_b200.pipeline import FarPosePipeline
    cfg = far_eval_cfg(0.0)
    model = LoFTR(cfg)
    model.load_state_dict(synth.synth_state_dict(model.state_dict(), 3), strict=True)
    model = model.to(DEV).eval()
    img0, img1 = synth.synth_pair_images(2, seed=9)
    K = cu(synth.mp3d_intrinsics(2))
    out = FarPosePipeline(model, K, K, prior_ransac=True, ransac_kwargs={"batch_size": 512})(cu(img0), cu(img1))
    assert torch.isfinite(out["pose"]).all() and torch.isfinite(out["loftr_rt"]).all()
    assert out["data"]["ransac_scores"].shape == (2, 512)
    R = out["loftr_rt"][:, :, :3].double().cpu()
    assert (R @ R.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-4
