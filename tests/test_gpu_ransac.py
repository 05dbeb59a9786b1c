"""Prior-guided RANSAC round on the GPU: the deterministic scoring step against the fixture the
UNMODIFIED reference produced (tests/golden/make_golden_ransac.py: RANSAC.verify / get_prior_estimate /
remove_bad_models of mp3d_loftr/third_party/prior_ransac/ransac.py), candidate selection, and the whole stochastic
round through size-independent properties (it must recover the true pose of synthetic two-view geometry with 30 %
outliers; ragged pairs, an unsolvable pair)."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import f_distance

from oracle import far_oracle as O
from far_b200 import ops, synth
from far_b200.ransac import prior_ransac_round, ransac_round, normalise_prior
from far_b200 import solver as fsolver

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
    assert torch.equal((mask0.cpu() & 1).bool(), mo[0]) and c30[0].tolist() == [int(m.sum()) for m in mo]
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
    assert c3[0].tolist() == [int(m.sum()) for m in mw_]
    assert torch.equal((mask.cpu() & 1).bool(), mw_[0]) and torch.equal((mask.cpu() & 2).bool(), mw_[1]) \
        and torch.equal((mask.cpu() & 4).bool(), mw_[2]), "tight / ultra-tight bits of the mask"


def test_segment_offsets_ragged():
    for sizes in ([3, 0, 0, 5, 1, 0], [0, 0, 4], [7], [0, 0]):
        bids = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
        off = ops.segment_offsets(cu(bids), len(sizes)).cpu()
        assert off.tolist() == np.concatenate([[0], np.cumsum(sizes)]).tolist(), sizes


def _pixel_pairs(sizes, f=517.97, c=(320.0, 240.0), seed0=50, outlier_frac=0.3, noise=2e-4):
    c = torch.tensor(c)
    K = torch.tensor([[f, 0, c[0]], [0, f, c[1]], [0, 0, 1.0]])[None].repeat(len(sizes), 1, 1)
    mk0, mk1, bids, Rs, ts = [], [], [], [], []
    for b, n in enumerate(sizes):
        p1, p2, w, R, t = synth.two_view_geometry(1, max(n, 8), seed=seed0 + b, noise=noise, outlier_frac=outlier_frac)
        mk0.append((p1[0] * f + c)[:n]); mk1.append((p2[0] * f + c)[:n])
        bids.append(torch.full((n,), b, dtype=torch.int64)); Rs.append(R[0]); ts.append(t[0])
    return torch.cat(mk0), torch.cat(mk1), torch.cat(bids), K, Rs, ts


def test_ransac_sampling_contract_uniform_and_minimal_models():
    """Without a prior the CDF is exact (1, 2, ..., n), so the device sampler must reproduce the numpy Philox contract
    of oracle.ransac_sample_indices bit for bit; each model is the in-repo 8-point of its own sample."""
    sizes, H, seed = [60, 3, 700], 64, 1234567890123
    mk0, mk1, bids, K, _, _ = _pixel_pairs(sizes)
    off = ops.segment_offsets(cu(bids), len(sizes))
    models, idx = ops.ransac_sample_models(cu(mk0), cu(mk1), off, cu(K), cu(K), None, 0.1, H, seed, return_indices=True)
    idx, models = idx.cpu(), models.cpu()
    o = np.concatenate([[0], np.cumsum(sizes)])
    for b, n in enumerate(sizes):
        ref = O.ransac_sample_indices(np.ones(n), H, seed, pair=b)
        assert np.array_equal(idx[b].numpy().astype(np.int64), ref), f"pair {b}: sample indices"
        if n < 8:
            assert (models[b] == 0).all()
            continue
        assert all(len(set(r.tolist())) == 8 for r in ref), "distinct indices within a sample"
        f, c = K[b, 0, 0], K[b, :2, 2]
        k0 = ((mk0[o[b]:o[b + 1]] - c) / f).double()
        k1 = ((mk1[o[b]:o[b + 1]] - c) / f).double()
        sel = torch.from_numpy(ref)
        Fo = O.run_8point(k0[sel], k1[sel], torch.ones(H, 8, dtype=torch.float64))
        d = f_distance(models[b], Fo)
        # an exact fit through 8 points: the smallest eigenvalue is ~0 and near-degenerate samples amplify fp32 noise
        assert d.median() < 1e-4 and (d < 1e-2).float().mean() > 0.9, (d.median(), (d < 1e-2).float().mean())


def test_ransac_sampling_contract_biased(golden_dir):
    """With a prior the weights are exp(-sym_epipolar / sigma^2) + 1e-4 (ransac.py:358-367, 166-168): the device draws
    must equal the oracle contract on the oracle's fp64 weights except where u*total falls within float noise of a CDF
    step, and the empirical distribution must follow the weights."""
    g = np.load(os.path.join(golden_dir, "ransac.npz"))
    kp1, kp2, prior_rt = torch.from_numpy(g["kp1"]), torch.from_numpy(g["kp2"]), torch.from_numpy(g["prior_rt"])
    n, H, seed = kp1.shape[0], 2048, 42
    K = torch.eye(3)[None]
    off = _offsets([n])
    _, idx = ops.ransac_sample_models(cu(kp1), cu(kp2), cu(off), cu(K), cu(K), cu(prior_rt[None] * 1.0), 0.1, H, seed,
                                      return_indices=True)
    w = (O.ransac_bias_weight(kp1.double(), kp2.double(), prior_rt.double(), 0.1) + 1e-4).numpy()
    assert np.abs(w - 1e-4 - g["bias_ref"]).max() < 1e-5          # the oracle's weights are the reference's
    ref = O.ransac_sample_indices(w, H, seed, pair=0)
    got = idx[0].cpu().numpy().astype(np.int64)
    assert (got != ref).mean() < 5e-3, (got != ref).mean()
    mass = np.bincount(got.reshape(-1), minlength=n) / got.size
    top = np.argsort(-w)[: n // 10]
    assert abs(mass[top].sum() - (w[top].sum() / w.sum())) < 0.03


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
    mk0, mk1, bids, K, Rs, ts = _pixel_pairs(sizes)
    data = {"mkpts0_f": cu(mk0), "mkpts1_f": cu(mk1), "m_bids": cu(bids)}
    # prior: truth rotated by ~3 degrees about z, translation perturbed
    ang = 0.05
    Rz = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]], dtype=torch.float32)
    prior = torch.stack([torch.cat([Rz @ Rs[b], (ts[b] + 0.05)[:, None]], 1) for b in range(4)])
    for pr in (cu(prior), None):                            # prior-guided round, then the `prior_ransac_noprior` round
        Rt = prior_ransac_round(data, cu(K), cu(K), pr, batch_size=1024, inl_th=3e-7 * 1e3, seed=1).cpu()
        for b in (0, 2, 3):
            cosang = ((Rt[b, :, :3].T @ Rs[b]).trace() - 1) / 2
            assert torch.rad2deg(torch.arccos(cosang.clamp(-1, 1))) < 6.0, f"pair {b}: rotation error"   # best MINIMAL-sample (8 noisy points) model of 1024, no local optimisation (max_lo_iters = 0)
            assert torch.rad2deg(torch.arccos((Rt[b, :, 3] @ ts[b]).clamp(-1, 1))) < 45.0, f"pair {b}: translation direction"   # weakly constrained by a minimal sample; the prior is 5 deg off too
            n_in = int(data["num_correspondences_after_ransac"][b])
            assert 0.6 * sizes[b] <= n_in <= 0.75 * sizes[b], (b, n_in)       # 70 % inliers by construction
            assert int(data["inliers_best_tight"][b]) <= n_in and int(data["inliers_best_ultra_tight"][b]) <= int(data["inliers_best_tight"][b])
        assert torch.equal(Rt[1], torch.eye(3, 4)) and int(data["num_correspondences_after_ransac"][1]) == 0
        assert int((data["ransac_inlier_mask"] & 1).sum()) == int(data["num_correspondences_after_ransac"].sum())
        assert data["num_correspondences_before_ransac"].tolist() == sizes
    # reproducible: same seed -> same winner; zero-translation prior must not poison the scores
    r1 = ransac_round(cu(mk0), cu(mk1), cu(bids), cu(K), cu(K), cu(prior), batch_size=256, inl_th=3e-4, seed=7)
    r2 = ransac_round(cu(mk0), cu(mk1), cu(bids), cu(K), cu(K), cu(prior), batch_size=256, inl_th=3e-4, seed=7)
    assert torch.equal(r1["best"], r2["best"]) and torch.equal(r1["scores"], r2["scores"])
    zero_t = prior.clone(); zero_t[:, :, 3] = 0
    r3 = ransac_round(cu(mk0), cu(mk1), cu(bids), cu(K), cu(K), cu(zero_t), batch_size=256, inl_th=3e-4, seed=7)
    assert int(r3["best"][0]) >= 0 and torch.isfinite(r3["scores"][0][r3["best"][0]])


def test_reference_signature_solver_shims():
    """estimate_pose (metrics.py:80-174), EssentialMatrixSolver.estimate_pose (pose_solver.py:30-97) and
    RANSAC.forward (ransac.py:340) with the argument shapes spvs_RT / RegressionModel.forward pass."""
    mk0, mk1, bids, K, Rs, ts = _pixel_pairs([600], outlier_frac=0.2)
    K0 = K[0].double()                                       # the reference hands K as [3,3] f64 (datasets/mp3d.py)

    def ang(R):
        return float(torch.rad2deg(torch.arccos((((R.float().cpu().T @ Rs[0]).trace() - 1) / 2).clamp(-1, 1))))

    # --- estimate_pose
    prior = torch.cat([Rs[0], ts[0][:, None]], 1).numpy()
    for solver, pr in (("ransac", None), ("prior_ransac", prior), ("prior_ransac", None), ("prior_ransac_noprior", None)):
        ret, n_after, n_tight, n_ultra = fsolver.estimate_pose(cu(mk0), cu(mk1), cu(K0), cu(K0), 0.5, conf=0.99999,
                                                               translation_scale=None, solver=solver, priorRT=pr)
        R, t, inl, E = ret
        assert R.shape == (3, 3) and t.shape == (3,) and R.is_cuda and R.dtype == torch.float64
        assert inl.dtype == np.bool_ and inl.shape == (600,) and int(n_after) == int(inl.sum())
        assert ang(R) < 3.0 and 0.6 * 600 <= int(n_after) <= 0.85 * 600, (solver, ang(R), int(n_after))
        if solver == "ransac" or pr is None and solver == "prior_ransac":
            assert n_tight == 0 and n_ultra == 0             # the cv2 branch never fills them (:96)
        else:
            assert int(n_after) >= n_tight >= n_ultra >= 0
    for n in (0, 4, 7):                                      # < 5 -> None (:82-84); 5..7 -> no model -> None (:157-159)
        assert fsolver.estimate_pose(cu(mk0[:n]), cu(mk1[:n]), cu(K0), cu(K0), 0.5) == (None, 0, 0, 0)
    # --- EssentialMatrixSolver (numpy keypoints, K in a data dict, as RegressionModel.forward passes them)
    cfg = {"EMAT_RANSAC": {"PIX_THRESHOLD": 2.0, "SCALE_THRESHOLD": 0.1, "CONFIDENCE": 0.9999}}
    data2 = {"K_color0": K[:1].clone(), "K_color1": K[:1].clone()}
    for use_prior, pr in ((False, None), (True, cu(torch.cat([Rs[0], ts[0][:, None]], 1)))):
        es = fsolver.EssentialMatrixSolver(cfg, use_prior)
        (R, t, n), tight, ultra = es.estimate_pose(mk0.numpy(), mk1.numpy(), data2, pr)
        assert isinstance(R, np.ndarray) and R.shape == (3, 3) and t.shape == (3,) and isinstance(n, int)
        assert ang(torch.from_numpy(R)) < 3.0 and n > 300
    (R, t, n), tight, ultra = fsolver.EssentialMatrixSolver(cfg).estimate_pose(mk0.numpy()[:3], mk1.numpy()[:3], data2)
    assert np.array_equal(R, np.eye(3)) and n == 0 and (tight, ultra) == (0, 0)
    # --- RANSAC.forward on K-normalised keypoints (metrics.py:114-128)
    kp1 = cu((mk0 - K[0, :2, 2]) / K[0, 0, 0])
    kp2 = cu((mk1 - K[0, :2, 2]) / K[0, 0, 0])
    pp = {"rotation_pcl_error": True, "rotation_error": False, "K1": cu(K[0]), "K2": cu(K[0]),
          "RT": cu(torch.from_numpy(prior)), "pcl": cu(torch.rand(300, 3) * 6 - 3), "lambda": 0.3, "biased_sampling": "biased"}
    rs = fsolver.RANSAC(model_type="essential_cv2", max_iter=1, inl_th=3e-7 * 1e3, prior_params=pp, max_lo_iters=0,
                        batch_size=2048, use_noexp_prior_scoring=True, use_linear_bias_sampling=True, bias_sigma_sq=0.1)
    E, mask, tight, ultra = rs.forward(kp1=kp1, kp2=kp2)
    assert E.shape == (3, 3) and mask.dtype == torch.bool and mask.shape == (600,)
    assert int(mask.sum()) >= int(tight.sum()) >= int(ultra.sum()) and int(mask.sum()) > 360
    err = O.sampson_epipolar_distance(kp1.cpu()[None], kp2.cpu()[None], E.cpu()[None])[0]
    assert int(((err <= 3e-4) ^ mask.cpu()).sum()) <= 2, "the returned mask is the Sampson inlier set of the returned model"


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


def _five_point_samples(S, seed):
    """S minimal samples (5 calibrated correspondences) of random two-view geometry, exact (no noise)."""
    p1, p2, _, R, t = synth.two_view_geometry(S, 5, seed=seed, noise=0.0, outlier_frac=0.0)
    return torch.cat([p1, p2], -1).double(), R, t


def _set_distance(Ea, Eb):
    """max over the matrices of Ea of the distance (up to sign, unit Frobenius norm) to the closest matrix of Eb."""
    if len(Ea) == 0:
        return 0.0
    if len(Eb) == 0:
        return float("inf")
    a = np.stack([e.reshape(-1) / np.linalg.norm(e) for e in Ea])
    b = np.stack([e.reshape(-1) / np.linalg.norm(e) for e in Eb])
    d = np.minimum(np.linalg.norm(a[:, None] - b[None], axis=-1), np.linalg.norm(a[:, None] + b[None], axis=-1))
    return float(d.min(1).max())


def test_five_point_solver_vs_oracle():
    """csrc/fivept.cuh (Nister 5-point, one thread per sample) against oracle.run_5point_nister (the numpy restatement of
    cv_geometry.py:861-1041) on 256 exact minimal samples: the same SET of real solutions (up to sign), every solution
    satisfies the five epipolar equations, the trace constraint and det E = 0, and one of them is the true E."""
    S = 256
    pts, R, t = _five_point_samples(S, 31)
    E, ns = ops.five_point(cu(pts))
    E, ns = E.cpu().numpy(), ns.cpu().numpy()
    assert ns.min() >= 1 and ns.max() <= 10
    mismatched, inexact, truth_d = 0, 0, []
    for s in range(S):
        p = pts[s].numpy()
        Eo = O.run_5point_nister(p[:, :2], p[:, 2:])
        Eg = E[s, :ns[s]]
        x1 = np.concatenate([p[:, :2], np.ones((5, 1))], 1)
        x2 = np.concatenate([p[:, 2:], np.ones((5, 1))], 1)
        worst = 0.0
        for e in Eg:
            assert abs(np.linalg.norm(e) - 1.0) < 1e-9
            assert np.abs(np.einsum("ni,ij,nj->n", x2, e, x1)).max() < 1e-8, "epipolar equations of the sample"
            worst = max(worst, abs(np.linalg.det(e)), np.abs(e @ e.T @ e - 0.5 * np.trace(e @ e.T) * e).max())
        inexact += worst > 1e-8        # det E = 0 / trace constraint: only a (near-)double root may miss 1e-8
        tx = np.array([[0, -t[s, 2], t[s, 1]], [t[s, 2], 0, -t[s, 0]], [-t[s, 1], t[s, 0], 0]], dtype=np.float64)
        truth_d.append(_set_distance([tx @ R[s].double().numpy()], Eg))
        # same solution set as the oracle; a pair of nearly coincident real roots may be classified differently by the
        # two root finders (companion-matrix eigenvalues vs Aberth iteration): counted, must be rare
        if len(Eo) != len(Eg) or _set_distance(Eo, Eg) > 1e-6 or _set_distance(Eg, Eo) > 1e-6:
            mismatched += 1
    truth_d = np.array(truth_d)
    print(f"[five_point] samples whose real-root set differs from the oracle's: {mismatched} / {S}; with a solution "
          f"missing the cubic constraints by > 1e-8: {inexact}; distance of the true E to the closest solution: median "
          f"{np.median(truth_d):.1e}, 90 % {np.percentile(truth_d, 90):.1e} (the synthetic points are fp32-rounded)")
    assert mismatched <= S // 50 and inexact <= S // 50
    assert np.median(truth_d) < 5e-6 and np.percentile(truth_d, 90) < 1e-4


def test_ransac_round_five_point_minimal_solver():
    """The round with the recipe's model type (5-point on 5 + 1 draws): sampling contract (6 distinct Philox draws,
    bit-equal to the numpy restatement), every hypothesis' model is an exact solution for its first five draws, and the
    round recovers the true pose of a ragged batch with 30 % outliers."""
    sizes = [900, 5, 400, 1300]
    mk0, mk1, bids, K, Rs, ts = _pixel_pairs(sizes)
    off = ops.segment_offsets(cu(bids), len(sizes))
    H = 512
    models, idx = ops.ransac_sample_models(cu(mk0), cu(mk1), off, cu(K), cu(K), None, 0.1, H, 5, return_indices=True,
                                           minimal_solver="5pt")
    models, idx = models.cpu(), idx.cpu()
    assert idx.shape == (4, H, 6) and (idx[1] == -1).all() and (models[1] == 0).all()
    for b in (0, 2, 3):
        ref = O.ransac_sample_indices(np.ones(sizes[b]), H, 5, pair=b, S=6)
        assert np.array_equal(idx[b].numpy().astype(np.int64), ref), f"pair {b}: Philox draws"
        assert all(len(set(row.tolist())) == 6 for row in idx[b])
    # K-normalised points of pair 0; each model satisfies the epipolar equation of its own first five draws
    o0 = int(off[0])
    x0 = (mk0[o0:o0 + sizes[0]] - K[0, :2, 2]) / torch.stack([K[0, 0, 0], K[0, 1, 1]])
    x1 = (mk1[o0:o0 + sizes[0]] - K[0, :2, 2]) / torch.stack([K[0, 0, 0], K[0, 1, 1]])
    solved = 0
    for h in range(H):
        Em = models[0, h].double()
        if Em.abs().max() == 0:
            continue
        solved += 1
        sel = idx[0, h, :5].long()
        a = torch.cat([x0[sel], torch.ones(5, 1)], 1).double()
        b = torch.cat([x1[sel], torch.ones(5, 1)], 1).double()
        assert torch.einsum("ni,ij,nj->n", b, Em, a).abs().max() < 1e-5, h     # fp32 model, fp32 pixel keypoints
    assert solved >= 0.98 * H
    data = {"mkpts0_f": cu(mk0), "mkpts1_f": cu(mk1), "m_bids": cu(bids)}
    Rt = prior_ransac_round(data, cu(K), cu(K), None, batch_size=1024, inl_th=3e-7 * 1e3, seed=1, minimal_solver="5pt").cpu()
    for b in (0, 2, 3):
        cosang = ((Rt[b, :, :3].T @ Rs[b]).trace() - 1) / 2
        assert torch.rad2deg(torch.arccos(cosang.clamp(-1, 1))) < 6.0, f"pair {b}: rotation error"
        n_in = int(data["num_correspondences_after_ransac"][b])
        assert 0.6 * sizes[b] <= n_in <= 0.75 * sizes[b], (b, n_in)
    assert torch.equal(Rt[1], torch.eye(3, 4))


def test_ransac_shim_five_point_model_type():
    """RANSAC('essential_cv2', minimal_solver='5pt').forward on K-normalised keypoints: the recipe's model type end to end."""
    mk0, mk1, bids, K, Rs, ts = _pixel_pairs([700], outlier_frac=0.25)
    f, c = K[0, 0, 0], K[0, :2, 2]
    kp1, kp2 = (mk0 - c) / f, (mk1 - c) / f
    r = fsolver.RANSAC('essential_cv2', inl_th=3e-4, batch_size=1024, max_iter=2, max_lo_iters=0, minimal_solver='5pt', seed=3)
    assert r.minimal_sample_size == 6
    E, inl, tight, ultra = r(cu(kp1), cu(kp2))
    assert E.shape == (3, 3) and 0.6 * 700 <= int(inl.sum()) <= 0.8 * 700
    assert int(tight.sum()) <= int(inl.sum()) and int(ultra.sum()) <= int(tight.sum())
    tx = torch.tensor([[0, -ts[0][2], ts[0][1]], [ts[0][2], 0, -ts[0][0]], [-ts[0][1], ts[0][0], 0]])
    assert f_distance(E[None].cpu(), (tx @ Rs[0])[None]).max() < 0.15   # best MINIMAL-sample model, no local optimisation
