"""Prior-guided RANSAC round on the GPU, batched over pairs.

What `estimate_pose(..., solver='prior_ransac', priorRT=...)` does for ONE pair on the CPU/GPU mix of the reference
(mp3d_loftr/src/utils/metrics.py:100-123 -> third_party/prior_ransac/ransac.py:340-442, `max_iter=1`, `batch_size=2048`):

    bias weights     exp(-symmetrical_epipolar(kp0, kp1, E_prior) / sigma^2) + 1e-4        ransac.py:358-367, 166-168
    biased sampling  2048 minimal samples, with replacement, p ~ bias weights             :161-175 (numpy RNG there)
    minimal solver   one model per sample                                                 :250-253
    prior score      -(pcl error of the decomposed model vs the prior pose)^2 / lambda    :203-231, 401-404
    verify           Sampson inliers at 3 thresholds, argmax(inliers + prior score)       :256-292
    pose             cv2.recoverPose on the winner                                         metrics.py:164-170

here runs for all pairs of the batch at once, entirely on the device, as THREE C-ABI calls with no torch math and no
host synchronisation in between:

    far_ransac_sample_models   bias weights + per-pair CDF + Philox inverse-CDF sampling + the minimal solver: the
                               in-repo normalised 8-point on each 8-sample (default), or -- the recipe's model type,
                               `essential_cv2` = OpenCV's 5-point on 6 points, 2048 CPU calls per pair there -- Nister's
                               5-point on 5 + 1 draws (minimal_solver='5pt'; csrc/fivept.cuh restates the in-tree
                               batched 5-point cv_geometry.py:861-1041; OpenCV's own arithmetic is un-vendored)
    far_prior_ransac_score     prior score, Sampson inliers, argmax, masks and the 3 counters
    far_pose_from_essential    cheirality vote over the inliers (cv2.recoverPose's criterion)

(plus `far_segment_offsets` to turn the sorted `m_bids` into segment offsets).  Without a prior
(`prior_rt=None`) the same round is the reference's `solver='prior_ransac_noprior'` branch (metrics.py:124-143):
uniform sampling, inlier count only.  Sampling is stochastic in the reference (numpy's global RNG); here it is a
counter-based generator keyed by (seed, pair, hypothesis), so the samples are reproducible and pinned by a numpy
Philox in tests/test_gpu_ransac.py; the deterministic scoring step is pinned to the unmodified
`RANSAC.verify` / `get_prior_estimate` (tests/golden/ransac.npz).
"""
import torch

from . import ops

_PCL = {}


def default_pcl(device, seed=0):
    """metrics.py:103: 300 points uniform in [-3, 3]^3 (the reference redraws them per call with numpy's RNG; a fixed,
    seeded cloud per device keeps the round deterministic and costs no launch)."""
    key = (str(device), seed)
    if key not in _PCL:
        g = torch.Generator().manual_seed(20240000 + seed)
        _PCL[key] = (torch.rand(300, 3, generator=g) * 6.0 - 3.0).to(device)
    return _PCL[key]


def normalise_prior(prior_rt):
    """setup_prior (ransac.py:176-186): unit-norm translation."""
    p = prior_rt.float().clone()
    p[:, :, 3] = p[:, :, 3] / p[:, :, 3].norm(dim=1, keepdim=True).clamp_min(1e-12)
    return p


@torch.no_grad()
def ransac_round(mkpts0, mkpts1, m_bids, K0, K1, prior_rt=None, batch_size=2048, inl_th=3e-7, prior_lambda=0.3,
                 bias_sigma_sq=0.1, biased=True, pcl=None, seed=0, offsets=None, minimal_solver="8pt"):
    """One (prior-guided) RANSAC round for every pair of a ragged batch.  minimal_solver: '8pt' (default: the in-repo
    normalised 8-point) or '5pt' (the recipe's model type: Nister's 5-point on 5 + 1 draws, ops.ransac_sample_models).  mkpts* [M,2] pixel keypoints, m_bids [M]
    sorted pair ids, K0/K1 [N,3,3], prior_rt [N,3,4] or None.  Returns a dict of device tensors:
    Rt [N,3,4], E [N,3,3], mask [M] uint8 (bit 0 inlier / 1 tight / 2 ultra tight), counts3 [N,3] int32, best [N],
    scores [N,H], counts [N] (matches per pair), n_pos [N]."""
    dev = mkpts0.device
    N = K0.shape[0]
    K0, K1 = K0.to(dev).float().contiguous(), K1.to(dev).float().contiguous()
    if offsets is None:
        offsets = ops.segment_offsets(m_bids, N)
    if mkpts0.shape[0] == 0:   # no match in the whole batch: the kernels still want non-null keypoint pointers
        mkpts0 = mkpts1 = torch.zeros(1, 2, device=dev)
    prior = prior_rt.to(dev).float().contiguous() if prior_rt is not None else None
    models = ops.ransac_sample_models(mkpts0, mkpts1, offsets, K0, K1, prior if biased else None, bias_sigma_sq,
                                      batch_size, seed, minimal_solver=minimal_solver)
    if prior is not None and pcl is None:
        pcl = default_pcl(dev)
    scores, best, best_E, counts3, mask = ops.prior_ransac_score(mkpts0, mkpts1, offsets, K0, K1, models, prior, pcl,
                                                                 prior_lambda, inl_th)
    Rt, npos = ops.pose_from_essential(mkpts0, mkpts1, mask, offsets, K0, K1, best_E)
    return {'Rt': Rt, 'E': best_E, 'mask': mask, 'counts3': counts3, 'best': best, 'scores': scores,
            'offsets': offsets, 'n_pos': npos}


@torch.no_grad()
def prior_ransac_round(data, K0, K1, prior_rt, batch_size=2048, inl_th=3e-7, prior_lambda=0.3, bias_sigma_sq=0.1,
                       biased=True, pcl=None, seed=0, minimal_solver="8pt"):
    """The round applied to a LoFTR `data` dict: reads m_bids, mkpts0_f, mkpts1_f; prior_rt [N,3,4] (e.g. the FAR
    head's prediction, loftr.py:187-192) or None for the reference's `prior_ransac_noprior` round.  Writes the keys
    spvs_RT writes (supervision.py:226-233: loftr_rt, num_correspondences*, inliers_best_tight / ultra_tight) with the
    reference's meaning -- `num_correspondences_after_ransac` IS the RANSAC inlier count -- and returns loftr_rt."""
    r = ransac_round(data['mkpts0_f'], data['mkpts1_f'], data['m_bids'], K0, K1, prior_rt, batch_size, inl_th,
                     prior_lambda, bias_sigma_sq, biased, pcl, seed, minimal_solver=minimal_solver)
    off = r['offsets']
    c3 = r['counts3'].to(torch.int64)
    data.update({'loftr_rt': r['Rt'], 'expec_rt': r['Rt'], 'expec_e': r['E'], 'ransac_inlier_mask': r['mask'],
                 'ransac_best_index': r['best'], 'ransac_scores': r['scores'],
                 'num_correspondences_before_ransac': off[1:] - off[:-1],
                 'num_correspondences_after_ransac': c3[:, 0], 'num_correspondences': c3[:, 0],
                 'inliers_best_tight': c3[:, 1], 'inliers_best_ultra_tight': c3[:, 2]})
    return r['Rt']
