"""Prior-guided RANSAC round on the GPU, batched over pairs.

What `estimate_pose(..., solver='prior_ransac', priorRT=...)` does for ONE pair on the CPU/GPU mix of the reference
(mp3d_loftr/src/utils/metrics.py:100-123 -> third_party/prior_ransac/ransac.py:340-442, `max_iter=1`, `batch_size=2048`):

    bias weights     exp(-symmetrical_epipolar(kp0, kp1, E_prior) / sigma^2) + 1e-4        ransac.py:358-367, 166-168
    biased sampling  2048 minimal samples, with replacement, p ~ bias weights             :161-175 (numpy RNG there)
    minimal solver   one model per sample                                                 :250-253
    prior score      -(pcl error of the decomposed model vs the prior pose)^2 / lambda    :203-231, 401-404
    verify           Sampson inliers at 3 thresholds, argmax(inliers + prior score)       :256-292
    pose             cv2.recoverPose on the winner                                         metrics.py:164-170

here runs for all pairs of the batch at once on the device: sampling with `torch.multinomial`, the minimal solver is the
in-repo normalised 8-point on 8 correspondences (`far_eight_point`; the recipe's `essential_cv2` model calls OpenCV's
5-point solver 2048 times per pair on the CPU -- un-vendored arithmetic), scoring and selection in
`far_prior_ransac_score`, candidate selection by the cheirality vote of `far_pose_from_essential` over the inliers.
Sampling is stochastic (as in the reference): parity is pinned on the deterministic scoring step
(tests/golden/ransac.npz, produced by the unmodified `RANSAC.verify` / `get_prior_estimate`).
"""
import torch

from . import ops


def _skew_times_R(rt):
    """E_prior = [t]_x R of a [N,3,4] pose (ransac.py:63-71 `fundamental_from_RT`, which returns E)."""
    R, t = rt[:, :, :3], rt[:, :, 3]
    z = torch.zeros_like(t[:, 0])
    tx = torch.stack([torch.stack([z, -t[:, 2], t[:, 1]], -1), torch.stack([t[:, 2], z, -t[:, 0]], -1),
                      torch.stack([-t[:, 1], t[:, 0], z], -1)], -2)
    return tx @ R


def normalise_prior(prior_rt):
    """setup_prior (ransac.py:176-186): unit-norm translation."""
    p = prior_rt.float().clone()
    p[:, :, 3] = p[:, :, 3] / p[:, :, 3].norm(dim=1, keepdim=True).clamp_min(1e-12)
    return p


def bias_weights(kp0, kp1, m_bids, prior_rt, sigma_sq=0.1):
    """exp(-symmetrical_epipolar_distance / sigma^2) per match (ransac.py:358-367, use_linear_bias_sampling)."""
    E = _skew_times_R(prior_rt)[m_bids]                                       # [M,3,3]
    p0 = torch.cat([kp0, torch.ones_like(kp0[:, :1])], 1)
    p1 = torch.cat([kp1, torch.ones_like(kp1[:, :1])], 1)
    l = torch.einsum('mij,mj->mi', E, p0)                                      # E p0: line in image 1
    m = torch.einsum('mji,mj->mi', E, p1)                                      # E^T p1
    num = (p1 * l).sum(-1) ** 2
    d = num * (1.0 / (l[:, 0] ** 2 + l[:, 1] ** 2) + 1.0 / (m[:, 0] ** 2 + m[:, 1] ** 2))
    return torch.exp(-d / sigma_sq)


@torch.no_grad()
def prior_ransac_round(data, K0, K1, prior_rt, batch_size=2048, inl_th=3e-7, prior_lambda=0.3, bias_sigma_sq=0.1,
                       biased=True, pcl=None, generator=None):
    """One prior-guided RANSAC round for every pair of the batch.  Reads m_bids, mkpts0_f, mkpts1_f from `data`;
    prior_rt [N,3,4] (e.g. the FAR head's prediction, loftr.py:187-192).  Writes the same keys estimate_pose_batched
    writes (loftr_rt, num_correspondences*, inliers_best_tight / ultra_tight) and returns loftr_rt [N,3,4]."""
    mk0, mk1, m_bids = data['mkpts0_f'], data['mkpts1_f'], data['m_bids']
    dev = mk0.device
    N = K0.shape[0]
    K0, K1 = K0.to(dev).float(), K1.to(dev).float()
    M = int(mk0.shape[0])
    counts = torch.bincount(m_bids, minlength=N)
    offsets = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(counts, 0)
    prior = normalise_prior(prior_rt.to(dev)) if prior_rt is not None else None
    if pcl is None:  # metrics.py:103: 300 points uniform in [-3, 3]^3
        pcl = torch.rand(300, 3, device=dev, generator=generator) * 6.0 - 3.0
    if M == 0:
        eye = torch.eye(3, 4, device=dev).repeat(N, 1, 1)
        zero = torch.zeros(N, dtype=torch.int64, device=dev)
        data.update({'loftr_rt': eye, 'expec_rt': eye, 'num_correspondences_before_ransac': counts,
                     'num_correspondences_after_ransac': zero, 'num_correspondences': zero,
                     'inliers_best_tight': zero, 'inliers_best_ultra_tight': zero})
        return eye
    # K-normalised keypoints (metrics.py:88-89)
    k0, k1 = K0[m_bids], K1[m_bids]
    kp0 = (mk0 - k0[:, :2, 2]) / torch.stack([k0[:, 0, 0], k0[:, 1, 1]], -1)
    kp1 = (mk1 - k1[:, :2, 2]) / torch.stack([k1[:, 0, 0], k1[:, 1, 1]], -1)
    # sampling weights, padded to [N, max matches]
    w = bias_weights(kp0, kp1, m_bids, prior, bias_sigma_sq) + 1e-4 if (biased and prior is not None) \
        else torch.ones(M, device=dev)
    max_m = int(counts.max())                       # one host sync (the reference round-trips through numpy here)
    local = torch.arange(M, device=dev) - offsets[m_bids]
    W = torch.zeros(N, max(max_m, 1), device=dev)
    W[m_bids, local] = w
    W[counts < 8, 0] = 1.0                          # pairs that cannot be solved still need a valid distribution
    idx = torch.multinomial(W, batch_size * 8, replacement=True, generator=generator)      # [N, H*8] local indices
    gidx = (idx + offsets[:N, None]).clamp_(max=M - 1)
    p0 = kp0[gidx].reshape(N * batch_size, 8, 2)
    p1 = kp1[gidx].reshape(N * batch_size, 8, 2)
    models = ops.eight_point(p0, p1, None).reshape(N, batch_size, 3, 3)
    scores, best, best_E, counts3, mask = ops.prior_ransac_score(mk0, mk1, offsets, K0, K1, models, prior, pcl,
                                                                 prior_lambda, inl_th)
    Rt, npos = ops.pose_from_essential(mk0, mk1, mask, offsets, K0, K1, best_E)
    c3 = counts3.to(torch.int64)
    data.update({'loftr_rt': Rt, 'expec_rt': Rt, 'expec_e': best_E, 'ransac_inlier_mask': mask.bool(),
                 'ransac_best_index': best, 'ransac_scores': scores,
                 'num_correspondences_before_ransac': counts,
                 'num_correspondences_after_ransac': c3[:, 0], 'num_correspondences': c3[:, 0],
                 'inliers_best_tight': c3[:, 1], 'inliers_best_ultra_tight': c3[:, 2]})
    return Rt
