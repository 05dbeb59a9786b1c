"""Geometric solver functions with the reference's names/signatures (SURVEY.md 8b "solver fns"):
  run_8point(points1, points2, weights=None) -> F            third_party/prior_ransac/cv_geometry.py:772-833
  decompose_essential_matrix(E) -> (R1, R2, t)               third_party/prior_ransac/essential.py:99-139
  motion_from_essential(E) -> (Rs [*,4,3,3], ts [*,4,3,1])    essential.py:41-64
  estimate_pose_batched(...)                                  the spvs_RT per-pair loop (supervision.py:184-233)
"""
import torch

from . import ops


def run_8point(points1, points2, weights=None):
    if points1.shape != points2.shape:
        raise AssertionError(points1.shape, points2.shape)
    if points1.shape[1] < 8:
        raise AssertionError(points1.shape)
    if weights is not None and not (len(weights.shape) == 2 and weights.shape[1] == points1.shape[1]):
        raise AssertionError(weights.shape)
    return ops.eight_point(points1, points2, weights)


def decompose_essential_matrix(E_mat):
    if not (len(E_mat.shape) >= 2 and E_mat.shape[-2:] == (3, 3)):
        raise AssertionError(E_mat.shape)
    return ops.essential_decompose(E_mat)


def motion_from_essential(E_mat):
    R1, R2, t = decompose_essential_matrix(E_mat)
    return torch.stack([R1, R1, R2, R2], dim=-3), torch.stack([t, -t, t, -t], dim=-3)


def estimate_pose_batched(data, K0, K1):
    """Vectorised replacement of spvs_RT's python loop over pairs (supervision.py:209-233) with the in-repo weighted
    8-point as the model solver: reads m_bids, mkpts0_f, mkpts1_f, mconf from `data`;
    writes loftr_rt [N,3,4], expec_e [N,3,3] and the counter keys FAR's head consumes.  Pairs with < 8 matches get
    the identity pose (the reference's `ret is None` fallback, :222-224)."""
    N = K0.shape[0]
    m_bids = data['m_bids']
    dev = m_bids.device
    counts = torch.bincount(m_bids, minlength=N)
    offsets = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(counts, 0)
    E, Rt, npos = ops.pose_from_matches(data['mkpts0_f'], data['mkpts1_f'], data['mconf'], offsets,
                                        K0.to(dev).float(), K1.to(dev).float())
    data.update({'loftr_rt': Rt, 'expec_rt': Rt, 'expec_e': E,
                 'num_correspondences_before_ransac': counts,
                 'num_correspondences_after_ransac': npos.to(torch.int64),
                 'num_correspondences': npos.to(torch.int64),
                 'inliers_best_tight': torch.zeros_like(counts),
                 'inliers_best_ultra_tight': torch.zeros_like(counts)})
    return Rt
