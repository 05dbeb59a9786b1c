"""Geometric solver functions with the reference's names/signatures (SURVEY.md 8b "solver fns"):
  run_8point(points1, points2, weights=None) -> F            third_party/prior_ransac/cv_geometry.py:772-833
  decompose_essential_matrix(E) -> (R1, R2, t)               third_party/prior_ransac/essential.py:99-139
  motion_from_essential(E) -> (Rs [*,4,3,3], ts [*,4,3,1])    essential.py:41-64
  estimate_pose(kpts0, kpts1, K0, K1, thresh, conf, translation_scale, solver, priorRT)
      -> (ret | None, n_after, n_tight, n_ultra)              mp3d_loftr/src/utils/metrics.py:80-174
  EssentialMatrixSolver(cfg, use_prior_ransac).estimate_pose(kpts0, kpts1, data, priorRT)
      -> ((R, t, n_inliers), n_tight, n_ultra)                mapfree_6dreg/lib/models/matching/pose_solver.py:20-97
  RANSAC(model_type, ...).forward(kp1, kp2)
      -> (E, inliers, inliers_tight, inliers_ultra_tight)     third_party/prior_ransac/ransac.py:74-442
  estimate_pose_batched(...) / estimate_pose_ransac_batched   the spvs_RT per-pair loop (supervision.py:184-233)

The three reference-signature entry points are thin per-pair adapters over the batched device ops (one pair = a batch
of one), so `spvs_RT`, `compute_pose_errors` and `RegressionModel.forward` can call them unchanged.  Every shipped
recipe solves with OpenCV (`cv2.findEssentialMat` + `cv2.recoverPose`), whose arithmetic is not under
the reference tree: here the robust estimator is the GPU RANSAC round of far_b200/ransac.py (uniform sampling for
solver='ransac' / 'prior_ransac_noprior', prior-biased sampling + prior scoring for 'prior_ransac'), the minimal solver
the in-repo normalised 8-point, candidate selection the cheirality vote cv2.recoverPose applies.  PARITY UNPINNED at
the OpenCV call sites; pinned for run_8point / decompose_essential_matrix / verify / get_prior_estimate.
"""
import numpy as np
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


def estimate_pose_ransac_batched(data, K0, K1, prior_rt=None, **kw):
    """spvs_RT's loop with the GPU RANSAC round as the robust estimator (far_b200/ransac.py): the counters it writes
    (`num_correspondences_after_ransac`, `inliers_best_tight`, ...) are RANSAC inlier counts, the quantity the FAR gate
    MLP was trained on (supervision.py:226-233)."""
    from .ransac import prior_ransac_round
    return prior_ransac_round(data, K0, K1, prior_rt, **kw)


# ----------------------------------------------------------------------------------------------------------------------
# reference-signature adapters (one pair per call, exactly what the reference's python loops pass)
_MIN_MODEL_POINTS = 8   # the minimal solver is the in-repo 8-point (cv_geometry.py:787 asserts N >= 8)


def _ransac_one_pair(kpts0, kpts1, K0, K1, prior_rt, inl_th, biased, batch_size=2048, seed=0):
    from .ransac import ransac_round
    dev = kpts0.device if kpts0.is_cuda else torch.device('cuda', torch.cuda.current_device())
    k0 = torch.as_tensor(kpts0, dtype=torch.float32, device=dev).reshape(-1, 2)
    k1 = torch.as_tensor(kpts1, dtype=torch.float32, device=dev).reshape(-1, 2)
    bids = torch.zeros(k0.shape[0], dtype=torch.int64, device=dev)
    K0 = torch.as_tensor(K0, dtype=torch.float32, device=dev).reshape(1, 3, 3)
    K1 = torch.as_tensor(K1, dtype=torch.float32, device=dev).reshape(1, 3, 3)
    pr = None
    if prior_rt is not None:
        pr = torch.as_tensor(prior_rt, dtype=torch.float32, device=dev).reshape(-1, 4)[:3].reshape(1, 3, 4)
    return ransac_round(k0, k1, bids, K0, K1, pr, batch_size=batch_size, inl_th=inl_th, biased=biased, seed=seed)


def estimate_pose(kpts0, kpts1, K0, K1, thresh, conf=0.99999, translation_scale=None, solver='ransac', priorRT=None):
    """Drop-in for mp3d_loftr/src/utils/metrics.py:80-174.  kpts [M,2] pixel keypoints (tensor), K [3,3].
    Returns (ret, num_correspondences_after_ransac, inliers_best_tight, inliers_best_ultra_tight) with
    ret = (R [3,3] f64 cuda, t [3] f64 cuda, inlier mask [M] numpy bool, E [3,3] f64 cpu) or None.

    * fewer than 5 keypoints -> (None, 0, 0, 0), the reference's rule (:82-84);
    * 5..7 keypoints: the in-repo minimal solver needs 8 (cv_geometry.py:787), so no model exists: returned through the
      reference's `E is None` exit (:157-159), which spvs_RT maps to the same identity pose (supervision.py:222-224);
    * solver == 'prior_ransac' with a prior: inl_th 3e-7, biased sampling, prior scoring (:100-123);
      'prior_ransac_noprior': inl_th 3e-7, uniform (:124-143); anything else ('ransac'): uniform sampling with the
      normalised pixel threshold (thresh / mean focal)^2 on the squared Sampson error, where the reference calls
      cv2.findEssentialMat(RANSAC) (:155-156; `conf` is OpenCV's early-exit knob: a fixed 2048-hypothesis round here)."""
    if len(kpts0) < 5:
        return None, 0, 0, 0
    if len(kpts0) < _MIN_MODEL_POINTS:
        return None, 0, 0, 0
    K0t, K1t = torch.as_tensor(K0).float(), torch.as_tensor(K1).float()
    if solver == 'prior_ransac' and priorRT is not None:
        r = _ransac_one_pair(kpts0, kpts1, K0t, K1t, priorRT, 3e-7, True)
    elif solver == 'prior_ransac_noprior':
        r = _ransac_one_pair(kpts0, kpts1, K0t, K1t, None, 3e-7, False)
    else:
        f = float(np.mean([float(K0t[0, 0]), float(K1t[1, 1]), float(K0t[0, 0]), float(K1t[1, 1])]))
        r = _ransac_one_pair(kpts0, kpts1, K0t, K1t, None, (float(thresh) / f) ** 2, False)
    best = int(r['best'][0])            # the reference's sync point (E.cpu().numpy(), :118)
    if best < 0:
        return None, 0, 0, 0
    bits = r['mask'].cpu().numpy()
    mask = (bits & 1) > 0
    c3 = r['counts3'][0].tolist()
    Rt = r['Rt'][0].double()
    t = Rt[:, 3]
    if translation_scale is not None:
        t = t * torch.as_tensor(translation_scale).to(t)
    ret = (Rt[:, :3].contiguous(), t.contiguous(), mask, r['E'][0].double().cpu())
    explicit = solver in ('prior_ransac', 'prior_ransac_noprior') and not (solver == 'prior_ransac' and priorRT is None)
    return ret, torch.tensor(int(mask.sum())), (c3[1] if explicit else 0), (c3[2] if explicit else 0)


class EssentialMatrixSolver:
    """Drop-in for mapfree_6dreg/lib/models/matching/pose_solver.py:20-97 (numpy in / numpy out, one pair)."""

    def __init__(self, cfg, use_prior_ransac=False):
        em = cfg['EMAT_RANSAC'] if isinstance(cfg, dict) else cfg.EMAT_RANSAC
        get = (lambda k: em[k]) if isinstance(em, dict) else (lambda k: getattr(em, k))
        self.ransac_pix_threshold = get('PIX_THRESHOLD')
        self.ransac_confidence = get('CONFIDENCE')
        self.use_prior_ransac = use_prior_ransac

    def estimate_pose(self, kpts0, kpts1, data, priorRT=None):
        R, t = np.eye(3), np.zeros((3))
        if len(kpts0) < 5 or len(kpts0) < _MIN_MODEL_POINTS:   # :33-34 (and the `E is None` exit :86-87 for 5..7)
            return (R, t, 0), 0, 0
        K0 = torch.as_tensor(data['K_color0']).reshape(3, 3).float()
        K1 = torch.as_tensor(data['K_color1']).reshape(3, 3).float()
        k0, k1 = torch.as_tensor(np.asarray(kpts0) if not torch.is_tensor(kpts0) else kpts0), \
            torch.as_tensor(np.asarray(kpts1) if not torch.is_tensor(kpts1) else kpts1)
        prior = self.use_prior_ransac and priorRT is not None
        if prior:
            r = _ransac_one_pair(k0, k1, K0, K1, priorRT, 3e-7, True)
        else:  # cv.findEssentialMat(USAC_MAGSAC, threshold = pix / mean focal) in the reference (:81-83)
            f = float(np.mean([float(K0[0, 0]), float(K1[1, 1]), float(K0[1, 1]), float(K1[0, 0])]))
            r = _ransac_one_pair(k0, k1, K0, K1, None, (self.ransac_pix_threshold / f) ** 2, False)
        if int(r['best'][0]) < 0:
            return (R, t, 0), 0, 0
        self.mask = ((r['mask'].cpu().numpy() & 1) > 0).astype(np.uint8)[:, None]
        c3 = r['counts3'][0].tolist()
        Rt = r['Rt'][0].double().cpu().numpy()
        n = int(r['n_pos'][0])                 # recoverPose's return value: cheirality-consistent inliers (:92-95)
        if n <= 0:
            return (R, t, 0), (c3[1] if prior else 0), (c3[2] if prior else 0)
        return (Rt[:, :3], Rt[:, 3], n), (c3[1] if prior else 0), (c3[2] if prior else 0)


class RANSAC(torch.nn.Module):
    """Drop-in for third_party/prior_ransac/ransac.py:74-442 for the essential-matrix model types the FAR recipes
    construct (`essential_cv2` at metrics.py:114 / pose_solver.py:60, `essential`, `fundamental`): forward(kp1, kp2) on
    K-normalised keypoints [N,2] -> (E [3,3], inliers [N] bool, inliers_tight [N] bool, inliers_ultra_tight [N] bool).
    `max_iter` rounds of `batch_size` hypotheses; a round replaces the running best only when its score is higher
    (:419-437).  Local optimisation, early stopping and the homography models are outside the FAR path."""

    def __init__(self, model_type='homography', inl_th=2.0, batch_size=2048, max_iter=10, confidence=0.99,
                 max_lo_iters=5, prior_params={}, use_noexp_prior_scoring=False, use_linear_bias_sampling=False,
                 bias_sigma_sq=1.0, compute_stopping_inlier_only=False, perform_early_stopping=False, l1_dist=False,
                 use_epipolar_error=False, K=None, normalize=False, seed=0, minimal_solver=None):
        super().__init__()
        if model_type not in ('essential_cv2', 'essential', 'fundamental'):
            raise NotImplementedError(f"{model_type}: only the essential/fundamental models are on the FAR path")
        if max_lo_iters or perform_early_stopping or l1_dist or use_epipolar_error:
            raise NotImplementedError("local optimisation / early stopping / L1 / epipolar-error variants are unused by "
                                      "the FAR recipes (metrics.py:114-123, pose_solver.py:60-70)")
        if prior_params and not (use_noexp_prior_scoring and use_linear_bias_sampling):
            raise NotImplementedError("prior scoring is implemented for use_noexp_prior_scoring + "
                                      "use_linear_bias_sampling (the shipped recipe)")
        self.model_type, self.inl_th, self.batch_size, self.max_iter = model_type, inl_th, batch_size, max_iter
        self.bias_sigma_sq, self.prior_params, self.seed = bias_sigma_sq, prior_params, seed
        self.use_prior = bool(prior_params)
        self.prior_lambda = prior_params['lambda'] if self.use_prior else 1.0
        # minimal solver of a hypothesis: the in-repo normalised 8-point on 8 draws (default for every model type: fastest),
        # or '5pt' = the sample sizes of the reference's essential models (ransac.py:146-157: `essential` 5 draws,
        # `essential_cv2` 6): Nister's 5-point on 5 + 1 draws (csrc/fivept.cuh).  FAR_RANSAC_MINIMAL=5pt selects it globally.
        import os
        self.minimal_solver = minimal_solver or os.environ.get("FAR_RANSAC_MINIMAL", "8pt")
        if self.minimal_solver not in ("8pt", "5pt") or (self.minimal_solver == "5pt" and model_type == "fundamental"):
            raise NotImplementedError(f"minimal_solver={self.minimal_solver!r} for model_type={model_type!r}")
        self.minimal_sample_size = 8 if self.minimal_solver == "8pt" else 6

    def forward(self, kp1, kp2, weights=None):
        from .ransac import ransac_round
        dev = kp1.device
        n = kp1.shape[0]
        eyeK = torch.eye(3, device=dev).reshape(1, 3, 3)
        bids = torch.zeros(n, dtype=torch.int64, device=dev)
        prior = pcl = None
        if self.use_prior:
            prior = torch.as_tensor(self.prior_params['RT'], dtype=torch.float32, device=dev)[:3].reshape(1, 3, 4)
            pcl = torch.as_tensor(self.prior_params['pcl'], dtype=torch.float32, device=dev)
        best_score, out = float(self.minimal_sample_size), None
        for i in range(self.max_iter):
            biased = self.use_prior and i % 2 == 0 and bool(self.prior_params.get('biased_sampling'))   # :375-378
            r = ransac_round(kp1.float(), kp2.float(), bids, eyeK, eyeK, prior, batch_size=self.batch_size,
                             inl_th=self.inl_th, prior_lambda=self.prior_lambda, bias_sigma_sq=self.bias_sigma_sq,
                             biased=biased, pcl=pcl, seed=self.seed + i, minimal_solver=self.minimal_solver)
            b = int(r['best'][0])
            if b >= 0 and float(r['scores'][0, b]) > best_score:
                best_score, out = float(r['scores'][0, b]), r
        if out is None:   # :346-348 initial values
            z = torch.zeros(n, dtype=torch.bool, device=dev)
            return torch.zeros(3, 3, device=dev), z.reshape(n, 1), z, z
        m = out['mask']
        return out['E'][0], (m & 1) > 0, (m & 2) > 0, (m & 4) > 0
