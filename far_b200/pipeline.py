"""The FAR-LoFTR per-pair pose pipeline, batched: what PL_LoFTR.test_step does for one pair
(mp3d_loftr/src/lightning/lightning_loftr.py:325-421), run as N independent B=1 evaluations in one pass.

    LoFTR.forward                       backbone (cuDNN) -> coarse transformer -> dual-softmax matching -> fine level
    solver call 1                       RANSAC round on the GPU, uniform sampling, no prior (the reference's
                                        `prior_ransac_noprior` branch, metrics.py:124-143; OpenCV RANSAC on the CPU in a
                                        python loop with 3 PCIe hops in the shipped recipe)
    for i in range(fine_pred_steps):    FAR head: LoFTR regress layers -> EMM bilinear attention -> gated MLP fusion
        forward_rt_prediction
        solver call 2 (i == 0)          prior-guided RANSAC round with the head's pose as the prior
                                        (lightning_loftr.py:343-344, metrics.py:100-123)

The counters the head's gate consumes (`num_correspondences_after_ransac`, `inliers_best_tight`,
`inliers_best_ultra_tight`) are RANSAC inlier counts in both solver calls, the quantity the gate was trained on.
`first_solver='weighted_8pt'` selects SURVEY.md 8d config 2's literal solver instead (one mconf-weighted 8-point fit,
no outlier rejection: its `after_ransac` counter is the cheirality vote, NOT an inlier count -- for solver benchmarks,
not for real checkpoints).
"""
import torch

from .solver import estimate_pose_batched, estimate_pose_ransac_batched
from .loftr.pose import pose_mean_6d, pose_std_6d, rotation_6d_to_matrix


class FarPosePipeline:
    def __init__(self, model, K0, K1, fine_pred_steps=None, prior_ransac=True, ransac_kwargs=None,
                 first_solver='ransac', overlap=True, graph=False):
        self.model = model
        self.overlap = overlap
        # graph=True: segment 1 replays from a CUDA graph (one per input shape).  Opt-in: its outputs (feature maps, the
        # `data` entries it writes) are static buffers reused by the next call.
        self.graph = graph
        self.graph_error = None
        self._graphs = {}
        self.K0, self.K1 = K0, K1
        self.steps = fine_pred_steps if fine_pred_steps is not None else model.config.get('fine_pred_steps', 1)
        # prior_ransac=True: the solver call between the two head invocations is the prior-guided RANSAC round of the
        # recipe of record run on the GPU (far_b200/ransac.py) with the first head prediction as the prior;
        # False: the first solver is simply re-run.
        self.prior_ransac = prior_ransac
        self.first_solver = first_solver
        self.ransac_kwargs = ransac_kwargs or {}
        if 'regress' in model.config:
            model.config['regress']['prior_rt_on_device'] = True   # no blocking D2H of the prior (loftr.py:187-192)

    def _solve(self, data, K0, K1, prior=None):
        if self.first_solver == 'weighted_8pt' and prior is None:
            return estimate_pose_batched(data, K0, K1)
        return estimate_pose_ransac_batched(data, K0, K1, prior, **self.ransac_kwargs)

    # ---- segment 1: everything before the host needs the match count --------------------------------------------------
    def _segment1(self, data, defer_readback=False):
        m = self.model
        m.forward_feature_extraction(data)
        handle = m.forward_coarse(data, defer_readback=defer_readback)
        # The head trunk needs only the coarse features: queued BEFORE the host waits for the match count, so the GPU
        # stays busy across the reference's torch.where sync and while the host issues the many small fine-level /
        # solver launches.  Same ops, same results, different issue order.
        if self.overlap and m.config.get('regress_rt') and m.config['regress'].get('reuse_trunk', True):
            m.head_trunk(data)
        return handle

    def _segment1_graphed(self, image0, image1):
        """Segment 1 (backbone -> coarse transformer -> score / decision kernels -> head trunk: static shapes, ~85 % of
        the step's launches) as one CUDA graph per input shape: captured once after two eager warm-up runs (cuDNN
        autotuning, weight-split / table caches), replayed afterwards.  The tensors it produces live in the graph's
        memory pool and are OVERWRITTEN by the next call with the same shape."""
        key = (tuple(image0.shape), tuple(image1.shape), image0.device.index)
        ent = self._graphs.get(key)
        if ent is None:
            s0, s1 = torch.empty_like(image0), torch.empty_like(image1)
            s0.copy_(image0)
            s1.copy_(image1)
            side = torch.cuda.Stream(device=image0.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    h = self._segment1({'image0': s0, 'image1': s1}, defer_readback=True)
                    del h
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            data = {'image0': s0, 'image1': s1}
            try:
                with torch.cuda.graph(g):
                    handle = self._segment1(data, defer_readback=True)
            except Exception as ex:   # an op that cannot be captured: stay eager, loudly
                import warnings
                warnings.warn(f"FarPosePipeline: CUDA-graph capture of segment 1 failed ({ex!r}); running eagerly")
                self.graph, self.graph_error = False, repr(ex)
                data = {'image0': image0, 'image1': image1}
                return data, self._segment1(data)
            ent = (g, s0, s1, data, handle)
            self._graphs[key] = ent
        else:
            g, s0, s1, data, handle = ent
            s0.copy_(image0)
            s1.copy_(image1)
        g.replay()                # (the capture itself does not execute anything)
        # the handle object is reused by every replay: start THIS call's read-back of the match count now
        from . import ops
        ops.match_count_readback(handle)
        return dict(data), handle

    @torch.no_grad()
    def __call__(self, image0, image1):
        """image0/1: [N,1,H,W] fp32 CUDA in [0,1].  Returns dict: pose [N,3,4] fused (R|t), regressed_rt [N,9],
        loftr_rt [N,3,4], num_matches [N] (before RANSAC), num_inliers [N] (after RANSAC), gating [N,2]."""
        from . import ops, _lib
        m = self.model
        if self.graph and ops.timer is None and not _lib.profiling_enabled():
            data, handle = self._segment1_graphed(image0, image1)
        else:
            data = {'image0': image0, 'image1': image1}
            handle = self._segment1(data)
        m.forward_fine(data, handle)
        N = image0.shape[0]
        K0 = self.K0[:N] if self.K0.shape[0] >= N else self.K0.expand(N, 3, 3)
        K1 = self.K1[:N] if self.K1.shape[0] >= N else self.K1.expand(N, 3, 3)
        self._solve(data, K0, K1)
        if m.config.get('regress_rt'):
            for i in range(self.steps):
                m.forward_rt_prediction(data)
                if i == 0 and self.steps > 1:
                    self._solve(data, K0, K1, data['priorRT_device'] if self.prior_ransac else None)
            rr = data['regressed_rt']
            dev = rr.device
            R = rotation_6d_to_matrix(rr[:, 3:] * pose_std_6d[3:].to(dev) + pose_mean_6d[3:].to(dev))
            t = rr[:, :3] * pose_std_6d[:3].to(dev) + pose_mean_6d[:3].to(dev)
            pose = torch.cat([R, t.unsqueeze(-1)], dim=-1)
        else:
            pose = data['loftr_rt']
        return {'pose': pose, 'regressed_rt': data.get('regressed_rt'), 'loftr_rt': data['loftr_rt'],
                'num_matches': data['num_correspondences_before_ransac'],
                'num_inliers': data['num_correspondences_after_ransac'],
                'gating': data.get('gating_reg_weights'), 'data': data}


def shard_pairs(n_pairs, world_size, rank):
    """Contiguous split of a pair batch over ranks: returns (start, stop)."""
    base, rem = divmod(n_pairs, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_poses(pose, counts, group=None):
    """The path's only collective: all-gather of [pairs_per_rank, 12] poses + [pairs_per_rank] match counts
    (replaces the reference's pickled gloo gather, mp3d_loftr/src/utils/comm.py:179-219).  Ragged shards are padded
    to the largest shard."""
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    n = torch.tensor([pose.shape[0]], device=pose.device, dtype=torch.int64)
    ns = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(ns, n, group=group)
    nmax = int(max(x.item() for x in ns))
    buf = torch.zeros((nmax, 13), device=pose.device, dtype=torch.float32)
    buf[:pose.shape[0], :12] = pose.reshape(-1, 12)
    buf[:pose.shape[0], 12] = counts.float()
    out = [torch.zeros_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf, group=group)
    full = torch.cat([o[:int(k.item())] for o, k in zip(out, ns)], 0)
    return full[:, :12].reshape(-1, 3, 4), full[:, 12].round().to(torch.int64)
