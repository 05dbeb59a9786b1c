#!/usr/bin/env python
"""Golden vectors for the map-free path (BASELINE configs[3], SURVEY.md 8b `LoFTR (upstream)` / `RegressionModel`),
produced by the UNMODIFIED reference imported through oracle/ref_import.py:

  upstream_loftr.npz   pristine LoFTR (mapfree_6dreg/etc/feature_matching_baselines/LoFTR/src/loftr/loftr.py:29-75,
                       default_cfg: 4 x (self, cross), temp_bug_fix False) on one 720 x 544 pair -> 90 x 68 = 6120
                       coarse tokens (not a multiple of the 128-row tiles), thr = 0; also pins the oracle there.
  mapfree_model.npz    RegressionModel.forward(data) (mapfree_6dreg/lib/models/regression/model.py:235-308) with
                       use_loftr_preds + use_vanilla_transformer + use_prior, B = 2: every stage of the image branch
                       (ResUNet, aggregator, head, TransformerEncoder) and the two solver / fusion loops.  The
                       solver of record is OpenCV (un-vendored); `pose_solver.estimate_pose` is replaced by a
                       deterministic stand-in (the oracle's in-repo 8-point + cheirality on the reference matcher's
                       keypoints) whose per-loop outputs are stored, so the GPU test can replay the reference's
                       control flow (prior from loop 0 into loop 1) on identical solver numbers.

Runs only in the build container:   python tests/golden/make_golden_mapfree_model.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import far_oracle as O  # noqa: E402
from oracle import ref_import as R  # noqa: E402
from far_b200 import synth  # noqa: E402
from far_b200.loftr import upstream_loftr_cfg  # noqa: E402

SEED = 4321


def npz(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrs.items()})
    print(f"  wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def mapfree_images(B, seed):
    return synth.synth_mapfree_images(B, seed)


def main():
    ns = R.load_mapfree()
    # ------------------------------------------------------------------ upstream LoFTR @ 720 x 544
    print("[upstream_loftr] pristine LoFTR.forward, 1 pair 720x544 (6120 coarse tokens, 8 coarse layers)")
    img0, img1, _, _ = mapfree_images(1, 20240003)
    out = {"seed": np.int64(SEED), "image_seed": np.int64(20240003)}
    # default_cfg has temp_bug_fix False: div_term = exp(-k) makes half the positional channels constant, and with
    # random-init weights the dual-softmax is nearly flat (conf ~ 4e-6, ~19 mutual-NN matches).  The same unmodified
    # class with temp_bug_fix True ("tb" keys) yields ~1.2k matches on the same 6120-token grid.
    for tag, tb in (("", False), ("tb_", True)):
        cfg = {k: (dict(v) if isinstance(v, dict) else v) for k, v in ns.upstream_default_cfg.items()}
        cfg["match_coarse"]["thr"] = 0.0
        cfg["coarse"]["temp_bug_fix"] = tb
        ref = ns.UpstreamLoFTR(config=cfg).eval()
        sd = synth.synth_state_dict(ref.state_dict(), SEED)
        print("  load_state_dict:", ref.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True))
        data = {"image0": img0, "image1": img1}
        with torch.no_grad():
            ref(data)
        M = data["b_ids"].shape[0]
        print(f"  temp_bug_fix={tb}: reference matches: {M}; hw0_c {tuple(data['hw0_c'])}; conf max {data['conf_matrix'].max():.3e}")
        ocfg = upstream_loftr_cfg()
        ocfg["match_coarse"]["thr"] = 0.0
        ocfg["coarse"]["temp_bug_fix"] = tb
        with torch.no_grad():
            od = O.loftr_forward(sd, img0, img1, ocfg)
        assert torch.equal(od["i_ids"], data["i_ids"]) and torch.equal(od["j_ids"], data["j_ids"]), "oracle match indices"
        for k, tol in (("mconf", 1e-6), ("mkpts1_f", 1e-4), ("expec_f", 1e-5)):
            err = (od[k] - data[k]).abs().max().item()
            print(f"  oracle vs reference {k}: max|diff| {err:.3e}")
            assert err <= tol, k
        conf = data["conf_matrix"][0]
        rmax, rarg = conf.max(dim=1)
        out.update({tag + "i_ids": data["i_ids"].to(torch.int32).numpy(), tag + "j_ids": data["j_ids"].to(torch.int32).numpy(),
                    tag + "mconf": data["mconf"].numpy(), tag + "mkpts1_f": data["mkpts1_f"].numpy(),
                    tag + "expec_f": data["expec_f"].numpy(), tag + "conf_rowmax": rmax.numpy(),
                    tag + "conf_rowarg": rarg.to(torch.int32).numpy(), tag + "conf_colmax": conf.max(dim=0)[0].numpy(),
                    tag + "feat_c0_sample": od["featmap0"][0, ::97, ::8].numpy()})
    npz("upstream_loftr.npz", **out)

    # ------------------------------------------------------------------ RegressionModel.forward
    print("[mapfree_model] RegressionModel.forward, B = 2, use_loftr_preds + use_vanilla_transformer + use_prior")
    torch.manual_seed(0)
    m = ns.build(use_loftr_preds=True, use_vanilla_transformer=True, use_prior=True, inference=True)
    ns.restore()
    m.matcher.config["match_coarse"]["thr"] = 0.0
    m.matcher.coarse_matching.thr = 0.0
    msd = synth.synth_state_dict(m.state_dict(), SEED)
    print("  load_state_dict:", m.load_state_dict({k: v.clone() for k, v in msd.items()}, strict=True))
    B = 2
    i0, i1, r0, r1 = mapfree_images(B, 20240004)
    K = synth.mapfree_intrinsics(B)
    log = {"loftr_rt": [], "n": [], "prior": []}

    def stub_estimate_pose(kpts0, kpts1, data2, priorRT=None):
        # deterministic stand-in for cv.findEssentialMat + cv.recoverPose (pose_solver.py:30-97)
        k0, k1 = torch.from_numpy(np.asarray(kpts0)).float(), torch.from_numpy(np.asarray(kpts1)).float()
        Kc = data2["K_color0"][0].float()
        Rm, tv, _ = O.pose_from_matches_8pt(k0, k1, torch.ones(k0.shape[0]), Kc, Kc)
        n = int(k0.shape[0])
        if priorRT is not None:   # make the second loop's numbers depend on the prior, like a prior-guided solver
            Rm = torch.as_tensor(priorRT)[:3, :3].float()
            log["prior"].append(torch.as_tensor(priorRT).float().clone())
            return (Rm.numpy(), tv.numpy(), n), n // 3, n // 7
        return (Rm.numpy(), tv.numpy(), n), 0, 0

    m.pose_solver.estimate_pose = stub_estimate_pose
    stages = {}
    hooks = [m.encoder.register_forward_hook(lambda mod, a, out: stages.setdefault("vol", []).append(out.detach().clone())),
             m.aggregator.register_forward_hook(lambda mod, a, out: stages.__setitem__("agg", out.detach().clone())),
             m.head.register_forward_hook(lambda mod, a, out: stages.__setitem__("head", out[2].detach().clone())),
             m.transformer.register_forward_hook(lambda mod, a, out: stages.__setitem__("tr", out.detach().clone()))]
    data = {"image0": i0, "image1": i1, "image0_reg": r0, "image1_reg": r1, "K_color0": K, "K_color1": K}
    orig_mlp = m.regression_mlp

    def logged_mlp(feats, loftr_rt, inliers, R=None, t=None):
        log["loftr_rt"].append(loftr_rt.clone())
        log["n"].append(inliers.clone().float())
        return orig_mlp(feats, loftr_rt, inliers, R, t)

    m.regression_mlp = logged_mlp
    with torch.no_grad():
        Rout, tout = m(data)
    for h in hooks:
        h.remove()
    print(f"  R {tuple(Rout.shape)} t {tuple(tout.shape)}; vol {tuple(stages['vol'][0].shape)} agg {tuple(stages['agg'].shape)} "
          f"head {tuple(stages['head'].shape)} transformer {tuple(stages['tr'].shape)}")
    # pin the oracle pieces that exist for this path against the reference stages
    agg_o = O.mapfree_correlation_aggregator(stages["vol"][0], stages["vol"][1])
    print("  oracle aggregator vs reference:", (agg_o - stages["agg"]).abs().max().item())
    assert (agg_o - stages["agg"]).abs().max() < 1e-5
    tr_in = stages["head"].reshape(B, 256, 108).permute(2, 0, 1)
    tr_o = O.torch_transformer_encoder(O._sub(msd, "transformer"), tr_in)
    print("  oracle transformer vs reference:", (tr_o - stages["tr"]).abs().max().item())
    assert (tr_o - stages["tr"]).abs().max() < 2e-5
    feats = stages["tr"].permute(1, 2, 0)                                   # [B, 256, 108]
    Ro, to, _ = O.mapfree_regression_mlp(msd, feats.reshape(B, -1), log["loftr_rt"][1], log["n"][1])
    print("  oracle regression_mlp vs reference:", (Ro - Rout).abs().max().item(), (to - tout).abs().max().item())
    assert (Ro - Rout).abs().max() < 1e-5 and (to - tout).abs().max() < 1e-5
    npz("mapfree_model.npz", seed=np.int64(SEED), image_seed=np.int64(20240004), K=K,
        vol0=stages["vol"][0][:, :, ::5, ::3], vol1=stages["vol"][1][:, :, ::5, ::3], agg=stages["agg"][:, :, ::5, ::3],
        head=stages["head"][:, ::3], transformer=feats[:, ::3],
        loftr_rt0=log["loftr_rt"][0], loftr_rt1=log["loftr_rt"][1], inliers0=log["n"][0], inliers1=log["n"][1],
        prior=torch.stack(log["prior"]), R=Rout, t=tout)


if __name__ == "__main__":
    main()
