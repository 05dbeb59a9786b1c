"""Generate the committed golden fixtures by running the UNMODIFIED reference (/root/reference, imported on CPU
through oracle/ref_import.py) on deterministic synthetic weights/inputs, and pin the restatement in
oracle/far_oracle.py against it in the same run.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Outputs: tests/golden/*.npz (small; large tensors are stored as strided samples).  The GPU box never needs the
reference: tests re-create the same weights/inputs from far_b200.synth seeds and compare against these files.
"""
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

torch.manual_seed(0)
SEED = 1234


def npz(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrs.items()})
    print(f"  wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def close(a, b, tol, what):
    a, b = a.double(), b.double()
    err = (a - b).abs().max().item() if a.numel() else 0.0
    scale = max(b.abs().max().item(), 1e-30) if b.numel() else 1.0
    print(f"  {what}: max|diff| {err:.3e} (ref scale {scale:.3e})")
    assert err <= tol * max(scale, 1.0), what


def gen_loftr_full():
    """One 640x480 pair through the reference LoFTR.forward (thr=0) + FAR head, vs the oracle."""
    print("[loftr_full] reference LoFTR.forward + forward_rt_prediction, 1 pair 640x480")
    ns = R.load_mp3d()
    cfg = R.mp3d_eval_config(thr=0.0)
    ref = ns.LoFTR(cfg).eval()
    sd = synth.synth_state_dict(ref.state_dict(), SEED)
    missing = ref.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    print("  load_state_dict:", missing)
    img0, img1 = synth.synth_pair_images(1, seed=20240001)
    data = {"image0": img0, "image1": img1}
    with torch.no_grad():
        ref(data)
    M = data["b_ids"].shape[0]
    print(f"  reference matches: {M}; conf max {data['conf_matrix'].max():.4f}")
    with torch.no_grad():
        od = O.loftr_forward(sd, img0, img1, cfg)
    # ---- pin the oracle against the reference
    assert torch.equal(od["b_ids"], data["b_ids"]) and torch.equal(od["i_ids"], data["i_ids"]) and \
        torch.equal(od["j_ids"], data["j_ids"]), "oracle match indices differ from the reference"
    close(od["featmap0"], data["featmap0"], 1e-5, "featmap0 (post coarse transformer)")
    close(od["mconf"], data["mconf"], 1e-5, "mconf")
    close(od["expec_f"], data["expec_f"], 1e-4, "expec_f")
    close(od["mkpts1_f"], data["mkpts1_f"], 1e-5, "mkpts1_f")
    close(od["conf_matrix"], data["conf_matrix"], 1e-5, "conf_matrix")

    # ---- FAR head on a synthetic solver pose (the reference solver of record is OpenCV RANSAC, not restated)
    g = np.random.default_rng(77)
    _, _, _, Rg, tg = synth.two_view_geometry(1, 16, seed=77)
    loftr_rt = torch.cat([Rg[0], tg[0][:, None]], dim=1).double()  # estimate_pose returns float64 (metrics.py:165-168)
    data.update({"loftr_rt": loftr_rt, "num_correspondences": torch.tensor([412]),
                 "num_correspondences_before_ransac": torch.tensor([M]),
                 "inliers_best_tight": torch.tensor([301]), "inliers_best_ultra_tight": torch.tensor([57])})
    with torch.no_grad():
        ref.forward_rt_prediction(data)
        lp, ilp = O.preprocess_helper(loftr_rt, 412, M, 301, 57)
        pose_o, wt_o = O.far_head_mp3d(O._sub(sd, "loftr_regress"), data["featmap0"], data["featmap1"], lp, ilp, cfg)
    close(pose_o, data["regressed_rt"], 1e-4, "regressed_rt (FAR head)")
    close(wt_o, data["gating_reg_weights"], 1e-4, "gating weights")
    close(O.prior_rt_from_regressed(pose_o), torch.from_numpy(data["priorRT"]), 1e-4, "priorRT")

    s = slice(None, None, 53)
    npz("loftr_full.npz", b_ids=data["b_ids"].int(), i_ids=data["i_ids"].int(), j_ids=data["j_ids"].int(),
        mconf=data["mconf"], mkpts0_c=data["mkpts0_c"], mkpts1_c=data["mkpts1_c"], mkpts1_f=data["mkpts1_f"],
        expec_f=data["expec_f"], featmap0_s=data["featmap0"][0, s, ::7], featmap1_s=data["featmap1"][0, s, ::7],
        conf_rows_s=data["conf_matrix"][0, ::601, ::11], loftr_rt=loftr_rt, counters=np.array([412, M, 301, 57]),
        regressed_rt=data["regressed_rt"], gating=data["gating_reg_weights"], priorRT=data["priorRT"],
        seed=np.array([SEED, 20240001]))


def gen_stages():
    """Stage-level goldens at small sizes (seconds on CPU): encoder layer, matching on a 12x16 grid with border,
    fine preprocess + fine matching, from the reference modules themselves."""
    print("[stages] reference modules at small sizes")
    ns = R.load_mp3d()
    cfg = R.mp3d_eval_config(thr=0.0)
    g = O.rng(11)
    # --- LoFTREncoderLayer / LocalFeatureTransformer, C=256 h=8, L=300 S=280
    lft = ns.LocalFeatureTransformer({**cfg["coarse"], "layer_names": ["self", "cross"]}).eval()
    sd = synth.synth_state_dict(lft.state_dict(), SEED)
    lft.load_state_dict(sd)
    f0, f1 = O.randn(g, 2, 300, 256), O.randn(g, 2, 280, 256)
    with torch.no_grad():
        r0, r1 = lft(f0, f1)
        o0, o1 = O.local_feature_transformer(sd, f0, f1, ["self", "cross"], 8)
    close(o0, r0, 1e-5, "LocalFeatureTransformer feat0")
    close(o1, r1, 1e-5, "LocalFeatureTransformer feat1")
    # --- CoarseMatching on a 12x16 / 10x14 grid, border 2, thr 0
    cm = ns.CoarseMatching({**cfg["match_coarse"], "thr": 0.0}).eval()
    c0, c1 = O.randn(g, 3, 12 * 16, 256, scale=4.0), O.randn(g, 3, 10 * 14, 256, scale=4.0)
    d = {"hw0_i": (96, 128), "hw1_i": (80, 112), "hw0_c": (12, 16), "hw1_c": (10, 14)}
    with torch.no_grad():
        cm(c0, c1, d)
        oc = O.coarse_matching(c0, c1, (12, 16), (10, 14), 0.0, 2, 0.1, 8.0)
    for k in ("b_ids", "i_ids", "j_ids"):
        assert torch.equal(oc[k], d[k]), k
    close(oc["mconf"], d["mconf"], 1e-5, "coarse mconf")
    print(f"  coarse matches on the small grid: {d['b_ids'].numel()}")
    # --- FinePreprocess + fine transformer + FineMatching on a 48x64 fine map (coarse 12x16, stride 4)
    fp = ns.FinePreprocess(cfg).eval()
    sdf = synth.synth_state_dict(fp.state_dict(), SEED)
    fp.load_state_dict(sdf)
    ff0, ff1 = O.randn(g, 3, 128, 48, 64), O.randn(g, 3, 128, 40, 56)
    d.update({"hw0_f": (48, 64), "hw1_f": (40, 56)})
    with torch.no_grad():
        u0, u1 = fp(ff0, ff1, c0, c1, d)
        # the reference's FinePreprocess needs equal-size maps only for torch.cat of the two unfolded tensors' rows
        q0, q1 = O.fine_preprocess(sdf, ff0, ff1, c0, c1, d["b_ids"], d["i_ids"], d["j_ids"], 5, 4)
    close(q0, u0, 1e-5, "fine_preprocess feat0")
    close(q1, u1, 1e-5, "fine_preprocess feat1")
    fm = ns.FineMatching(cfg).eval()
    with torch.no_grad():
        fm(u0, u1, d)
        e, m0, m1 = O.fine_matching(u0, u1, d["mkpts0_c"], d["mkpts1_c"], 2.0)
    close(e, d["expec_f"], 1e-5, "expec_f")
    close(m1, d["mkpts1_f"], 1e-5, "mkpts1_f")
    npz("stages.npz", lft_out0=r0[:, ::3, ::5], lft_out1=r1[:, ::3, ::5], cm_b=d["b_ids"].int(), cm_i=d["i_ids"].int(), cm_j=d["j_ids"].int(),
        cm_mconf=d["mconf"], fp_out0=u0[::5, ::3, ::9], fp_out1=u1[::5, ::3, ::9], expec_f=d["expec_f"],
        mkpts1_f=d["mkpts1_f"], seed=np.array([SEED, 11]))


def gen_solver():
    """run_8point / decompose_essential_matrix from third_party/prior_ransac on synthetic two-view geometry."""
    print("[solver] reference run_8point + decompose_essential_matrix, P=16 N=256")
    pr = R.load_prior_ransac()
    p1, p2, w, Rg, tg = synth.two_view_geometry(16, 256, seed=5)
    with torch.no_grad():
        F_w = pr.run_8point(p1, p2, w)
        F_u = pr.run_8point(p1, p2, None)
        R1, R2, t = pr.decompose_essential_matrix(F_w)
        Fo_w = O.run_8point(p1, p2, w, dense_diag=True)
        Fo_u = O.run_8point(p1, p2, None)
        o1, o2, ot = O.decompose_essential_matrix(F_w)
    close(Fo_w, F_w, 1e-5, "run_8point weighted (oracle dense_diag)")
    close(O.run_8point(p1, p2, w), F_w, 1e-3, "run_8point weighted (oracle (w*X)^T X form)")
    close(Fo_u, F_u, 1e-5, "run_8point unweighted")
    close(o1, R1, 1e-5, "decompose R1")
    close(o2, R2, 1e-5, "decompose R2")
    close(ot, t, 1e-5, "decompose t")
    npz("solver.npz", F_w=F_w, F_u=F_u, R1=R1, R2=R2, t=t, R_gt=Rg, t_gt=tg, seed=np.array([5, 16, 256]))


def gen_vit8pt():
    """Reference ViTEss.forward (interiornetStreetlearn_8ptVit/src/model.py:165-217), B=2, 640x480 BGR 0-255."""
    print("[vit8pt] reference ViTEss.forward, B=2")
    ns = R.load_vit8pt()
    mean = torch.tensor([0.1, -0.05, 0.2, 0.8, 0.02, -0.1, -0.02, 0.9, 0.05])
    std = torch.tensor([0.9, 0.4, 0.8, 0.3, 0.2, 0.5, 0.2, 0.1, 0.15])
    ref = ns.ViTEss(R.vit8pt_args(), mean, std).eval()
    sd = synth.synth_state_dict(ref.state_dict(), SEED)
    ref.load_state_dict(sd, strict=True)
    g = np.random.default_rng(20240003)
    B = 2
    images = torch.from_numpy(g.integers(0, 256, size=(B, 2, 3, 448, 448)).astype(np.float32))  # exact 2x nearest resize
    intr = torch.tensor([[[400., 400., 224., 224.]] * 2, [[517.97, 517.97, 224., 224.]] * 2])
    _, _, _, Rg, tg = synth.two_view_geometry(B, 16, seed=78)
    loftr_preds = torch.cat([Rg, tg[:, :, None]], dim=2).double()
    num_corr = torch.tensor([350, 12])
    with torch.no_grad():
        feats, intr_s = ref.extract_features(images.clone(), intr.clone())
        t, rot, Rm, r6 = ref(images.clone(), intr.clone(), loftr_num_corr=num_corr, loftr_preds=loftr_preds)
        pos = O.emm_positional_encodings_vit(intr_s)
        to, Ro, r6o, wto = O.vit_fusion_head(sd, feats, pos, loftr_preds, num_corr, mean, std)
    close(to, t, 1e-4, "ViTEss t")
    close(Ro, Rm, 1e-4, "ViTEss R")
    close(r6o, r6, 1e-4, "ViTEss r6d")
    npz("vit8pt.npz", t=t, R=Rm, r6d=r6, rot=rot, feats_s=feats[:, ::17, ::5], intr_scaled=intr_s, mean=mean, std=std,
        loftr_preds=loftr_preds, num_corr=num_corr, intr=intr, seed=np.array([SEED, 20240003, 78]))


def gen_mapfree_mlp():
    """RegressionModel.regression_mlp (mapfree_6dreg/lib/models/regression/model.py:198-233): the method's own source
    lines are executed on a stand-in `self` (building the whole model needs datasets/checkpoints that are not here)."""
    print("[mapfree] reference RegressionModel.regression_mlp source, B=3")
    import re
    import types
    src = open(os.path.join(R.REF_ROOT, "mapfree_6dreg/lib/models/regression/model.py")).read()
    m = re.search(r"    def regression_mlp\(self.*?\n(?=    def )", src, flags=re.S)
    body = "\n".join(line[4:] for line in m.group(0).splitlines())
    loss_src = open(os.path.join(R.REF_ROOT, "mapfree_6dreg/lib/utils/loss.py")).read()
    c6 = re.search(r"def compute_6d\(r\):.*?\n(?=\n)", loss_src, flags=re.S).group(0)
    R.install_shims()
    nsd = {"torch": torch}
    exec(c6, nsd)
    exec(body, nsd)
    import torch.nn as nn
    H, H2 = 256 * 12 * 9, 512
    fake = types.SimpleNamespace(
        use_prior=True, use_vanilla_transformer=True, num_corr_size=3,
        pose_regressor=nn.Sequential(nn.Linear(H, H2), nn.ReLU(), nn.Linear(H2, H2), nn.ReLU(), nn.Linear(H2, 9)),
        moe_predictor=nn.Sequential(nn.Linear(H + 18 + 3, H2), nn.ReLU(), nn.Linear(H2, H2), nn.ReLU(),
                                    nn.Linear(H2, 2), nn.Sigmoid()))
    sd = {}
    for name in ("pose_regressor", "moe_predictor"):
        mod = getattr(fake, name)
        part = synth.synth_state_dict({f"{name}.{k}": v for k, v in mod.state_dict().items()}, SEED)
        mod.load_state_dict({k[len(name) + 1:]: v for k, v in part.items()})
        sd.update(part)
    g = O.rng(31)
    B = 3
    feats = O.randn(g, B, 256, 12, 9)
    _, _, _, Rg, tg = synth.two_view_geometry(B, 16, seed=79)
    loftr_rt = torch.cat([Rg, 3.0 * tg[:, :, None]], dim=2)
    inl = torch.tensor([[400., 120., 9.], [30., 2., 0.], [0., 0., 0.]])
    with torch.no_grad():
        Rr, tr = nsd["regression_mlp"](fake, feats, loftr_rt, inl)
        Ro, to, wo = O.mapfree_regression_mlp(sd, feats.reshape(B, -1), loftr_rt, inl)
    close(Ro, Rr, 1e-5, "mapfree R6d")
    close(to, tr, 1e-5, "mapfree t")
    npz("mapfree_mlp.npz", R=Rr, t=tr, loftr_rt=loftr_rt, inliers=inl, seed=np.array([SEED, 31, 79]))


if __name__ == "__main__":
    assert R.have_reference(), "needs /root/reference (build container only)"
    torch.set_num_threads(os.cpu_count() or 1)
    only = sys.argv[1:] or ["solver", "stages", "loftr_full", "vit8pt", "mapfree"]
    if "solver" in only:
        gen_solver()
    if "stages" in only:
        gen_stages()
    if "loftr_full" in only:
        gen_loftr_full()
    if "vit8pt" in only:
        gen_vit8pt()
    if "mapfree" in only:
        gen_mapfree_mlp()
    print("done")
