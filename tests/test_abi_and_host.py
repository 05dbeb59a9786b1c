"""`-m "not gpu"`: the C-ABI library loads and exports every symbol include/far_sm100.h declares (no compute calls
without a GPU), the host-side modules keep the reference's parameter names, and the product path fails loudly
instead of falling back to the CPU."""
import os
import re
import subprocess

import pytest
import torch

import far_b200
from far_b200 import _lib, ops, synth
from far_b200.loftr import LoFTR, far_eval_cfg, upstream_loftr_cfg, get_positional_encodings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "far_sm100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(far_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/far_sm100.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes signature table out of sync with the header"
    assert lib.far_abi_version() == 2
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (far_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_library_is_sm100a_with_no_other_arch():
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump / library not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback():
    x = torch.zeros(4, 8)
    with pytest.raises(_lib.FarError):
        ops.linear(x, torch.zeros(3, 8))
    with pytest.raises(_lib.FarError):
        ops.layernorm(x, torch.ones(8), torch.zeros(8), 1e-5)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "far_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "/root/reference" not in src


def test_state_dict_names_match_reference_contract():
    """Parameter names of SURVEY.md 8(b): checkpoints of the reference must load with strict=True."""
    m = LoFTR(far_eval_cfg())
    sd = m.state_dict()
    must = {
        "backbone.conv1.weight": (128, 1, 7, 7),
        "loftr_coarse.layers.0.q_proj.weight": (256, 256), "loftr_coarse.layers.5.mlp.0.weight": (512, 512),
        "loftr_coarse.layers.5.mlp.2.weight": (256, 512), "loftr_coarse.layers.0.norm1.bias": (256,),
        "fine_preprocess.down_proj.weight": (128, 256), "fine_preprocess.merge_feat.bias": (128,),
        "loftr_fine.layers.1.merge.weight": (128, 128),
        "loftr_regress.encoder.0.weight": (512, 35840), "loftr_regress.moe_predictor.0.weight": (512, 35862),
        "loftr_regress.moe_predictor.4.weight": (2, 512), "loftr_regress.pose_regressor_simple_moe.2.weight": (9, 512),
        "loftr_regress.emm.pos_embed": (1, 4800, 256), "loftr_regress.emm.cross_attn.qkv.weight": (768, 256),
        "loftr_regress.emm.cross_attn.qkv.bias": (768,), "loftr_regress.emm.cross_attn.proj_fundamental.weight": (256, 280),
        "loftr_regress.emm.mlp.fc1.weight": (1024, 256), "loftr_regress.loftr.layers.1.v_proj.weight": (256, 256),
        "loftr_regress.norm.weight": (256,),
    }
    for k, shp in must.items():
        assert k in sd and tuple(sd[k].shape) == shp, (k, sd.get(k, torch.empty(0)).shape)
    assert "pos_encoding.pe" not in sd  # non-persistent buffer (position_encoding.py:35)
    assert sum(p.numel() for p in m.parameters()) == 51092731  # 51.1 M params
    # `matcher.`-prefixed Lightning checkpoints load (loftr.py:207-211)
    m.load_state_dict({"matcher." + k: v for k, v in sd.items()}, strict=True)
    up = LoFTR(upstream_loftr_cfg())
    assert len(up.loftr_coarse.layers) == 8 and not hasattr(up, "loftr_regress")


def test_synth_is_deterministic_and_emm_pos_table():
    a = synth.synth_tensor("loftr_coarse.layers.0.q_proj.weight", (256, 256), 1234)
    b = synth.synth_tensor("loftr_coarse.layers.0.q_proj.weight", (256, 256), 1234)
    assert torch.equal(a, b) and abs(a.std().item() - 1 / 16) < 5e-3
    from oracle import far_oracle as O
    t = get_positional_encodings(1, 4800)[0]
    assert torch.allclose(t, O.emm_positional_encodings_mp3d(), atol=1e-6)
    assert t.shape == (4800, 6) and torch.all(t[:, 5] == 1)


def test_mapfree_regression_model_contract():
    """SURVEY.md 8(b) RegressionModel: constructor keywords of the reference, 540 state-dict keys / 66,249,761
    parameters (the probe count of the reference model incl. the frozen upstream LoFTR), reference key prefixes."""
    from far_b200.mapfree import RegressionModel
    m = RegressionModel(use_loftr_preds=True, use_vanilla_transformer=True, use_prior=True, inference=True)
    sd = m.state_dict()
    assert len(sd) == 540 and sum(p.numel() for p in m.parameters()) == 66_249_761
    for k in ("encoder.firstconv.weight", "encoder.encoder3.2.conv3.weight", "encoder.upconv4.conv1.conv.weight",
              "encoder.outconv.normalize.running_var", "head.resblock1.shortcut.0.weight",
              "transformer.layers.5.self_attn.in_proj_weight", "transformer.layers.0.linear1.weight",
              "matcher.loftr_coarse.layers.7.mlp.2.weight", "pose_regressor.0.weight", "moe_predictor.0.weight"):
        assert k in sd, k
    assert tuple(sd["moe_predictor.0.weight"].shape) == (512, 27648 + 18 + 3)
    assert tuple(sd["head.resblock1.conv1.weight"].shape) == (64, 67, 3, 3)
    assert all(not p.requires_grad for p in m.matcher.parameters())


def test_solver_shims_degenerate_inputs_need_no_gpu():
    """The reference's return-by-value error behaviour (metrics.py:82-84, pose_solver.py:33-34): fewer than 5 keypoints
    -> (None, 0, 0, 0) / identity, decided on the host before any kernel is involved."""
    import numpy as np
    from far_b200 import solver
    k = torch.zeros(4, 2)
    K = torch.eye(3)
    assert solver.estimate_pose(k, k, K, K, 0.5) == (None, 0, 0, 0)
    assert solver.estimate_pose(k[:0], k[:0], K, K, 0.5, solver='prior_ransac', priorRT=np.eye(4)[:3]) == (None, 0, 0, 0)
    es = solver.EssentialMatrixSolver({"EMAT_RANSAC": {"PIX_THRESHOLD": 2.0, "CONFIDENCE": 0.9999}}, True)
    (R, t, n), a, b = es.estimate_pose(np.zeros((3, 2)), np.zeros((3, 2)), {"K_color0": K[None], "K_color1": K[None]})
    assert np.array_equal(R, np.eye(3)) and np.array_equal(t, np.zeros(3)) and (n, a, b) == (0, 0, 0)
    with pytest.raises(NotImplementedError):
        solver.RANSAC(model_type='homography')


def test_bench_reference_arm_runs_every_workload_contract():
    """`bench.py --impl reference` prints the contract's JSON line for a cheap workload (the CPU port, no GPU)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "micro_4096x2048", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["workload"] == "micro_4096x2048" and line["scaling"] == "strong"
