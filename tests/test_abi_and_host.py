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
    assert lib.far_abi_version() == 1
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
