"""The tcgen05 3xTF32 engine (far_b200/csrc/tc_gemm.cu: TMA -> UMMA kind::tf32 -> TMEM) against fp64 and against the
fp32 CUDA-core engine.  3xTF32 must be fp32-accurate: the bar here is the SAME tolerance the CUDA-core engine meets."""
import pytest
import torch

from oracle import far_oracle as O
from far_b200 import ops
from far_b200._lib import ACT_NONE, ACT_RELU, ACT_ELU1, ENGINE_SIMT, ENGINE_TCGEN05
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("M,N,K", [(4800, 512, 512), (4800, 256, 256), (1000, 768, 256), (5000, 256, 280),
                                   (9600, 256, 512), (130, 4100, 64)])
@pytest.mark.parametrize("act", [ACT_NONE, ACT_ELU1])
def test_tc_linear_matches_fp64(M, N, K, act):
    g = O.rng(M + N + K)
    x, w, b = O.randn(g, M, K), O.randn(g, N, K, scale=K ** -0.5), O.randn(g, N, scale=0.1)
    y = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), act, engine=ENGINE_TCGEN05)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    if act == ACT_ELU1:
        ref = torch.nn.functional.elu(ref) + 1
    y2 = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), act, engine=ENGINE_SIMT)
    for name, out in (("tcgen05", y), ("cuda-core", y2)):
        e = out.double().cpu() - ref
        print(f"{name} {M}x{N}x{K}: max|err| {e.abs().max():.2e}  rms {e.pow(2).mean().sqrt():.2e}  "
              f"signed bias/|ref| {(e * torch.sign(ref)).mean() / ref.abs().mean():.2e}")
    assert_close(y, ref, 2e-5, 1e-5, f"tcgen05 linear {M}x{N}x{K}")
    # the two engines agree to accumulation-order noise (tensor-core fp32 accumulation truncates, see tc_gemm.cu)
    assert_close(y, y2, 2e-5, 1e-5, "tcgen05 vs CUDA-core engine")


def test_tc_linear_two_segments():
    g = O.rng(77)
    x1, x2 = O.randn(g, 4800, 256), O.randn(g, 4800, 256)
    w = O.randn(g, 512, 512, scale=0.05)
    y = ops.linear(x1.to(DEV), w.to(DEV), None, ACT_RELU, x2=x2.to(DEV), engine=ENGINE_TCGEN05)
    ref = torch.relu(torch.nn.functional.linear(torch.cat([x1, x2], -1).double(), w.double()))
    assert_close(y, ref, 2e-5, 1e-5, "tcgen05 two-segment linear")


def test_tc_linear_adversarial_magnitudes():
    """Mixed magnitudes / exact cancellation: the hi/lo split must not lose the small terms."""
    g = O.rng(78)
    x = O.randn(g, 1024, 256)
    x[:, ::2] *= 1e3
    x[:, 1::2] *= 1e-3
    w = O.randn(g, 256, 256)
    y = ops.linear(x.to(DEV), w.to(DEV), None, ACT_NONE, engine=ENGINE_TCGEN05)
    ref = torch.nn.functional.linear(x.double(), w.double())
    scale = ref.abs().max().item()
    assert (y.double().cpu() - ref).abs().max().item() < 3e-6 * scale


@pytest.mark.parametrize("engine", [ENGINE_SIMT, 0])
def test_dual_softmax_match_both_engines(engine):
    """CoarseMatching on the CUDA-core engine and on the tcgen05 score passes: identical ordered indices vs the oracle."""
    g = O.rng(61)
    c0, c1 = O.randn(g, 2, 40 * 50, 256, scale=3.0), O.randn(g, 2, 36 * 44, 256, scale=3.0)
    m = ops.dual_softmax_match(c0.to(DEV), c1.to(DEV), (40, 50), (36, 44), 0.0, 2, 0.1, 8.0, 8.0,
                               return_conf_matrix=True, engine=engine)
    o = O.coarse_matching(c0, c1, (40, 50), (36, 44), 0.0, 2, 0.1, 8.0)
    for k in ("b_ids", "i_ids", "j_ids"):
        assert torch.equal(m[k].cpu(), o[k]), (k, m[k].numel(), o[k].numel())
    assert_close(m["mconf"], o["mconf"], 1e-6, 1e-4, "mconf")
    assert_close(m["conf_matrix"], o["conf_matrix"], 1e-7, 1e-4, "conf_matrix")


@pytest.mark.parametrize("N,L,S", [(3, 1000, 1100), (2, 4800, 4800), (5, 777, 777)])
def test_fused_encoder_layer_schedule(N, L, S):
    """The fused tensor-core schedule of the encoder layer ([K'|V] in one GEMM, Z in the q epilogue, attention-apply
    folded into a per-batch-element merge operand; ragged L: tiles must not straddle batch elements) against the
    fp64 oracle and against the kernel-per-op schedule (engine 3)."""
    from oracle import far_oracle as O
    C, H = 256, 8
    g = O.rng(100 + N)
    x, src = O.randn(g, N, L, C, scale=1.5), O.randn(g, N, S, C, scale=1.5)
    w = {}
    for k, shp in (("q_proj", (C, C)), ("k_proj", (C, C)), ("v_proj", (C, C)), ("merge", (C, C)),
                   ("mlp0", (2 * C, 2 * C)), ("mlp2", (C, 2 * C))):
        w[k] = O.randn(g, *shp, scale=(2.0 / (shp[0] + shp[1])) ** 0.5)
    for k in ("norm1_w", "norm2_w"):
        w[k] = 1.0 + O.randn(g, C, scale=0.1)
    for k in ("norm1_b", "norm2_b"):
        w[k] = O.randn(g, C, scale=0.1)
    sd = {"q_proj.weight": w["q_proj"], "k_proj.weight": w["k_proj"], "v_proj.weight": w["v_proj"],
          "merge.weight": w["merge"], "mlp.0.weight": w["mlp0"], "mlp.2.weight": w["mlp2"],
          "norm1.weight": w["norm1_w"], "norm1.bias": w["norm1_b"], "norm2.weight": w["norm2_w"],
          "norm2.bias": w["norm2_b"]}
    ref = O.loftr_encoder_layer({k: v.double() for k, v in sd.items()}, x.double(), src.double(), H)
    wd = {k: v.cuda() for k, v in w.items()}
    fused = ops.loftr_encoder_layer(x.cuda(), src.cuda(), wd, H, 0)
    per_op = ops.loftr_encoder_layer(x.cuda(), src.cuda(), wd, H, 3)
    assert_close(fused, ref, 5e-5, 1e-5, "fused schedule vs fp64 oracle")
    assert_close(per_op, ref, 5e-5, 1e-5, "per-op schedule vs fp64 oracle")
    assert (fused - per_op).abs().max().item() < 2e-5


@pytest.mark.parametrize("cross16", [1, 0])
def test_tc_linear_cross_term_modes(cross16):
    """far_tc_set_cross16: bf16 cross terms (default; 2/3 of the tensor cycles) vs all-tf32 cross terms.  Both must meet
    the engine's fp32-level bar against fp64, on well-scaled, two-segment and adversarial mixed-magnitude operands; the
    measured error statistics are printed (pytest -s) for profiles/."""
    from far_b200 import _lib
    lib = _lib.load()
    prev = lib.far_tc_set_cross16(cross16)
    try:
        g = O.rng(500)
        for (M, N, K) in ((4800, 512, 512), (9600, 256, 256), (2000, 768, 1024)):
            x, w = O.randn(g, M, K, scale=1.5), O.randn(g, N, K, scale=K ** -0.5)
            y = ops.linear(x.to(DEV), w.to(DEV), None, ACT_NONE, engine=ENGINE_TCGEN05).double().cpu()
            ref = torch.nn.functional.linear(x.double(), w.double())
            e = y - ref
            print(f"cross16={cross16} {M}x{N}x{K}: max|err| {e.abs().max():.2e} rms {e.pow(2).mean().sqrt():.2e} "
                  f"rms/|ref|rms {e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt():.2e} "
                  f"signed bias/|ref| {(e * torch.sign(ref)).mean() / ref.abs().mean():.2e}")
            assert_close(y, ref, 2e-5, 1e-5, f"cross16={cross16} linear {M}x{N}x{K}")
            assert e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt() < 5e-6     # measured 0.9e-6 .. 2.7e-6 (tf32 cross terms)
        x = O.randn(g, 1024, 256)
        x[:, ::2] *= 1e3
        x[:, 1::2] *= 1e-3
        w = O.randn(g, 256, 256)
        y = ops.linear(x.to(DEV), w.to(DEV), None, ACT_NONE, engine=ENGINE_TCGEN05)
        ref = torch.nn.functional.linear(x.double(), w.double())
        assert (y.double().cpu() - ref).abs().max().item() < 3e-6 * ref.abs().max().item()
        # exact cancellation: y = x w^T - x w^T must stay at rounding level relative to the summands
        xx = O.randn(g, 512, 256)
        ww = O.randn(g, 128, 128)
        wcat = torch.cat([ww, -ww], dim=1)
        y0 = ops.linear(torch.cat([xx[:, :128], xx[:, :128]], 1).contiguous().to(DEV), wcat.to(DEV), None, ACT_NONE,
                        engine=ENGINE_TCGEN05)
        assert y0.abs().max().item() < 2e-5 * (xx[:, :128].abs().max() * ww.abs().max() * 128) ** 0.5 * 16
    finally:
        lib.far_tc_set_cross16(prev)
