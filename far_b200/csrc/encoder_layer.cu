// LoFTREncoderLayer.forward (mp3d_loftr/src/loftr/loftr_module/transformer.py:44-67), masks None.
// Host-side composition of the library's own kernels so the Python side makes one C-ABI call per layer.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace far {
int linear_dispatch(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
                    const float* bias, float* y, int ldy, int M, int N, int act, int act_cols, int engine,
                    float* workspace, size_t workspace_bytes, cudaStream_t st);
int linear_dispatch_ps(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
                       const float* presplit, float* y, int ldy, int M, int N, int act, int act_cols, int engine,
                       float* workspace, size_t workspace_bytes, cudaStream_t st);
int linear_attention_dispatch(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out,
                              int ldo, int N, int L, int S, int H, int D, float eps, int applied, float* workspace,
                              size_t workspace_bytes, cudaStream_t st);
size_t linear_attention_ws_bytes(int N, int S, int H, int D);
int la_reduce_summed(const float* k, int ldk, const float* v, int ldv, int N, int S, int applied, float* workspace,
                     size_t workspace_bytes, const float** summed_out, cudaStream_t st);
int la_fold_merge(const float* summed, const float* Wm, int N, int C, int S, float* bhi, float* blo, bool cat,
                  cudaStream_t st);
int la_reduce_fold(const float* kv, int N, int S, const float* Wm, float* workspace, size_t workspace_bytes, float* bhi,
                   float* blo, bool cat, const float** summed_out, cudaStream_t st);

static inline size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

struct LayerPlan {
  size_t q, k, v, attn, msg, hid, la, bn, lin, lin_bytes, total;
};
static LayerPlan plan_layer(long long N, int L, int S, int C, int nhead) {
  LayerPlan p;
  size_t off = 0;
  const size_t rowsL = (size_t)N * L, rowsS = (size_t)N * S;
  p.q = off; off += align_up(rowsL * C * 4);
  p.k = off; off += align_up(rowsS * C * 4);
  p.v = off; off += align_up(rowsS * C * 4);
  p.attn = off; off += align_up(rowsL * C * 4);
  p.msg = off; off += align_up(rowsL * C * 4);
  p.hid = off; off += align_up(rowsL * 2 * C * 4);
  p.la = off; off += align_up(linear_attention_ws_bytes((int)N, S, nhead, C / nhead));
  p.bn = off; off += 2 * align_up((size_t)N * C * C * 4) + 1024;  // per-batch-element folded merge operand (hi | lo)
  // hi/lo operand split of the tcgen05 engine, sized for the layer's largest GEMM (mlp.0: [rows, 2C] x [2C, 2C])
  const size_t rows = rowsL > rowsS ? rowsL : rowsS;
  p.lin_bytes = tc_linear_workspace_bytes((int)rows, 2 * C, 2 * C);
  p.lin = off; off += align_up(p.lin_bytes);
  p.total = off;
  return p;
}
}  // namespace far

using namespace far;

extern "C" size_t far_loftr_encoder_layer_workspace_bytes(int N, int L, int S, int C, int nhead) {
  return plan_layer(N, L, S, C, nhead).total + 256;
}

extern "C" int far_loftr_encoder_layer(const float* x, const float* source, float* out, int N, int L, int S, int C,
                                       int nhead, const far_encoder_layer_weights* w, int engine, float* workspace,
                                       size_t workspace_bytes, void* stream) {
  if (N <= 0 || L <= 0) return FAR_OK;
  FAR_REQUIRE(x && source && out && w && workspace && C % nhead == 0 && C % 4 == 0);
  const LayerPlan p = plan_layer(N, L, S, C, nhead);
  if (workspace_bytes < p.total) return FAR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* base = reinterpret_cast<char*>(workspace);
  float* q = reinterpret_cast<float*>(base + p.q);
  float* k = reinterpret_cast<float*>(base + p.k);
  float* v = reinterpret_cast<float*>(base + p.v);
  float* attn = reinterpret_cast<float*>(base + p.attn);
  float* msg = reinterpret_cast<float*>(base + p.msg);
  float* hid = reinterpret_cast<float*>(base + p.hid);
  float* la = reinterpret_cast<float*>(base + p.la);
  float* lw = reinterpret_cast<float*>(base + p.lin);
  const size_t lwb = p.lin_bytes;
  const int ML = N * L, MS = N * S, D = C / nhead;
  const float eps1 = w->eps1 > 0.f ? w->eps1 : 1e-5f, eps2 = w->eps2 > 0.f ? w->eps2 : 1e-5f;   // nn.LayerNorm default
  // cached operand splits (far_tc_weight_split): tf32 (hi, lo) form, only meaningful with tf32 cross terms
  const bool ps_on = !tc::tc_cross16_on();
  const float* ps_wq = ps_on ? w->ps_wq : nullptr;
  const float* ps_wkv = ps_on ? w->ps_wkv : nullptr;
  const float* ps_wmerge = ps_on ? w->ps_wmerge : nullptr;
  const float* ps_wmlp0 = ps_on ? w->ps_wmlp0 : nullptr;
  const float* ps_wmlp2 = ps_on ? w->ps_wmlp2 : nullptr;
  auto ps_lo = [](const float* ps, size_t n, size_t k) {
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(ps) + ((n * k * 4 + 1023) & ~size_t(1023)));
  };
  int rc;
  // ---- fused tensor-core schedule (coarse / regress layers: D = 32, 8 heads) ----------------------------------------
  //   [K' | V] = source [Wk; Wv]^T  (one GEMM, elu+1 on the K' half)      -> KV, Ksum  (la_reduce, fixed-order merge)
  //   Bn = S * KV Wm^T per batch element (la_fold_merge)                   Q'' = (elu(x Wq^T)+1) * Z   (GEMM epilogue)
  //   message = Q'' Bn^T  (grouped GEMM: the attention apply step folded into `merge`)  -> norm1 -> mlp -> norm2 + x
  // [K' | V] is one [MS, 2C] row-major block that spans the workspace's k and v slots.
  static const bool fuse_env_off = getenv("FAR_ENC_UNFUSED") != nullptr;
  bool fuse_off = fuse_env_off;
  if (engine == 3) { fuse_off = true; engine = 2; }  // engine 3: tensor-core GEMMs, kernel-per-op schedule (A/B testing)
  const bool fused = !fuse_off && D == 32 && nhead == 8 && engine != 1 && (engine == 2 || tc_engine_default_on()) &&
                     tc_linear_supported(x, C, C, nullptr, 0, 0, w->wq, C, ML, C) && p.v == p.k + (size_t)MS * C * 4 &&
                     (long long)ML * C >= 128LL * 128 * 32 && (long long)MS * C >= 128LL * 128 * 32;
  if (fused) {
    TcLinearEx a{};
    a.x1 = source; a.ldx1 = C; a.K1 = C; a.W = w->wk; a.ldw = C; a.W2 = w->wv; a.N1 = C;
    if (ps_wkv) { a.Whi = ps_wkv; a.Wlo = ps_lo(ps_wkv, 2 * C, C); }
    a.y = k; a.ldy = 2 * C; a.M = MS; a.N = 2 * C; a.act = FAR_ACT_ELU1; a.act_cols = C;
    a.workspace = lw; a.workspace_bytes = lwb;
    if ((rc = tc_linear_ex(a, st))) return rc;
    const float* summed = nullptr;
    char* bnb = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(base + p.bn) + 1023) & ~uintptr_t(1023));
    float* bhi = reinterpret_cast<float*>(bnb);
    float* blo = reinterpret_cast<float*>(bnb + align_up((size_t)N * C * C * 4));
    const bool cat = tc::tc_cross16_on();   // the grouped `message` GEMM reads q in place (raw-A): cross16 operand form
    static const bool la_sync_env = getenv("FAR_LA_SYNC") != nullptr;   // A/B: the round-1 synchronous reduce + partial sum
    if (la_sync_env) {
      if ((rc = la_reduce_summed(k, 2 * C, k + C, 2 * C, N, S, 1, la, workspace_bytes - p.la, &summed, st))) return rc;
      if ((rc = la_fold_merge(summed, w->wmerge, N, C, S, bhi, blo, cat, st))) return rc;
    } else if ((rc = la_reduce_fold(k, N, S, w->wmerge, la, workspace_bytes - p.la, bhi, blo, cat, &summed, st))) {
      return rc;
    }
    TcLinearEx b{};
    b.x1 = x; b.ldx1 = C; b.K1 = C; b.W = w->wq; b.ldw = C; b.y = q; b.ldy = C; b.M = ML; b.N = C;
    if (ps_wq) { b.Whi = ps_wq; b.Wlo = ps_lo(ps_wq, C, C); }
    b.act = FAR_ACT_ELU1; b.act_cols = -1; b.G = N; b.L = L;
    b.ksum = summed; b.ksum_rec = 32 * 32 + 32; b.ksum_off = 32 * 32; b.eps = 1e-6f;
    b.workspace = lw; b.workspace_bytes = lwb;
    if ((rc = tc_linear_ex(b, st))) return rc;
    TcLinearEx c{};
    c.x1 = q; c.ldx1 = C; c.K1 = C; c.Whi = bhi; c.Wlo = blo; c.wlo_is_cat = cat ? 1 : 0; c.b_grouped = 1; c.y = msg; c.ldy = C; c.M = ML; c.N = C;
    c.act = FAR_ACT_NONE; c.act_cols = -1; c.G = N; c.L = L; c.workspace = lw; c.workspace_bytes = lwb;
    // norm1 without a pass of its own: the `message` GEMM's epilogue leaves per-(row, 32-column) (mean, M2) partials in
    // the (now free) attn slot and the mlp.0 GEMM's converter threads normalise its second A segment on the fly
    const bool ln_fused = tc_ln_fusion_available() && C % 32 == 0 && ps_wmlp0 != nullptr;
    if (ln_fused) c.ln_stats_out = attn;
    if ((rc = tc_linear_ex(c, st))) return rc;
    if (ln_fused) {
      TcLinearEx d{};
      d.x1 = x; d.ldx1 = C; d.K1 = C; d.x2 = msg; d.ldx2 = C; d.K2 = C; d.W = w->wmlp0; d.ldw = 2 * C;
      d.Whi = ps_wmlp0; d.Wlo = ps_lo(ps_wmlp0, 2 * C, 2 * C);
      d.y = hid; d.ldy = 2 * C; d.M = ML; d.N = 2 * C; d.act = FAR_ACT_RELU; d.act_cols = -1;
      d.ln_stats_in = attn; d.ln_gamma = w->g1; d.ln_beta = w->b1; d.ln_eps = eps1;
      d.workspace = lw; d.workspace_bytes = lwb;
      if ((rc = tc_linear_ex(d, st))) return rc;
    } else {
      if ((rc = far_layernorm(msg, w->g1, w->b1, nullptr, attn, ML, C, eps1, stream))) return rc;  // attn := LN(msg)
      if ((rc = linear_dispatch_ps(x, C, C, attn, C, C, w->wmlp0, 2 * C, ps_wmlp0, hid, 2 * C, ML, 2 * C, FAR_ACT_RELU, -1, engine, lw, lwb, st))) return rc;
    }
    if ((rc = linear_dispatch_ps(hid, 2 * C, 2 * C, nullptr, 0, 0, w->wmlp2, 2 * C, ps_wmlp2, msg, C, ML, C, FAR_ACT_NONE, -1, engine, lw, lwb, st))) return rc;
    return far_layernorm(msg, w->g2, w->b2, x, out, ML, C, eps2, stream);
  }
  // q/k projections with the elu(x)+1 feature map fused into the epilogue; v plain (:55-57, linear_attention.py:33-34)
  if ((rc = linear_dispatch_ps(x, C, C, nullptr, 0, 0, w->wq, C, ps_wq, q, C, ML, C, FAR_ACT_ELU1, -1, engine, lw, lwb, st))) return rc;
  const bool kv_fused = !fuse_off && engine != 1 && (engine == 2 || tc_engine_default_on()) &&
                        tc_linear_supported(source, C, C, nullptr, 0, 0, w->wk, C, MS, 2 * C) && C % 32 == 0 &&
                        p.v == p.k + (size_t)MS * C * 4 && (long long)MS * 2 * C >= 128LL * 128 * 32;
  if (kv_fused) {  // [K' | V] = source [Wk; Wv]^T in one tensor-core GEMM (source read and split once)
    TcLinearEx a{};
    a.x1 = source; a.ldx1 = C; a.K1 = C; a.W = w->wk; a.ldw = C; a.W2 = w->wv; a.N1 = C;
    if (ps_wkv) { a.Whi = ps_wkv; a.Wlo = ps_lo(ps_wkv, 2 * C, C); }
    a.y = k; a.ldy = 2 * C; a.M = MS; a.N = 2 * C; a.act = FAR_ACT_ELU1; a.act_cols = C;
    a.workspace = lw; a.workspace_bytes = lwb;
    if ((rc = tc_linear_ex(a, st))) return rc;
    if ((rc = linear_attention_dispatch(q, C, k, 2 * C, k + C, 2 * C, attn, C, N, L, S, nhead, D, 1e-6f, 1, la,
                                        workspace_bytes - p.la, st))) return rc;
  } else {
    if ((rc = linear_dispatch(source, C, C, nullptr, 0, 0, w->wk, C, nullptr, k, C, MS, C, FAR_ACT_ELU1, -1, engine, lw, lwb, st))) return rc;
    if ((rc = linear_dispatch(source, C, C, nullptr, 0, 0, w->wv, C, nullptr, v, C, MS, C, FAR_ACT_NONE, -1, engine, lw, lwb, st))) return rc;
    if ((rc = linear_attention_dispatch(q, C, k, C, v, C, attn, C, N, L, S, nhead, D, 1e-6f, 1, la,
                                        workspace_bytes - p.la, st))) return rc;
  }
  // merge + norm1 (:58-59)
  if ((rc = linear_dispatch_ps(attn, C, C, nullptr, 0, 0, w->wmerge, C, ps_wmerge, msg, C, ML, C, FAR_ACT_NONE, -1, engine, lw, lwb, st))) return rc;
  if ((rc = far_layernorm(msg, w->g1, w->b1, nullptr, attn, ML, C, eps1, stream))) return rc;  // attn := LN(msg)
  // mlp([x | message]) (:62-63): two K-segments instead of a materialised concat
  if ((rc = linear_dispatch_ps(x, C, C, attn, C, C, w->wmlp0, 2 * C, ps_wmlp0, hid, 2 * C, ML, 2 * C, FAR_ACT_RELU, -1, engine, lw, lwb, st))) return rc;
  if ((rc = linear_dispatch_ps(hid, 2 * C, 2 * C, nullptr, 0, 0, w->wmlp2, 2 * C, ps_wmlp2, msg, C, ML, C, FAR_ACT_NONE, -1, engine, lw, lwb, st))) return rc;
  // norm2 + residual (:64-66)
  return far_layernorm(msg, w->g2, w->b2, x, out, ML, C, eps2, stream);
}
