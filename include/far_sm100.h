/* libfar_sm100.so -- C ABI of the B200-native FAR per-pair pose hot path.
 *
 * The reference (crockwell/far) is pure Python/PyTorch and has no FFI layer: its "plugin interface" for
 * this path is a set of nn.Module.forward() methods.  Each entry point below replaces the op
 * sequence of one of those methods; the comment on each cites the reference file:line it replaces
 * (paths relative to /root/reference).  The Python host side (far_b200/) keeps the reference's module /
 * parameter names and forward signatures and binds these symbols with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions: plain pointers + sizes, no torch types.  All tensors are fp32, row-major, device memory,
 * 16-byte aligned unless stated; index tensors are int64 (what torch.where yields).  `stream` is a
 * cudaStream_t passed as void*.  Functions never allocate and never synchronise; scratch memory comes from
 * the caller (`*_workspace_bytes` says how much).  Return: FAR_OK or an FAR_ERR_* code (never throws).
 */
#ifndef FAR_SM100_H_
#define FAR_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FAR_OK 0
#define FAR_ERR_ARG 1       /* shape / alignment / null-pointer precondition violated */
#define FAR_ERR_CUDA 2      /* a kernel launch failed (cudaGetLastError) */
#define FAR_ERR_WORKSPACE 3 /* workspace too small */

#define FAR_ACT_NONE 0
#define FAR_ACT_RELU 1    /* nn.ReLU */
#define FAR_ACT_GELU 2    /* nn.GELU (exact erf) */
#define FAR_ACT_ELU1 3    /* F.elu(x) + 1  (linear_attention.py:10-11) */
#define FAR_ACT_SIGMOID 4 /* nn.Sigmoid */

/* ABI version of this header (bumped on any signature change). */
int far_abi_version(void);
/* Number of kernels this library has launched in this process (bench.py's `gpu_launches`). */
unsigned long long far_launch_count(void);

/* ---- per-kernel device timing (measurement only; off by default) ---------------------------------------
 * far_profile_enable(1) makes the library bracket every launch of its heavy kernels with a CUDA-event pair on the
 * launching stream (bench.py's `roofline` entry: the dominant KERNEL's own duration, live, inside the timed region);
 * far_profile_enable(0) stops and frees the events.  far_profile_read() synchronises the recorded events and returns,
 * for kernel class `id` (0 <= id < far_profile_num_ids()), total device ms, launches, and the ALGORITHMIC flops and
 * HBM bytes of those launches (the per-unit figures of SURVEY.md 8d / DESIGN.md 4).  Returns 0, or FAR_ERR_ARG. */
int far_profile_num_ids(void);
const char* far_profile_name(int id);
int far_profile_enable(int on);
int far_profile_read(int id, double* total_ms, unsigned long long* launches, double* flops, double* bytes);

/* Precision mode of the tcgen05 GEMM engine's correction terms (returns the previous setting).  x = hi + lo per operand
 * (hi = tf32); the product is hi*hi + (hi*lo + lo*hi).  on = 0 (default): all three products in tf32 (relative accuracy
 * 2^-21 per product).  on = 1 (env FAR_TC_CROSS=bf16 at start-up): the two 2^-11-times-smaller cross products run as one
 * bf16 MMA over K-concatenated operands (2^-19 per product, 2/3 of the tensor cycles; measured 2 % faster on B200 because
 * the kernel is paced by the shared-memory port).  Either way the engine is held to 2e-5 + 1e-5 |ref| against fp64
 * (tests/test_gpu_tcgen05.py); the score / EMM kernels always use tf32 cross terms. */
int far_tc_set_cross16(int on);

/* Diagnostics of the TMEM-operand GEMM kernels (kernel = 0: csrc/tc_gemm_ts.cu, 1: the CTA-pair kernel
 * csrc/tc_gemm_pair.cu): with env FAR_TC_DBG bit 256 set, CTA 0 of every launch accounts, per warp role, the cycles spent
 * waiting on each pipeline barrier; this reads the 16 counters of the last launch (layout documented at g_ts_prof).
 * kernel = 2: `out16` must hold 256 entries and receives the clock64 timeline of one tile of the pair kernel (g_tp_trace).
 * kernel = 3 (FAR_TC_DBG bit 1024): 320 entries, (total cycles, tiles) of every CTA of the last pair-kernel launch.
 * Returns 0, or -1 on a CUDA error. */
int far_tc_debug_counters(int kernel, unsigned long long* out16);

/* ---- nn.Linear family --------------------------------------------------------------------------------
 * y[M,N] = act( [x1 | x2] * W^T + bias ),  x1:[M,K1] (ld ldx1), x2:[M,K2] (ld ldx2, may be NULL with K2=0),
 * W:[N,K1+K2] (ld ldw), bias:[N] or NULL, act applied to columns < act_cols only (act_cols<0: all).
 * Replaces every nn.Linear on the path, e.g. q/k/v/merge/mlp of LoFTREncoderLayer
 * (mp3d_loftr/src/loftr/loftr_module/transformer.py:23-38,55-63; the [x|message] concat of :63 is the
 * x1|x2 pair), FinePreprocess.down_proj/merge_feat (fine_preprocess.py:19-20), CrossAttention.qkv /
 * proj_fundamental (transformer.py:261-263), Mlp.fc1/fc2 (vit_layers/mlp.py:16-18), the FAR MLP heads
 * (transformer.py:385-405; interiornetStreetlearn_8ptVit/src/model.py:94-108;
 * mapfree_6dreg/lib/models/regression/model.py:66-84).
 * `engine`: 0 = auto, 1 = fp32 CUDA-core tile engine, 2 = tcgen05 3xTF32 (TMA-fed; needs K%32==0,
 * contiguous x1 (x2 NULL) and W).  Split-K is chosen internally when M is small and K large. */
size_t far_linear_workspace_bytes(int M, int N, int K);
int far_linear(const float* x1, int ldx1, int K1, const float* x2, int ldx2, int K2, const float* W, int ldw,
               const float* bias, float* y, int ldy, int M, int N, int act, int act_cols, int engine,
               float* workspace, size_t workspace_bytes, void* stream);

/* y[r,:] = (residual ? residual[r,:] : 0) + LayerNorm(x[r,:]) * gamma + beta ; C <= 1024.
 * Replaces nn.LayerNorm + the `x + message` residual (transformer.py:59,64-66; CrossBlock norms :339-347). */
int far_layernorm(const float* x, const float* gamma, const float* beta, const float* residual, float* y,
                  int rows, int C, float eps, void* stream);
/* Same with a broadcast pre-add: LayerNorm(x[r,:] + pre_add[r % pre_rows, :]) -- CrossBlock's
 * `x = x + self.pos_embed` followed by norm1 (transformer.py:337,343); pre_add may be NULL. */
int far_layernorm_pre(const float* x, const float* pre_add, int pre_rows, const float* gamma, const float* beta,
                      const float* residual, float* y, int rows, int C, float eps, void* stream);

/* out[n, h*W+w, c] = feat[n,c,h,w] + pe[c,h,w]  (PositionEncodingSine.forward + rearrange
 * 'n c h w -> n (h w) c': mp3d_loftr/src/loftr/utils/position_encoding.py:37-42, loftr.py:100-101).
 * feat given with element strides (sn, sc, sh, sw) so NCHW and channels_last both work;
 * pe_hwc is the sine table laid out [H*W, C]. */
int far_pos_encode_flatten(const float* feat, long long sn, long long sc, long long sh, long long sw,
                           const float* pe_hwc, float* out, int N, int C, int H, int W, void* stream);

/* ---- ResNet-FPN top-down glue on channels_last (NHWC) maps ----------------------------------------------
 * out = skip + F.interpolate(low, scale_factor=2., mode='bilinear', align_corners=True)
 * (mp3d_loftr/src/loftr/backbone/resnet_fpn.py:106-112; skip may be NULL = plain upsample).
 * low:[N,Hin,Win,C], skip/out:[N,2Hin,2Win,C], all dense NHWC, C % 4 == 0, 16-byte aligned. */
int far_upsample2x_add_nhwc(const float* low, const float* skip, float* out, int N, int Hin, int Win, int C,
                            void* stream);
/* In place x[p,c] = leaky_relu(x[p,c]*scale[c] + shift[c], negative_slope): eval-mode BatchNorm2d (scale =
 * gamma/sqrt(var+eps), shift = beta - mean*scale) followed by nn.LeakyReLU / nn.ReLU (slope 0)
 * (resnet_fpn.py:84-95 layer{1,2}_outconv2).  scale may be NULL (bias only).  x:[pixels, C] NHWC, C % 4 == 0. */
int far_scale_shift_act_nhwc(float* x, const float* scale, const float* shift, long long pixels, int C,
                             float negative_slope, void* stream);
/* Out-of-place variant (y != x): the pre-activation `relu(bn(x))` of PreActBlock / PreActBottleneck, where x itself stays
 * alive as the identity shortcut (mapfree_6dreg/lib/models/regression/encoder/preact.py:35-41, 68-74). */
int far_scale_shift_act_nhwc_out(const float* x, float* y, const float* scale, const float* shift, long long pixels,
                                 int C, float negative_slope, void* stream);
/* 8pt-ViT input preprocessing (interiornetStreetlearn_8ptVit/src/model.py:131-141): images [n,3,H,W] fp32 BGR 0..255 ->
 * out [n,3,OH,OW] RGB = nearest-resize(((x / 255) - mean) / std), the reference's arithmetic on exactly the pixels
 * F.interpolate(mode='nearest') samples (resize and per-pixel arithmetic commute: bit-identical, 6x fewer bytes read).
 * mean3 / std3: HOST pointers to the 3 RGB constants. */
int far_vit_preprocess(const float* images, float* out, long long n, int H, int W, int OH, int OW, const float* mean3,
                       const float* std3, void* stream);
/* Backbone stem: y = relu(conv2d(x, w, stride 2, padding 3) + bias) for a ONE-channel input and a 7x7 kernel,
 * x:[N,1,H,W] fp32, w:[Cout,1,7,7] (eval BatchNorm folded in), y:[N,OH,OW,Cout] NHWC, Cout == 128.
 * Replaces `self.relu(self.bn1(self.conv1(x)))` (mp3d_loftr/src/loftr/backbone/resnet_fpn.py:52-54,80); cuDNN has no
 * tensor-core engine for C_in = 1.  Exact fp32 FMA arithmetic. */
size_t far_stem_conv_workspace_bytes(int Cout);
int far_stem_conv7x7s2_relu_nhwc(const float* x, const float* w, const float* bias, float* y, int N, int H, int W,
                                 int Cout, float* workspace, size_t workspace_bytes, void* stream);

/* ---- LinearAttention.forward (mp3d_loftr/src/loftr/loftr_module/linear_attention.py:20-52) -----------
 * q:[N,L,H*D], k,v:[N,S,H*D] (row strides ldq/ldk/ldv), out:[N,L,H*D] (ld ldo).
 * feature_map_applied != 0 means q,k already hold elu(x)+1 (fused into the projection epilogue). */
size_t far_linear_attention_workspace_bytes(int N, int S, int H, int D);
int far_linear_attention(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out,
                         int ldo, int N, int L, int S, int H, int D, float eps, int feature_map_applied,
                         float* workspace, size_t workspace_bytes, void* stream);

/* ---- LoFTREncoderLayer.forward (transformer.py:44-67), masks None ------------------------------------
 * x:[N,L,C] source:[N,S,C] -> out:[N,L,C] (out may alias x: every read of x precedes the final write).
 * Weights are the layer's own tensors: wq,wk,wv,wmerge [C,C]; wmlp0 [2C,2C]; wmlp2 [C,2C]; LayerNorm
 * gamma/beta [C] x2.  Composition: (q|k|v projections with fused elu+1) -> far_linear_attention ->
 * merge -> LN -> mlp.0([x|msg]) ReLU -> mlp.2 -> LN + residual. */
typedef struct {
  const float *wq, *wk, *wv, *wmerge, *wmlp0, *wmlp2;
  const float *g1, *b1, *g2, *b2;
  /* ABI version 2 */
  float eps1, eps2;   /* eps of norm1 / norm2 (<= 0: nn.LayerNorm's default 1e-5) */
  /* Optional cached tcgen05 operands of the static weights: outputs of far_tc_weight_split for wq [C,C], the stacked
   * [wk; wv] [2C,C], wmerge [C,C], wmlp0 [2C,2C], wmlp2 [C,2C]; NULL = split on every call.  The caller owns the buffers
   * and must refresh them when the weights change. */
  const float *ps_wq, *ps_wkv, *ps_wmerge, *ps_wmlp0, *ps_wmlp2;
} far_encoder_layer_weights;
/* tcgen05 operand form of a static weight matrix W [N,K] (row pitch ldw): tf32 hi | lo pair, computed once per weight
 * instead of on every GEMM call.  out: 1024-byte aligned, far_tc_weight_split_bytes(N, K) bytes. */
size_t far_tc_weight_split_bytes(int N, int K);
int far_tc_weight_split(const float* W, int ldw, int N, int K, float* out, size_t out_bytes, void* stream);
size_t far_loftr_encoder_layer_workspace_bytes(int N, int L, int S, int C, int nhead);
int far_loftr_encoder_layer(const float* x, const float* source, float* out, int N, int L, int S, int C,
                            int nhead, const far_encoder_layer_weights* w, int engine, float* workspace,
                            size_t workspace_bytes, void* stream);

/* ---- CoarseMatching.forward + get_coarse_match (mp3d_loftr/src/loftr/utils/coarse_matching.py:86-265),
 * dual_softmax branch, eval, no padding masks.  Two calls because M is data dependent (the reference
 * synchronises in torch.where, :193):
 *   _select: sim = (f0/sqrt(C)).(f1/sqrt(C))^T / temperature; row/col log-sum-exp; conf = softmax(sim,1) *
 *            softmax(sim,2); threshold (strict >), border_rm frame on both grids, mutual-nearest test; writes
 *            per-row decisions into the workspace, the total count to *num_matches (device int64) and, when
 *            conf_out != NULL, the dense conf matrix [N,L,S] (`data['conf_matrix']`, :144).
 *   _gather: order-preserving compaction into b_ids,i_ids,j_ids (ascending (b,i)), mconf, mkpts0_c,
 *            mkpts1_c = (id % w, id / w) * scale  (:246-254).  Caller sizes outputs from *num_matches. */
size_t far_dual_softmax_match_workspace_bytes(int N, int L, int S);
int far_dual_softmax_match_select(const float* feat0, const float* feat1, int N, int L, int S, int C,
                                  float temperature, float thr, int border_rm, int h0c, int w0c, int h1c,
                                  int w1c, float* conf_out, long long* num_matches, int engine,
                                  float* workspace, size_t workspace_bytes, void* stream);
int far_dual_softmax_match_gather(int N, int L, int w0c, int w1c, float scale0, float scale1,
                                  long long num_matches, long long* b_ids, long long* i_ids, long long* j_ids,
                                  float* mconf, float* mkpts0_c, float* mkpts1_c, const float* workspace,
                                  size_t workspace_bytes, void* stream);

/* ---- FinePreprocess.forward (mp3d_loftr/src/loftr/loftr_module/fine_preprocess.py:29-59) -------------
 * Gathers the WxW (stride `stride`, padding W/2) windows of the fine maps at the M matches directly
 * (no F.unfold materialisation), projects the matched coarse tokens with down_proj, and applies merge_feat
 * to [window | coarse] -> out0,out1 [M, W*W, Cf].  feat_f* given with element strides (channels_last or
 * NCHW).  feat_c*: [N, L, Cc] post-transformer coarse features. */
size_t far_fine_preprocess_workspace_bytes(long long M, int WW, int Cf, int Cc);
int far_fine_preprocess(const float* feat_f0, const float* feat_f1, long long sn, long long sc, long long sh,
                        long long sw, int Hf, int Wf, int Cf, const float* feat_c0, const float* feat_c1, int L0,
                        int L1, int Cc, const long long* b_ids, const long long* i_ids, const long long* j_ids,
                        long long M, int W, int stride, int w0c, int w1c, const float* down_w,
                        const float* down_b, const float* merge_w, const float* merge_b, float* out0, float* out1,
                        float* workspace, size_t workspace_bytes, void* stream);

/* ---- FineMatching.forward + get_fine_match (mp3d_loftr/src/loftr/utils/fine_matching.py:15-76) -------
 * One warp per match: centre-token correlation, softmax(./sqrt(C)) over the WW window, spatial expectation
 * on the [-1,1] grid (kornia dsnt semantics), std; expec_f [M,3]; mkpts1_f = mkpts1_c + coords*offset_scale
 * with offset_scale = (W//2) * (hw0_i/hw0_f)  (:71). */
int far_fine_match(const float* feat_f0, const float* feat_f1, long long M, int WW, int C,
                   const float* mkpts1_c, float offset_scale, float* expec_f, float* mkpts1_f, void* stream);

/* ---- run_8point (third_party/prior_ransac/cv_geometry.py:772-833, incl. normalize_points :713-750 and
 * normalize_transformation :753-769).  One warp per pair, register-resident 9x9 accumulation, cyclic-Jacobi
 * eigen-decomposition and 3x3 SVD (no cuSOLVER).  pts1,pts2:[P,N,2]; weights:[P,N] or NULL;
 * counts:[P] (int32) valid correspondences per pair or NULL (= N).  F:[P,3,3]. */
size_t far_eight_point_workspace_bytes(int P);
int far_eight_point(const float* pts1, const float* pts2, const float* weights, const int* counts, int P, int N,
                    float* F, float* workspace, size_t workspace_bytes, void* stream);

/* ---- decompose_essential_matrix / motion_from_essential (third_party/prior_ransac/essential.py:41-139)
 * E:[P,3,3] -> R1,R2:[P,3,3], t:[P,3].  One thread per matrix (3x3 Jacobi SVD in registers). */
int far_essential_decompose(const float* E, int P, float* R1, float* R2, float* t, void* stream);

/* ---- per-pair solver glue for ragged matches (replaces the python loop of spvs_RT,
 * mp3d_loftr/src/loftr/utils/supervision.py:184-233 -> estimate_pose metrics.py:80-174, with the in-repo
 * weighted 8-point as the model solver).  Matches of pair b are the contiguous
 * segment [offsets[b], offsets[b+1]) of mkpts0/mkpts1/mconf (m_bids is sorted).  K0,K1:[N,3,3] fp32.
 * Outputs per pair: E [3,3], Rt [3,4] (cheirality-selected candidate; identity if < 8 matches),
 * n_pos [int32] = cheirality votes of the winner. */
int far_pose_from_matches(const float* mkpts0, const float* mkpts1, const float* mconf,
                          const long long* offsets, int N, const float* K0, const float* K1, float* E,
                          float* Rt, int* n_pos, float* workspace /* far_eight_point_workspace_bytes(N) */,
                          size_t workspace_bytes, void* stream);

/* ---- prior-guided RANSAC round, scoring step --------------------------------------------
 * For every pair p (ragged matches [offsets[p], offsets[p+1]) of pixel keypoints, K-normalised on the fly as
 * mp3d_loftr/src/utils/metrics.py:88-89) and every hypothesis h of models [P,H,3,3]:
 *   scores[p,h] = #{ i : sampson_sq(E, x0_i, x1_i) <= inl_th } + prior(E)          (ransac.py:256-275 `verify`)
 *   prior(E)    = -(min_k mean |[R_k | T] pcl - prior_rt[p] pcl|)^2 / prior_lambda  (ransac.py:203-231, :401-404;
 *                 (R_1, R_2, T) = decompose_essential_matrix(E); prior_rt translation normalised as in :180;
 *                 prior_rt == NULL: no prior term)
 *   -inf for degenerate models (min |diag E| <= 1e-4, ransac.py:303-308) and for pairs with < 8 matches.
 * Then per pair: best_idx = argmax_h (lowest index on ties; -1 if none), best_E [P,3,3], counts3 [P,3] = inliers of the
 * winner at inl_th, inl_th/10, inl_th/100 (:279-283), inlier_mask [M] (uint8): bit 0 = inlier at inl_th, bit 1 = at
 * inl_th/10 (`inliers_best_tight`), bit 2 = at inl_th/100 (`inliers_best_ultra_tight`). */
int far_prior_ransac_score(const float* mkpts0, const float* mkpts1, const long long* offsets, int P, const float* K0,
                           const float* K1, const float* models, int H, const float* prior_rt, const float* pcl,
                           int npcl, float prior_lambda, float inl_th, float* scores, int* best_idx, float* best_E,
                           int* counts3, unsigned char* inlier_mask, void* stream);
/* ---- prior-guided RANSAC round, sampling + minimal solver (no eager-torch math, no host sync) -----------------------
 * far_segment_offsets: m_bids [M] int64, sorted ascending (what get_coarse_match yields, coarse_matching.py:193) ->
 *   offsets [P+1] int64 with offsets[b] = first match of pair b (replaces bincount + cumsum).
 * far_ransac_sample_models: for every pair p and hypothesis h < H draws `sample_size` (8, or 6: below) correspondences of the
 *   pair's segment and solves the minimal model with the in-repo normalised 8-point (cv_geometry.py:772-833, unit
 *   weights as ransac.py:250-253) -> models [P,H,3,3].  Sampling distribution (ransac.py:161-175, :358-367):
 *     prior_rt != NULL: p_i ~ exp(-symmetrical_epipolar_distance(x0_i, x1_i, [t]_x R) / bias_sigma_sq) + 1e-4 with the
 *                       prior translation normalised to unit length (`use_linear_bias_sampling`, 'biased');
 *     prior_rt == NULL: uniform.
 *   Draw k of (p,h) = searchsorted(cdf_p, u * cdf_p[-1], 'right'), u = Philox4x32-10(key = seed; counter = (p*H+h,
 *   block, 0, 0))[k mod 4] * 2^-32, blocks consumed in order; an index already in the sample is redrawn (at most 4
 *   times).  The reference uses numpy's global RNG there, so only the distribution is the contract; the counter-based
 *   generator makes samples reproducible and testable.  sample_idx [P,H,8] int32 (segment-local; -1 for pairs with
 *   fewer than 8 matches) may be NULL.  Pairs with < 8 matches get all-zero models (rejected by the scoring step).
 *   sample_size == 6 selects the recipe's model type instead (`essential_cv2`: a 5-point solver on 6 sampled points,
 *   ransac.py:250-253 / cv_geometry.py:836-859): Nister's 5-point (the in-tree batched version cv_geometry.py:861-1041,
 *   restated in csrc/fivept.cuh) on the first five draws, the candidate with the smallest squared Sampson distance at
 *   the sixth is the hypothesis' model (E with unit Frobenius norm); sample_idx is then [P,H,6].  OpenCV's own 5-point
 *   arithmetic is un-vendored, so that mode is pinned to the restated algorithm, not to cv2.
 * far_five_point: every real solution of the 5-point solver for S explicit minimal samples (tests / diagnostics):
 *   pts5 [S,5,4] fp64 calibrated (x1, y1, x2, y2) -> E [S,10,9] fp64 row-major with x2h^T E x1h = 0 (zero padded),
 *   nsol [S]. */
int far_five_point(const double* pts5, int S, double* E, int* nsol, void* stream);
int far_segment_offsets(const long long* m_bids, long long M, int P, long long* offsets, void* stream);
size_t far_ransac_sample_models_workspace_bytes(long long M, int P, int H);
int far_ransac_sample_models(const float* mkpts0, const float* mkpts1, const long long* offsets, long long M, int P,
                             const float* K0, const float* K1, const float* prior_rt, float bias_sigma_sq, int H,
                             int sample_size, unsigned long long seed, float* models, int* sample_idx,
                             float* workspace, size_t workspace_bytes, void* stream);
/* (R | t) [P,3,4] of essential matrices E [P,3,3] by the cheirality vote of far_pose_from_matches (the criterion of
 * cv2.recoverPose, metrics.py:164-170) over the matches with mask != 0 (mask NULL: all); E == 0 -> identity pose. */
size_t far_pose_from_essential_workspace_bytes(int P);
int far_pose_from_essential(const float* mkpts0, const float* mkpts1, const unsigned char* mask,
                            const long long* offsets, int P, const float* K0, const float* K1, const float* E,
                            float* Rt, int* n_pos, float* workspace, size_t workspace_bytes, void* stream);

/* ---- CrossAttention.forward core (mp3d: transformer.py:266-303; 8pt-ViT: vision_transformer.py:177-208)
 * qkv1,qkv2: [B,Ntok,3,h,d] (the output of the shared qkv Linear on the two images), pos: [Bpos,Ntok,6]
 * (Bpos = 1 broadcasts).  Computes, per (b,head), for X in {1,2}:
 *   S_X = q_other k_X^T * scale;  P_X = softmax(S_X,-1) * softmax(S_X,-2);  V'_X = [v_X | pos];
 *   F_X = V'_X^T P_X V'_X   -> F1,F2 [B,h,d+6,d+6]
 * Two passes over S (LSE pass, recompute pass); S and P are never written to HBM. */
size_t far_emm_bilinear_attn_workspace_bytes(int B, int Ntok, int h, int d);
int far_emm_bilinear_attn(const float* qkv1, const float* qkv2, const float* pos, int Bpos, int B, int Ntok,
                          int h, int d, float scale, float* F1, float* F2, int engine, float* workspace,
                          size_t workspace_bytes, void* stream);

/* ---- CorrelationVolumeWarping.forward (mapfree_6dreg/lib/models/regression/aggregator.py:42-116) with the shipped
 * recipe's flags (POSITION_ENCODER + MAX_SCORE_CHANNEL; config/regression/mapfree/rot6d_trans_with_loftr.yaml:8-11):
 *   C = softmax_j(vol0^T vol1) [B,N,N];  out = cat[vol0, vol1 C^T, grid C^T, max_j C]  -> [B, 2D+3, N]
 * vol0, vol1: [B, D, N] given with element strides (sb batch, sc channel, sp pixel): NCHW (sp == 1) or channels_last
 * (sc == 1); D must be 32 (ENCODER.NUM_OUT_LAYERS of the recipe); grid: [2, N] = the meshgrid(linspace(-1,1,H),
 * linspace(-1,1,W)) table of :84-88.  Flash-style tcgen05 kernel: the N x N volume (157 MB/pair at N = 6256) is never
 * written; two sweeps over the key tiles per query tile (exact row maximum, then exponentials + P V'). */
size_t far_corr_volume_warp_workspace_bytes(int B, int N);
int far_corr_volume_warp(const float* vol0, const float* vol1, long long sb, long long sc, long long sp,
                         const float* grid, int B, int N, int D, float* out, float* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- timm-style softmax attention of the 8pt-ViT blocks (vision_transformer.py:236-262):
 * qkv:[B,Ntok,3,h,d] -> out:[B,Ntok,h*d] = softmax(q k^T * scale) v, heads re-interleaved. */
size_t far_softmax_attention_workspace_bytes(int B, int Ntok, int h, int d);
int far_softmax_attention(const float* qkv, int B, int Ntok, int h, int d, float scale, float* out,
                          float* workspace, size_t workspace_bytes, void* stream);

/* ---- FAR gated fusion epilogues -----------------------------------------------------------------------
 * mp3d (forward_emm, transformer.py:427-473, use_simple_moe + use_2wt [+ scale_8pt]): per row b,
 *   t_s = scale_8pt ? renorm(solver t to |regressed t| in un-normalised space, clamp 1e-3..100) : solver t
 *   out[b] = [ w0*pred_t + (1-w0)*t_s | w1*pred_R6 + (1-w1)*solver_R6 ]
 * pred:[B,9], solver:[B,ld_solver>=9] (normalised 9-D first), wt:[B,2], mean/std:[9]. */
int far_pose_blend_mp3d(const float* pred, const float* solver, int ld_solver, const float* wt,
                        const float* mean9, const float* std9, int scale_8pt, float* out, int B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FAR_SM100_H_ */
