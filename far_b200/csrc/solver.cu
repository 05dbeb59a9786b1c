// Weighted normalised 8-point + essential-matrix decomposition, no cuSOLVER.
//   run_8point                  */third_party/prior_ransac/cv_geometry.py:772-833 (+ normalize_points :713-750,
//                               normalize_transformation :753-769)
//   decompose_essential_matrix  */third_party/prior_ransac/essential.py:99-139, motion_from_essential :41-64
//   per-pair python loop        mp3d_loftr/src/loftr/utils/supervision.py:184-233 -> metrics.py:80-174
//
// Stage 1 (HBM-bound, one warp per pair): two coalesced passes over the pair's correspondences: means, then
//   mean distance + the 45 unique entries of  A_c = sum_i w_i xc_i xc_i^T  on CENTRED coordinates, accumulated in
//   fp64 registers.  The Hartley scale is applied afterwards as a diagonal congruence (A = D A_c D), which is
//   exact algebra, so the reference's dense [N,N] diag_embed(w) (16.8 MB/pair at N=2048) never exists.
// Stage 2 (one THREAD per pair, 32 pairs per warp in lock-step): register-resident cyclic Jacobi on the 9x9
//   (fp32, like the reference's fp32 SVD), smallest eigenvector -> F_hat; rank-2 projection through the smallest
//   right-singular vector of F_hat (3x3 Jacobi in fp64 on F^T F):  F_hat - (F_hat v3) v3^T  ==  U diag(s1,s2,0) V^T;
//   de-normalise  T2^T F T1 ; divide by (F22 + 1e-8) when |F22| > 1e-8.
#include "common.cuh"
#include "fivept.cuh"

namespace far {

// ------------------------------------------------------------------------------------------- Jacobi kernels
template <typename T, int N>
__device__ __forceinline__ void jacobi_eig(T (&A)[N][N], T (&V)[N][N], int max_sweeps, T tol) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? T(1) : T(0);
#pragma unroll 1
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    T off = T(0), diag = T(0);
#pragma unroll
    for (int p = 0; p < N; ++p) {
      diag += A[p][p] * A[p][p];
#pragma unroll
      for (int q = p + 1; q < N; ++q) off += A[p][q] * A[p][q];
    }
    if (off <= tol * tol * diag || off == T(0)) break;
#pragma unroll
    for (int p = 0; p < N - 1; ++p) {
#pragma unroll
      for (int q = p + 1; q < N; ++q) {
        const T apq = A[p][q];
        if (apq != T(0)) {
          const T theta = (A[q][q] - A[p][p]) / (T(2) * apq);
          const T t = (theta >= T(0) ? T(1) : T(-1)) / (fabs(theta) + sqrt(theta * theta + T(1)));
          const T c = T(1) / sqrt(t * t + T(1));
          const T s = t * c;
          A[p][p] -= t * apq;
          A[q][q] += t * apq;
          A[p][q] = T(0);
          A[q][p] = T(0);
#pragma unroll
          for (int k = 0; k < N; ++k) {
            if (k != p && k != q) {
              const T akp = A[k][p], akq = A[k][q];
              const T np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
              A[k][p] = np_; A[p][k] = np_;
              A[k][q] = nq_; A[q][k] = nq_;
            }
            const T vkp = V[k][p], vkq = V[k][q];
            V[k][p] = c * vkp - s * vkq;
            V[k][q] = s * vkp + c * vkq;
          }
        }
      }
    }
  }
}

// Eigen-decomposition of the 3x3 symmetric M = F^T F (fp64).  Returns V columns sorted by DESCENDING eigenvalue.
__device__ __forceinline__ void sym3_eig_sorted(const double (&M)[3][3], double (&V)[3][3], double (&lam)[3]) {
  double A[3][3], W[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A[i][j] = M[i][j];
  jacobi_eig<double, 3>(A, W, 30, 1e-15);
  double l0 = A[0][0], l1 = A[1][1], l2 = A[2][2];
  int i0 = 0, i1 = 1, i2 = 2;
  if (l0 < l1) { double tl = l0; l0 = l1; l1 = tl; int ti = i0; i0 = i1; i1 = ti; }
  if (l1 < l2) { double tl = l1; l1 = l2; l2 = tl; int ti = i1; i1 = i2; i2 = ti; }
  if (l0 < l1) { double tl = l0; l0 = l1; l1 = tl; int ti = i0; i0 = i1; i1 = ti; }
  lam[0] = l0; lam[1] = l1; lam[2] = l2;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    V[k][0] = (i0 == 0) ? W[k][0] : (i0 == 1 ? W[k][1] : W[k][2]);
    V[k][1] = (i1 == 0) ? W[k][0] : (i1 == 1 ? W[k][1] : W[k][2]);
    V[k][2] = (i2 == 0) ? W[k][0] : (i2 == 1 ? W[k][1] : W[k][2]);
  }
}

// ------------------------------------------------------------------------------------------- stage 1
// per-pair record in the workspace: 45 doubles (upper triangle of A_c, row-major) + 6 doubles
// (mu1x, mu1y, mu2x, mu2y, s1, s2) + count
constexpr int kRec = 52;

struct DensePts {  // far_eight_point: pts [P,N,2], weights [P,N] or NULL, counts [P] or NULL
  const float* p1; const float* p2; const float* w; const int* counts; int N;
  __device__ __forceinline__ int count(int pair) const { return counts ? min(counts[pair], N) : N; }
  __device__ __forceinline__ void load(int pair, int i, float& x1, float& y1, float& x2, float& y2, float& wt) const {
    const size_t o = (size_t)pair * N + i;
    const float2 a = reinterpret_cast<const float2*>(p1)[o], b = reinterpret_cast<const float2*>(p2)[o];
    x1 = a.x; y1 = a.y; x2 = b.x; y2 = b.y;
    wt = w ? w[o] : 1.f;
  }
};
struct RaggedPts {  // far_pose_from_matches: segment [off[b], off[b+1]) of pixel keypoints, K-normalised on the fly
  const float* mk0; const float* mk1; const float* conf; const long long* off; const float* K0; const float* K1;
  __device__ __forceinline__ int count(int pair) const { return (int)(off[pair + 1] - off[pair]); }
  __device__ __forceinline__ void load(int pair, int i, float& x1, float& y1, float& x2, float& y2, float& wt) const {
    const size_t o = (size_t)off[pair] + i;
    const float2 a = reinterpret_cast<const float2*>(mk0)[o], b = reinterpret_cast<const float2*>(mk1)[o];
    const float* k0 = K0 + (size_t)pair * 9;
    const float* k1 = K1 + (size_t)pair * 9;
    // (kpts - K[[0,1],[2,2]]) / K[[0,1],[0,1]]   (metrics.py:88-89)
    x1 = (a.x - k0[2]) / k0[0]; y1 = (a.y - k0[5]) / k0[4];
    x2 = (b.x - k1[2]) / k1[0]; y2 = (b.y - k1[5]) / k1[4];
    wt = conf ? conf[o] : 1.f;
  }
};

template <class Pts>
__global__ void __launch_bounds__(128) eightpt_accumulate_kernel(Pts pts, int P, double* __restrict__ rec) {
  const int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (pair >= P) return;
  const int n = pts.count(pair);
  double* r = rec + (size_t)pair * kRec;
  if (n < 8) {
    if (lane == 0) r[51] = (double)n;
    return;
  }
  // pass 1: means (normalize_points :736)
  double sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0;
  for (int i = lane; i < n; i += 32) {
    float x1, y1, x2, y2, w;
    pts.load(pair, i, x1, y1, x2, y2, w);
    sx1 += x1; sy1 += y1; sx2 += x2; sy2 += y2;
  }
  const double m1x = warp_sum(sx1) / n, m1y = warp_sum(sy1) / n, m2x = warp_sum(sx2) / n, m2y = warp_sum(sy2) / n;
  // pass 2: mean distance to the centroid (:738) + centred weighted moments
  double d1 = 0, d2 = 0;
  double a[45];
#pragma unroll
  for (int k = 0; k < 45; ++k) a[k] = 0.0;
  for (int i = lane; i < n; i += 32) {
    float fx1, fy1, fx2, fy2, fw;
    pts.load(pair, i, fx1, fy1, fx2, fy2, fw);
    const double x1 = fx1 - m1x, y1 = fy1 - m1y, x2 = fx2 - m2x, y2 = fy2 - m2y, w = fw;
    d1 += sqrt(x1 * x1 + y1 * y1);
    d2 += sqrt(x2 * x2 + y2 * y2);
    // X row = [x2*x1, x2*y1, x2, y2*x1, y2*y1, y2, x1, y1, 1]   (:810)
    const double X[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
    int k = 0;
#pragma unroll
    for (int p = 0; p < 9; ++p) {
      const double wp = w * X[p];
#pragma unroll
      for (int q = p; q < 9; ++q) { a[k] = fma(wp, X[q], a[k]); ++k; }
    }
  }
  d1 = warp_sum(d1) / n;
  d2 = warp_sum(d2) / n;
#pragma unroll
  for (int k = 0; k < 45; ++k) a[k] = warp_sum(a[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 45; ++k) r[k] = a[k];
    r[45] = m1x; r[46] = m1y; r[47] = m2x; r[48] = m2y;
    r[49] = sqrt(2.0) / (d1 + 1e-8);  // scale (:739)
    r[50] = sqrt(2.0) / (d2 + 1e-8);
    r[51] = (double)n;
  }
}

// Same record, one CTA (4 warps) per pair: for long correspondence lists (N >= 256) a single warp per pair leaves the
// kernel latency bound (64 dependent iterations per lane at N = 2048: 8 % of HBM peak); with 128 threads per pair the
// 2 x 41 KB of a pair are in flight at once and the second pass hits L2.
template <class Pts>
__global__ void __launch_bounds__(128) eightpt_accumulate_cta_kernel(Pts pts, int P, double* __restrict__ rec) {
  __shared__ double sred[4][48];
  const int pair = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
  if (pair >= P) return;
  const int n = pts.count(pair);
  double* r = rec + (size_t)pair * kRec;
  if (n < 8) {
    if (t == 0) r[51] = (double)n;
    return;
  }
  double sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0;
  for (int i = t; i < n; i += 128) {
    float x1, y1, x2, y2, w;
    pts.load(pair, i, x1, y1, x2, y2, w);
    sx1 += x1; sy1 += y1; sx2 += x2; sy2 += y2;
  }
  sx1 = warp_sum(sx1); sy1 = warp_sum(sy1); sx2 = warp_sum(sx2); sy2 = warp_sum(sy2);
  if (lane == 0) { sred[wid][0] = sx1; sred[wid][1] = sy1; sred[wid][2] = sx2; sred[wid][3] = sy2; }
  __syncthreads();
  // fixed-order merge: every thread computes the same means
  const double m1x = (sred[0][0] + sred[1][0] + sred[2][0] + sred[3][0]) / n;
  const double m1y = (sred[0][1] + sred[1][1] + sred[2][1] + sred[3][1]) / n;
  const double m2x = (sred[0][2] + sred[1][2] + sred[2][2] + sred[3][2]) / n;
  const double m2y = (sred[0][3] + sred[1][3] + sred[2][3] + sred[3][3]) / n;
  __syncthreads();
  double d1 = 0, d2 = 0;
  double a[45];
#pragma unroll
  for (int k = 0; k < 45; ++k) a[k] = 0.0;
  for (int i = t; i < n; i += 128) {
    float fx1, fy1, fx2, fy2, fw;
    pts.load(pair, i, fx1, fy1, fx2, fy2, fw);
    const double x1 = fx1 - m1x, y1 = fy1 - m1y, x2 = fx2 - m2x, y2 = fy2 - m2y, w = fw;
    d1 += sqrt(x1 * x1 + y1 * y1);
    d2 += sqrt(x2 * x2 + y2 * y2);
    const double X[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
    int k = 0;
#pragma unroll
    for (int p = 0; p < 9; ++p) {
      const double wp = w * X[p];
#pragma unroll
      for (int q = p; q < 9; ++q) { a[k] = fma(wp, X[q], a[k]); ++k; }
    }
  }
#pragma unroll
  for (int k = 0; k < 45; ++k) {
    const double v = warp_sum(a[k]);
    if (lane == 0) sred[wid][k] = v;
  }
  d1 = warp_sum(d1); d2 = warp_sum(d2);
  if (lane == 0) { sred[wid][45] = d1; sred[wid][46] = d2; }
  __syncthreads();
  if (t < 47) {
    const double v = sred[0][t] + sred[1][t] + sred[2][t] + sred[3][t];
    if (t < 45) r[t] = v;
    else r[49 + (t - 45)] = sqrt(2.0) / (v / n + 1e-8);   // scale (:739)
  }
  if (t == 64) { r[45] = m1x; r[46] = m1y; r[47] = m2x; r[48] = m2y; r[51] = (double)n; }
}

// ------------------------------------------------------------------------------------------- stage 2
// F (row-major 3x3, fp32) for one pair from its record.  Returns false when the pair had < 8 points.
__device__ __forceinline__ bool eightpt_solve(const double* __restrict__ r, float (&F)[9]) {
  if (r[51] < 8.0) return false;
  const double s1 = r[49], s2 = r[50];
  // centred -> normalised: x~_n = D x~_c, D = diag(s2 s1, s2 s1, s2, s2 s1, s2 s1, s2, s1, s1, 1)
  const double D[9] = {s2 * s1, s2 * s1, s2, s2 * s1, s2 * s1, s2, s1, s1, 1.0};
  float A[9][9], V[9][9];
  {
    int k = 0;
#pragma unroll
    for (int p = 0; p < 9; ++p)
#pragma unroll
      for (int q = p; q < 9; ++q) {
        const float v = (float)(r[k++] * D[p] * D[q]);
        A[p][q] = v;
        A[q][p] = v;
      }
  }
  jacobi_eig<float, 9>(A, V, 10, 1e-8f);
  // smallest eigenvalue's eigenvector == V[..., -1] of the SVD of the PSD matrix (:819-821)
  int idx = 0;
  float best = A[0][0];
#pragma unroll
  for (int c = 1; c < 9; ++c)
    if (A[c][c] < best) { best = A[c][c]; idx = c; }
  double Fh[3][3];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    float v = V[k][0];
#pragma unroll
    for (int c = 1; c < 9; ++c) v = (idx == c) ? V[k][c] : v;
    Fh[k / 3][k % 3] = (double)v;
  }
  // rank-2 projection (:824-827): F_hat - (F_hat v3) v3^T with v3 = right-singular vector of the smallest sigma
  double M[3][3], Vs[3][3], lam[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) M[i][j] = Fh[0][i] * Fh[0][j] + Fh[1][i] * Fh[1][j] + Fh[2][i] * Fh[2][j];
  sym3_eig_sorted(M, Vs, lam);
  double Fp[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double fv = Fh[i][0] * Vs[0][2] + Fh[i][1] * Vs[1][2] + Fh[i][2] * Vs[2][2];
#pragma unroll
    for (int j = 0; j < 3; ++j) Fp[i][j] = Fh[i][j] - fv * Vs[j][2];
  }
  // F = T2^T Fp T1, T = [[s,0,-s mx],[0,s,-s my],[0,0,1]]   (:828)
  const double m1x = r[45], m1y = r[46], m2x = r[47], m2y = r[48];
  const double T1[3][3] = {{s1, 0, -s1 * m1x}, {0, s1, -s1 * m1y}, {0, 0, 1}};
  const double T2[3][3] = {{s2, 0, -s2 * m2x}, {0, s2, -s2 * m2y}, {0, 0, 1}};
  double G[3][3], Fe[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) G[i][j] = Fp[i][0] * T1[0][j] + Fp[i][1] * T1[1][j] + Fp[i][2] * T1[2][j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Fe[i][j] = T2[0][i] * G[0][j] + T2[1][i] * G[1][j] + T2[2][i] * G[2][j];
  // normalize_transformation (:753-769): M / (M22 + eps) where |M22| > eps
  const float f22 = (float)Fe[2][2];
  const bool nz = fabsf(f22) > 1e-8f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float v = (float)Fe[k / 3][k % 3];
    F[k] = nz ? v / (f22 + 1e-8f) : v;
  }
  return true;
}

// R1, R2 (row-major) and t from E (essential.py:99-139).  SVD via the eigen-decomposition of E^T E:
// V sorted by descending sigma, u_i = E v_i / sigma_i (i = 1,2), u3 = u1 x u2  (=> det U = +1, the state the
// reference reaches after its det(U) < 0 fix), v3 negated when det V < 0.
__device__ __forceinline__ void essential_decompose(const float (&Ef)[9], float (&R1)[9], float (&R2)[9], float (&t)[3]) {
  double E[3][3], M[3][3], V[3][3], lam[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) E[k / 3][k % 3] = (double)Ef[k];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) M[i][j] = E[0][i] * E[0][j] + E[1][i] * E[1][j] + E[2][i] * E[2][j];
  sym3_eig_sorted(M, V, lam);
  double U[3][3];
  // u1
  double n1 = 0, u1[3], u2[3], u3[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { u1[i] = E[i][0] * V[0][0] + E[i][1] * V[1][0] + E[i][2] * V[2][0]; n1 += u1[i] * u1[i]; }
  n1 = sqrt(n1);
  if (n1 > 1e-300) { u1[0] /= n1; u1[1] /= n1; u1[2] /= n1; } else { u1[0] = 1; u1[1] = 0; u1[2] = 0; }
  double n2 = 0, dp = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) u2[i] = E[i][0] * V[0][1] + E[i][1] * V[1][1] + E[i][2] * V[2][1];
  dp = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) { u2[i] -= dp * u1[i]; n2 += u2[i] * u2[i]; }
  n2 = sqrt(n2);
  if (n2 > 1e-300) { u2[0] /= n2; u2[1] /= n2; u2[2] /= n2; }
  else {  // rank <= 1: any unit vector orthogonal to u1
    const double ex = fabs(u1[0]) < 0.9 ? 1.0 : 0.0, ey = 1.0 - ex;
    const double d = ex * u1[0] + ey * u1[1];
    u2[0] = ex - d * u1[0]; u2[1] = ey - d * u1[1]; u2[2] = -d * u1[2];
    const double nn = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    u2[0] /= nn; u2[1] /= nn; u2[2] /= nn;
  }
  u3[0] = u1[1] * u2[2] - u1[2] * u2[1];
  u3[1] = u1[2] * u2[0] - u1[0] * u2[2];
  u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
#pragma unroll
  for (int i = 0; i < 3; ++i) { U[i][0] = u1[i]; U[i][1] = u2[i]; U[i][2] = u3[i]; }
  const double detV = V[0][0] * (V[1][1] * V[2][2] - V[1][2] * V[2][1]) - V[0][1] * (V[1][0] * V[2][2] - V[1][2] * V[2][0]) +
                      V[0][2] * (V[1][0] * V[2][1] - V[1][1] * V[2][0]);
  if (detV < 0) { V[0][2] = -V[0][2]; V[1][2] = -V[1][2]; V[2][2] = -V[2][2]; }
  // R1 = U W V^T, R2 = U W^T V^T, W = [[0,-1,0],[1,0,0],[0,0,1]]:  U W = [u2, -u1, u3],  U W^T = [-u2, u1, u3]
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double a = U[i][1] * V[j][0] - U[i][0] * V[j][1];
      const double c = U[i][2] * V[j][2];
      R1[i * 3 + j] = (float)(a + c);
      R2[i * 3 + j] = (float)(-a + c);
    }
  t[0] = (float)u3[0]; t[1] = (float)u3[1]; t[2] = (float)u3[2];
}

__global__ void __launch_bounds__(64) eightpt_solve_kernel(const double* __restrict__ rec, int P, float* __restrict__ Fout) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= P) return;
  float F[9];
  if (!eightpt_solve(rec + (size_t)pair * kRec, F)) {
#pragma unroll
    for (int k = 0; k < 9; ++k) F[k] = 0.f;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) Fout[(size_t)pair * 9 + k] = F[k];
}

__global__ void __launch_bounds__(128) essential_decompose_kernel(const float* __restrict__ E, int P, float* __restrict__ R1,
                                                                  float* __restrict__ R2, float* __restrict__ t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float e[9], r1[9], r2[9], tt[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) e[k] = E[(size_t)i * 9 + k];
  essential_decompose(e, r1, r2, tt);
#pragma unroll
  for (int k = 0; k < 9; ++k) { R1[(size_t)i * 9 + k] = r1[k]; R2[(size_t)i * 9 + k] = r2[k]; }
  t[(size_t)i * 3 + 0] = tt[0]; t[(size_t)i * 3 + 1] = tt[1]; t[(size_t)i * 3 + 2] = tt[2];
}

// pose_from_matches stage 2: E + the 4 candidates into cand[pair][ 2*9 + 3 ] floats (R1, R2, t)
__global__ void __launch_bounds__(64) pose_solve_kernel(const double* __restrict__ rec, int P, float* __restrict__ Eout,
                                                        float* __restrict__ cand) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= P) return;
  float F[9], r1[9], r2[9], tt[3];
  if (!eightpt_solve(rec + (size_t)pair * kRec, F)) {
    // < 8 matches: identity pose, E = I  (supervision.py:222-224 fallback)
#pragma unroll
    for (int k = 0; k < 9; ++k) { F[k] = (k % 4 == 0) ? 1.f : 0.f; r1[k] = F[k]; r2[k] = F[k]; }
    tt[0] = tt[1] = tt[2] = 0.f;
  } else {
    essential_decompose(F, r1, r2, tt);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    Eout[(size_t)pair * 9 + k] = F[k];
    cand[(size_t)pair * 21 + k] = r1[k];
    cand[(size_t)pair * 21 + 9 + k] = r2[k];
  }
  cand[(size_t)pair * 21 + 18] = tt[0]; cand[(size_t)pair * 21 + 19] = tt[1]; cand[(size_t)pair * 21 + 20] = tt[2];
}

// stage 3 (one warp per pair): cheirality votes of the 4 candidates (R1,t),(R1,-t),(R2,t),(R2,-t) -- the order of
// motion_from_essential (essential.py:60-62) -- the first candidate with the most votes wins (metrics.py:165-170
// keeps the first strictly-better recoverPose result).
__global__ void __launch_bounds__(128) pose_select_kernel(RaggedPts pts, int P, const float* __restrict__ cand,
                                                          float* __restrict__ Rt, int* __restrict__ npos,
                                                          const unsigned char* __restrict__ mask = nullptr) {
  const int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (pair >= P) return;
  const int n = pts.count(pair);
  const float* c = cand + (size_t)pair * 21;
  int votes[4] = {0, 0, 0, 0};
  if (n >= 8) {
    for (int i = lane; i < n; i += 32) {
      if (mask != nullptr && !mask[(size_t)pts.off[pair] + i]) continue;   // RANSAC round: inliers of the best model only
      float x0, y0, x1, y1, w;
      pts.load(pair, i, x0, y0, x1, y1, w);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* R = c + (k >> 1) * 9;
        const float sg = (k & 1) ? -1.f : 1.f;
        const float tx = sg * c[18], ty = sg * c[19], tz = sg * c[20];
        const float rx = R[0] * x0 + R[1] * y0 + R[2], ry = R[3] * x0 + R[4] * y0 + R[5], rz = R[6] * x0 + R[7] * y0 + R[8];
        // a = (R x0h) x x1h ; b = t x x1h ; z0 = -(a.b)/(a.a) ; X1 = z0 R x0h + t
        const float ax = ry - rz * y1, ay = rz * x1 - rx, az = rx * y1 - ry * x1;
        const float bx = ty - tz * y1, by = tz * x1 - tx, bz = tx * y1 - ty * x1;
        const float z0 = -(ax * bx + ay * by + az * bz) / fmaxf(ax * ax + ay * ay + az * az, 1e-20f);
        const float z1 = z0 * rz + tz;
        votes[k] += (z0 > 0.f && z1 > 0.f) ? 1 : 0;
      }
    }
  }
  int best = -1, bk = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int v = warp_sum(votes[k]);
    if (v > best) { best = v; bk = k; }
  }
  if (lane == 0) {
    const float* R = c + (bk >> 1) * 9;
    const float sg = (bk & 1) ? -1.f : 1.f;
    float* o = Rt + (size_t)pair * 12;
    for (int i = 0; i < 3; ++i) {
      o[i * 4 + 0] = R[i * 3 + 0]; o[i * 4 + 1] = R[i * 3 + 1]; o[i * 4 + 2] = R[i * 3 + 2];
      o[i * 4 + 3] = (n >= 8) ? sg * c[18 + i] : 0.f;
    }
    npos[pair] = (n >= 8) ? best : 0;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Prior-guided RANSAC, scoring step (mp3d_loftr/third_party/prior_ransac/ransac.py:203-231 `get_prior_estimate`,
// :256-292 `verify`, :303-308 `remove_bad_models`): one warp per (pair, hypothesis).
//   prior score = -(min_k mean |[R_k | T] p - prior_rt p|)^2 / lambda   over a fixed point cloud, (R_1, R_2, T) from
//                 decompose_essential_matrix(E)  (use_noexp_prior_scoring, :401-404)
//   inlier count = #{ i : sampson(E, x0_i, x1_i) <= inl_th }            (squared Sampson distance, K-normalised points)
// score[p,h] = count + prior, or -inf for a degenerate model (min |diag E| <= 1e-4).
__device__ __forceinline__ float sampson_sq(const float (&E)[9], float x0, float y0, float x1, float y1) {
  // l = E x0h (epipolar line in image 1), m = E^T x1h
  const float lx = E[0] * x0 + E[1] * y0 + E[2], ly = E[3] * x0 + E[4] * y0 + E[5], lz = E[6] * x0 + E[7] * y0 + E[8];
  const float mx = E[0] * x1 + E[3] * y1 + E[6], my = E[1] * x1 + E[4] * y1 + E[7];
  const float num = x1 * lx + y1 * ly + lz;
  return num * num / (lx * lx + ly * ly + mx * mx + my * my);
}

__global__ void __launch_bounds__(256) ransac_score_kernel(RaggedPts pts, int P, int H, const float* __restrict__ models,
                                                           const float* __restrict__ prior_rt,
                                                           const float* __restrict__ pcl, int npcl, float prior_lambda,
                                                           float inl_th, float* __restrict__ scores) {
  const int h = blockIdx.x * 8 + (threadIdx.x >> 5), pair = blockIdx.y, lane = threadIdx.x & 31;
  if (h >= H) return;
  const int n = pts.count(pair);
  float E[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) E[k] = models[((size_t)pair * H + h) * 9 + k];
  float* out = scores + (size_t)pair * H + h;
  if (n < 8 || !(fminf(fminf(fabsf(E[0]), fabsf(E[4])), fabsf(E[8])) > 1e-4f)) {
    if (lane == 0) *out = -INFINITY;
    return;
  }
  float prior = 0.f;
  if (prior_rt != nullptr) {
    float r1[9], r2[9], tt[3];
    essential_decompose(E, r1, r2, tt);
    const float* pr = prior_rt + (size_t)pair * 12;
    // setup_prior: unit translation (:180); a zero prior translation stays zero instead of poisoning every score with NaN
    const float tn = 1.f / fmaxf(sqrtf(pr[3] * pr[3] + pr[7] * pr[7] + pr[11] * pr[11]), 1e-12f);
    // The sign of T = U[:, 2] is an artefact of the SVD routine (E fixes t only up to sign).  The reference scores
    // whichever sign LAPACK returns (ransac.py:215-224); here both signs are scored and the better one counts, which
    // equals the reference's value whenever LAPACK's sign is the better one and is never worse.
    float e1p = 0.f, e2p = 0.f, e1m = 0.f, e2m = 0.f;
    for (int j = lane; j < npcl; j += 32) {
      const float px = pcl[j * 3], py = pcl[j * 3 + 1], pz = pcl[j * 3 + 2];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float tg = pr[k * 4] * px + pr[k * 4 + 1] * py + pr[k * 4 + 2] * pz + pr[k * 4 + 3] * tn;
        const float a1 = r1[k * 3] * px + r1[k * 3 + 1] * py + r1[k * 3 + 2] * pz - tg;
        const float a2 = r2[k * 3] * px + r2[k * 3 + 1] * py + r2[k * 3 + 2] * pz - tg;
        e1p += fabsf(a1 + tt[k]); e1m += fabsf(a1 - tt[k]);
        e2p += fabsf(a2 + tt[k]); e2m += fabsf(a2 - tt[k]);
      }
    }
    e1p = warp_sum(e1p); e2p = warp_sum(e2p); e1m = warp_sum(e1m); e2m = warp_sum(e2m);
    const float e = fminf(fminf(e1p, e2p), fminf(e1m, e2m)) / (3.f * (float)npcl);
    prior = -e * e / prior_lambda;
  }
  int cnt = 0;
  for (int i = lane; i < n; i += 32) {
    float x0, y0, x1, y1, w;
    pts.load(pair, i, x0, y0, x1, y1, w);
    cnt += sampson_sq(E, x0, y0, x1, y1) <= inl_th ? 1 : 0;
  }
  cnt = warp_sum(cnt);
  if (lane == 0) *out = (float)cnt + prior;
}

// argmax over the hypotheses of a pair (lowest index among equal scores, like torch.argmax on CPU), then the inlier
// masks / counts of the winner at inl_th, inl_th/10, inl_th/100 (:279-283).  One CTA per pair.
__global__ void __launch_bounds__(256) ransac_select_kernel(RaggedPts pts, int P, int H, const float* __restrict__ models,
                                                            const float* __restrict__ scores, float inl_th,
                                                            int* __restrict__ best_idx, float* __restrict__ best_E,
                                                            int* __restrict__ counts3, unsigned char* __restrict__ mask) {
  __shared__ float sv[256];
  __shared__ int si[256];
  __shared__ int sc[3];
  const int pair = blockIdx.x, t = threadIdx.x;
  const int n = pts.count(pair);
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int h = t; h < H; h += 256) {
    const float v = scores[(size_t)pair * H + h];
    if (v > bv) { bv = v; bi = h; }   // ascending h per thread: keeps the lowest index
  }
  sv[t] = bv; si[t] = bi;
  if (t < 3) sc[t] = 0;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) {
      const float v = sv[t + o];
      const int i = si[t + o];
      if (v > sv[t] || (v == sv[t] && i < si[t])) { sv[t] = v; si[t] = i; }
    }
    __syncthreads();
  }
  const bool ok = sv[0] > -INFINITY && n >= 8;
  const int best = ok ? si[0] : -1;
  float E[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) E[k] = ok ? models[((size_t)pair * H + best) * 9 + k] : 0.f;
  int c0 = 0, c1 = 0, c2 = 0;
  for (int i = t; i < n; i += 256) {
    unsigned char m = 0;
    if (ok) {
      float x0, y0, x1, y1, w;
      pts.load(pair, i, x0, y0, x1, y1, w);
      const float e = sampson_sq(E, x0, y0, x1, y1);
      const int t1 = e <= inl_th / 10.0f ? 1 : 0, t2 = e <= inl_th / 100.0f ? 1 : 0;
      m = e <= inl_th ? 1 : 0;
      c0 += m;
      c1 += t1;
      c2 += t2;
      m |= (unsigned char)((t1 << 1) | (t2 << 2));   // bit 0: inlier, bit 1: tight, bit 2: ultra tight (:279-283)
    }
    mask[(size_t)pts.off[pair] + i] = m;
  }
  c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
  if ((t & 31) == 0) { atomicAdd(&sc[0], c0); atomicAdd(&sc[1], c1); atomicAdd(&sc[2], c2); }
  __syncthreads();
  if (t == 0) {
    best_idx[pair] = best;
    counts3[pair * 3 + 0] = sc[0]; counts3[pair * 3 + 1] = sc[1]; counts3[pair * 3 + 2] = sc[2];
  }
  if (t < 9) best_E[(size_t)pair * 9 + t] = E[t];
}

// E -> the 4 motion candidates in the cand[21] layout of pose_select_kernel; E == 0 (no valid model) -> identity.
__global__ void __launch_bounds__(64) essential_to_cand_kernel(const float* __restrict__ E, int P, float* __restrict__ cand) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= P) return;
  float e[9], r1[9], r2[9], tt[3];
  float nrm = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { e[k] = E[(size_t)pair * 9 + k]; nrm += e[k] * e[k]; }
  if (nrm > 0.f) {
    essential_decompose(e, r1, r2, tt);
  } else {
#pragma unroll
    for (int k = 0; k < 9; ++k) { r1[k] = (k % 4 == 0) ? 1.f : 0.f; r2[k] = r1[k]; }
    tt[0] = tt[1] = tt[2] = 0.f;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) { cand[(size_t)pair * 21 + k] = r1[k]; cand[(size_t)pair * 21 + 9 + k] = r2[k]; }
  cand[(size_t)pair * 21 + 18] = tt[0]; cand[(size_t)pair * 21 + 19] = tt[1]; cand[(size_t)pair * 21 + 20] = tt[2];
}


// ------------------------------------------------------------------------------------------------------------------
// Prior-guided RANSAC, sampling step + minimal solver (ransac.py:161-175 `sample`, :358-367 bias weights, :250-253
// `estimate_model_from_minsample`), all on the device: no eager-torch math, no host sync, no [P, max_m] scatter.
//   segment offsets      m_bids (sorted) -> offsets[P+1]                                      segment_offsets_kernel
//   weights + CDF        w_i = exp(-symmetrical_epipolar(x0_i, x1_i, E_prior)/sigma^2) + 1e-4  (1 without a prior),
//                        inclusive prefix sum per pair in fp64                                 ransac_cdf_kernel
//   sample + 8-point     one thread per (pair, hypothesis): counter-based Philox4x32-10 keyed by (seed; pair, hyp),
//                        inverse-CDF draws, the S sampled correspondences -> the 52-double moment record of
//                        eightpt_accumulate (so eightpt_solve_kernel finishes the model)         ransac_sample_kernel
// The reference draws with numpy's global RNG (np.random.choice(..., replace=True, p=w), or rand().topk without a
// prior); bit-parity with that stream is meaningless, so the contract is: index k of hypothesis (p, h) is
// searchsorted(cdf_p, u * cdf_p[-1], right) with u = philox(seed, p*H+h)[k] * 2^-32 -- reproducible, pinned by a numpy
// Philox in the tests.  A draw that repeats an index already in the sample is redrawn (up to 4 times, next counter
// block): duplicated correspondences make the 8-point system rank deficient.
struct Philox {
  uint32_t key0, key1;
  __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) const {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ __forceinline__ void block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) const {
    uint32_t c[4] = {c0, c1, c2, c3};
    uint32_t k0 = key0, k1 = key1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = c[i];
  }
};

__global__ void segment_offsets_kernel(const long long* __restrict__ m_bids, long long M, int P,
                                       long long* __restrict__ offsets) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > M) return;
  // offsets[b] = first index whose batch id is >= b; thread i fills every b in (bid[i-1], bid[i]]
  const long long lo = (i == 0) ? -1 : m_bids[i - 1];
  const long long hi = (i == M) ? (long long)P : m_bids[i];
  for (long long b = lo + 1; b <= hi && b <= P; ++b) offsets[b] = i;
}

__device__ __forceinline__ float sym_epipolar_sq(const float (&E)[9], float x0, float y0, float x1, float y1) {
  const float lx = E[0] * x0 + E[1] * y0 + E[2], ly = E[3] * x0 + E[4] * y0 + E[5], lz = E[6] * x0 + E[7] * y0 + E[8];
  const float mx = E[0] * x1 + E[3] * y1 + E[6], my = E[1] * x1 + E[4] * y1 + E[7];
  const float num = x1 * lx + y1 * ly + lz;
  return num * num * (1.f / (lx * lx + ly * ly) + 1.f / (mx * mx + my * my));
}

// one CTA (256 threads) per pair
__global__ void __launch_bounds__(256) ransac_cdf_kernel(RaggedPts pts, int P, const float* __restrict__ prior_rt,
                                                         float sigma_sq, double* __restrict__ cdf) {
  __shared__ double warp_tot[8];
  __shared__ double carry_s;
  const int pair = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int n = pts.count(pair);
  double* out = cdf + pts.off[pair];
  float E[9];
  const bool biased = prior_rt != nullptr;
  if (biased) {  // E_prior = [t]_x R with the unit-norm translation of setup_prior (ransac.py:63-71,180)
    const float* pr = prior_rt + (size_t)pair * 12;
    const float tn = 1.f / fmaxf(sqrtf(pr[3] * pr[3] + pr[7] * pr[7] + pr[11] * pr[11]), 1e-12f);
    const float tx = pr[3] * tn, ty = pr[7] * tn, tz = pr[11] * tn;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      E[j] = -tz * pr[4 + j] + ty * pr[8 + j];
      E[3 + j] = tz * pr[j] - tx * pr[8 + j];
      E[6 + j] = -ty * pr[j] + tx * pr[4 + j];
    }
  }
  if (t == 0) carry_s = 0.0;
  __syncthreads();
  for (int base = 0; base < n; base += 256) {
    const int i = base + t;
    double w = 0.0;
    if (i < n) {
      if (biased) {
        float x0, y0, x1, y1, cf;
        pts.load(pair, i, x0, y0, x1, y1, cf);
        w = (double)(expf(-sym_epipolar_sq(E, x0, y0, x1, y1) / sigma_sq) + 1e-4f);
        if (!(w == w)) w = 1e-4;  // degenerate prior (zero lines): keep a valid distribution
      } else {
        w = 1.0;
      }
    }
    double v = w;  // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    double pre = carry_s;
    for (int k = 0; k < wid; ++k) pre += warp_tot[k];
    if (i < n) out[i] = pre + v;
    __syncthreads();
    if (t == 255) carry_s = pre + v;
    __syncthreads();
  }
}

// S distinct indices of a pair's n correspondences for hypothesis g: inverse-CDF draws from the counter-based stream
// (seed; g, block), a draw that repeats an index already in the sample is redrawn (up to 4 times, next words).
template <int S>
__device__ __forceinline__ void draw_sample(const double* __restrict__ c, int n, unsigned long long seed, int g, int (&idx)[S]) {
  const double total = c[n - 1];
  Philox rng{(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t u4[4];
  int blk = 0, used = 4;
  for (int k = 0; k < S; ++k) {
    int pick = 0;
    for (int attempt = 0; attempt < 5; ++attempt) {
      if (used == 4) { rng.block((uint32_t)g, (uint32_t)blk, 0u, 0u, u4); ++blk; used = 0; }
      const double target = ((double)u4[used++] * (1.0 / 4294967296.0)) * total;
      int lo = 0, hi = n - 1;  // first i with cdf[i] > target  (searchsorted side='right')
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (c[mid] > target) hi = mid; else lo = mid + 1;
      }
      pick = lo;
      bool dup = false;
      for (int q = 0; q < k; ++q) dup |= (idx[q] == pick);
      if (!dup) break;
    }
    idx[k] = pick;
  }
}

template <int S>
__global__ void __launch_bounds__(128) ransac_sample_kernel(RaggedPts pts, int P, int H, const double* __restrict__ cdf,
                                                            unsigned long long seed, double* __restrict__ rec,
                                                            int* __restrict__ idx_out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P * H) return;
  const int pair = g / H;
  const int n = pts.count(pair);
  double* r = rec + (size_t)g * kRec;
  if (n < S) {
    r[51] = 0.0;
    if (idx_out)
      for (int k = 0; k < S; ++k) idx_out[(size_t)g * S + k] = -1;
    return;
  }
  int idx[S];
  draw_sample<S>(cdf + pts.off[pair], n, seed, g, idx);
  if (idx_out)
    for (int k = 0; k < S; ++k) idx_out[(size_t)g * S + k] = idx[k];
  // the moment record of eightpt_accumulate_kernel for these S correspondences, unit weights (:253 passes ones)
  float X1[S], Y1[S], X2[S], Y2[S];
  double m1x = 0, m1y = 0, m2x = 0, m2y = 0;
#pragma unroll
  for (int k = 0; k < S; ++k) {
    float w;
    pts.load(pair, idx[k], X1[k], Y1[k], X2[k], Y2[k], w);
    m1x += X1[k]; m1y += Y1[k]; m2x += X2[k]; m2y += Y2[k];
  }
  m1x /= S; m1y /= S; m2x /= S; m2y /= S;
  double a[45];
#pragma unroll
  for (int k = 0; k < 45; ++k) a[k] = 0.0;
  double d1 = 0, d2 = 0;
#pragma unroll 1
  for (int k = 0; k < S; ++k) {
    const double x1 = X1[k] - m1x, y1 = Y1[k] - m1y, x2 = X2[k] - m2x, y2 = Y2[k] - m2y;
    d1 += sqrt(x1 * x1 + y1 * y1);
    d2 += sqrt(x2 * x2 + y2 * y2);
    const double X[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
    int e = 0;
#pragma unroll
    for (int p = 0; p < 9; ++p)
#pragma unroll
      for (int q = p; q < 9; ++q) { a[e] = fma(X[p], X[q], a[e]); ++e; }
  }
#pragma unroll
  for (int k = 0; k < 45; ++k) r[k] = a[k];
  r[45] = m1x; r[46] = m1y; r[47] = m2x; r[48] = m2y;
  r[49] = sqrt(2.0) / (d1 / S + 1e-8);
  r[50] = sqrt(2.0) / (d2 / S + 1e-8);
  r[51] = (double)S;
}


// The recipe's model type (`essential_cv2`: a 5-point solver on 6 sampled correspondences, ransac.py:250-253 with
// cv_geometry.py:836-859 / the in-tree batched 5-point :861-1041): one thread per (pair, hypothesis) draws 6 distinct
// correspondences, solves Nister's 5-point on the first five (fivept.cuh, up to 10 essential matrices) and keeps the
// candidate with the smallest squared Sampson distance at the sixth -- one model per hypothesis, so the scoring step is
// the same as for the 8-point models.  Writes E row-major (unit Frobenius norm), zeros when there is no solution.
__global__ void __launch_bounds__(64) ransac_sample5_kernel(RaggedPts pts, int P, int H, const double* __restrict__ cdf,
                                                            unsigned long long seed, float* __restrict__ models,
                                                            int* __restrict__ idx_out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P * H) return;
  const int pair = g / H;
  const int n = pts.count(pair);
  float* out = models + (size_t)g * 9;
  if (n < 8) {   // the scoring step treats pairs with fewer than 8 correspondences as unsolved (same bound as the 8-point mode)
#pragma unroll
    for (int k = 0; k < 9; ++k) out[k] = 0.f;
    if (idx_out)
      for (int k = 0; k < 6; ++k) idx_out[(size_t)g * 6 + k] = -1;
    return;
  }
  int idx[6];
  draw_sample<6>(cdf + pts.off[pair], n, seed, g, idx);
  if (idx_out)
    for (int k = 0; k < 6; ++k) idx_out[(size_t)g * 6 + k] = idx[k];
  double X1[6], Y1[6], X2[6], Y2[6];
  for (int k = 0; k < 6; ++k) {
    float a, b, c, d, w;
    pts.load(pair, idx[k], a, b, c, d, w);
    X1[k] = a; Y1[k] = b; X2[k] = c; Y2[k] = d;
  }
  double Es[10][9];
  const int ns = fivept::solve(X1, Y1, X2, Y2, Es);
  int best = -1;
  double best_err = 1e300;
  for (int q = 0; q < ns; ++q) {
    const double* E = Es[q];
    const double lx = E[0] * X1[5] + E[1] * Y1[5] + E[2], ly = E[3] * X1[5] + E[4] * Y1[5] + E[5],
                 lz = E[6] * X1[5] + E[7] * Y1[5] + E[8];
    const double mx = E[0] * X2[5] + E[3] * Y2[5] + E[6], my = E[1] * X2[5] + E[4] * Y2[5] + E[7];
    const double num = X2[5] * lx + Y2[5] * ly + lz;
    const double err = num * num / (lx * lx + ly * ly + mx * mx + my * my);
    if (err < best_err) { best_err = err; best = q; }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) out[k] = best >= 0 ? (float)Es[best][k] : 0.f;
}

// diagnostics / tests: all solutions of the 5-point solver for explicit minimal samples.  pts [S,5,4] = (x1, y1, x2, y2)
// calibrated, fp64; E [S,10,9]; nsol [S]
__global__ void fivept_solve_kernel(const double* __restrict__ pts5, int S, double* __restrict__ E, int* __restrict__ nsol) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double x1[5], y1[5], x2[5], y2[5];
  for (int k = 0; k < 5; ++k) {
    x1[k] = pts5[(s * 5 + k) * 4]; y1[k] = pts5[(s * 5 + k) * 4 + 1];
    x2[k] = pts5[(s * 5 + k) * 4 + 2]; y2[k] = pts5[(s * 5 + k) * 4 + 3];
  }
  double Es[10][9];
  const int ns = fivept::solve(x1, y1, x2, y2, Es);
  nsol[s] = ns;
  for (int q = 0; q < 10; ++q)
    for (int k = 0; k < 9; ++k) E[((size_t)s * 10 + q) * 9 + k] = q < ns ? Es[q][k] : 0.0;
}
}  // namespace far

using namespace far;

extern "C" size_t far_eight_point_workspace_bytes(int P) { return (size_t)P * (kRec * 8 + 21 * 4) + 256; }

extern "C" int far_eight_point(const float* pts1, const float* pts2, const float* weights, const int* counts, int P,
                               int N, float* F, float* workspace, size_t workspace_bytes, void* stream) {
  if (P <= 0) return FAR_OK;
  FAR_REQUIRE(pts1 && pts2 && F && workspace && N > 0);
  if (workspace_bytes < (size_t)P * kRec * 8) return FAR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  double* rec = reinterpret_cast<double*>(workspace);
  DensePts pts{pts1, pts2, weights, counts, N};
  // algorithmic bytes: 20 N in (two point sets + weight) + 36 out per pair; ~120 N FLOP
  ProfScope prof(PROF_EIGHTPT, 120.0 * N * (double)P, ((weights ? 20.0 : 16.0) * N + 36.0) * (double)P, st);
  if (N >= 256) eightpt_accumulate_cta_kernel<DensePts><<<P, 128, 0, st>>>(pts, P, rec);
  else eightpt_accumulate_kernel<DensePts><<<ceil_div(P, 4), 128, 0, st>>>(pts, P, rec);
  FAR_CHECK_LAUNCH();
  eightpt_solve_kernel<<<ceil_div(P, 64), 64, 0, st>>>(rec, P, F);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

extern "C" int far_essential_decompose(const float* E, int P, float* R1, float* R2, float* t, void* stream) {
  if (P <= 0) return FAR_OK;
  FAR_REQUIRE(E && R1 && R2 && t);
  essential_decompose_kernel<<<ceil_div(P, 128), 128, 0, (cudaStream_t)stream>>>(E, P, R1, R2, t);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

extern "C" int far_pose_from_matches(const float* mkpts0, const float* mkpts1, const float* mconf,
                                     const long long* offsets, int N, const float* K0, const float* K1, float* E,
                                     float* Rt, int* n_pos, float* workspace, size_t workspace_bytes, void* stream) {
  if (N <= 0) return FAR_OK;
  FAR_REQUIRE(offsets && K0 && K1 && E && Rt && n_pos && workspace);
  if (workspace_bytes < (size_t)N * (kRec * 8 + 21 * 4)) return FAR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  double* rec = reinterpret_cast<double*>(workspace);
  float* cand = reinterpret_cast<float*>(rec + (size_t)N * kRec);
  RaggedPts pts{mkpts0, mkpts1, mconf, offsets, K0, K1};
  eightpt_accumulate_kernel<RaggedPts><<<ceil_div(N, 4), 128, 0, st>>>(pts, N, rec);
  FAR_CHECK_LAUNCH();
  pose_solve_kernel<<<ceil_div(N, 64), 64, 0, st>>>(rec, N, E, cand);
  FAR_CHECK_LAUNCH();
  pose_select_kernel<<<ceil_div(N, 4), 128, 0, st>>>(pts, N, cand, Rt, n_pos);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

// ---- prior-guided RANSAC round: scoring of H hypotheses per pair + pose of the winner ----------
extern "C" int far_prior_ransac_score(const float* mkpts0, const float* mkpts1, const long long* offsets, int P,
                                      const float* K0, const float* K1, const float* models, int H,
                                      const float* prior_rt, const float* pcl, int npcl, float prior_lambda,
                                      float inl_th, float* scores, int* best_idx, float* best_E, int* counts3,
                                      unsigned char* inlier_mask, void* stream) {
  if (P <= 0 || H <= 0) return FAR_OK;
  FAR_REQUIRE(mkpts0 && mkpts1 && offsets && K0 && K1 && models && scores && best_idx && best_E && counts3 &&
              inlier_mask && (prior_rt == nullptr || (pcl != nullptr && npcl > 0 && prior_lambda > 0.f)));
  cudaStream_t st = (cudaStream_t)stream;
  RaggedPts pts{mkpts0, mkpts1, nullptr, offsets, K0, K1};
  {
    ProfScope prof(PROF_SOLVER, 0.0, 0.0, st);
    ransac_score_kernel<<<dim3(ceil_div(H, 8), P), 256, 0, st>>>(pts, P, H, models, prior_rt, pcl, npcl, prior_lambda,
                                                                 inl_th, scores);
  }
  FAR_CHECK_LAUNCH();
  ransac_select_kernel<<<P, 256, 0, st>>>(pts, P, H, models, scores, inl_th, best_idx, best_E, counts3, inlier_mask);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

extern "C" size_t far_pose_from_essential_workspace_bytes(int P) { return (size_t)P * 21 * 4 + 256; }

// (R | t) of an essential matrix by the cheirality vote of pose_from_matches, restricted to `mask` (NULL: all matches).
extern "C" int far_pose_from_essential(const float* mkpts0, const float* mkpts1, const unsigned char* mask,
                                       const long long* offsets, int P, const float* K0, const float* K1,
                                       const float* E, float* Rt, int* n_pos, float* workspace, size_t workspace_bytes,
                                       void* stream) {
  if (P <= 0) return FAR_OK;
  FAR_REQUIRE(mkpts0 && mkpts1 && offsets && K0 && K1 && E && Rt && n_pos && workspace);
  if (workspace_bytes < (size_t)P * 21 * 4) return FAR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  RaggedPts pts{mkpts0, mkpts1, nullptr, offsets, K0, K1};
  essential_to_cand_kernel<<<ceil_div(P, 64), 64, 0, st>>>(E, P, workspace);
  FAR_CHECK_LAUNCH();
  pose_select_kernel<<<ceil_div(P, 4), 128, 0, st>>>(pts, P, workspace, Rt, n_pos, mask);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

// ---- prior-guided RANSAC round, sampling + minimal solver ---------------------------------------------------------
extern "C" int far_segment_offsets(const long long* m_bids, long long M, int P, long long* offsets, void* stream) {
  FAR_REQUIRE(offsets && P >= 0 && M >= 0 && (M == 0 || m_bids));
  const int blocks = (int)((M + 1 + 255) / 256);
  segment_offsets_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(m_bids, M, P, offsets);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

extern "C" size_t far_ransac_sample_models_workspace_bytes(long long M, int P, int H) {
  return (size_t)(M + 64) * 8 + (size_t)P * H * kRec * 8 + 512;
}

extern "C" int far_ransac_sample_models(const float* mkpts0, const float* mkpts1, const long long* offsets, long long M,
                                        int P, const float* K0, const float* K1, const float* prior_rt,
                                        float bias_sigma_sq, int H, int sample_size, unsigned long long seed,
                                        float* models, int* sample_idx, float* workspace, size_t workspace_bytes,
                                        void* stream) {
  if (P <= 0 || H <= 0) return FAR_OK;
  FAR_REQUIRE(offsets && K0 && K1 && models && workspace && (sample_size == 8 || sample_size == 6) &&
              (M == 0 || (mkpts0 && mkpts1)) && (prior_rt == nullptr || bias_sigma_sq > 0.f));
  if (workspace_bytes < far_ransac_sample_models_workspace_bytes(M, P, H)) return FAR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  double* cdf = reinterpret_cast<double*>(base);
  double* rec = cdf + (((size_t)M + 1 + 31) & ~size_t(31));
  RaggedPts pts{mkpts0, mkpts1, nullptr, offsets, K0, K1};
  ProfScope prof(PROF_SOLVER, 0.0, 0.0, st);
  ransac_cdf_kernel<<<P, 256, 0, st>>>(pts, P, prior_rt, bias_sigma_sq, cdf);
  FAR_CHECK_LAUNCH();
  if (sample_size == 6) {   // 5-point minimal solver + one disambiguating correspondence
    ransac_sample5_kernel<<<ceil_div(P * H, 64), 64, 0, st>>>(pts, P, H, cdf, seed, models, sample_idx);
    FAR_CHECK_LAUNCH();
    return FAR_OK;
  }
  ransac_sample_kernel<8><<<ceil_div(P * H, 128), 128, 0, st>>>(pts, P, H, cdf, seed, rec, sample_idx);
  FAR_CHECK_LAUNCH();
  eightpt_solve_kernel<<<ceil_div(P * H, 64), 64, 0, st>>>(rec, P * H, models);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}

extern "C" int far_five_point(const double* pts5, int S, double* E, int* nsol, void* stream) {
  if (S <= 0) return FAR_OK;
  FAR_REQUIRE(pts5 && E && nsol);
  fivept_solve_kernel<<<ceil_div(S, 64), 64, 0, (cudaStream_t)stream>>>(pts5, S, E, nsol);
  FAR_CHECK_LAUNCH();
  return FAR_OK;
}
