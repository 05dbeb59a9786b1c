// Nister's 5-point essential-matrix solver, one thread per minimal sample, fp64, register / local-memory resident.
// Restates third_party/prior_ransac/cv_geometry.py:861-1041 `run_5point_our_kornia` (the in-tree batched 5-point; its
// polynomial helpers come from un-vendored kornia 0.7.1 and are restated in oracle/far_oracle.py:_mul_deg_one /
// _mul_deg_two_one):
//   1. null space of the 5 epipolar equations  x2^T E x1 = 0            (:886-896; here Gauss-Jordan with full pivoting +
//      modified Gram-Schmidt instead of an SVD of X^T X: any basis of the same 4-D space gives the same solutions)
//   2. E(x,y,z) = x N0 + y N1 + z N2 + N3; the ten cubic constraints  det E = 0,  E E^T E - 1/2 tr(E E^T) E = 0  as a
//      10 x 20 matrix over Nister's monomial order                        (:901-945)
//   3. Gauss-Jordan on the first ten monomials                            (:952-956)
//   4. the 3 x 3 polynomial matrix A(z) from rows 4..9 (k = e - z f, ...) (:958-969)
//   5. det A(z): degree-10 polynomial; its real roots                     (:971-984: companion-matrix eigenvalues there;
//      Aberth-Ehrlich iteration on all ten complex roots here, then a Newton polish of the real ones)
//   6. x, y from the 3 x 2 system at each root (least squares), E normalised to unit Frobenius norm   (:1008-1022)
#pragma once
#include <cuda_runtime.h>

namespace far {
namespace fivept {

// product of two linear forms in (x, y, z, 1) -> [x^2, xy, xz, x, y^2, yz, y, z^2, z, 1]
__device__ __forceinline__ void mul11(const double* a, const double* b, double* o) {
  o[0] = a[0] * b[0]; o[1] = a[0] * b[1] + a[1] * b[0]; o[2] = a[0] * b[2] + a[2] * b[0]; o[3] = a[0] * b[3] + a[3] * b[0];
  o[4] = a[1] * b[1]; o[5] = a[1] * b[2] + a[2] * b[1]; o[6] = a[1] * b[3] + a[3] * b[1]; o[7] = a[2] * b[2];
  o[8] = a[2] * b[3] + a[3] * b[2]; o[9] = a[3] * b[3];
}
// o += s * (degree-2 poly a) * (linear form b) over [x^3, y^3, x^2y, xy^2, x^2z, x^2, y^2z, y^2, xyz, xy | xz^2, xz, x,
// yz^2, yz, y, z^3, z^2, z, 1]
__device__ __forceinline__ void mul21_acc(const double* a, const double* b, double s, double* o) {
  o[0] += s * (a[0] * b[0]);
  o[1] += s * (a[4] * b[1]);
  o[2] += s * (a[0] * b[1] + a[1] * b[0]);
  o[3] += s * (a[1] * b[1] + a[4] * b[0]);
  o[4] += s * (a[0] * b[2] + a[2] * b[0]);
  o[5] += s * (a[0] * b[3] + a[3] * b[0]);
  o[6] += s * (a[4] * b[2] + a[5] * b[1]);
  o[7] += s * (a[4] * b[3] + a[6] * b[1]);
  o[8] += s * (a[1] * b[2] + a[2] * b[1] + a[5] * b[0]);
  o[9] += s * (a[1] * b[3] + a[3] * b[1] + a[6] * b[0]);
  o[10] += s * (a[2] * b[2] + a[7] * b[0]);
  o[11] += s * (a[2] * b[3] + a[3] * b[2] + a[8] * b[0]);
  o[12] += s * (a[3] * b[3] + a[9] * b[0]);
  o[13] += s * (a[5] * b[2] + a[7] * b[1]);
  o[14] += s * (a[5] * b[3] + a[6] * b[2] + a[8] * b[1]);
  o[15] += s * (a[6] * b[3] + a[9] * b[1]);
  o[16] += s * (a[7] * b[2]);
  o[17] += s * (a[7] * b[3] + a[8] * b[2]);
  o[18] += s * (a[8] * b[3] + a[9] * b[2]);
  o[19] += s * (a[9] * b[3]);
}

// c (degree da, lowest power first) * d (degree db) -> out (degree da + db), out must not alias
__device__ __forceinline__ void polymul(const double* c, int da, const double* d, int db, double* out) {
  for (int i = 0; i <= da + db; ++i) out[i] = 0.0;
  for (int i = 0; i <= da; ++i)
    for (int j = 0; j <= db; ++j) out[i + j] = fma(c[i], d[j], out[i + j]);
}

// Real roots of  sum_k c[k] z^k  (degree 10).  Returns the count; roots[] ascending is NOT guaranteed.
__device__ inline int real_roots_deg10(const double* c, double* roots) {
  int deg = 10;
  while (deg > 0 && c[deg] == 0.0) --deg;
  if (deg < 1) return 0;
  double cmax = 0.0;
  for (int k = 0; k <= deg; ++k) cmax = fmax(cmax, fabs(c[k]));
  if (!(cmax > 0.0) || !isfinite(cmax)) return 0;
  // starting circle: the geometric mean of the root moduli |c0 / c_deg|^(1/deg), clamped
  double r0 = pow(fabs(c[0] / c[deg]), 1.0 / deg);
  if (!isfinite(r0) || r0 < 1e-3) r0 = 1e-3;
  if (r0 > 1e3) r0 = 1e3;
  double zr[10], zi[10];
  for (int k = 0; k < deg; ++k) {  // slightly irregular so no initial symmetry survives
    const double ang = 6.283185307179586 * (k + 0.35) / deg + 0.4;
    zr[k] = r0 * cos(ang) * (1.0 + 0.02 * k);
    zi[k] = r0 * sin(ang) * (1.0 + 0.02 * k);
  }
  unsigned live = (1u << deg) - 1u;   // roots still moving: a converged root is frozen (its update is skipped)
  for (int it = 0; it < 64 && live; ++it) {   // typically 15-25 iterations; solve()'s Gauss-Newton polish restores the last digits
    for (int k = 0; k < deg; ++k) {
      if (!(live >> k & 1u)) continue;
      // Horner for p and p' at z_k
      double pr = c[deg], pi = 0.0, dr = 0.0, di = 0.0;
      for (int j = deg - 1; j >= 0; --j) {
        const double ndr = dr * zr[k] - di * zi[k] + pr, ndi = dr * zi[k] + di * zr[k] + pi;
        dr = ndr; di = ndi;
        const double npr = pr * zr[k] - pi * zi[k] + c[j], npi = pr * zi[k] + pi * zr[k];
        pr = npr; pi = npi;
      }
      const double dn = dr * dr + di * di;
      if (!(dn > 0.0)) continue;
      // w = p / p'
      const double idn = 1.0 / dn;
      const double wr = (pr * dr + pi * di) * idn, wi = (pi * dr - pr * di) * idn;
      // s = sum_{j != k} 1 / (z_k - z_j)
      double sr = 0.0, si = 0.0;
      for (int j = 0; j < deg; ++j) {
        if (j == k) continue;
        const double er = zr[k] - zr[j], ei = zi[k] - zi[j];
        const double en = er * er + ei * ei;
        if (en > 0.0) { const double ien = 1.0 / en; sr = fma(er, ien, sr); si = fma(-ei, ien, si); }
      }
      // step = w / (1 - w s)
      const double qr = 1.0 - (wr * sr - wi * si), qi = -(wr * si + wi * sr);
      const double qn = qr * qr + qi * qi;
      if (!(qn > 0.0)) continue;
      const double iqn = 1.0 / qn;
      const double stepr = (wr * qr + wi * qi) * iqn, stepi = (wi * qr - wr * qi) * iqn;
      zr[k] -= stepr; zi[k] -= stepi;
      if ((fabs(stepr) + fabs(stepi)) < 1e-13 * (1.0 + fabs(zr[k]) + fabs(zi[k]))) live &= ~(1u << k);
    }
  }
  int n = 0;
  for (int k = 0; k < deg; ++k) {
    // the reference keeps a companion-matrix eigenvalue as a real root by the same relative test (oracle: 1e-9)
    if (!(fabs(zi[k]) < 1e-9 * (1.0 + fabs(zr[k])))) continue;
    double x = zr[k];
    for (int it = 0; it < 8; ++it) {   // Newton polish on the real axis
      double pv = c[deg], dv = 0.0;
      for (int j = deg - 1; j >= 0; --j) { dv = dv * x + pv; pv = pv * x + c[j]; }
      if (dv == 0.0 || !isfinite(pv / dv)) break;
      const double step = pv / dv;
      if (fabs(step) > 1e-6 * (1.0 + fabs(x))) break;   // not a simple root nearby: keep the Aberth value
      x -= step;
      if (fabs(step) < 1e-16 * (1.0 + fabs(x))) break;
    }
    if (isfinite(x)) roots[n++] = x;
  }
  return n;
}

// Essential matrices (row-major, x2h^T E x1h = 0, unit Frobenius norm) through five correspondences.
// x1, y1: image-0 points, x2, y2: image-1 points (calibrated).  Writes at most 10 solutions, returns their number.
__device__ inline int solve(const double* x1, const double* y1, const double* x2, const double* y2, double (*Eout)[9]) {
  // ---- 1. null space
  double X[5][9];
  for (int i = 0; i < 5; ++i) {
    X[i][0] = x2[i] * x1[i]; X[i][1] = x2[i] * y1[i]; X[i][2] = x2[i];
    X[i][3] = y2[i] * x1[i]; X[i][4] = y2[i] * y1[i]; X[i][5] = y2[i];
    X[i][6] = x1[i]; X[i][7] = y1[i]; X[i][8] = 1.0;
  }
  int perm[9];
  for (int j = 0; j < 9; ++j) perm[j] = j;
  for (int r = 0; r < 5; ++r) {
    int pr = r, pc = r;
    double best = -1.0;
    for (int i = r; i < 5; ++i)
      for (int j = r; j < 9; ++j)
        if (fabs(X[i][j]) > best) { best = fabs(X[i][j]); pr = i; pc = j; }
    if (!(best > 1e-300)) return 0;   // rank-deficient sample
    for (int j = 0; j < 9; ++j) { const double t = X[r][j]; X[r][j] = X[pr][j]; X[pr][j] = t; }
    for (int i = 0; i < 5; ++i) { const double t = X[i][r]; X[i][r] = X[i][pc]; X[i][pc] = t; }
    { const int t = perm[r]; perm[r] = perm[pc]; perm[pc] = t; }
    const double inv = 1.0 / X[r][r];
    for (int j = r; j < 9; ++j) X[r][j] *= inv;
    for (int i = 0; i < 5; ++i) {
      if (i == r) continue;
      const double f = X[i][r];
      if (f != 0.0)
        for (int j = r; j < 9; ++j) X[i][j] = fma(-f, X[r][j], X[i][j]);
    }
  }
  double N[4][9];
  for (int k = 0; k < 4; ++k) {
    for (int j = 0; j < 9; ++j) N[k][j] = 0.0;
    N[k][perm[5 + k]] = 1.0;
    for (int i = 0; i < 5; ++i) N[k][perm[i]] = -X[i][5 + k];
  }
  for (int k = 0; k < 4; ++k) {   // modified Gram-Schmidt: orthonormal rows keep the constraint matrix well scaled
    for (int q = 0; q < k; ++q) {
      double d = 0.0;
      for (int j = 0; j < 9; ++j) d = fma(N[k][j], N[q][j], d);
      for (int j = 0; j < 9; ++j) N[k][j] = fma(-d, N[q][j], N[k][j]);
    }
    double nn = 0.0;
    for (int j = 0; j < 9; ++j) nn = fma(N[k][j], N[k][j], nn);
    if (!(nn > 1e-300)) return 0;
    const double inv = rsqrt(nn);
    for (int j = 0; j < 9; ++j) N[k][j] *= inv;
  }
  // ---- 2. constraints.  Entry (i, j) of E as a linear form in (x, y, z, 1): e[i][j][0..3]
  double e[3][3][4];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 4; ++k) e[i][j][k] = N[k][3 * i + j];
  double C[10][20];
  for (int r = 0; r < 10; ++r)
    for (int j = 0; j < 20; ++j) C[r][j] = 0.0;
  {
    double t1[10], t2[10];
    // det E = sum over the first column cofactors (row 9, :908-924)
    mul11(e[0][1], e[1][2], t1); mul11(e[0][2], e[1][1], t2);
    for (int k = 0; k < 10; ++k) t1[k] -= t2[k];
    mul21_acc(t1, e[2][0], 1.0, C[9]);
    mul11(e[0][2], e[1][0], t1); mul11(e[0][0], e[1][2], t2);
    for (int k = 0; k < 10; ++k) t1[k] -= t2[k];
    mul21_acc(t1, e[2][1], 1.0, C[9]);
    mul11(e[0][0], e[1][1], t1); mul11(e[0][1], e[1][0], t2);
    for (int k = 0; k < 10; ++k) t1[k] -= t2[k];
    mul21_acc(t1, e[2][2], 1.0, C[9]);
  }
  {
    // D = E E^T - 1/2 tr(E E^T) I  (degree 2, symmetric), rows 0..8 = D E  (:928-950)
    double D[3][3][10];
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j) {
        double t[10];
        for (int k = 0; k < 10; ++k) D[i][j][k] = 0.0;
        for (int q = 0; q < 3; ++q) {
          mul11(e[i][q], e[j][q], t);
          for (int k = 0; k < 10; ++k) D[i][j][k] += t[k];
        }
      }
    for (int k = 0; k < 10; ++k) {
      const double tr = 0.5 * (D[0][0][k] + D[1][1][k] + D[2][2][k]);
      D[0][0][k] -= tr; D[1][1][k] -= tr; D[2][2][k] -= tr;
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int q = 0; q < 3; ++q) mul21_acc(i <= q ? D[i][q] : D[q][i], e[q][j], 1.0, C[3 * i + j]);
  }
  double C0[10][20];   // the constraints before elimination: the final Gauss-Newton polish of (x, y, z) evaluates them
  for (int r = 0; r < 10; ++r)
    for (int j = 0; j < 20; ++j) C0[r][j] = C[r][j];
  // ---- 3. Gauss-Jordan on columns 0..9 (partial pivoting), keeping Nister's row order: row r ends with a 1 in column r
  for (int col = 0; col < 10; ++col) {
    int piv = col;
    double best = fabs(C[col][col]);
    for (int r = col + 1; r < 10; ++r)
      if (fabs(C[r][col]) > best) { best = fabs(C[r][col]); piv = r; }
    if (!(best > 1e-300)) return 0;
    if (piv != col)
      for (int j = col; j < 20; ++j) { const double t = C[col][j]; C[col][j] = C[piv][j]; C[piv][j] = t; }
    const double inv = 1.0 / C[col][col];
    for (int j = col; j < 20; ++j) C[col][j] *= inv;
    for (int r = 0; r < 10; ++r) {
      if (r == col) continue;
      const double f = C[r][col];
      if (f != 0.0)
        for (int j = col; j < 20; ++j) C[r][j] = fma(-f, C[col][j], C[r][j]);
    }
  }
  // ---- 4. A(z): row i = (row 4+2i) - z (row 5+2i); polynomials stored lowest power first
  double Ax[3][4], Ay[3][4], A1[3][5];
  for (int i = 0; i < 3; ++i) {
    const double* u = C[4 + 2 * i] + 10;   // [x z^2, x z, x, y z^2, y z, y, z^3, z^2, z, 1]
    const double* v = C[5 + 2 * i] + 10;
    // x-part: u0 z^2 + u1 z + u2 - z (v0 z^2 + v1 z + v2)
    Ax[i][0] = u[2]; Ax[i][1] = u[1] - v[2]; Ax[i][2] = u[0] - v[1]; Ax[i][3] = -v[0];
    Ay[i][0] = u[5]; Ay[i][1] = u[4] - v[5]; Ay[i][2] = u[3] - v[4]; Ay[i][3] = -v[3];
    A1[i][0] = u[9]; A1[i][1] = u[8] - v[9]; A1[i][2] = u[7] - v[8]; A1[i][3] = u[6] - v[7]; A1[i][4] = -v[6];
  }
  // ---- 5. det A(z)
  double det[11];
  for (int k = 0; k < 11; ++k) det[k] = 0.0;
  {
    double m[8], t[11];
    // + Ax0 (Ay1 A1_2 - A1_1 Ay2)
    double m2[8];
    polymul(Ay[1], 3, A1[2], 4, m); polymul(A1[1], 4, Ay[2], 3, m2);
    for (int k = 0; k < 8; ++k) m[k] -= m2[k];
    polymul(Ax[0], 3, m, 7, t);
    for (int k = 0; k < 11; ++k) det[k] += t[k];
    // - Ay0 (Ax1 A1_2 - A1_1 Ax2)
    polymul(Ax[1], 3, A1[2], 4, m); polymul(A1[1], 4, Ax[2], 3, m2);
    for (int k = 0; k < 8; ++k) m[k] -= m2[k];
    polymul(Ay[0], 3, m, 7, t);
    for (int k = 0; k < 11; ++k) det[k] -= t[k];
    // + A1_0 (Ax1 Ay2 - Ay1 Ax2)
    double m6[7], m6b[7];
    polymul(Ax[1], 3, Ay[2], 3, m6); polymul(Ay[1], 3, Ax[2], 3, m6b);
    for (int k = 0; k < 7; ++k) m6[k] -= m6b[k];
    polymul(A1[0], 4, m6, 6, t);
    for (int k = 0; k < 11; ++k) det[k] += t[k];
  }
  double roots[10];
  const int nr = real_roots_deg10(det, roots);
  // ---- 6. back-substitution
  int ns = 0;
  for (int q = 0; q < nr; ++q) {
    const double z = roots[q];
    double bx[3], by[3], b1[3];
    for (int i = 0; i < 3; ++i) {
      bx[i] = ((Ax[i][3] * z + Ax[i][2]) * z + Ax[i][1]) * z + Ax[i][0];
      by[i] = ((Ay[i][3] * z + Ay[i][2]) * z + Ay[i][1]) * z + Ay[i][0];
      b1[i] = (((A1[i][4] * z + A1[i][3]) * z + A1[i][2]) * z + A1[i][1]) * z + A1[i][0];
    }
    // A(z) (x, y, 1)^T = 0: three equations of rank 2 at a root.  Solve the best-conditioned pair of rows exactly
    // (normal equations would square the condition number; the reference solves the first two rows, :1008)
    double x = 0.0, y = 0.0, bestdet = 0.0;
    for (int a = 0; a < 3; ++a) {
      const int b = (a + 1) % 3;
      const double dd = bx[a] * by[b] - bx[b] * by[a];
      if (fabs(dd) > fabs(bestdet)) {
        bestdet = dd;
        x = -(b1[a] * by[b] - b1[b] * by[a]) / dd;
        y = -(bx[a] * b1[b] - bx[b] * b1[a]) / dd;
      }
    }
    if (!(fabs(bestdet) > 1e-300)) continue;
    // Two Gauss-Newton steps on the ten cubic constraints in (x, y, z): removes what the elimination and the degree-10
    // root finding lose on ill-conditioned samples (near-double roots keep ~1e-8, everything else reaches ~1e-14)
    double zz = z;
    for (int it = 0; it < 2; ++it) {
      // monomials [x^3, y^3, x^2y, xy^2, x^2z, x^2, y^2z, y^2, xyz, xy, xz^2, xz, x, yz^2, yz, y, z^3, z^2, z, 1] and gradients
      const double m[20] = {x * x * x, y * y * y, x * x * y, x * y * y, x * x * zz, x * x, y * y * zz, y * y, x * y * zz, x * y,
                            x * zz * zz, x * zz, x, y * zz * zz, y * zz, y, zz * zz * zz, zz * zz, zz, 1.0};
      const double mx[20] = {3 * x * x, 0, 2 * x * y, y * y, 2 * x * zz, 2 * x, 0, 0, y * zz, y, zz * zz, zz, 1, 0, 0, 0, 0, 0, 0, 0};
      const double my[20] = {0, 3 * y * y, x * x, 2 * x * y, 0, 0, 2 * y * zz, 2 * y, x * zz, x, 0, 0, 0, zz * zz, zz, 1, 0, 0, 0, 0};
      const double mz[20] = {0, 0, 0, 0, x * x, 0, y * y, 0, x * y, 0, 2 * x * zz, x, 0, 2 * y * zz, y, 0, 3 * zz * zz, 2 * zz, 1, 0};
      double JtJ[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Jtf[3] = {0, 0, 0};
      for (int r = 0; r < 10; ++r) {
        double f = 0, jx = 0, jy = 0, jz = 0;
        for (int j = 0; j < 20; ++j) {
          f = fma(C0[r][j], m[j], f); jx = fma(C0[r][j], mx[j], jx);
          jy = fma(C0[r][j], my[j], jy); jz = fma(C0[r][j], mz[j], jz);
        }
        const double J[3] = {jx, jy, jz};
        for (int a = 0; a < 3; ++a) {
          Jtf[a] = fma(J[a], f, Jtf[a]);
          for (int b = 0; b < 3; ++b) JtJ[a][b] = fma(J[a], J[b], JtJ[a][b]);
        }
      }
      const double c00 = JtJ[1][1] * JtJ[2][2] - JtJ[1][2] * JtJ[2][1], c01 = JtJ[1][2] * JtJ[2][0] - JtJ[1][0] * JtJ[2][2],
                   c02 = JtJ[1][0] * JtJ[2][1] - JtJ[1][1] * JtJ[2][0];
      const double dj = JtJ[0][0] * c00 + JtJ[0][1] * c01 + JtJ[0][2] * c02;
      if (!(fabs(dj) > 1e-300)) break;
      const double c11 = JtJ[0][0] * JtJ[2][2] - JtJ[0][2] * JtJ[2][0], c12 = JtJ[0][1] * JtJ[2][0] - JtJ[0][0] * JtJ[2][1],
                   c22 = JtJ[0][0] * JtJ[1][1] - JtJ[0][1] * JtJ[1][0];
      // delta = -inv(JtJ) Jtf  (symmetric adjugate)
      const double dx = -(c00 * Jtf[0] + c01 * Jtf[1] + c02 * Jtf[2]) / dj;
      const double dy = -(c01 * Jtf[0] + c11 * Jtf[1] + c12 * Jtf[2]) / dj;
      const double dz = -(c02 * Jtf[0] + c12 * Jtf[1] + c22 * Jtf[2]) / dj;
      if (!isfinite(dx + dy + dz)) break;
      const double scale = 1.0 + fabs(x) + fabs(y) + fabs(zz);
      if (fabs(dx) + fabs(dy) + fabs(dz) > 1e-2 * scale) break;   // far from a solution (spurious root): leave it
      x += dx; y += dy; zz += dz;
    }
    double nn = 0.0, E[9];
    for (int j = 0; j < 9; ++j) {
      E[j] = x * N[0][j] + y * N[1][j] + zz * N[2][j] + N[3][j];
      nn = fma(E[j], E[j], nn);
    }
    if (!(nn > 0.0) || !isfinite(nn)) continue;
    const double inv = rsqrt(nn);
    for (int j = 0; j < 9; ++j) Eout[ns][j] = E[j] * inv;
    ++ns;
  }
  return ns;
}

}  // namespace fivept
}  // namespace far
