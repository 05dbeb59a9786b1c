// Host build of far_b200/csrc/fivept.cuh (the device code is plain C++ apart from the qualifiers): lets the CPU test
// suite pin the 5-point solver to the oracle without a GPU.  stdin: S, then S x 5 lines "x1 y1 x2 y2"; stdout: per
// sample the number of solutions and the row-major matrices.
#include <cmath>
#include <cstdio>
#define __device__
#define __forceinline__ inline
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::isfinite;
#define FAR_FIVEPT_HOST_BUILD 1
#include "fivept_nocuda.cuh"
int main() {
  int S;
  if (scanf("%d", &S) != 1) return 1;
  for (int s = 0; s < S; ++s) {
    double x1[5], y1[5], x2[5], y2[5];
    for (int k = 0; k < 5; ++k)
      if (scanf("%lf %lf %lf %lf", &x1[k], &y1[k], &x2[k], &y2[k]) != 4) return 1;
    double E[10][9];
    const int n = far::fivept::solve(x1, y1, x2, y2, E);
    printf("%d\n", n);
    for (int q = 0; q < n; ++q) {
      for (int k = 0; k < 9; ++k) printf("%.17g ", E[q][k]);
      printf("\n");
    }
  }
  return 0;
}
