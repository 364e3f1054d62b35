/* C restatement of the reference's chamfer semantics — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Semantics restated from /root/reference/model/loss.py:45-70 (squared distance map, row/column minima) and the
 * torch behaviour the reference relies on implicitly: min(dim) returns the FIRST minimum and autograd routes the
 * gradient to that single index.  Evaluated in DIFFERENCE form, not the reference's
 * expanded form, in two precisions:
 *   *_f64 : double arithmetic on the fp32 inputs — the tight parity target (O2 of SURVEY.md §8c);
 *   *_f32 : float arithmetic with exactly the CUDA kernels' operation order  dx*dx -> fmaf(dy,dy,.) -> fmaf(dz,dz,.)
 *           so indices and distances can be compared BIT-exactly with the GPU at full benchmark sizes.
 * Pinned (through tests/test_oracle_golden.py) against tests/golden/chamfer_*.npz, which hold outputs of the
 * reference's own code.  Build: make -C oracle   (gcc -O3 -mavx2 -mfma -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* x [B,N,3], y [B,M,3] contiguous fp32. dist1/idx1 [B,N], dist2/idx2 [B,M]. */
void chamfer_fwd_f64(const float* x, const float* y, int B, int N, int M,
                     double* dist1, double* dist2, int32_t* idx1, int32_t* idx2) {
#pragma omp parallel for schedule(static) collapse(2)
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < N; ++i) {
      const float* p = x + ((size_t)b * N + i) * 3;
      const float* q = y + (size_t)b * M * 3;
      double best = INFINITY; int bj = 0;
      for (int j = 0; j < M; ++j) {
        double dx = (double)p[0] - q[3 * j], dy = (double)p[1] - q[3 * j + 1], dz = (double)p[2] - q[3 * j + 2];
        double d = dx * dx + dy * dy + dz * dz;
        if (d < best) { best = d; bj = j; }
      }
      dist1[(size_t)b * N + i] = best; idx1[(size_t)b * N + i] = bj;
    }
#pragma omp parallel for schedule(static) collapse(2)
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < M; ++j) {
      const float* q = y + ((size_t)b * M + j) * 3;
      const float* p = x + (size_t)b * N * 3;
      double best = INFINITY; int bi = 0;
      for (int i = 0; i < N; ++i) {
        double dx = (double)p[3 * i] - q[0], dy = (double)p[3 * i + 1] - q[1], dz = (double)p[3 * i + 2] - q[2];
        double d = dx * dx + dy * dy + dz * dz;
        if (d < best) { best = d; bi = i; }
      }
      dist2[(size_t)b * M + j] = best; idx2[(size_t)b * M + j] = bi;
    }
}

static inline float sqd32(const float* p, const float* q) {
  /* same order as rma::sqdist_neg / the packed search loop: (x - y) then mul, fma, fma */
  float dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
  float d = dx * dx;
  d = fmaf(dy, dy, d);
  d = fmaf(dz, dz, d);
  return d;
}

void chamfer_fwd_f32(const float* x, const float* y, int B, int N, int M,
                     float* dist1, float* dist2, int32_t* idx1, int32_t* idx2) {
#pragma omp parallel for schedule(static) collapse(2)
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < N; ++i) {
      const float* p = x + ((size_t)b * N + i) * 3;
      const float* q = y + (size_t)b * M * 3;
      float best = INFINITY; int bj = 0;
      for (int j = 0; j < M; ++j) {
        float d = sqd32(p, q + 3 * j);
        if (d < best) { best = d; bj = j; }
      }
      dist1[(size_t)b * N + i] = best; idx1[(size_t)b * N + i] = bj;
    }
#pragma omp parallel for schedule(static) collapse(2)
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < M; ++j) {
      const float* q = y + ((size_t)b * M + j) * 3;
      const float* p = x + (size_t)b * N * 3;
      float best = INFINITY; int bi = 0;
      for (int i = 0; i < N; ++i) {
        float d = sqd32(p + 3 * i, q);
        if (d < best) { best = d; bi = i; }
      }
      dist2[(size_t)b * M + j] = best; idx2[(size_t)b * M + j] = bi;
    }
}

/* Analytic backward in double (serial scatter; deterministic).  g1 [B,N], g2 [B,M] upstream gradients. */
void chamfer_bwd_f64(const float* x, const float* y, int B, int N, int M, const double* g1, const double* g2,
                     const int32_t* idx1, const int32_t* idx2, double* gx, double* gy) {
  for (size_t t = 0; t < (size_t)B * N * 3; ++t) gx[t] = 0.0;
  for (size_t t = 0; t < (size_t)B * M * 3; ++t) gy[t] = 0.0;
  for (int b = 0; b < B; ++b) {
    const float* xb = x + (size_t)b * N * 3; const float* yb = y + (size_t)b * M * 3;
    double* gxb = gx + (size_t)b * N * 3; double* gyb = gy + (size_t)b * M * 3;
    for (int i = 0; i < N; ++i) {
      int j = idx1[(size_t)b * N + i]; double g = 2.0 * g1[(size_t)b * N + i];
      for (int c = 0; c < 3; ++c) {
        double v = g * ((double)xb[3 * i + c] - (double)yb[3 * j + c]);
        gxb[3 * i + c] += v; gyb[3 * j + c] -= v;
      }
    }
    for (int j = 0; j < M; ++j) {
      int i = idx2[(size_t)b * M + j]; double g = 2.0 * g2[(size_t)b * M + j];
      for (int c = 0; c < 3; ++c) {
        double v = g * ((double)yb[3 * j + c] - (double)xb[3 * i + c]);
        gyb[3 * j + c] += v; gxb[3 * i + c] -= v;
      }
    }
  }
}
