/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Plain-C restatement of the two leaf kernels of the reference's
 * fake-quantization path, used by tests/ (cross-check of the numpy/torch oracle and of the CUDA kernels) and by
 * bench.py's cpu_baseline leg (OpenMP over the host cores).  The product never links or loads this file.
 *
 *   oracle_float2gemmlowp  follows kernels/gemmlowp.cu:8-45 literally (fp32, one rounding per C operator,
 *                          roundf = half away from zero, fminf/fmaxf, `out*scale - shift` as ONE fma because that
 *                          is what the reference's own nvcc build emits - see DESIGN.md "a1 contraction").
 *   oracle_quantize1_rows  follows pytorch_quantizer/quantization/qtypes/int_quantizer.py:557-603
 *                          (enforce_true_zero branch, per-row delta/offset/bit_alloc, scale floor 1e-8,
 *                          round-half-even).
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (oracle/build_oracle.py).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* returns 1 when the reference would hand back its input (range <= 0), 0 otherwise */
int oracle_float2gemmlowp(const float* in, float* out, int64_t n, float range, float offset, int num_bits,
                          int int_exp, int enforce_true_zero, const float* noise) {
  if (range <= 0) return 1;
  long long qmax_i = (0x1l << num_bits) - 1;
  float qmax = (float)qmax_i;
  float scale = range / qmax;
  if (int_exp) scale = powf(2, (int)ceilf(log2f(scale)));
  float zero_point = roundf(-offset / scale);
  float shift = enforce_true_zero ? zero_point : -offset;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    float v = in[i];
    if (enforce_true_zero)
      v = (v / scale) + shift;
    else
      v = (v + shift) / scale;
    if (noise) v += noise[i];
    v = fminf(v, qmax);
    v = fmaxf(v, 0.f);
    v = roundf(v);
    if (enforce_true_zero)
      v = (v - shift) * scale;
    else
      v = fmaf(v, scale, -shift);
    out[i] = v;
  }
  return 0;
}

/* x: [rows][cols]; delta/offset: rows entries (per_row) or 1; bits: NULL or rows entries */
void oracle_quantize1_rows(const float* x, float* y, float* grid, int64_t rows, int64_t cols, const float* delta,
                           const float* offset, const float* bits, int per_row, int num_bits) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const int64_t pi = per_row ? r : 0;
    float qmax, scale;
    if (bits) {
      qmax = exp2f(bits[r]) - 1.f;
      scale = (qmax > 0.f) ? delta[pi] / qmax : 0.f;
    } else {
      qmax = (float)(exp2((double)num_bits) - 1.0);
      scale = delta[pi] / qmax;
    }
    scale = fmaxf(scale, 1e-8f);
    const float zp = nearbyintf(0.f - offset[pi] / scale);
    for (int64_t c = 0; c < cols; ++c) {
      float v = x[r * cols + c] / scale;
      v = v + zp;
      if (v > qmax) v = qmax;
      if (v < 0.f) v = 0.f;
      v = nearbyintf(v);
      if (grid) grid[r * cols + c] = v;
      v = v - zp;
      y[r * cols + c] = v * scale;
    }
  }
}
