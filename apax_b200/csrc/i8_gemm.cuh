// Exact readout GEMMs on the integer tensor cores (tcgen05.mma kind::i8, int32 TMEM accumulators).
//
// Why integers: the reference's readout is fp32 (apax/layers/ntk_linear.py:40) and north_star asks for total
// energies within 1e-6.  The floating-point tensor-core path (tc_gemm.cuh, 3xTF32) is fast but its fp32
// accumulator truncates, which biases ~1M-atom energy sums by 8e-6.  Integer MMAs have
// no rounding at all, so the only errors left are the operand quantisation (round to nearest, 2^-30 of the
// row maximum) and the dropped digit products below 2^-32 -- both unbiased and below fp32 resolution.
//
// Scheme (Ozaki-style slicing, fixed 8-bit slices):
//   * every activation row (one atom) is scaled by a power of two 2^-e[m] and a per-feature power of two c[k]
//     such that |x[m][k] c[k] 2^-e[m]| < 64, then written as a 32-bit fixed-point number I = rn(v 2^24) and
//     split into four balanced base-256 digits d0..d3 in [-128,127]:  v = sum_s d_s 256^-s;
//   * the weights get the reciprocal per-feature scale (w[k][n] / c[k], row maximum in [32,64)) and the
//     same four digits, once per parameter set;
//   * D_o[m][n] = sum_{s+t=o} sum_k a_s[m][k] b_t[k][n] for o = 0..3 -- ten int8 MMAs per 32-wide k-step into
//     four int32 accumulators of 128 TMEM columns each (all 512 columns of the SM);
//   * epilogue: y[m][n] = 2^e[m] (D_0 + D_1/256 + D_2/256^2 + D_3/256^3), then bias / swish / multipliers.
//
// Data path: digit planes are K-major int8 matrices [row][Kp]; TMA loads {64 bytes of k, 128 rows} boxes with
// SWIZZLE_64B (the canonical UMMA K-major SW64 layout: 8-row x 64-byte atoms, 512 bytes apart), three
// pipeline stages of 8 planes x 8 KB.  One CTA (128 threads: TMA thread, MMA thread, 4 epilogue warps) per
// 128 atoms x 128 outputs tile; TMEM lane == atom, so epilogue stores are coalesced along atoms.
#pragma once
#include <cuda.h>

#include <vector>

#include "common.cuh"

namespace apax {

constexpr int I8_BM = 128;      // atoms per CTA (UMMA M)
constexpr int I8_BN = 128;      // outputs per CTA (UMMA N); 4 accumulators x 128 columns = whole TMEM
constexpr int I8_KC = 64;       // k bytes per pipeline stage (two K = 32 MMAs per digit pair)
constexpr int I8_STAGES = 3;
constexpr int I8_DIGITS = 4;

enum { I8_EPI_STORE = 0, I8_EPI_ACT = 1, I8_EPI_MULD = 2 };

struct I8Planes {        // four digit planes of a [rows][kp] matrix, plane p at base + p * plane_stride
  const int8_t* base;
  int64_t rows;          // multiple of 128
  int64_t kp;            // padded reduction length in bytes, multiple of 64
  int64_t plane_stride;  // bytes between planes
};

struct I8Epilogue {
  int mode;
  int64_t ld;            // atom stride of the feature-major fp32 outputs
  int n_valid;           // outputs >= n_valid are not stored
  const int32_t* row_exp;  // [rows of A] e[m]
  float* out;            // [n][ld]  STORE: y;  ACT: swish(y + bias);  MULD: y * aux * row_mul
  float* aux;            // ACT: swish'(y + bias) written;  MULD: multiplier read
  const float* bias;     // ACT: [n]
  const float* row_mul;  // MULD: per-atom factor [ld] (may be null)
};

// activations: feature-major fp32 [K][ld] -> digit planes [4][ld][kp] + per-atom exponents
int i8_quantize_rows(cudaStream_t stream, const float* xt, int64_t ld, int64_t n_valid_rows, int K, int kp,
                     const float* col_scale /* [kp] powers of two */, int8_t* planes, int32_t* row_exp);

// weights (host): W[k][n] (row stride ldw) -> planes [4][ncp][kp] of W[k][n] / c[k] and the per-k scales c[k]
void i8_quantize_weights(const float* W, int K, int NC, int ldw, int kp, int ncp, std::vector<int8_t>& planes,
                         std::vector<float>& col_scale);

int i8_gemm(cudaStream_t stream, const I8Planes& act, const I8Planes& wgt, const I8Epilogue& epi, int32_t* err_flag);

}  // namespace apax
