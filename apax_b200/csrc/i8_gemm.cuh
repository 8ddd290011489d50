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
//     such that |x[m][k] c[k] 2^-e[m]| < 64, then written as a fixed-point number and split into balanced
//     base-256 digits d0, d1, ... in [-128,127]:  v = sum_s d_s 256^-s  (4 digits, or 3 for the 360-feature
//     descriptor whose 24-bit slices equal fp32's own significand);
//   * the weights get the reciprocal per-feature scale (w[k][n] / c[k], row maximum in [32,64)) and four
//     digits, once per parameter set;
//   * D_o[m][n] = sum_{s+t=o} sum_k a_s[m][k] b_t[k][n] for o = 0..3 -- ten (nine) int8 MMAs per 32-wide k-step
//     into four int32 accumulators;
//   * epilogue: y[m][n] = 2^e[m] (D_0 + D_1/256 + D_2/256^2 + D_3/256^3), then the layer's fused tail.
//
// Kernel shape (one persistent CTA per SM; TMA warp, MMA warp, eight epilogue warps):
//   * activation-stationary: the 128-atom tile's digit planes for the whole reduction (<= 144 KB) sit in shared
//     memory while the CTA sweeps the layer's outputs in 64-wide sub-tiles, streaming only weight planes
//     (16 KB per 64-byte k-block) from L2 through a ring -- the L2 -> SM stream is what bounded the first version;
//   * two TMEM accumulator sets (2 x 4 x 64 columns = all 512): the epilogue of sub-tile j overlaps the MMAs of
//     j + 1; draining tensor memory (64 B/clk/SM) is the epilogue's floor;
//   * digit planes are K-major [row][kp] int8; TMA boxes {64 bytes of k, rows} with SWIZZLE_64B land in the
//     canonical UMMA K-major SW64 layout (8-row x 64-byte atoms, 512 bytes apart);
//   * TMEM lane == atom: fp32 stores are coalesced along atoms, and one thread sees a whole output row of its
//     atom over the sweep, so row maxima, the energy head and the next layer's digit planes come out of the
//     epilogue directly (fused modes below) -- activations never exist in fp32 in HBM.
#pragma once
#include <cuda.h>

#include <vector>

#include "common.cuh"

namespace apax {

constexpr int I8_BM = 128;      // atoms per CTA tile (UMMA M)
constexpr int I8_BN = 64;       // outputs per sub-tile (UMMA N)
constexpr int I8_KC = 64;       // k bytes per k-block (two K = 32 MMAs per digit pair)
constexpr int I8_DIGITS = 4;
constexpr int I8_MAX_KB = 6;    // reduction length <= 384
constexpr int I8_MAX_HEADS = 16;

enum {
  I8_EPI_STORE = 0,     // out[n][m] = y
  I8_EPI_ACT_Q = 1,     // h = swish(y + bias): aux[n][m] = swish'(.), h -> digit planes of the next layer (row-max scaled)
  I8_EPI_ACT_HEAD = 2,  // h = swish(y + bias): h_out[k][m] = sum_n h w_head[n][k]; swish'(.) -> digit planes (fixed scale)
  I8_EPI_MULD_Q = 3     // t = y * aux[n][m] * row_mul[m] -> digit planes of the next layer (row-max scaled)
};

struct I8Planes {        // digit planes of a [rows][kp] matrix, plane p at base + p * plane_stride
  const int8_t* base;
  int64_t rows;          // activations: multiple of 128; weights: multiple of 64
  int64_t kp;            // padded reduction length in bytes, multiple of 64
  int64_t plane_stride;  // bytes between planes
  int n_planes;          // 3 or 4 (weights: 4)
};

struct I8Epilogue {
  int mode;
  int64_t ld;              // atom stride of the feature-major fp32 matrices
  int n_valid;             // outputs >= n_valid are ignored
  const int32_t* row_exp;  // [rows] exponents e[m] of the activation planes; null: fixed_exp for every row
  int fixed_exp;
  float* out;              // STORE: [n][ld]
  const float* bias;       // ACT_*: [n]
  float* aux;              // ACT_Q: swish' written, MULD_Q: multiplier read -- tile-major [atom tile][n_outputs][128 atoms]
  const float* row_mul;    // MULD_Q: per-atom factor [ld] (may be null)
  // digit planes produced for the next GEMM (always four planes of a [rows][q_kp] matrix)
  int8_t* q_planes;
  int64_t q_kp, q_plane_stride;
  const float* q_col_scale;  // [n] per-feature powers of two of the consuming layer
  int32_t* q_row_exp;        // [rows] written by ACT_Q / MULD_Q
  int q_fixed_exp;           // ACT_HEAD: exponent shared by all rows
  float* scratch;            // ACT_Q / MULD_Q: [grid][n_outputs][128] fp32 staging (stays in L2)
  // energy head (ACT_HEAD)
  const float* w_head;       // [n][n_out]
  int n_out;
  float* h_out;              // [n_out][ld]
  long long* dbg;            // optional [grid][8] cycle counters
  int dbg_flags;             // measurement only: 1 = epilogue skips the drain, 2 = producer signals without loading
};

// activations: feature-major fp32 [K][ld] -> digit planes [n_planes][ld][kp] + per-atom exponents
#ifdef __CUDACC__
// ---- device helpers shared by the GEMM epilogues, the stand-alone quantiser and the fused contraction + quantiser
__device__ __forceinline__ float exp2_int(int e) {
  if (e >= -126 && e <= 127) return __int_as_float((e + 127) << 23);
  return ldexpf(1.0f, e);
}

// ---- digits.  Balanced base-256 digits of I = rn(v), all at once: adding 128 to each low byte (with the
// carries that produces) turns the balanced digits into offset-binary bytes, flipping their top bits back gives
// two's complement.  P digits: byte P-1 = leading digit d0 ... byte 0 = d_{P-1}.
template <int P>
__device__ __forceinline__ uint32_t i8_digit_word(float v) {
  constexpr uint32_t kOff = P == 4 ? 0x00808080u : 0x00008080u;
  return (uint32_t(__float2int_rn(v)) + kOff) ^ kOff;   // |I| < 2^(8 P - 2)
}
// 4 x 4 byte transpose: words e[j] of elements j = 0..3  ->  tb[b] = byte b of the four elements (element 0 lowest)
__device__ __forceinline__ void i8_transpose4(const uint32_t (&e)[4], uint32_t (&tb)[4]) {
  const uint32_t t0 = __byte_perm(e[0], e[1], 0x5140);  // e0.b0 e1.b0 e0.b1 e1.b1
  const uint32_t t1 = __byte_perm(e[2], e[3], 0x5140);
  const uint32_t t2 = __byte_perm(e[0], e[1], 0x7362);  // e0.b2 e1.b2 e0.b3 e1.b3
  const uint32_t t3 = __byte_perm(e[2], e[3], 0x7362);
  tb[0] = __byte_perm(t0, t1, 0x5410);
  tb[1] = __byte_perm(t0, t1, 0x7632);
  tb[2] = __byte_perm(t2, t3, 0x5410);
  tb[3] = __byte_perm(t2, t3, 0x7632);
}

// exponent e (row maximum 2^-e in [32, 64)) and two in-range powers of two whose product is 2^(8 (P - 1) - e)
__device__ __forceinline__ void i8_row_scale(float mx, int P, int& e, float& sc, float& sc2) {
  e = 0;
  sc = 0.0f;
  sc2 = 0.0f;                         // zero / padding / non-finite rows quantise to all-zero digits
  if (mx > 0.0f && mx < 3.0e38f) {
    int ex;
    frexpf(mx, &ex);                  // mx = f 2^ex, f in [0.5, 1)
    e = ex - 6;
    const int tot = 8 * (P - 1) - e, h = tot / 2;
    sc = exp2_int(h);
    sc2 = exp2_int(tot - h);
  }
}

#endif  // __CUDACC__

int i8_quantize_rows(cudaStream_t stream, const float* xt, int64_t ld, int64_t n_valid_rows, int K, int kp, int n_planes,
                     const float* col_scale /* [kp] powers of two */, int8_t* planes, int32_t* row_exp,
                     int32_t* rowmax /* [ld] float bits of the scaled row maxima */, bool have_rowmax = false,
                     int64_t n_rows = 0 /* rows to emit, multiple of 128; 0 = ld */);

// weights (host): W[k][n] (row stride ldw) -> planes [4][ncp][kp] of W[k][n] / c[k].  col_scale: computed here from
// the row maxima of W (given_scale = false), or taken as given (shared between several weight matrices).
void i8_quantize_weights(const float* W, int K, int NC, int ldw, int kp, int ncp, std::vector<int8_t>& planes,
                         std::vector<float>& col_scale, bool given_scale = false);
// per-k scale c[k] = 2^(exponent of max_n |W[k][n]| - 6); merges into an existing scale vector by maximum
void i8_weight_scales(const float* W, int K, int NC, int ldw, int kp, std::vector<float>& col_scale, bool merge);

int i8_gemm(cudaStream_t stream, const I8Planes& act, const I8Planes& wgt, const I8Epilogue& epi, int32_t* err_flag);

// per-CTA staging floats needed by the fused modes for n_outputs outputs and the given number of atom rows
inline int64_t i8_scratch_floats(int64_t rows, int n_sm, int n_outputs) {
  const int64_t tiles = rows / I8_BM;
  return (tiles < n_sm ? tiles : n_sm) * int64_t(n_outputs) * I8_BM;
}

}  // namespace apax
