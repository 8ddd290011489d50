// Readout GEMMs on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a only.
//
//   D[atom][n] = sum_k X^T[k][atom] * W[k][n]      (feature-major activations, row-major weights)
//
// Precision: the reference's readout is fp32 (apax/layers/ntk_linear.py:40).  Each operand is kept as
// an exact pair x = hi + lo with hi = tf32-rounded value (low 13 mantissa bits zero) and lo = x - hi, and
// every k-step issues three kind::tf32 MMAs -- hi*hi + hi*lo + lo*hi -- into one fp32 TMEM accumulator.
// The dropped lo*lo term is 2^-22 relative, i.e. the result is fp32-accurate ("3xTF32").
//
// Data path: both operands are MN-major in shared memory (atoms resp. output features contiguous),
// loaded by TMA as 3-d boxes {32 elements = 128 B, 16 k-rows, chunks} with SWIZZLE_128B, which is exactly
// the canonical UMMA MN-major SW128 layout ((8 x 16B, chunks), (8 k-rows, groups)); the accumulator
// (128 atoms = 128 TMEM lanes x BN fp32 columns) is read back with tcgen05.ld so that lane == atom and
// the epilogue's global stores are coalesced along atoms -- the next layer's feature-major input.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace apax {

constexpr int TC_BM = 128;      // atoms per CTA (UMMA M, cta_group::1)
constexpr int TC_BK = 16;       // k-rows per pipeline stage (two UMMA K=8 groups)
constexpr int TC_STAGES = 2;    // 2 x 48 KB (BN = 256): two CTAs per SM overlap MMA and epilogue

enum { TC_EPI_STORE = 0, TC_EPI_ACT = 1, TC_EPI_MULD = 2 };

struct TcEpilogue {
  int mode;
  int64_t ld;         // atom stride of every feature-major matrix
  float* out_hi;      // [NC][ld]  STORE: plain fp32 result;  ACT / MULD: tf32-rounded part
  float* out_lo;      // [NC][ld]  ACT / MULD: residual
  float* aux;         // ACT: swish'(y) written here;  MULD: multiplier read from here
  const float* bias;  // ACT: [NC]
  int dbg_flags;      // debug only: 1 = K-major instruction descriptor, 2 = LBO/SBO swapped, 4 = no swizzle bits
  float* dbg;         // debug only: first stage of shared memory is dumped here (16384 floats) when non-null
};

struct TcOperand {   // one feature-major / row-major fp32 matrix [rows][cols], as hi and lo parts
  const float* hi;
  const float* lo;
  int64_t rows, cols;
};

// Y^T[NC][ld] (+ epilogue) from activations [K][ld] and weights [K][NC]; NC % 128 == 0, ld % 128 == 0.
int tc_gemm(cudaStream_t stream, const TcOperand& act, const TcOperand& wgt, int64_t n_atoms, const TcEpilogue& epi,
            int32_t* err_flag);

// tf32 split of an fp32 value (device + host): hi = rna_tf32(x), lo = rna_tf32(x - hi).  Both parts have
// zero low mantissa bits, so whatever the tensor core does with bits below tf32 precision is irrelevant;
// x - (hi + lo) is at most 2^-22 |x|.
__host__ __device__ inline float round_tf32(float x) {
#ifdef __CUDA_ARCH__
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u & 0xffffe000u);
#else
  union { float f; uint32_t u; } v;
  v.f = x;
  v.u = (v.u + 0x1000u) & 0xffffe000u;  // round half away from zero (rna), finite inputs
  return v.f;
#endif
}
__host__ __device__ inline void split_tf32(float x, float& hi, float& lo) {
  hi = round_tf32(x);
  lo = round_tf32(x - hi);
}

}  // namespace apax
