// Exact fixed-point readout GEMM on tcgen05.mma kind::i8.  See i8_gemm.cuh.
#include "i8_gemm.cuh"

#include <cmath>

#include "tc_ptx.cuh"

namespace apax {

namespace {

using namespace tcptx;

constexpr uint32_t I8_PLANE_TILE = I8_BM * I8_KC;                       // 8 KB: 128 rows x 64 bytes
constexpr uint32_t I8_STAGE_BYTES = 2 * I8_DIGITS * I8_PLANE_TILE;       // 4 A planes + 4 B planes = 64 KB
constexpr uint32_t I8_SMEM_BYTES = I8_STAGES * I8_STAGE_BYTES + 1024;   // + slack for 1024-byte alignment
constexpr uint32_t I8_TMEM_COLS = I8_DIGITS * I8_BN;                     // 512
constexpr float kSwishI8 = 1.6765324703310907f;                          // apax/layers/activation.py:8

static_assert(I8_BM == I8_BN, "A and B plane tiles share one size");

// UMMA shared-memory descriptor, K-major, SWIZZLE_64B (cute::UMMA::LayoutType::SWIZZLE_64B = 4): rows of 64
// bytes, 8-row atoms 512 bytes apart (stride byte offset); the leading byte offset is unused for swizzled
// K-major operands.  Advancing the start address by 32 bytes selects the second K = 32 slice of the row.
__device__ __forceinline__ uint64_t make_desc_k_sw64(uint32_t addr) {
  return uint64_t((addr >> 4) & 0x3fffu) | (uint64_t(1) << 16) | (uint64_t(512u >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(4) << 61);
}

// cute::UMMA::InstrDescriptor for kind::i8: c_format S32 (2) at [4,6), a/b format signed 8 bit (1) at
// [7,10)/[10,13), both operands K-major (bits 15/16 = 0), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

__device__ __forceinline__ float exp2_int(int e) {
  if (e >= -126 && e <= 127) return __int_as_float((e + 127) << 23);
  return ldexpf(1.0f, e);
}

__global__ void __launch_bounds__(128) k_i8_gemm(const __grid_constant__ CUtensorMap tm_a,
                                                 const __grid_constant__ CUtensorMap tm_b, int num_k_blocks,
                                                 I8Epilogue epi, int32_t* err) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[I8_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[I8_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int m0 = blockIdx.x * I8_BM;
  const int n0 = blockIdx.y * I8_BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < I8_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(I8_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    for (int kb = 0; kb < num_k_blocks; ++kb) {
      const int s = kb % I8_STAGES;
      const uint32_t phase = (kb / I8_STAGES) & 1;
      if (!mbar_wait(smem_u32(&empty_bar[s]), phase ^ 1, err)) break;
      const uint32_t stage = smem_base + s * I8_STAGE_BYTES;
      const uint32_t bar = smem_u32(&full_bar[s]);
      mbar_expect_tx(bar, I8_STAGE_BYTES);
      const int k0 = kb * I8_KC;
#pragma unroll
      for (int d = 0; d < I8_DIGITS; ++d) {
        tma_load_3d(stage + d * I8_PLANE_TILE, &tm_a, bar, k0, m0, d);
        tma_load_3d(stage + (I8_DIGITS + d) * I8_PLANE_TILE, &tm_b, bar, k0, n0, d);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    constexpr uint32_t idesc = make_idesc_i8(I8_BM, I8_BN);
    bool ok = true;
    for (int kb = 0; kb < num_k_blocks && ok; ++kb) {
      const int s = kb % I8_STAGES;
      const uint32_t phase = (kb / I8_STAGES) & 1;
      ok = mbar_wait(smem_u32(&full_bar[s]), phase, err);
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t stage = smem_base + s * I8_STAGE_BYTES;
#pragma unroll
      for (int kk = 0; kk < I8_KC / 32; ++kk) {
        uint64_t a[I8_DIGITS], b[I8_DIGITS];
#pragma unroll
        for (int d = 0; d < I8_DIGITS; ++d) {
          a[d] = make_desc_k_sw64(stage + d * I8_PLANE_TILE + kk * 32);
          b[d] = make_desc_k_sw64(stage + (I8_DIGITS + d) * I8_PLANE_TILE + kk * 32);
        }
        const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
        // digit pair (s, t) contributes to order o = s + t; orders above 3 are below 2^-32 and dropped
#pragma unroll
        for (int o = 0; o < I8_DIGITS; ++o) {
#pragma unroll
          for (int sa = 0; sa <= o; ++sa) umma_i8(tmem_d + o * I8_BN, a[sa], b[o - sa], idesc, sa == 0 ? acc : 1u);
        }
      }
      umma_commit(smem_u32(&empty_bar[s]));  // frees the stage when the MMAs above have read it
    }
    umma_commit(smem_u32(&accum_bar));
  }
  __syncwarp();

  // ---------------------------------------------------------------------- epilogue: lane == atom
  const bool have = mbar_wait(smem_u32(&accum_bar), 0, err);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (have) {
    const int64_t m = int64_t(m0) + threadIdx.x;
    const float scale = exp2_int(epi.row_exp[m]);
    const float rmul = (epi.mode == I8_EPI_MULD && epi.row_mul) ? epi.row_mul[m] : 1.0f;
    const uint32_t taddr = tmem_d + (uint32_t(warp * 32) << 16);
    constexpr float kInv256 = 1.0f / 256.0f;
#pragma unroll 1
    for (int c = 0; c < I8_BN; c += 16) {
      uint32_t r0[16], r1[16], r2[16], r3[16];
      tmem_ld16_u32(taddr + c, r0);
      tmem_ld16_u32(taddr + I8_BN + c, r1);
      tmem_ld16_u32(taddr + 2 * I8_BN + c, r2);
      tmem_ld16_u32(taddr + 3 * I8_BN + c, r3);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int n = n0 + c + i;
        if (n >= epi.n_valid) break;
        float t = fmaf(float(int(r3[i])), kInv256, float(int(r2[i])));
        t = fmaf(t, kInv256, float(int(r1[i])));
        t = fmaf(t, kInv256, float(int(r0[i])));
        const float v = t * scale;
        const int64_t o = int64_t(n) * epi.ld + m;
        if (epi.mode == I8_EPI_STORE) {
          epi.out[o] = v;
        } else if (epi.mode == I8_EPI_ACT) {
          const float y = v + __ldg(epi.bias + n);
          const float sg = 1.0f / (1.0f + expf(-y));
          epi.out[o] = kSwishI8 * y * sg;                            // variance_preserving_swish (activation.py:7-9)
          epi.aux[o] = kSwishI8 * sg * (1.0f + y * (1.0f - sg));     // its derivative, for the backward pass
        } else {
          epi.out[o] = v * epi.aux[o] * rmul;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(I8_TMEM_COLS) : "memory");
  }
}

// ---- activations -> digit planes.  Thread == atom (coalesced feature-major reads); two passes over the
// atom's features: scaled row maximum, then fixed-point conversion and balanced base-256 digit extraction.
__global__ void __launch_bounds__(128) k_i8_quantize_rows(const float* __restrict__ xt, int64_t ld, int64_t n_valid_rows,
                                                          int K, int kp, const float* __restrict__ col_scale,
                                                          int8_t* __restrict__ planes, int32_t* __restrict__ row_exp) {
  const int64_t m = int64_t(blockIdx.x) * 128 + threadIdx.x;
  const float* x = xt + m;
  float mx = 0.0f;
  if (m < n_valid_rows) {
#pragma unroll 8
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, fabsf(x[int64_t(k) * ld] * __ldg(col_scale + k)));
  }
  int e = 0;
  float sc = 0.0f;  // zero / padding / non-finite rows quantise to all-zero digits
  if (mx > 0.0f && mx < 3.0e38f) {
    int ex;
    frexpf(mx, &ex);        // mx = f 2^ex, f in [0.5, 1)  =>  mx 2^-(ex-6) in [32, 64)
    e = ex - 6;
    sc = exp2_int(24 - e);  // v 2^24 with v = x c 2^-e
  }
  row_exp[m] = e;
  const int64_t plane_stride = ld * int64_t(kp);
  int8_t* dst = planes + m * int64_t(kp);
  for (int k0 = 0; k0 < kp; k0 += 16) {
    uint32_t w[I8_DIGITS][4];
#pragma unroll
    for (int d = 0; d < I8_DIGITS; ++d)
#pragma unroll
      for (int q = 0; q < 4; ++q) w[d][q] = 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int k = k0 + j;
      float v = 0.0f;
      if (k < K && sc != 0.0f) v = (x[int64_t(k) * ld] * __ldg(col_scale + k)) * sc;
      int I = __float2int_rn(v);        // |I| <= 2^30
      const int d3 = (I << 24) >> 24;   // balanced digit: sign-extended low byte
      I = (I - d3) >> 8;
      const int d2 = (I << 24) >> 24;
      I = (I - d2) >> 8;
      const int d1 = (I << 24) >> 24;
      I = (I - d1) >> 8;                // leading digit, |I| <= 65
      const int sh = (j & 3) * 8;
      w[0][j >> 2] |= uint32_t(I & 0xff) << sh;
      w[1][j >> 2] |= uint32_t(d1 & 0xff) << sh;
      w[2][j >> 2] |= uint32_t(d2 & 0xff) << sh;
      w[3][j >> 2] |= uint32_t(d3 & 0xff) << sh;
    }
#pragma unroll
    for (int d = 0; d < I8_DIGITS; ++d)
      *reinterpret_cast<uint4*>(dst + d * plane_stride + k0) = make_uint4(w[d][0], w[d][1], w[d][2], w[d][3]);
  }
}

// planes [4][rows][kp] int8 as a 3-d tensor {kp, rows, 4}; box {64 bytes, 128 rows, 1 plane}, SWIZZLE_64B
int make_plane_tmap(CUtensorMap* tm, const I8Planes& p) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return APAX_ERR_CUDA;
  const cuuint64_t dims[3] = {cuuint64_t(p.kp), cuuint64_t(p.rows), cuuint64_t(I8_DIGITS)};
  const cuuint64_t strides[2] = {cuuint64_t(p.kp), cuuint64_t(p.plane_stride)};
  const cuuint32_t box[3] = {I8_KC, I8_BM, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(p.base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? APAX_OK : APAX_ERR_INVALID_ARG;
}

}  // namespace

int i8_quantize_rows(cudaStream_t stream, const float* xt, int64_t ld, int64_t n_valid_rows, int K, int kp,
                     const float* col_scale, int8_t* planes, int32_t* row_exp) {
  if (ld % 128 != 0 || kp % 64 != 0 || K > kp) return APAX_ERR_INVALID_ARG;
  APAX_LAUNCH(k_i8_quantize_rows, dim3((unsigned)(ld / 128)), dim3(128), 0, stream, xt, ld, n_valid_rows, K, kp, col_scale,
              planes, row_exp);
  return APAX_OK;
}

void i8_quantize_weights(const float* W, int K, int NC, int ldw, int kp, int ncp, std::vector<int8_t>& planes,
                         std::vector<float>& col_scale) {
  planes.assign(size_t(I8_DIGITS) * ncp * kp, 0);
  col_scale.assign(size_t(kp), 1.0f);
  const size_t plane_stride = size_t(ncp) * kp;
  for (int k = 0; k < K; ++k) {
    double mx = 0.0;
    for (int n = 0; n < NC; ++n) mx = std::fmax(mx, std::fabs(double(W[size_t(k) * ldw + n])));
    if (!(mx > 0.0) || !std::isfinite(mx)) continue;
    int ex;
    std::frexp(mx, &ex);                       // mx / 2^(ex-6) in [32, 64)
    const double c = std::ldexp(1.0, ex - 6);
    col_scale[k] = float(c);                   // power of two: exact in fp32, exact to multiply by
    for (int n = 0; n < NC; ++n) {
      long long I = std::llrint(double(W[size_t(k) * ldw + n]) / c * 16777216.0);
      int d[I8_DIGITS];
      for (int s = I8_DIGITS - 1; s > 0; --s) {
        const int lo = int(((I & 0xff) ^ 0x80) - 0x80);  // sign-extended low byte
        d[s] = lo;
        I = (I - lo) >> 8;
      }
      d[0] = int(I);
      for (int s = 0; s < I8_DIGITS; ++s) planes[s * plane_stride + size_t(n) * kp + k] = int8_t(d[s]);
    }
  }
}

int i8_gemm(cudaStream_t stream, const I8Planes& act, const I8Planes& wgt, const I8Epilogue& epi, int32_t* err_flag) {
  if (act.kp != wgt.kp || act.kp % I8_KC != 0 || act.rows % I8_BM != 0 || wgt.rows % I8_BN != 0) return APAX_ERR_INVALID_ARG;
  CUtensorMap ta, tb;
  APAX_TRY(make_plane_tmap(&ta, act));
  APAX_TRY(make_plane_tmap(&tb, wgt));
  static bool attr_set = false;
  if (!attr_set) {
    APAX_CUDA_TRY(cudaFuncSetAttribute(k_i8_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, int(I8_SMEM_BYTES)));
    attr_set = true;
  }
  const int num_k_blocks = int(act.kp / I8_KC);
  const dim3 grid((unsigned)(act.rows / I8_BM), (unsigned)(wgt.rows / I8_BN));
  APAX_LAUNCH(k_i8_gemm, grid, dim3(128), I8_SMEM_BYTES, stream, ta, tb, num_k_blocks, epi, err_flag);
  return APAX_OK;
}

}  // namespace apax

// Test / measurement entry: out[nc][ld] = act^T . wgt through the integer tensor-core path (plain store
// epilogue).  act: device fp32 [k][ld]; wgt: device fp32 [k][nc].  Allocates its own scratch (test only).
extern "C" int apax_i8_gemm_test(apax_stream_t stream_, const float* act, int64_t k, int64_t ld, int64_t n_atoms,
                                 const float* wgt, int64_t nc, float* out, int32_t* err_flag) {
  if (!act || !wgt || !out || !err_flag || k <= 0 || nc <= 0 || ld % 128 != 0) return APAX_ERR_INVALID_ARG;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int kp = int((k + 63) / 64 * 64), ncp = int((nc + 127) / 128 * 128);
  std::vector<float> hw(size_t(k) * nc);
  APAX_CUDA_TRY(cudaStreamSynchronize(stream));
  APAX_CUDA_TRY(cudaMemcpy(hw.data(), wgt, hw.size() * sizeof(float), cudaMemcpyDeviceToHost));
  std::vector<int8_t> wp;
  std::vector<float> cs;
  apax::i8_quantize_weights(hw.data(), int(k), int(nc), int(nc), kp, ncp, wp, cs);
  int8_t *d_wp = nullptr, *d_ap = nullptr;
  float* d_cs = nullptr;
  int32_t* d_e = nullptr;
  int rc = APAX_OK;
  if (cudaMalloc(&d_wp, wp.size()) != cudaSuccess || cudaMalloc(&d_ap, size_t(4) * ld * kp) != cudaSuccess ||
      cudaMalloc(&d_cs, cs.size() * sizeof(float)) != cudaSuccess || cudaMalloc(&d_e, ld * sizeof(int32_t)) != cudaSuccess)
    rc = APAX_ERR_CUDA;
  if (rc == APAX_OK) {
    cudaMemcpy(d_wp, wp.data(), wp.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_cs, cs.data(), cs.size() * sizeof(float), cudaMemcpyHostToDevice);
    rc = apax::i8_quantize_rows(stream, act, ld, n_atoms, int(k), kp, d_cs, d_ap, d_e);
  }
  if (rc == APAX_OK) {
    apax::I8Planes a{d_ap, ld, kp, ld * int64_t(kp)}, w{d_wp, ncp, kp, int64_t(ncp) * kp};
    apax::I8Epilogue e{};
    e.mode = apax::I8_EPI_STORE;
    e.ld = ld;
    e.n_valid = int(nc);
    e.row_exp = d_e;
    e.out = out;
    rc = apax::i8_gemm(stream, a, w, e, err_flag);
  }
  if (cudaStreamSynchronize(stream) != cudaSuccess && rc == APAX_OK) rc = APAX_ERR_CUDA;
  cudaFree(d_wp);
  cudaFree(d_ap);
  cudaFree(d_cs);
  cudaFree(d_e);
  return rc;
}
