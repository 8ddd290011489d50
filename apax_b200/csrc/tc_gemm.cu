// tcgen05 / TMEM / TMA readout GEMM with 3xTF32 operand splitting.  See tc_gemm.cuh.
#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace apax {

namespace {

using namespace tcptx;

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), MN-major, SWIZZLE_128B:
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (stride between 128-byte MN chunks)
//   [32,46) stride byte offset >> 4 (stride between groups of 8 k-rows)   [46,48) version = 1
//   [61,64) layout type = 1 (SWIZZLE_128B_BASE32B): the only shared-memory layout the tensor core accepts for
//   MN-major 32-bit operands -- 128-byte rows, 32-byte swizzle granules, pattern period 4 k-rows
//   (cute::UMMA::Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> over 128 B x 4 rows); TMA writes it with
//   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return uint64_t((addr >> 4) & 0x3fffu) | (uint64_t((lbo_bytes >> 4) & 0x3fffu) << 16) |
         (uint64_t((sbo_bytes >> 4) & 0x3fffu) << 32) | (uint64_t(1) << 46) | (uint64_t(1) << 61);
}

// cute::UMMA::InstrDescriptor: c_format F32 (1) at [4,6), a/b format TF32 (2) at [7,10)/[10,13),
// a_major / b_major = MN (1) at bits 15 / 16, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

constexpr float kSwishTc = 1.6765324703310907f;

template <int BN>
struct TcShape {
  static constexpr uint32_t A_BYTES = TC_BM * TC_BK * 4;           // one of hi / lo
  static constexpr uint32_t B_BYTES = BN * TC_BK * 4;
  static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr uint32_t SMEM_BYTES = TC_STAGES * STAGE_BYTES + 1024;  // + slack for 1024-byte alignment
  static constexpr uint32_t CHUNK_STRIDE = TC_BK * 128;            // one 32-element chunk: 16 k-rows x 128 B
};

template <int BN>
__global__ void __launch_bounds__(128) k_tc_gemm(const __grid_constant__ CUtensorMap tm_a_hi,
                                                 const __grid_constant__ CUtensorMap tm_a_lo,
                                                 const __grid_constant__ CUtensorMap tm_b_hi,
                                                 const __grid_constant__ CUtensorMap tm_b_lo, int num_k_blocks,
                                                 TcEpilogue epi, int32_t* err) {
  using S = TcShape<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int m0 = blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(uint32_t(BN))
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
      const int s = kb % TC_STAGES;
      const uint32_t phase = (kb / TC_STAGES) & 1;
      if (!mbar_wait(smem_u32(&empty_bar[s]), phase ^ 1, err)) break;
      const uint32_t stage = smem_base + s * S::STAGE_BYTES;
      const uint32_t bar = smem_u32(&full_bar[s]);
      mbar_expect_tx(bar, S::STAGE_BYTES);
      const int k0 = kb * TC_BK;
      tma_load_3d(stage, &tm_a_hi, bar, 0, k0, m0 / 32);
      tma_load_3d(stage + S::A_BYTES, &tm_a_lo, bar, 0, k0, m0 / 32);
      tma_load_3d(stage + 2 * S::A_BYTES, &tm_b_hi, bar, 0, k0, n0 / 32);
      tma_load_3d(stage + 2 * S::A_BYTES + S::B_BYTES, &tm_b_lo, bar, 0, k0, n0 / 32);
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    uint32_t idesc = make_idesc(TC_BM, BN);
    if (epi.dbg_flags & 1) idesc &= ~((1u << 15) | (1u << 16));
    // MN chunk (32 atoms / features) stride, and stride between groups of 4 k-rows
    const uint32_t lbo = (epi.dbg_flags & 2) ? 512u : S::CHUNK_STRIDE;
    const uint32_t sbo = (epi.dbg_flags & 2) ? S::CHUNK_STRIDE : 512u;
    const uint64_t desc_mask = (epi.dbg_flags & 4) ? ~(uint64_t(7) << 61) : ~uint64_t(0);
    bool ok = true;
    for (int kb = 0; kb < num_k_blocks && ok; ++kb) {
      const int s = kb % TC_STAGES;
      const uint32_t phase = (kb / TC_STAGES) & 1;
      ok = mbar_wait(smem_u32(&full_bar[s]), phase, err);
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t stage = smem_base + s * S::STAGE_BYTES;
      if (epi.dbg && kb == 0) {
        for (uint32_t i = 0; i < S::STAGE_BYTES / 4 && i < 16384; ++i) {
          float v;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(stage + i * 4));
          epi.dbg[i] = v;
        }
        epi.dbg[16384] = __uint_as_float(tmem_d);
        epi.dbg[16385] = __uint_as_float(stage);
      }
#pragma unroll
      for (int g = 0; g < TC_BK / 8; ++g) {
        const uint32_t koff = g * 1024;  // 8 k-rows x 128 B
        const uint64_t a_hi = make_smem_desc(stage + koff, lbo, sbo) & desc_mask;
        const uint64_t a_lo = make_smem_desc(stage + S::A_BYTES + koff, lbo, sbo) & desc_mask;
        const uint64_t b_hi = make_smem_desc(stage + 2 * S::A_BYTES + koff, lbo, sbo) & desc_mask;
        const uint64_t b_lo = make_smem_desc(stage + 2 * S::A_BYTES + S::B_BYTES + koff, lbo, sbo) & desc_mask;
        umma_tf32(tmem_d, a_hi, b_hi, idesc, (kb | g) != 0 ? 1u : 0u);
        umma_tf32(tmem_d, a_hi, b_lo, idesc, 1u);
        umma_tf32(tmem_d, a_lo, b_hi, idesc, 1u);
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
    const uint32_t taddr = tmem_d + (uint32_t(warp * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      float v[16];
      tmem_ld16(taddr + c, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int n = n0 + c + i;
        const int64_t o = int64_t(n) * epi.ld + m;
        if (epi.mode == TC_EPI_STORE) {
          epi.out_hi[o] = v[i];
        } else if (epi.mode == TC_EPI_ACT) {
          const float y = v[i] + __ldg(epi.bias + n);
          const float sg = 1.0f / (1.0f + expf(-y));
          const float h = kSwishTc * y * sg;                        // variance_preserving_swish (activation.py:7-9)
          epi.aux[o] = kSwishTc * sg * (1.0f + y * (1.0f - sg));    // its derivative, for the backward pass
          float hi, lo;
          split_tf32(h, hi, lo);
          epi.out_hi[o] = hi;
          epi.out_lo[o] = lo;
        } else {
          const float t = v[i] * epi.aux[o];
          float hi, lo;
          split_tf32(t, hi, lo);
          epi.out_hi[o] = hi;
          epi.out_lo[o] = lo;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(uint32_t(BN)) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host side
// [rows][cols] fp32 row-major viewed as {32 elements, rows, cols/32}; box {32, 16 rows, box_chunks}
int make_tmap(CUtensorMap* tm, const float* base, int64_t rows, int64_t cols, int box_chunks) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return APAX_ERR_CUDA;
  const cuuint64_t dims[3] = {32, cuuint64_t(rows), cuuint64_t(cols / 32)};
  const cuuint64_t strides[2] = {cuuint64_t(cols) * 4, 128};
  const cuuint32_t box[3] = {32, TC_BK, cuuint32_t(box_chunks)};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? APAX_OK : APAX_ERR_INVALID_ARG;
}

template <int BN>
int launch_tc(cudaStream_t stream, const TcOperand& act, const TcOperand& wgt, int64_t n_atoms, const TcEpilogue& epi,
              int32_t* err_flag) {
  using S = TcShape<BN>;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  APAX_TRY(make_tmap(&ta_hi, act.hi, act.rows, act.cols, TC_BM / 32));
  APAX_TRY(make_tmap(&ta_lo, act.lo, act.rows, act.cols, TC_BM / 32));
  APAX_TRY(make_tmap(&tb_hi, wgt.hi, wgt.rows, wgt.cols, BN / 32));
  APAX_TRY(make_tmap(&tb_lo, wgt.lo, wgt.rows, wgt.cols, BN / 32));
  static bool attr_set = false;
  if (!attr_set) {
    APAX_CUDA_TRY(cudaFuncSetAttribute(k_tc_gemm<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(S::SMEM_BYTES)));
    attr_set = true;
  }
  const int num_k_blocks = int(cdiv(act.rows, TC_BK));
  const dim3 grid((unsigned)cdiv(n_atoms, TC_BM), (unsigned)(wgt.cols / BN));
  APAX_LAUNCH(k_tc_gemm<BN>, grid, dim3(128), S::SMEM_BYTES, stream, ta_hi, ta_lo, tb_hi, tb_lo, num_k_blocks, epi,
              err_flag);
  return APAX_OK;
}

}  // namespace

int tc_gemm(cudaStream_t stream, const TcOperand& act, const TcOperand& wgt, int64_t n_atoms, const TcEpilogue& epi,
            int32_t* err_flag) {
  if (act.rows != wgt.rows || act.cols % 128 != 0 || wgt.cols % 128 != 0) return APAX_ERR_INVALID_ARG;
  if (wgt.cols % 256 == 0) return launch_tc<256>(stream, act, wgt, n_atoms, epi, err_flag);
  return launch_tc<128>(stream, act, wgt, n_atoms, epi, err_flag);
}

}  // namespace apax

// Test / measurement entry: out[NC][ld] = act^T . wgt through the tensor-core path (plain store epilogue).
extern "C" int apax_tc_gemm_test(apax_stream_t stream, const float* act_hi, const float* act_lo, int64_t k, int64_t ld,
                                 int64_t n_atoms, const float* wgt_hi, const float* wgt_lo, int64_t nc, float* out,
                                 int32_t* err_flag, float* dbg, int32_t dbg_flags) {
  if (!act_hi || !act_lo || !wgt_hi || !wgt_lo || !out || !err_flag) return APAX_ERR_INVALID_ARG;
  apax::TcOperand a{act_hi, act_lo, k, ld}, w{wgt_hi, wgt_lo, k, nc};
  apax::TcEpilogue e{};
  e.mode = apax::TC_EPI_STORE;
  e.ld = ld;
  e.out_hi = out;
  e.dbg = dbg;
  e.dbg_flags = dbg_flags;
  return apax::tc_gemm(stream, a, w, n_atoms, e, err_flag);
}
