// Exact fixed-point readout GEMM on tcgen05.mma kind::i8.  See i8_gemm.cuh.
#include "i8_gemm.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "tc_ptx.cuh"

namespace apax {

namespace {

using namespace tcptx;

constexpr uint32_t I8_A_TILE = I8_BM * I8_KC;                  // 8 KB: 128 atoms x 64 bytes of one plane
constexpr uint32_t I8_B_TILE = I8_BN * I8_KC;                  // 4 KB: 64 outputs x 64 bytes of one plane
constexpr uint32_t I8_B_STAGE = I8_DIGITS * I8_B_TILE;         // 16 KB: one k-block of the four weight planes
constexpr uint32_t I8_ACC_COLS = I8_DIGITS * I8_BN;            // 256 columns: one accumulator set (orders 0..3)
constexpr uint32_t I8_TMEM_COLS = 2 * I8_ACC_COLS;             // two sets = the SM's whole tensor memory
constexpr int I8_MAX_B_STAGES = 8;
// Warp roles.  W = 0 (352 threads): warp 0 TMA, warps 1 and 10 MMA issuers, warps 2-9 epilogue (two per TMEM lane quarter, 32
// output columns each).  W = 1 (640 threads, "wide epilogue"): the epilogue is what bounds these GEMMs (66 instructions per
// output element issued by two warps per scheduler at 0.4 IPC), so sixteen epilogue warps (four per
// quarter, 16 columns each) in four aligned warp groups, the control warps (0 TMA, 1 and 2 MMA, 3 idle) in a group of their
// own, and setmaxnreg moves registers from the control group (96 -> 32) to the epilogue groups (96 -> 112).
// MEASURED (1M atoms, four GEMMs, profiles/r02p_gemm_wide_epilogue.txt): 8 warps 4.14 ms; 16 warps at the 96 registers a
// 640-thread CTA gets, no setmaxnreg: 4.74 ms (the epilogue spills); with setmaxnreg 32 / 112 (the register pool is only what
// the CTA's own warps release, so 112 for the epilogue forces the control group down to 32): 10.0 ms -- the MMA-issuing warp
// spills on the critical path.  The register file cannot hold sixteen epilogue warps of this epilogue; W = 0 ships.
template <int W>
struct I8Roles {
  static constexpr int kEpiBase = W ? 4 : 2;                   // first epilogue warp
  static constexpr int kEpiWarps = W ? 16 : 8;
  static constexpr int kEpiThreads = 32 * kEpiWarps;
  static constexpr int kMma2 = W ? 2 : 10;                     // second MMA-issuing warp (odd tiles)
  static constexpr int kThreads = W ? 640 : 352;
  static constexpr int kParts = kEpiWarps / 4;                 // warps sharing a lane quarter split the 64 columns
  static constexpr int kPW = I8_BN / kParts;                   // columns per epilogue thread and sub-tile: 32 or 16
};
constexpr int I8_THREADS = I8Roles<0>::kThreads;
constexpr uint32_t I8_SMEM_LIMIT = 232448;                     // 227 KB per CTA
constexpr float kSwishI8 = 1.6765324703310907f;                // apax/layers/activation.py:8

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

// 1 / x for finite x >= 1: SFU approximation (1 ulp) + one Newton step
__device__ __forceinline__ float rcp_newton(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.0f), r);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// one elected lane of a fully converged warp (the MMA / TMA warps run warp-uniform code so that descriptors and
// loop state stay in uniform registers; only the issuing instructions are predicated on the elected lane)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// whole-warp wait with a warp-uniform result
__device__ __forceinline__ bool mbar_wait_warp(uint32_t bar, uint32_t parity, int32_t* err) {
  return __all_sync(0xffffffffu, mbar_wait(bar, parity, err));
}
// Bounded wait that never changes control flow: the TMA and MMA warps must stay provably warp-uniform (descriptors
// in uniform registers, no per-MMA register -> uniform-register moves), so a time-out only raises the error flag
// and marks the CTA dead, after which every later wait returns at once and the kernel drains quickly.
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_soft(uint32_t bar, uint32_t parity, int32_t* err, volatile int* dead) {
  if (*dead) return;
  // non-blocking polls: the pipeline hand-offs are on the critical path, a suspended try_wait wakes up too late
  for (uint32_t it = 0; it < (kSpinLimit << 4); ++it)
    if (mbar_test_wait(bar, parity)) return;
  atomicExch(err, 1);
  *dead = 1;
}
template <int NT>
__device__ __forceinline__ void epi_bar_sync() {   // named barrier 1: the epilogue threads only
  asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

// All MMAs of up to two 64-byte k-blocks as ONE instruction burst.  The tensor pipe queues only an MMA or two, so
// whatever the issuing warp does between MMAs (barrier polls, descriptor arithmetic, commits) is time the pipe
// idles: every operand is therefore an input of a single asm statement (descriptors sit in uniform registers before
// the burst starts), two k-blocks share one hand-shake, and the burst ends with its widest MMAs (N = 256, 128
// cycles each) so that the pipe still has work while the warp goes through the next hand-shake.
//   activation digit s (descriptor a_s) x weight planes 0 .. 3-s (descriptor b, N = 64 (4 - s)) -> accumulators s..3
//   order: a0 of the first k-slice (it may zero-initialise all four accumulators), then digits 3, 2, 1, then the rest of a0
__device__ __forceinline__ void umma_i8_burst(uint32_t tm0, uint32_t tm1, uint32_t tm2, uint32_t tm3, uint64_t a00,
                                              uint64_t a01, uint64_t a02, uint64_t a03, uint64_t a10, uint64_t a11,
                                              uint64_t a12, uint64_t a13, uint64_t b0, uint64_t b1, uint32_t id256,
                                              uint32_t id192, uint32_t id128, uint32_t id64, uint32_t acc_first,
                                              uint32_t have3, uint32_t have_kb1) {
  asm volatile(
      "{\n\t"
      ".reg .pred pacc, p3, p1, p31, pt;\n\t"
      ".reg .b64 a0b, a1b, a2b, a3b, a0c, a1c, a2c, a3c, a0d, a1d, a2d, a3d, b0b, b1b;\n\t"
      "setp.ne.b32 pacc, %18, 0;\n\t"
      "setp.ne.b32 p3, %19, 0;\n\t"
      "setp.ne.b32 p1, %20, 0;\n\t"
      "and.pred p31, p3, p1;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "add.s64 a0b, %4, 2;\n\t"
      "add.s64 a1b, %5, 2;\n\t"
      "add.s64 a2b, %6, 2;\n\t"
      "add.s64 a3b, %7, 2;\n\t"
      "add.s64 a0d, %8, 2;\n\t"
      "add.s64 a1d, %9, 2;\n\t"
      "add.s64 a2d, %10, 2;\n\t"
      "add.s64 a3d, %11, 2;\n\t"
      "add.s64 b0b, %12, 2;\n\t"
      "add.s64 b1b, %13, 2;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %4, %12, %14, pacc;\n\t"
      "@p3 tcgen05.mma.cta_group::1.kind::i8 [%3], %7, %12, %17, pt;\n\t"
      "@p3 tcgen05.mma.cta_group::1.kind::i8 [%3], a3b, b0b, %17, pt;\n\t"
      "@p31 tcgen05.mma.cta_group::1.kind::i8 [%3], %11, %13, %17, pt;\n\t"
      "@p31 tcgen05.mma.cta_group::1.kind::i8 [%3], a3d, b1b, %17, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%2], %6, %12, %16, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%2], a2b, b0b, %16, pt;\n\t"
      "@p1 tcgen05.mma.cta_group::1.kind::i8 [%2], %10, %13, %16, pt;\n\t"
      "@p1 tcgen05.mma.cta_group::1.kind::i8 [%2], a2d, b1b, %16, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%1], %5, %12, %15, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%1], a1b, b0b, %15, pt;\n\t"
      "@p1 tcgen05.mma.cta_group::1.kind::i8 [%1], %9, %13, %15, pt;\n\t"
      "@p1 tcgen05.mma.cta_group::1.kind::i8 [%1], a1d, b1b, %15, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], a0b, b0b, %14, pt;\n\t"
      "@p1 tcgen05.mma.cta_group::1.kind::i8 [%0], %8, %13, %14, pt;\n\t"
      "@p1 tcgen05.mma.cta_group::1.kind::i8 [%0], a0d, b1b, %14, pt;\n\t"
      "}"
      ::"r"(tm0), "r"(tm1), "r"(tm2), "r"(tm3), "l"(a00), "l"(a01), "l"(a02), "l"(a03), "l"(a10), "l"(a11), "l"(a12),
      "l"(a13), "l"(b0), "l"(b1), "r"(id256), "r"(id192), "r"(id128), "r"(id64), "r"(acc_first), "r"(have3), "r"(have_kb1)
      : "memory");
}

// sixteen values (already multiplied by their column scales) -> four 16-byte plane pieces at planes + p * stride
__device__ __forceinline__ void i8_emit16(const float (&v)[16], float sc, float sc2, int8_t* dst, int64_t plane_stride) {
  uint32_t w[I8_DIGITS][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t ew[4], tb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ew[j] = i8_digit_word<4>((v[4 * q + j] * sc) * sc2);
    i8_transpose4(ew, tb);
#pragma unroll
    for (int d = 0; d < I8_DIGITS; ++d) w[d][q] = tb[3 - d];
  }
#pragma unroll
  for (int d = 0; d < I8_DIGITS; ++d)
    __stcs(reinterpret_cast<uint4*>(dst + d * plane_stride), make_uint4(w[d][0], w[d][1], w[d][2], w[d][3]));
}

__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
// thirty-two values (already multiplied by their column scales) -> one full 32-byte sector per plane (256-bit stores:
// a thread owns 32 consecutive outputs of its atom, i.e. exactly one sector of each plane's row)
__device__ __forceinline__ void i8_emit32(const float (&v)[32], float sc, float sc2, int8_t* dst, int64_t plane_stride) {
  uint32_t w[I8_DIGITS][8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint32_t ew[4], tb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ew[j] = i8_digit_word<4>((v[4 * q + j] * sc) * sc2);
    i8_transpose4(ew, tb);
#pragma unroll
    for (int d = 0; d < I8_DIGITS; ++d) w[d][q] = tb[3 - d];
  }
#pragma unroll
  for (int d = 0; d < I8_DIGITS; ++d) st_global_v8(dst + d * plane_stride, w[d]);
}

template <int N>
__device__ __forceinline__ void i8_emit_cols(const float (&v)[N], float sc, float sc2, int8_t* dst, int64_t plane_stride) {
  static_assert(N == 16 || N == 32, "16 or 32 columns per epilogue thread");
  if constexpr (N == 32) i8_emit32(v, sc, sc2, dst, plane_stride); else i8_emit16(v, sc, sc2, dst, plane_stride);
}

// int32 accumulators of the four orders -> fp32.  Order 0 (|.| < 2^22) and the rounded-down tails of orders 2 and 3
// go through the exact magic-number conversion on the ALU pipes, order 1 through I2F; the shifts drop at most
// 2^-15 of an order-0 unit, far below the fp32 rounding of the result.
__device__ __forceinline__ float i8_combine(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  constexpr float kMagic = 12582912.0f;   // 1.5 * 2^23
  const float f0 = __int_as_float(int(r0) + 0x4B400000) - kMagic;
  const float f1 = float(int(r1));
  const float f2 = __int_as_float(((int(r2) + 2) >> 2) + 0x4B400000) - kMagic;   // units of 4
  const float f3 = __int_as_float(((int(r3) + 8) >> 4) + 0x4B400000) - kMagic;   // units of 16
  float t = fmaf(f3, 16.0f / 256.0f / 4.0f, f2);   // in units of 4 * 256^-2
  t = fmaf(t, 4.0f / 256.0f, f1);                   // in units of 256^-1
  return fmaf(t, 1.0f / 256.0f, f0);
}

// ================================================================================================
// The GEMM.  MODE: epilogue (I8_EPI_*), NOUT: heads accumulated by I8_EPI_ACT_HEAD (1 or I8_MAX_HEADS).
// ================================================================================================
template <int MODE, int NOUT, int W = 0>
__global__ void __launch_bounds__(I8Roles<W>::kThreads, 1) k_i8_gemm(const __grid_constant__ CUtensorMap tm_a,
                                                           const __grid_constant__ CUtensorMap tm_b, int num_kb,
                                                           int a_planes, int a_res, int num_m_tiles, int num_n_sub,
                                                           int b_stages, I8Epilogue epi, int32_t* err) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[I8_MAX_KB];
  __shared__ __align__(8) uint64_t a_empty[I8_MAX_KB];
  __shared__ __align__(8) uint64_t b_full[I8_MAX_B_STAGES];
  __shared__ __align__(8) uint64_t b_empty[I8_MAX_B_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ int cta_dead;
  using R = I8Roles<W>;
  constexpr int PARTS = R::kParts, PW = R::kPW;
  __shared__ float xmax[PARTS][I8_BM];                   // row-maximum exchange between the column parts
  __shared__ float xhead[(PARTS - 1) * I8_BM * NOUT];    // head partial sums of the parts 1 ..
  // per-output constants of the epilogue (bias, next layer's feature scales, head weights for one head): broadcast
  // shared-memory reads instead of global loads inside the element loop
  __shared__ float s_bias[I8_MAX_KB * I8_KC], s_qcs[I8_MAX_KB * I8_KC], s_whead[I8_MAX_KB * I8_KC];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Two MMA-issuing warps (even / odd tiles) when the weight ring can be split into two private rings of at least two
  // stages: a ring must have ONE consumer walking it in order, or a parity wait could alias a phase that is a whole
  // revolution away.
  const bool dual = b_stages >= 4 && num_n_sub >= 2;
  const uint32_t ring_size = dual ? uint32_t(b_stages) / 2 : uint32_t(b_stages);
  const uint32_t smem_a = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // a_res planes of the atom tile are resident; the others travel with the weight stages (they are used least)
  const uint32_t smem_b = smem_a + uint32_t(num_kb * a_res) * I8_A_TILE;
  const uint32_t stage_bytes = I8_B_STAGE + uint32_t(a_planes - a_res) * I8_A_TILE;

  for (int n = threadIdx.x; n < num_n_sub * I8_BN; n += R::kThreads) {
    s_bias[n] = (MODE == I8_EPI_ACT_Q || MODE == I8_EPI_ACT_HEAD) ? epi.bias[n] : 0.0f;
    s_qcs[n] = MODE != I8_EPI_STORE ? epi.q_col_scale[n] : 0.0f;
    s_whead[n] = (MODE == I8_EPI_ACT_HEAD && NOUT == 1) ? epi.w_head[n] : 0.0f;
  }
  if (threadIdx.x == 0) {
    cta_dead = 0;
    for (int i = 0; i < I8_MAX_KB; ++i) {
      mbar_init(smem_u32(&a_full[i]), 1);
      mbar_init(smem_u32(&a_empty[i]), dual ? 2 : 1);   // with two MMA warps both sweep every atom tile
    }
    for (int i = 0; i < I8_MAX_B_STAGES; ++i) {
      mbar_init(smem_u32(&b_full[i]), 1);
      mbar_init(smem_u32(&b_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), R::kEpiThreads);
    }
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
  if (W) {   // registers follow the work: control warp group down, epilogue warp groups up (whole aligned groups of four warps)
    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");   // frees 4 x 32 x 64 registers ...
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory");  // ... exactly what 16 x 32 x 16 more need (the pool is the CTA's own)
    }
  }
  const uint32_t tmem_d = tmem_base_slot;
  const int my_m_tiles = (num_m_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform, elected lane issues)
    uint32_t cnt[2] = {0, 0};   // k-blocks sent into each ring
    for (int mt = 0; mt < my_m_tiles; ++mt) {
      const int m0 = (int(blockIdx.x) + mt * int(gridDim.x)) * I8_BM;
      for (int j = 0; j < num_n_sub; ++j) {
        const uint32_t w = dual ? uint32_t(mt * num_n_sub + j) & 1u : 0u;   // the ring of the warp that owns this tile
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint32_t bit = w == 0 ? cnt[0]++ : cnt[1]++;
          if (j == 0) {   // the atom tile's planes of this k-block: resident for the whole output sweep
            mbar_wait_soft(smem_u32(&a_empty[kb]), (uint32_t(mt) & 1) ^ 1, err, &cta_dead);
            if (epi.dbg_flags & 2) {
              if (elect_one()) mbar_arrive(smem_u32(&a_full[kb]));
            } else if (elect_one()) {
              const uint32_t bar = smem_u32(&a_full[kb]);
              mbar_expect_tx(bar, uint32_t(a_res) * I8_A_TILE);
              for (int d = 0; d < a_res; ++d)
                tma_load_3d(smem_a + uint32_t(kb * a_res + d) * I8_A_TILE, &tm_a, bar, kb * I8_KC, m0, d);
            }
            __syncwarp();
          }
          const uint32_t s = w * ring_size + bit % ring_size, phase = (bit / ring_size) & 1;
          mbar_wait_soft(smem_u32(&b_empty[s]), phase ^ 1, err, &cta_dead);
          if (epi.dbg_flags & 2) {
            if (elect_one()) mbar_arrive(smem_u32(&b_full[s]));
          } else if (elect_one()) {
            const uint32_t bar = smem_u32(&b_full[s]);
            mbar_expect_tx(bar, stage_bytes);
#pragma unroll
            for (int d = 0; d < I8_DIGITS; ++d)
              tma_load_3d(smem_b + s * stage_bytes + d * I8_B_TILE, &tm_b, bar, kb * I8_KC, j * I8_BN, d);
            for (int d = a_res; d < a_planes; ++d)
              tma_load_3d(smem_b + s * stage_bytes + I8_B_STAGE + uint32_t(d - a_res) * I8_A_TILE, &tm_a, bar, kb * I8_KC, m0, d);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1 || warp == R::kMma2) {
    // ------------------------------------------------------------------ MMA issuers (warp-uniform, elected lane issues)
    // Two warps: one owns the even tiles (accumulator set 0), the other the odd ones (set 1).  The tensor pipe queues
    // only an MMA or two, so a single issuer leaves it idle during every hand-shake (barrier polls, descriptors,
    // commits); with two, one warp's burst executes while the other goes through its hand-shake.  Each warp's MMAs
    // touch only its own accumulator set and tcgen05.commit tracks the issuing thread's MMAs, so no ordering between
    // the two is needed; the atom tile's slots are released when both have committed (barrier count 2).
    const uint32_t mma_id = warp == 1 ? 0u : 1u;
    const uint32_t total_tiles = (dual || mma_id == 0) ? uint32_t(my_m_tiles) * uint32_t(num_n_sub) : 0u;   // single mode: warp 1 only
    const uint32_t tile_step = dual ? 2u : 1u;
    uint32_t bit = 0;   // k-blocks this warp has taken from its ring
    const long long t_begin = clock64();
    long long w_acc = 0, w_b = 0, w_a = 0;   // cycle counters (trace only): waits for a drained accumulator set, weight stages, atom planes
    // 16-byte-unit offsets of the four activation planes: resident ones relative to the k-block's resident slot,
    // streamed ones relative to the weight stage
    uint32_t a_off[I8_DIGITS];
#pragma unroll
    for (int sa = 0; sa < I8_DIGITS; ++sa)
      a_off[sa] = sa < a_res ? uint32_t(sa) * (I8_A_TILE >> 4) : (I8_B_STAGE >> 4) + uint32_t(sa - a_res) * (I8_A_TILE >> 4);
    for (uint32_t tile = mma_id; tile < total_tiles; tile += tile_step) {
      {
        const int mt = int(tile / uint32_t(num_n_sub));
        const bool first_in_mt = tile < tile_step || int((tile - tile_step) / uint32_t(num_n_sub)) != mt;   // this warp's first sweep of the atom tile
        const bool last_in_mt = int((tile + tile_step) / uint32_t(num_n_sub)) != mt;                        // ... and its last
        const uint32_t buf = tile & 1, use = tile >> 1;
        long long c0 = clock64();
        mbar_wait_soft(smem_u32(&acc_empty[buf]), (use & 1) ^ 1, err, &cta_dead);   // epilogue has drained this set
        w_acc += clock64() - c0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc_base = tmem_d + buf * I8_ACC_COLS;
        for (int kb = 0; kb < num_kb; kb += 2, bit += 2) {
          const bool two = kb + 1 < num_kb;   // k-blocks are issued in pairs (one hand-shake per sixteen MMAs)
          const uint32_t ring0 = dual ? mma_id * ring_size : 0u;
          const uint32_t s0 = ring0 + bit % ring_size, ph0 = (bit / ring_size) & 1;
          const uint32_t s1 = ring0 + (bit + 1) % ring_size, ph1 = ((bit + 1) / ring_size) & 1;
          c0 = clock64();
          if (first_in_mt) {
            mbar_wait_soft(smem_u32(&a_full[kb]), uint32_t(mt) & 1, err, &cta_dead);
            if (two) mbar_wait_soft(smem_u32(&a_full[kb + 1]), uint32_t(mt) & 1, err, &cta_dead);
          }
          const long long c1 = clock64();
          mbar_wait_soft(smem_u32(&b_full[s0]), ph0, err, &cta_dead);
          if (two) mbar_wait_soft(smem_u32(&b_full[s1]), ph1, err, &cta_dead);
          w_a += c1 - c0;
          w_b += clock64() - c1;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // Activation digit s meets the weight digits 0 .. 3-s in ONE MMA: the four weight planes of a stage are
          // contiguous 64-row blocks of one K-major tile, so N = 64 (4 - s) output columns are the products
          // a_s b_0 | a_s b_1 | ... and land on the accumulators of orders s, s+1, ..., 3.  Products of order above 3
          // (below 2^-32) are never formed.
          const uint64_t ar0 = make_desc_k_sw64(smem_a + uint32_t(kb * a_res) * I8_A_TILE);
          const uint64_t ar1 = ar0 + uint64_t((uint32_t(a_res) * I8_A_TILE) >> 4);
          const uint64_t b0 = make_desc_k_sw64(smem_b + s0 * stage_bytes);
          const uint64_t b1 = make_desc_k_sw64(smem_b + s1 * stage_bytes);
          uint64_t ad0[I8_DIGITS], ad1[I8_DIGITS];
#pragma unroll
          for (int sa = 0; sa < I8_DIGITS; ++sa) {
            ad0[sa] = (sa < a_res ? ar0 : b0) + uint64_t(a_off[sa]);
            ad1[sa] = (sa < a_res ? ar1 : b1) + uint64_t(a_off[sa]);
          }
          if (elect_one()) {
            umma_i8_burst(acc_base, acc_base + I8_BN, acc_base + 2 * I8_BN, acc_base + 3 * I8_BN, ad0[0], ad0[1], ad0[2], ad0[3],
                          ad1[0], ad1[1], ad1[2], ad1[3], b0, b1, make_idesc_i8(I8_BM, 4 * I8_BN),
                          make_idesc_i8(I8_BM, 3 * I8_BN), make_idesc_i8(I8_BM, 2 * I8_BN), make_idesc_i8(I8_BM, I8_BN),
                          kb != 0 ? 1u : 0u, a_planes > 3 ? 1u : 0u, two ? 1u : 0u);
            umma_commit(smem_u32(&b_empty[s0]));                 // weight stages free once these MMAs have read them
            if (two) umma_commit(smem_u32(&b_empty[s1]));
            if (last_in_mt) {                                    // this warp's last sweep over these atom k-blocks
              umma_commit(smem_u32(&a_empty[kb]));
              if (two) umma_commit(smem_u32(&a_empty[kb + 1]));
            }
            if (kb + 2 >= num_kb) umma_commit(smem_u32(&acc_full[buf]));
          }
          __syncwarp();
          if (!two) --bit;   // an odd last k-block advanced the ring by one stage only
        }
      }
    }
    if (epi.dbg && lane == 0 && mma_id == 0) {
      epi.dbg[blockIdx.x * 8 + 0] = clock64() - t_begin;
      epi.dbg[blockIdx.x * 8 + 1] = w_acc;
      epi.dbg[blockIdx.x * 8 + 2] = w_b;
      epi.dbg[blockIdx.x * 8 + 3] = w_a;
    }
  } else if (warp >= R::kEpiBase && warp < R::kEpiBase + R::kEpiWarps) {
    // -------------------------------------------------------------------- epilogue warps: lane == atom
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int half = (warp - R::kEpiBase) >> 2;   // column part of this warp: the warps of a quarter split the 64 columns
    const int row = q * 32 + lane;
    float* scratch = (MODE == I8_EPI_ACT_Q || MODE == I8_EPI_MULD_Q)
                         ? epi.scratch + int64_t(blockIdx.x) * num_n_sub * I8_BN * I8_BM + row
                         : nullptr;
    uint32_t tile = 0;
    bool ok = true;
    long long w_wait = 0, w_phase2 = 0, t_begin = clock64();
    for (int mt = 0; mt < my_m_tiles && ok; ++mt) {
      const int64_t m = int64_t(int(blockIdx.x) + mt * int(gridDim.x)) * I8_BM + row;
      const float scale = exp2_int(epi.row_exp ? epi.row_exp[m] : epi.fixed_exp);
      const float rmul = (MODE == I8_EPI_MULD_Q && epi.row_mul) ? epi.row_mul[m] : 1.0f;
      // swish' of layer 0 (written by ACT_Q, read by MULD_Q): tile-major [atom tile][output][128]
      float* aux_tile = (MODE == I8_EPI_ACT_Q || MODE == I8_EPI_MULD_Q)
                            ? epi.aux + int64_t(int(blockIdx.x) + mt * int(gridDim.x)) * (num_n_sub * I8_BN) * I8_BM + row
                            : nullptr;
      float mx = 0.0f;
      float hacc[NOUT];
#pragma unroll
      for (int k = 0; k < NOUT; ++k) hacc[k] = 0.0f;
      float qsc = 0.0f, qsc2 = 0.0f;
      if (MODE == I8_EPI_ACT_HEAD) {   // v 2^24 = d c 2^(24 - e) with the shared exponent
        const int tot = 24 - epi.q_fixed_exp, h = tot / 2;
        qsc = exp2_int(h);
        qsc2 = exp2_int(tot - h);
      }
      for (int j = 0; j < num_n_sub && ok; ++j, ++tile) {
        const uint32_t buf = tile & 1, use = tile >> 1;
        float auxv[PW];   // MULD_Q: the multipliers of this sub-tile, requested before the wait so that their latency hides
        if (MODE == I8_EPI_MULD_Q) {
          const float* ap = aux_tile + int64_t(j * I8_BN + half * PW) * I8_BM;   // one 64-bit base, immediate offsets below
#pragma unroll
          for (int i = 0; i < PW; ++i) auxv[i] = __ldcs(ap + i * I8_BM);
        }
        const long long c0 = clock64();
        ok = mbar_wait(smem_u32(&acc_full[buf]), use & 1, err);
        w_wait += clock64() - c0;
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_d + buf * I8_ACC_COLS + (uint32_t(q * 32) << 16);
        if (epi.dbg_flags & 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(smem_u32(&acc_empty[buf]));
          continue;
        }
        float dq[PW];                                       // ACT_HEAD: swish' values to be quantised
#pragma unroll
        for (int cc = 0; cc < PW / 16; ++cc) {
          const int c = half * PW + cc * 16;
          uint32_t r0[16], r1[16], r2[16], r3[16];          // sixteen columns of the four orders
          tmem_ld16_u32(taddr + c, r0);
          tmem_ld16_u32(taddr + I8_BN + c, r1);
          tmem_ld16_u32(taddr + 2 * I8_BN + c, r2);
          tmem_ld16_u32(taddr + 3 * I8_BN + c, r3);
          tmem_ld_wait();
          if (cc == PW / 16 - 1) {  // everything is in registers: hand the accumulator set back before the math
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(smem_u32(&acc_empty[buf]));
          }
          const int nb = j * I8_BN + c;                     // first output of this chunk
          float* gp = epi.out + int64_t(nb) * epi.ld + m;   // STORE: walks down the feature rows of the [n][ld] output
          // ACT_Q / MULD_Q: the derivative matrix and the staging rows are tile-major [atom tile][output][128 atoms], so
          // every element of the chunk is one 64-bit base plus a compile-time offset (with [n][ld] addressing the sweep
          // spent 7 of its 44 instructions per element on 64-bit address arithmetic)
          float* ap = MODE == I8_EPI_ACT_Q ? aux_tile + int64_t(nb) * I8_BM : nullptr;
          float* sp = (MODE == I8_EPI_ACT_Q || MODE == I8_EPI_MULD_Q) ? scratch + int64_t(nb) * I8_BM : nullptr;
          const bool all_valid = MODE != I8_EPI_STORE || nb + 16 <= epi.n_valid;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = nb + i, ri = cc * 16 + i;
            const float y = i8_combine(r0[i], r1[i], r2[i], r3[i]) * scale;
            if (MODE == I8_EPI_STORE) {
              if (all_valid || n < epi.n_valid) __stcs(gp, y);
            } else if (MODE == I8_EPI_ACT_Q || MODE == I8_EPI_ACT_HEAD) {
              // variance_preserving_swish (activation.py:7-9) and its derivative: SFU exponential (ex2.approx, 2 ulp
              // like libm's expf) and an SFU reciprocal refined by one Newton step (< 1 ulp) -- the epilogue is bound by
              // instruction issue, and libm expf + IEEE division cost three times as many instructions.
              const float z = y + s_bias[n];
              const float sg = rcp_newton(1.0f + __expf(fminf(-z, 80.0f)));
              const float cs = kSwishI8 * sg;
              const float h = cs * z;
              const float d = cs * fmaf(z, 1.0f - sg, 1.0f);
              if (MODE == I8_EPI_ACT_Q) {
                __stcs(ap + i * I8_BM, d);   // streaming (read once, by the backward GEMM): the staging rows own the L2
                sp[i * I8_BM] = h;
                mx = fmaxf(mx, fabsf(h * s_qcs[n]));
              } else {
                if (NOUT == 1) {
                  hacc[0] = fmaf(h, s_whead[n], hacc[0]);
                } else {
#pragma unroll
                  for (int k = 0; k < NOUT; ++k)
                    if (k < epi.n_out) hacc[k] = fmaf(h, __ldg(epi.w_head + n * epi.n_out + k), hacc[k]);
                }
                dq[ri] = d * s_qcs[n];
              }
            } else {
              const float t = y * auxv[ri] * rmul;
              sp[i * I8_BM] = t;
              mx = fmaxf(mx, fabsf(t * s_qcs[n]));
            }
            if (MODE == I8_EPI_STORE) gp += epi.ld;
          }
        }
        if (MODE == I8_EPI_ACT_HEAD)
          i8_emit_cols<PW>(dq, qsc, qsc2, epi.q_planes + m * epi.q_kp + j * I8_BN + half * PW, epi.q_plane_stride);
      }
      if (!ok) break;
      // ---- end of the output sweep of this atom tile
      if (MODE == I8_EPI_ACT_Q || MODE == I8_EPI_MULD_Q) {
        // This thread staged the columns [j*64 + half*PW, +PW) of every sub-tile itself: re-read them (L2) and emit
        // digits, one 32-byte sector (16 bytes with the wide epilogue) per plane and sub-tile.  The loads do not depend on
        // the row maximum, so the first batch is requested before the exchange.
        const long long p0 = clock64();
        float cur[PW];
        const float* rp = scratch + int64_t(half * PW) * I8_BM;
#pragma unroll
        for (int i = 0; i < PW; ++i) cur[i] = rp[i * I8_BM];
        xmax[half][row] = mx;
        epi_bar_sync<R::kEpiThreads>();
#pragma unroll
        for (int pp = 0; pp < PARTS; ++pp) mx = fmaxf(mx, xmax[pp][row]);
        epi_bar_sync<R::kEpiThreads>();
        int e;
        float sc, sc2;
        i8_row_scale(mx, 4, e, sc, sc2);
        if (half == 0) epi.q_row_exp[m] = e;
#pragma unroll 1
        for (int j = 0; j < num_n_sub; ++j) {
          const int nb = j * I8_BN + half * PW;
          float nxt[PW];                     // the next sub-tile's staged values travel while this one is converted
          if (j + 1 < num_n_sub) {
            rp += I8_BN * I8_BM;
#pragma unroll
            for (int i = 0; i < PW; ++i) nxt[i] = rp[i * I8_BM];
          }
#pragma unroll
          for (int i = 0; i < PW; ++i) cur[i] *= s_qcs[nb + i];
          i8_emit_cols<PW>(cur, sc, sc2, epi.q_planes + m * epi.q_kp + nb, epi.q_plane_stride);
          if (j + 1 < num_n_sub) {
#pragma unroll
            for (int i = 0; i < PW; ++i) cur[i] = nxt[i];
          }
        }
        w_phase2 += clock64() - p0;
      } else if (MODE == I8_EPI_ACT_HEAD) {
        if (half > 0) {
#pragma unroll
          for (int k = 0; k < NOUT; ++k) xhead[((half - 1) * I8_BM + row) * NOUT + k] = hacc[k];
        }
        epi_bar_sync<R::kEpiThreads>();
        if (half == 0) {
#pragma unroll
          for (int k = 0; k < NOUT; ++k)
            if (k < epi.n_out) {
              float h = hacc[k];
#pragma unroll
              for (int pp = 1; pp < PARTS; ++pp) h += xhead[((pp - 1) * I8_BM + row) * NOUT + k];
              epi.h_out[int64_t(k) * epi.ld + m] = h;
            }
        }
        epi_bar_sync<R::kEpiThreads>();
      }
    }
    if (epi.dbg && threadIdx.x == R::kEpiBase * 32) {
      epi.dbg[blockIdx.x * 8 + 4] = clock64() - t_begin;
      epi.dbg[blockIdx.x * 8 + 5] = w_wait;
      epi.dbg[blockIdx.x * 8 + 6] = w_phase2;
    }
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(I8_TMEM_COLS) : "memory");
  }
}

// ---- activations -> digit planes, two tile-parallel kernels (grid: 128-atom tiles x 64-feature chunks).
// Thread == atom, so the feature-major reads are coalesced.
//   k_i8_rowmax:   scaled row maxima, combined across chunks with an integer atomicMax on the float bits
//   k_i8_quantize: fixed-point conversion + balanced base-256 digits; the 128 x 64-byte tile of each plane is
//                  staged in shared memory and written out as full 64-byte row segments.
__global__ void __launch_bounds__(128) k_i8_rowmax(const float* __restrict__ xt, int64_t ld, int64_t n_valid_rows, int K,
                                                   const float* __restrict__ col_scale, int32_t* __restrict__ rowmax_bits) {
  const int64_t m = int64_t(blockIdx.x) * 128 + threadIdx.x;
  if (m >= n_valid_rows) return;
  const int k0 = blockIdx.y * 64, k1 = min(K, k0 + 64);
  const float* x = xt + m;
  float mx = 0.0f;
#pragma unroll 8
  for (int k = k0; k < k1; ++k) mx = fmaxf(mx, fabsf(x[int64_t(k) * ld] * __ldg(col_scale + k)));
  if (!(mx < 3.0e38f)) mx = 3.4e38f;   // NaN / inf rows are flagged by a huge maximum and quantise to zero
  atomicMax(rowmax_bits + m, __float_as_int(mx));
}

template <int P>
__global__ void __launch_bounds__(128, 7) k_i8_quantize(const float* __restrict__ xt, int64_t ld, int64_t n_valid_rows, int K, int kp,
                                                        const float* __restrict__ col_scale,
                                                        const int32_t* __restrict__ rowmax_bits,
                                                        int8_t* __restrict__ planes, int32_t* __restrict__ row_exp) {
  __shared__ __align__(16) uint8_t tile[P][128 * 64];
  const int tid = threadIdx.x;
  const int64_t m0 = int64_t(blockIdx.x) * 128, m = m0 + tid;
  const int kc = blockIdx.y * 64;
  int e;
  float sc, sc2;
  i8_row_scale(m < n_valid_rows ? __int_as_float(rowmax_bits[m]) : 0.0f, P, e, sc, sc2);   // padding rows: zero digits
  if (blockIdx.y == 0) row_exp[m] = e;
  const float* x = xt + m;
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {
    const int k0 = kc + g * 16;
    float v[16];
    if (k0 + 16 <= K) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = x[int64_t(k0 + j) * ld];
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (k0 + j < K) ? x[int64_t(k0 + j) * ld] : 0.0f;
    }
    uint32_t w[P][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t ew[4], tb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + 4 * q + j;
        const float c = (k < K) ? __ldg(col_scale + k) : 0.0f;
        ew[j] = i8_digit_word<P>(sc != 0.0f ? ((v[4 * q + j] * c) * sc) * sc2 : 0.0f);
      }
      i8_transpose4(ew, tb);
#pragma unroll
      for (int d = 0; d < P; ++d) w[d][q] = tb[P - 1 - d];
    }
    const int piece = g ^ ((tid >> 1) & 3);   // 16-byte pieces of a row are xor-swizzled: conflict-free both ways
#pragma unroll
    for (int d = 0; d < P; ++d)
      *reinterpret_cast<uint4*>(&tile[d][tid * 64 + piece * 16]) = make_uint4(w[d][0], w[d][1], w[d][2], w[d][3]);
  }
  __syncthreads();
  const int64_t plane_stride = ld * int64_t(kp);
#pragma unroll
  for (int d = 0; d < P; ++d) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = tid + 128 * j, r = idx >> 2, p = idx & 3;
      const uint4 val = *reinterpret_cast<const uint4*>(&tile[d][r * 64 + (p ^ ((r >> 1) & 3)) * 16]);
      *reinterpret_cast<uint4*>(planes + d * plane_stride + (m0 + r) * int64_t(kp) + kc + p * 16) = val;
    }
  }
}

// planes [n][rows][kp] int8 as a 3-d tensor {kp, rows, n}; box {64 bytes, box_rows, 1 plane}, SWIZZLE_64B
int make_plane_tmap(CUtensorMap* tm, const I8Planes& p, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return APAX_ERR_CUDA;
  const cuuint64_t dims[3] = {cuuint64_t(p.kp), cuuint64_t(p.rows), cuuint64_t(p.n_planes)};
  const cuuint64_t strides[2] = {cuuint64_t(p.kp), cuuint64_t(p.plane_stride)};
  const cuuint32_t box[3] = {I8_KC, cuuint32_t(box_rows), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(p.base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? APAX_OK : APAX_ERR_INVALID_ARG;
}

struct I8Launch {
  CUtensorMap ta, tb;
  int num_kb, a_planes, a_res, num_m_tiles, num_n_sub, b_stages, grid;
  uint32_t smem;
};

template <class Kern>
int prepare_i8(Kern kern) {
  cudaFuncAttributes fa;
  APAX_CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
  APAX_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(I8_SMEM_LIMIT - fa.sharedSizeBytes)));
  return APAX_OK;
}

// one named launch per epilogue so that launch lists and the per-kernel profile tell the four GEMMs apart
#define I8_LAUNCH_W(name, MODE, NOUT, WIDE)                                                                          \
  do {                                                                                                               \
    auto name = k_i8_gemm<MODE, NOUT, WIDE>;                                                                         \
    static bool prepared = false;                                                                                    \
    if (!prepared) {                                                                                                 \
      APAX_TRY(prepare_i8(name));                                                                                    \
      prepared = true;                                                                                               \
    }                                                                                                                \
    APAX_LAUNCH(name, dim3((unsigned)L.grid), dim3(I8Roles<WIDE>::kThreads), L.smem, stream, L.ta, L.tb, L.num_kb,   \
                L.a_planes, L.a_res, L.num_m_tiles, L.num_n_sub, L.b_stages, epi, err_flag);                         \
    return APAX_OK;                                                                                                  \
  } while (0)
// the eight-warp kernels, or (apax_i8_set_debug flag 16) the wide-epilogue variant
#define I8_LAUNCH(name, MODE, NOUT)                 \
  do {                                              \
    if (wide) I8_LAUNCH_W(name, MODE, NOUT, 1);     \
    I8_LAUNCH_W(name, MODE, NOUT, 0);               \
  } while (0)

}  // namespace

int i8_quantize_rows(cudaStream_t stream, const float* xt, int64_t ld, int64_t n_valid_rows, int K, int kp, int n_planes,
                     const float* col_scale, int8_t* planes, int32_t* row_exp, int32_t* rowmax_bits, bool have_rowmax,
                     int64_t n_rows) {
  if (n_rows <= 0) n_rows = ld;   // rows to emit (whole 128-row tiles; rows >= n_valid_rows get zero digits); default: all
  if (ld % 128 != 0 || n_rows % 128 != 0 || n_rows > ld || kp % 64 != 0 || K > kp || (n_planes != 3 && n_planes != 4))
    return APAX_ERR_INVALID_ARG;
  const dim3 grid((unsigned)(n_rows / 128), (unsigned)(kp / 64));
  if (!have_rowmax) {   // otherwise the producer of xt already left the maxima of the first n_valid_rows rows
    APAX_CUDA_TRY(cudaMemsetAsync(rowmax_bits, 0, sizeof(int32_t) * ld, stream));
    APAX_LAUNCH(k_i8_rowmax, grid, dim3(128), 0, stream, xt, ld, n_valid_rows, K, col_scale, rowmax_bits);
  }
  if (n_planes == 3) {
    APAX_LAUNCH(k_i8_quantize<3>, grid, dim3(128), 0, stream, xt, ld, n_valid_rows, K, kp, col_scale, rowmax_bits, planes, row_exp);
  } else {
    APAX_LAUNCH(k_i8_quantize<4>, grid, dim3(128), 0, stream, xt, ld, n_valid_rows, K, kp, col_scale, rowmax_bits, planes, row_exp);
  }
  return APAX_OK;
}

void i8_weight_scales(const float* W, int K, int NC, int ldw, int kp, std::vector<float>& col_scale, bool merge) {
  if (!merge) col_scale.assign(size_t(kp), 0.0f);
  for (int k = 0; k < K; ++k) {
    double mx = 0.0;
    for (int n = 0; n < NC; ++n) mx = std::fmax(mx, std::fabs(double(W[size_t(k) * ldw + n])));
    if (!(mx > 0.0) || !std::isfinite(mx)) continue;
    int ex;
    std::frexp(mx, &ex);                                   // mx / 2^(ex-6) in [32, 64)
    col_scale[k] = std::fmax(col_scale[k], float(std::ldexp(1.0, ex - 6)));   // power of two: exact in fp32
  }
}

void i8_quantize_weights(const float* W, int K, int NC, int ldw, int kp, int ncp, std::vector<int8_t>& planes,
                         std::vector<float>& col_scale, bool given_scale) {
  planes.assign(size_t(I8_DIGITS) * ncp * kp, 0);
  if (!given_scale) i8_weight_scales(W, K, NC, ldw, kp, col_scale, false);
  const size_t plane_stride = size_t(ncp) * kp;
  float min_scale = 0.0f;
  for (int k = 0; k < K; ++k)
    if (col_scale[k] > 0.0f && (min_scale == 0.0f || col_scale[k] < min_scale)) min_scale = col_scale[k];
  if (min_scale == 0.0f) min_scale = 1.0f;
  for (int k = 0; k < K; ++k) {
    if (!(col_scale[k] > 0.0f)) {
      // all-zero weight row (a feature dropped by n_contr < 8): its digits stay zero; the scale is the smallest real one
      // so that the ignored activation cannot dominate the row maximum that fixes the atoms' exponents
      col_scale[k] = min_scale * 0.00390625f;
      continue;
    }
    const double c = double(col_scale[k]);
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
  for (int k = K; k < kp; ++k) col_scale[k] = 1.0f;
}

// measurement switches of tools/trace_i8.py (apax_i8_set_debug); both 0 in normal operation
static std::atomic<int> g_i8_dbg_flags{0}, g_i8_trace{0};

int i8_gemm(cudaStream_t stream, const I8Planes& act, const I8Planes& wgt, const I8Epilogue& epi_in, int32_t* err_flag) {
  const I8Epilogue& epi0 = epi_in;
  (void)epi0;
  if (act.kp != wgt.kp || act.kp % I8_KC != 0 || act.kp / I8_KC > I8_MAX_KB || act.rows % I8_BM != 0 ||
      wgt.rows % I8_BN != 0 || wgt.n_planes != I8_DIGITS || (act.n_planes != 3 && act.n_planes != 4))
    return APAX_ERR_INVALID_ARG;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    APAX_CUDA_TRY(cudaGetDevice(&dev));
    APAX_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  I8Epilogue epi = epi_in;
  if (g_i8_dbg_flags.load(std::memory_order_relaxed)) epi.dbg_flags = g_i8_dbg_flags.load(std::memory_order_relaxed);
  const bool trace = g_i8_trace.load(std::memory_order_relaxed) != 0;   // measurement only (apax_i8_set_debug): per-launch cycle counters to stderr
  static long long* trace_buf = nullptr;
  if (trace) {
    if (!trace_buf) APAX_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&trace_buf), 256 * 8 * sizeof(long long)));
    APAX_CUDA_TRY(cudaMemsetAsync(trace_buf, 0, 256 * 8 * sizeof(long long), stream));
    epi.dbg = trace_buf;
  }
  struct TraceDump {
    bool on; cudaStream_t st; long long* buf; int mode; int* grid;
    ~TraceDump() {
      if (!on) return;
      std::vector<long long> h(256 * 8);
      cudaStreamSynchronize(st);
      cudaMemcpy(h.data(), buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      double m[8] = {0};
      for (int c = 0; c < *grid; ++c)
        for (int i = 0; i < 8; ++i) m[i] += double(h[c * 8 + i]) / *grid;
      fprintf(stderr, "[i8 trace] mode %d grid %d: mma warp %.0f clk (waits: accumulator %.0f, weight stages %.0f, atom planes %.0f) | "
                      "epilogue total %.0f wait %.0f phase2 %.0f\n", mode, *grid, m[0], m[1], m[2], m[3], m[4], m[5], m[6]);
    }
  };
  int trace_grid = 0;
  TraceDump trace_dump{trace, stream, trace_buf, epi.mode, &trace_grid};
  I8Launch L;
  APAX_TRY(make_plane_tmap(&L.ta, act, I8_BM));
  APAX_TRY(make_plane_tmap(&L.tb, wgt, I8_BN));
  L.num_kb = int(act.kp / I8_KC);
  L.a_planes = act.n_planes;
  L.num_m_tiles = int(act.rows / I8_BM);
  L.num_n_sub = int(wgt.rows / I8_BN);
  L.grid = L.num_m_tiles < n_sm ? L.num_m_tiles : n_sm;   // one persistent CTA per SM (it owns the SM's whole TMEM)
  trace_grid = L.grid;
  // as many activation planes as fit stay resident next to a ring of at least three weight stages; the least
  // significant planes (used by the fewest MMAs) travel with the stages otherwise
  // the eight-warp epilogue ships; apax_i8_set_debug flag 16 selects the sixteen-warp kernels (measurement: slower, see I8Roles)
  const bool wide = (g_i8_dbg_flags.load(std::memory_order_relaxed) & 16) != 0;
  // static shared memory + alignment: the many-head epilogue of the wide kernel keeps three partial head sums per row
  const uint32_t static_reserve = (wide && epi.mode == I8_EPI_ACT_HEAD && epi.n_out > 1) ? 36864 : 10240;
  const uint32_t budget = I8_SMEM_LIMIT - static_reserve;
  L.a_res = L.a_planes;
  uint32_t a_bytes, stage_bytes;
  for (;; --L.a_res) {
    a_bytes = uint32_t(L.num_kb * L.a_res) * I8_A_TILE;
    stage_bytes = I8_B_STAGE + uint32_t(L.a_planes - L.a_res) * I8_A_TILE;
    if (a_bytes + 3 * stage_bytes <= budget) break;
    if (L.a_res == 0) return APAX_ERR_INVALID_ARG;
  }
  // measurement (apax_i8_set_debug flags 32 / 64): one / two fewer resident planes than fit, i.e. a deeper weight ring
  // (more bytes in flight) paid for with re-streamed activation planes
  {
    const int dbgf = g_i8_dbg_flags.load(std::memory_order_relaxed);
    int drop = ((dbgf & 32) ? 1 : 0) + ((dbgf & 64) ? 2 : 0);
    while (drop-- > 0 && L.a_res > 1 && L.a_res * L.num_kb * int(I8_A_TILE) > 100 * 1024) {
      --L.a_res;
      a_bytes = uint32_t(L.num_kb * L.a_res) * I8_A_TILE;
      stage_bytes = I8_B_STAGE + uint32_t(L.a_planes - L.a_res) * I8_A_TILE;
    }
  }
  L.b_stages = int((budget - a_bytes) / stage_bytes);
  if (L.b_stages > I8_MAX_B_STAGES) L.b_stages = I8_MAX_B_STAGES;
  L.smem = a_bytes + uint32_t(L.b_stages) * stage_bytes + 1024;
  switch (epi.mode) {
    case I8_EPI_STORE: I8_LAUNCH(k_i8_gemm_store, I8_EPI_STORE, 1);
    case I8_EPI_ACT_Q: I8_LAUNCH(k_i8_gemm_act_q, I8_EPI_ACT_Q, 1);
    case I8_EPI_MULD_Q: I8_LAUNCH(k_i8_gemm_muld_q, I8_EPI_MULD_Q, 1);
    case I8_EPI_ACT_HEAD:
      if (epi.n_out == 1) I8_LAUNCH(k_i8_gemm_act_head, I8_EPI_ACT_HEAD, 1);
      if (epi.n_out <= I8_MAX_HEADS) I8_LAUNCH(k_i8_gemm_act_heads, I8_EPI_ACT_HEAD, I8_MAX_HEADS);
      return APAX_ERR_INVALID_ARG;
    default: return APAX_ERR_INVALID_ARG;
  }
}

}  // namespace apax

// Test / measurement entry: out[nc][ld] = act^T . wgt through the integer tensor-core path (plain store
// epilogue).  act: device fp32 [k][ld]; wgt: device fp32 [k][nc].  Allocates its own scratch (test only).
extern "C" int apax_i8_set_debug(int32_t dbg_flags, int32_t trace) {
  apax::g_i8_dbg_flags.store(dbg_flags);
  apax::g_i8_trace.store(trace);
  return APAX_OK;
}

extern "C" int apax_i8_gemm_test(apax_stream_t stream_, const float* act, int64_t k, int64_t ld, int64_t n_atoms,
                                 const float* wgt, int64_t nc, float* out, int32_t* err_flag, int64_t* dbg,
                                 int32_t act_planes) {
  if (!act || !wgt || !out || !err_flag || k <= 0 || nc <= 0 || ld % 128 != 0) return APAX_ERR_INVALID_ARG;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int kp = int((k + 63) / 64 * 64), ncp = int((nc + 63) / 64 * 64);
  std::vector<float> hw(size_t(k) * nc);
  APAX_CUDA_TRY(cudaStreamSynchronize(stream));
  APAX_CUDA_TRY(cudaMemcpy(hw.data(), wgt, hw.size() * sizeof(float), cudaMemcpyDeviceToHost));
  std::vector<int8_t> wp;
  std::vector<float> cs;
  apax::i8_quantize_weights(hw.data(), int(k), int(nc), int(nc), kp, ncp, wp, cs);
  int8_t *d_wp = nullptr, *d_ap = nullptr;
  float* d_cs = nullptr;
  int32_t *d_e = nullptr, *d_mx = nullptr;
  int rc = APAX_OK;
  if (cudaMalloc(&d_wp, wp.size()) != cudaSuccess || cudaMalloc(&d_ap, size_t(4) * ld * kp) != cudaSuccess ||
      cudaMalloc(&d_cs, cs.size() * sizeof(float)) != cudaSuccess || cudaMalloc(&d_e, ld * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc(&d_mx, ld * sizeof(int32_t)) != cudaSuccess)
    rc = APAX_ERR_CUDA;
  if (rc == APAX_OK) {
    cudaMemcpy(d_wp, wp.data(), wp.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_cs, cs.data(), cs.size() * sizeof(float), cudaMemcpyHostToDevice);
    rc = apax::i8_quantize_rows(stream, act, ld, n_atoms, int(k), kp, act_planes, d_cs, d_ap, d_e, d_mx);
  }
  if (rc == APAX_OK) {
    apax::I8Planes a{d_ap, ld, kp, ld * int64_t(kp), act_planes}, w{d_wp, ncp, kp, int64_t(ncp) * kp, 4};
    apax::I8Epilogue e{};
    e.mode = apax::I8_EPI_STORE;
    e.ld = ld;
    e.n_valid = int(nc);
    e.row_exp = d_e;
    e.out = out;
    e.dbg = reinterpret_cast<long long*>(dbg);
    rc = apax::i8_gemm(stream, a, w, e, err_flag);
  }
  if (cudaStreamSynchronize(stream) != cudaSuccess && rc == APAX_OK) rc = APAX_ERR_CUDA;
  cudaFree(d_wp);
  cudaFree(d_ap);
  cudaFree(d_cs);
  cudaFree(d_e);
  cudaFree(d_mx);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// Tensor-pipe / tensor-memory probe (measurement only): one CTA per SM.  Warp 4 issues `iters` rounds of ten
// tcgen05.mma kind::i8 (pattern 0: one accumulator; 1: four accumulators; 2: the GEMM's digit-pair schedule;
// < 0: none) on whatever shared memory holds; warps 0-3 meanwhile drain all 512 TMEM columns `reads` times with
// tcgen05.ld.32x32b.x32.  out[cta][0] = MMA cycles, out[cta][1] = drain cycles of warp 0.
// ------------------------------------------------------------------------------------------------
namespace apax {
namespace {
using namespace tcptx;
template <int N>
__global__ void __launch_bounds__(160, 1) k_i8_mma_probe(int pattern, int iters, int reads, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ __align__(8) uint64_t side_bar;
  __shared__ __align__(8) uint64_t ready_bar;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const bool random_data = (pattern & 16) != 0;   // constant operands toggle few wires; random ones load the data paths
  if (pattern >= 0) pattern &= 15;
  for (uint32_t i = threadIdx.x; i < 96 * 1024 / 4; i += 160) {
    uint32_t v = 0x01010101u;
    if (random_data) {
      v = (i + blockIdx.x * 7919u) * 2654435761u;
      v ^= v >> 15; v *= 2246822519u; v ^= v >> 13;
    }
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + i * 4), "r"(v) : "memory");
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&done_bar), 1);
    mbar_init(smem_u32(&side_bar), 1);
    mbar_init(smem_u32(&ready_bar), 1);
    mbar_arrive(smem_u32(&ready_bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;
  if (warp == 4) {
    constexpr uint32_t idesc = make_idesc_i8(128, N);
    constexpr uint32_t B_TILE = N * 64;
    const uint64_t a0 = make_desc_k_sw64(base), b0 = make_desc_k_sw64(base + 4 * I8_A_TILE);
    const long long t0 = clock64();
    if (pattern >= 0) {
      for (int it = 0; it < iters; ++it) {
        if (elect_one()) {
          if (pattern == 0) {
#pragma unroll
            for (int i = 0; i < 10; ++i) umma_i8(tmem_d, a0, b0, idesc, 1u);
          } else if (pattern == 1) {
#pragma unroll
            for (int i = 0; i < 10; ++i) umma_i8(tmem_d + (N > 128 ? (i & 1) * 256 : (i & 3) * N), a0, b0, idesc, 1u);
          } else if (pattern == 2) {
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
              for (int sa = 0; sa <= o; ++sa)
                umma_i8(tmem_d + (N > 128 ? (o & 1) * 256 : o * N), a0 + uint64_t((sa * I8_A_TILE + (it & 1) * 32) >> 4),
                        b0 + uint64_t(((o - sa) * B_TILE + (it & 1) * 32) >> 4), idesc, 1u);
          } else {   // the GEMM's merged schedule: digit s of the activations against 4 - s weight planes (ten products)
#pragma unroll
            for (int sa = 0; sa < 4; ++sa)
              umma_i8(tmem_d + sa * 64, a0 + uint64_t((sa * I8_A_TILE + (it & 1) * 32) >> 4), b0 + uint64_t(((it & 1) * 32) >> 4),
                      make_idesc_i8(128, (4 - sa) * 64), 1u);
            if (pattern >= 4 && (it & 1)) umma_commit(smem_u32(&side_bar));   // one commit per "k-block" of 8 MMAs
          }
        }
        __syncwarp();
        if (pattern >= 5 && (it & 1)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (pattern >= 6 && (it & 1)) {
          int32_t dummy = 0;
          mbar_wait(smem_u32(&ready_bar), 0, &dummy);   // long since complete
        }
        if (pattern >= 7 && (it & 1)) {
          int32_t dummy = 0;
          mbar_wait(smem_u32(&side_bar), (uint32_t(it >> 1) & 1), &dummy);   // wait for this k-block's own commit
        }
      }
      if (elect_one()) umma_commit(smem_u32(&done_bar));
      __syncwarp();
      int32_t dummy = 0;
      mbar_wait_warp(smem_u32(&done_bar), 0, &dummy);
    }
    if (threadIdx.x == 128) out[blockIdx.x * 2] = clock64() - t0;
  } else if (reads > 0) {
    const uint32_t taddr = tmem_d + (uint32_t(warp * 32) << 16);
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int it = 0; it < reads; ++it) {
#pragma unroll 1
      for (int c = 0; c < 512; c += 128) {
        uint32_t r0[32], r1[32], r2[32], r3[32];
        tmem_ld32_u32(taddr + c, r0);
        tmem_ld32_u32(taddr + c + 32, r1);
        tmem_ld32_u32(taddr + c + 64, r2);
        tmem_ld32_u32(taddr + c + 96, r3);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) sink ^= r0[i] ^ r1[i] ^ r2[i] ^ r3[i];
      }
    }
    if (threadIdx.x == 0) out[blockIdx.x * 2 + 1] = clock64() - t0;
    if (sink == 0x12345678u) out[blockIdx.x * 2 + 1] = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
}
}  // namespace
}  // namespace apax

extern "C" int apax_i8_mma_probe(apax_stream_t stream_, int32_t n, int32_t pattern, int32_t iters, int32_t reads,
                                 int64_t* out_cycles) {
  using namespace apax;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int smem = 100 * 1024;
  long long* out = reinterpret_cast<long long*>(out_cycles);
  if (n == 64) {
    APAX_CUDA_TRY(cudaFuncSetAttribute(k_i8_mma_probe<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_i8_mma_probe<64><<<148, 160, smem, stream>>>(pattern, iters, reads, out);
  } else if (n == 128) {
    APAX_CUDA_TRY(cudaFuncSetAttribute(k_i8_mma_probe<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_i8_mma_probe<128><<<148, 160, smem, stream>>>(pattern, iters, reads, out);
  } else if (n == 256) {
    APAX_CUDA_TRY(cudaFuncSetAttribute(k_i8_mma_probe<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_i8_mma_probe<256><<<148, 160, smem, stream>>>(pattern, iters, reads, out);
  } else {
    return APAX_ERR_INVALID_ARG;
  }
  APAX_CUDA_TRY(cudaGetLastError());
  return APAX_OK;
}
