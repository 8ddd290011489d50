// Internal layout of the GMNN hot-path workspace and parameter block.
#pragma once
#include "common.cuh"

namespace apax {

constexpr int NR = 5;        // radial channels
constexpr int NB = 7;        // Gaussian basis functions
constexpr int NPK = 100;     // packed moments per atom: 5 x (1+3+6+10)
constexpr int NF = 360;      // descriptor features
constexpr int NFP = 384;     // feature rows padded to a multiple of the GEMM tile
constexpr int NH = 256;      // hidden width
constexpr int EMB_STRIDE = 36;  // 35 coefficients padded to 9 float4
constexpr int MAXNB_GEN = 16;   // generic radial bases (Bessel, exponential spacing, n_basis != 7): up to 16 functions
constexpr int EMB_STRIDE_GEN = NR * MAXNB_GEN;   // [5][16] coefficients per species pair, zero padded
constexpr int MAX_SP_SMEM = 8;  // compact species tables up to 8x8 are staged in shared memory
constexpr int MAX_OUT = 16;

// pairwise empirical corrections (apax/layers/empirical.py:25-126), parameters already in their positive form
struct CorrConst {
  int zbl_on;
  float zbl_rmax, zbl_scale;      // rep_scale after softplus
  float zbl_c[4], zbl_e[4];       // coefficients after softplus, exponents
  int exp_on;
  float exp_rmax;
  const float* exp_r;             // [S] |rep_scale|   (device)
  const float* exp_a;             // [S] |rep_prefactor|
};

struct apax_gm_params_impl {
  apax_gm_config cfg;
  CorrConst corr;                 // all zero: no corrections
  int readout_mode;               // APAX_READOUT_* (apax_gm_params_set_option)
  int pair_block;                 // block shape of the pair kernels, 0..2 (measurement switch)
  int generic_basis;              // radial basis other than the 7 linearly spaced Gaussians: pair kernels with GEN = 1
  int contract_split;             // central atoms up to which the contraction kernels run one atom on four warps (0 = never)
  int mean_forces;                // n_out > 1: forces[1][N][3] of the MEAN energy in one backward pass instead of per head
  // device copies
  float* emb;    // [S][S][5][7], pre-multiplied by 1/sqrt(n_basis) when use_embed_norm
  float* w0;     // [360][256]
  float* w0t;    // [256][384]  (transposed, zero padded) for the backward GEMM
  float* b0;     // [256]
  float* w1;     // [256][256]
  float* w1t;    // [256][256]
  float* w1t_head;  // [n_out + 1][256][256]  W1^T with the head's output weights folded in (backward GEMM); last: the mean head
  float* b1;     // [256]
  float* w2;     // [256][n_out]
  float* b2;     // [n_out]
  double* scale; // [S]
  double* shift; // [S]
  // exact tf32 (hi, lo) pairs of the same weights for the tensor-core readout (tc_gemm.cuh)
  float *w0_hi, *w0_lo;    // [360][256]
  float *w1_hi, *w1_lo;    // [256][256]
  float *w1t_hi, *w1t_lo;  // [256][256]
  float *w0t_hi, *w0t_lo;  // [256][384]
  int32_t* tc_err;         // [1] set by the tensor-core pipeline on a barrier time-out
  // integer tensor-core readout (i8_gemm.cuh): weight digit planes [4][outputs][k] and per-feature scales
  int8_t* q_w0;     // [4][256][384]          layer 0
  int8_t* q_w1;     // [4][256][256]          layer 1
  int8_t* q_w1th;   // [n_out + 1][4][256][256]   backward through layer 1, head weights folded in (last: the mean head)
  int8_t* q_w0t;    // [4][384][256]          backward through layer 0
  float* cs_g;      // [384] scales of the descriptor features (layer 0 reduction)
  float* cs_h1;     // [256] scales of swish(y1)          (layer 1 reduction)
  float* cs_d2;     // [256] scales of swish'(y2)         (backward layer 1 reduction), shared by all heads
  float* cs_x1;     // [256] scales of dE/dy1             (backward layer 0 reduction)
  int d2_exp;       // exponent shared by every row of the swish'(y2) planes (|swish'| <= 1.85)
  double* eye;             // [3][3] identity (displacement box of Cartesian batched evaluation)
};

// Everything the compute kernels need, carved out of the caller's workspace.  All feature-major
// ("transposed") matrices have leading dimension ld = round_up(n_atoms, 128) along the atom axis.
struct GmWorkspace {
  int64_t n_atoms, n_pairs, ld;
  int64_t n_central;  // atoms [0, n_central) are evaluated as centrals; the rest only act as neighbours
  const int32_t* seg_ptr;  // optional [n_seg+1]: atom ranges of independent structures (batched evaluation)
  int64_t n_seg;           // 0: one structure, energy[n_out]; else energy[n_seg][n_out]
  int32_t n_out;
  // list (CSR by central atom) + transpose
  int32_t* rowptr;   // [N+1]
  int32_t* trowptr;  // [N+1]
  int32_t* nbr;      // [P] neighbour index of CSR slot
  int32_t* src;      // [P] original COO position of CSR slot
  int32_t* tpos;     // [P] for atom a: CSR slots in which a is the neighbour, ascending
  int32_t* tmp_a;    // [P]
  int32_t* tmp_b;    // [P]
  int32_t* cnt;      // [N+1]
  int32_t* tcnt;     // [N+1]
  int32_t* scan_scratch;
  // species compaction
  int32_t* sp_present;  // [128]
  int32_t* sp_map;      // [128] Z -> compact id
  int32_t* n_sp;        // [1]
  int32_t* sp_z;        // [128] compact id -> atomic number
  float* embc;          // [n_sp*n_sp][36]
  // per atom: (x, y, z, compact species id as int bits) -- one aligned 32-byte gather per neighbour
  double4* P;
  // per pair
  float4* G;  // (dx,dy,dz, pair-table index as int bits)
  float4* D;  // dE/d(dr_vec) per pair
  // per atom, feature-major
  float* MT;   // [100][ld] packed moments
  float* AT;   // [100][ld] packed moment adjoints
  float* GT;   // [384][ld] features g, later overwritten by dE/dg
  float* H1T;  // [256][ld] act(y1), later dE/dy1
  float* D1T;  // [256][ld] act'(y1)
  float* H2T;  // [256][ld]
  float* D2T;  // [256][ld]
  // residual (lo) parts of GT / H1T / H2T when the readout runs on the tensor cores (3xTF32)
  float* GLT;   // [384][ld]
  float* H1LT;  // [256][ld]
  float* H2LT;  // [256][ld]
  float* ebar;     // [ld]   dE/dh per atom (scale*mask) in fp32
  double* Eatom;   // [n_out][ld]
  double* partial; // [n_out][256]
  double* Frow;    // [N][3] sum of dbar over the rows of the central atom
  // integer tensor-core readout: the contraction kernel also leaves each atom's scaled feature maximum (set per call)
  const float* q_cs;   // [384] per-feature scales of layer 0, or null
  int32_t* q_rowmax;   // [ld] float bits of max_k |g[k] c[k]|
  int8_t* q_planes;    // fused contraction + quantiser: layer-0 digit planes [4][ld][384] (null: separate quantiser)
  int32_t* q_row_exp;  // ... and the row exponents [ld]
};

size_t carve_gm_workspace(void* base, int64_t n_atoms, int64_t n_pairs, int32_t n_out, int32_t n_species,
                          GmWorkspace* ws);

// exported to ctx.cu
int gm_prepare_from_csr(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w, const int32_t* Z,
                        const int32_t* rowptr_in, const int32_t* nbr_in, bool symmetric_sorted = false,
                        int64_t n_central = -1 /* symmetric lists: rows >= n_central are not evaluated (ghosts) */);
struct PairGeom;
int gm_compute_device(cudaStream_t stream, const apax_gm_params_impl* p, const GmWorkspace& w, const double* R,
                      const int32_t* Z, const double* box, const double* offsets, int mode, double* energy,
                      double* forces, int64_t n_central = -1);

}  // namespace apax
