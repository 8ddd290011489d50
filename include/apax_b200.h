/* apax_b200 -- C ABI of the B200-native GMNN energy+forces path.
 *
 * Drop-in boundary for ONE hot path of apax (https://github.com/apax-hub/apax): periodic neighbour
 * list -> GaussianMomentDescriptor -> per-element MLP readout -> scale/shift -> fp64 energy -> forces.
 * apax itself has no native layer; every function below replaces a piece of JAX/Flax code that the
 * reference traces into XLA.  The citation after each prototype names that piece (path:line relative to
 * the apax repository).  INTEGRATION.md shows the XLA-FFI / ctypes stubs that bind these symbols.
 *
 * Conventions
 *   - plain C, no C++/torch types; all pointers are DEVICE pointers unless the name ends in `_host`.
 *   - every device-side function is asynchronous on `stream`, never allocates, never synchronises and
 *     never writes beyond a caller-given capacity; it returns 0 or a negative apax_status.
 *   - conditions only the device can see (list overflow) are reported through device-side flags,
 *     like jax_md's PartitionErrorCode (apax/utils/jax_md_reduced/partition.py:168-191).
 *   - neighbour lists cross the boundary in apax's own COO form: int32 idx[2][P], idx[0] = neighbour,
 *     idx[1] = central atom, dr = R[idx[1]] - R[idx[0]] (+ offsets)  (apax/layers/distances.py:48-67).
 */
#ifndef APAX_B200_H
#define APAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APAX_B200_ABI_VERSION 3 /* 3: options, descriptor VJP, prepare_minimage, ctx stress, 512-byte halo header, slab-MD entry points */

typedef struct CUstream_st* apax_stream_t; /* == cudaStream_t */

typedef enum {
  APAX_OK = 0,
  APAX_ERR_INVALID_ARG = -1,
  APAX_ERR_UNSUPPORTED = -2, /* statics outside what the kernels are specialised for */
  APAX_ERR_WORKSPACE = -3,   /* caller workspace too small */
  APAX_ERR_CUDA = -4,        /* a CUDA runtime call failed (see apax_last_cuda_error) */
  APAX_ERR_OVERFLOW = -5     /* host-visible capacity exceeded */
} apax_status;

/* Displacement branch of make_distance_fn (apax/layers/distances.py:32-67). */
typedef enum {
  APAX_DISP_GAS = 0,      /* dr = R[i1]-R[i0]                               (:59-62) */
  APAX_DISP_OFFSETS = 1,  /* dr = box.(R[i1]-R[i0]) + offsets               (:12-14,:65-66) */
  APAX_DISP_MINIMAGE = 2, /* dr = box.wrap(R[i1]-R[i0]) + offsets, R fractional
                             (apax/utils/jax_md_reduced/space.py:260-282) */
  APAX_DISP_DRVEC = 3     /* apax_gm_descriptor only: `R` is dr_vec f64 [n_pairs][3] itself, the argument
                             of GaussianMomentDescriptor.__call__ (gaussian_moment_descriptor.py:31) */
} apax_disp_mode;

/* Hot-path statics (apax/config/model_config.py:12-34,147-227; Flax defaults in SURVEY.md s8). */
typedef struct {
  int32_t n_species;      /* 119 */
  int32_t n_radial;       /* 5   (kernels specialised: must be 5) */
  int32_t n_basis;        /* 7   (must be 7) */
  int32_t n_contr;        /* 8   (1..8: the first 5/20/35/50/85/160/235/360 features, gaussian_moment_descriptor.py:101) */
  int32_t n_hidden1;      /* 256 (must be 256) */
  int32_t n_hidden2;      /* 256 (must be 256) */
  int32_t n_out;          /* 1, or n_shallow_ensemble (<= 16) */
  int32_t use_ntk;        /* NTKLinear scaling (apax/layers/ntk_linear.py:42-45) */
  int32_t use_embed_norm; /* 1/sqrt(n_basis) (apax/layers/descriptor/basis_functions.py:145-146) */
  int32_t mask_atoms;     /* mask_by_atom on E_i (apax/nn/models.py:111-112) */
  float r_min;            /* 0.5 */
  float r_max;            /* 6.0 */
  int32_t basis_kind;     /* radial basis (apax/layers/descriptor/basis_functions.py): 0 GaussianBasis spacing="linear" (:22-28),
                             1 GaussianBasis spacing="exponential" (:29-37), 2 BesselBasis (:59-79).  The apax_gm_* fast path is
                             specialised for kind 0 with n_basis = 7; the other kinds (n_basis <= 16) run on the table-driven
                             kernels behind apax_train_* (energy+forces: apax_train_energy_forces). */
} apax_gm_config;

/* ------------------------------------------------------------------------------------------------
 * Parameters.  Replaces: the Flax parameter tree consumed by model.apply (layout in SURVEY.md s5;
 * apax/layers/descriptor/basis_functions.py:110-120, apax/layers/ntk_linear.py:37-38,
 * apax/layers/scaling.py:33-38).  Host arrays in, opaque device-resident handle out (weights are
 * uploaded once, transposed/padded copies for the backward GEMMs are made here).
 *   emb_host   f32 [n_species][n_species][n_radial][n_basis]
 *   w0_host    f32 [n_feat][256] (n_feat = 360 for n_contr = 8), b0_host f32 [256]; w1_host f32 [256][256], b1_host f32 [256]
 *   w2_host    f32 [256][n_out], b2_host f32 [n_out]
 *   scale_host, shift_host  f64 [n_species]
 * ---------------------------------------------------------------------------------------------- */
typedef struct apax_gm_params apax_gm_params;
int apax_gm_params_create(const apax_gm_config* cfg, const float* emb_host, const float* w0_host,
                          const float* b0_host, const float* w1_host, const float* b1_host,
                          const float* w2_host, const float* b2_host, const double* scale_host,
                          const double* shift_host, apax_gm_params** out);
void apax_gm_params_destroy(apax_gm_params* p);
/* Synchronising health query: 0, or APAX_ERR_CUDA if a tensor-core pipeline barrier ever timed out
 * while these parameters were in use (results of that call are then invalid). */
int apax_gm_params_status(const apax_gm_params* p);

/* Implementation switches of a parameter handle (never read from the environment; the numerics path of a process is what
 * its host code set here).  APAX_OPT_READOUT: which kernels run the per-element MLP (apax/layers/readout.py:22-50) --
 * APAX_READOUT_I8 (default) exact integer tensor-core GEMMs, APAX_READOUT_SIMT fp32 CUDA-core GEMMs (cross-check),
 * APAX_READOUT_TC tcgen05 3xTF32 (biased accumulation, NOT parity-safe; measurement only).  APAX_OPT_PAIR_BLOCK: block shape
 * of the pair kernels, 0..2 (results are bitwise identical; measurement only).  APAX_OPT_MEAN_FORCES (shallow ensembles,
 * n_out > 1): 1 = every later evaluation returns forces f64 [1][n_atoms][3] = -d(mean over heads of E)/dR from ONE backward
 * pass -- ShallowEnsembleModel with force_variance = False (apax/nn/models.py:265-267) and the mean-energy gradient of the MD
 * energy function (apax/md/simulate.py:55-58) -- instead of the per-head forces of jacrev.  APAX_OPT_CONTRACT_SPLIT: systems of
 * at most `value` central atoms run the contraction kernels with one atom on four warps (small systems are bound by the
 * latency of one warp walking the whole contraction program); 0 = never; default 8192 (one wave of 32-atom groups).  Results differ from the one-thread
 * form only in the summation order of the moment adjoints (fp32 rounding).  Not to be called while an evaluation with
 * these parameters is being enqueued from another thread. */
typedef enum { APAX_OPT_READOUT = 0, APAX_OPT_PAIR_BLOCK = 1, APAX_OPT_MEAN_FORCES = 2, APAX_OPT_CONTRACT_SPLIT = 3 } apax_gm_option;
typedef enum { APAX_READOUT_SIMT = 0, APAX_READOUT_TC = 1, APAX_READOUT_I8 = 2 } apax_readout_kind;
int apax_gm_params_set_option(apax_gm_params* p, int32_t option, int32_t value);

/* Pairwise empirical corrections added to the model energy (EnergyModel.corrections, apax/nn/models.py:127-130), with
 * the parameters as they sit in the Flax tree (raw, i.e. before softplus / abs):
 *   ZBLRepulsion        (apax/layers/empirical.py:25-86):  coefficients f32 [4], rep_scale, r_max
 *   ExponentialRepulsion (:89-126): rep_scale f32 [n_species], rep_prefactor f32 [n_species] (host arrays), r_max
 * They take part in every later apax_gm_* evaluation with these parameters (energy, forces, stress, all heads). */
typedef struct {
  int32_t zbl_on;
  float zbl_r_max;
  float zbl_coefficients[4];
  float zbl_rep_scale;
  int32_t exp_on;
  float exp_r_max;
  const float* exp_rep_scale_host;
  const float* exp_rep_prefactor_host;
} apax_gm_corrections;
int apax_gm_params_set_corrections(apax_gm_params* p, const apax_gm_corrections* corrections);

/* ------------------------------------------------------------------------------------------------
 * Energy + forces given a neighbour list.  Replaces EnergyDerivativeModel.apply
 * (apax/nn/models.py:146-171 = jax.value_and_grad of EnergyModel.__call__, :87-134).
 *   R       f64 [n_atoms][3]  (fractional for OFFSETS/MINIMAGE, Cartesian for GAS)
 *   Z       i32 [n_atoms]     (0 = padding atom)
 *   idx     i32 [2][n_pairs]  apax COO list; entries with idx[0]==idx[1] (the (0,0) / (N,N) padding)
 *                             contribute exactly zero (apax/layers/masking.py:10-16)
 *   box     f64 [3][3]        as passed to model.apply (ignored for GAS)
 *   offsets f64 [n_pairs][3]  Cartesian, or NULL for all-zero
 *   energy  f64 [n_out]       out (n_out == 1: the scalar energy)
 *   forces  f64 [n_out][n_atoms][3] out; for n_out == 1 this is F = -dE/dR; for a shallow ensemble
 *                             the per-head forces -jacrev(E_ens) (apax/nn/models.py:236)
 *   workspace: >= apax_gm_workspace_bytes(n_atoms, n_pairs, n_out) bytes, 256-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
size_t apax_gm_workspace_bytes(int64_t n_atoms, int64_t n_pairs, int32_t n_out);
int apax_gm_energy_forces(apax_stream_t stream, const apax_gm_params* params, const double* R,
                          const int32_t* Z, const int32_t* idx, const double* box,
                          const double* offsets, int64_t n_atoms, int64_t n_pairs, int32_t disp_mode,
                          double* energy, double* forces, void* workspace, size_t workspace_bytes);

/* The same computation split at the list-preparation boundary, so that a list that did not change
 * (MD between rebuilds, apax/md/simulate.py:313-354) is prepared once:
 *   apax_gm_prepare : COO -> CSR by central atom + transpose map + compact species table, in workspace
 *   apax_gm_compute : E, F from the prepared workspace (R may have changed; Z, idx may not).      */
int apax_gm_prepare(apax_stream_t stream, const apax_gm_params* params, const int32_t* Z,
                    const int32_t* idx, int64_t n_atoms, int64_t n_pairs, void* workspace,
                    size_t workspace_bytes);
int apax_gm_compute(apax_stream_t stream, const apax_gm_params* params, const double* R,
                    const int32_t* Z, const double* box, const double* offsets, int64_t n_atoms,
                    int64_t n_pairs, int32_t disp_mode, double* energy, double* forces,
                    void* workspace, size_t workspace_bytes);

/* Per-atom energies E_i of the last apax_gm_compute / apax_gm_energy_forces on this workspace: what EnergyModel.__call__
 * sums at apax/nn/models.py:109-122 (scale/shift and atom mask applied, empirical corrections included), i.e. the
 * reference's atomic energies before fp64_sum.  e_atoms f64 [n_out][n_atoms], device.  With apax_gm_compute_partial the
 * entries of ghost atoms are unspecified. */
int apax_gm_atomic_energies(apax_stream_t stream, const apax_gm_params* params, int64_t n_atoms, int64_t n_pairs,
                            double* e_atoms, void* workspace, size_t workspace_bytes);

/* Neighbour-list build AND list preparation in one asynchronous call, everything on the device: replaces
 * partition.neighbor_list(...).allocate / update (apax/utils/jax_md_reduced/partition.py:478-787, Sparse, min-image,
 * fractional coordinates) followed by the consumption of neighbor.idx in the model (apax/nn/models.py:100-107) for callers
 * whose positions live on the device.  The list goes straight from the cell-list builder's CSR into the prepared workspace
 * (no COO round trip); afterwards apax_gm_compute / apax_gm_compute_partial / apax_md_nvt_run are called with
 * n_pairs = capacity.  Atoms [0, n_central) are evaluated as centrals (n_central = n_atoms for a whole box; the owned atoms
 * of a slab otherwise, ghosts behind them).
 *   frac f64 [n_atoms][3], Z i32 [n_atoms], box f64 [3][3] (columns = lattice vectors), cutoff = r_max + dr_threshold
 *   n_found i32 [1] out: list entries found (over ALL n_atoms rows); overflow u8 [1] out: 1 if n_found > capacity -- the
 *   caller must check it before trusting results (PEC.NEIGHBOR_LIST_OVERFLOW, partition.py:734)
 *   gm_workspace >= apax_gm_workspace_bytes(n_atoms, capacity, n_out); nl_workspace >= apax_nl_workspace_bytes(n_atoms, capacity) */
int apax_gm_prepare_minimage(apax_stream_t stream, const apax_gm_params* params, const double* frac, const int32_t* Z,
                             const double* box, int64_t n_atoms, int64_t n_central, double cutoff, int64_t capacity,
                             int32_t* n_found, uint8_t* overflow, void* gm_workspace, size_t gm_workspace_bytes,
                             void* nl_workspace, size_t nl_workspace_bytes);

/* Stress times volume, dU/d(eps) at eps = 0.  Replaces stress_times_vol (apax/layers/properties.py:11-42) as called by
 * EnergyDerivativeModel with calc_stress (apax/nn/models.py:160-169) for the MINIMAGE displacement, where the reference
 * applies the strain as dr' = (I + eps) . dr (apax/utils/jax_md_reduced/space.py:277-278):
 *   stress[i][j] = sum over list entries of dE/d(dr)_i * dr_j.
 * Call it after apax_gm_compute / apax_gm_energy_forces (n_out == 1) on the SAME workspace: it reduces the per-pair
 * quantities the force pass left there.  (On the OFFSETS path the reference's disp_fn applies dr' = dr + (I + eps) . dr,
 * apax/layers/distances.py:16-17, i.e. it differentiates at doubled displacements: apax_gm_stress_times_vol_offsets below.)
 * stress f64 [3][3], device. */
int apax_gm_stress_times_vol(apax_stream_t stream, const apax_gm_params* params, int64_t n_atoms, int64_t n_pairs,
                             double* stress, void* workspace, size_t workspace_bytes);

/* The same quantity on the OFFSETS displacement, exactly as the reference computes it: under a perturbation its disp_fn
 * returns dR + (I + eps) . dR with dR = box . (R[i1] - R[i0]) and the offsets are added afterwards
 * (apax/layers/distances.py:12-22, 62-66), i.e. dU/d(eps) is taken at dr' = 2 dR + offsets:
 *   stress[i][j] = sum over list entries of dE/d(dr')_i * dR_j          (dR in fp64).
 * Protocol: (1) apax_gm_energy_forces with box2 = 2 * box, the same offsets and APAX_DISP_OFFSETS into scratch energy / force
 * buffers -- the evaluation at dr'; (2) this call with the ORIGINAL box and the same R on the SAME workspace.
 * R f64 [n_atoms][3], box f64 [3][3], stress f64 [3][3]: device. */
int apax_gm_stress_times_vol_offsets(apax_stream_t stream, const apax_gm_params* params, const double* R, const double* box,
                                     int64_t n_atoms, int64_t n_pairs, double* stress, void* workspace, size_t workspace_bytes);

/* Batched evaluation of many independent structures in one call.  Replaces the vmapped model of
 * ASECalculator.batch_eval (apax/md/ase_calc.py:395-481, jax.vmap over padded structures) and of the training
 * data path (apax/train/run.py:204).  Atoms of all structures are concatenated (no padding needed):
 * struct_ptr i32 [n_structs+1] gives each structure's atom range, idx holds GLOBAL atom indices, R is
 * Cartesian and offsets (f64 [n_pairs][3], Cartesian, NULL = zero) carry the periodic images, i.e.
 * dr = R[idx[1]] - R[idx[0]] + offsets.  energy f64 [n_structs][n_out]; forces f64 [n_out][n_atoms][3]. */
int apax_gm_energy_forces_batched(apax_stream_t stream, const apax_gm_params* params, const double* R,
                                  const int32_t* Z, const int32_t* idx, const double* offsets,
                                  const int32_t* struct_ptr, int64_t n_structs, int64_t n_atoms, int64_t n_pairs,
                                  double* energy, double* forces, void* workspace, size_t workspace_bytes);

/* Shallow-ensemble statistics.  Replaces the tail of ShallowEnsembleModel.__call__ (apax/nn/models.py:221-263) for the
 * outputs of apax_gm_energy_forces(_batched) with n_out = n_ens heads: energy_ens f64 [n_structs][n_ens], forces_ens f64
 * [n_ens][n_atoms][3] -> e_mean, e_unc f64 [n_structs] (unbiased: divisor 1/(n_ens-1)), f_mean, f_unc f64 [n_atoms][3],
 * and f_ens_t f64 [n_atoms][3][n_ens] (`forces_ensemble`; may be NULL). */
int apax_ensemble_stats(apax_stream_t stream, const double* energy_ens, const double* forces_ens, int64_t n_structs,
                        int64_t n_atoms, int32_t n_ens, double* e_mean, double* e_unc, double* f_mean, double* f_unc,
                        double* f_ens_t);

/* Domain-decomposed evaluation (no counterpart in the reference, which is single-device:
 * apax/md/simulate.py:313-354).  Atoms [0, n_central) are this rank's owned atoms and are the only
 * centrals (rows of the prepared list); atoms [n_central, n_atoms) are ghosts that only act as
 * neighbours.  energy = sum over the centrals; forces f64 [n_out][n_atoms][3] = dE_local/dR on ALL
 * local atoms, the ghost rows being what must be sent back to their owners. */
int apax_gm_compute_partial(apax_stream_t stream, const apax_gm_params* params, const double* R,
                            const int32_t* Z, const double* box, const double* offsets, int64_t n_atoms,
                            int64_t n_pairs, int64_t n_central, int32_t disp_mode, double* energy,
                            double* forces, void* workspace, size_t workspace_bytes);

/* Descriptor only.  Replaces GaussianMomentDescriptor.__call__ behind the builder seam
 * ModelBuilder.build_descriptor (apax/nn/builder.py:79-83,248-260;
 * apax/layers/descriptor/gaussian_moment_descriptor.py:31-103).  g: f32 [n_atoms][360] row-major out; with
 * n_contr < 8 the reference's output is the first n_feat columns. */
int apax_gm_descriptor(apax_stream_t stream, const apax_gm_params* params, const double* R,
                       const int32_t* Z, const int32_t* idx, const double* box, const double* offsets,
                       int64_t n_atoms, int64_t n_pairs, int32_t disp_mode, float* g, void* workspace,
                       size_t workspace_bytes);

/* Vector-Jacobian product of the descriptor: the backward half of the same seam, so that a JAX host can wrap the
 * descriptor ALONE in jax.custom_vjp and keep its own readout (ModelBuilder.build_descriptor, apax/nn/builder.py:79-83,
 * 248-260; the call it replaces is `representation(dr_vec, Z, idx)` at apax/nn/models.py:107 under jax.grad,
 * apax/utils/jax_md_reduced/quantity.py:52-54).  Inputs as apax_gm_descriptor, plus
 *   g_bar      f32 [n_atoms][360] row-major: the cotangent of g (columns beyond the n_contr truncation must be zero)
 *   dr_vec_bar f64 [n_pairs][3] out, COO order, or NULL: cotangent of dr_vec -- the VJP of `representation` itself; with
 *              APAX_DISP_DRVEC (`R` is dr_vec) this is the whole backward rule.  Entries with idx[0]==idx[1] (padding) get 0.
 *   R_bar      f64 [n_atoms][3] out, or NULL: cotangent of the positions for the three displacement branches, with the identity
 *              Jacobian of space.transform (apax/utils/jax_md_reduced/space.py:93-97), i.e. Cartesian also for fractional R.
 * The forward pair pass is recomputed here (nothing is kept from apax_gm_descriptor), so the call is self-contained. */
int apax_gm_descriptor_vjp(apax_stream_t stream, const apax_gm_params* params, const double* R, const int32_t* Z,
                           const int32_t* idx, const double* box, const double* offsets, int64_t n_atoms, int64_t n_pairs,
                           int32_t disp_mode, const float* g_bar, double* dr_vec_bar, double* R_bar, void* workspace,
                           size_t workspace_bytes);

/* ------------------------------------------------------------------------------------------------
 * Neighbour lists.
 * (1) apax_nl_build_minimage replaces partition.neighbor_list(...).allocate/update with
 *     format=Sparse, fractional_coordinates=True, disable_cell_list=True
 *     (apax/utils/jax_md_reduced/partition.py:478-787; predicate :640-664).  Output order is the
 *     reference's: idx[1] (central) ascending, idx[0] ascending within it; tail padded with n_atoms.
 *       frac     f64 [n_atoms][3] fractional positions;  box f64 [3][3] (columns = lattice vectors,
 *                i.e. atoms.cell.array.T as at apax/md/ase_calc.py:42-43)
 *       cutoff   r_max + dr_threshold (:574)
 *       idx      i32 [2][capacity] out;  n_found i32 [1] out = occupancy (may exceed capacity)
 *       overflow u8 [1] out: |= 1 (PEC.NEIGHBOR_LIST_OVERFLOW, :734) when occupancy > capacity
 * (2) apax_nl_build_images replaces the vesin NeighborList(cutoff, full_list=True).compute(..."ijS")
 *     calls (apax/md/ase_calc.py:337-355, apax/data/preprocessing.py:47-54): every (i, j, S) with
 *     |r_j - r_i + S.cell| < cutoff except (i,i,0), for any cell (also L < 2 cutoff, triclinic).
 *       pos      f64 [n_atoms][3] Cartesian; cell f64 [3][3] rows = lattice vectors; pbc 0/1
 *       idx      i32 [2][capacity] out (idx[0]=i, idx[1]=j, rows grouped by j, sorted by (i,S));
 *                tail padded with (0,0) like ase_calc.py:349-355
 *       shifts   i32 [capacity][3] out (S);  offsets f64 [capacity][3] out = S @ cell (:354), or NULL
 * (3) apax_nl_needs_rebuild replaces the skin test (partition.py:771-779): flag[0] = any atom moved
 *     more than dr_threshold/2 (min-image metric) since `ref`.
 * ---------------------------------------------------------------------------------------------- */
size_t apax_nl_workspace_bytes(int64_t n_atoms, int64_t capacity);
int apax_nl_build_minimage(apax_stream_t stream, const double* frac, const double* box,
                           int64_t n_atoms, double cutoff, int64_t capacity, int32_t* idx,
                           int32_t* n_found, uint8_t* overflow, void* workspace,
                           size_t workspace_bytes);
int apax_nl_build_images(apax_stream_t stream, const double* pos, const double* cell, int32_t pbc,
                         int64_t n_atoms, double cutoff, int64_t capacity, int32_t* idx,
                         int32_t* shifts, double* offsets, int32_t* n_found, uint8_t* overflow,
                         void* workspace, size_t workspace_bytes);
/* (2b) the same list for a batch of independent periodic cells (cells f64 [n_structs][3][3], rows = lattice
 *     vectors; struct_ptr i32 [n_structs+1]); idx holds global atom indices, rows grouped by central atom. */
size_t apax_nl_batched_workspace_bytes(int64_t n_atoms, int64_t capacity, int64_t n_structs);
int apax_nl_build_images_batched(apax_stream_t stream, const double* pos, const double* cells,
                                 const int32_t* struct_ptr, int64_t n_structs, int64_t n_atoms, double cutoff,
                                 int64_t capacity, int32_t* idx, int32_t* shifts, double* offsets, int32_t* n_found,
                                 uint8_t* overflow, void* workspace, size_t workspace_bytes);
int apax_nl_needs_rebuild(apax_stream_t stream, const double* frac, const double* ref,
                          const double* box, int64_t n_atoms, double dr_threshold, int32_t* flag);

/* ------------------------------------------------------------------------------------------------
 * NVT Nose-Hoover-chain integrator, on the device.  Replaces the per-step body of apax's MD loop
 * (apax/md/simulate.py:313-354) for the ensemble built at simulate.py:100-119, i.e. jax_md's
 * nvt_nose_hoover.apply_fn (apax/utils/jax_md_reduced/simulate.py:611-637):
 *   apax_md_nvt_pre_force : chain half step (:436-487), P += dt/2 F, R_frac = mod(R_frac + inv(box).(dt P/m), 1)
 *   -- caller evaluates F(R_frac) with apax_gm_compute --
 *   apax_md_nvt_post_force: P += dt/2 F, kinetic energy, chain half step, momentum rescale
 * R_frac, P, F f64 [n][3]; mass f64 [n]; box f64 [3][3] (columns = lattice vectors); state: device buffer of
 * apax_md_nvt_state_bytes() bytes initialised by apax_md_nvt_init (chain at rest, KE from P).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t chain_length; /* 5 (simulate.py:541) */
  int32_t chain_steps;  /* 2 */
  int32_t sy_steps;     /* 3 (1, 3, 5 or 7) */
  double dt;            /* time step (rounded to fp32 like the reference's f32(dt)) */
  double kT;
  double tau;           /* thermostat period; the reference default is 100 dt */
} apax_md_nvt_config;
size_t apax_md_nvt_state_bytes(void);
int apax_md_nvt_init(apax_stream_t stream, const apax_md_nvt_config* cfg, const double* P, const double* mass,
                     int64_t n_atoms, void* state);
int apax_md_nvt_pre_force(apax_stream_t stream, const apax_md_nvt_config* cfg, double* R_frac, double* P,
                          const double* F, const double* mass, const double* box, int64_t n_atoms, void* state);
int apax_md_nvt_post_force(apax_stream_t stream, const apax_md_nvt_config* cfg, double* P, const double* F,
                           const double* mass, int64_t n_atoms, void* state);
/* The same step for a box whose atoms are spread over several ranks (spatial domain decomposition): the Nose-Hoover chain
 * is replicated and driven by the GLOBAL kinetic energy, so apax_md_nvt_post_force splits at the reduction:
 *   apax_md_nvt_kick_ke      : P += dt/2 F on this rank's atoms (apply_kick != 0), ke_local f64 [1] = their kinetic energy
 *   -- caller sums ke_local over the ranks (apax_peer_allreduce_sum) --
 *   apax_md_nvt_chain_finish : chain half step with ke_global f64 [1] and dof = 3 N_global, momentum rescale of this rank's atoms
 * apax_md_nvt_pre_force is used unchanged on the owned atoms (the chain state then already holds the global KE and dof).
 * For the initial state: apax_md_nvt_init on the owned atoms, apax_md_nvt_kick_ke with apply_kick = 0, the reduction, then
 * apax_md_nvt_set_global. */
int apax_md_nvt_kick_ke(apax_stream_t stream, const apax_md_nvt_config* cfg, double* P, const double* F, const double* mass,
                        int64_t n_atoms, int32_t apply_kick, void* state, double* ke_local);
int apax_md_nvt_set_global(apax_stream_t stream, void* state, const double* ke_global, int64_t n_atoms_global);
int apax_md_nvt_chain_finish(apax_stream_t stream, const apax_md_nvt_config* cfg, double* P, int64_t n_atoms, void* state,
                             const double* ke_global, int64_t n_atoms_global);
/* n_steps steps in one call -- the jax.lax.fori_loop over apply_fn between two neighbour-list decisions
 * (apax/md/simulate.py:313-354): pre_force -> apax_gm_compute (MINIMAGE, list already prepared in `workspace` with
 * apax_gm_prepare, n_out == 1) -> post_force, repeated.  F holds the forces of the current positions on entry and on
 * exit; energy f64 [1] receives the last step's potential energy. */
int apax_md_nvt_run(apax_stream_t stream, const apax_md_nvt_config* cfg, const apax_gm_params* params, double* R_frac,
                    double* P, double* F, const double* mass, const int32_t* Z, const double* box, int64_t n_atoms,
                    int64_t n_pairs, void* state, double* energy, void* workspace, size_t workspace_bytes, int32_t n_steps);

/* ------------------------------------------------------------------------------------------------
 * Training step (BASELINE configs[1]).  Replaces, for the GMNN model with an energy + forces MSE loss,
 *   calc_loss + jax.value_and_grad w.r.t. params   apax/train/trainer.py:253-263, 332-336
 *   the vmapped model                              apax/train/run.py:204
 *   Loss.__call__ ("mse", atoms_exponent)           apax/train/loss.py:12-23, 199-244
 *   get_opt: clip -> adam(linear schedule) -> zero_nans per parameter group, lr <= 1e-7 freezes a group
 *                                                   apax/optimizer/get_optimizer.py:62-176
 * The gradient of the force loss (second order: F = -dE/dR differentiated w.r.t. the weights) is hand-derived
 * (forward-mode sweep along dL/dF, then one reverse sweep); no floating-point atomics, bitwise reproducible.
 *
 * All state is CALLER-owned device memory, so that a host can all-reduce the gradient blocks between the two
 * calls (data parallelism over structures, apax/train/trainer.py:101-106) and capture the step in a CUDA graph:
 *   theta32 / grad32 / m32 / v32 : f32 blocks of offsets32[7] elements laid out as apax_train_param_layout says
 *                                  (atomic_type_embedding [S][S][5][n_basis], dense_0 w [360][256], b, dense_1 w, b, dense_2 w, b)
 *   theta64 / grad64 / m64 / v64 : f64 [2 S] = scale_per_element, shift_per_element
 * Batch layout = apax_gm_energy_forces_batched: atoms of all structures concatenated, Cartesian positions,
 * idx i32 [2][n_pairs] with GLOBAL atom indices, offsets f64 [n_pairs][3] or NULL; struct_ptr / pair_ptr i32
 * [n_structs+1] give each structure's atom and COO-entry ranges; n_real i32 [n_structs] is inputs["n_atoms"]
 * (the divisor of loss.py:233; padding atoms have Z = 0).  Entries with idx[0]==idx[1] are padding.  The pair_ptr ranges
 * must tile [0, n_pairs) (pair_ptr[0] = 0, pair_ptr[n_structs] = n_pairs) and an entry must name two atoms of its own structure.
 *   loss f64 [1], energy f64 [n_structs], forces f64 [n_atoms][3] : outputs (the step's predictions)
 *   batch_divisor : batch size of the mean at loss.py:241 -- the GLOBAL batch under data parallelism, so that the
 *                   per-rank loss and gradients simply add up
 * ---------------------------------------------------------------------------------------------- */
typedef struct apax_train_plan apax_train_plan;
typedef struct {
  double energy_weight, energy_atoms_exponent; /* Loss(name="energy"): weight, atoms_exponent (train_config.py:296-299) */
  double forces_weight, forces_atoms_exponent; /* Loss(name="forces") */
} apax_train_loss_config;
typedef struct {
  double emb_lr, nn_lr, scale_lr, shift_lr; /* get_opt arguments (get_optimizer.py:90-97) */
  double gradient_clipping;                 /* optax.clip, element-wise */
  double end_value;                         /* LinearLR.end_value (lr_config.py:26) */
  int64_t transition_steps;                 /* n_epochs * steps_per_epoch (get_optimizer.py:54-56) */
  int64_t transition_begin;                 /* LinearLR.transition_begin */
  double b1, b2, eps;                       /* optax.adam defaults 0.9, 0.999, 1e-8 */
} apax_train_opt_config;
int apax_train_plan_create(const apax_gm_config* cfg, apax_train_plan** out);
void apax_train_plan_destroy(apax_train_plan* plan);
/* offsets32[0..6] = element offsets of emb, w0, b0, w1, b1, w2, b2; offsets32[7] = block length.
 * offsets64[0..1] = scale, shift; offsets64[2] = block length. */
int apax_train_param_layout(const apax_train_plan* plan, int64_t* offsets32, int64_t* offsets64);
size_t apax_train_workspace_bytes(const apax_train_plan* plan, int64_t n_atoms, int64_t n_pairs, int64_t n_structs);
int apax_train_loss_and_grad(apax_stream_t stream, const apax_train_plan* plan, const float* theta32,
                             const double* theta64, const double* R, const int32_t* Z, const int32_t* idx,
                             const double* offsets, const int32_t* struct_ptr, const int32_t* pair_ptr,
                             const int32_t* n_real, int64_t n_structs, int64_t n_atoms, int64_t n_pairs,
                             const double* E_label, const double* F_label, const apax_train_loss_config* loss_cfg,
                             double batch_divisor, double* loss, double* energy, double* forces, float* grad32,
                             double* grad64, void* workspace, size_t workspace_bytes);
/* Energy and forces only, on the same kernels and batch layout (no labels, no gradients): the evaluation path for
 * model statics the specialised apax_gm_* kernels do not cover (Bessel / exponential-spacing bases, n_basis up to 16);
 * replaces the vmapped EnergyDerivativeModel.apply (apax/nn/models.py:146-171, apax/train/run.py:204) for those. */
int apax_train_energy_forces(apax_stream_t stream, const apax_train_plan* plan, const float* theta32, const double* theta64,
                             const double* R, const int32_t* Z, const int32_t* idx, const double* offsets,
                             const int32_t* struct_ptr, const int32_t* pair_ptr, int64_t n_structs, int64_t n_atoms,
                             int64_t n_pairs, double* energy, double* forces, void* workspace, size_t workspace_bytes);
/* One optimizer update, in place.  opt_state i64 [8], zero-initialised by the caller: [0] is the optimizer's step counter
 * (optax `count`), read for the schedule / bias correction and incremented on the device; the rest is scratch. */
int apax_train_adam_step(apax_stream_t stream, const apax_train_plan* plan, const apax_train_opt_config* opt_cfg,
                         float* theta32, double* theta64, const float* grad32, const double* grad64, float* m32,
                         float* v32, double* m64, double* v64, int64_t* opt_state);

/* Data-parallel update without a library collective: all-reduce + Adam fused in one kernel over NVLink peer memory.
 * Replaces the gradient all-reduce XLA inserts for the sharded batch (apax/train/trainer.py:101-106) plus
 * state.apply_gradients (:335).  Every rank allocates its gradient blocks with apax_peer_alloc, publishes the 64-byte IPC
 * handle to its peers (any host channel), and maps theirs with apax_peer_open.  grad32[r] / grad64[r] / flags[r] are rank
 * r's fp32 block, fp64 block (scale, shift gradients and, behind them, the step's loss) and flag array (u32 [16], zeroed:
 * eight epoch flags, then a sticky error word)
 * AS MAPPED IN THIS PROCESS.  `epoch` must increase by one per step on every rank, and consecutive steps must use two
 * alternating sets of gradient blocks (the kernel of step t+1 may start while a peer still reads step t's block).
 * The gradients are summed in rank order, so all replicas stay bitwise identical; loss_out f64 [1] receives the global loss. */
typedef struct {
  const float* grad32[8];
  const double* grad64[8];
  uint32_t* flags[8];
  int32_t world, rank;
} apax_train_peers;
int apax_peer_alloc(int64_t bytes, void** dev_ptr, void* ipc_handle_out /* 64 bytes */);
int apax_peer_open(const void* ipc_handle /* 64 bytes */, void** dev_ptr);
int apax_peer_close(void* dev_ptr);
int apax_peer_free(void* dev_ptr);
int apax_train_allreduce_adam_step(apax_stream_t stream, const apax_train_plan* plan, const apax_train_opt_config* opt_cfg,
                                   const apax_train_peers* peers, uint32_t epoch, float* theta32, double* theta64, float* m32,
                                   float* v32, double* m64, double* v64, int64_t* opt_state, double* loss_out);
/* The cross-GPU waits inside the peer-memory kernels are bounded (30 s): a peer that died, raised on the host or issued a
 * different number of steps cannot hang this GPU.  A time-out sets a sticky word in this rank's own flag block and the
 * results of that step are invalid; this synchronising query returns APAX_ERR_CUDA once that has happened, else 0. */
int apax_train_peers_status(const apax_train_peers* peers);

/* Halo exchange of the slab-decomposed evaluation (apax_gm_compute_partial) as kernels over NVLink peer memory; no
 * counterpart in the reference, which is single-device (apax/md/simulate.py:313-354).  Every rank keeps, in one
 * apax_peer_alloc block mapped by its peers: a 512-byte zeroed header, then its local position rows and its local force
 * rows (f64 [n_local][3], owned atoms first, then ghosts grouped by owner rank).  `header[r]` is rank r's header and
 * `rows[r]` the first row, inside rank r's block, of the ghosts that THIS rank owns -- both as mapped in this process
 * (forward: in r's position rows; reverse: in r's force rows).  send_local i64 [n_send] (device) lists this rank's owned
 * atoms that are ghosts elsewhere, grouped by destination rank; send_ptr_host i64 [world+1] (host) delimits the groups.
 * `epoch` increases by one per evaluation on every rank.
 *   apax_halo_push_rows     : positions of my boundary atoms -> the ghost rows of my neighbours; returns (on the stream)
 *                             once every peer's rows have landed here as well.
 *   apax_halo_pull_add_rows : waits until every peer's forces are complete, then adds to my owned atoms the contributions
 *                             they received as ghosts elsewhere (rank order, deterministic) and sums the ranks' partial
 *                             energies (energy_local / energy_global f64 [n_out], device). */
typedef struct {
  double* rows[8];
  uint32_t* header[8];
  int32_t world, rank;
} apax_halo_peers;
int apax_halo_push_rows(apax_stream_t stream, const double* local_rows, const int64_t* send_local,
                        const int64_t* send_ptr_host, const apax_halo_peers* peers, uint32_t epoch);
int apax_halo_pull_add_rows(apax_stream_t stream, double* local_rows, const int64_t* send_local,
                            const int64_t* send_ptr_host, const apax_halo_peers* peers, uint32_t epoch,
                            const double* energy_local, double* energy_global, int32_t n_out);
/* Synchronising health query (see apax_train_peers_status): APAX_ERR_CUDA once one of this rank's halo barriers timed out. */
int apax_halo_status(const apax_halo_peers* peers);
/* Sum of n <= 16 doubles over the ranks through the same peer headers (kinetic energy and the skin flag of the
 * domain-decomposed MD loop): global_out[i] = sum over ranks of local[i], added in rank order so that every rank holds the
 * bitwise identical result.  `epoch` is this reduction's own counter (+1 per call on every rank); only peers->header is used. */
int apax_peer_allreduce_sum(apax_stream_t stream, const apax_halo_peers* peers, uint32_t epoch, const double* local,
                            double* global_out, int32_t n);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer convenience context: what ASECalculator.calculate does per call
 * (apax/md/ase_calc.py:357-393): positions/cell on the HOST in, neighbour list (min-image when
 * min cell height / 2 > r_max, image list otherwise, :501-522), E+F, results on the HOST out.
 * Owns its device buffers and stream; includes the H2D/D2H copies.  Used for the end-to-end
 * measurement in bench.py and by non-Python hosts.
 * ---------------------------------------------------------------------------------------------- */
typedef struct apax_ctx apax_ctx;
int apax_ctx_create(const apax_gm_params* params, int64_t max_atoms, int64_t max_pairs, apax_ctx** out);
void apax_ctx_destroy(apax_ctx* ctx);
/* positions_host f64 [n][3] Cartesian, numbers_host i32 [n], cell_host f64 [3][3] rows = lattice
 * vectors (all-zero = gas phase is not supported here: use apax_gm_energy_forces with a list),
 * dr_threshold = skin; rebuild = 1 forces a list rebuild, 0 applies the reference's skin test.
 * energy_host f64 [n_out], forces_host f64 [n_out][n][3] ([1][n][3] with APAX_OPT_MEAN_FORCES). */
int apax_ctx_calculate(apax_ctx* ctx, const double* positions_host, const int32_t* numbers_host,
                       const double* cell_host, int64_t n_atoms, double dr_threshold, int32_t rebuild,
                       double* energy_host, double* forces_host, int64_t* n_pairs_host);
/* The `stress` entry of the calculator's results for the structure of the LAST apax_ctx_calculate (which must have asked for
 * forces; n_out == 1): EnergyDerivativeModel with calc_stress (apax/nn/models.py:160-169 -> stress_times_vol,
 * apax/layers/properties.py:11-42) followed by ProcessStress (apax/md/function_transformations.py:140-159),
 * stress = (dU/d eps)^T / det(box).  stress_host f64 [3][3].  On the image-list path the reference differentiates at doubled
 * displacements (apax/layers/distances.py:16-17), which costs a second force pass; a further stress call then needs a new
 * apax_ctx_calculate first (APAX_ERR_INVALID_ARG otherwise). */
int apax_ctx_stress(apax_ctx* ctx, double* stress_host);

/* Test entry of the exact integer tensor-core readout GEMM (tcgen05.mma kind::i8, four base-256 digit planes per
 * operand, int32 accumulators): out[nc][ld] = act^T . wgt with act a device fp32 [k][ld] feature-major matrix and
 * wgt a device fp32 [k][nc] matrix.  Replaces the fp32 matmuls of apax/layers/ntk_linear.py:40 inside the
 * readout; allocates its own scratch, so it is for tests and measurements only. */
int apax_i8_gemm_test(apax_stream_t stream, const float* act, int64_t k, int64_t ld, int64_t n_atoms, const float* wgt,
                      int64_t nc, float* out, int32_t* err_flag, int64_t* dbg_cycles /* nullable, [n_sm][8] */,
                      int32_t act_planes /* 3 or 4 digit planes for the activations */);

/* Measurement only (tools/trace_i8.py, tools/gemm_ab.py): dbg_flags != 0 overrides the kernels' debug switches -- 1 the
 * epilogue skips the drain, 2 the producer signals without loading, 16 the sixteen-warp epilogue kernels, 32 / 64 one / two
 * fewer resident activation planes than fit (a deeper weight ring) --, trace != 0 makes every readout GEMM print its per-role
 * cycle counters (MMA warp: waits for accumulator / weight stages / atom planes; epilogue: waits, digit emission) to stderr
 * (synchronising).  Both 0 in normal operation; results with flags 1 or 2 are meaningless. */
int apax_i8_set_debug(int32_t dbg_flags, int32_t trace);

/* Measurement only: tcgen05.mma kind::i8 issue rate (M = 128, N = n; `iters` rounds of ten MMAs; pattern 0 / 1 / 2:
 * one accumulator / four accumulators / the readout GEMM's digit-pair schedule; negative: none) and tensor-memory
 * drain rate (`reads` passes of tcgen05.ld over all 512 columns by four warps), alone or concurrently, one CTA per
 * SM.  out_cycles[cta][0] = MMA cycles, [cta][1] = drain cycles. */
int apax_i8_mma_probe(apax_stream_t stream, int32_t n, int32_t pattern, int32_t iters, int32_t reads, int64_t* out_cycles);

/* Readout GEMM building block on the tcgen05 tensor cores, exposed for tests and measurement:
 * out[nc][ld] = sum_k act[k][atom] * wgt[k][n] with both operands given as exact tf32 (hi, lo) pairs
 * (3xTF32, fp32-accurate; replaces jnp.dot in apax/layers/ntk_linear.py:40).  act_* f32 [k][ld],
 * wgt_* f32 [k][nc], nc % 128 == 0, ld % 128 == 0; err_flag i32 [1] is set to 1 on a pipeline time-out. */
int apax_tc_gemm_test(apax_stream_t stream, const float* act_hi, const float* act_lo, int64_t k, int64_t ld,
                      int64_t n_atoms, const float* wgt_hi, const float* wgt_lo, int64_t nc, float* out,
                      int32_t* err_flag, float* debug_dump /* NULL, or f32 [16386] */, int32_t debug_flags /* 0 */);

/* Diagnostics */
int apax_abi_version(void);
const char* apax_last_cuda_error(void);
/* number of kernels launched by this library since process start (bench.py's gpu_launches) */
int64_t apax_kernel_launch_count(void);
/* Per-launch device timing for measurement: while enabled, every kernel of this library is
 * bracketed by CUDA events on its launching stream.  apax_profile_collect waits for them and writes
 * up to max_records (kernel name, milliseconds) pairs in launch order; returns the count (>= 0).
 * apax_profile_enable(0/1) also clears the record list.  A measurement aid behind one process-wide switch (guarded by a
 * mutex); off by default, and the only mutable global state of the library besides the launch counter. */
int apax_profile_enable(int on);
int apax_profile_collect(char* names_host, int name_stride, float* ms_host, int max_records);

#ifdef __cplusplus
}
#endif
#endif /* APAX_B200_H */
