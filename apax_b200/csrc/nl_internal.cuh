// Internal layout of the neighbour-list workspace.
#pragma once
#include "common.cuh"

namespace apax {

// Cell grid, computed ON THE DEVICE from the (device-resident) box so that no host round trip is
// needed; all launches are sized by n_atoms and the number of cells is capped at n_atoms.
struct NlGrid {
  double B[9];     // columns = lattice vectors: r = B f
  double Binv[9];  // f = Binv r
  double cutoff_sq;
  int nc[3];       // cells per fractional axis
  int m[3];        // stencil half-width per axis (in cells, may wrap several times for tiny boxes)
  int ncell;
  int periodic;
};

struct NlWorkspace {
  int64_t n_atoms, capacity;
  NlGrid* grid;
  int32_t* cell_of;     // [N]
  int32_t* wrap;        // [N][3] integer part of the fractional coordinate (images mode)
  int32_t* cell_cnt;    // [N+1]
  int32_t* cellptr;     // [N+1]
  int32_t* cell_tmp;    // [N]
  int32_t* cell_atoms;  // [N] atoms of each cell, ascending
  float4* cell_pos;     // [N] per cell slot: wrapped fractional position in fp32 + atom index (pre-filter stream)
  int32_t* rowcnt;      // [N+1]
  int32_t* rowptr;      // [N+1] (unclamped occupancy prefix)
  int32_t* rowptr_c;    // [N+1] clamped to capacity
  int32_t* nbr_tmp;     // [cap]
  int32_t* nbr;         // [cap]
  int32_t* shift_tmp;   // [cap] packed image shift
  int32_t* shift;       // [cap]
  int32_t* scan_scratch;
};

size_t carve_nl_workspace(void* base, int64_t n_atoms, int64_t capacity, NlWorkspace* ws);

// device-pointer variants used by the host-buffer context (ctx.cu)
int nl_build_minimage_csr(cudaStream_t stream, const double* frac, const double* box, int64_t n_atoms, double cutoff,
                          int64_t capacity, const NlWorkspace& w, int32_t* idx, int32_t* n_found, uint8_t* overflow);
int nl_build_images_csr(cudaStream_t stream, const double* pos, const double* cell, int32_t pbc, int64_t n_atoms,
                        double cutoff, int64_t capacity, const NlWorkspace& w, int32_t* idx, int32_t* shifts,
                        double* offsets, int32_t* n_found, uint8_t* overflow);

// Cartesian offsets S @ cell per CSR slot of the last image-list build (apax/md/ase_calc.py:354)
int nl_emit_offsets(cudaStream_t stream, const NlWorkspace& w, int64_t n_atoms, double* offsets);

}  // namespace apax
