// XLA-FFI custom-call layer over the C ABI of libapax_b200.so (include/apax_b200.h) -- the binding `north_star` names for
// keeping apax's host code in Python/JAX.  One handler per C entry point; every handler only unpacks XLA buffers and
// forwards them (no arithmetic here).  Buffers are XLA-owned device memory, the stream comes from the call frame, scratch
// is a result buffer sized by the host with apax_*_workspace_bytes.  The library's device-side entry points never allocate, never synchronise and
// keep no mutable global state besides a launch counter and the opt-in measurement switches, so the handlers may run on any XLA runtime thread and inside CUDA-graph capture.
//
// jaxlib's headers (xla/ffi/api/ffi.h) are NOT in this repository's build image, so this file is compiled only by
// apax_b200/ffi/build_ffi.py when `jax.ffi.include_dir()` exists; INTEGRATION.md s2 shows the Python side
// (jax.ffi.register_ffi_target / jax.ffi.ffi_call / jax.custom_vjp).
//
// Replaces, behind apax's Flax modules: EnergyDerivativeModel.apply (apax/nn/models.py:146-171), the jax_md neighbour list
// (apax/utils/jax_md_reduced/partition.py:478-787), GaussianMomentDescriptor.__call__
// (apax/layers/descriptor/gaussian_moment_descriptor.py:31-103), stress_times_vol (apax/layers/properties.py:11-42) and update_step
// (apax/train/trainer.py:332-336).  Empirical corrections travel with the parameter handle (apax_gm_params_set_corrections).
#include <cstdint>

#include "apax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error status(int rc, const char* where) {
  if (rc == APAX_OK) return ffi::Error::Success();
  return ffi::Error(rc == APAX_ERR_INVALID_ARG ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    std::string(where) + ": " + (rc == APAX_ERR_CUDA ? apax_last_cuda_error() : "apax_status " + std::to_string(rc)));
}

template <class T>
const T* opt(const ffi::AnyBuffer& b) {   // zero-sized buffer == "not given" (e.g. offsets on the min-image path)
  return b.element_count() == 0 ? nullptr : reinterpret_cast<const T*>(b.untyped_data());
}

// ---- energy + forces given a list -------------------------------------------------------------------------------------
ffi::Error EnergyForces(cudaStream_t stream, int64_t params, int32_t disp_mode, ffi::Buffer<ffi::F64> R, ffi::Buffer<ffi::S32> Z,
                        ffi::Buffer<ffi::S32> idx, ffi::Buffer<ffi::F64> box, ffi::AnyBuffer offsets,
                        ffi::ResultBuffer<ffi::F64> energy, ffi::ResultBuffer<ffi::F64> forces, ffi::ResultBuffer<ffi::U8> workspace) {
  const int64_t n = R.dimensions()[0], p = idx.dimensions()[1];
  return status(apax_gm_energy_forces(reinterpret_cast<apax_stream_t>(stream), reinterpret_cast<const apax_gm_params*>(params),
                                      R.typed_data(), Z.typed_data(), idx.typed_data(), box.typed_data(), opt<double>(offsets), n, p,
                                      disp_mode, energy->typed_data(), forces->typed_data(), workspace->typed_data(),
                                      workspace->element_count()),
                "apax_gm_energy_forces");
}

// ---- energy + forces + stress x volume (EnergyDerivativeModel with calc_stress, apax/nn/models.py:160-169) ------------------
// MINIMAGE: the strain enters as (I + eps) . dr (space.py:277-278) and the virial is reduced from the records the force pass
// left in the workspace.  OFFSETS: the reference's disp_fn differentiates at doubled displacements
// (apax/layers/distances.py:12-22); the caller passes box2 = 2 * box (a trivial XLA op) and two scratch results for the second
// force pass, as include/apax_b200.h describes for apax_gm_stress_times_vol_offsets.
ffi::Error EnergyForcesStress(cudaStream_t stream, int64_t params, int32_t disp_mode, ffi::Buffer<ffi::F64> R, ffi::Buffer<ffi::S32> Z,
                              ffi::Buffer<ffi::S32> idx, ffi::Buffer<ffi::F64> box, ffi::AnyBuffer offsets, ffi::AnyBuffer box2,
                              ffi::ResultBuffer<ffi::F64> energy, ffi::ResultBuffer<ffi::F64> forces,
                              ffi::ResultBuffer<ffi::F64> stress, ffi::ResultBuffer<ffi::F64> energy2,
                              ffi::ResultBuffer<ffi::F64> forces2, ffi::ResultBuffer<ffi::U8> workspace) {
  const int64_t n = R.dimensions()[0], p = idx.dimensions()[1];
  auto s = reinterpret_cast<apax_stream_t>(stream);
  auto prm = reinterpret_cast<const apax_gm_params*>(params);
  int rc = apax_gm_energy_forces(s, prm, R.typed_data(), Z.typed_data(), idx.typed_data(), box.typed_data(), opt<double>(offsets), n, p,
                                 disp_mode, energy->typed_data(), forces->typed_data(), workspace->typed_data(),
                                 workspace->element_count());
  if (rc != APAX_OK) return status(rc, "apax_gm_energy_forces");
  if (disp_mode == APAX_DISP_MINIMAGE)
    return status(apax_gm_stress_times_vol(s, prm, n, p, stress->typed_data(), workspace->typed_data(), workspace->element_count()),
                  "apax_gm_stress_times_vol");
  if (disp_mode != APAX_DISP_OFFSETS || box2.element_count() != 9)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "stress needs a periodic displacement (and box2 = 2 box on the offsets path)");
  rc = apax_gm_energy_forces(s, prm, R.typed_data(), Z.typed_data(), idx.typed_data(), opt<double>(box2), opt<double>(offsets), n, p,
                             disp_mode, energy2->typed_data(), forces2->typed_data(), workspace->typed_data(),
                             workspace->element_count());
  if (rc != APAX_OK) return status(rc, "apax_gm_energy_forces (doubled box)");
  return status(apax_gm_stress_times_vol_offsets(s, prm, R.typed_data(), box.typed_data(), n, p, stress->typed_data(),
                                                 workspace->typed_data(), workspace->element_count()),
                "apax_gm_stress_times_vol_offsets");
}

// ---- descriptor only (the ModelBuilder.build_descriptor seam, apax/nn/builder.py:79-83) -------------------------------
ffi::Error Descriptor(cudaStream_t stream, int64_t params, int32_t disp_mode, ffi::Buffer<ffi::F64> R, ffi::Buffer<ffi::S32> Z,
                      ffi::Buffer<ffi::S32> idx, ffi::AnyBuffer box, ffi::AnyBuffer offsets, ffi::ResultBuffer<ffi::F32> g,
                      ffi::ResultBuffer<ffi::U8> workspace) {
  const int64_t n = Z.dimensions()[0], p = idx.dimensions()[1];
  return status(apax_gm_descriptor(reinterpret_cast<apax_stream_t>(stream), reinterpret_cast<const apax_gm_params*>(params),
                                   R.typed_data(), Z.typed_data(), idx.typed_data(), opt<double>(box), opt<double>(offsets), n, p,
                                   disp_mode, g->typed_data(), workspace->typed_data(), workspace->element_count()),
                "apax_gm_descriptor");
}

// ---- VJP of the descriptor: the backward rule a jax.custom_vjp around the same seam needs -------------------------------
// want = 1: cotangent of dr_vec [P,3] (result d_dr); 2: cotangent of the positions [N,3] (result d_R); 3: both.  The result
// that is not wanted is a zero-sized buffer.
ffi::Error DescriptorVjp(cudaStream_t stream, int64_t params, int32_t disp_mode, ffi::Buffer<ffi::F64> R, ffi::Buffer<ffi::S32> Z,
                         ffi::Buffer<ffi::S32> idx, ffi::AnyBuffer box, ffi::AnyBuffer offsets, ffi::Buffer<ffi::F32> g_bar,
                         ffi::ResultBuffer<ffi::F64> d_dr, ffi::ResultBuffer<ffi::F64> d_R, ffi::ResultBuffer<ffi::U8> workspace) {
  const int64_t n = Z.dimensions()[0], p = idx.dimensions()[1];
  return status(apax_gm_descriptor_vjp(reinterpret_cast<apax_stream_t>(stream), reinterpret_cast<const apax_gm_params*>(params),
                                       R.typed_data(), Z.typed_data(), idx.typed_data(), opt<double>(box), opt<double>(offsets), n, p,
                                       disp_mode, g_bar.typed_data(), d_dr->element_count() ? d_dr->typed_data() : nullptr,
                                       d_R->element_count() ? d_R->typed_data() : nullptr, workspace->typed_data(),
                                       workspace->element_count()),
                "apax_gm_descriptor_vjp");
}

// ---- min-image neighbour list (partition.neighbor_list allocate/update) ------------------------------------------------
ffi::Error NeighborListMinImage(cudaStream_t stream, double cutoff, ffi::Buffer<ffi::F64> frac, ffi::Buffer<ffi::F64> box,
                                ffi::ResultBuffer<ffi::S32> idx, ffi::ResultBuffer<ffi::S32> n_found,
                                ffi::ResultBuffer<ffi::U8> overflow, ffi::ResultBuffer<ffi::U8> workspace) {
  const int64_t n = frac.dimensions()[0], cap = idx->dimensions()[1];
  return status(apax_nl_build_minimage(reinterpret_cast<apax_stream_t>(stream), frac.typed_data(), box.typed_data(), n, cutoff, cap,
                                       idx->typed_data(), n_found->typed_data(), overflow->typed_data(), workspace->typed_data(),
                                       workspace->element_count()),
                "apax_nl_build_minimage");
}

ffi::Error NeedsRebuild(cudaStream_t stream, double dr_threshold, ffi::Buffer<ffi::F64> frac, ffi::Buffer<ffi::F64> ref,
                        ffi::Buffer<ffi::F64> box, ffi::ResultBuffer<ffi::S32> flag) {
  return status(apax_nl_needs_rebuild(reinterpret_cast<apax_stream_t>(stream), frac.typed_data(), ref.typed_data(), box.typed_data(),
                                      frac.dimensions()[0], dr_threshold, flag->typed_data()),
                "apax_nl_needs_rebuild");
}

// ---- training step: loss + gradients, then the optimizer update (parameter / moment blocks are aliased in place by the
// caller with input_output_aliases) ---------------------------------------------------------------------------------------
ffi::Error TrainLossAndGrad(cudaStream_t stream, int64_t plan, double w_e, double exp_e, double w_f, double exp_f, double batch_divisor,
                            ffi::Buffer<ffi::F32> theta32, ffi::Buffer<ffi::F64> theta64, ffi::Buffer<ffi::F64> R,
                            ffi::Buffer<ffi::S32> Z, ffi::Buffer<ffi::S32> idx, ffi::AnyBuffer offsets, ffi::Buffer<ffi::S32> struct_ptr,
                            ffi::Buffer<ffi::S32> pair_ptr, ffi::Buffer<ffi::S32> n_real, ffi::Buffer<ffi::F64> e_label,
                            ffi::Buffer<ffi::F64> f_label, ffi::ResultBuffer<ffi::F64> loss, ffi::ResultBuffer<ffi::F64> energy,
                            ffi::ResultBuffer<ffi::F64> forces, ffi::ResultBuffer<ffi::F32> grad32, ffi::ResultBuffer<ffi::F64> grad64,
                            ffi::ResultBuffer<ffi::U8> workspace) {
  const apax_train_loss_config lc{w_e, exp_e, w_f, exp_f};
  const int64_t n = R.dimensions()[0], p = idx.dimensions()[1], b = e_label.dimensions()[0];
  return status(apax_train_loss_and_grad(reinterpret_cast<apax_stream_t>(stream), reinterpret_cast<const apax_train_plan*>(plan),
                                         theta32.typed_data(), theta64.typed_data(), R.typed_data(), Z.typed_data(), idx.typed_data(),
                                         opt<double>(offsets), struct_ptr.typed_data(), pair_ptr.typed_data(), n_real.typed_data(), b, n, p,
                                         e_label.typed_data(), f_label.typed_data(), &lc, batch_divisor, loss->typed_data(),
                                         energy->typed_data(), forces->typed_data(), grad32->typed_data(), grad64->typed_data(),
                                         workspace->typed_data(), workspace->element_count()),
                "apax_train_loss_and_grad");
}

}  // namespace

#define APAX_FFI_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(ApaxEnergyForces, EnergyForces,
                              APAX_FFI_STREAM.Attr<int64_t>("params").Attr<int32_t>("disp_mode")
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>(),
                              {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(ApaxEnergyForcesStress, EnergyForcesStress,
                              APAX_FFI_STREAM.Attr<int64_t>("params").Attr<int32_t>("disp_mode")
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>(),
                              {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(ApaxDescriptor, Descriptor,
                              APAX_FFI_STREAM.Attr<int64_t>("params").Attr<int32_t>("disp_mode")
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::U8>>(),
                              {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(ApaxDescriptorVjp, DescriptorVjp,
                              APAX_FFI_STREAM.Attr<int64_t>("params").Attr<int32_t>("disp_mode")
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>(),
                              {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(ApaxNeighborListMinImage, NeighborListMinImage,
                              APAX_FFI_STREAM.Attr<double>("cutoff").Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::U8>>(),
                              {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(ApaxNeedsRebuild, NeedsRebuild,
                              APAX_FFI_STREAM.Attr<double>("dr_threshold").Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::S32>>(),
                              {ffi::Traits::kCmdBufferCompatible});
XLA_FFI_DEFINE_HANDLER_SYMBOL(ApaxTrainLossAndGrad, TrainLossAndGrad,
                              APAX_FFI_STREAM.Attr<int64_t>("plan").Attr<double>("energy_weight").Attr<double>("energy_atoms_exponent")
                                  .Attr<double>("forces_weight").Attr<double>("forces_atoms_exponent").Attr<double>("batch_divisor")
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>(),
                              {ffi::Traits::kCmdBufferCompatible});
