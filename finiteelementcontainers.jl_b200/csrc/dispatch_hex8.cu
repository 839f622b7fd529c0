// dispatch_hex8.cu -- HEX8 instantiations of the element kernels (3-D, 8 nodes).
// Hot configurations (BASELINE.json configs 2, 3, 5): Poisson NF=1 and neo-Hookean NF=3 with 8-point rules.
#include "kernel_mat2.cuh"
#include "kernel_mat_scalar.cuh"

// Warps per k_mat2 CTA.  The warps of a CTA start together and walk the phases (FP64-bound G/K, RED-bound S) in
// lockstep, so the FP64 pipes idle while a whole CTA scatters: one-shot CTAs of 8 / 4 / 2 / 1 warps measured
// 22.9 / 17.9 / 16.8 / 17.2 ms for the fused kernel at 192^3 (same 8 warps per SM in every case).
#ifndef FEC_MAT2_WARPS
#define FEC_MAT2_WARPS 2
#endif
// Column-split variant (kernel_mat2c.cuh): FEC_MAT2C = 1 selects it, FEC_MAT2C_EPC elements per CTA (12 threads each),
// FEC_MAT2C_MINB CTAs per SM (sets the register cap: 5 -> 136, 4 -> 168).
#ifndef FEC_MAT2C
#define FEC_MAT2C 0
#endif
#ifndef FEC_MAT2C_EPC
#define FEC_MAT2C_EPC 8
#endif
#ifndef FEC_MAT2C_MINB
#define FEC_MAT2C_MINB 5
#endif
// Warp-specialised persistent variant (kernel_mat2w.cuh): FEC_MAT2W = 1 selects it; producer teams, consumer teams,
// ring stages and the register cap are FEC_MAT2W_PT / _CT / _NST / _REG.
#ifndef FEC_MAT2W
#define FEC_MAT2W 0
#endif
#ifndef FEC_MAT2W_PT
#define FEC_MAT2W_PT 3
#endif
#ifndef FEC_MAT2W_CT
#define FEC_MAT2W_CT 3
#endif
#ifndef FEC_MAT2W_NST
#define FEC_MAT2W_NST 5
#endif
#ifndef FEC_MAT2W_REG
#define FEC_MAT2W_REG 128
#endif
// the two measured-slower variants live outside the product tree (tools/variants/, evidence in profiles/r01s_*, r01y_*);
// they are compiled only by tools/build_variants.sh
#if FEC_MAT2C
#include "../../tools/variants/kernel_mat2c.cuh"
#endif
#if FEC_MAT2W
#include "../../tools/variants/kernel_mat2w.cuh"
#endif
#ifndef FEC_MAT3
#define FEC_MAT3 0
#endif
#if FEC_MAT3   // warp-specialised persistent variant (4 producer + 4 consumer warps, mbarrier ring): 22.2 vs 16.0 ms
#include "../../tools/variants/kernel_mat3.cuh"
#endif
#include <cstdlib>

namespace fec {

template <class Phys, int NF, int MINB>
static void vec8(fecb200_handle* h, BlockPlan& b, const VecLaunch& a) {
  FEC_REQUIRE(b.nq == 8, "HEX8: this physics is compiled for 8-point quadrature rules only");
  run_vec_modes<3, 8, NF, 8, Phys, kTE, MINB>(h, b, a);
}
template <class Phys, int NF, int EPB>
static void mat8(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  FEC_REQUIRE(b.nq == 8, "HEX8: this physics is compiled for 8-point quadrature rules only");
  // symmetric tangents (all shipped mechanics physics): pair-owner kernel with staged, coalesced REDs
  if (a.kind == FECB200_STIFFNESS && matrix_kernel_fuses_residual(h, b)) {
#if FEC_MAT2W
    run_mat2w<3, 8, NF, 8, Phys, Mat2wShape<FEC_MAT2W_PT, FEC_MAT2W_CT, FEC_MAT2W_NST>, FEC_MAT2W_REG>(h, b, a);
#elif FEC_MAT2C
    run_mat2c<3, 8, NF, 8, Phys, FEC_MAT2C_EPC, FEC_MAT2C_MINB>(h, b, a);
#else
#if FEC_MAT3
    if constexpr (NF == 3) { run_mat3<3, 8, NF, 8, Phys>(h, b, a); return; }
#endif
    run_mat2<3, 8, NF, 8, Phys, FEC_MAT2_WARPS>(h, b, a);
#endif
  }
  else run_mat<3, 8, NF, 8, Phys, EPB>(h, b, a);
}

// true when launch_matrix for this block runs k_mat2, which can produce the residual in the same pass
bool matrix_kernel_fuses_residual(fecb200_handle* h, const BlockPlan& b) {
  return b.elem_type == FECB200_HEX8 && h->nf == 3 && b.nq == 8 && !getenv("FECB200_KMAT1") && b.d_emeta.p && (int64_t)nz_alloc_len(h) < (int64_t)0xFFFFFFFFll &&
         (b.physics == FECB200_PHYS_LINEAR_ELASTIC || b.physics == FECB200_PHYS_NEOHOOKEAN ||
          b.physics == FECB200_PHYS_NEOHOOKEAN_AS_WRITTEN || b.physics == FECB200_PHYS_J2_PLASTICITY);
}

void launch_vector_hex8(fecb200_handle* h, BlockPlan& b, const VecLaunch& a) {
  switch (b.physics) {
    case FECB200_PHYS_POISSON:
      FEC_REQUIRE(h->nf == 1, "Poisson needs NF = 1");
      if (b.nq == 8) run_vec_modes<3, 8, 1, 8, PhysPoisson<3>, kTE, kMinB1>(h, b, a);
      else run_vec_modes<3, 8, 1, 0, PhysPoisson<3>, kTE, kMinB1>(h, b, a);
      break;
    case FECB200_PHYS_LINEAR_ELASTIC: FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND"); vec8<PhysLinearElastic<3>, 3, kMinB3>(h, b, a); break;
    case FECB200_PHYS_NEOHOOKEAN: FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND"); vec8<PhysNeoHookean<3>, 3, kMinB3>(h, b, a); break;
    case FECB200_PHYS_NEOHOOKEAN_AS_WRITTEN: FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND"); vec8<PhysNeoHookeanAsWritten<3>, 3, kMinB3>(h, b, a); break;
    case FECB200_PHYS_J2_PLASTICITY: FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND"); vec8<PhysJ2<3>, 3, kMinB3>(h, b, a); break;
    case FECB200_PHYS_TEST_NONSYMMETRIC: FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND"); vec8<PhysNonSymmetricTest<3>, 3, kMinB3>(h, b, a); break;
    default: throw Error("fecb200: unsupported physics for HEX8");
  }
}

void launch_matrix_hex8(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  switch (b.physics) {
    case FECB200_PHYS_POISSON:
      FEC_REQUIRE(h->nf == 1, "Poisson needs NF = 1");
      if (b.nq == 8 && b.d_emeta.p && !b.emeta_trash_rows && h->nnz + 4096 < (int64_t)0xFFFFFFFFll && !getenv("FECB200_KMAT1"))
        run_mat_scalar<3, 8, 8, PhysPoisson<3>>(h, b, a);   // register-resident thread-per-element kernel
      else if (b.nq == 8) run_mat<3, 8, 1, 8, PhysPoisson<3>, 32>(h, b, a);
      else run_mat<3, 8, 1, 0, PhysPoisson<3>, 32>(h, b, a);
      break;
    case FECB200_PHYS_LINEAR_ELASTIC: mat8<PhysLinearElastic<3>, 3, 16>(h, b, a); break;
    case FECB200_PHYS_NEOHOOKEAN: mat8<PhysNeoHookean<3>, 3, 16>(h, b, a); break;
    case FECB200_PHYS_NEOHOOKEAN_AS_WRITTEN: mat8<PhysNeoHookeanAsWritten<3>, 3, 16>(h, b, a); break;
    case FECB200_PHYS_J2_PLASTICITY: mat8<PhysJ2<3>, 3, 16>(h, b, a); break;
    case FECB200_PHYS_TEST_NONSYMMETRIC: run_mat<3, 8, 3, 8, PhysNonSymmetricTest<3>, 16>(h, b, a); break;   // k_mat honours the transposed convention
    default: throw Error("fecb200: unsupported physics for HEX8");
  }
}

void launch_scalar_hex8(fecb200_handle* h, BlockPlan& b, const double* U) {
  switch (b.physics) {
    case FECB200_PHYS_POISSON: FEC_REQUIRE(h->nf == 1, "Poisson needs NF = 1"); run_energy<3, 8, 1, 0, PhysPoisson<3>>(h, b, U); break;
    case FECB200_PHYS_LINEAR_ELASTIC: FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND"); run_energy<3, 8, 3, 0, PhysLinearElastic<3>>(h, b, U); break;
    case FECB200_PHYS_NEOHOOKEAN: FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND"); run_energy<3, 8, 3, 0, PhysNeoHookean<3>>(h, b, U); break;
    case FECB200_PHYS_NEOHOOKEAN_AS_WRITTEN: FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND"); run_energy<3, 8, 3, 0, PhysNeoHookeanAsWritten<3>>(h, b, U); break;
    case FECB200_PHYS_J2_PLASTICITY: run_energy<3, 8, 3, 0, PhysJ2<3>>(h, b, U); break;
    default: throw Error("fecb200: unsupported physics for HEX8");
  }
}

}  // namespace fec
