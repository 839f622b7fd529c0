// dispatch_tet.cu -- TET4 / TET10 instantiations (BASELINE.json config 4: stateful J2 on tet10).
#include "kernels.cuh"

namespace fec {

template <int NNPE>
static void vec3d(fecb200_handle* h, BlockPlan& b, const VecLaunch& a) {
  switch (b.physics) {
    case FECB200_PHYS_POISSON:
      FEC_REQUIRE(h->nf == 1, "Poisson needs NF = 1");
      run_vec_modes<3, NNPE, 1, 0, PhysPoisson<3>, kTE, kMinB1>(h, b, a);
      break;
    case FECB200_PHYS_LINEAR_ELASTIC:
      FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND");
      run_vec_modes<3, NNPE, 3, 0, PhysLinearElastic<3>, kTE, kMinB3>(h, b, a);
      break;
    case FECB200_PHYS_NEOHOOKEAN:
      FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND");
      run_vec_modes<3, NNPE, 3, 0, PhysNeoHookean<3>, kTE, kMinB3>(h, b, a);
      break;
    case FECB200_PHYS_J2_PLASTICITY:
      FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND");
      if (b.nq == 4) run_vec_modes<3, NNPE, 3, 4, PhysJ2<3>, kTE, kMinB3>(h, b, a);
      else run_vec_modes<3, NNPE, 3, 0, PhysJ2<3>, kTE, kMinB3>(h, b, a);
      break;
    default: throw Error("fecb200: unsupported physics for TET4/TET10");
  }
}
template <int NNPE, int EPB>
static void mat3d(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  switch (b.physics) {
    case FECB200_PHYS_POISSON:
      FEC_REQUIRE(h->nf == 1, "Poisson needs NF = 1");
      run_mat<3, NNPE, 1, 0, PhysPoisson<3>, 32>(h, b, a);
      break;
    case FECB200_PHYS_LINEAR_ELASTIC:
      FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND");
      run_mat<3, NNPE, 3, 0, PhysLinearElastic<3>, EPB>(h, b, a);
      break;
    case FECB200_PHYS_NEOHOOKEAN:
      FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND");
      run_mat<3, NNPE, 3, 0, PhysNeoHookean<3>, EPB>(h, b, a);
      break;
    case FECB200_PHYS_J2_PLASTICITY:
      FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND");
      run_mat<3, NNPE, 3, 0, PhysJ2<3>, EPB>(h, b, a);
      break;
    default: throw Error("fecb200: unsupported physics for TET4/TET10");
  }
}

template <int NNPE>
static void scalar3d(fecb200_handle* h, BlockPlan& b, const double* U) {
  switch (b.physics) {
    case FECB200_PHYS_POISSON: FEC_REQUIRE(h->nf == 1, "Poisson needs NF = 1"); run_energy<3, NNPE, 1, 0, PhysPoisson<3>>(h, b, U); break;
    case FECB200_PHYS_LINEAR_ELASTIC: FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND"); run_energy<3, NNPE, 3, 0, PhysLinearElastic<3>>(h, b, U); break;
    case FECB200_PHYS_NEOHOOKEAN: FEC_REQUIRE(h->nf == 3, "mechanics needs NF = ND"); run_energy<3, NNPE, 3, 0, PhysNeoHookean<3>>(h, b, U); break;
    case FECB200_PHYS_J2_PLASTICITY: run_energy<3, NNPE, 3, 0, PhysJ2<3>>(h, b, U); break;
    default: throw Error("fecb200: unsupported physics for TET4/TET10");
  }
}
void launch_scalar_tet(fecb200_handle* h, BlockPlan& b, const double* U) {
  if (b.elem_type == FECB200_TET4) scalar3d<4>(h, b, U); else scalar3d<10>(h, b, U);
}

void launch_vector_tet(fecb200_handle* h, BlockPlan& b, const VecLaunch& a) {
  if (b.elem_type == FECB200_TET4) vec3d<4>(h, b, a); else vec3d<10>(h, b, a);
}
void launch_matrix_tet(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  if (b.elem_type == FECB200_TET4) mat3d<4, 32>(h, b, a); else mat3d<10, 12>(h, b, a);
}

}  // namespace fec
