// dispatch_quad_tri.cu -- QUAD4 / TRI3 instantiations (2-D; the reference's test fixtures:
// test/poisson/poisson.g, multi_block_mesh_quad4_tri3.g, test/mechanics/mechanics_coarse.g).
#include "kernels.cuh"

namespace fec {

template <int NNPE>
static void vec2d(fecb200_handle* h, BlockPlan& b, const VecLaunch& a) {
  switch (b.physics) {
    case FECB200_PHYS_POISSON:
      FEC_REQUIRE(h->nf == 1, "Poisson needs NF = 1");
      run_vec_modes<2, NNPE, 1, 0, PhysPoisson<2>, kTE, kMinB1>(h, b, a);
      break;
    case FECB200_PHYS_LINEAR_ELASTIC:  // PlaneStrain (src/Formulations.jl:318-447)
      FEC_REQUIRE(h->nf == 2, "plane-strain mechanics needs NF = 2");
      run_vec_modes<2, NNPE, 2, 0, PhysLinearElastic<2>, kTE, kMinB1>(h, b, a);
      break;
    case FECB200_PHYS_NEOHOOKEAN:
      FEC_REQUIRE(h->nf == 2, "plane-strain mechanics needs NF = 2");
      run_vec_modes<2, NNPE, 2, 0, PhysNeoHookean<2>, kTE, kMinB1>(h, b, a);
      break;
    case FECB200_PHYS_TEST_NONSYMMETRIC:
      FEC_REQUIRE(h->nf == 2, "plane-strain mechanics needs NF = 2");
      run_vec_modes<2, NNPE, 2, 0, PhysNonSymmetricTest<2>, kTE, kMinB1>(h, b, a);
      break;
    default: throw Error("fecb200: unsupported physics for QUAD4/TRI3");
  }
}
template <int NNPE>
static void mat2d(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  switch (b.physics) {
    case FECB200_PHYS_POISSON:
      FEC_REQUIRE(h->nf == 1, "Poisson needs NF = 1");
      run_mat<2, NNPE, 1, 0, PhysPoisson<2>, 32>(h, b, a);
      break;
    case FECB200_PHYS_LINEAR_ELASTIC:
      FEC_REQUIRE(h->nf == 2, "plane-strain mechanics needs NF = 2");
      run_mat<2, NNPE, 2, 0, PhysLinearElastic<2>, 32>(h, b, a);
      break;
    case FECB200_PHYS_NEOHOOKEAN:
      FEC_REQUIRE(h->nf == 2, "plane-strain mechanics needs NF = 2");
      run_mat<2, NNPE, 2, 0, PhysNeoHookean<2>, 32>(h, b, a);
      break;
    case FECB200_PHYS_TEST_NONSYMMETRIC:
      FEC_REQUIRE(h->nf == 2, "plane-strain mechanics needs NF = 2");
      run_mat<2, NNPE, 2, 0, PhysNonSymmetricTest<2>, 32>(h, b, a);
      break;
    default: throw Error("fecb200: unsupported physics for QUAD4/TRI3");
  }
}

template <int NNPE>
static void scalar2d(fecb200_handle* h, BlockPlan& b, const double* U) {
  switch (b.physics) {
    case FECB200_PHYS_POISSON: FEC_REQUIRE(h->nf == 1, "Poisson needs NF = 1"); run_energy<2, NNPE, 1, 0, PhysPoisson<2>>(h, b, U); break;
    case FECB200_PHYS_LINEAR_ELASTIC: FEC_REQUIRE(h->nf == 2, "plane-strain mechanics needs NF = 2"); run_energy<2, NNPE, 2, 0, PhysLinearElastic<2>>(h, b, U); break;
    case FECB200_PHYS_NEOHOOKEAN: FEC_REQUIRE(h->nf == 2, "plane-strain mechanics needs NF = 2"); run_energy<2, NNPE, 2, 0, PhysNeoHookean<2>>(h, b, U); break;
    default: throw Error("fecb200: unsupported physics for QUAD4/TRI3");
  }
}
void launch_scalar_quad_tri(fecb200_handle* h, BlockPlan& b, const double* U) {
  if (b.elem_type == FECB200_QUAD4) scalar2d<4>(h, b, U); else scalar2d<3>(h, b, U);
}

void launch_vector_quad_tri(fecb200_handle* h, BlockPlan& b, const VecLaunch& a) {
  if (b.elem_type == FECB200_QUAD4) vec2d<4>(h, b, a); else vec2d<3>(h, b, a);
}
void launch_matrix_quad_tri(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  if (b.elem_type == FECB200_QUAD4) mat2d<4>(h, b, a); else mat2d<3>(h, b, a);
}

}  // namespace fec
