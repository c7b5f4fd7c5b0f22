#include "../../pic1dp_b200/csrc/particle_kernels.cuh"
using namespace pic1dp;
template __global__ void pic1dp::k_push<3, false, DEP_SMEM_ATOMIC, true, 25>(const ParticleArgs);
template __global__ void pic1dp::k_push<3, true, DEP_SMEM_ATOMIC, true, 25>(const ParticleArgs);
