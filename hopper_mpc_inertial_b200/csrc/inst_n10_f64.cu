// inst_n10_f64.cu -- one instantiation of the solver kernel (see hmpc_kernel.cuh), its own translation unit.
#include "hmpc_kernel.cuh"

namespace hmpc {
cudaError_t mpc_set_smem_n10_f64(int bytes) { return mpc_set_smem<128, 4, true, double, false>(bytes); }
void mpc_launch_n10_f64(const MpcLaunch& l, const QpConst& qc, const MpcIo& io) { mpc_launch<128, 4, true, double, false>(l, qc, io); }
}  // namespace hmpc

#ifdef HMPC_PHASE_TIMING
// debug hook (not part of include/hmpc.h): cycles per phase summed over the CTAs' thread 0, then reset
extern "C" int hmpc_debug_phases(unsigned long long* out) {
    cudaError_t e = cudaMemcpyFromSymbol(out, hmpc::g_phase, sizeof(unsigned long long) * 16);
    if (e != cudaSuccess) return -1;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(hmpc::g_phase, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif
