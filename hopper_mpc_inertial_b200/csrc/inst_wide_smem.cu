// inst_wide_smem.cu -- one instantiation of the solver kernel (see hmpc_kernel.cuh), its own translation unit.
#include "hmpc_kernel.cuh"

namespace hmpc {
cudaError_t mpc_set_smem_wide_smem(int bytes) { return mpc_set_smem<256, 1, true, double, true>(bytes); }
void mpc_launch_wide_smem(const MpcLaunch& l, const QpConst& qc, const MpcIo& io) { mpc_launch<256, 1, true, double, true>(l, qc, io); }
}  // namespace hmpc
