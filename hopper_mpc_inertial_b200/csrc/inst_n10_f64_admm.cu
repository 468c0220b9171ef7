// inst_n10_f64_admm.cu -- one instantiation of the solver kernel (see hmpc_kernel.cuh), its own translation unit.
#include "hmpc_kernel.cuh"

namespace hmpc {
cudaError_t mpc_set_smem_n10_f64_admm(int bytes) { return mpc_set_smem<128, 4, true, double, true>(bytes); }
void mpc_launch_n10_f64_admm(const MpcLaunch& l, const QpConst& qc, const MpcIo& io) { mpc_launch<128, 4, true, double, true>(l, qc, io); }
}  // namespace hmpc
