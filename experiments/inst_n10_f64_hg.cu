// inst_n10_f64_hg.cu -- EXPERIMENT (HMPC_H_GLOBAL=1): Hessian in the per-CTA L2 workspace, factor in shared
// memory, five CTAs per SM.  Thread count of the variant: HMPC_HG_THREADS at compile time (default 128).
#include "hmpc_kernel.cuh"
#ifndef HMPC_HG_THREADS
#define HMPC_HG_THREADS 128
#endif
namespace hmpc {
cudaError_t mpc_set_smem_n10_f64_hg(int bytes) { return mpc_set_smem<HMPC_HG_THREADS, 5, true, double, false, true>(bytes); }
void mpc_launch_n10_f64_hg(const MpcLaunch& l, const QpConst& qc, const MpcIo& io) { mpc_launch<HMPC_HG_THREADS, 5, true, double, false, true>(l, qc, io); }
}  // namespace hmpc
