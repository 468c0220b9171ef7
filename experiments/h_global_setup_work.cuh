// Companion of h_global_5cta.patch: the setup_work variant (hmpc_mpc.cuh) the experiment used.
template <bool SMEM_MATS, bool H_GLOBAL = false>
__device__ inline void setup_work(Work& w, const QpConst& c, double* smem, double* ws, int fsize = 8) {
    carve(w, smem, c.N);
    const size_t n = 6 * (size_t)c.N;
    if (SMEM_MATS && H_GLOBAL) {          // Hessian in the per-CTA global (L2-resident) workspace, factor in shared memory
        w.H = ws + (size_t)blockIdx.x * (n * (n + 1) / 2);
        w.Lm = smem + ((work_vec_doubles(c.N) + 1) & ~(size_t)1);
        return;
    }
    double* mat;
    if (SMEM_MATS) mat = smem + ((work_vec_doubles(c.N) + 1) & ~(size_t)1);
    else mat = ws + (size_t)blockIdx.x * mat_doubles(c.N, fsize);
    w.H = mat; w.Lm = mat + n * (n + 1) / 2;
}
