"""hopper_mpc_inertial_b200 -- B200-native batched hopper MPC hot path.

Drop-in surface of bbokser/hopper-mpc-inertial's cvxpy/OSQP path:
  mpc_cvx_euler_3f.Mpc / mpc_cvx_euler_2f.Mpc  (same constructor and mpcontrol signature)
  robotrunner.Runner, robotrunner.convert, run (CLI)
plus the batched interface the GPU exists for: batch.BatchMpc, scenarios.make_batch.
The compute path is libhmpc_b200.so (hand-written sm_100a CUDA behind the C ABI of include/hmpc.h);
there is no CPU fallback.
"""
__all__ = ["batch", "planner", "scenarios", "mpc_cvx_euler_2f", "mpc_cvx_euler_3f", "robotrunner", "utils"]
