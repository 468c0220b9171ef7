"""CLI drop-in for the reference's run.py (run.py:7-27): ``python -m hopper_mpc_inertial_b200.run 3f --curve``.

Same positional ``dyn`` and ``--curve`` / ``--N_run`` flags; ``--runtime`` is accepted as an alias of
``--N_run`` because the reference README uses that spelling (README.md:58).
Extra flags select the GPU and the horizon; plotting is opt-in (headless by default)."""
import argparse


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("dyn", help="choose 2f or 3f", choices=['2f', '3f'], type=str)
    parser.add_argument("--curve", help="make the ref traj curved", action="store_true")
    parser.add_argument("--N_run", "--runtime", dest="N_run", help="sim run time in ms (integer)",
                        type=int, default=5000)
    parser.add_argument("--horizon", type=int, default=60, help="MPC horizon (reference: 60)")
    parser.add_argument("--device", type=int, default=0)
    parser.add_argument("--plot", action="store_true", help="save plots (needs matplotlib)")
    parser.add_argument("--fused", action="store_true", help="use the fused rollout path")
    parser.add_argument("--gate", choices=["off", "schedule", "detect"], default="off",
                        help="contact gate of the applied control (reference robotrunner.py:111 has `* s` commented out = off)")
    parser.add_argument("--leg_max", type=float, default=None, help="leg reach for --gate detect [m]")
    args = parser.parse_args(argv)

    from .robotrunner import Runner
    dt = 1e-3
    runner = Runner(dt=dt, dyn=args.dyn, curve=bool(args.curve), N_run=args.N_run, N=args.horizon,
                    device=args.device, contact_gate=args.gate, leg_max=args.leg_max)
    if args.fused:
        X_log, U_log = runner.run_fused()
        print("final state:", X_log[-1])
    else:
        runner.run(plot=args.plot)
        print("final state:", runner.X_traj[-1])
    return runner


if __name__ == "__main__":
    main()
