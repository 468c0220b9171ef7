"""Top-level CLI shim so that the reference's README command works as written (run.py:1-27, README.md:58):

    python run.py 3f --curve
    python run.py 2f --N_run 2000

Everything lives in hopper_mpc_inertial_b200/run.py; this file only forwards the arguments."""
from hopper_mpc_inertial_b200.run import main

if __name__ == "__main__":
    main()
