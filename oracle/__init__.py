"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (never imported by hopper_mpc_inertial_b200)."""
