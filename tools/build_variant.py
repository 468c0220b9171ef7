"""Build an experiment variant of the library:  python tools/build_variant.py NAME -DHMPC_PHASE_TIMING ...
-> hopper_mpc_inertial_b200/variants/libhmpc_b200_NAME.so (git-ignored; use with HMPC_LIB_PATH=...)."""
import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hopper_mpc_inertial_b200 import build as b   # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
vdir = os.path.join(ROOT, "hopper_mpc_inertial_b200", "variants")
odir = os.path.join(ROOT, "hopper_mpc_inertial_b200", "build", "var_" + name)
os.makedirs(vdir, exist_ok=True)
os.makedirs(odir, exist_ok=True)
nvcc = b._nvcc()


def cu(u):
    obj = os.path.join(odir, u.replace(".cu", ".o"))
    subprocess.run([nvcc] + b.NVCC_FLAGS + flags + ["-c", os.path.join(b._CSRC, u), "-o", obj], check=True)
    return obj


with cf.ThreadPoolExecutor(max_workers=len(b.UNITS)) as ex:
    objs = list(ex.map(cu, b.UNITS))
out = os.path.join(vdir, f"libhmpc_b200_{name}.so")
subprocess.run([nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", out] + objs, check=True)
print(out)
