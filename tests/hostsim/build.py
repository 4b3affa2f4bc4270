"""TEST INFRASTRUCTURE: builds tests/hostsim/libpb_hostsim.so = the unmodified host engine
(diffusion_pullback_b200/csrc/pb_engine.cpp) + the plain-C++ leaf-kernel double (pbk_hostsim.cpp).
Used only by `pytest -m "not gpu"` to exercise the engine's host logic without a GPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "diffusion_pullback_b200", "csrc")
LIB = os.path.join(HERE, "libpb_hostsim.so")


def build(force=False):
    srcs = [os.path.join(CSRC, "pb_engine.cpp"), os.path.join(HERE, "pbk_hostsim.cpp")]
    deps = srcs + [os.path.join(ROOT, "include", f) for f in ("pb_kernels.h", "pb_gemm.h", "pullback_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["g++", "-O3", "-march=native", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
           "-I", CSRC, "-I", os.path.join(ROOT, "include"), "-o", LIB] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("hostsim build failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
