"""Builds the test-only host emulator of the CUDA walkers (g++, no GPU needed)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "varanneal_b200", "csrc")
OUT = os.path.join(HERE, "libvab_emul.so")


def build(force=False):
    srcs = [os.path.join(HERE, "ode_emul.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if (not force and os.path.exists(OUT)
            and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps)):
        return OUT
    cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-x", "c++", "-I", CSRC] + srcs + ["-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
