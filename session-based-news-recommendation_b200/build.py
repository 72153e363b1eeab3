"""Compile libtcar_b200.so for sm_100a with nvcc (cross-compiles on a GPU-less host)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libtcar_b200.so")


def build(force: bool = False, verbose: bool = False) -> str:
    csrc = os.path.join(HERE, "csrc")
    cmd = ["make", "-C", csrc, "-j4"] + (["-B"] if force else [])
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0 or not os.path.exists(LIB):
        raise RuntimeError("nvcc build of libtcar_b200.so failed:\n" + res.stdout)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
