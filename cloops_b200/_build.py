"""Build libcloops_b200.so in-tree (nvcc, sm_100a).  No GPU needed: nvcc cross-compiles."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcloops_b200.so")


def build(force: bool = False, verbose: bool = False) -> str:
    csrc = os.path.join(HERE, "csrc")
    if force:
        subprocess.run(["make", "-C", csrc, "clean"], check=True, capture_output=not verbose)
    r = subprocess.run(["make", "-C", csrc, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libcloops_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
