"""Builds amc-slam_b200/libgpba.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgpba.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
         "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++", "-Xptxas", "-v"]


def sources():
    return [os.path.join(SRC, f) for f in sorted(os.listdir(SRC))] + [os.path.join(os.path.dirname(HERE), "include", h) for h in ("gpba.h", "gpba_map.h")]


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in sources()):
        return OUT
    flags = [f for f in FLAGS if f != "--use_fast_math=false"]
    cmd = [NVCC] + flags + ["-o", OUT, os.path.join(SRC, "gpba_host.cu"), os.path.join(SRC, "gpba_map.cc"), "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print("built", OUT)
