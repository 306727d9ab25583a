"""Builds libwildcat_b200.so (hand-written sm_100a CUDA + the extern "C" ABI of include/wildcat_b200.h) in-tree.

nvcc cross-compiles without a GPU.  wc_match.cu is compiled with -fmad=false: its discrete decisions (kNN order,
gates) must see the same un-contracted fp64 arithmetic as the reference's x86-64 build.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libwildcat_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
          *os.environ.get("WC_NVCC_EXTRA", "").split()]  # e.g. WC_NVCC_EXTRA=-DWC_LM_TIMING for the in-kernel phase clocks
SOURCES = {"wc_api.cu": [], "wc_extract.cu": [], "wc_match.cu": ["-fmad=false"], "wc_solve.cu": [], "wc_spline.cu": [],
           "wc_comm.cu": [], "wc_pass.cu": [], "wc_sweep.cu": ["-fmad=false"]}


def _newer(src, dst):
    return not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "wildcat_b200.h"))
    objs, procs = [], []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(s, o) or any(_newer(h, o) for h in headers):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:  # a runaway cicc/ptxas must not hang the caller
            p.kill()
            out, _ = p.communicate()
            out += f"\n{src}: nvcc timed out after 600 s\n".encode()
            p.returncode = 1
        if p.returncode != 0 or verbose:
            sys.stderr.write(out.decode())
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or not os.path.exists(OUT) or any(_newer(o, OUT) for o in objs):
        subprocess.check_call([nvcc, *ARCH, "-shared", "-o", OUT, *objs, "-lcudart"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
