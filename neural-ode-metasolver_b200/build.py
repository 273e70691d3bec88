"""Build libmetasolver_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python neural-ode-metasolver_b200/build.py [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libmetasolver_b200.so")
SOURCES = ["odeblock.cu", "blocks.cu", "netlayers.cu", "elementwise.cu", "conv_simt.cu", "conv_tc.cu", "conv_tcp.cu", "conv_tcp2.cu", "conv_tct.cu", "wgrad_tc.cu",
           "groupnorm.cu", "train_aux.cu", "head.cu", "mnist_fused.cu", "peer.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
              "-I", os.path.join(ROOT, "include"), "-I", CSRC] + os.environ.get("MSB_NVCC_EXTRA", "").split()


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False, debug=False):
    """debug=True: the instrumented variant (-DMSB_CONV_DEBUG: bottleneck-decomposition switches, bounded waits that
    report instead of trapping) as libmetasolver_b200_dbg.so; select it with MSB_LIB_PATH."""
    if debug:
        return _build(LIB.replace(".so", "_dbg.so"), NVCC_FLAGS + ["-DMSB_CONV_DEBUG"], "_dbg.o", force, verbose)
    return _build(LIB, NVCC_FLAGS, ".o", force, verbose)


def _build(LIB, NVCC_FLAGS, osuffix, force, verbose):
    stamp = LIB + ".stamp"
    dig = _digest() + "".join(NVCC_FLAGS[-1:])
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace(".cu", osuffix))
        objs.append(o)
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (s, out))
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (s, out))
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + []
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    open(os.path.join(CSRC, "build.log"), "w").write("\n".join(log))
    open(stamp, "w").write(dig)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
