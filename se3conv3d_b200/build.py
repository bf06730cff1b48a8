"""Builds libse3conv3d_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

The library has no torch dependency: only the CUDA runtime (and CUB, header-only, for the device
radix sorts / scans).  `python -m se3conv3d_b200.build` or `__graft_entry__.build()` runs this.
"""
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libse3conv3d_b200.so")
SOURCES = ["grid_ops.cu", "knn_frames.cu", "conv_simt.cu", "conv_tc.cu", "conv_fused.cu", "proj_tcgen05.cu", "proj_tma.cu", "conv.cu", "hierarchy.cu", "block_ops.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/se3conv3d_b200.h"]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(name, defines):
    """Builds lib/libse3conv3d_b200.<name>.so with extra -D flags (kernel-tuning A/B runs; load it with
    SE3CONV3D_LIB=<path>)."""
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libse3conv3d_b200.%s.so" % name)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(LIBDIR, "%s.%s.o" % (s, name))
        objs.append(o)
        cmd = ["nvcc", "-c", os.path.join(CSRC, s), "-o", o] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + \
              ["-D" + d for d in defines]
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed")
    subprocess.check_call(["nvcc", "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    for o in objs:
        os.remove(o)
    return out


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.sha256")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    objs, procs = [], []
    for s in SOURCES:
        o = os.path.join(LIBDIR, s + ".o")
        objs.append(o)
        cmd = ["nvcc", "-c", os.path.join(CSRC, s), "-o", o] + NVCC_FLAGS
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        log.append("==== %s ====\n%s" % (s, out))
        if p.returncode != 0:
            failed = True
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see " + os.path.join(LIBDIR, "build.log"))
    subprocess.check_call(["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                  "-lcudart"])
    for o in objs:
        os.remove(o)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
