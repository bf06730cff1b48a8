"""TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference CUDA ops into oracle/_ref/.

The reference's `point_cloud_lib_ops` extension (10 translation units, see
/root/reference/point_cloud_lib/setup.py:14-28) is compiled here from the sources where
they lie under /root/reference; nothing is copied into this repository and the reference's
own build system (setup.py) is not run.  Output goes to oracle/_ref/ only (git-ignored, but
it travels to the GPU box with the gpurun snapshot), where `tests/` use it as the GPU-side
oracle for ball_query / knn_query / compute_keys / feat_basis_proj[_grad].

Usage:  python oracle/build_ref.py            (about 12 min on 8 cores; cached afterwards)
"""
import os
import sys
import glob
import shutil
import subprocess
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("SE3_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "point_cloud_lib", "custom_ops")

SOURCES = [
    "feature_aggregation/feat_basis_proj.cu",
    "feature_aggregation/feat_basis_proj_grads.cu",
    "ball_query/ball_query.cu",
    "ball_query/compute_keys.cu",
    "ball_query/build_grid_ds.cu",
    "ball_query/count_neighbors.cu",
    "ball_query/store_neighbors.cu",
    "ball_query/find_ranges_grid_ds.cu",
    "knn_query/knn_query.cu",
    "ops_list.cpp",
]


def ref_so_path():
    c = glob.glob(os.path.join(OUT, "point_cloud_lib_ops*.so"))
    return c[0] if c else None


def build(force=False, jobs=None):
    if ref_so_path() and not force:
        return ref_so_path()
    if not os.path.isdir(SRC):
        return None
    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce

    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-I", p]
    inc += ["-I", sysconfig.get_paths()["include"]]
    common = [
        "-DTORCH_EXTENSION_NAME=point_cloud_lib_ops",
        "-DTORCH_API_INCLUDE_EXTENSION_H",
        "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
        "-std=c++17", "-O3",
    ]
    procs = []
    objs = []
    jobs = jobs or os.cpu_count() or 4
    for s in SOURCES:
        o = os.path.join(OUT, "obj", s.replace("/", "_") + ".o")
        objs.append(o)
        if os.path.exists(o) and not force:
            continue
        src = os.path.join(SRC, s)
        if s.endswith(".cu"):
            cmd = ["nvcc", "-c", src, "-o", o, "-gencode", "arch=compute_100,code=sm_100",
                   "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
                   "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                   "-D__CUDA_NO_HALF2_OPERATORS__"] + common + inc
        else:
            cmd = ["g++", "-c", src, "-o", o, "-fPIC"] + common + inc
        procs.append((s, subprocess.Popen(cmd)))
        while len([p for _, p in procs if p.poll() is None]) >= jobs:
            for _, p in procs:
                if p.poll() is None:
                    p.wait()
                    break
    for s, p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference unit failed to compile: " + s)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(OUT, "point_cloud_lib_ops" + ext)
    libdir = ce.library_paths("cuda")
    link = ["g++", "-shared", "-o", so] + objs
    for d in libdir:
        link += ["-L", d, "-Wl,-rpath," + d]
    link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    subprocess.check_call(link)
    shutil.rmtree(os.path.join(OUT, "obj"), ignore_errors=True)
    return so


if __name__ == "__main__":
    so = build(force="--force" in sys.argv)
    print("reference ops:", so)
