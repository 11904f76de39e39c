#!/usr/bin/env python
"""Compile the UNMODIFIED reference extensions into oracle/_ref/ (test infrastructure only).

This is the recipe the task calls `oracle/_ref`: the three native extensions of the
reference are compiled from the sources where they lie under /root/reference (nothing is
copied into this repository), with plain nvcc / g++ command lines written here (the
reference's own setup.py files are not executed).  Outputs go only to oracle/_ref/:

    oracle/_ref/gbref_pointnet2_ext.so     <- PointNet/_ext_src/src/*.cpp,*.cu      (module "A", pointnet2._ext)
    oracle/_ref/gbref_pointnet2_batch.so   <- pointnet2_batch/src/*.cpp,*.cu        (module "B", pointnet2_batch_cuda)
    oracle/_ref/gbref_knn.so               <- KNN/Pytorch_CUDA_KNN/**               (module "C", KNN._C, WITH_CUDA)

Flags mirror what the reference's setup.py files ask torch's CUDAExtension for
(PointNet/setup.py:19-22 `-O2`; pointnet2_batch/setup.py:18-19 `nvcc -O2`;
KNN/setup.py:30-39 WITH_CUDA + half-precision defines) plus the sm_100a gencode.

The built modules are the *real reference kernels*; tests/ load them on the GPU box to pin
the C oracle (oracle/gb_oracle.c) and to check the product kernels bit for bit.  The product
(graspbalance_b200/) never imports anything from here.

/root/reference only exists in the build container: on the GPU box the prebuilt .so files
that travelled with the snapshot are used as they are.
"""
import glob
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("GB_REFERENCE_ROOT", "/root/reference")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

EXTS = {
    "gbref_pointnet2_ext": dict(
        srcs=sorted(glob.glob(f"{REF}/PointNet/_ext_src/src/*.cpp") + glob.glob(f"{REF}/PointNet/_ext_src/src/*.cu")),
        inc=[f"{REF}/PointNet/_ext_src/include"], defs=[], nvcc=["-O2"], cxx=["-O2"]),
    "gbref_pointnet2_batch": dict(
        srcs=sorted(glob.glob(f"{REF}/pointnet2_batch/src/*.cpp") + glob.glob(f"{REF}/pointnet2_batch/src/*.cu")),
        inc=[], defs=[], nvcc=["-O2"], cxx=["-g"]),
    "gbref_knn": dict(
        srcs=sorted(glob.glob(f"{REF}/KNN/Pytorch_CUDA_KNN/*.cpp") + glob.glob(f"{REF}/KNN/Pytorch_CUDA_KNN/cpu/*.cpp")
                    + glob.glob(f"{REF}/KNN/Pytorch_CUDA_KNN/cuda/*.cu")),
        inc=[f"{REF}/KNN/Pytorch_CUDA_KNN"], defs=["-DWITH_CUDA"],
        nvcc=["-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
              "-D__CUDA_NO_HALF2_OPERATORS__"], cxx=[]),
}


def _torch_paths():
    import torch
    from torch.utils import cpp_extension as ce
    tdir = os.path.dirname(torch.__file__)
    incs = ce.include_paths() + ["/usr/local/cuda/include", sysconfig.get_paths()["include"]]
    return tdir, incs


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout[-4000:]))
    return r.stdout


def build(force=False, verbose=True):
    """Build the three reference modules.  Returns {name: path}.  No-op when up to date."""
    if not os.path.isdir(REF):
        # GPU box: use whatever travelled.
        return {n: os.path.join(OUT, n + ".so") for n in EXTS if os.path.exists(os.path.join(OUT, n + ".so"))}
    os.makedirs(OUT, exist_ok=True)
    tdir, incs = _torch_paths()
    common = ["-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1", "-std=c++17"]
    inc_flags = [f"-I{p}" for p in incs]
    built = {}
    jobs = []
    for name, e in EXTS.items():
        so = os.path.join(OUT, name + ".so")
        built[name] = so
        if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in e["srcs"]):
            continue
        objs = []
        for s in e["srcs"]:
            o = os.path.join(OUT, name + "__" + os.path.basename(s).replace(".", "_") + ".o")
            objs.append(o)
            loc_inc = [f"-I{p}" for p in e["inc"]]
            defs = common + e["defs"] + [f"-DTORCH_EXTENSION_NAME={name}"]
            if s.endswith(".cu"):
                cmd = [NVCC, "-c", s, "-o", o, "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                       "-Xcompiler", "-fPIC", "-w"] + e["nvcc"] + defs + loc_inc + inc_flags
            else:
                cmd = ["g++", "-c", s, "-o", o, "-fPIC", "-w"] + e["cxx"] + defs + loc_inc + inc_flags
            jobs.append(cmd)
        e["_objs"] = objs
    if jobs:
        if verbose:
            print(f"[build_ref] compiling {len(jobs)} reference translation units ...", file=sys.stderr)
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
            list(ex.map(_run, jobs))
    for name, e in EXTS.items():
        if "_objs" not in e:
            continue
        so = built[name]
        cmd = ["g++", "-shared", "-o", so] + e["_objs"] + [
            f"-L{tdir}/lib", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
            "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{tdir}/lib"]
        _run(cmd)
        for o in e["_objs"]:
            os.remove(o)
        if verbose:
            print(f"[build_ref] built {so}", file=sys.stderr)
    return built


if __name__ == "__main__":
    build(force="--force" in sys.argv)
