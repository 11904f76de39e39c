"""ctypes binding of libgbops.so (the C ABI declared in include/gbops.h).

The library is built in-tree by `make -C graspbalance_b200/csrc` (see __graft_entry__.build).  There is NO fallback:
if the shared object is missing or a CUDA call fails, the operators raise.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GBOPS_LIB") or os.path.join(_HERE, "libgbops.so")  # GBOPS_LIB: experiments with an alternative build
CSRC = os.path.join(_HERE, "csrc")
ABI_VERSION = 1

_lib = None

_vp, _i, _f, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong

# name -> argtypes (restype is int unless listed in _RESTYPES); must match include/gbops.h
SIGNATURES = {
    "gb_abi_version": [],
    "gb_error_string": [_i],
    "gb_fps": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gb_fps_xyz": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gb_fps_xyz_hint": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "gb_fps_segments": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gb_gather_fwd": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gb_gather_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gb_ball_query": [_vp, _vp, _vp, _i, _i, _i, _f, _i, _vp],
    "gb_cylinder_query": [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _i, _vp],
    "gb_cylinder_query_multi": [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, ctypes.POINTER(ctypes.c_float), _i, _i, _vp],
    "gb_cylinder_query_multi_radius": [_vp, _vp, _vp, _vp, _i, _i, _i, ctypes.POINTER(ctypes.c_float), _i, _f,
                                       ctypes.POINTER(ctypes.c_float), _i, _i, _vp],
    "gb_group_fwd": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "gb_group_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "gb_group_bwd_set": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "gb_group_fwd_strided": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _vp],
    "gb_group_bwd_strided": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _i, _vp],
    "gb_group_xyz": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _ll, _vp],
    "gb_group_xyz_feat": [_vp, _vp, _vp, _vp, _ll, _f, _i, _vp, _vp, _ll, _i, _i, _i, _i, _i, _vp],
    "gb_three_nn": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "gb_three_nn_weights": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "gb_three_interp_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gb_three_interp_bwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gb_three_interp_bwd_set": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gb_knn": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "gb_group_max_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "gb_group_max_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gb_three_interpolation": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "gb_collision_counts": [_vp, _i, _vp, _vp, _vp, _i, _vp, _vp],
    "gb_collision_counts_host": [_vp, _i, _vp, _vp, _vp, _i, _vp],
    "gb_collision_counts_batched": [_vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp],
    "gb_collision_detect": [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, ctypes.POINTER(ctypes.c_double), _vp, _vp, _vp, _vp, _vp],
    "gb_voxel_means": [_vp, _vp, _vp, _vp, _i, _vp],
    "gb_set_tuning": [ctypes.c_char_p, _i],
    "gb_get_tuning": [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)],
    "gb_launch_count": [],
}
_RESTYPES = {"gb_error_string": ctypes.c_char_p, "gb_launch_count": ctypes.c_uint64}


def build(force=False, verbose=False):
    """Compile libgbops.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "gbops.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    r = subprocess.run(["make", "-C", CSRC, "-j8"] + (["-B"] if force else []), stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("building libgbops.so failed (see output above)")
    return LIB_PATH


def lib():
    """The loaded library.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(graspbalance_b200 has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        if L.gb_abi_version() != ABI_VERSION:
            raise RuntimeError("libgbops.so ABI version mismatch: rebuild it")
        _lib = L
    return _lib


def check(err, what):
    if err != 0:
        msg = lib().gb_error_string(err)
        raise RuntimeError(f"{what}: CUDA error {err} ({msg.decode() if msg else '?'})")


# ---- optional per-launch profiler (bench.py): name -> list of (start_event, end_event, algorithmic_bytes) ----
PROFILER = None


# ALGORITHMIC bytes of one launch from its integer arguments (SURVEY.md 8d / DESIGN.md "Roofline accounting"):
# the bytes the op must read and write at minimum, whatever the implementation does.
ALGO_BYTES = {
    "gb_fps": lambda a: a[3] * (12 * a[4] + 4 * a[5]),                                   # b*(12n + 4m)
    "gb_fps_xyz": lambda a: a[4] * (12 * a[5] + 16 * a[6]),                               # b*(12n + 4m + 12m)
    "gb_fps_xyz_hint": lambda a: a[4] * (12 * a[5] + 16 * a[6]),
    "gb_fps_segments": lambda a: a[4] * (12 * a[5] + 4 * a[6]),                           # <= nseg*(12 max_n + 4 max_m)
    "gb_gather_fwd": lambda a: a[3] * (4 * a[4] * a[5] + 4 * a[6] + 4 * a[4] * a[6]),     # b*(4cn + 4m + 4cm)
    "gb_gather_bwd": lambda a: a[3] * (4 * a[4] * a[5] + 4 * a[6] + 4 * a[4] * a[6]),
    "gb_ball_query": lambda a: a[3] * (12 * a[4] + 12 * a[5] + 4 * a[5] * a[7]),          # b*(12n + 12m + 4 m ns)
    "gb_cylinder_query": lambda a: a[4] * (12 * a[5] + 48 * a[6] + 4 * a[6] * a[10]),     # b*(12n + 48m + 4 m ns)
    "gb_cylinder_query_multi": lambda a: a[4] * (12 * a[5] + 48 * a[6] + 4 * a[6] * a[10] * a[11]),  # b*(12n + 48m + 4 m nd ns)
    "gb_cylinder_query_multi_radius": lambda a: a[4] * (12 * a[5] + 48 * a[6] + 4 * a[6] * a[8] * a[11] * a[12]),  # b*(12n+48m+4 m nr nd ns)
    "gb_group_fwd": lambda a: a[3] * (4 * a[4] * a[5] + 4 * a[6] * a[7] + 4 * a[4] * a[6] * a[7]),
    "gb_group_bwd": lambda a: a[3] * (4 * a[4] * a[5] + 4 * a[6] * a[7] + 4 * a[4] * a[6] * a[7]),
    "gb_three_nn": lambda a: a[4] * (12 * a[5] + 12 * a[6] + 24 * a[5]),                  # b*(12n + 12m + 24n)
    "gb_three_nn_weights": lambda a: a[5] * (12 * a[6] + 12 * a[7] + 36 * a[6]),         # b*(12n + 12m + 36n)
    "gb_three_interp_fwd": lambda a: a[4] * (4 * a[5] * a[6] + 24 * a[7] + 4 * a[5] * a[7]),   # b*(4cm + 24n + 4cn)
    "gb_three_interp_bwd": lambda a: a[4] * (4 * a[5] * a[7] + 24 * a[6] + 4 * a[5] * a[6]),   # args (b,c,n,m)
    "gb_group_bwd_set": lambda a: a[3] * (4 * a[4] * a[5] + 4 * a[6] * a[7] + 4 * a[4] * a[6] * a[7]),
    "gb_three_interp_bwd_set": lambda a: a[4] * (4 * a[5] * a[7] + 24 * a[6] + 4 * a[5] * a[6]),
    "gb_group_fwd_strided": lambda a: a[3] * (4 * a[4] * a[5] + 4 * a[6] * a[7] + 4 * a[4] * a[6] * a[7]),
    "gb_group_bwd_strided": lambda a: a[3] * (4 * a[4] * a[5] + 4 * a[6] * a[7] + 4 * a[4] * a[6] * a[7]),
    "gb_group_xyz": lambda a: a[5] * (12 * a[6] + 12 * a[7] + (36 * a[7] if a[3] else 0) + 16 * a[7] * a[8]),  # b*(12n+12m(+36m)+(4+12) m ns)
    # b*(4cn + 4 m ns + 4c m ns) for the features + b*(12n + 12m + 12 m ns) for the coordinates (idx read once)
    "gb_group_xyz_feat": lambda a: a[10] * (4 * a[11] * a[12] + 4 * a[13] * a[14] + 4 * a[11] * a[13] * a[14]
                                            + 12 * a[12] + 12 * a[13] + 12 * a[13] * a[14]),
    "gb_knn": lambda a: a[3] * (4 * a[4] * (a[5] + a[6]) + 8 * a[7] * a[6]),              # b*(4d(R+Q) + 8kQ)
    "gb_group_max_fwd": lambda a: a[4] * (4 * a[5] * a[6] + 4 * a[7] * a[8] + (8 if a[3] else 4) * a[5] * a[7]),
    "gb_group_max_bwd": lambda a: a[3] * (8 * a[4] * a[6] + 4 * a[4] * a[5]),
    "gb_three_interpolation": lambda a: a[6] * (12 * a[8] + 12 * a[9] + 4 * a[7] * a[9] + 4 * a[7] * a[8] + (24 * a[8] if a[4] else 0)),
    "gb_collision_counts": lambda a: 24 * a[1] + 176 * a[5] + 48 * a[5],
    "gb_collision_counts_batched": lambda a: a[2] * (24 * a[3] + 176 * a[7] + 48 * a[7]),
    "gb_collision_detect": lambda a: 24 * a[1] + 8 * 15 * a[4] + a[4],                   # 24 N' + grasp rows + masks
    "gb_voxel_means": lambda a: 48 * a[4] * 4 + 24 * a[4],  # ~4 points per voxel read (24 B + 8 B index), 24 B written
}


def call(name, ref_tensor, *args):
    """Launch entry point `name` on torch's current stream of ref_tensor's device (making that device current for the
    launch), raise on error, and -- when PROFILER is a dict -- bracket the launch with CUDA events on that stream."""
    import torch
    dev = ref_tensor.device
    prev = None
    if dev.index is not None and dev.index != torch.cuda.current_device():
        prev = torch.cuda.current_device()
        torch.cuda.set_device(dev.index)
    try:
        stream = torch.cuda.current_stream(dev)
        fn = getattr(lib(), name)
        if PROFILER is None:
            err = fn(*args, stream.cuda_stream)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            err = fn(*args, stream.cuda_stream)
            e1.record(stream)
            PROFILER.setdefault(name, []).append((e0, e1, ALGO_BYTES[name](args), tuple(a if isinstance(a, (int, float)) else None for a in args)))
    finally:
        if prev is not None:
            torch.cuda.set_device(prev)
    check(err, name)


def set_tuning(key, value):
    check(lib().gb_set_tuning(key.encode(), int(value)), f"gb_set_tuning({key})")


def get_tuning(key):
    v = ctypes.c_int(0)
    check(lib().gb_get_tuning(key.encode(), ctypes.byref(v)), f"gb_get_tuning({key})")
    return v.value


def launch_count():
    return int(lib().gb_launch_count())
