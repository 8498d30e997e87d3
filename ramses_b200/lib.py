"""ctypes binding of libramses_gpu.so (the C-ABI declared in include/ramses_gpu.h)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RGPU_LIB", os.path.join(_HERE, "libramses_gpu.so"))   # RGPU_LIB: kernel-variant experiments
HEADER = os.path.join(os.path.dirname(_HERE), "include", "ramses_gpu.h")

RIEMANN = {"llf": 0, "exact": 1, "acoustic": 2, "hllc": 3, "hll": 4}
# MHD build: iriemann / iriemann2d (hydro/read_hydro_params.f90:190-220)
MHD_RIEMANN = {"llf": 0, "roe": 1, "hll": 2, "hlld": 3, "upwind": 4, "hydro": 5}
MHD_RIEMANN2D = {"llf": 0, "roe": 1, "upwind": 2, "hll": 3, "hlla": 4, "hlld": 5}


class RgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rgpu error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """struct rgpu_params (include/ramses_gpu.h)."""
    _fields_ = [("ndim", C.c_int), ("nvar", C.c_int), ("nvector", C.c_int), ("slope_type", C.c_int),
                ("niter_riemann", C.c_int), ("scheme", C.c_int), ("riemann", C.c_int), ("pressure_fix", C.c_int),
                ("gamma", C.c_double), ("smallr", C.c_double), ("smallc", C.c_double), ("slope_theta", C.c_double),
                ("difmag", C.c_double), ("courant_factor", C.c_double), ("boxlen", C.c_double),
                ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("icoarse_min", C.c_int), ("icoarse_max", C.c_int), ("jcoarse_min", C.c_int), ("jcoarse_max", C.c_int),
                ("kcoarse_min", C.c_int), ("kcoarse_max", C.c_int), ("nlevelmax", C.c_int),
                ("mhd", C.c_int), ("riemann2d", C.c_int), ("slope_mag_type", C.c_int), ("fast", C.c_int),
                ("poisson", C.c_int), ("beta_fix", C.c_double)]


class LevelInfo(C.Structure):
    _fields_ = [("dense", C.c_int), ("ncell_box", C.c_int * 3), ("own_lo", C.c_int * 3), ("own_hi", C.c_int * 3),
                ("wrap", C.c_int * 3), ("nslot", C.c_longlong), ("kernel_launches", C.c_longlong),
                ("last_sweep_ms", C.c_double), ("last_steps_ms", C.c_double), ("pipeline_slabs", C.c_int), ("sweep_variant", C.c_int)]


def build(verbose=False):
    """Compile every CUDA source for sm_100a into ramses_b200/libramses_gpu.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    r = subprocess.run(cmd, capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libramses_gpu.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


_lib = None


def _preload_bundled_nccl():
    """libramses_gpu.so needs `libnccl.so.2`.  When the Python environment ships its own NCCL (the `nvidia-nccl-cu12` wheel
    PyTorch links against), load THAT copy first: the dynamic loader then binds both this library and a later
    `import torch` to the same, newer NCCL.  Otherwise the system libnccl would be bound first and a subsequent
    `import torch` fails with an undefined NCCL symbol.  No torch import happens here."""
    try:
        import glob
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            for f in sorted(glob.glob(os.path.join(d, "lib", "libnccl.so*"))):
                C.CDLL(f, mode=C.RTLD_GLOBAL)
                return f
    except Exception:
        pass
    return None


def load():
    """Load the C-ABI library.  Fails loudly if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(ramses_b200 has no CPU fallback)")
    _preload_bundled_nccl()
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    ipp = C.POINTER(ip)
    L.rgpu_last_error.restype = C.c_char_p
    L.rgpu_init.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int]
    L.rgpu_bind_tree.argtypes = [C.c_int, C.c_int, ip, ip, ip]
    L.rgpu_bind_level.argtypes = [C.c_int, C.c_int, ip, C.c_int, ip, ipp, ip, ipp, C.c_int, ip, ip, ipp]
    L.rgpu_plan_level.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, ip, C.c_int, C.c_int, ip, C.c_int, ip, ipp, ip, ipp,
                                  C.c_int, ip, ip, ipp, C.POINTER(LevelInfo), ip, C.c_longlong]
    L.rgpu_godunov_fine.argtypes = [C.c_int, C.c_double, dp, dp]
    L.rgpu_host_register.argtypes = [C.c_void_p, C.c_size_t]
    L.rgpu_host_unregister.argtypes = [C.c_void_p]
    L.rgpu_upload_state.argtypes = [C.c_int, dp]
    L.rgpu_download_state.argtypes = [C.c_int, dp]
    L.rgpu_set_unew.argtypes = [C.c_int]
    L.rgpu_godunov_fine_dev.argtypes = [C.c_int, C.c_double]
    L.rgpu_set_uold.argtypes = [C.c_int]
    L.rgpu_courant_fine.argtypes = [C.c_int, dp, dp]
    L.rgpu_make_boundary_hydro.argtypes = [C.c_int]
    L.rgpu_make_virtual_fine.argtypes = [C.c_int]
    L.rgpu_make_virtual_reverse.argtypes = [C.c_int]
    L.rgpu_level_steps.argtypes = [C.c_int, C.c_int, dp, dp]
    L.rgpu_upload_fine.argtypes = [C.c_int]
    L.rgpu_amr_steps.argtypes = [C.c_int, ip, C.c_int, dp]
    L.rgpu_set_amr.argtypes = [C.c_int, C.c_int, C.c_int]
    L.rgpu_comm_unique_id.argtypes = [C.c_void_p]
    L.rgpu_comm_init.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.rgpu_get_level_info.argtypes = [C.c_int, C.POINTER(LevelInfo)]
    L.rgpu_set_timing.argtypes = [C.c_int]
    L.rgpu_set_pipeline.argtypes = [C.c_int]
    L.rgpu_level_totals.argtypes = [C.c_int, ip]
    L.rgpu_hydro_flag.argtypes = [C.c_int, dp, dp, ip]
    L.rgpu_upload_force.argtypes = [dp]
    L.rgpu_set_interpol_mag.argtypes = [C.c_int]
    L.rgpu_set_boundary_var.argtypes = [C.c_int, dp]
    L.rgpu_download_pressure_fix.argtypes = [dp, dp]
    L.rgpu_selftest_div.argtypes = [C.c_longlong, C.c_ulonglong, C.POINTER(C.c_longlong)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RgpuError(rc, load().rgpu_last_error().decode())


def exported_symbols():
    """Entry points declared in include/ramses_gpu.h (parsed from the header)."""
    import re
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"\b(rgpu_[a-z_0-9]+)\s*\(", txt)))
