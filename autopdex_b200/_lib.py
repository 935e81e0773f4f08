"""ctypes binding of libapdx_b200.so (include/apdx_b200.h).

The library is the product; this module only declares its C ABI.  There is no CPU
fallback: if the shared library is missing or no CUDA device is present, every compute
entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("APDX_LIB") or os.path.join(_HERE, "lib", "libapdx_b200.so")   # APDX_LIB: A/B builds

APDX_PARAM = {"coefficient": 0, "source": 1, "youngs_modulus": 2, "poisson_ratio": 3, "body_load": 4, "traction": 5}
APDX_MODEL = {"poisson_potential": 0, "poisson_weak": 1, "linear_elasticity": 2, "neo_hooke": 3, "neumann": 4,
              "capacity": 5, "pattern_only": 6}
APDX_MODE = {None: 0, "plain strain": 1, "plain stress": 2, "3d": 3, "lame": 4}
APDX_KIND = {"domain": 0, "surface": 1, "intpoint": 2}
APDX_LAYOUT = {"const": 0, "per_gp": 1, "per_row_gp": 2}
APDX_KRYLOV = {"cg": 0, "bicgstab": 1}


class ApdxError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libapdx_b200 error %d: %s" % (code, message))
        self.code = code


class SetDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("model", C.c_int32), ("mode", C.c_int32), ("nen", C.c_int32),
                ("n_gp", C.c_int32), ("dim_ref", C.c_int32), ("conn_itemsize", C.c_int32), ("reserved", C.c_int32),
                ("n_rows", C.c_int64), ("conn_h", C.c_void_p), ("shape_n_h", C.c_void_p),
                ("shape_dn_h", C.c_void_p), ("gp_w_h", C.c_void_p)]


class KrylovOpts(C.Structure):
    _fields_ = [("method", C.c_int32), ("maxiter", C.c_int32), ("rtol", C.c_double), ("atol", C.c_double),
                ("jacobi", C.c_int32), ("check_every", C.c_int32)]


# name -> (restype, argtypes); mirrors include/apdx_b200.h one to one
_P, _I32, _I64, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_double
SIGNATURES = {
    "apdx_abi_version": (C.c_int, []),
    "apdx_last_error": (C.c_char_p, []),
    "apdx_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "apdx_set_device": (C.c_int, [C.c_int]),
    "apdx_malloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "apdx_free": (C.c_int, [_P]),
    "apdx_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "apdx_host_free": (C.c_int, [_P]),
    "apdx_host_register": (C.c_int, [_P, C.c_size_t]),
    "apdx_host_unregister": (C.c_int, [_P]),
    "apdx_memcpy_h2d": (C.c_int, [_P, _P, C.c_size_t]),
    "apdx_memcpy_d2h": (C.c_int, [_P, _P, C.c_size_t]),
    "apdx_memset": (C.c_int, [_P, C.c_int, C.c_size_t]),
    "apdx_synchronize": (C.c_int, []),
    "apdx_mem_info": (C.c_int, [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "apdx_plan_create": (C.c_int, [C.POINTER(_P), _I32, _I64, _I32, _I32, C.POINTER(SetDesc), _P]),
    "apdx_plan_destroy": (C.c_int, [_P]),
    "apdx_plan_query": (C.c_int, [_P, C.POINTER(_I64)]),
    "apdx_plan_get_csr": (C.c_int, [_P, C.c_int, _P, _P]),
    "apdx_plan_get_elem_map": (C.c_int, [_P, _I64, _I64, _P]),
    "apdx_set_coords": (C.c_int, [_P, _P]),
    "apdx_set_param": (C.c_int, [_P, _I32, _I32, _I32, _I32, _P]),
    "apdx_set_intpoint_tables": (C.c_int, [_P, _I32, _P, _P, _P]),
    "apdx_set_time_increment": (C.c_int, [_P, _D]),
    "apdx_set_dofs_n": (C.c_int, [_P, _P]),
    "apdx_assemble": (C.c_int, [_P, _P, C.c_int, _P]),
    "apdx_get_values": (C.c_int, [_P, C.c_int, _P]),
    "apdx_get_coo_values": (C.c_int, [_P, _I64, _I64, _P]),
    "apdx_plan_newton_history": (C.c_int, [_P, _P, _I32, C.POINTER(_I32)]),
    "apdx_plan_set_coarse": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "apdx_plan_set_coarse_structured": (C.c_int, [_P, _P, _I32, _P, _P, _I64, _I64]),
    "apdx_plan_get_transfer": (C.c_int, [_P, _I32, C.POINTER(_I64), C.POINTER(_I64), _P, _P, _P, _P]),
    "apdx_plan_set_multigrid": (C.c_int, [_P, _I32, _I32, _I32, _D, _D]),
    "apdx_spmv": (C.c_int, [_P, _P, _P]),
    "apdx_krylov": (C.c_int, [_P, C.POINTER(KrylovOpts), _P, _P, C.POINTER(_I32), C.POINTER(_D)]),
    "apdx_linear_step": (C.c_int, [_P, C.POINTER(KrylovOpts), _P, _P, _P, C.POINTER(_I32)]),
    "apdx_newton": (C.c_int, [_P, C.POINTER(KrylovOpts), _P, _P, _D, _I32, _D, C.POINTER(_I32), C.POINTER(_D),
                              C.POINTER(_I32)]),
    "apdx_tangent_solve": (C.c_int, [_P, C.POINTER(KrylovOpts), _P, _P, C.c_int, _P, C.POINTER(_I32)]),
    "apdx_plan_stats": (C.c_int, [_P, C.POINTER(_D)]),
    "apdx_plan_last_krylov": (C.c_int, [_P, C.POINTER(_D), C.POINTER(_I32)]),
    "apdx_plan_sell_info": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "apdx_plan_set_stream": (C.c_int, [_P, _P]),
    "apdx_assemble_async": (C.c_int, [_P, _P, C.c_int, _P]),
    "apdx_spmv_async": (C.c_int, [_P, _P, _P]),
    "apdx_stream_create": (C.c_int, [C.POINTER(_P)]),
    "apdx_stream_synchronize": (C.c_int, [_P]),
    "apdx_stream_destroy": (C.c_int, [_P]),
    "apdx_time_spmv": (C.c_int, [_P, _I32, C.POINTER(_D)]),
    "apdx_measure_fp64_peak": (C.c_int, [C.POINTER(_D)]),
    "apdx_comm_allreduce_host": (C.c_int, [_P, _I32, _I32]),
    "apdx_comm_unique_id": (C.c_int, [_P]),
    "apdx_comm_init": (C.c_int, [_P, _I32, _I32]),
    "apdx_comm_destroy": (C.c_int, []),
    "apdx_plan_set_partition": (C.c_int, [_P, _I64, _I64, _I32, _I32]),
    "apdx_plan_comm_info": (C.c_int, [_P, C.POINTER(_D)]),
    "apdx_comm_exchange_planes": (C.c_int, [_P, _I64, _I64, _I64, _I32, _I32]),
    "apdx_plan_set_partition_lists": (C.c_int, [_P, _I64, _I32, _P, _P, _P, _P, _P]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "autopdex_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C autopdex_b200/csrc`). The b200 backend has no CPU path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.apdx_abi_version() != 1:
        raise ImportError("autopdex_b200: ABI version mismatch between _lib.py and %s" % LIB_PATH)
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise ApdxError(rc, load().apdx_last_error().decode("utf-8", "replace"))


def device_count():
    n = C.c_int(0)
    rc = load().apdx_device_count(C.byref(n))
    return n.value if rc == 0 else 0
