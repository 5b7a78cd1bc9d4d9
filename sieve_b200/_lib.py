"""ctypes binding of libsieve_b200.so (include/sieve_b200.h).  The library is built in-tree by
sieve_b200/csrc/Makefile (nvcc, sm_100a).  Loading fails loudly when it is missing: there is no CPU fallback."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SIEVE_B200_LIB") or os.path.join(_HERE, "libsieve_b200.so")   # override: kernel-variant experiments only
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "sieve_b200.h")

OK, ERR_INVALID, ERR_CUDA, ERR_OOM, ERR_STATE = 0, 1, 2, 3, 4
FLAG_LOG_PARTIALS, FLAG_CONST_PARTIALS, FLAG_SINGLE_ADO, FLAG_EPHEMERAL, FLAG_SPARSE, FLAG_SINGLE_LEAF_BUFFER = 1, 2, 4, 8, 16, 32
FLAG_EXTERNAL_COUNTS = 64
FLAG_PRIVATE_SEQ_COV = 128
UPDATE_FLIP, UPDATE_PER_LEVEL = 1, 2
ASC_NONE, ASC_FELSENSTEIN_MEAN, ASC_FELSENSTEIN_GEOMETRIC, ASC_LEWIS = 0, 1, 2, 3
ASC_MODES = {"none": 0, "felsenstein_mean": 1, "felsenstein_geometric": 2, "lewis": 3}


class SieveError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sieve_b200 error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("n_leaves", C.c_int32), ("n_sites", C.c_int32), ("n_categories", C.c_int32), ("n_states", C.c_int32),
                ("flags", C.c_uint32), ("device", C.c_int32)]


class LeafParams(C.Structure):
    _fields_ = [("ado_rate", C.c_double), ("allelic_seq_cov", C.c_double), ("allelic_seq_cov_raw_var", C.c_double),
                ("eff_seq_err_rate", C.c_double), ("shape_ctrl1", C.c_double), ("shape_ctrl2", C.c_double)]


class ShardResult(C.Structure):
    _fields_ = [("sum_log_l", C.c_double), ("const_lse", C.c_double), ("const_sum", C.c_double),
                ("lewis_sum", C.c_double), ("weight_sum", C.c_double)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)

# callbacks of the distributed size-factor protocol (include/sieve_b200.h, sieve_sf_ops_t)
SF_MOMENTS_FN = C.CFUNCTYPE(C.c_int, _vp, _u64p)
SF_WEIGHT_LE_FN = C.CFUNCTYPE(C.c_int, _vp, C.c_int32, _u64p, _i64p)
SF_ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, _vp, _i64p, C.c_int32)


class SfOps(C.Structure):
    _fields_ = [("user", _vp), ("n_cells", C.c_int32), ("moments", SF_MOMENTS_FN), ("weight_le", SF_WEIGHT_LE_FN),
                ("allreduce_sum_i64", SF_ALLREDUCE_FN)]


# name -> (restype, argtypes); every symbol include/sieve_b200.h declares
SIGNATURES = {
    "sieve_last_error": (C.c_char_p, []),
    "sieve_version": (C.c_int, []),
    "sieve_device_count": (C.c_int, []),
    "sieve_create": (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "sieve_destroy": (C.c_int, [_vp]),
    "sieve_device_bytes": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "sieve_set_counts": (C.c_int, [_vp, _ip, C.c_int32]),
    "sieve_set_counts_cellmajor": (C.c_int, [_vp, _vp, C.c_int32]),
    "sieve_set_counts_device": (C.c_int, [_vp, _vp]),
    "sieve_set_pattern_weights": (C.c_int, [_vp, _ip]),
    "sieve_compute_size_factors": (C.c_int, [_vp, C.c_int32, _dp, _dp]),
    "sieve_set_size_factors": (C.c_int, [_vp, _dp]),
    "sieve_sf_open": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _vp, _ip, C.c_int32, C.POINTER(_vp)]),
    "sieve_sf_open_core": (C.c_int, [_vp, C.c_int32, C.POINTER(_vp)]),
    "sieve_sf_close": (C.c_int, [_vp]),
    "sieve_sf_moments": (C.c_int, [_vp, _u64p]),
    "sieve_sf_weight_le": (C.c_int, [_vp, C.c_int32, _u64p, _i64p]),
    "sieve_sf_solve": (C.c_int, [C.POINTER(SfOps), _dp, _dp]),
    "sieve_sf_solve_parts": (C.c_int, [C.c_int32, C.POINTER(_vp), SF_ALLREDUCE_FN, _vp, _dp, _dp]),
    "sieve_set_private_seq_cov": (C.c_int, [_vp, _dp, _dp]),
    "sieve_get_private_seq_cov": (C.c_int, [_vp, _dp, _dp]),
    "sieve_private_update_from_ml": (C.c_int, [_vp, C.c_int32, _ip, _vp]),
    "sieve_private_retraverse_changed": (C.c_int, [_vp, C.c_int32, _ip, _ip]),
    "sieve_set_leaf_params": (C.c_int, [_vp, C.POINTER(LeafParams)]),
    "sieve_update_leaves": (C.c_int, [_vp, C.c_int32]),
    "sieve_set_node_matrix_for_update": (C.c_int, [_vp, C.c_int32]),
    "sieve_set_node_matrix": (C.c_int, [_vp, C.c_int32, C.c_int32, _dp]),
    "sieve_set_matrices": (C.c_int, [_vp, C.c_int32, _ip, _dp, C.c_int32]),
    "sieve_set_distances": (C.c_int, [_vp, C.c_int32, _ip, _dp, C.c_int32]),
    "sieve_set_node_partials_for_update": (C.c_int, [_vp, C.c_int32]),
    "sieve_calculate_partials": (C.c_int, [_vp] + [C.c_int32] * 6),
    "sieve_update_nodes": (C.c_int, [_vp, C.c_int32, _ip, _ip, C.c_int32, C.c_uint32]),
    "sieve_set_root": (C.c_int, [_vp, C.c_int32, _dp, C.c_int32]),
    "sieve_evaluate": (C.c_int, [_vp, C.POINTER(ShardResult)]),
    "sieve_step": (C.c_int, [_vp, C.c_int32, C.c_int32, _ip, _dp, C.c_int32, _ip, C.POINTER(ShardResult)]),
    "sieve_step_distances": (C.c_int, [_vp, C.c_int32, C.c_int32, _ip, _dp, C.c_int32, _ip, C.POINTER(ShardResult)]),
    "sieve_last_stage_bytes": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "sieve_replay": (C.c_int, [_vp, C.c_int32, C.c_int32]),
    "sieve_get_site_log_likelihoods": (C.c_int, [_vp, _dp, _dp]),
    "sieve_get_result_device_ptr": (C.c_int, [_vp, C.POINTER(_vp)]),
    "sieve_get_node_partials": (C.c_int, [_vp, C.c_int32, C.c_int32, _dp]),
    "sieve_set_elision": (C.c_int, [_vp, C.c_int32]),
    "sieve_sparse_info": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "sieve_elided_count": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "sieve_walk_blocks_per_sm": (C.c_int, [C.c_int64]),
    "sieve_ml_trace": (C.c_int, [_vp]),
    "sieve_ml_call": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "sieve_ml_get_categories": (C.c_int, [_vp, _vp]),
    "sieve_ml_get_node": (C.c_int, [_vp, C.c_int32, _dp, _vp, _vp]),
    "sieve_ml_get_site": (C.c_int, [_vp, C.c_int32, _vp, _dp]),
    "sieve_store": (C.c_int, [_vp]),
    "sieve_unstore": (C.c_int, [_vp]),
    "sieve_restore": (C.c_int, [_vp]),
    "sieve_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "sieve_comm_init": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]),
    "sieve_comm_size": (C.c_int, [_vp]),
    "sieve_comm_exchange_mode": (C.c_int, [_vp]),
    "sieve_last_global_results": (C.c_int, [_vp, C.POINTER(ShardResult)]),
    "sieve_combine_shards": (C.c_double, [C.POINTER(ShardResult), C.c_int32, C.c_int32, C.c_double]),
    "sieve_pattern_points": (None, [C.c_int32, C.c_int32, _ip]),
    "sieve_background_log_likelihood": (C.c_int, [C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _ip, C.c_int64,
                                                   C.c_int32, C.c_double, C.c_double, _dp]),
    "sieve_last_timing": (C.c_int, [_vp, C.POINTER(C.c_float), _ip]),
    "sieve_enable_timing": (C.c_int, [_vp, C.c_int32]),
    "sieve_launch_count": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "sieve_get_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
}

_lib = None


def build(force=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    newest = max(os.path.getmtime(os.path.join(src_dir, f)) for f in os.listdir(src_dir))
    newest = max(newest, os.path.getmtime(HEADER_PATH))
    if force or not os.path.exists(SO_PATH) or os.path.getmtime(SO_PATH) < newest:
        subprocess.check_call(["make", "-C", src_dir, "-s"])
    return SO_PATH


def lib():
    """The loaded library; raises if it was not built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("libsieve_b200.so is missing: run `make -C sieve_b200/csrc` (or __graft_entry__.build())")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code):
    if code != OK:
        raise SieveError(code, lib().sieve_last_error().decode("utf-8", "replace"))
