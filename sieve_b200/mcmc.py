"""ctypes binding of the native MCMC replay driver (include/sieve_b200_mcmc.h, sieve_b200/csrc/sieve_mcmc.cpp)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ASC_MODES, LeafParams, check

MIX = {"stage1": 1, "stage2": 2, "relaxed2": 3}
OP_NAMES = ["tree_scale", "root_scale", "uniform", "subtree_slide", "narrow", "wide", "wilson_balding", "prior_only",
            "coverage_params", "nucleotide_params", "ado_rate", "gamma_shape", "branch_rate"]


class McmcConfig(C.Structure):
    _fields_ = [("n_leaves", C.c_int32), ("n_categories", C.c_int32), ("asc_mode", C.c_int32),
                ("n_background_sites", C.c_double), ("gamma_shape", C.c_double), ("clock_rate", C.c_double),
                ("leaf", LeafParams), ("mix", C.c_int32), ("seed", C.c_uint64)]


class McmcStats(C.Structure):
    _fields_ = [("states", C.c_int64), ("accepted", C.c_int64), ("likelihood_evaluations", C.c_int64),
                ("ops_total", C.c_int64), ("matrices_total", C.c_int64), ("leaf_updates", C.c_int64),
                ("seconds", C.c_double), ("log_p", C.c_double), ("proposed", C.c_int64 * 13),
                ("accepted_by_op", C.c_int64 * 13), ("seconds_by_op", C.c_double * 13)]


COMBINE_FN = C.CFUNCTYPE(C.c_double, C.POINTER(_lib.ShardResult), C.c_void_p)
_bound = False


def _bind():
    global _bound
    L = _lib.lib()
    if not _bound:
        vp, ip, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
        L.sieve_mcmc_create.restype = C.c_int
        L.sieve_mcmc_create.argtypes = [C.POINTER(McmcConfig), vp, ip, ip, dp, dp, C.POINTER(vp)]
        L.sieve_mcmc_destroy.argtypes = [vp]
        L.sieve_mcmc_run.restype = C.c_int
        L.sieve_mcmc_run.argtypes = [vp, C.c_int64, C.c_int32, C.POINTER(McmcStats)]
        L.sieve_mcmc_get_state.restype = C.c_int
        L.sieve_mcmc_get_state.argtypes = [vp, ip, ip, dp, dp, C.POINTER(LeafParams), dp, dp]
        L.sieve_mcmc_full_recompute.restype = C.c_int
        L.sieve_mcmc_full_recompute.argtypes = [vp, dp]
        L.sieve_mcmc_set_combine.restype = C.c_int
        L.sieve_mcmc_set_combine.argtypes = [vp, COMBINE_FN, vp]
        L.sieve_gamma_category_rates.restype = C.c_int
        L.sieve_gamma_category_rates.argtypes = [C.c_double, C.c_int32, dp]
        _bound = True
    return L


def gamma_category_rates(shape, k=4):
    L = _bind()
    out = np.empty(k)
    check(L.sieve_gamma_category_rates(float(shape), k, out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


class McmcReplay:
    """Replays the reference's operator schedule against one core (one GPU shard)."""

    def __init__(self, core, tree, mix="stage1", asc_mode="none", n_background_sites=0.0, gamma_shape=1.0,
                 clock_rate=1.0, leaf=(0.3, 50.0, 50.0, 0.008, 150.0, 2.0), branch_rates=None, seed=1):
        self.L = _bind()
        self.core = core
        self.n = tree.n
        cfg = McmcConfig(tree.n, core.K, ASC_MODES[asc_mode], float(n_background_sites), float(gamma_shape),
                         float(clock_rate), LeafParams(*leaf), MIX[mix], int(seed))
        left = np.ascontiguousarray(tree.left, dtype=np.int32)
        right = np.ascontiguousarray(tree.right, dtype=np.int32)
        height = np.ascontiguousarray(tree.height, dtype=np.float64)
        br = None if branch_rates is None else np.ascontiguousarray(branch_rates, dtype=np.float64)
        ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        h = C.c_void_p()
        check(self.L.sieve_mcmc_create(C.byref(cfg), core.h, left.ctypes.data_as(ip), right.ctypes.data_as(ip),
                                       height.ctypes.data_as(dp), None if br is None else br.ctypes.data_as(dp),
                                       C.byref(h)))
        self.h = h

    def set_combine(self, fn):
        """fn(ShardResult) -> global log-likelihood (e.g. sieve_b200.distributed.global_log_p); None = single shard."""
        if fn is None:
            self._cb = COMBINE_FN()
        else:
            self._cb = COMBINE_FN(lambda r, user: float(fn(r.contents)))
        check(self.L.sieve_mcmc_set_combine(self.h, self._cb, None))

    def close(self):
        if getattr(self, "h", None):
            self.L.sieve_mcmc_destroy(self.h)
            self.h = None

    def run(self, n_states, reset=False):
        st = McmcStats()
        check(self.L.sieve_mcmc_run(self.h, int(n_states), int(reset), C.byref(st)))
        d = {k: getattr(st, k) for k in ("states", "accepted", "likelihood_evaluations", "ops_total", "matrices_total",
                                         "leaf_updates", "seconds", "log_p")}
        d["by_op"] = {OP_NAMES[i]: {"proposed": st.proposed[i], "accepted": st.accepted_by_op[i],
                                    "seconds": st.seconds_by_op[i]} for i in range(13) if st.proposed[i]}
        return d

    def state(self):
        from .tree import ScsTree
        N = 2 * self.n
        left = np.empty(N, dtype=np.int32)
        right = np.empty(N, dtype=np.int32)
        height = np.empty(N)
        rates = np.empty(N)
        leaf = LeafParams()
        gs = C.c_double()
        lp = C.c_double()
        ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        check(self.L.sieve_mcmc_get_state(self.h, left.ctypes.data_as(ip), right.ctypes.data_as(ip),
                                          height.ctypes.data_as(dp), rates.ctypes.data_as(dp), C.byref(leaf),
                                          C.byref(gs), C.byref(lp)))
        return dict(tree=ScsTree(self.n, left, right, height), branch_rates=rates,
                    leaf=(leaf.ado_rate, leaf.allelic_seq_cov, leaf.allelic_seq_cov_raw_var, leaf.eff_seq_err_rate,
                          leaf.shape_ctrl1, leaf.shape_ctrl2), gamma_shape=gs.value, log_p=lp.value)

    def full_recompute(self):
        lp = C.c_double()
        check(self.L.sieve_mcmc_full_recompute(self.h, C.byref(lp)))
        return lp.value
