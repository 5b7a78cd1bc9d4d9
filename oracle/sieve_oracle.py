"""ctypes wrapper around oracle/libsieve_oracle.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY -- see the header of sieve_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.  PARITY UNPINNED (the Java
reference cannot run here and holds no numeric golden vectors); pinned by independent derivations in
tests/test_oracle_pinning.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsieve_oracle.so")


def build(force=False):
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "sieve_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class LeafParams(C.Structure):
    _fields_ = [
        ("ado_rate", C.c_double),
        ("allelic_cov", C.c_double),
        ("allelic_raw_var", C.c_double),
        ("eff_seq_err", C.c_double),
        ("shape1", C.c_double),
        ("shape2", C.c_double),
        ("single_ado", C.c_int),
        ("use_log", C.c_int),
    ]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    L.so_log_gamma.restype = C.c_double
    L.so_log_gamma.argtypes = [C.c_double]
    L.so_log_beta.restype = C.c_double
    L.so_log_beta.argtypes = [C.c_double, C.c_double]
    L.so_log_sum_exp.restype = C.c_double
    L.so_log_sum_exp.argtypes = [_dp, C.c_int]
    L.so_dm_log_density.restype = C.c_double
    L.so_dm_log_density.argtypes = [_ip, _dp]
    L.so_nb_log_density.restype = C.c_double
    L.so_nb_log_density.argtypes = [C.c_double, C.c_double, C.c_double]
    L.so_size_factors.restype = None
    L.so_size_factors.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, _dp, _dp]
    L.so_seq_cov_likelihood.restype = C.c_double
    L.so_seq_cov_likelihood.argtypes = [C.c_int, C.c_int, C.c_double, C.POINTER(LeafParams)]
    L.so_nuc_likelihoods.restype = None
    L.so_nuc_likelihoods.argtypes = [_ip, C.POINTER(LeafParams), _dp]
    L.so_leaf_likelihood.restype = None
    L.so_leaf_likelihood.argtypes = [_ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(LeafParams), _dp]
    L.so_core_create.restype = C.c_void_p
    L.so_core_create.argtypes = [C.c_int] * 6
    L.so_core_free.argtypes = [C.c_void_p]
    L.so_core_set_use_scaling.argtypes = [C.c_void_p, C.c_double]
    L.so_core_set_node_partials.argtypes = [C.c_void_p, C.c_int, _dp, C.c_size_t]
    L.so_core_store_leaf_partials.argtypes = [C.c_void_p, C.c_int]
    L.so_core_set_node_matrix_for_update.argtypes = [C.c_void_p, C.c_int]
    L.so_core_set_node_partials_for_update.argtypes = [C.c_void_p, C.c_int]
    L.so_core_set_node_matrix.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    L.so_core_get_node_partials.argtypes = [C.c_void_p, C.c_int, _dp]
    L.so_core_get_const_partials.argtypes = [C.c_void_p, C.c_int, _dp]
    L.so_core_store.argtypes = [C.c_void_p]
    L.so_core_unstore.argtypes = [C.c_void_p]
    L.so_core_restore.argtypes = [C.c_void_p]
    L.so_core_calculate_partials.argtypes = [C.c_void_p] + [C.c_int] * 6
    L.so_core_integrate_partials.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, _dp, _dp]
    L.so_core_calculate_log_likelihoods.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    L.so_sum_pattern_logl.restype = C.c_double
    L.so_sum_pattern_logl.argtypes = [_dp, _ip, C.c_int]
    L.so_asc_bias_lewis.restype = C.c_double
    L.so_asc_bias_lewis.argtypes = [_dp, _ip, C.c_int]
    L.so_asc_bias_felsenstein.restype = C.c_double
    L.so_asc_bias_felsenstein.argtypes = [_dp, _ip, C.c_int, C.c_int, C.c_double, C.c_int]
    L.so_asc_bias_threaded.restype = C.c_double
    L.so_asc_bias_threaded.argtypes = [_dp, C.c_int, C.c_int, C.c_long, C.c_double]
    L.so_pattern_points.argtypes = [C.c_int, C.c_int, _ip]
    L.so_full_eval.restype = C.c_double
    L.so_full_eval.argtypes = [_ip, _ip, C.c_int, C.c_int, C.c_int, _dp, C.POINTER(LeafParams), _ip, C.c_int, _dp,
                               _dp, C.c_int, C.c_int, C.c_double, C.c_int, _dp, _dp]
    L.so_max_threads.restype = C.c_int
    L.so_background_log_likelihood.restype = C.c_double
    L.so_background_log_likelihood.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _ip, C.c_longlong,
                                               C.c_int, C.c_double, C.c_double]
    _lib = L
    return L


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


ASC_MODES = {"none": 0, "felsenstein_mean": 1, "felsenstein_geometric": 2, "lewis": 3}


def leaf_params(theta=0.3, t=50.0, v=50.0, eps=0.008, w1=150.0, w2=2.0, single_ado=True, use_log=True):
    """Start values of examples/templates/smc_stage_1.xml:30-38."""
    return LeafParams(theta, t, v, eps, w1, w2, int(single_ado), int(use_log))


def counts5(counts4):
    """[..., 4] (alt1, alt2, alt3, cov) -> [..., 5] (alt1, alt2, alt3, ref, cov), datatype/FullSupsCov.java:74-77."""
    c = np.asarray(counts4, dtype=np.int32)
    out = np.empty(c.shape[:-1] + (5,), dtype=np.int32)
    out[..., :3] = c[..., :3]
    out[..., 3] = c[..., 3] - c[..., :3].sum(-1)
    out[..., 4] = c[..., 3]
    return out


def log_gamma(x):
    return lib().so_log_gamma(float(x))


def log_sum_exp(a):
    a, p = _d(a)
    return lib().so_log_sum_exp(p, a.size)


def dm_log_density(counts_5, alpha_5):
    c, cp = _i(counts_5)
    a, ap = _d(alpha_5)
    return lib().so_dm_log_density(cp, ap)


def nb_log_density(x, r, p):
    return lib().so_nb_log_density(float(x), float(r), float(p))


def size_factors(counts_s_n_5, weights=None, zero_cov_mode=0):
    c, cp = _i(counts_s_n_5)
    S, n = c.shape[0], c.shape[1]
    wp = None
    if weights is not None:
        w, wp = _i(weights)
    sf = np.empty(n)
    st = np.empty(2)
    lib().so_size_factors(cp, wp, S, n, zero_cov_mode, sf.ctypes.data_as(_dp), st.ctypes.data_as(_dp))
    return sf, st


def leaf_likelihood(counts_s_n_5, taxon, size_factor, params, p0=0, p1=None):
    """RawReadCountsModelInterface.computeLeafLikelihood for one taxon -> [S, 4]."""
    c, cp = _i(counts_s_n_5)
    S, n = c.shape[0], c.shape[1]
    p1 = S if p1 is None else p1
    out = np.empty((p1 - p0, 4))
    lib().so_leaf_likelihood(cp, n, p0, p1, taxon, float(size_factor), C.byref(params), out.ctypes.data_as(_dp))
    return out


class Core:
    """Mirror of ScsBeerLikelihoodCore's MCMC-time API (likelihood/ScsBeerLikelihoodCore.java)."""

    def __init__(self, node_count, leaf_count, pattern_count, matrix_count=4, state_count=4, use_log=True):
        self.L = lib()
        self.n_nodes, self.n_leaves, self.S, self.K, self.G = node_count, leaf_count, pattern_count, matrix_count, state_count
        self.use_log = use_log
        self.h = self.L.so_core_create(node_count, leaf_count, pattern_count, matrix_count, state_count, int(use_log))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.so_core_free(self.h)
            self.h = None

    def set_use_scaling(self, scale):
        self.L.so_core_set_use_scaling(self.h, float(scale))

    def set_node_partials(self, node, partials):
        a, p = _d(partials)
        self.L.so_core_set_node_partials(self.h, node, p, a.size)

    def store_leaf_partials(self, node):
        self.L.so_core_store_leaf_partials(self.h, node)

    def set_node_matrix_for_update(self, node):
        self.L.so_core_set_node_matrix_for_update(self.h, node)

    def set_node_partials_for_update(self, node):
        self.L.so_core_set_node_partials_for_update(self.h, node)

    def set_node_matrix(self, node, cat, P16):
        a, p = _d(P16)
        assert a.size == self.G * self.G
        self.L.so_core_set_node_matrix(self.h, node, cat, p)

    def get_node_partials(self, node):
        out = np.empty((self.K, self.S, self.G))
        self.L.so_core_get_node_partials(self.h, node, out.ctypes.data_as(_dp))
        return out

    def get_const_partials(self, node):
        out = np.empty((self.K, self.S, self.G))
        self.L.so_core_get_const_partials(self.h, node, out.ctypes.data_as(_dp))
        return out

    def calculate_partials(self, c1, leaf1, c2, leaf2, parent, const_genotype=0):
        self.L.so_core_calculate_partials(self.h, c1, int(leaf1), c2, int(leaf2), parent, const_genotype)

    def integrate_partials(self, root, proportions, root_genotype=0):
        a, p = _d(proportions)
        out = np.empty(self.S)
        cr = np.empty(self.S)
        self.L.so_core_integrate_partials(self.h, root, p, root_genotype, out.ctypes.data_as(_dp), cr.ctypes.data_as(_dp))
        return out, cr

    def calculate_log_likelihoods(self, root_partials, const_root):
        a, ap = _d(root_partials)
        c, cp = _d(const_root)
        out = np.empty(self.S)
        lcr = np.empty(self.S)
        self.L.so_core_calculate_log_likelihoods(self.h, ap, out.ctypes.data_as(_dp), cp, lcr.ctypes.data_as(_dp))
        return out, lcr

    def store(self):
        self.L.so_core_store(self.h)

    def unstore(self):
        self.L.so_core_unstore(self.h)

    def restore(self):
        self.L.so_core_restore(self.h)


def sum_pattern_logl(ll, weights=None):
    a, p = _d(ll)
    wp = None
    if weights is not None:
        w, wp = _i(weights)
    return lib().so_sum_pattern_logl(p, wp, a.size)


def asc_bias(log_const_root, weights, mode, M, return_const_sum=False):
    a, p = _d(log_const_root)
    wp = None
    if weights is not None:
        w, wp = _i(weights)
    m = ASC_MODES[mode] if isinstance(mode, str) else mode
    if m == 0:
        return 0.0
    if m == 3:
        return lib().so_asc_bias_lewis(p, wp, a.size)
    return lib().so_asc_bias_felsenstein(p, wp, a.size, m - 1, float(M), int(return_const_sum))


def asc_bias_threaded(const_sums, mode, loci_nr, M):
    a, p = _d(const_sums)
    m = ASC_MODES[mode] if isinstance(mode, str) else mode
    return lib().so_asc_bias_threaded(p, a.size, m - 1, int(loci_nr), float(M))


def pattern_points(n_sites, T):
    pts = np.zeros(T + 1, dtype=np.int32)
    lib().so_pattern_points(n_sites, T, pts.ctypes.data_as(_ip))
    return pts


def max_threads():
    return lib().so_max_threads()


def full_eval(counts_s_n_5, size_factors_n, params, ops, matrices, proportions, root, asc_mode="none", M=0.0,
              n_threads=1, weights=None, want_sites=True):
    """One complete evaluation the way ThreadedScsTreeLikelihood does it (site ranges over threads).

    counts [S][n][5]; ops [n_ops][3] (parent, c1, c2|-1) post-order; matrices [2n][K][16].
    Returns (logP, site_logl, site_log_const)."""
    c, cp = _i(counts_s_n_5)
    S, n = c.shape[0], c.shape[1]
    sf, sfp = _d(size_factors_n)
    o, op = _i(ops)
    m, mp = _d(matrices)
    K = m.shape[1]
    pr, prp = _d(proportions)
    wp = None
    if weights is not None:
        w, wp = _i(weights)
    sl = np.empty(S) if want_sites else None
    sc = np.empty(S) if want_sites else None
    am = ASC_MODES[asc_mode] if isinstance(asc_mode, str) else asc_mode
    logp = lib().so_full_eval(cp, wp, n, S, K, sfp, C.byref(params), op, o.shape[0], mp, prp, int(root), am, float(M),
                              int(n_threads), sl.ctypes.data_as(_dp) if want_sites else None,
                              sc.ctypes.data_as(_dp) if want_sites else None)
    return logp, sl, sc


def background_hist_arrays(hists):
    """list of 5 {value: count} dicts (variant1, variant2, variant3, normal, coverage) -> (values, counts, offsets)."""
    vals, cnts, off = [], [], [0]
    for h in hists:
        for v in sorted(h):
            vals.append(v)
            cnts.append(h[v])
        off.append(len(vals))
    return (np.array(vals, dtype=np.int64), np.array(cnts, dtype=np.int64), np.array(off, dtype=np.int32))


def background_log_likelihood(hists, n_background_sites, n_taxa, eps, w1):
    """ScsBackgroundLikelihood.calculateLogPDirichletMultinomial (likelihood/ScsBackgroundLikelihood.java:169-268)."""
    v, c, off = background_hist_arrays(hists)
    ll = C.POINTER(C.c_longlong)
    return lib().so_background_log_likelihood(v.ctypes.data_as(ll), c.ctypes.data_as(ll), off.ctypes.data_as(_ip),
                                              int(n_background_sites), int(n_taxa), float(eps), float(w1))
