"""`ScsCudaLikelihoodCore` -- the device-backed stand-in for SIEVE's likelihood core plus leaf model.

Method names and argument meaning follow the reference's `ScsBeerLikelihoodCore` / `LikelihoodCore`
(likelihood/ScsBeerLikelihoodCore.java; the calls `ScsTreeLikelihood.traverse` makes, likelihood/ScsTreeLikelihood.java:545-935)
so that a test written against the reference reads the same here.  Everything is a thin call into the C-ABI of
include/sieve_b200.h; no arithmetic of the hot path happens in Python.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (ASC_MODES, FLAG_CONST_PARTIALS, FLAG_EPHEMERAL, FLAG_SPARSE, FLAG_LOG_PARTIALS, FLAG_SINGLE_ADO, UPDATE_FLIP, UPDATE_PER_LEVEL,
                   LeafParams, ShardResult, check)


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class ScsCudaLikelihoodCore:
    def __init__(self, n_leaves, n_sites, n_categories=4, n_states=4, use_log_partials=True, asc_bias=False,
                 single_ado=True, device=-1, ephemeral=False, sparse=False, single_leaf_buffer=False,
                 external_counts=False, private_seq_cov=False):
        """initialize(nodeCount = 2n, leafCount = n, internalCount = n, patternCount, matrixCount, stateCount, ...)
        (likelihood/ScsBeerLikelihoodCore.java:91-128)."""
        self.L = _lib.lib()
        self.n, self.S, self.K, self.G = int(n_leaves), int(n_sites), int(n_categories), int(n_states)
        self.n_nodes = 2 * self.n
        self.use_log = bool(use_log_partials)
        self.asc_bias = bool(asc_bias)
        flags = (FLAG_LOG_PARTIALS if use_log_partials else 0) | (FLAG_CONST_PARTIALS if asc_bias else 0) | \
                (FLAG_SINGLE_ADO if single_ado else 0) | (FLAG_EPHEMERAL if ephemeral else 0) | \
                (FLAG_SPARSE if sparse else 0) | (_lib.FLAG_SINGLE_LEAF_BUFFER if single_leaf_buffer else 0) | \
                (_lib.FLAG_EXTERNAL_COUNTS if external_counts else 0) | \
                (_lib.FLAG_PRIVATE_SEQ_COV if private_seq_cov else 0)
        cfg = _lib.Config(self.n, self.S, self.K, self.G, flags, device)
        h = C.c_void_p()
        check(self.L.sieve_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.sieve_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data ----------------------------------------------------------------------------------------------
    def set_counts(self, counts):
        """int32 [site][cell][4 | 5] as the alignment holds it (alignment/ScsAlignment.java:188)."""
        c = np.ascontiguousarray(counts, dtype=np.int32)
        assert c.shape[:2] == (self.S, self.n) and c.shape[2] in (4, 5), c.shape
        check(self.L.sieve_set_counts(self.h, _iptr(c), c.shape[2]))

    def set_counts_cellmajor(self, counts):
        """[cell][site][4] int32 or uint16 in host memory (e.g. the memory-mapped tensor of sieve_b200.ingest)."""
        assert counts.shape == (self.n, self.S, 4) and counts.dtype in (np.int32, np.uint16) and counts.flags.c_contiguous
        check(self.L.sieve_set_counts_cellmajor(self.h, C.c_void_p(counts.ctypes.data), 0 if counts.dtype == np.int32 else 1))

    def set_counts_device(self, device_ptr):
        check(self.L.sieve_set_counts_device(self.h, C.c_void_p(int(device_ptr))))

    def set_pattern_weights(self, weights):
        if weights is None:
            check(self.L.sieve_set_pattern_weights(self.h, None))
            return
        w = np.ascontiguousarray(weights, dtype=np.int32)
        assert w.shape == (self.S,)
        check(self.L.sieve_set_pattern_weights(self.h, _iptr(w)))

    def compute_size_factors(self, zero_cov_mode=0):
        sf = np.empty(self.n)
        st = np.empty(2)
        check(self.L.sieve_compute_size_factors(self.h, zero_cov_mode, _dptr(sf), _dptr(st)))
        return sf, st

    def set_size_factors(self, sf):
        a = np.ascontiguousarray(sf, dtype=np.float64)
        assert a.shape == (self.n,)
        check(self.L.sieve_set_size_factors(self.h, _dptr(a)))

    # ---- leaves --------------------------------------------------------------------------------------------
    def set_leaf_params(self, theta, t, v, eps, w1, w2):
        p = LeafParams(theta, t, v, eps, w1, w2)
        check(self.L.sieve_set_leaf_params(self.h, C.byref(p)))

    # ---- PrivateAllelicSeqCovModel (rawreadcountsmodel/seqcovmodel/PrivateAllelicSeqCovModel.java) ----------------
    def set_private_seq_cov(self, t_ks, v_ks):
        """allelicSeqCovPerPattern / allelicSeqCovRawVarPerPattern, both [K][S] (index matrixIndex * nrOfPatterns + pattern)."""
        t = np.ascontiguousarray(t_ks, dtype=np.float64)
        v = np.ascontiguousarray(v_ks, dtype=np.float64)
        assert t.shape == (self.K, self.S) and v.shape == (self.K, self.S)
        check(self.L.sieve_set_private_seq_cov(self.h, _dptr(t), _dptr(v)))

    def get_private_seq_cov(self):
        t = np.empty((self.K, self.S))
        v = np.empty((self.K, self.S))
        check(self.L.sieve_get_private_seq_cov(self.h, _dptr(t), _dptr(v)))
        return t, v

    def private_update_from_ml(self, zero_cov_mode=0):
        """updateAllelicSeqCovAndRawVar (:312-432) from the max-sum trace of the current state, as
        ScsTreeLikelihood.postProcess does after an accepted move (likelihood/ScsTreeLikelihood.java:2454-2520).
        -> (number of (category, pattern) pairs whose (t, v) moved, tied uint8 [K][S])"""
        changed = C.c_int32()
        tied = np.zeros((self.K, self.S), dtype=np.uint8)
        check(self.L.sieve_private_update_from_ml(self.h, int(zero_cov_mode), C.byref(changed), tied.ctypes.data))
        return int(changed.value), tied

    def private_retraverse_changed(self, ops):
        """ScsTreeLikelihood.postProcess step 3 (:2493-2498): the patterns whose (t, v) moved are re-evaluated in place
        (leaves, whole tree, site log-likelihoods); the accepted state becomes the stored state.  -> number of sites"""
        o = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 3)
        n = C.c_int32()
        check(self.L.sieve_private_retraverse_changed(self.h, int(o.shape[0]), o.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n)))
        return int(n.value)

    def update_leaves(self, for_update=True):
        check(self.L.sieve_update_leaves(self.h, int(bool(for_update))))

    # ---- matrices ------------------------------------------------------------------------------------------
    def set_node_matrix_for_update(self, node):
        check(self.L.sieve_set_node_matrix_for_update(self.h, int(node)))

    def set_node_matrix(self, node, category, matrix):
        m = np.ascontiguousarray(matrix, dtype=np.float64).reshape(-1)
        assert m.size == 16
        check(self.L.sieve_set_node_matrix(self.h, int(node), int(category), _dptr(m)))

    def set_matrices(self, nodes, P, for_update=True):
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        m = np.ascontiguousarray(P, dtype=np.float64)
        assert m.size == nd.size * self.K * 16
        check(self.L.sieve_set_matrices(self.h, nd.size, _iptr(nd), _dptr(m), int(bool(for_update))))

    def set_distances(self, nodes, distances, for_update=True):
        """Fast path for the fixed rate matrix: distances [len(nodes)][K] = branch length * branch rate * site rate."""
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        d = np.ascontiguousarray(distances, dtype=np.float64)
        assert d.size == nd.size * self.K
        check(self.L.sieve_set_distances(self.h, nd.size, _iptr(nd), _dptr(d), int(bool(for_update))))

    # ---- partials ------------------------------------------------------------------------------------------
    def set_node_partials_for_update(self, node):
        check(self.L.sieve_set_node_partials_for_update(self.h, int(node)))

    def calculate_partials(self, child1, is_leaf1, child2, is_leaf2, parent, const_genotype=0):
        """calculateLogPartials / calculatePartials (queued; launched with the next evaluate)."""
        check(self.L.sieve_calculate_partials(self.h, int(child1), int(bool(is_leaf1)), int(child2), int(bool(is_leaf2)),
                                              int(parent), int(const_genotype)))

    calculate_log_partials = calculate_partials

    def update_nodes(self, ops, level_offsets=None, flip=True, per_level=False):
        o = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 3)
        mode = (UPDATE_FLIP if flip else 0) | (UPDATE_PER_LEVEL if per_level else 0)
        lo, nl = None, 0
        if level_offsets is not None:
            lo_a = np.ascontiguousarray(level_offsets, dtype=np.int32)
            lo, nl = _iptr(lo_a), lo_a.size - 1
        check(self.L.sieve_update_nodes(self.h, o.shape[0], _iptr(o), lo, nl, mode))

    def set_root(self, root, proportions, root_genotype=0):
        p = np.zeros(4)
        p[:self.K] = np.asarray(proportions, dtype=np.float64)[:self.K]
        check(self.L.sieve_set_root(self.h, int(root), _dptr(p), int(root_genotype)))

    def evaluate(self):
        r = ShardResult()
        check(self.L.sieve_evaluate(self.h, C.byref(r)))
        return r

    def step(self, matrix_nodes, P, ops, update_leaves=False):
        nd = np.ascontiguousarray(matrix_nodes, dtype=np.int32)
        m = np.ascontiguousarray(P, dtype=np.float64)
        o = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 3)
        r = ShardResult()
        check(self.L.sieve_step(self.h, int(bool(update_leaves)), nd.size, _iptr(nd), _dptr(m), o.shape[0], _iptr(o),
                                C.byref(r)))
        return r

    def step_distances(self, nodes, distances, ops, update_leaves=False):
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        d = np.ascontiguousarray(distances, dtype=np.float64)
        o = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 3)
        assert d.size == nd.size * self.K
        r = ShardResult()
        check(self.L.sieve_step_distances(self.h, int(bool(update_leaves)), nd.size, _iptr(nd), _dptr(d), o.shape[0],
                                          _iptr(o), C.byref(r)))
        return r

    def traverse_full(self, tree_root, ops, P_all_nodes, update_leaves=False, want_sites=True):
        """The reference's own traverse call sequence, issued natively call by call (include/sieve_b200_mcmc.h,
        sieve_traverse_full): the no-extra-edit integration.  P_all_nodes [2n][K][16]."""
        o = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 3)
        m = np.ascontiguousarray(P_all_nodes, dtype=np.float64)
        assert m.size == self.n_nodes * self.K * 16
        site = np.empty(self.S) if want_sites else None
        r = ShardResult()
        fn = self.L.sieve_traverse_full
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                       C.c_int32, C.POINTER(C.c_double), C.POINTER(ShardResult)]
        check(fn(self.h, self.n_nodes, self.K, int(tree_root), o.shape[0], _iptr(o), _dptr(m), int(bool(update_leaves)),
                 _dptr(site) if want_sites else None, C.byref(r)))
        return r, site

    def last_stage_bytes(self):
        n = C.c_uint64()
        check(self.L.sieve_last_stage_bytes(self.h, C.byref(n)))
        return n.value

    def replay(self, update_leaves=False, n_times=1):
        """Re-launch the last step's kernels from what is staged in HBM (asynchronous, no copies)."""
        check(self.L.sieve_replay(self.h, int(bool(update_leaves)), int(n_times)))

    def site_log_likelihoods(self):
        """getPatternLogLikelihoods + logConstRoot (likelihood/ScsTreeLikelihood.java:2958-2970, 915-920)."""
        a = np.empty(self.S)
        b = np.empty(self.S)
        check(self.L.sieve_get_site_log_likelihoods(self.h, _dptr(a), _dptr(b)))
        return a, b

    def get_node_partials(self, node, const=False):
        out = np.empty((self.K, self.S, self.G))
        check(self.L.sieve_get_node_partials(self.h, int(node), int(bool(const)), _dptr(out)))
        return out

    def set_elision(self, on=True):
        """Skip dirty nodes whose inputs did not change (see include/sieve_b200.h, "Change tracking")."""
        check(self.L.sieve_set_elision(self.h, 1 if on else 0))

    def sparse_info(self):
        """-> dict(pool_slots, keep_leaves, free_slots) (keep_leaves == 0 for a dense core)."""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        check(self.L.sieve_sparse_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"pool_slots": int(a.value), "keep_leaves": int(b.value), "free_slots": int(c.value)}

    def elided_count(self):
        v = C.c_uint64()
        check(self.L.sieve_elided_count(self.h, C.byref(v)))
        return int(v.value)

    # ---- max-sum pass (variant calling) --------------------------------------------------------------------
    def ml_trace(self):
        """The traceMLGenotypes evaluation of the current state (likelihood/ScsTreeLikelihood.java:700-722, 947-1460)."""
        check(self.L.sieve_ml_trace(self.h))

    def ml_call(self):
        """-> (genotypes int8 [node][S], ado int8 [leaf][S], category uint8 [S], ambiguous uint8 [S]); -1 = tied."""
        g = np.empty((self.n_nodes, self.S), dtype=np.int8)
        a = np.empty((self.n, self.S), dtype=np.int8)
        c = np.empty(self.S, dtype=np.uint8)
        m = np.empty(self.S, dtype=np.uint8)
        check(self.L.sieve_ml_call(self.h, g.ctypes.data, a.ctypes.data, c.ctypes.data, m.ctypes.data))
        return g, a, c, m

    def ml_categories(self):
        """getPatternCategories (likelihood/ScsBeerLikelihoodCore.java:278-296) as a bit mask per site."""
        out = np.empty(self.S, dtype=np.uint8)
        check(self.L.sieve_ml_get_categories(self.h, out.ctypes.data))
        return out

    def ml_node(self, node):
        """-> (ML partials [K][S][4], back-pointer bytes uint8 [K][S][4], genotype collection uint8 [K][S])."""
        p = np.empty((self.K, self.S, 4))
        b = np.empty((self.K, self.S, 4), dtype=np.uint8)
        c = np.empty((self.K, self.S), dtype=np.uint8)
        check(self.L.sieve_ml_get_node(self.h, int(node), _dptr(p), b.ctypes.data, c.ctypes.data))
        return p, b, c

    def ml_site(self, site):
        """-> (back-pointer bytes uint8 [node][K][4], sum-product log partials [node][K][4]) of one site."""
        b = np.empty((self.n_nodes, self.K, 4), dtype=np.uint8)
        p = np.empty((self.n_nodes, self.K, 4))
        check(self.L.sieve_ml_get_site(self.h, int(site), b.ctypes.data, _dptr(p)))
        return b, p

    def store(self):
        check(self.L.sieve_store(self.h))

    def unstore(self):
        check(self.L.sieve_unstore(self.h))

    def restore(self):
        check(self.L.sieve_restore(self.h))

    # ---- multi-process shards ---------------------------------------------------------------------------------
    def comm_init(self, rank, world, unique_id_bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id_bytes))
        check(self.L.sieve_comm_init(self.h, int(rank), int(world), buf))

    def comm_size(self):
        return self.L.sieve_comm_size(self.h)

    def comm_exchange_mode(self):
        """0: no communicator; 1: ncclAllGather behind the reduction; 2: fused into the reduction kernel over peer memory"""
        return self.L.sieve_comm_exchange_mode(self.h)

    def last_global_results(self):
        n = self.comm_size()
        arr = (ShardResult * n)()
        check(self.L.sieve_last_global_results(self.h, arr))
        return list(arr)

    # ---- instrumentation -------------------------------------------------------------------------------------
    def enable_timing(self, on=True):
        check(self.L.sieve_enable_timing(self.h, int(on)))

    def last_timing(self):
        ms = (C.c_float * 4)()
        n = C.c_int32()
        check(self.L.sieve_last_timing(self.h, ms, C.byref(n)))
        return dict(leaf_ms=ms[0], prune_ms=ms[1], reduce_ms=ms[2], total_ms=ms[3], launches=n.value)

    def launch_count(self):
        n = C.c_uint64()
        check(self.L.sieve_launch_count(self.h, C.byref(n)))
        return n.value

    def device_bytes(self):
        n = C.c_uint64()
        check(self.L.sieve_device_bytes(self.h, C.byref(n)))
        return n.value

    def stream(self):
        p = C.c_void_p()
        check(self.L.sieve_get_stream(self.h, C.byref(p)))
        return p.value


def combine_shards(results, asc_mode="none", n_background_sites=0.0):
    """ThreadedScsTreeLikelihood.calculateLogPByBeagle (likelihood/ThreadedScsTreeLikelihood.java:366-389)."""
    arr = (ShardResult * len(results))(*results)
    m = ASC_MODES[asc_mode] if isinstance(asc_mode, str) else int(asc_mode)
    return _lib.lib().sieve_combine_shards(arr, len(results), m, float(n_background_sites))


def pattern_points(n_sites, n_shards):
    pts = np.zeros(n_shards + 1, dtype=np.int32)
    _lib.lib().sieve_pattern_points(int(n_sites), int(n_shards), _iptr(pts))
    return pts


def background_log_likelihood(hists, n_background_sites, n_taxa, eps, w1, device=-1):
    """ScsBackgroundLikelihood.calculateLogP for the Dirichlet-multinomial model
    (likelihood/ScsBackgroundLikelihood.java:169-268). `hists`: 5 {value: count} dicts in the order
    (variant1, variant2, variant3, normal, coverage)."""
    vals, cnts, off = [], [], [0]
    for h in hists:
        for v in sorted(h):
            vals.append(v)
            cnts.append(h[v])
        off.append(len(vals))
    v = np.array(vals, dtype=np.int64)
    c = np.array(cnts, dtype=np.int64)
    o = np.array(off, dtype=np.int32)
    out = C.c_double()
    lp = C.POINTER(C.c_int64)
    check(_lib.lib().sieve_background_log_likelihood(device, v.ctypes.data_as(lp), c.ctypes.data_as(lp), _iptr(o),
                                                     int(n_background_sites), int(n_taxa), float(eps), float(w1),
                                                     C.byref(out)))
    return out.value


def comm_unique_id():
    buf = (C.c_uint8 * 128)()
    check(_lib.lib().sieve_comm_unique_id(buf))
    return bytes(buf)
