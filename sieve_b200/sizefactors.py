"""Global size factors of a site-sharded alignment (include/sieve_b200.h, `sieve_sf_*`).

The reference initialises the coverage model on the FULL alignment before it splits the sites over its workers
(likelihood/ThreadedScsTreeLikelihood.java:119-131, seqcovmodel/SeqCovModelInterface.java:291-351), so every worker
uses the same per-cell size factors.  Here every GPU only ever holds its own site range; the per-cell median of the
coverage ratios over all sites is found by exact distributed selection: per round each part counts on its device, the
integer counts are summed over parts / processes.  All arithmetic lives in the library; this module only wires

  * `SizeFactorPart`     -- one device-resident count tensor (a part of the alignment) -> the two local primitives;
  * `solve_parts`        -- parts driven by this process (+ an optional all-reduce across processes);
  * `solve_with_ops`     -- the protocol over arbitrary callbacks (CPU tests drive it with stand-in primitives);
  * `torch_allreduce`    -- the all-reduce callback over torch.distributed (gloo or NCCL).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


class SizeFactorPart:
    """A part of the alignment on one device: `counts` is a DEVICE pointer to int32 [cell][site][4] (the core's layout),
    borrowed for the lifetime of the part; or a core whose counts and pattern weights are used."""

    def __init__(self, n_cells=None, n_sites=None, counts_device_ptr=None, weights=None, zero_cov_mode=0, device=-1, core=None):
        self.L = _lib.lib()
        h = C.c_void_p()
        if core is not None:
            check(self.L.sieve_sf_open_core(core.h, int(zero_cov_mode), C.byref(h)))
            self.n = core.n
        else:
            w = None
            if weights is not None:
                w = np.ascontiguousarray(weights, dtype=np.int32)
                assert w.shape == (n_sites,)
            check(self.L.sieve_sf_open(int(device), int(n_cells), int(n_sites), C.c_void_p(int(counts_device_ptr)),
                                       None if w is None else w.ctypes.data_as(C.POINTER(C.c_int32)), int(zero_cov_mode),
                                       C.byref(h)))
            self.n = int(n_cells)
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.sieve_sf_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def moments(self):
        out = (C.c_uint64 * 3)()
        check(self.L.sieve_sf_moments(self.h, out))
        return [int(x) for x in out]

    def weight_le(self, pivots):
        """pivots uint64 [n][p] (bit patterns of doubles), p in {1, 2} -> int64 [n][p]."""
        pv = np.ascontiguousarray(pivots, dtype=np.uint64)
        out = np.zeros(pv.shape, dtype=np.int64)
        check(self.L.sieve_sf_weight_le(self.h, pv.shape[1], pv.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out


def torch_allreduce(group=None):
    """-> SF_ALLREDUCE_FN summing an int64 buffer over the ranks of a torch.distributed group (in place)."""
    import torch
    import torch.distributed as dist

    def fn(_user, buf, count):
        try:
            a = np.ctypeslib.as_array(buf, shape=(count,))
            if dist.get_backend(group) == "nccl":
                t = torch.from_numpy(a).cuda()
                dist.all_reduce(t, group=group)
                a[:] = t.cpu().numpy()
            else:
                t = torch.from_numpy(a)      # shares memory with the caller's buffer
                dist.all_reduce(t, group=group)
            return 0
        except Exception:   # no exception may cross the C boundary
            return 1
    return _lib.SF_ALLREDUCE_FN(fn)


def solve_parts(parts, allreduce=None):
    """-> (size_factors [n], stats [2] = (mean(cov) / 2, var(cov) / 4)) over all parts of all processes."""
    L = _lib.lib()
    n = parts[0].n
    arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
    sf = np.empty(n)
    st = np.empty(2)
    cb = allreduce if allreduce is not None else _lib.SF_ALLREDUCE_FN(0)
    check(L.sieve_sf_solve_parts(len(parts), arr, cb, None, sf.ctypes.data_as(C.POINTER(C.c_double)),
                                 st.ctypes.data_as(C.POINTER(C.c_double))))
    return sf, st


def solve_with_ops(n_cells, moments, weight_le, allreduce=None):
    """The protocol over Python callables: moments() -> (sum, sum of squares, entries) of what THIS process holds;
    weight_le(pivots uint64 [n][p]) -> int64 [n][p]; allreduce: an SF_ALLREDUCE_FN or None."""
    L = _lib.lib()

    def c_moments(_u, out3):
        try:
            m = moments()
            for i in range(3):
                out3[i] = int(m[i])
            return 0
        except Exception:
            return 1

    def c_weight_le(_u, n_piv, piv, out):
        try:
            p = np.ctypeslib.as_array(piv, shape=(n_cells, n_piv))
            o = np.ctypeslib.as_array(out, shape=(n_cells, n_piv))
            o[:] = weight_le(p)
            return 0
        except Exception:
            return 1
    ops = _lib.SfOps(None, int(n_cells), _lib.SF_MOMENTS_FN(c_moments), _lib.SF_WEIGHT_LE_FN(c_weight_le),
                     allreduce if allreduce is not None else _lib.SF_ALLREDUCE_FN(0))
    sf = np.empty(n_cells)
    st = np.empty(2)
    check(L.sieve_sf_solve(C.byref(ops), sf.ctypes.data_as(C.POINTER(C.c_double)), st.ctypes.data_as(C.POINTER(C.c_double))))
    return sf, st
