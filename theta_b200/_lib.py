"""ctypes binding of libtheta_b200.so (include/theta_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device
is usable, the product path raises ThetaError.  Nothing here imports oracle/.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# THETA_LIB: a differently built copy of the same library (kernel variants during development, tools/build_variants.py)
LIB_PATH = os.environ.get("THETA_LIB") or os.path.join(HERE, "libtheta_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_EMPTY, ERR_OVERFLOW, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
ST_OK, ST_ROOT_LB, ST_ROOT_UB, ST_FALLBACK_X0, ST_DEGENERATE, ST_FALLBACK_MOVED = 0, 1, 2, 3, 4, 5
ST_NONE_SIGN, ST_NONE_NAN, ST_NONE_NOCONV, ST_NONE_RANGE = 8, 9, 10, 11
TIE_MARGIN = 10e-4  # Misc.isClose default margin (python/Misc.py:36)


class ThetaError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "theta_b200 error %d: %s" % (code, msg))
        self.code = code


class EmptySpaceError(ThetaError):
    """No admissible copy-number profile (RunTHetA.py:154-156, 217-219)."""


class Problem(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("tau", C.c_int32), ("reserved", C.c_int32),
                ("r", C.POINTER(C.c_int64)), ("rN", C.POINTER(C.c_int64)),
                ("lower", C.POINTER(C.c_int32)), ("upper", C.POINTER(C.c_int32)),
                ("max_normal", C.c_double), ("min_normal", C.c_double)]


class Cand(C.Structure):
    _fields_ = [("rank", C.c_uint64), ("nll", C.c_double), ("mu", C.c_double * 3),
                ("status", C.c_int32), ("flags", C.c_int32)]


CAND_DTYPE = np.dtype([("rank", np.uint64), ("nll", np.float64), ("mu", np.float64, (3,)),
                       ("status", np.int32), ("flags", np.int32)])
assert CAND_DTYPE.itemsize == C.sizeof(Cand)


class SearchStats(C.Structure):
    _fields_ = [("n_candidates", C.c_uint64), ("n_solved", C.c_uint64), ("n_stripes", C.c_uint64),
                ("n_rescanned", C.c_uint64), ("min_nll", C.c_double), ("newton_iters", C.c_double),
                ("ms_pass1", C.c_float), ("ms_pass2", C.c_float),
                ("n_exact", C.c_uint64), ("n_bound_violations", C.c_uint64)]


# every symbol include/theta_b200.h declares: (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "theta_init": (C.c_int, [C.c_int]),
    "theta_device_count": (C.c_int, []),
    "theta_shutdown": (None, []),
    "theta_last_error": (C.c_char_p, []),
    "theta_version": (C.c_char_p, []),
    "theta_fp64_peak": (C.c_int, [C.POINTER(C.c_double), _VP]),
    "theta_plan_create": (C.c_int, [C.POINTER(Problem), C.POINTER(_VP)]),
    "theta_plan_destroy": (None, [_VP]),
    "theta_plan_count": (C.c_int, [_VP, C.POINTER(C.c_uint64)]),
    "theta_plan_bounds": (C.c_int, [_VP, _VP, _VP]),
    "theta_plan_null_nll": (C.c_int, [_VP, C.POINTER(C.c_double)]),
    "theta_n3_rows": (C.c_int, [_VP, _VP, _VP, C.POINTER(C.c_int32)]),
    "theta_n3_prepare": (C.c_int, [_VP, C.c_int64, C.c_int64, C.POINTER(C.c_uint64)]),
    "theta_n3_item_count": (C.c_int, [_VP, C.POINTER(C.c_uint64)]),
    "theta_n3_item_weights": (C.c_int, [_VP, _VP]),
    "theta_n3_truncated": (C.c_int, [_VP, C.POINTER(C.c_int32)]),
    "theta_search_n3_min": (C.c_int, [_VP, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.POINTER(C.c_double), _VP]),
    "theta_search_n3_window": (C.c_int, [_VP, C.c_double, C.c_double, _VP, C.c_uint64, C.POINTER(C.c_uint64), _VP, _VP]),
    "theta_enumerate_n3": (C.c_int, [_VP, C.c_uint64, C.c_uint64, _VP, C.c_uint64, C.POINTER(C.c_uint64)]),
    "theta_solve_rows_n3": (C.c_int, [_VP, _VP, C.c_int64, _VP, _VP, _VP, _VP, _VP, C.c_int32]),
    "theta_enumerate_n2": (C.c_int, [_VP, C.c_uint64, C.c_int64, _VP, C.c_int64]),
    "theta_enumerate_n2_dev": (C.c_int, [_VP, C.c_uint64, C.c_int64, _VP, C.c_int64, _VP]),
    "theta_tile_pitch": (C.c_int64, [C.c_int32]),
    "theta_search_n2_pass1": (C.c_int, [_VP, C.c_uint64, C.c_uint64, C.POINTER(C.c_double), _VP]),
    "theta_search_n2_pass1_begin": (C.c_int, [_VP, C.c_uint64, C.c_uint64, _VP, _VP]),
    "theta_search_n2_pass1_end": (C.c_int, [_VP, C.POINTER(C.c_double)]),
    "theta_search_n2_pass2": (C.c_int, [_VP, C.c_double, C.c_double, _VP, C.c_uint64, C.POINTER(C.c_uint64), _VP]),
    "theta_search_n2_pass2_dev": (C.c_int, [_VP, _VP, C.c_double, _VP, C.c_uint64, C.POINTER(C.c_uint64),
                                            C.POINTER(C.c_double), _VP]),
    "theta_search_n2_pass2_enqueue": (C.c_int, [_VP, _VP, C.c_double, _VP, C.c_uint64, _VP]),
    "theta_search_n2_async_done": (C.c_int, [_VP, _VP]),
    "theta_search_n2_graph_replay": (C.c_int, [_VP, C.c_uint64, C.c_uint64, _VP]),
    "theta_filter_records": (C.c_int, [_VP, C.c_uint64, C.c_double, C.c_double, C.POINTER(C.c_uint64)]),
    "theta_search_stats_get": (C.c_int, [_VP, C.POINTER(SearchStats)]),
    "theta_search_n2": (C.c_int, [_VP, C.c_uint64, C.c_uint64, _VP, C.c_uint64, C.POINTER(C.c_uint64)]),
    "theta_optimize_n2": (C.c_int, [C.POINTER(Problem), C.c_uint64, _VP, _VP, C.c_uint64, C.POINTER(C.c_uint64), _VP, _VP, _VP,
                                    C.POINTER(SearchStats), C.POINTER(C.c_uint64)]),
    "theta_eval_ranks_n2": (C.c_int, [_VP, _VP, C.c_int64, _VP, _VP]),
    "theta_columns_of_ranks": (C.c_int, [_VP, _VP, C.c_int64, _VP]),
    "theta_nccl_unique_id": (C.c_int, [_VP]),
    "theta_comm_create": (C.c_int, [C.c_int32, C.c_int32, _VP, C.POINTER(_VP)]),
    "theta_comm_destroy": (None, [_VP]),
    "theta_comm_ipc_handle": (C.c_int, [_VP, _VP]),
    "theta_comm_open_peers": (C.c_int, [_VP, _VP]),
    "theta_comm_peer_exchange": (C.c_int, [_VP, C.c_int32]),
    "theta_search_n2_dist": (C.c_int, [_VP, _VP, C.c_uint64, C.c_uint64, _VP, C.c_uint64, C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_double), C.POINTER(C.c_uint64), _VP]),
    "theta_search_n3_dist": (C.c_int, [_VP, _VP, C.c_uint64, C.c_uint64, C.c_double, C.c_double, _VP, _VP, C.c_uint64,
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_double), _VP, _VP]),
    "theta_mgpu_create": (C.c_int, [C.POINTER(Problem), C.c_int32, _VP, C.POINTER(_VP)]),
    "theta_mgpu_destroy": (None, [_VP]),
    "theta_mgpu_devices": (C.c_int, [_VP]),
    "theta_mgpu_optimize": (C.c_int, [_VP, C.c_uint64, _VP, _VP, C.c_uint64, C.POINTER(C.c_uint64), _VP, _VP, _VP,
                                      C.POINTER(SearchStats), C.POINTER(C.c_uint64)]),
    "theta_mgpu_last_ms": (C.c_int, [_VP, _VP]),
    "theta_optimize": (C.c_int, [C.POINTER(Problem), C.c_int32, _VP, C.c_uint64, _VP, _VP, C.c_uint64, C.POINTER(C.c_uint64),
                                 _VP, _VP, _VP, C.POINTER(SearchStats), C.POINTER(C.c_uint64)]),
    "theta_solve_tiles_n2_dev": (C.c_int, [_VP, _VP, C.c_int64, C.c_int64, _VP, _VP, _VP, _VP]),
    "theta_solve_tiles_n2": (C.c_int, [_VP, _VP, C.c_int64, C.c_int64, _VP, _VP, _VP]),
    "theta_nll_tiles_n2_dev": (C.c_int, [_VP, _VP, C.c_int64, C.c_int64, _VP, _VP, _VP]),
    "theta_nll_tiles_n2": (C.c_int, [_VP, _VP, C.c_int64, C.c_int64, _VP, _VP]),
    "theta_nll_rows_n3_dev": (C.c_int, [_VP, _VP, C.c_int64, C.c_int64, _VP, _VP, _VP]),
    "theta_nll_rows_n3": (C.c_int, [_VP, _VP, C.c_int64, C.c_int64, _VP, _VP]),
    "theta_nll": (C.c_int, [C.c_int32, C.c_int32, _VP, _VP, _VP, C.POINTER(C.c_double), _VP, _VP]),
}

_lib = None


def load():
    """Load the shared library and bind every declared symbol (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ThetaError(ERR_CUDA, "%s not built: run `python -m theta_b200.build` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        msg = load().theta_last_error().decode("utf-8", "replace")
        if rc == ERR_EMPTY:
            raise EmptySpaceError(rc, msg)
        raise ThetaError(rc, msg)


_initialised = None


def init(device=None):
    """theta_init on `device` (default: LOCAL_RANK or 0). Raises ThetaError when no GPU is usable."""
    global _initialised
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _initialised == device:
        return
    check(load().theta_init(int(device)))
    _initialised = device


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _int_array(v, dtype):
    """int(x) for every element (the reference's lists hold Python ints; strings / floats are converted the same way)."""
    try:
        a = np.asarray(v)
        if a.dtype.kind in "iu":
            return np.ascontiguousarray(a, dtype=dtype)
    except (TypeError, ValueError):
        pass
    return np.ascontiguousarray(np.asarray([int(x) for x in v], dtype=dtype))


class Plan(object):
    """Owns a theta_plan: device tables of one problem at the do_optimization* boundary."""

    def __init__(self, n, r, rN, lower, upper, tau=2, max_normal=1.0, device=None, min_normal=0.0, tiles_only=False):
        init(device)
        self.n = int(n)
        self.m = len(r)
        self.tau = int(tau)
        self.max_normal = float(max_normal)
        self._r = _int_array(r, np.int64)
        self._rN = _int_array(rN, np.int64)
        self._lo = _int_array(lower, np.int32)
        self._up = _int_array(upper, np.int32)
        if not (len(self._rN) == len(self._lo) == len(self._up) == self.m):
            raise ValueError("r, rN, lower, upper must have the same length")
        prob = Problem(self.n, self.m, self.tau, 1 if tiles_only else 0,
                       self._r.ctypes.data_as(C.POINTER(C.c_int64)), self._rN.ctypes.data_as(C.POINTER(C.c_int64)),
                       self._lo.ctypes.data_as(C.POINTER(C.c_int32)), self._up.ctypes.data_as(C.POINTER(C.c_int32)),
                       self.max_normal, float(min_normal))
        h = C.c_void_p()
        check(load().theta_plan_create(C.byref(prob), C.byref(h)))
        self._h = h
        self.pitch = int(load().theta_tile_pitch(self.m))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:  # (module globals may already be gone at interpreter shutdown)
            try:
                _lib.theta_plan_destroy(h)
            except Exception:
                pass

    __del__ = close

    @property
    def handle(self):
        return self._h

    def table_bytes(self):
        """Size of the n=2 table blob copied to the device by theta_plan_create (TableLayout, theta_n2.cuh)."""
        K = max(1, int(self._up.max()))
        KT = next(o for o in (2, 3, 4, 5, 6, 7, 8, 10, 12, 15) if K <= o)
        o = 8 * (self.m + 1) * 2 + 8 * KT * (self.m + 2) + 4 * (self.m + 1) + 4 * KT * 2
        return (o + 15) & ~15

    def count(self):
        c = C.c_uint64()
        check(load().theta_plan_count(self._h, C.byref(c)))
        return int(c.value)

    def bounds(self):
        lo = np.zeros(self.m, dtype=np.int32)
        up = np.zeros(self.m, dtype=np.int32)
        check(load().theta_plan_bounds(self._h, ptr(lo), ptr(up)))
        return lo, up

    def null_nll(self):
        g = C.c_double()
        check(load().theta_plan_null_nll(self._h, C.byref(g)))
        return float(g.value)

    # ---- n = 3 ----
    def rows(self):
        a = np.zeros(64, dtype=np.int32)
        b = np.zeros(64, dtype=np.int32)
        n = C.c_int32()
        check(load().theta_n3_rows(self._h, ptr(a), ptr(b), C.byref(n)))
        return [(int(x), int(y)) for x, y in zip(a[:n.value], b[:n.value])]

    def prepare(self, leaf_cap=0, max_items=1 << 22):
        """Cut the DFS forest into work items (theta_n3_prepare).  The list is a function of the plan and the two
        arguments only, so a repeated call with the same arguments keeps the list that is already on the device."""
        if getattr(self, "_prepared", None) == (int(leaf_cap), int(max_items)):
            return self.n_items
        n = C.c_uint64()
        check(load().theta_n3_prepare(self._h, int(leaf_cap), int(max_items), C.byref(n)))
        self.n_items = int(n.value)
        t = C.c_int32()
        check(load().theta_n3_truncated(self._h, C.byref(t)))
        self.truncated = bool(t.value)
        self._prepared = (int(leaf_cap), int(max_items))
        return self.n_items

    def columns_of_ranks(self, ranks):
        ranks = np.ascontiguousarray(np.asarray(ranks, dtype=np.uint64))
        cols = np.zeros((len(ranks), self.m), dtype=np.int8)
        if len(ranks):
            check(load().theta_columns_of_ranks(self._h, ptr(ranks), len(ranks), ptr(cols)))
        return cols

    def item_weights(self):
        w = np.zeros(self.n_items, dtype=np.float64)
        if self.n_items:
            check(load().theta_n3_item_weights(self._h, ptr(w)))
        return w

    def search_min_n3(self, item_begin, item_end, t_seed=float("inf"), window=0.0, stream=None):
        """Exact minimum NLL of the items [item_begin, item_end) (theta_search_n3_min)."""
        g = C.c_double()
        check(load().theta_search_n3_min(self._h, int(item_begin), int(item_end), float(t_seed), float(window),
                                         C.byref(g), stream))
        return float(g.value)

    def window_n3(self, global_min, window=0.0, stream=None):
        """Records + row-index paths of every candidate within `window` of global_min (theta_search_n3_window)."""
        cap = 1 << 12
        while True:
            buf = np.zeros(cap, dtype=CAND_DTYPE)
            paths = np.zeros((cap, self.m), dtype=np.uint8)
            n = C.c_uint64()
            rc = load().theta_search_n3_window(self._h, float(global_min), float(window), ptr(buf), cap, C.byref(n),
                                               ptr(paths), stream)
            if rc == ERR_OVERFLOW and n.value > cap:
                cap = int(n.value) + 16
                continue
            check(rc)
            return buf[:int(n.value)], paths[:int(n.value)]

    def enumerate_n3(self, item_begin=0, item_end=None):
        """Row-index strings [count, m] of every candidate of the given items, in enumeration order."""
        if item_end is None:
            item_end = self.n_items
        cap = 1 << 16
        while True:
            paths = np.zeros((cap, self.m), dtype=np.uint8)
            n = C.c_uint64()
            rc = load().theta_enumerate_n3(self._h, int(item_begin), int(item_end), ptr(paths), cap, C.byref(n))
            if rc == ERR_OVERFLOW and n.value > cap:
                cap = int(n.value) + 16
                continue
            check(rc)
            return paths[:int(n.value)]

    def solve_rows(self, paths, want_vals=False, mode=0):
        """Optimizer.solve (n=3) on row-index strings [count, m] -> (mu[count,3], muhat, nll, status, vals).
        mode: 0 auto (reference-order solver for rank-deficient candidates), 1 reference-order always, 2 grouped."""
        paths = np.ascontiguousarray(np.asarray(paths, dtype=np.uint8))
        cnt = paths.shape[0]
        mu = np.zeros((cnt, 3))
        mh = np.zeros((cnt, 3))
        nll = np.zeros(cnt)
        st = np.zeros(cnt, dtype=np.int32)
        vals = np.zeros((cnt, self.m)) if want_vals else None
        if cnt:
            check(load().theta_solve_rows_n3(self._h, ptr(paths), cnt, ptr(mu), ptr(mh), ptr(nll), ptr(st),
                                             ptr(vals) if want_vals else None, int(mode)))
        return mu, mh, nll, st, vals

    def nll_rows(self, paths, mu):
        """Optimizer.L3 at given mu[count,3] on row-index strings [count, m] (streaming NLL kernel, n=3)."""
        paths = np.asarray(paths, dtype=np.uint8)
        cnt = paths.shape[0]
        t = np.zeros((cnt, self.pitch), dtype=np.uint8)
        t[:, :self.m] = paths
        mu = np.ascontiguousarray(np.asarray(mu, dtype=np.float64).reshape(cnt, 3))
        nll = np.zeros(cnt)
        if cnt:
            check(load().theta_nll_rows_n3(self._h, ptr(t), self.pitch, cnt, ptr(mu), ptr(nll)))
        return nll

    def enumerate(self, rank_begin, count):
        """int8 [count, m] tumour columns of consecutive ranks (Enumerator.generate_next_C order)."""
        out = np.zeros((int(count), self.pitch), dtype=np.int8)
        check(load().theta_enumerate_n2(self._h, int(rank_begin), int(count), ptr(out), self.pitch))
        return out[:, :self.m]

    def pass1(self, rank_begin, rank_end, stream=None):
        g = C.c_double()
        check(load().theta_search_n2_pass1(self._h, int(rank_begin), int(rank_end), C.byref(g), stream))
        return float(g.value)

    def pass1_begin(self, rank_begin, rank_end, local_min_dev=None, stream=None):
        """Asynchronous pass 1; local_min_dev: device address of a double that receives the shard minimum."""
        check(load().theta_search_n2_pass1_begin(self._h, int(rank_begin), int(rank_end), local_min_dev, stream))

    def pass1_end(self):
        g = C.c_double()
        check(load().theta_search_n2_pass1_end(self._h, C.byref(g)))
        return float(g.value)

    def pass2(self, global_min, prior_min=float("inf"), stream=None):
        cap = 1 << 10
        while True:
            buf = np.empty(cap, dtype=CAND_DTYPE)
            n = C.c_uint64()
            rc = load().theta_search_n2_pass2(self._h, float(global_min), float(prior_min), ptr(buf), cap, C.byref(n), stream)
            if rc == ERR_OVERFLOW and n.value > cap:
                cap = int(n.value) + 16
                continue
            check(rc)
            return buf[:int(n.value)]

    def pass2_dev(self, global_min_dev=None, prior_min=float("inf"), stream=None):
        """Pass 2 with the global minimum taken from device memory (None: this shard's own); also completes a
        pending pass1_begin.  Returns (records, global_min)."""
        cap = 1 << 10
        while True:
            buf = np.empty(cap, dtype=CAND_DTYPE)
            n = C.c_uint64()
            g = C.c_double()
            rc = load().theta_search_n2_pass2_dev(self._h, global_min_dev, float(prior_min), ptr(buf), cap, C.byref(n),
                                                  C.byref(g), stream)
            if rc == ERR_OVERFLOW and n.value > cap:
                cap = int(n.value) + 16
                continue
            check(rc)
            return buf[:int(n.value)], float(g.value)

    def pass2_enqueue(self, global_min_dev, out_dev, cap, prior_min=float("inf"), stream=None):
        """Asynchronous pass 2 into a device buffer [u64 status[32] | records[cap]] (see include/theta_b200.h)."""
        check(load().theta_search_n2_pass2_enqueue(self._h, global_min_dev, float(prior_min), out_dev, int(cap), stream))

    def async_done(self, status_words):
        st = np.ascontiguousarray(status_words, dtype=np.uint64)
        check(load().theta_search_n2_async_done(self._h, ptr(st)))

    def stats(self):
        s = SearchStats()
        check(load().theta_search_stats_get(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in SearchStats._fields_}

    def eval_ranks(self, ranks, want_vals=True):
        ranks = np.ascontiguousarray(np.asarray(ranks, dtype=np.uint64))
        out = np.zeros(len(ranks), dtype=CAND_DTYPE)
        vals = np.zeros((len(ranks), self.m), dtype=np.float64) if want_vals else None
        if len(ranks):
            check(load().theta_eval_ranks_n2(self._h, ptr(ranks), len(ranks), ptr(out), ptr(vals) if want_vals else None))
        return out, vals

    def _tiles(self, cols):
        cols = np.asarray(cols)
        if cols.ndim != 2 or cols.shape[1] != self.m:
            raise ValueError("cols must be [count, m]")
        t = np.zeros((cols.shape[0], self.pitch), dtype=np.int8)
        t[:, :self.m] = cols
        return t

    def solve_tiles(self, cols):
        """Optimizer.solve on explicit tumour columns [count, m] -> (mu0, nll, status)."""
        t = self._tiles(cols)
        cnt = t.shape[0]
        mu = np.zeros(cnt)
        nll = np.zeros(cnt)
        st = np.zeros(cnt, dtype=np.int32)
        if cnt:
            check(load().theta_solve_tiles_n2(self._h, ptr(t), self.pitch, cnt, ptr(mu), ptr(nll), ptr(st)))
        return mu, nll, st

    def nll_tiles(self, cols, mu):
        t = self._tiles(cols)
        cnt = t.shape[0]
        mu = np.ascontiguousarray(np.asarray(mu, dtype=np.float64))
        nll = np.zeros(cnt)
        if cnt:
            check(load().theta_nll_tiles_n2(self._h, ptr(t), self.pitch, cnt, ptr(mu), ptr(nll)))
        return nll


def optimize_n2(r, rN, lower, upper, tau=2, max_normal=1.0, rank_limit=None, cap_best=64):
    """theta_optimize_n2: the whole single-GPU n=2 search in one C call.
    Returns (lower, upper, best records [fold order], cols [n_best, m] int8, vals [n_best, m], stats dict, n_candidates)."""
    init()
    ra, rNa = _int_array(r, np.int64), _int_array(rN, np.int64)
    lo, up = _int_array(lower, np.int32), _int_array(upper, np.int32)
    m = len(ra)
    if not (len(rNa) == len(lo) == len(up) == m):
        raise ValueError("r, rN, lower, upper must have the same length")
    prob = Problem(2, m, int(tau), 0, ra.ctypes.data_as(C.POINTER(C.c_int64)), rNa.ctypes.data_as(C.POINTER(C.c_int64)),
                   lo.ctypes.data_as(C.POINTER(C.c_int32)), up.ctypes.data_as(C.POINTER(C.c_int32)), float(max_normal), 0.0)
    lo_out = np.empty(m, dtype=np.int32)
    up_out = np.empty(m, dtype=np.int32)
    while True:
        best = np.empty(cap_best, dtype=CAND_DTYPE)
        cols = np.empty((cap_best, m), dtype=np.int8)
        vals = np.empty((cap_best, m), dtype=np.float64)
        n = C.c_uint64()
        nc = C.c_uint64()
        st = SearchStats()
        rc = load().theta_optimize_n2(C.byref(prob), int(rank_limit or 0), ptr(lo_out), ptr(up_out), cap_best, C.byref(n),
                                      ptr(best), ptr(cols), ptr(vals), C.byref(st), C.byref(nc))
        if rc == ERR_OVERFLOW and n.value > cap_best:
            cap_best = int(n.value) + 8
            continue
        check(rc)
        k = int(n.value)
        return (lo_out, up_out, best[:k], cols[:k], vals[:k], {f: getattr(st, f) for f, _ in SearchStats._fields_},
                int(nc.value))


def device_count():
    return int(load().theta_device_count())


class Comm(object):
    """theta_comm: this rank's NCCL communicator + exchange buffers for the sharded search steps
    (theta_search_n2_dist / theta_search_n3_dist).  Under torch.distributed the NCCL id made on rank 0 travels
    through torch's own broadcast; world == 1 needs no id and no NCCL."""

    def __init__(self, rank=0, world=1, nccl_id=None, device=None):
        init(device)
        self.rank, self.world = int(rank), int(world)
        h = C.c_void_p()
        idbuf = None
        if self.world > 1:
            idbuf = np.ascontiguousarray(np.frombuffer(bytes(nccl_id), dtype=np.uint8))
            assert idbuf.size == 128
        check(load().theta_comm_create(self.rank, self.world, ptr(idbuf) if idbuf is not None else None, C.byref(h)))
        self._h = h

    @staticmethod
    def unique_id():
        buf = np.zeros(128, dtype=np.uint8)
        check(load().theta_nccl_unique_id(ptr(buf)))
        return buf.tobytes()

    @classmethod
    def from_torch(cls, dist):
        """Communicator over the ranks of an initialised torch.distributed NCCL group."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        c = cls(rank, world, bytes(t.cpu().numpy().tobytes()), device=torch.cuda.current_device())
        # exchange areas over NVLink peer memory (CUDA IPC): handles all-gathered once through torch
        mine = np.zeros(128, dtype=np.uint8)
        check(load().theta_comm_ipc_handle(c._h, ptr(mine)))
        hs = torch.zeros(world * 128, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(hs, torch.from_numpy(mine).to(dev))
        allh = np.ascontiguousarray(hs.cpu().numpy())
        check(load().theta_comm_open_peers(c._h, ptr(allh)))
        # all ranks or none: a rank that could not map a peer sends everybody back to the NCCL collectives
        ok = torch.tensor([load().theta_comm_peer_exchange(c._h, -1)], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        load().theta_comm_peer_exchange(c._h, int(ok.item()))
        return c

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            try:
                _lib.theta_comm_destroy(h)
            except Exception:
                pass

    __del__ = close

    def search_n2(self, plan, rank_begin, rank_end, stream=None):
        """One sharded n=2 step in a single C call: pass 1, all-reduce(min), pass 2, record all-gather.
        Returns (records of all shards as the fold replay needs them, global_min, solved candidates of all shards)."""
        cap = 1 << 10
        while True:
            buf = np.empty(cap, dtype=CAND_DTYPE)
            n = C.c_uint64()
            g = C.c_double()
            ns = C.c_uint64()
            rc = load().theta_search_n2_dist(plan.handle, self._h, int(rank_begin), int(rank_end), ptr(buf), cap, C.byref(n),
                                             C.byref(g), C.byref(ns), stream)
            if rc == ERR_OVERFLOW and n.value > cap:
                cap = int(n.value) + 16
                continue
            check(rc)
            return buf[:int(n.value)], float(g.value), int(ns.value)

    def search_n3(self, plan, item_begin, item_end, window, extra_min=float("inf"), stream=None):
        """One sharded n=3 step: returns (records, paths, global_min, (solved, candidates) over all shards)."""
        cap = 1 << 10
        while True:
            buf = np.empty(cap, dtype=CAND_DTYPE)
            paths = np.zeros((cap, plan.m), dtype=np.uint8)
            n = C.c_uint64()
            g = C.c_double()
            tot = np.zeros(2, dtype=np.uint64)
            rc = load().theta_search_n3_dist(plan.handle, self._h, int(item_begin), int(item_end), float(window),
                                             float(extra_min), ptr(buf), ptr(paths), cap, C.byref(n), C.byref(g), ptr(tot),
                                             stream)
            if rc == ERR_OVERFLOW and n.value > cap:
                cap = int(n.value) + 16
                continue
            check(rc)
            k = int(n.value)
            return buf[:k], paths[:k], float(g.value), (int(tot[0]), int(tot[1]))


class MultiGpu(object):
    """theta_mgpu: do_optimization on several GPUs from ONE process (one host thread + stream per GPU, NCCL
    all-reduce(min) in between) -- what RunTHetA's max_processes / --NUM_PROCESSES maps to."""

    def __init__(self, n, r, rN, lower, upper, tau=2, max_normal=1.0, n_devices=None, devices=None):
        load()
        self.n = int(n)
        self.m = len(r)
        self._r, self._rN = _int_array(r, np.int64), _int_array(rN, np.int64)
        self._lo, self._up = _int_array(lower, np.int32), _int_array(upper, np.int32)
        if devices is not None:
            devs = np.ascontiguousarray(np.asarray(devices, dtype=np.int32))
            n_devices = len(devs)
        else:
            devs = None
            n_devices = int(n_devices or device_count())
        prob = Problem(self.n, self.m, int(tau), 0, self._r.ctypes.data_as(C.POINTER(C.c_int64)),
                       self._rN.ctypes.data_as(C.POINTER(C.c_int64)), self._lo.ctypes.data_as(C.POINTER(C.c_int32)),
                       self._up.ctypes.data_as(C.POINTER(C.c_int32)), float(max_normal), 0.0)
        h = C.c_void_p()
        check(load().theta_mgpu_create(C.byref(prob), n_devices, ptr(devs) if devs is not None else None, C.byref(h)))
        self._h = h
        self.n_devices = n_devices

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            try:
                _lib.theta_mgpu_destroy(h)
            except Exception:
                pass

    __del__ = close

    def optimize(self, budget=0, cap_best=64):
        """Returns (lower, upper, best records [fold order], rows [n_best, m] uint8, vals [n_best, m], stats, n_candidates)."""
        m = self.m
        lo_out = np.empty(m, dtype=np.int32)
        up_out = np.empty(m, dtype=np.int32)
        while True:
            best = np.empty(cap_best, dtype=CAND_DTYPE)
            rows = np.empty((cap_best, m), dtype=np.uint8)
            vals = np.empty((cap_best, m), dtype=np.float64)
            n = C.c_uint64()
            nc = C.c_uint64()
            st = SearchStats()
            rc = load().theta_mgpu_optimize(self._h, int(budget or 0), ptr(lo_out), ptr(up_out), cap_best, C.byref(n),
                                            ptr(best), ptr(rows), ptr(vals), C.byref(st), C.byref(nc))
            if rc == ERR_OVERFLOW and n.value > cap_best:
                cap_best = int(n.value) + 8
                continue
            check(rc)
            k = int(n.value)
            return (lo_out, up_out, best[:k], rows[:k], vals[:k], {f: getattr(st, f) for f, _ in SearchStats._fields_},
                    int(nc.value))

    def last_ms(self):
        ms = np.zeros(self.n_devices, dtype=np.float32)
        check(load().theta_mgpu_last_ms(self._h, ptr(ms)))
        return ms


STATUS_WORDS = 32          # uint64 words of the device status block (theta_n2.cu)
ST_SOLVED, ST_EMITTED, ST_LIST, ST_LOCAL_MIN, ST_GLOBAL_MIN = 1, 8 + 3, 16, 17, 18


def filter_records(recs, near_thr, prior_min=float("inf")):
    """theta_filter_records: sort raw pass-2 records by rank, keep NaN / strict prefix minima / near-minimum ones."""
    recs = np.ascontiguousarray(recs)
    n = C.c_uint64()
    check(load().theta_filter_records(ptr(recs), len(recs), float(prior_min), float(near_thr), C.byref(n)))
    return recs[:int(n.value)]


def nll_single(mu, Cw, r):
    """theta_nll: masked NLL of one weighted matrix (CalcAllC.L2/L3 semantics). Returns (nll, vals, valid)."""
    init()
    Cw = np.ascontiguousarray(np.asarray(Cw, dtype=np.float64))
    m, n = Cw.shape
    mu = np.ascontiguousarray(np.asarray(mu, dtype=np.float64))
    rr = np.ascontiguousarray(np.asarray([int(v) for v in r], dtype=np.int64))
    nll = C.c_double()
    vals = np.zeros(m)
    valid = np.zeros(m, dtype=np.uint8)
    check(load().theta_nll(n, m, ptr(mu), ptr(Cw), ptr(rr), C.byref(nll), ptr(vals), ptr(valid)))
    return float(nll.value), vals, valid.astype(bool)


def fp64_peak_gflops(stream=None):
    init()
    g = C.c_double()
    check(load().theta_fp64_peak(C.byref(g), stream))
    return float(g.value)
