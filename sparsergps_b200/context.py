"""Context: one per GPU.  Owns the device-side state behind the C ABI (streams, scratch, resident shard)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class Context:
    def __init__(self, device: int = 0):
        self._lib = L.load()
        h = L.vp()
        L.check(self._lib.srgp_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.device = device
        self.n = 0
        self.d = 0

    def close(self):
        if getattr(self, "handle", None):
            self._lib.srgp_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- resident shard + fused evaluations -------------------------------------------------
    def set_data(self, xy, y, mu=None):
        """Upload this rank's rows: xy (n x d), y (n), mu (n or None = 0).  include/srgp.h: srgp_set_data."""
        xy = L.fmat(xy)
        y = L.fvec(y)
        n, d = xy.shape
        assert y.size == n
        mup = None
        if mu is not None:
            mu = L.fvec(np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), (n,)))
            mup = L.ptr(mu)
        L.check(self._lib.srgp_set_data(self.handle, L.ptr(xy), n, d, L.ptr(y), mup))
        self.n, self.d = n, d

    def set_data_dev(self, xy_dev, n, d, y_dev, mu_dev=None):
        L.check(self._lib.srgp_set_data_dev(self.handle, xy_dev, int(n), int(d), y_dev, mu_dev))
        self.n, self.d = int(n), int(d)

    def gauss_obj_grad(self, model, cov_fun, xu, sigma, l, tau, delta, want_grad=True):
        """One objective (+ gradient wrt log theta) evaluation on the resident shard.
        model: "vi" | "fic"; returns (obj, grad ndarray ordered sigma, l / l1..ld, tau)."""
        xu = L.fmat(xu)
        m, d = xu.shape
        assert d == self.d, "knots and data disagree on the input dimension"
        lv = L.fvec(l)
        p = (d + 2) if cov_fun == "ard" else 3
        obj = L.cd()
        grad = np.zeros(p)
        L.check(self._lib.srgp_gauss_obj_grad(self.handle, L.VI if model == "vi" else L.FIC, L.KERNELS[cov_fun],
                                              L.ptr(xu), m, float(sigma), L.ptr(lv), float(tau), float(delta),
                                              C.byref(obj), L.ptr(grad) if want_grad else None))
        return obj.value, (grad if want_grad else None)

    def gauss_obj_grad_knots(self, model, cov_fun, xu, sigma, l, tau, delta, knot_bounds=None, knot_opt=None):
        """Objective, gradient wrt log theta and the knot-location gradient in one evaluation.
        knot_bounds: d x 2 array [lb, ub] (None = transform FALSE); knot_opt: 0-based indices (None = all).
        Returns (obj, grad, knot_gradient (m*d, knot-major), trans_knot (m x d))."""
        xu = L.fmat(xu)
        m, d = xu.shape
        assert d == self.d, "knots and data disagree on the input dimension"
        lv = L.fvec(l)
        p = (d + 2) if cov_fun == "ard" else 3
        obj = L.cd()
        grad, kgrad, tk = np.zeros(p), np.zeros(m * d), np.zeros((m, d), order="F")
        lb = ub = None
        if knot_bounds is not None:
            kb = np.asarray(knot_bounds, dtype=np.float64).reshape(d, 2)
            lb, ub = L.fvec(kb[:, 0]), L.fvec(kb[:, 1])
        opt, n_opt = None, 0
        if knot_opt is not None:
            opt = np.ascontiguousarray(np.asarray(list(knot_opt), dtype=np.int32))
            n_opt = len(opt)
        L.check(self._lib.srgp_gauss_obj_grad_knots(
            self.handle, L.VI if model == "vi" else L.FIC, L.KERNELS[cov_fun], L.ptr(xu), m, float(sigma), L.ptr(lv),
            float(tau), float(delta), L.ptr(lb) if lb is not None else None, L.ptr(ub) if ub is not None else None,
            opt.ctypes.data_as(C.POINTER(C.c_int)) if opt is not None and n_opt else None, n_opt, C.byref(obj),
            L.ptr(grad), L.ptr(kgrad), L.ptr(tk)))
        return obj.value, grad, kgrad, tk

    def gauss_fit(self, model, cov_fun, xu, sigma, l, tau, delta, opt=None, opt_theta=True, opt_knots=False,
                  knot_bounds=None, knot_opt=None):
        """norm_grad_ascent_vi / norm_grad_ascent on the resident shard (one fused evaluation per iteration).
        opt: dict with the reference's option names (optim_method, decay, epsilon, eta, learn_rate, maxit, obj_tol,
        grad_tol).  Returns dict(sigma, l, tau, xu, iter, obj_fun, cov_par_history, grad)."""
        o = {"optim_method": "adadelta", "decay": 0.95, "epsilon": 1e-6, "learn_rate": 1e-2, "eta": 1e3, "maxit": 1000,
             "obj_tol": 1e-3, "grad_tol": float("inf")}
        o.update({k: v for k, v in (opt or {}).items() if k in o})
        xu = np.array(L.fmat(xu), order="F", copy=True)
        m, d = xu.shape
        assert d == self.d
        nl = d if cov_fun == "ard" else 1
        lv = np.array(np.broadcast_to(np.asarray(l, dtype=np.float64).reshape(-1), (nl,)), copy=True)
        p = nl + 2
        fo = L.FitOpt({"adadelta": L.OPT_ADADELTA, "ga": L.OPT_GA}[o["optim_method"]], o["decay"], o["epsilon"],
                      o["eta"], o["learn_rate"], int(o["maxit"]), o["obj_tol"], o["grad_tol"], int(bool(opt_theta)),
                      int(bool(opt_knots)))
        lb = ub = None
        if knot_bounds is not None:
            kb = np.asarray(knot_bounds, dtype=np.float64).reshape(d, 2)
            lb, ub = L.fvec(kb[:, 0]), L.fvec(kb[:, 1])
        ko, n_opt = None, 0
        if knot_opt is not None:
            ko = np.ascontiguousarray(np.asarray(list(knot_opt), dtype=np.int32))
            n_opt = len(ko)
        sg, ta, it = L.cd(float(sigma)), L.cd(float(tau)), C.c_int(0)
        maxit = int(o["maxit"])
        obj_hist, par_hist, grad_hist = np.full(maxit, np.nan), np.full((maxit, p), np.nan), np.full((maxit, p), np.nan)
        L.check(self._lib.srgp_gauss_fit(
            self.handle, L.VI if model == "vi" else L.FIC, L.KERNELS[cov_fun], L.ptr(xu), m, C.byref(sg), L.ptr(lv),
            C.byref(ta), float(delta), C.byref(fo), L.ptr(lb) if lb is not None else None,
            L.ptr(ub) if ub is not None else None, ko.ctypes.data_as(C.POINTER(C.c_int)) if n_opt else None, n_opt,
            C.byref(it), L.ptr(obj_hist), L.ptr(par_hist), L.ptr(grad_hist)))
        k = it.value
        return {"sigma": sg.value, "l": lv, "tau": ta.value, "xu": xu, "iter": k, "obj_fun": obj_hist[:k],
                "cov_par_history": par_hist[:k], "grad": grad_hist[:k]}

    def oat_scores(self, model, cov_fun, xu, cand, sigma, l, tau, delta):
        """Objective with each candidate row appended to the knots (theta fixed) on the resident shard.
        Returns (objective with the knots alone, scores ndarray; NaN = that candidate's Cholesky failed)."""
        xu, cand = L.fmat(xu), L.fmat(cand)
        assert xu.shape[1] == self.d and cand.shape[1] == self.d
        lv = L.fvec(l)
        obj0 = L.cd()
        scores = np.zeros(cand.shape[0])
        L.check(self._lib.srgp_oat_scores(self.handle, L.VI if model == "vi" else L.FIC, L.KERNELS[cov_fun], L.ptr(xu),
                                          xu.shape[0], L.ptr(cand), cand.shape[0], float(sigma), L.ptr(lv), float(tau),
                                          float(delta), C.byref(obj0), L.ptr(scores)))
        return obj0.value, scores

    def gauss_obj_grad_host(self, model, cov_fun, xy, y, mu, xu, sigma, l, tau, delta, want_grad=True):
        """One-shot call with the reference's argument list (uploads xy / y / mu inside the call)."""
        xy, y, xu, lv = L.fmat(xy), L.fvec(y), L.fmat(xu), L.fvec(l)
        n, d = xy.shape
        mup = None
        if mu is not None:
            mu = L.fvec(np.broadcast_to(np.asarray(mu, dtype=np.float64).reshape(-1), (n,)))
            mup = L.ptr(mu)
        p = (d + 2) if cov_fun == "ard" else 3
        obj = L.cd()
        grad = np.zeros(p)
        L.check(self._lib.srgp_gauss_obj_grad_host(self.handle, L.VI if model == "vi" else L.FIC,
                                                   L.KERNELS[cov_fun], L.ptr(xy), n, d, L.ptr(y), mup, L.ptr(xu),
                                                   xu.shape[0], float(sigma), L.ptr(lv), float(tau), float(delta),
                                                   C.byref(obj), L.ptr(grad) if want_grad else None))
        self.n, self.d = n, d
        return obj.value, (grad if want_grad else None)

    # ---- multi-GPU ---------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(L.UNIQUE_ID_BYTES)
        L.check(L.load().srgp_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, world: int, rank: int, unique_id: bytes):
        L.check(self._lib.srgp_comm_init(self.handle, int(world), int(rank), unique_id))

    # ---- instrumentation -------------------------------------------------------------------
    def sync(self):
        L.check(self._lib.srgp_ctx_sync(self.handle))

    def timer_start(self):
        L.check(self._lib.srgp_timer_start(self.handle))

    def timer_stop_ms(self) -> float:
        ms = L.cd()
        L.check(self._lib.srgp_timer_stop_ms(self.handle, C.byref(ms)))
        return ms.value

    def prof_enable(self, on=True):
        L.check(self._lib.srgp_prof_enable(self.handle, int(bool(on))))

    def prof_reset(self):
        L.check(self._lib.srgp_prof_reset(self.handle))

    def prof_get(self, name):
        n, ms = L.i64(), L.cd()
        L.check(self._lib.srgp_prof_get(self.handle, L.PROF[name], C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def launch_count(self) -> int:
        return int(self._lib.srgp_launch_count(self.handle))

    def flush_l2(self):
        L.check(self._lib.srgp_flush_l2(self.handle))

    def i8_slices(self) -> int:
        return int(self._lib.srgp_i8_slices())

    def probe_i8_peak(self, iters=4000):
        """(INT8 TOP/s, cycles per 128x128x32 MMA) of the resident-operand issue-rate probe (csrc/probe.cu)."""
        tops, cyc = C.c_double(), C.c_double()
        L.check(self._lib.srgp_probe_i8_peak(self.handle, int(iters), C.byref(tops), C.byref(cyc)))
        return tops.value, cyc.value

    # ---- raw device memory (bench / tests) --------------------------------------------------
    def dev_alloc(self, nbytes: int):
        p = L.vp()
        L.check(self._lib.srgp_dev_alloc(self.handle, int(nbytes), C.byref(p)))
        return p

    def dev_free(self, p):
        L.check(self._lib.srgp_dev_free(self.handle, p))

    def h2d(self, dev, arr: np.ndarray):
        L.check(self._lib.srgp_memcpy_h2d(self.handle, dev, arr.ctypes.data_as(L.vp), arr.nbytes))

    def d2h(self, arr: np.ndarray, dev):
        L.check(self._lib.srgp_memcpy_d2h(self.handle, arr.ctypes.data_as(L.vp), dev, arr.nbytes))

    def fill_normal(self, dev, n, seed, mean=0.0, sd=1.0):
        L.check(self._lib.srgp_fill_normal_dev(self.handle, dev, int(n), int(seed), float(mean), float(sd)))


_default = None


def default_context() -> Context:
    """Process-wide context on cuda:0 used by the reference-shaped functional API."""
    global _default
    if _default is None:
        _default = Context(0)
    return _default
