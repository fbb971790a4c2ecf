"""numpy-level wrapper of the pmcgpu_* C-ABI, used by the tests and bench.py.  All compute happens in libpmcb200.so."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .lib import ALLREDUCE_FN, PmcError, StepResultC, load_library


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


@dataclass
class StepResult:
    perplexity: float
    ess: float
    logSum: float
    ln_evidence: float
    maxW: float
    maxR: float
    nok: int
    n_reject: int
    retry: int
    n_alive: int
    used_precise_path: int
    wght: np.ndarray
    mean: np.ndarray
    L: np.ndarray
    detL: np.ndarray
    alive: np.ndarray


def host_cholesky(cov):
    """mvdens_cholesky_decomp on the host side of the library (GSL-1.14 semantics): returns (L with mirrored upper, detL)."""
    L = load_library()
    a = _f64(cov).copy()
    det = C.c_double()
    rc = L.pmcgpu_cholesky(a.shape[0], _p(a), C.byref(det))
    if rc:
        raise PmcError(rc, "Cholesky decomposition failed")
    return a, det.value


class PmcGpu:
    """One GPU context = one shard of the PMC sample."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.pmcgpu_create(device, C.byref(h))
        if rc:
            raise PmcError(rc, "pmcgpu_create failed (no CUDA device?)")
        self.h = h
        self.K = self.d = 0
        self.N = 0
        self._cb = None

    def close(self):
        if self.h:
            self.lib.pmcgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise PmcError(rc, self.lib.pmcgpu_last_error(self.h).decode())

    def set_option(self, name, value):
        self._ck(self.lib.pmcgpu_set_option(self.h, name.encode(), float(value)))

    def get_timing(self):
        ms, n, tot = C.c_double(), C.c_long(), C.c_long()
        self._ck(self.lib.pmcgpu_get_timing(self.h, C.byref(ms), C.byref(n), C.byref(tot)))
        return ms.value, n.value, tot.value

    def get_draw_timing(self):
        ms = C.c_double()
        self._ck(self.lib.pmcgpu_get_draw_timing(self.h, C.byref(ms)))
        return ms.value

    def get_stats_timing(self):
        ms = C.c_double()
        self._ck(self.lib.pmcgpu_get_stats_timing(self.h, C.byref(ms)))
        return ms.value

    @property
    def stream(self):
        return self.lib.pmcgpu_stream(self.h)

    # ---- model --------------------------------------------------------------------------------------
    def set_proposal(self, wght, mean, L, detL=None, df=None):
        wght, mean, L = _f64(wght), _f64(mean), _f64(L)
        self.K, self.d = mean.shape
        detL = None if detL is None else _f64(detL)
        df = None if df is None else np.ascontiguousarray(df, np.int32)
        self._ck(self.lib.pmcgpu_set_proposal(self.h, self.K, self.d, _p(wght), _p(mean), _p(L), _p(detL), _p(df)))

    def set_proposal_cov(self, wght, mean, cov, df=None):
        """Convenience: factorise covariances with the library's host Cholesky, like the reference does lazily."""
        Ls, dets = zip(*[host_cholesky(c) for c in cov])
        self.set_proposal(wght, mean, np.array(Ls), np.array(dets), df)

    def set_workload(self, wl):
        self.set_proposal_cov(wl["wght"], wl["mean"], wl["cov"], wl["df"])
        t = wl["target"]
        if t["kind"] == 1:
            self._ck(self.lib.pmcgpu_set_target_banana(self.h, t["d"], t["b"], t["sig1"]))
        elif t["kind"] == 5:
            self.set_target_combine(t["d"], t["terms"])
        else:
            Ls, dets = zip(*[host_cholesky(c) for c in t["cov"]])
            w, m, Lc, dt = _f64(t["wght"]), _f64(t["mean"]), _f64(np.array(Ls)), _f64(np.array(dets))
            df = np.ascontiguousarray(t["df"], np.int32)
            self._ck(self.lib.pmcgpu_set_target_mix(self.h, t["kind"], t["K"], t["d"], _p(w), _p(m), _p(Lc), _p(dt), _p(df)))
        self.set_box(wl["faces"])

    def get_info(self, name):
        v = C.c_double()
        self._ck(self.lib.pmcgpu_get_info(self.h, name.encode(), C.byref(v)))
        return v.value

    def set_target_combine(self, d, terms):
        """Sum of densities on index-mapped coordinates (combine_lkl / add_gaussian_prior): terms = list of dicts with
        kind (1 banana: b, sig1; 2 mvdens / 3 mix: wght, mean, cov, df), d and dims."""
        from .lib import CombineTermC
        arr = (CombineTermC * len(terms))()
        keep = []
        for i, t in enumerate(terms):
            dims = np.ascontiguousarray(t["dims"], np.int32)
            keep.append(dims)
            arr[i].kind, arr[i].ndim, arr[i].dim_idx = t["kind"], len(dims), dims.ctypes.data
            if t["kind"] == 1:
                arr[i].b, arr[i].sig1 = t["b"], t["sig1"]
            else:
                Ls, dets = zip(*[host_cholesky(c) for c in t["cov"]])
                w, m, Lc, dt = _f64(t["wght"]), _f64(t["mean"]), _f64(np.array(Ls)), _f64(np.array(dets))
                df = np.ascontiguousarray(t["df"], np.int32)
                keep += [w, m, Lc, dt, df]
                arr[i].K = len(w)
                arr[i].wght, arr[i].mean, arr[i].L, arr[i].detL, arr[i].df = (a.ctypes.data for a in (w, m, Lc, dt, df))
        self._ck(self.lib.pmcgpu_set_target_combine(self.h, d, len(terms), C.addressof(arr)))

    def set_target_banana(self, d, b, sig1):
        self._ck(self.lib.pmcgpu_set_target_banana(self.h, d, b, sig1))

    def set_target_defaults(self, is_def, value):
        if is_def is None:
            self._ck(self.lib.pmcgpu_set_target_defaults(self.h, 0, None, None))
            return
        i = np.ascontiguousarray(is_def, np.int32)
        v = _f64(value)
        self._ck(self.lib.pmcgpu_set_target_defaults(self.h, len(i), _p(i), _p(v)))

    def set_target_host(self, d):
        self._ck(self.lib.pmcgpu_set_target_host(self.h, d))

    def set_box(self, faces):
        if faces is None:
            self._ck(self.lib.pmcgpu_set_box(self.h, self.d, 0, None))
        else:
            f = _f64(faces)
            self._ck(self.lib.pmcgpu_set_box(self.h, f.shape[1] - 1, f.shape[0], _p(f)))

    def set_band_limit(self, bl):
        b = None if bl is None else np.ascontiguousarray(bl, np.uint64)
        self._ck(self.lib.pmcgpu_set_band_limit(self.h, _p(b)))

    # ---- store --------------------------------------------------------------------------------------
    def resize(self, N, d=None):
        self.N = int(N)
        self._ck(self.lib.pmcgpu_resize(self.h, self.N, d or self.d))

    def upload(self, X=None, indices=None, weights=None, log_rho=None, flg=None):
        X = None if X is None else _f64(X)
        indices = None if indices is None else np.ascontiguousarray(indices, np.uint64)
        weights = None if weights is None else _f64(weights)
        log_rho = None if log_rho is None else _f64(log_rho)
        flg = None if flg is None else np.ascontiguousarray(flg, np.int16)
        self._ck(self.lib.pmcgpu_upload(self.h, _p(X), _p(indices), _p(weights), _p(log_rho), _p(flg)))

    def download(self, X=True, indices=True, weights=True, log_rho=True, flg=True):
        out = {}
        if X:
            out["X"] = np.empty((self.N, self.d))
        if indices:
            out["indices"] = np.empty(self.N, np.uint64)
        if weights:
            out["weights"] = np.empty(self.N)
        if log_rho:
            out["log_rho"] = np.empty(self.N)
        if flg:
            out["flg"] = np.empty(self.N, np.int16)
        self._ck(self.lib.pmcgpu_download(self.h, _p(out.get("X")), _p(out.get("indices")), _p(out.get("weights")),
                                          _p(out.get("log_rho")), _p(out.get("flg"))))
        return out

    # ---- phases -------------------------------------------------------------------------------------
    def set_host_sink(self, X=None, indices=None, weights=None, log_rho=None, flg=None):
        """Host arrays pmc_step fills while the iteration runs (D2H overlapped with the kernels); all None detaches.
        The caller keeps the arrays alive (and ideally page-locked) while the sink is attached."""
        self._sink = (X, indices, weights, log_rho, flg)
        self._ck(self.lib.pmcgpu_set_host_sink(self.h, _p(X), _p(indices), _p(weights), _p(log_rho), _p(flg)))

    def draw(self, seed, it=0, first_index=0):
        nrej = C.c_long()
        self._ck(self.lib.pmcgpu_draw(self.h, seed, it, first_index, C.byref(nrej)))
        return nrej.value

    def weights(self, t=1.0, first_index=0):
        nok, mw, mr = C.c_long(), C.c_double(), C.c_double()
        self._ck(self.lib.pmcgpu_weights(self.h, t, first_index, C.byref(nok), C.byref(mw), C.byref(mr)))
        return nok.value, mw.value, mr.value

    def weights_host_post(self, post, ok=None, t=1.0, first_index=0):
        post = _f64(post)
        ok = None if ok is None else np.ascontiguousarray(ok, np.int16)
        nok, mw, mr = C.c_long(), C.c_double(), C.c_double()
        self._ck(self.lib.pmcgpu_weights_host_post(self.h, _p(post), _p(ok), t, first_index, C.byref(nok),
                                                    C.byref(mw), C.byref(mr)))
        return nok.value, mw.value, mr.value

    def proposal_log_pdf(self, x):
        x = _f64(x)
        out = np.empty(len(x))
        self._ck(self.lib.pmcgpu_proposal_log_pdf(self.h, len(x), _p(x), _p(out)))
        return out

    def target_log_pdf(self, x):
        x = _f64(x)
        out = np.empty(len(x))
        self._ck(self.lib.pmcgpu_target_log_pdf(self.h, len(x), _p(x), _p(out)))
        return out

    def normalize(self, maxW):
        ls, cnt = C.c_double(), C.c_long()
        self._ck(self.lib.pmcgpu_normalize(self.h, maxW, C.byref(ls), C.byref(cnt)))
        return ls.value, cnt.value

    def perplexity_ess(self):
        p, e, n = C.c_double(), C.c_double(), C.c_long()
        self._ck(self.lib.pmcgpu_perplexity_ess(self.h, C.byref(p), C.byref(e), C.byref(n)))
        return p.value, e.value, n.value

    def filter_histogram(self, nbins, clipleft, clipright, first_index=0):
        """filter_histogram (pmc.c:709-770) on the resident log-weights; returns the number of samples flagged out."""
        n = C.c_long()
        self._ck(self.lib.pmcgpu_filter_histogram(self.h, nbins, clipleft, clipright, first_index, C.byref(n)))
        return n.value

    def set_filter_histogram(self, nbins, clipleft=0, clipright=0):
        """Apply filter_histogram inside pmc_step / step_given (the pmc_filter hook, pmc.c:203-206); nbins=0: off."""
        self._ck(self.lib.pmcgpu_set_filter_histogram(self.h, nbins, clipleft, clipright))

    def weighted_mean(self, a):
        """mean_from_psim (pmc.c:1108-1120) of coordinate a."""
        m = C.c_double()
        self._ck(self.lib.pmcgpu_weighted_mean(self.h, a, C.byref(m)))
        return m.value

    def clip_weights(self, n):
        """clip_weights (pmc.c:1134-1188): returns (threshold, number clipped, sum of the remaining weights before rescaling)."""
        T, nc, S = C.c_double(), C.c_long(), C.c_double()
        self._ck(self.lib.pmcgpu_clip_weights(self.h, n, C.byref(T), C.byref(nc), C.byref(S)))
        return T.value, nc.value, S.value

    def sample_indices(self, seed, stream=0):
        """sample_from_pmc_simu (pmc.c:1585-1601): N indices drawn with probability proportional to the weights."""
        out = np.empty(self.N, np.uint64)
        self._ck(self.lib.pmcgpu_sample_indices(self.h, seed, stream, _p(out)))
        return out

    def resample_residual(self, seed, stream=0):
        """resample_residual (pmc.c:1605-1626): multinomial multiplicities of the residual weights."""
        out = np.empty(self.N, np.uint32)
        self._ck(self.lib.pmcgpu_resample_residual(self.h, seed, stream, _p(out)))
        return out

    def count_indices(self):
        c = np.zeros(self.K, np.uint64)
        self._ck(self.lib.pmcgpu_count_indices(self.h, _p(c)))
        return c

    def _outs(self):
        return (np.zeros(self.K), np.zeros((self.K, self.d)), np.zeros((self.K, self.d, self.d)), np.zeros(self.K),
                np.zeros(self.K, np.int32))

    def update_rb(self, maxR, N_total=0, retry=0, hard=False):
        r = C.c_int(retry)
        w, m, L, dt, al = self._outs()
        if hard:
            self._ck(self.lib.pmcgpu_update_hard(self.h, N_total, C.byref(r), _p(w), _p(m), _p(L), _p(dt), _p(al)))
        else:
            self._ck(self.lib.pmcgpu_update_rb(self.h, maxR, N_total, C.byref(r), _p(w), _p(m), _p(L), _p(dt), _p(al)))
        return dict(wght=w, mean=m, L=L, detL=dt, alive=al, retry=r.value)

    def ranindx(self, z):
        z = _f64(z)
        idx = np.empty(len(z), np.uint64)
        self._ck(self.lib.pmcgpu_ranindx(self.h, len(z), _p(z), _p(idx)))
        return idx

    def isinbox(self, x):
        x = _f64(x)
        out = np.empty(len(x), np.int32)
        self._ck(self.lib.pmcgpu_isinbox(self.h, len(x), _p(x), _p(out)))
        return out

    def _result(self, rs, r, w, m, L, dt, al):
        return StepResult(rs.perplexity, rs.ess, rs.logSum, rs.ln_evidence, rs.maxW, rs.maxR, rs.nok, rs.n_reject,
                          r.value, rs.n_alive, rs.used_precise_path, w, m, L, dt, al)

    def pmc_step(self, seed, it=0, first_index=0, N_total=0, do_update=1, retry=0):
        r = C.c_int(retry)
        rs = StepResultC()
        w, m, L, dt, al = self._outs()
        self._ck(self.lib.pmcgpu_pmc_step(self.h, seed, it, first_index, N_total, do_update, C.byref(r), C.byref(rs),
                                          _p(w), _p(m), _p(L), _p(dt), _p(al)))
        return self._result(rs, r, w, m, L, dt, al)

    def step_given(self, first_index=0, N_total=0, t=1.0, do_update=1, retry=0):
        r = C.c_int(retry)
        rs = StepResultC()
        w, m, L, dt, al = self._outs()
        self._ck(self.lib.pmcgpu_step_given(self.h, first_index, N_total, t, do_update, C.byref(r), C.byref(rs),
                                            _p(w), _p(m), _p(L), _p(dt), _p(al)))
        return self._result(rs, r, w, m, L, dt, al)

    def get_proposal(self):
        w, m, L, dt, al = self._outs()
        self._ck(self.lib.pmcgpu_get_proposal(self.h, _p(w), _p(m), _p(L), _p(dt), _p(al)))
        return dict(wght=w, mean=m, L=L, detL=dt, alive=al)

    # ---- multi-GPU ----------------------------------------------------------------------------------
    def comm_init_nccl(self, unique_id: bytes, rank, nranks):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.lib.pmcgpu_comm_init(self.h, buf, rank, nranks))

    def comm_set_callback(self, fn, rank, nranks):
        """fn(np_array_view, op) reduces in place over ranks (op 0 sum, 1 max)."""
        def tramp(_user, ptr, n, op):
            try:
                arr = np.ctypeslib.as_array(ptr, shape=(n,))
                fn(arr, op)
                return 0
            except Exception:  # pragma: no cover
                return 1
        self._cb = ALLREDUCE_FN(tramp)
        self._ck(self.lib.pmcgpu_comm_set_callback(self.h, self._cb, None, rank, nranks))


def nccl_unique_id() -> bytes:
    L = load_library()
    buf = C.create_string_buffer(128)
    rc = L.pmcgpu_nccl_unique_id(buf)
    if rc:
        raise PmcError(rc, "ncclGetUniqueId failed / NCCL not loadable")
    return buf.raw


def shard_range(N_total, rank, nranks):
    L = load_library()
    f, n = C.c_long(), C.c_long()
    L.pmcgpu_shard_range(N_total, rank, nranks, C.byref(f), C.byref(n))
    return f.value, n.value
