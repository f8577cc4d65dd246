"""ctypes binding of libqldpc_b200.so (C ABI declared in include/qldpc_b200.h).

There is deliberately no CPU fallback: if the shared library is missing or no CUDA device is
visible, every compute entry point raises.
"""
import ctypes as C
import hashlib
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QLDPC_B200_LIB", os.path.join(_HERE, "libqldpc_b200.so"))   # override: instrumented builds

QB_ALPHA_FIXED, QB_ALPHA_DYNAMIC, QB_ALPHA_SEQUENCE = 0, 1, 2
QB_PRECISION_F32, QB_PRECISION_HALF2 = 0, 1

_lib = None
_lock = threading.Lock()

c_void_p, c_int, c_i64, c_u64, c_double, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_float


class QbError(RuntimeError):
    pass


class DecodeConfig(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("alpha_mode", C.c_int32), ("alpha_z", C.c_float), ("alpha_x", C.c_float),
                ("alpha_seq_z_h", C.c_void_p), ("alpha_seq_x_h", C.c_void_p),
                ("alpha_len_z", C.c_int32), ("alpha_len_x", C.c_int32), ("clip_llr", C.c_float), ("use_osd", C.c_int32),
                ("precision", C.c_int32)]


class PipelineStats(C.Structure):
    _fields_ = [("ms_total", C.c_float), ("ms_sample", C.c_float), ("ms_minsum", C.c_float), ("ms_osd", C.c_float),
                ("ms_logical", C.c_float), ("kernel_launches", C.c_int64), ("edge_messages", C.c_int64),
                ("osd_sides", C.c_int64)]


EXPORTS = [
    "qb_last_error", "qb_device_count", "qb_version", "qb_decoder_create", "qb_decoder_set_prior",
    "qb_decoder_destroy", "qb_edge_layout_probe", "qb_minsum_batch", "qb_minsum_decode_host", "qb_minsum_core_host", "qb_alpha_messages_host", "qb_bp_decode_host",
    "qb_syndrome_check_host", "qb_osd0_batch", "qb_osd0_host", "qb_gf2_eliminate_host", "qb_sampler_create",
    "qb_sampler_destroy", "qb_syndrome_from_events", "qb_syndrome_from_events_host", "qb_sample_syndromes",
    "qb_pipeline_create", "qb_pipeline_destroy", "qb_pipeline_set_stream", "qb_pipeline_run", "qb_pipeline_run_events_host",
    "qb_pipeline_decode_host", "qb_pipeline_last_stats", "qb_pipeline_enable_detail", "qb_pipeline_last_batch_detail",
    "qb_osd0_pipeline_host", "qb_decoder_osd_stats", "qb_decoder_set_precision", "qb_decoder_minsum_path",
    "qb_sampler_geometric_table",
]


def load():
    """Load the shared library (no CUDA call is made by loading)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise QbError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`."
                              " There is no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            lib.qb_last_error.restype = C.c_char_p
            lib.qb_version.restype = C.c_char_p
            for name in EXPORTS:
                fn = getattr(lib, name)
                if name not in ("qb_last_error", "qb_version", "qb_decoder_destroy", "qb_sampler_destroy", "qb_pipeline_destroy"):
                    fn.restype = C.c_int
            lib.qb_decoder_destroy.restype = None
            lib.qb_sampler_destroy.restype = None
            lib.qb_pipeline_destroy.restype = None
            _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().qb_last_error().decode(errors="replace")
        if rc == -1:
            raise ValueError(msg)
        raise QbError(f"libqldpc_b200 error {rc}: {msg}")


def require_gpu():
    lib = load()
    if lib.qb_device_count() <= 0:
        raise QbError("no CUDA device visible: qldpc_b200 has no CPU fallback")
    return lib


def default_device():
    """One process per GPU: LOCAL_RANK picks the device (torchrun), else device 0."""
    return int(os.environ.get("QLDPC_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Decoder:
    """One decoding side on the device (qb_decoder)."""

    def __init__(self, indptr, indices, n, prior, logical_rows=None, device=None):
        lib = require_gpu()
        self.device = default_device() if device is None else device
        self.indptr, self.indices = i32(indptr), i32(indices)
        self.m, self.n = len(self.indptr) - 1, int(n)
        self.nnz = int(self.indptr[-1])
        prior = np.ascontiguousarray(prior, dtype=np.float64)
        assert prior.shape == (self.n,)
        if logical_rows is None:
            lptr, lidx, self.k = np.zeros(1, np.int32), np.zeros(1, np.int32), 0
        else:
            L = np.asarray(logical_rows) & 1
            self.k = L.shape[0]
            lptr = np.concatenate([[0], np.cumsum(L.sum(axis=1))]).astype(np.int32)
            lidx = np.nonzero(L)[1].astype(np.int32)
            if lidx.size == 0:
                lidx = np.zeros(1, np.int32)
        self._h = C.c_void_p()
        check(lib.qb_decoder_create(self.device, self.m, self.n, ptr(self.indptr), ptr(self.indices), ptr(prior),
                                    self.k, ptr(lptr), ptr(lidx), C.byref(self._h)))
        self._prior_key = prior.tobytes()

    @classmethod
    def from_dense(cls, H, prior, logical_rows=None, device=None):
        H = np.asarray(H)
        mask = H != 0
        indptr = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int32)
        indices = np.nonzero(mask)[1].astype(np.int32)
        return cls(indptr, indices, H.shape[1], prior, logical_rows, device)

    def set_precision(self, precision):
        """QB_PRECISION_F32 (default) or QB_PRECISION_HALF2 (packed opt-in mode) for this handle's min-sum calls."""
        check(load().qb_decoder_set_precision(self._h, int(precision)))

    def minsum_path(self):
        """0 = compressed-state kernel, 1 = per-edge kernel on one SM, N >= 2 = per-edge kernel on a cluster of N SMs."""
        return int(load().qb_decoder_minsum_path(self._h))

    def set_prior(self, prior):
        prior = np.ascontiguousarray(prior, dtype=np.float64)
        key = prior.tobytes()
        if key != self._prior_key:
            check(load().qb_decoder_set_prior(self._h, ptr(prior)))
            self._prior_key = key

    # -- host-buffer entry points ---------------------------------------------------------------
    def minsum(self, syndromes, max_iter, alpha_mode, alpha=1.0, alpha_seq=None, damping=1.0, clip_llr=20.0,
               dense_variant=False, want_values=True):
        syn = np.ascontiguousarray(syndromes, dtype=np.int8).reshape(-1, self.m)
        B = syn.shape[0]
        hard = np.zeros((B, self.n), dtype=np.int8)
        conv = np.zeros(B, dtype=np.uint8)
        fin = np.zeros(B, dtype=np.int32)
        values = np.zeros((B, self.n), dtype=np.float64) if want_values else None
        seq = None if alpha_seq is None else np.ascontiguousarray(alpha_seq, dtype=np.float64)
        check(load().qb_minsum_decode_host(self._h, ptr(syn), B, int(max_iter), int(alpha_mode), c_double(float(alpha)),
                                           ptr(seq), 0 if seq is None else len(seq), c_double(damping), c_double(clip_llr),
                                           int(bool(dense_variant)), ptr(hard), ptr(conv), ptr(fin), ptr(values)))
        return hard, conv.astype(bool), values, fin

    def minsum_core(self, Q, syndrome_sign, alpha):
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(-1, self.nnz)
        ss = np.ascontiguousarray(syndrome_sign, dtype=np.float64).reshape(-1, self.m)
        B = Q.shape[0]
        R = np.zeros((B, self.nnz)); Rs = np.zeros((B, self.n))
        check(load().qb_minsum_core_host(self._h, ptr(Q), ptr(ss), B, c_double(alpha), ptr(R), ptr(Rs)))
        return R, Rs

    def alpha_messages(self, syndromes, prior, alpha_prev, damping=1.0, clip_llr=20.0):
        """Unscaled check messages after len(alpha_prev) full iterations (reference alpha.py:206-253)."""
        syn = np.ascontiguousarray(syndromes, dtype=np.int8).reshape(-1, self.m)
        prior = np.ascontiguousarray(prior, dtype=np.float64)
        ap = np.ascontiguousarray(alpha_prev, dtype=np.float64).reshape(-1)
        R = np.zeros((syn.shape[0], self.nnz), dtype=np.float64)
        check(load().qb_alpha_messages_host(self._h, ptr(syn), syn.shape[0], ptr(prior), len(ap), ptr(ap) if len(ap) else None,
                                            c_double(damping), c_double(clip_llr), ptr(R)))
        return R

    def bp(self, syndromes, max_iter):
        syn = np.ascontiguousarray(syndromes, dtype=np.int8).reshape(-1, self.m)
        B = syn.shape[0]
        hard = np.zeros((B, self.n), dtype=np.int8); conv = np.zeros(B, dtype=np.uint8)
        fin = np.zeros(B, dtype=np.int32); values = np.zeros((B, self.n), dtype=np.float64)
        check(load().qb_bp_decode_host(self._h, ptr(syn), B, int(max_iter), ptr(hard), ptr(conv), ptr(fin), ptr(values)))
        return hard, conv.astype(bool), values, fin

    def syndrome_check(self, candidates):
        cand = np.ascontiguousarray(candidates, dtype=np.int8).reshape(-1, self.n)
        out = np.zeros((cand.shape[0], self.m), dtype=np.int8)
        check(load().qb_syndrome_check_host(self._h, ptr(cand), cand.shape[0], ptr(out)))
        return out

    def osd0(self, syndromes, hard, llr=None, ordering=None, want_pivots=False):
        syn = np.ascontiguousarray(syndromes, dtype=np.int8).reshape(-1, self.m)
        hd = np.ascontiguousarray(np.asarray(hard) & 1, dtype=np.int8).reshape(-1, self.n)
        B = syn.shape[0]
        llr_a = None if llr is None else np.ascontiguousarray(llr, dtype=np.float64).reshape(B, self.n)
        ord_a = None if ordering is None else np.ascontiguousarray(ordering, dtype=np.int32).reshape(B, self.n)
        sol = np.zeros((B, self.n), dtype=np.int64)
        rank = np.zeros(B, dtype=np.int32)
        rcap = max(1, min(self.m, self.n))
        piv = np.full((B, rcap), -1, dtype=np.int32) if want_pivots else None
        check(load().qb_osd0_host(self._h, ptr(syn), ptr(hd), ptr(llr_a), ptr(ord_a), B, ptr(sol), ptr(rank), ptr(piv)))
        return (sol, rank, piv) if want_pivots else (sol, rank)

    def osd_stats(self):
        """Tier exits of the free-row OSD kernel since the last call, by reason (qb_decoder_osd_stats)."""
        out = np.zeros(10, dtype=np.int32)
        check(load().qb_decoder_osd_stats(self._h, ptr(out)))
        names = ("window", "rows", "slots", "records", "not_materialised")
        return {"tier_a": dict(zip(names, out[:5].tolist())), "tier_b": dict(zip(names, out[5:].tolist()))}

    def osd0_pipeline(self, syndromes, hard, post):
        """The pipeline's OSD-0 kernels on host float32 posteriors -> (solution int8 [B, n], osd_info int32 [B])."""
        syn = np.ascontiguousarray(syndromes, dtype=np.int8).reshape(-1, self.m)
        hd = np.ascontiguousarray(np.asarray(hard) & 1, dtype=np.int8).reshape(-1, self.n)
        B = syn.shape[0]
        pf = np.ascontiguousarray(post, dtype=np.float32).reshape(B, self.n)
        sol = np.zeros((B, self.n), dtype=np.int8)
        info = np.zeros(B, dtype=np.int32)
        check(load().qb_osd0_pipeline_host(self._h, ptr(syn), ptr(hd), ptr(pf), B, ptr(sol), ptr(info)))
        return sol, info

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            load().qb_decoder_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Sampler:
    """Fault tables of one (code, circuit) on the device (qb_sampler)."""

    def __init__(self, ft, device=None):
        lib = require_gpu()
        self.device = default_device() if device is None else device
        self.L, self.k = int(ft.L), int(ft.Z.k)
        self.mZ, self.nZ, self.mX, self.nX = ft.Z.m, ft.Z.n_cols, ft.X.m, ft.X.n_cols
        if self.k > 32:
            raise ValueError("at most 32 logical observables are supported")
        self._h = C.c_void_p()
        a = dict(kind=i32(ft.loc_kind), cz=i32(ft.loc_colZ), cx=i32(ft.loc_colX),
                 pz=i32(ft.Z.col_ptr), rz=i32(ft.Z.col_rows), lz=np.ascontiguousarray(ft.Z.col_logmask, dtype=np.uint32),
                 px=i32(ft.X.col_ptr), rx=i32(ft.X.col_rows), lx=np.ascontiguousarray(ft.X.col_logmask, dtype=np.uint32))
        check(lib.qb_sampler_create(self.device, self.L, ptr(a["kind"]), ptr(a["cz"]), ptr(a["cx"]),
                                    self.mZ, self.nZ, ptr(a["pz"]), ptr(a["rz"]), ptr(a["lz"]),
                                    self.mX, self.nX, ptr(a["px"]), ptr(a["rx"]), ptr(a["lx"]), self.k, C.byref(self._h)))

    def syndromes_from_events(self, ev_ptr, events):
        ev_ptr = i32(ev_ptr)
        events = np.ascontiguousarray(events, dtype=np.uint32)
        B = len(ev_ptr) - 1
        sz = np.zeros((B, self.mZ), dtype=np.int8); sx = np.zeros((B, self.mX), dtype=np.int8)
        tz = np.zeros((B, self.k), dtype=np.int8); tx = np.zeros((B, self.k), dtype=np.int8)
        check(load().qb_syndrome_from_events_host(self._h, ptr(ev_ptr), ptr(events), B, ptr(sz), ptr(tz), ptr(sx), ptr(tx)))
        return sz, tz, sx, tx

    def sample(self, seed, first_shot, B, error_rate):
        """K1+K2 on the device, results copied to the host (torch only provides the device buffers):
        (synZ_bits uint32 [B, mwZ], trueZ uint32 [B], synX_bits, trueX, nfaults int32 [B])."""
        import torch
        dev = torch.device("cuda", self.device)
        mwZ, mwX = max(1, (self.mZ + 31) // 32), max(1, (self.mX + 31) // 32)
        sz = torch.zeros((B, mwZ), dtype=torch.int32, device=dev); sx = torch.zeros((B, mwX), dtype=torch.int32, device=dev)
        tz = torch.zeros(B, dtype=torch.int32, device=dev); tx = torch.zeros(B, dtype=torch.int32, device=dev)
        nf = torch.zeros(B, dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        check(load().qb_sample_syndromes(self._h, c_u64(int(seed)), c_u64(int(first_shot)), int(B), c_double(error_rate),
                                         c_void_p(sz.data_ptr()), c_void_p(tz.data_ptr()), c_void_p(sx.data_ptr()),
                                         c_void_p(tx.data_ptr()), c_void_p(nf.data_ptr()), c_void_p(st)))
        torch.cuda.synchronize(dev)
        u = lambda t: t.cpu().numpy().view(np.uint32)
        return u(sz), u(tz), u(sx), u(tx), nf.cpu().numpy()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            load().qb_sampler_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def geometric_table(p):
    """Jump table of the fault sampler for error rate p (uint32 [1024], entry 0 unused): qb_sampler_geometric_table."""
    out = np.zeros(1024, dtype=np.uint32)
    n = load().qb_sampler_geometric_table(c_double(float(p)), ptr(out), 1024)
    if n < 0:
        check(n)
    return out


def make_config(max_iter, alpha_mode, alpha_z=1.0, alpha_x=1.0, clip_llr=20.0, use_osd=True, precision=QB_PRECISION_F32):
    """Build a qb_decode_config; returns (config, keepalive) -- keep both until the call returns."""
    cfg = DecodeConfig()
    keep = []
    cfg.max_iter = int(max_iter)
    cfg.clip_llr = float(clip_llr)
    cfg.use_osd = 1 if use_osd else 0
    cfg.precision = int(precision)
    if alpha_mode == QB_ALPHA_SEQUENCE:
        sz = np.ascontiguousarray(alpha_z, dtype=np.float32)
        sx = np.ascontiguousarray(alpha_x, dtype=np.float32)
        if sz.ndim != 1 or sz.size == 0 or sx.ndim != 1 or sx.size == 0:
            raise ValueError("alpha must be a non-empty 1D sequence for alvarado-autoregressive")
        keep += [sz, sx]
        cfg.alpha_mode = QB_ALPHA_SEQUENCE
        cfg.alpha_seq_z_h, cfg.alpha_seq_x_h = sz.ctypes.data, sx.ctypes.data
        cfg.alpha_len_z, cfg.alpha_len_x = len(sz), len(sx)
    else:
        cfg.alpha_mode = int(alpha_mode)
        cfg.alpha_z, cfg.alpha_x = float(alpha_z), float(alpha_x)
    return cfg, keep


class Pipeline:
    """Sampler + both decoders + batch workspaces (qb_pipeline)."""

    def __init__(self, sampler, decZ, decX, max_batch):
        lib = require_gpu()
        self.sampler, self.decZ, self.decX = sampler, decZ, decX
        self.max_batch = int(max_batch)
        self._h = C.c_void_p()
        check(lib.qb_pipeline_create(None if sampler is None else sampler._h, decZ._h, decX._h, self.max_batch, C.byref(self._h)))

    def run(self, seed, first_shot, n_shots, error_rate, cfg, want_flags=False):
        counts = np.zeros(8, dtype=np.int64)
        flags = np.zeros(int(n_shots), dtype=np.uint8) if want_flags else None
        c, keep = cfg
        check(load().qb_pipeline_run(self._h, c_u64(int(seed)), c_u64(int(first_shot)), c_i64(int(n_shots)),
                                     c_double(error_rate), C.byref(c), ptr(counts), ptr(flags)))
        return counts, flags

    def run_events(self, ev_ptr, events, cfg, want_detail=False, flags_out=None):
        ev_ptr = ev_ptr if (isinstance(ev_ptr, np.ndarray) and ev_ptr.dtype == np.int32 and ev_ptr.flags.c_contiguous) else i32(ev_ptr)
        events = events if (isinstance(events, np.ndarray) and events.dtype == np.uint32 and events.flags.c_contiguous) else np.ascontiguousarray(events, dtype=np.uint32)
        B = len(ev_ptr) - 1
        counts = np.zeros(8, dtype=np.int64)
        flags = np.zeros(B, dtype=np.uint8) if flags_out is None else flags_out
        conv = np.zeros((2, B), dtype=np.uint8) if want_detail else None
        fin = np.zeros((2, B), dtype=np.int32) if want_detail else None
        c, keep = cfg
        check(load().qb_pipeline_run_events_host(self._h, ptr(ev_ptr), ptr(events), B, C.byref(c), ptr(counts), ptr(flags),
                                                 ptr(conv), ptr(fin)))
        return (counts, flags, conv, fin) if want_detail else (counts, flags)

    def decode(self, sparse_z, true_z_mask, sparse_x, true_x_mask, cfg):
        sz = np.ascontiguousarray(sparse_z, dtype=np.int8); sx = np.ascontiguousarray(sparse_x, dtype=np.int8)
        tz = np.ascontiguousarray(true_z_mask, dtype=np.uint32); tx = np.ascontiguousarray(true_x_mask, dtype=np.uint32)
        B = sz.shape[0]
        counts = np.zeros(8, dtype=np.int64); flags = np.zeros(B, dtype=np.uint8)
        c, keep = cfg
        check(load().qb_pipeline_decode_host(self._h, ptr(sz), ptr(tz), ptr(sx), ptr(tx), B, C.byref(c), ptr(counts), ptr(flags)))
        return counts, flags

    def enable_detail(self, on=True):
        check(load().qb_pipeline_enable_detail(self._h, 1 if on else 0))

    def last_batch_detail(self, side, B, want_post=True, want_info=True):
        """(hard_bits uint32 [B, nw], posteriors float32 [B, n] or None, osd_info int32 [B] or None) of the last batch."""
        dec = self.decX if side else self.decZ
        nw = max(1, (dec.n + 31) // 32)
        hard = np.zeros((B, nw), dtype=np.uint32)
        post = np.zeros((B, dec.n), dtype=np.float32) if want_post else None
        info = np.zeros(B, dtype=np.int32) if want_info else None
        check(load().qb_pipeline_last_batch_detail(self._h, int(side), ptr(hard), ptr(post), ptr(info)))
        return hard, post, info

    def set_stream(self, cuda_stream=None):
        """Issue the pipeline's work on ``cuda_stream`` (a cudaStream_t as int, e.g.
        ``torch.cuda.current_stream().cuda_stream``); ``None`` restores the pipeline's own stream."""
        check(load().qb_pipeline_set_stream(self._h, c_void_p(0 if cuda_stream is None else int(cuda_stream)),
                                            0 if cuda_stream is None else 1))

    def stats(self):
        s = PipelineStats()
        check(load().qb_pipeline_last_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in PipelineStats._fields_}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            load().qb_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- small handle cache for the per-call reference API ----------------------------------------------
_decoder_cache = {}
_CACHE_MAX = 8


def graph_key(indptr, indices, n):
    h = hashlib.blake2b(digest_size=16)
    h.update(np.ascontiguousarray(indptr, dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(indices, dtype=np.int32).tobytes())
    h.update(str(int(n)).encode())
    return h.hexdigest()


def cached_decoder(indptr, indices, n, prior=None):
    """Decoder handle for a CSR graph, reused across calls (the reference API passes H every call).

    ``prior=None`` is for entry points that never read the priors (OSD-0, syndrome check, single check pass):
    the handle is reused as it is -- no prior upload, no re-layout of the per-edge plan -- so the reference's
    per-shot pattern min-sum -> OSD (engine.py:84-97) does not flip the handle's prior back and forth."""
    key = (graph_key(indptr, indices, n), default_device())
    dec = _decoder_cache.get(key)
    if dec is None:
        if len(_decoder_cache) >= _CACHE_MAX:
            _decoder_cache.pop(next(iter(_decoder_cache))).close()
        dec = Decoder(indptr, indices, n, np.zeros(int(n)) if prior is None else prior)
        _decoder_cache[key] = dec
    elif prior is not None:
        dec.set_prior(prior)
    return dec


# The reference API hands over the dense H (or the scipy CSR) on every call; scanning 1008 x 8785 entries and hashing
# the CSR per call would dominate a per-shot loop.  Results are memoised per array object: the key holds a weak
# fingerprint (id, shape, dtype, data pointer) and the value keeps a reference to the array so the id cannot be
# recycled while the entry lives.  Callers that mutate H in place between calls must pass a new array.
_csr_cache = {}
_CSR_CACHE_MAX = 16


def dense_csr(H):
    """(indptr, indices) int32 of the non-zero pattern of a dense matrix, memoised per array object."""
    H = np.asarray(H)
    key = (id(H), H.shape, H.dtype.str, H.__array_interface__["data"][0])
    hit = _csr_cache.get(key)
    if hit is not None and hit[0] is H:
        return hit[1], hit[2]
    mask = H != 0
    indptr = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int32)
    indices = np.nonzero(mask)[1].astype(np.int32)
    if len(_csr_cache) >= _CSR_CACHE_MAX:
        _csr_cache.pop(next(iter(_csr_cache)))
    _csr_cache[key] = (H, indptr, indices)
    return indptr, indices


_handle_by_arrays = {}


def cached_decoder_for(indptr, indices, n, prior=None):
    """cached_decoder() without re-hashing the CSR arrays when the same array objects come back."""
    key = (id(indptr), id(indices), int(n), default_device())
    hit = _handle_by_arrays.get(key)
    if hit is not None and hit[0] is indptr and hit[1] is indices and hit[2]._h.value:
        dec = hit[2]
        if prior is not None:
            dec.set_prior(prior)
        return dec
    dec = cached_decoder(indptr, indices, n, prior)
    if len(_handle_by_arrays) >= _CSR_CACHE_MAX:
        _handle_by_arrays.pop(next(iter(_handle_by_arrays)))
    _handle_by_arrays[key] = (indptr, indices, dec)
    return dec
