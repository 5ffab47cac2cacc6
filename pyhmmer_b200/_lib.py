"""ctypes binding of ``libb2h.so`` (the C ABI declared in ``include/b2h.h``).

There is deliberately no fallback: if the CUDA library has not been built this module raises,
and if no CUDA device is present :func:`context` raises -- nothing in this package computes a
score on the CPU.
"""
import ctypes
import os
import threading

import numpy as np

# One hardware work queue per stream instead of the default 8: the engine keeps ~30 streams busy and queue aliasing would
# serialise its high-priority survivor lane behind queued cascade launches.  Read by the driver when the CUDA context is
# created, so it only takes effect if this package is imported before anything else initialises CUDA.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2h.so")

B2H_OK, B2H_EMEM, B2H_EINVAL, B2H_ERANGE, B2H_ENORESULT, B2H_ECUDA = 0, 5, 11, 16, 19, 100

c_void_p, c_int, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
c_u8, c_i16, c_i32, c_i64, c_u64 = ctypes.c_uint8, ctypes.c_int16, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
P = ctypes.POINTER


class OProfileDesc(ctypes.Structure):
    """``b2h_oprofile_desc`` (include/b2h.h)."""
    _fields_ = [
        ("M", c_i32), ("K", c_i32), ("Kp", c_i32), ("L", c_i32), ("mode_multihit", c_i32), ("max_length", c_i32),
        ("msv_cost", c_void_p),
        ("tbm_b", c_u8), ("tec_b", c_u8), ("tjb_b", c_u8), ("base_b", c_u8), ("bias_b", c_u8),
        ("scale_b", c_float),
        ("vit_rsc", c_void_p), ("vit_tsc", c_void_p),
        ("xw", (c_i16 * 2) * 4), ("base_w", c_i16), ("ddbound_w", c_i16), ("scale_w", c_float),
        ("fwd_rsc", c_void_p), ("fwd_tsc", c_void_p),
        ("xf", (c_float * 2) * 4),
        ("evparam", c_float * 6), ("cutoff", c_float * 6), ("compo", c_float * 20), ("bgf", c_float * 20),
        ("degen", c_void_p),
    ]


class LenParams(ctypes.Structure):
    """``b2h_len_params`` (include/b2h.h)."""
    _fields_ = [("tjb_b", c_u8), ("xw_move", c_i16), ("pmove", c_float), ("ploop", c_float),
                ("null1", c_float), ("p1", c_float), ("flt_len_a", c_float), ("flt_len_b", c_float)]


class SearchParams(ctypes.Structure):
    """``b2h_search_params`` (include/b2h.h)."""
    _fields_ = [("F1", ctypes.c_double), ("F2", ctypes.c_double), ("F3", ctypes.c_double),
                ("do_biasfilter", c_i32), ("do_null2", c_i32), ("seed", ctypes.c_uint32), ("host_threads", c_i32),
                ("seq_counters", c_i32), ("reserved", c_i32)]


class HitRec(ctypes.Structure):
    """``b2h_hit`` (include/b2h.h)."""
    _fields_ = [("profile", c_i32), ("seq", c_i32), ("score", c_float), ("pre_score", c_float), ("sum_score", c_float),
                ("lnP", ctypes.c_double), ("pre_lnP", ctypes.c_double), ("sum_lnP", ctypes.c_double),
                ("nexpected", c_float), ("nregions", c_i32), ("nclustered", c_i32), ("noverlaps", c_i32),
                ("nenvelopes", c_i32), ("ndom", c_i32), ("best_domain", c_i32), ("dom_offset", c_i64)]


class DomainRec(ctypes.Structure):
    """``b2h_domain`` (include/b2h.h)."""
    _fields_ = [("ienv", c_i32), ("jenv", c_i32), ("iali", c_i32), ("jali", c_i32),
                ("envsc", c_float), ("domcorrection", c_float), ("dombias", c_float), ("oasc", c_float), ("bitscore", c_float),
                ("lnP", ctypes.c_double), ("hmmfrom", c_i32), ("hmmto", c_i32), ("sqfrom", c_i32), ("sqto", c_i32),
                ("N", c_i32), ("text_offset", c_i64), ("has_rf", c_i32), ("has_cs", c_i32)]


class WindowRec(ctypes.Structure):
    """``b2h_window`` (include/b2h.h)."""
    _fields_ = [("seq", c_i32), ("k", c_i32), ("n", c_i64), ("length", c_i32), ("score", c_float)]


class PressedModel(ctypes.Structure):
    """``b2h_pressed_model`` (include/b2h.h)."""
    _fields_ = [("desc", OProfileDesc), ("alphabet_type", c_i32), ("reserved", c_i32),
                ("name", c_i64), ("acc", c_i64), ("descr", c_i64), ("rf", c_i64), ("mm", c_i64), ("cs", c_i64), ("consensus", c_i64)]


class HMMDesc(ctypes.Structure):
    """``b2h_hmm_desc`` (include/b2h.h)."""
    _fields_ = [("M", c_i32), ("max_length", c_i32), ("t", c_void_p), ("mat", c_void_p),
                ("evparam", c_float * 6), ("cutoff", c_float * 6), ("compo", c_float * 20)]


class LtWindow(ctypes.Structure):
    """``b2h_lt_window`` (include/b2h.h)."""
    _fields_ = [("dsq", c_void_p), ("L", c_i32), ("fwd_xmx", c_void_p), ("bck_xmx", c_void_p),
                ("window_start", c_i64), ("seq_start", c_i64), ("complement", c_i32), ("seq", c_i32),
                ("bck_own_scales", c_i32), ("reserved", c_i32)]


class B2HError(RuntimeError):
    def __init__(self, status, fn, detail=""):
        self.status = status
        super().__init__("%s failed with status %d%s" % (fn, status, (": " + detail) if detail else ""))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "pyhmmer_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)

    def sig(name, restype, *argtypes):
        f = getattr(lib, name)
        f.restype = restype
        f.argtypes = list(argtypes)
        return f

    sig("b2h_hmm_decode_probs", c_int, c_void_p, c_void_p, c_size_t)
    sig("b2h_profile_config", c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
        c_int, c_int, c_void_p, c_void_p, c_void_p)
    sig("b2h_oprofile_convert", c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, P(OProfileDesc))
    sig("b2h_destripe_oprofile", c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p)
    sig("b2h_length_params", c_int, c_int, c_float, P(LenParams))
    sig("b2h_ctx_create", c_int, c_int, P(c_void_p))
    sig("b2h_ctx_destroy", None, c_void_p)
    sig("b2h_ctx_set_stream", c_int, c_void_p, c_void_p)
    sig("b2h_ctx_synchronize", c_int, c_void_p)
    sig("b2h_ctx_last_error", ctypes.c_char_p, c_void_p)
    sig("b2h_ctx_launch_count", c_u64, c_void_p)
    sig("b2h_ctx_set_profiling", c_int, c_void_p, c_int)
    sig("b2h_ctx_stage_ms", c_int, c_void_p, c_void_p, c_int)
    sig("b2h_seqdb_create", c_int, c_void_p, c_void_p, c_void_p, c_size_t, P(c_void_p))
    sig("b2h_seqdb_create_packed", c_int, c_void_p, c_void_p, c_void_p, c_size_t, P(c_void_p))
    sig("b2h_seqdb_destroy", None, c_void_p)
    sig("b2h_seqdb_nseq", c_size_t, c_void_p)
    sig("b2h_seqdb_nres", c_i64, c_void_p)
    sig("b2h_profile_upload", c_int, c_void_p, P(OProfileDesc), P(c_void_p))
    sig("b2h_profile_destroy", None, c_void_p)
    sig("b2h_profile_upload_many", c_int, c_void_p, c_void_p, c_size_t, c_void_p)
    for name in ("b2h_ssv_filter", "b2h_msv_filter", "b2h_viterbi_filter", "b2h_forward_parser",
                 "b2h_backward_parser"):
        if hasattr(lib, name):
            sig(name, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p)
    sig("b2h_search", c_int, c_void_p, c_void_p, c_size_t, c_void_p, P(SearchParams), P(c_void_p))
    sig("b2h_results_nhits", c_size_t, c_void_p)
    sig("b2h_results_hits", P(HitRec), c_void_p)
    sig("b2h_results_ndomains", c_size_t, c_void_p)
    sig("b2h_results_domains", P(DomainRec), c_void_p)
    sig("b2h_results_text", c_void_p, c_void_p, P(c_size_t))
    sig("b2h_results_counters", P(c_i64), c_void_p)
    sig("b2h_results_seq_counters", P(c_i64), c_void_p)
    sig("b2h_results_destroy", None, c_void_p)
    sig("b2h_search_begin", c_int, c_void_p, c_void_p, c_size_t, c_void_p, P(SearchParams), P(c_void_p), P(c_size_t))
    sig("b2h_search_next", c_int, c_void_p, P(c_void_p))
    sig("b2h_search_end", c_int, c_void_p)
    sig("b2h_results_profiles", P(c_i32), c_void_p, P(c_size_t))
    sig("b2h_profile_set_annotation", c_int, c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p)
    sig("b2h_seqdb_h2d_bytes", ctypes.c_size_t, c_void_p)
    sig("b2h_profile_h2d_bytes", ctypes.c_size_t, c_void_p)
    sig("b2h_pack_windows", c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int)
    sig("b2h_ssv_tile_info", c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_double))
    sig("b2h_generic_scores", c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_float,
        c_void_p, c_void_p, c_void_p, c_void_p)
    sig("b2h_generic_decoding", c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int,
        c_void_p, c_void_p, P(c_float), P(c_float), c_void_p, c_void_p, c_void_p)
    sig("b2h_longtarget_windows", c_int, c_void_p, c_void_p, c_void_p, ctypes.c_double, P(c_void_p), P(c_size_t), P(c_void_p), P(c_size_t))
    sig("b2h_longtarget_scan_info", c_int, c_void_p, ctypes.c_double, P(c_int), P(c_int))
    sig("b2h_free", None, c_void_p)
    sig("b2h_window_lengths", c_int, c_void_p, c_void_p, c_void_p)
    sig("b2h_extend_merge_windows", c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_float, P(c_size_t))
    sig("b2h_longtarget_viterbi_windows", c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_double,
        P(c_void_p), P(c_size_t), P(c_void_p), P(c_size_t))
    sig("b2h_longtarget_vit_finish", c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, P(c_void_p), P(c_size_t))
    sig("b2h_longtarget_domains", c_int, c_void_p, c_void_p, c_size_t, P(SearchParams), P(c_void_p))
    sig("b2h_longtarget_hits", c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, P(SearchParams), P(c_void_p))
    sig("b2h_longtarget_vit_threshold", c_int, c_void_p, c_int, c_float, ctypes.c_double, P(c_i32), P(c_i32))
    sig("b2h_profile_set_model_mask", c_int, c_void_p, ctypes.c_char_p)
    sig("b2h_hmm_parse_body", c_int, ctypes.c_char_p, c_size_t, c_int, c_int, c_int, c_void_p, P(c_i32), c_void_p, c_void_p, c_void_p,
        c_void_p, c_void_p)
    sig("b2h_hmm_max_length", c_int, c_int, c_void_p, ctypes.c_double, P(c_i32))
    sig("b2h_hmm_convert_many", c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_int,
        P(c_void_p), P(c_void_p), P(c_size_t))
    sig("b2h_pressed_open", c_int, ctypes.c_char_p, P(c_void_p))
    sig("b2h_pressed_close", None, c_void_p)
    sig("b2h_pressed_rewind", c_int, c_void_p)
    sig("b2h_pressed_last_error", ctypes.c_char_p, c_void_p)
    sig("b2h_pressed_read", c_int, c_void_p, c_size_t, P(c_void_p), P(c_size_t), P(c_void_p), P(c_size_t), P(c_void_p), P(c_size_t))
    sig("b2h_profile_create_host", c_int, P(OProfileDesc), P(c_void_p))
    sig("b2h_debug_domaindef", c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_float, P(SearchParams), P(c_void_p))
    if hasattr(lib, "b2h_null_scores"):
        sig("b2h_null_scores", c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p)
    return lib


lib = _load()


class RecList(list):
    """A list of ctypes records that remembers the array they are views of (``raw``), for bulk (de)serialisation."""
    raw = None


def read_results(handle):
    """Copy a ``b2h_results`` into python objects: (hits, domains, text, counters_flat)."""
    nh = lib.b2h_results_nhits(handle)
    nd = lib.b2h_results_ndomains(handle)
    hp = lib.b2h_results_hits(handle)
    dp = lib.b2h_results_domains(handle)
    # one copy per record array; the list elements are views into those copies
    hits, doms = RecList(), RecList()
    if nh:
        hits.raw = (HitRec * nh).from_buffer_copy(ctypes.string_at(hp, nh * ctypes.sizeof(HitRec)))
        hits.extend(hits.raw)
    if nd:
        doms.raw = (DomainRec * nd).from_buffer_copy(ctypes.string_at(dp, nd * ctypes.sizeof(DomainRec)))
        doms.extend(doms.raw)
    nb = c_size_t()
    tp = lib.b2h_results_text(handle, ctypes.byref(nb))
    text = ctypes.string_at(tp, nb.value) if nb.value else b""
    return hits, doms, text


def ptr(a):
    """Raw data pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def check(status, fn, ctx=None):
    if status != B2H_OK:
        detail = ""
        if ctx is not None:
            detail = (lib.b2h_ctx_last_error(ctx) or b"").decode()
        raise B2HError(status, fn, detail)


class Context:
    """One ``b2h_ctx`` = one GPU of this process."""

    def __init__(self, device=0):
        h = c_void_p()
        st = lib.b2h_ctx_create(device, ctypes.byref(h))
        if st != B2H_OK:
            raise B2HError(st, "b2h_ctx_create",
                           "no usable CUDA device %d (sm_100a required; this package has no CPU path)" % device)
        self.handle = h
        self.device = device

    def set_stream(self, cuda_stream):
        check(lib.b2h_ctx_set_stream(self.handle, c_void_p(cuda_stream or 0)), "b2h_ctx_set_stream", self.handle)

    def synchronize(self):
        check(lib.b2h_ctx_synchronize(self.handle), "b2h_ctx_synchronize", self.handle)

    def set_profiling(self, on):
        check(lib.b2h_ctx_set_profiling(self.handle, int(on)), "b2h_ctx_set_profiling", self.handle)

    def stage_ms(self, reset=True):
        """{stage: milliseconds} accumulated by b2h_search since the last reset (needs set_profiling(True))."""
        a = np.zeros(8, dtype=np.float64)
        check(lib.b2h_ctx_stage_ms(self.handle, ptr(a), int(reset)), "b2h_ctx_stage_ms", self.handle)
        return dict(zip(("ssv", "msv", "bias", "viterbi", "forward", "fwdbck_survivors", "grouping", "reserved"), a.tolist()))

    @property
    def launch_count(self):
        return int(lib.b2h_ctx_launch_count(self.handle))

    def close(self):
        if self.handle:
            lib.b2h_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts = {}
_lock = threading.Lock()


def context(device=None):
    """The process-wide :class:`Context` for ``device`` (default: ``LOCAL_RANK`` or 0)."""
    if device is None:
        device = int(os.environ.get("B2H_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = _contexts[device] = Context(device)
        return ctx
