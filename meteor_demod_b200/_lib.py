"""ctypes binding of the C ABI in include/lrpt_b200.h (liblrpt_b200.so, built in-tree).

The library is the product; this module only loads it. If the shared object is
missing the import fails loudly -- there is no Python or CPU implementation to
fall back to.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("LRPT_SO") or os.path.join(HERE, "liblrpt_b200.so")   # LRPT_SO: A/B builds while tuning

LRPT_OK, LRPT_ERR_ARG, LRPT_ERR_CUDA, LRPT_ERR_NOMEM, LRPT_ERR_CAP, LRPT_ERR_STATE = 0, -1, -2, -3, -4, -5
KERNELS = {"auto": 0, "simple": 1, "ws": 2, "spec": 3, "lane": 4}


class Params(C.Structure):
    _fields_ = [("pll_bw", C.c_float), ("sym_bw", C.c_float), ("freq_max", C.c_float),
                ("samplerate", C.c_int32), ("symrate", C.c_int32), ("interp_factor", C.c_int32),
                ("rrc_order", C.c_int32), ("oqpsk", C.c_int32), ("bps", C.c_int32),
                ("device", C.c_int32), ("nstreams", C.c_int32), ("kernel", C.c_int32)]


class State(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("taps", C.c_uint32),
                ("t_phase", C.c_float), ("t_freq", C.c_float), ("t_prev", C.c_float),
                ("t_dual_state", C.c_int32), ("oq_inphase", C.c_float),
                ("agc_gain", C.c_float), ("agc_bias_re", C.c_float), ("agc_bias_im", C.c_float),
                ("p_phase", C.c_float), ("p_freq", C.c_float), ("p_err", C.c_float),
                ("p_locked", C.c_int32), ("p_locked_once", C.c_int32), ("p_updown", C.c_int32),
                ("nsamples", C.c_int64), ("nsymbols", C.c_int64), ("first_lock_symbol", C.c_int64)]


class Status(C.Structure):
    _fields_ = [("pll_freq", C.c_float), ("mm_omega", C.c_float), ("agc_gain", C.c_float),
                ("locked", C.c_int32), ("locked_once", C.c_int32),
                ("nsamples", C.c_int64), ("nsymbols", C.c_int64), ("first_lock_symbol", C.c_int64)]


class ShardPlan(C.Structure):
    _fields_ = [("chunk", C.c_uint64), ("warm", C.c_uint64), ("overlap", C.c_uint64), ("seed_nfft", C.c_uint64)]


class ShardReport(C.Structure):
    _fields_ = [("nchunks", C.c_int32), ("launches", C.c_int32), ("min_agreement_scan", C.c_float),
                ("min_agreement_final", C.c_float), ("aligned", C.c_int32), ("reserved", C.c_int32),
                ("first_lock_symbol", C.c_int64)]


# name -> (restype, argtypes); one entry per symbol include/lrpt_b200.h declares
SYMBOLS = {
    "lrpt_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Params)]),
    "lrpt_destroy": (None, [C.c_void_p]),
    "lrpt_reset": (C.c_int, [C.c_void_p]),
    "lrpt_reset_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lrpt_process": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                               C.POINTER(C.c_size_t), C.POINTER(C.c_longlong)]),
    "lrpt_process_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                     C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]),
    "lrpt_process_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                            C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                            C.c_void_p]),
    "lrpt_set_symbol_index_output": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "lrpt_sync": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lrpt_stream_wait": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lrpt_stream_release": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lrpt_get_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "lrpt_status": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Status)]),
    "lrpt_state_size": (C.c_size_t, [C.c_void_p]),
    "lrpt_export_state": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_size_t)]),
    "lrpt_import_state": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "lrpt_states_size": (C.c_size_t, [C.c_void_p]),
    "lrpt_export_states_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lrpt_import_states_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "lrpt_snapshot": (C.c_int, [C.c_void_p]),
    "lrpt_restore": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lrpt_shard_find_cuts_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lrpt_shard_quadrants_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lrpt_shard_quadrants_oqpsk_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                                                    C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lrpt_shard_ranges_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lrpt_shard_gather_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lrpt_sharded_process": (C.c_int, [C.POINTER(Params), C.POINTER(ShardPlan), C.c_void_p, C.c_size_t, C.c_void_p,
                                       C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(ShardReport)]),
    "lrpt_fe_sync_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lrpt_fe_sync_words": (C.c_size_t, [C.c_size_t]),
    "lrpt_fe_peaks_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "lrpt_fe_viterbi_scratch_bytes": (C.c_size_t, [C.c_int]),
    "lrpt_fe_viterbi_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "lrpt_alloc_host": (C.c_void_p, [C.c_size_t]),
    "lrpt_free_host": (None, [C.c_void_p]),
    "lrpt_pin_host": (C.c_int, [C.c_void_p, C.c_size_t]),
    "lrpt_unpin_host": (C.c_int, [C.c_void_p]),
    "lrpt_carrier_estimate_device": (C.c_int, [C.POINTER(Params), C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_double,
                                               C.c_void_p, C.c_void_p]),
    "lrpt_fir_stage_device": (C.c_int, [C.POINTER(Params), C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_void_p,
                                        C.c_size_t, C.c_int, C.c_void_p]),
    "lrpt_sharded_release": (None, []),
    "lrpt_sharded_process_multi": (C.c_int, [C.POINTER(Params), C.POINTER(ShardPlan), C.c_void_p, C.c_size_t, C.c_void_p,
                                             C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(ShardReport), C.c_void_p, C.c_int]),
    "lrpt_describe": (C.c_int, [C.POINTER(Params), C.POINTER(State), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "lrpt_get_taps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "lrpt_get_tanh_lut": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lrpt_launch_count": (C.c_ulonglong, [C.c_void_p]),
    "lrpt_fir_fallbacks": (C.c_ulonglong, [C.c_void_p]),
    "lrpt_kernel_name": (C.c_char_p, [C.c_void_p]),
    "lrpt_last_error": (C.c_char_p, [C.c_void_p]),
    "lrpt_strerror": (C.c_char_p, [C.c_int]),
    "lrpt_abi_version": (C.c_int, []),
    "lrpt_freq_delta_from_hz": (C.c_float, [C.c_float, C.c_float]),
}

_lib = None


def load():
    """Load liblrpt_b200.so and bind every declared symbol. Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError("%s not found: build it with `python -m meteor_demod_b200.build` "
                          "(there is no CPU fallback)" % SO_PATH)
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class LrptError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = load().lrpt_strerror(code).decode()
        super().__init__("lrpt error %d (%s)%s" % (code, msg, (": " + detail) if detail else ""))
