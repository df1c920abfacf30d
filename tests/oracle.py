"""ctypes binding of the CPU oracle (oracle/build/libcz_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(_ORACLE_DIR, "build", "libcz_oracle.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
szp = C.POINTER(C.c_size_t)


def build():
    subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR])


def _load():
    if not os.path.exists(_SO):
        build()
    lib = C.CDLL(_SO)
    lib.czo_expf.restype = C.c_float
    lib.czo_expf.argtypes = [C.c_float]
    lib.czo_expf_checksum.restype = C.c_uint64
    lib.czo_expf_checksum.argtypes = [C.c_uint64, C.c_uint64]
    lib.czo_ac_p_min.restype = C.c_double
    lib.czo_softmax_pdf.argtypes = [f32p, C.c_size_t, f64p]
    lib.czo_softmax_pdf_floor.argtypes = [f32p, C.c_size_t, C.c_double, f64p]
    lib.czo_combined_pdf_with_literals.argtypes = [f32p, C.c_size_t, f64p]
    lib.czo_quantize_pdf_to_cdf.argtypes = [f64p, C.c_size_t, u32p]
    lib.czo_logits_to_cdf.argtypes = [f32p, C.c_size_t, C.c_int, u32p]
    lib.czo_encoder_new.restype = C.c_void_p
    lib.czo_encoder_free.argtypes = [C.c_void_p]
    lib.czo_encoder_encode_counts.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
    lib.czo_encoder_bytes_written.restype = C.c_uint64
    lib.czo_encoder_bytes_written.argtypes = [C.c_void_p]
    lib.czo_encoder_finish.restype = u8p
    lib.czo_encoder_finish.argtypes = [C.c_void_p, szp]
    lib.czo_decoder_new.restype = C.c_void_p
    lib.czo_decoder_new.argtypes = [u8p, C.c_size_t]
    lib.czo_decoder_free.argtypes = [C.c_void_p]
    lib.czo_decoder_decode_symbol_counts.restype = C.c_size_t
    lib.czo_decoder_decode_symbol_counts.argtypes = [C.c_void_p, u32p, C.c_size_t, C.c_uint32]
    lib.czo_decoder_peek_value.restype = C.c_uint32
    lib.czo_decoder_peek_value.argtypes = [C.c_void_p, C.c_uint32]
    lib.czo_read_header_v2.restype = C.c_size_t
    lib.czo_read_header_v2.argtypes = [u8p, C.c_size_t, C.c_void_p, szp]
    lib.czo_write_header_v2.restype = C.c_size_t
    lib.czo_write_header_v2.argtypes = [u8p, C.c_size_t, C.c_void_p, u8p]
    lib.czo_flags_pack.restype = C.c_uint32
    lib.czo_flags_pack.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32]
    lib.czo_llama_new.restype = C.c_void_p
    lib.czo_llama_new.argtypes = [C.c_void_p]
    lib.czo_rwkv7_new.restype = C.c_void_p
    lib.czo_rwkv7_new.argtypes = [C.c_void_p]
    lib.czo_table_session_new.restype = C.c_void_p
    lib.czo_table_session_new.argtypes = [C.c_size_t, f32p, C.c_size_t]
    lib.czo_session_free.argtypes = [C.c_void_p]
    lib.czo_session_set_tensor.argtypes = [C.c_void_p, C.c_char_p, f32p, C.c_size_t]
    lib.czo_session_vocab_size.restype = C.c_size_t
    lib.czo_session_vocab_size.argtypes = [C.c_void_p]
    lib.czo_session_index_pos.restype = C.c_size_t
    lib.czo_session_index_pos.argtypes = [C.c_void_p]
    lib.czo_session_step_logits.restype = f32p
    lib.czo_session_step_logits.argtypes = [C.c_void_p, C.c_uint32]
    lib.czo_session_reprime.restype = f32p
    lib.czo_session_reprime.argtypes = [C.c_void_p, u32p, C.c_size_t]
    lib.czo_encode_tokens.argtypes = [C.c_void_p, u32p, C.c_size_t, C.c_void_p, C.POINTER(u8p), szp]
    lib.czo_decode_tokens.argtypes = [C.c_void_p, u8p, C.c_size_t, C.c_uint32, C.c_size_t, C.c_void_p, u32p]
    lib.czo_xe_bits_over_span.restype = C.c_double
    lib.czo_xe_bits_over_span.argtypes = [C.c_void_p, C.c_int, u32p, C.c_size_t, u32p, C.c_size_t, u32p, C.c_size_t]
    lib.czo_free.argtypes = [C.c_void_p]
    return lib


lib = _load()
AC_CDF_TOTAL = 1 << 30


class HeaderV2(C.Structure):
    _fields_ = [
        ("bos_token_id", C.c_uint32),
        ("token_count", C.c_uint64),
        ("orig_len_bytes", C.c_uint64),
        ("model_hash16", C.c_uint8 * 16),
        ("tokenizer_hash16", C.c_uint8 * 16),
        ("orig_hash16", C.c_uint8 * 16),
        ("reserved_flags", C.c_uint32),
        ("context_window", C.c_uint32),
        ("vocab_size", C.c_uint32),
        ("model_file_repr_len", C.c_uint32),
        ("reprime_interval", C.c_uint32),
    ]


class LlamaConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("vocab", "d_model", "n_layers", "n_heads", "n_kv_heads", "head_dim", "d_ffn")] + [
        ("rms_eps", C.c_float),
        ("rope_theta", C.c_float),
        ("max_pos", C.c_int),
        ("round_bf16", C.c_int),
    ]


class Rwkv7Config(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("vocab", "d_model", "n_layers", "head_dim", "d_ffn", "lora_w", "lora_a", "lora_v", "lora_g")] + [
        ("norm_eps", C.c_float),
        ("round_bf16", C.c_int),
    ]


class PrimeEvent(C.Structure):
    _fields_ = [("i", C.c_size_t), ("prime", u32p), ("prime_len", C.c_size_t), ("hold_until", C.c_size_t)]


class LoopOpts(C.Structure):
    _fields_ = [
        ("backend", C.c_int),
        ("context", C.c_size_t),
        ("reprime_interval", C.c_size_t),
        ("events", C.POINTER(PrimeEvent)),
        ("n_events", C.c_size_t),
        ("reprime_log", szp),
        ("reprime_log_cap", C.c_size_t),
        ("n_reprimes", C.c_size_t),
    ]


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(f32p)


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(u32p)


def logits_to_cdf(logits, mode=0):
    a, p = _f32(logits)
    v = a.shape[0]
    n = v + 256 if mode == 1 else v
    cdf = np.empty(n + 1, dtype=np.uint32)
    lib.czo_logits_to_cdf(p, v, mode, cdf.ctypes.data_as(u32p))
    return cdf


def softmax_pdf_floor(logits, p_floor=None):
    a, p = _f32(logits)
    pdf = np.empty(a.shape[0], dtype=np.float64)
    lib.czo_softmax_pdf_floor(p, a.shape[0], lib.czo_ac_p_min() if p_floor is None else p_floor, pdf.ctypes.data_as(f64p))
    return pdf


def combined_pdf_with_literals(logits):
    a, p = _f32(logits)
    pdf = np.empty(a.shape[0] + 256, dtype=np.float64)
    lib.czo_combined_pdf_with_literals(p, a.shape[0], pdf.ctypes.data_as(f64p))
    return pdf


def ac_encode(bounds, total=AC_CDF_TOTAL):
    """bounds: iterable of (c_lo, c_hi). Returns payload bytes."""
    e = lib.czo_encoder_new()
    try:
        for lo, hi in bounds:
            if lib.czo_encoder_encode_counts(e, int(lo), int(hi), total) != 0:
                raise ValueError("zero-width interval")
        n = C.c_size_t()
        p = lib.czo_encoder_finish(e, C.byref(n))
        return bytes(p[: n.value])
    finally:
        lib.czo_encoder_free(e)


class Decoder:
    def __init__(self, payload: bytes):
        self._buf = (C.c_uint8 * max(1, len(payload))).from_buffer_copy(payload or b"\0")
        self._d = lib.czo_decoder_new(self._buf, len(payload))

    def peek(self, total=AC_CDF_TOTAL):
        return lib.czo_decoder_peek_value(self._d, total)

    def decode(self, cdf, total=AC_CDF_TOTAL):
        a, p = _u32(cdf)
        return lib.czo_decoder_decode_symbol_counts(self._d, p, a.shape[0], total)

    def __del__(self):
        if getattr(self, "_d", None):
            lib.czo_decoder_free(self._d)
            self._d = None


class Session:
    """Mirror of the reference's LanguageModelSession trait (src/models.rs:28-33) over the oracle."""

    def __init__(self, handle, keep=()):
        if not handle:
            raise RuntimeError("oracle session construction failed")
        self._h = handle
        self._keep = keep

    @classmethod
    def llama(cls, cfg: dict, tensors: dict, round_bf16=0, max_pos=1100):
        c = LlamaConfig(cfg["vocab"], cfg["d_model"], cfg["n_layers"], cfg["n_heads"], cfg["n_kv_heads"], cfg["head_dim"],
                        cfg["d_ffn"], cfg.get("rms_eps", 1e-5), cfg.get("rope_theta", 1e5), max_pos, round_bf16)
        s = cls(lib.czo_llama_new(C.byref(c)))
        for name, arr in tensors.items():
            a, p = _f32(arr)
            if lib.czo_session_set_tensor(s._h, name.encode(), p, a.size) != 0:
                raise KeyError(f"oracle rejected tensor {name} with {a.size} elements")
        return s

    @classmethod
    def rwkv7(cls, cfg: dict, tensors: dict, round_bf16=0):
        c = Rwkv7Config(cfg["vocab"], cfg["d_model"], cfg["n_layers"], cfg["head_dim"], cfg["d_ffn"], cfg["lora_w"], cfg["lora_a"],
                        cfg["lora_v"], cfg["lora_g"], cfg.get("norm_eps", 1e-5), round_bf16)
        s = cls(lib.czo_rwkv7_new(C.byref(c)))
        for name, arr in tensors.items():
            a, p = _f32(arr)
            if lib.czo_session_set_tensor(s._h, name.encode(), p, a.size) != 0:
                raise KeyError(f"oracle rejected tensor {name} with {a.size} elements")
        return s

    @classmethod
    def table(cls, table):
        a, p = _f32(table)
        return cls(lib.czo_table_session_new(a.shape[1], p, a.shape[0]), keep=(a,))

    def vocab_size(self):
        return lib.czo_session_vocab_size(self._h)

    def index_pos(self):
        return lib.czo_session_index_pos(self._h)

    def step_logits(self, token):
        p = lib.czo_session_step_logits(self._h, int(token))
        return np.ctypeslib.as_array(p, shape=(self.vocab_size(),)).copy()

    def reprime(self, history):
        a, p = _u32(history)
        r = lib.czo_session_reprime(self._h, p, a.shape[0])
        return np.ctypeslib.as_array(r, shape=(self.vocab_size(),)).copy()

    def _opts(self, backend, context, reprime_interval, events):
        o = LoopOpts()
        o.backend = backend
        o.context = context
        o.reprime_interval = reprime_interval
        keep = []
        if events:
            arr = (PrimeEvent * len(events))()
            for k, (i, prime, hold) in enumerate(events):
                a, p = _u32(prime)
                keep.append(a)
                arr[k] = PrimeEvent(i, p, a.shape[0], hold)
            o.events = arr
            o.n_events = len(events)
            keep.append(arr)
        log = (C.c_size_t * 65536)()
        o.reprime_log = log
        o.reprime_log_cap = 65536
        keep.append(log)
        return o, keep

    def encode_tokens(self, ids, backend=0, context=512, reprime_interval=512, events=None):
        a, p = _u32(ids)
        o, keep = self._opts(backend, context, reprime_interval, events)
        out = u8p()
        n = C.c_size_t()
        rc = lib.czo_encode_tokens(self._h, p, a.shape[0], C.byref(o), C.byref(out), C.byref(n))
        if rc != 0:
            raise ValueError(f"oracle encode failed rc={rc}")
        payload = bytes(out[: n.value])
        lib.czo_free(out)
        return payload, list(o.reprime_log[: o.n_reprimes])

    def decode_tokens(self, payload, bos, token_count, backend=0, context=512, reprime_interval=512, events=None):
        buf = (C.c_uint8 * max(1, len(payload))).from_buffer_copy(payload or b"\0")
        o, keep = self._opts(backend, context, reprime_interval, events)
        ids = np.empty(token_count + 1, dtype=np.uint32)
        rc = lib.czo_decode_tokens(self._h, buf, len(payload), bos, token_count, C.byref(o), ids.ctypes.data_as(u32p))
        if rc != 0:
            raise ValueError(f"oracle decode failed rc={rc}")
        return ids, list(o.reprime_log[: o.n_reprimes])

    def xe_bits(self, history, targets, hint=None, backend=0):
        h, hp = _u32(history)
        t, tp = _u32(targets)
        hint = np.zeros(0, np.uint32) if hint is None else hint
        q, qp = _u32(hint)
        return lib.czo_xe_bits_over_span(self._h, backend, hp, h.shape[0], tp, t.shape[0], qp, q.shape[0])

    def __del__(self):
        if getattr(self, "_h", None):
            lib.czo_session_free(self._h)
            self._h = None
