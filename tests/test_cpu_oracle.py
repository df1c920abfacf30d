"""CPU suite (-m "not gpu"): the oracle against the reference's golden vectors / fixtures, the host logic of the
product library (schedule, container, BLAKE3, weight generator), and that the C-ABI library loads and exports every
symbol include/candlezip_b200.h declares.  No GPU compute is called here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL = 1 << 30


# ------------------------------------------------------------------ coder KATs
def test_coder_kat_empty_stream():
    assert oracle.ac_encode([]) == bytes([0x40])


def test_coder_kat_uniform4():
    cdf = [0, 1 << 28, 1 << 29, 3 << 28, 1 << 30]
    syms = [0, 1, 2, 3, 3, 2, 1, 0]
    pay = oracle.ac_encode([(cdf[s], cdf[s + 1]) for s in syms])
    assert pay.hex() == "1be440"
    d = oracle.Decoder(pay)
    assert [d.decode(cdf) for _ in syms] == syms


def test_coder_random_cdf_roundtrip():
    rng = np.random.default_rng(7)
    for trial in range(200):
        n = int(rng.integers(2, 40))
        w = rng.integers(1, 1000, n).astype(np.float64)
        cdf = np.concatenate([[0], np.floor(np.cumsum(w) / w.sum() * TOTAL)]).astype(np.uint32)
        cdf[-1] = TOTAL
        cdf = np.maximum.accumulate(cdf)
        ok = np.nonzero(np.diff(cdf.astype(np.int64)) > 0)[0]
        syms = rng.choice(ok, 300)
        pay = oracle.ac_encode([(int(cdf[s]), int(cdf[s + 1])) for s in syms])
        d = oracle.Decoder(pay)
        assert [d.decode(cdf) for _ in syms] == list(syms)


def test_coder_rejects_zero_width():
    with pytest.raises(ValueError):
        oracle.ac_encode([(5, 5)])


# ------------------------------------------------------------------ expf / pdf / CDF
def test_expf_matches_host_libm_on_samples():
    # the exhaustive proof is oracle/expf_exhaustive.c (run in test_expf_exhaustive_negative below)
    rng = np.random.default_rng(0)
    xs = np.concatenate([-rng.exponential(10, 20000), [-0.0, 0.0, -103.9, -104.5, -1e-30, -87.5, -88.5, -np.inf]]).astype(np.float32)
    libm = C.CDLL("libm.so.6")
    libm.expf.restype = C.c_float
    libm.expf.argtypes = [C.c_float]
    for x in xs:
        a = np.float32(oracle.lib.czo_expf(float(x)))
        b = np.float32(libm.expf(float(x)))
        assert a.tobytes() == b.tobytes(), (x, a, b)


def test_expf_exhaustive_negative():
    exe = os.path.join(ROOT, "oracle", "build", "expf_exhaustive")
    if not os.path.exists(exe):
        oracle.build()
    out = subprocess.run([exe, "--neg-only"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout
    assert "mismatches=0" in out.stdout


def _numpy_cdf_smollm(logits):
    """independent restatement of src/main.rs:784-824 in numpy (sequential f64 via cumsum on float64 is sequential)."""
    l = logits.astype(np.float32)
    mx = l.max()
    e = np.array([oracle.lib.czo_expf(float(np.float32(v) - mx)) for v in l], dtype=np.float64)
    s = 0.0
    for v in e:
        s += v
    p = e / s
    acc = 0.0
    cdf = [0]
    for v in p:
        acc += v
        c = int(np.floor(acc * TOTAL))
        c = min(max(c, 0), TOTAL)
        cdf.append(max(c, cdf[-1]))
    cdf[-1] = TOTAL
    return np.array(cdf, dtype=np.uint32)


def test_cdf_properties_and_numpy_restatement():
    rng = np.random.default_rng(3)
    for scale in (0.1, 3.0, 15.0):
        logits = (rng.normal(0, scale, 700)).astype(np.float32)
        cdf = oracle.logits_to_cdf(logits, 0)
        assert cdf[0] == 0 and cdf[-1] == TOTAL and np.all(np.diff(cdf.astype(np.int64)) >= 0)
        assert np.array_equal(cdf, _numpy_cdf_smollm(logits))
    # literal mode: V + 256 symbols, every symbol has mass >= ~2^-29 * total = 2 counts
    logits = rng.normal(0, 8, 500).astype(np.float32)
    cdf = oracle.logits_to_cdf(logits, 1)
    assert cdf.shape[0] == 500 + 257 and cdf[-1] == TOTAL
    assert np.all(np.diff(cdf.astype(np.int64)) >= 1)


def test_cdf_all_equal_and_extreme_logits():
    cdf = oracle.logits_to_cdf(np.zeros(64, np.float32), 0)
    assert np.array_equal(np.diff(cdf.astype(np.int64)), np.full(64, TOTAL // 64))
    l = np.full(100, -1e30, np.float32)
    l[17] = 5.0
    cdf = oracle.logits_to_cdf(l, 0)
    assert cdf[17] == 0 and cdf[18] == TOTAL  # all mass on one symbol; every other interval has zero width


# ------------------------------------------------------------------ reprime schedule vs the reference's shipped traces
def _events_from_gates(gates, n_tokens, chunk=512, lookahead=512):
    ev = []
    for chunk_index, gate, _cand, _bud in gates:
        i = chunk_index * chunk - 1
        if gate == 1 and i < n_tokens:
            ev.append((i, np.array([1, 2, 3], np.uint32), i + lookahead))  # prime content is irrelevant to the schedule
    return ev


@pytest.mark.parametrize("run", ["asyoulik", "alice29", "enwik8_128kb_0"])
def test_schedule_matches_reference_traces(fixtures, run):
    syms = fixtures[f"run_{run}_syms"].astype(np.uint32)
    want = list(fixtures[f"run_{run}_reprimes"])
    gates = fixtures[f"run_{run}_gates"]
    n = len(syms)
    rng = np.random.default_rng(1)
    small_v = 64
    tab = rng.normal(0, 1, (16, small_v)).astype(np.float32)
    ids = np.concatenate([[0], syms % small_v]).astype(np.uint32)
    ev = _events_from_gates(gates, n)
    s = oracle.Session.table(tab)
    payload, rep_enc = s.encode_tokens(ids, events=ev)
    assert rep_enc == want, "encode-side context_reprime positions differ from the reference's trace"
    s2 = oracle.Session.table(tab)
    out, rep_dec = s2.decode_tokens(payload, 0, n, events=ev)
    assert rep_dec == want
    assert np.array_equal(out, ids)


def test_product_schedule_chunks_match_oracle_loop():
    import candlezip_b200 as cz
    from candlezip_b200 import _lib

    rng = np.random.default_rng(5)
    tab = rng.normal(0, 1, (8, 32)).astype(np.float32)
    for n, ctx_, R in [(1, 512, 512), (511, 512, 512), (512, 512, 512), (513, 512, 512), (5000, 512, 512), (3000, 512, 256), (2000, 600, 512),
                       (1500, 100, 64)]:
        ids = np.concatenate([[0], rng.integers(0, 32, n)]).astype(np.uint32)
        _, reprimes = oracle.Session.table(tab).encode_tokens(ids, context=ctx_, reprime_interval=R)
        cap = n + 8
        first = np.zeros(cap, np.uint64); nc = np.zeros(cap, np.uint32); ps = np.zeros(cap, np.uint64); pl = np.zeros(cap, np.uint32)
        k = _lib.lib.cz_schedule_chunks(n, ctx_, R, first.ctypes.data_as(_lib.u64p), nc.ctypes.data_as(_lib.u32p),
                                        ps.ctypes.data_as(_lib.u64p), pl.ctypes.data_as(_lib.u32p), cap)
        assert list(first[1:k]) == reprimes
        assert int(nc[:k].sum()) == n and first[0] == 0 and pl[0] == 1 and ps[0] == 0
        eff = min(ctx_, 511)
        for c in range(1, k):
            end = int(first[c]) + 1
            assert int(ps[c]) == max(0, end - eff) and int(pl[c]) == end - int(ps[c])


# ------------------------------------------------------------------ container + BLAKE3 vs the shipped .canz files
@pytest.mark.parametrize("name", ["asyoulik", "fields", "alice29", "enwik8_128kb_0"])
def test_container_headers_of_shipped_canz(fixtures, name):
    from candlezip_b200 import container

    hb = fixtures[f"canz_{name}_header"].tobytes()
    # oracle parser
    h = oracle.HeaderV2()
    ro = C.c_size_t()
    buf = (C.c_uint8 * len(hb)).from_buffer_copy(hb)
    n = oracle.lib.czo_read_header_v2(buf, len(hb), C.byref(h), C.byref(ro))
    assert n > 0 and h.vocab_size == 49152 and h.bos_token_id == 0 and h.context_window == 512 and h.reprime_interval == 512
    assert hb[ro.value : ro.value + h.model_file_repr_len] == b"model.safetensors"
    # product parser agrees field by field and re-emits identical bytes
    f, rep, gates, eng, st, pay = container.read_container(hb)
    for k in ("token_count", "orig_len_bytes", "reserved_flags", "vocab_size"):
        assert f[k] == getattr(h, k)
    again = container.write_container({**f, "reserved_flags": f["reserved_flags"] & ~(1 << 2)}, rep, [b""], gates=gates)
    assert again == hb
    if name != "alice29":  # alice29.canz was made from the CRLF original, not the shipped LF file
        assert f["orig_hash16"] == fixtures[f"canz_{name}_src_blake3_16"].tobytes()
        assert f["orig_len_bytes"] == int(fixtures[f"canz_{name}_src_size"])
        assert f["reserved_flags"] == 0x02000000
    tokens = {"asyoulik": 39915, "fields": 4324, "alice29": 41933, "enwik8_128kb_0": 34803}[name]
    assert f["token_count"] == tokens


def test_blake3_matches_reference_package():
    blake3 = pytest.importorskip("blake3")
    from candlezip_b200 import container

    rng = np.random.default_rng(0)
    for n in [0, 1, 63, 64, 65, 1023, 1024, 1025, 2048, 3072, 3073, 7 * 1024, 100_003]:
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert container.blake3_16(d) == blake3.blake3(d).digest()[:16]


def test_segment_container_roundtrip():
    from candlezip_b200 import container

    f = dict(token_count=1000, orig_len_bytes=1000, vocab_size=49152)
    pays = [b"abc", b"", b"defgh"]
    blob = container.write_container(f, b"model.safetensors", pays, seg_tokens=[400, 300, 300], engine=1)
    g, rep, gates, eng, st, pay = container.read_container(blob)
    assert pay == pays and list(st) == [400, 300, 300] and eng == 1 and g["reserved_flags"] & container.CZ_FLAG_SEGMENTS
    # a one-segment file without the extension is the reference layout: header | payload
    blob1 = container.write_container(f, b"model.safetensors", [b"xyz"])
    g1, _, _, _, st1, pay1 = container.read_container(blob1)
    assert st1 is None and pay1 == [b"xyz"] and not (g1["reserved_flags"] & container.CZ_FLAG_SEGMENTS)


# ------------------------------------------------------------------ C ABI surface
def test_abi_exports_every_declared_symbol():
    from candlezip_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "candlezip_b200.h")).read()
    declared = set(re.findall(r"\b(cz_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    missing = [s for s in sorted(declared) if not hasattr(_lib.lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert declared <= set(_lib.SIGNATURES), f"python binding lacks: {sorted(declared - set(_lib.SIGNATURES))}"
    assert _lib.lib.cz_abi_version() == 1


def test_no_cpu_fallback_without_device():
    """The product must fail loudly without a GPU (no oracle / CPU route)."""
    import candlezip_b200 as cz
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cz.CzError) as e:
        cz.Context(0)
    assert e.value.code == -2
    host = cz.Context(-1)
    with pytest.raises(cz.CzError) as e:
        host.cdf_bounds(np.zeros((8, 2), np.float32), [0, 1])
    assert e.value.code == -2


def test_product_never_links_the_oracle():
    so = os.path.join(ROOT, "candlezip_b200", "libcandlezip_b200.so")
    out = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    assert "czo_" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "candlezip_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, fn), errors="replace").read()
                assert "cz_oracle" not in txt and "import oracle" not in txt and "libcz_oracle" not in txt, fn


# ------------------------------------------------------------------ weight generator + oracle LLaMA
def test_random_init_is_deterministic_and_bf16():
    import candlezip_b200 as cz

    host = cz.Context(-1)
    a = cz.Model(host, cz.SMOLLM_TINY).random_init(3).tensors()
    b = cz.Model(host, cz.SMOLLM_TINY).random_init(3).tensors()
    c = cz.Model(host, cz.SMOLLM_TINY).random_init(4).tensors()
    assert all(np.array_equal(a[k], b[k]) for k in a)
    w = a["model.layers.1.mlp.up_proj.weight"]
    assert not np.array_equal(w, c["model.layers.1.mlp.up_proj.weight"])
    assert abs(w.std() - 0.02) < 0.002 and abs(w.mean()) < 0.001
    bits = w.view(np.uint32)
    assert np.all(bits & 0xFFFF == 0), "weights must be exactly bf16-representable"
    assert np.all(a["model.norm.weight"] == 1.0)


def _tiny_oracle(round_bf16=0, seed=11, embed_std=0.2):
    import candlezip_b200 as cz

    host = cz.Context(-1)
    m = cz.Model(host, cz.SMOLLM_TINY).random_init(seed, 0.05, embed_std)
    cfg = dict(cz.SMOLLM_TINY)
    cfg["rms_eps"] = cfg.pop("norm_eps")
    return oracle.Session.llama(cfg, m.tensors(), round_bf16=round_bf16), cfg


def test_oracle_llama_step_equals_prefill():
    """KV-cache stepping and a teacher-forced prefill give the same last-token logits (same arithmetic, f32)."""
    s, cfg = _tiny_oracle()
    rng = np.random.default_rng(2)
    toks = rng.integers(0, cfg["vocab"], 40).astype(np.uint32)
    a = s.reprime(toks)
    s2, _ = _tiny_oracle()
    b = None
    for t in toks:
        b = s2.step_logits(t)
    assert np.allclose(a, b, rtol=1e-4, atol=1e-5)
    assert s.index_pos() == 40 and s2.index_pos() == 40


def test_oracle_llama_roundtrip_with_reprimes():
    # flat-ish logits: with uniformly random tokens a peaky model hits symbols of mass < 2^-30, where the reference's
    # SmolLM path (no pdf floor) has a zero-width interval -- the oracle reports that as an error (see next test)
    s, cfg = _tiny_oracle(embed_std=0.05)
    rng = np.random.default_rng(9)
    n = 1300
    ids = np.concatenate([[0], rng.integers(0, cfg["vocab"], n)]).astype(np.uint32)
    payload, rep = s.encode_tokens(ids)
    assert rep == [512, 1024]
    s2, _ = _tiny_oracle(embed_std=0.05)
    out, rep2 = s2.decode_tokens(payload, 0, n)
    assert rep2 == rep and np.array_equal(out, ids)
    bits_per_token = 8 * len(payload) / n
    assert bits_per_token < np.log2(cfg["vocab"]) + 1.5  # uniform tokens under a non-uniform model cost a little over log2 V


def test_oracle_flags_zero_width_interval():
    s, cfg = _tiny_oracle(embed_std=0.5)  # very peaky: random tokens land on zero-width intervals
    rng = np.random.default_rng(9)
    ids = np.concatenate([[0], rng.integers(0, cfg["vocab"], 400)]).astype(np.uint32)
    with pytest.raises(ValueError):
        s.encode_tokens(ids)


def test_oracle_llama_matches_transformers_golden():
    """golden logits produced by transformers.LlamaForCausalLM (f32, eager) on the same seeded weights:
    tests/golden/make_llama_golden.py"""
    path = os.path.join(ROOT, "tests", "golden", "llama_tiny_golden.npz")
    z = np.load(path)
    s, cfg = _tiny_oracle(seed=int(z["seed"]), embed_std=float(z["embed_std"]))
    toks = z["tokens"].astype(np.uint32)
    got = s.reprime(toks)
    want = z["last_logits"]
    assert np.max(np.abs(got - want)) < 2e-4 * max(1.0, np.abs(want).max())
    # stepping one more token
    got2 = s.step_logits(int(z["next_token"]))
    assert np.max(np.abs(got2 - z["next_logits"])) < 2e-4 * max(1.0, np.abs(z["next_logits"]).max())


def test_oracle_xe_is_sum_of_neg_log2():
    s, cfg = _tiny_oracle()
    rng = np.random.default_rng(4)
    hist = rng.integers(0, cfg["vocab"], 30).astype(np.uint32)
    tg = rng.integers(0, cfg["vocab"], 5).astype(np.uint32)
    bits = s.xe_bits(hist, tg)
    s2, _ = _tiny_oracle()
    l = s2.reprime(hist)
    acc = 0.0
    for t in tg:
        p = oracle.softmax_pdf_floor(l)
        acc += -np.log2(max(p[t], 1e-300))
        l = s2.step_logits(t)
    assert abs(bits - acc) < 1e-9


# ------------------------------------------------------------------ RWKV-7 oracle
def _rwkv_tiny_oracle(seed=7, round_bf16=0):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from rwkv7_weights import RWKV7_TINY, make_weights

    return oracle.Session.rwkv7(RWKV7_TINY, make_weights(RWKV7_TINY, seed), round_bf16=round_bf16), RWKV7_TINY


def test_oracle_rwkv7_matches_fla_golden():
    """golden logits from an independent implementation assembled from flash-linear-attention's pure-torch reference
    functions (tests/golden/make_rwkv7_golden.py).  Criterion of the reference's own check
    (candle_rwkv7/compare_with_reference.py:80,130): max|delta| / std(ref) < 1e-5... stated here as 2e-5."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "rwkv7_tiny_golden.npz"))
    s, cfg = _rwkv_tiny_oracle(seed=int(z["seed"]))
    for i, t in enumerate(z["tokens"]):
        got = s.step_logits(int(t))
        want = z["logits"][i]
        assert np.abs(got - want).max() / want.std() < 2e-5, i


def test_oracle_rwkv7_reprime_replays_history_from_zero_state():
    """src/models.rs:162-170: reprime == fresh state + step over every history token."""
    s, cfg = _rwkv_tiny_oracle()
    rng = np.random.default_rng(3)
    toks = rng.integers(0, cfg["vocab"], 17).astype(np.uint32)
    for t in toks[:5]:
        s.step_logits(t)  # dirty the state first
    a = s.reprime(toks)
    s2, _ = _rwkv_tiny_oracle()
    b = None
    for t in toks:
        b = s2.step_logits(t)
    assert np.array_equal(a, b)
    assert s.index_pos() == 17


def test_oracle_rwkv7_roundtrip_with_literal_escapes():
    """RWKV coding loop: V+256 symbols, literal symbols (>= V) do not step the model (src/main.rs:2301-2326, 2347-2349,
    2812-2834)."""
    s, cfg = _rwkv_tiny_oracle()
    rng = np.random.default_rng(5)
    n = 300
    ids = rng.integers(0, cfg["vocab"], n + 1).astype(np.uint32)
    ids[0] = 0
    lit = rng.random(n + 1) < 0.1
    lit[0] = False
    ids[lit] = cfg["vocab"] + rng.integers(0, 256, int(lit.sum()))
    payload, rep = s.encode_tokens(ids, backend=1)
    assert rep == []  # no context re-prime for RWKV (src/main.rs:2275 guard)
    s2, _ = _rwkv_tiny_oracle()
    out, _ = s2.decode_tokens(payload, 0, n, backend=1)
    assert np.array_equal(out, ids)
    # a literal costs about -log2(2^-29) = 29 bits
    assert 8 * len(payload) > 29 * int(lit.sum())


# ------------------------------------------------------------------ agentic gate host logic
def test_gate_candidates_budgets_and_selection():
    from candlezip_b200 import gate

    c = gate.build_candidates("Alice met Bob\tin 1865.\nThe Rabbit ran")
    assert c[0] == "Alice met Bob in 1865. The Rabbit ran"       # control characters -> spaces (main.rs:2019)
    assert c[1] == c[0] and c[2] == c[0]                          # < 2000 bytes; the single line contains a digit
    assert c[3] == "Alice Bob The Rabbit"                         # Capitalised words (main.rs:2026-2030)
    assert gate.build_candidates("no digits here")[2] == ""
    assert len(gate.build_candidates("x" * 3000)[1]) == 2000
    assert gate.budgets(511) == [63, 127, 255, 383]               # main.rs:2034-2035

    ids = np.arange(1, 1301, dtype=np.uint32) % 250
    texts = {1: "Hint One 42", 2: ""}
    tok = lambda s, m: [ord(ch) for ch in s][:m]
    jobs, plan = gate.plan_scan(ids, texts, tok, agent_chunk=512, scan_lookahead=512)
    assert [e["i"] for e in plan] == [511, 1023] and plan[0]["chunk_index"] == 1
    # boundary 1: baseline + 16 conditioned; boundary 2: empty hint -> all conditioned slots reuse the baseline
    assert plan[0]["job0"] == 0 and all(s is not None for s in plan[0]["slots"]) and plan[1]["slots"] == [None] * 16
    prime, targets = jobs[0]
    assert len(prime) == 511 and prime[0] == 0 and len(targets) == 512 and targets[0] == ids[510]   # seq[0] = BOS
    assert list(jobs[1][0][-11:]) == [ord(ch) for ch in "Hint One 42"] and len(jobs[1][0]) == 511
    bits = np.full(len(jobs), 1000.0)
    bits[plan[0]["slots"][2 * 4 + 1]] = 900.0   # cand 2 / budget 1 saves the most
    bits[plan[0]["slots"][3 * 4 + 3]] = 950.0
    rec, ev, rows = gate.decide(plan, bits)
    assert rec == [1 | (2 << 1) | (1 << 3), 0 | (0 << 1) | (0 << 3)]  # no improvement -> first (cand 0, budget 0) wins with saved == 0, gate 0
    assert len(ev) == 1 and ev[0][0] == 511 and ev[0][2] == 511 + 512 and ev[0][3] == 511 - 127
    assert rows[0]["bits_saved"] == 100.0 and rows[1]["gate"] == 0
    # thresholds (main.rs:2069-2071)
    rec2, ev2, _ = gate.decide(plan, bits, thr_pct=20.0)
    assert rec2[0] & 1 == 0 and not ev2
    # decode side rebuilds the same events from the records
    hints = [e["hints"] for e in plan]
    ev3 = gate.events_from_records(rec, hints, 512, len(ids))
    assert len(ev3) == 1 and ev3[0][0] == 511 and np.array_equal(ev3[0][1], ev[0][1]) and ev3[0][3] == ev[0][3]


# ------------------------------------------------------------------ safetensors loader (SmolLmSession::load, src/models.rs:48-61)
def test_safetensors_loader_roundtrip(tmp_path):
    """write HF-named tensors with the `safetensors` package (F32, BF16 and F16 files, a tied lm_head and an unknown tensor),
    load them through cz_model_load_safetensors on a host-only ctx and read them back"""
    import candlezip_b200 as cz
    import torch
    from safetensors.torch import save_file

    host = cz.Context(-1)
    ref = cz.Model(host, cz.SMOLLM_TINY).random_init(3, 0.05, 0.1)
    want = ref.tensors()
    shapes = {}
    for name, n in ref.tensor_names():
        if name.endswith("norm.weight"):
            shapes[name] = (n,)
        elif "embed_tokens" in name:
            shapes[name] = (cz.SMOLLM_TINY["vocab"], cz.SMOLLM_TINY["d_model"])
        else:
            shapes[name] = (n // cz.SMOLLM_TINY["d_model"], cz.SMOLLM_TINY["d_model"]) if "down_proj" not in name else (cz.SMOLLM_TINY["d_model"], n // cz.SMOLLM_TINY["d_model"])
    names = list(want)
    parts = [names[0::3], names[1::3], names[2::3]]
    dtypes = [torch.float32, torch.bfloat16, torch.float16]
    paths = []
    for k, (part, dt) in enumerate(zip(parts, dtypes)):
        tensors = {nm: torch.from_numpy(want[nm].reshape(shapes[nm]).copy()).to(dt) for nm in part}
        if k == 0:
            tensors["some.unknown.tensor"] = torch.zeros(3)
        p = str(tmp_path / f"model-{k}.safetensors")
        save_file(tensors, p)
        paths.append(p)
    m = cz.Model(host, cz.SMOLLM_TINY).load_safetensors(paths)
    got = m.tensors()
    for nm in names:
        if nm in parts[2]:  # went through f16: bf16-exact values are not all f16-exact
            assert np.allclose(got[nm], want[nm], rtol=2e-3, atol=1e-4), nm
        else:
            assert np.array_equal(got[nm], want[nm]), nm
    with pytest.raises(cz.CzError):
        cz.Model(host, cz.SMOLLM_TINY).load_safetensors([str(tmp_path / "missing.safetensors")])


# ------------------------------------------------------------------ tokenisation glue
def _toy_rwkv_vocab():
    toks = {bytes([b]): b + 1 for b in range(256) if b not in (0xFF,)}  # 0xFF has no token: forces the literal path
    extra = [b"th", b"the", b"the ", b"he", b"in", b"ing", b"\xe2\x82\xac", b"ab", b"abc", b"abcd"]
    for k, t in enumerate(extra):
        toks[t] = 300 + k
    return toks


def test_rwkv_trie_tokenizer_semantics():
    from candlezip_b200 import codec

    tok = codec.RwkvTokenizer(_toy_rwkv_vocab())
    # first match in DESCENDING id order among tokens sharing the first two bytes (rwkv7.rs:566-573, 584-589):
    # "the " (302) beats "the" (301) beats "th" (300); "abcd" (309) beats "abc" / "ab"
    assert list(tok.encode_bytes(b"the cat")) == [302, ord("c") + 1, ord("a") + 1, ord("t") + 1]
    assert list(tok.encode_bytes(b"then")) == [301, ord("n") + 1]
    assert list(tok.encode_bytes(b"abcab")) == [308, 307]
    data = "the thing in the € abcd ab".encode()
    ids = tok.encode_bytes(data)
    assert tok.decode_bytes(ids) == data
    with pytest.raises(KeyError):
        tok.encode_bytes(b"x\xffy")
    # plan_rwkv_symbols: a vocabulary gap makes EVERY byte a literal escape (main.rs:857-862)
    V = 320
    sym = codec.plan_rwkv_symbols(b"x\xffy", tok, V)
    assert list(sym) == [V + ord("x"), V + 0xFF, V + ord("y")]
    assert codec.rwkv_detok_with_literals(tok, sym, V) == b"x\xffy"
    mixed = np.array([302, V + 0xFF, 308, V + 0], np.uint32)
    assert codec.rwkv_detok_with_literals(tok, mixed, V) == b"the \xffabc\x00"
    assert codec.ByteTokenizer().decode_bytes(codec.ByteTokenizer().encode_bytes(bytes(range(256)))) == bytes(range(256))
