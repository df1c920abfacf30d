"""Shared helpers of the deep-parity GPU tests and scripts/parity_report.py (TEST INFRASTRUCTURE: uses the oracle).

Reference semantics being checked: SmolLmSession::{step_logits_tensor, reprime_with_history_and_get_last_logits_tensor}
(src/models.rs:92-119) -> candle-transformers llama forward, restated in oracle/cz_models.c."""
import numpy as np

import oracle


def bf16_bits(a):
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return u.astype(np.uint16)


def bf16_to_f32(b):
    return (np.asarray(b).astype(np.uint32) << 16).view(np.float32)


def oracle_llama(model, round_bf16, max_pos=1100):
    cfg = dict(model.cfg)
    cfg["rms_eps"] = cfg.pop("norm_eps")
    return oracle.Session.llama(cfg, model.tensors(), round_bf16=round_bf16, max_pos=max_pos)


def oracle_chunk_logits(orc, prime, targets):
    """logits[j] = the oracle's logits after prime ++ targets[:j] (what codes targets[j]): reprime then steps."""
    out = [orc.reprime(prime)]
    for t in targets[:-1]:
        out.append(orc.step_logits(int(t)))
    return np.stack(out)


def logits_parity(got, want, scale_from=None):
    """per position: max|got - want| / std(reference logits of that position)"""
    ref = want if scale_from is None else scale_from
    sd = ref.std(axis=1).astype(np.float64)
    return np.abs(got.astype(np.float64) - want.astype(np.float64)).max(axis=1) / np.maximum(sd, 1e-30)


def sample_from_oracle(orc, n, rng, bos=0, context=512, reprime_interval=512, vocab=None):
    """Runs the reference's encode loop on the oracle (first logits from BOS, src/main.rs:1916; context re-prime every
    `reprime_interval` tokens on the last 511, :2275-2290; step after each token, :2344-2350) and draws every token from the
    model's own softmax: a text the model predicts well, so bits/token << log2 V and a logits error shows up as size.
    Returns the n coded tokens (no BOS)."""
    seq = [bos]
    logits = orc.step_logits(bos)
    max_ctx = min(context, 511)
    pos = 1
    for i in range(n):
        if pos >= max_ctx and i % reprime_interval == 0 and i > 0:
            logits = orc.reprime(np.asarray(seq[i + 1 - max_ctx:i + 1], np.uint32))
            pos = max_ctx
        z = logits.astype(np.float64)
        p = np.exp(z - z.max())
        p /= p.sum()
        s = int(rng.choice(len(p), p=p))
        seq.append(s)
        logits = orc.step_logits(s)
        pos += 1
    return np.asarray(seq[1:], np.uint32)


def attention_ref_f64(q, k, v, nh, nkv):
    """causal GQA attention of one sequence in f64: q [n][nh*64], k / v [n][nkv*64] (f32 views of the bf16 inputs)"""
    n = q.shape[0]
    g = nh // nkv
    q = q.astype(np.float64).reshape(n, nh, 64)
    k = k.astype(np.float64).reshape(n, nkv, 64)
    v = v.astype(np.float64).reshape(n, nkv, 64)
    out = np.zeros((n, nh, 64))
    mask = np.tril(np.ones((n, n), bool))
    for h in range(nh):
        s = q[:, h] @ k[:, h // g].T / 8.0
        s = np.where(mask, s, -np.inf)
        s -= s.max(axis=1, keepdims=True)
        p = np.exp(s)
        p /= p.sum(axis=1, keepdims=True)
        out[:, h] = p @ v[:, h // g]
    return out.reshape(n, nh * 64)
