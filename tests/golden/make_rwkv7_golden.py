#!/usr/bin/env python3
"""Generate tests/golden/rwkv7_tiny_golden.npz: RWKV-7 logits of a short token sequence on the deterministic test
weights (rwkv7_weights.py), computed by an INDEPENDENT implementation assembled from flash-linear-attention's pure-torch
reference functions (fla 0.5.1, importable on CPU):
    fla.ops.rwkv7.fused_addcmul.torch_addcmul_rwkv7          token-shift mixes
    fla.ops.generalized_delta_rule.dplr.naive.dplr_recurrence  S_t = S_t diag(w) + (S_t a) b^T + v k^T with a=-kk, b=kk*a
    fla.ops.rwkv7.fused_k_update.k_update_ref                k * (1 + (a-1) k_a)
    fla.ops.rwkv7.gate_output_correction.gate_output_correction_ref   (o + (r k r_k).sum * v) * g
    fla.ops.rwkv7.channel_mixing.rwkv_mix_torch / rwkv_relu_and_square_torch
wired together the way fla/layers/rwkv7.py:233-350 and fla/models/rwkv7 do (w = -0.6065306597126334 * sigmoid(lora) as the
log decay; GroupNorm eps = head_dim * norm_eps = 64e-5).  The oracle (oracle/cz_rwkv7.c, restated from
candle_rwkv7/src/models/rwkv7.rs) and the CUDA path are both checked against these logits.

Run in the build container:  python tests/golden/make_rwkv7_golden.py
"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from rwkv7_weights import RWKV7_TINY, make_weights  # noqa: E402

from fla.ops.generalized_delta_rule.dplr.naive import dplr_recurrence  # noqa: E402
from fla.ops.rwkv7.channel_mixing import rwkv_mix_torch, rwkv_relu_and_square_torch  # noqa: E402
from fla.ops.rwkv7.fused_addcmul import torch_addcmul_rwkv7  # noqa: E402
from fla.ops.rwkv7.fused_k_update import k_update_ref  # noqa: E402
from fla.ops.rwkv7.gate_output_correction import gate_output_correction_ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rwkv7_tiny_golden.npz")


def reference_logits(cfg, W, tokens):
    """All positions at once (T tokens, zero initial state): logits [T, V] float32."""
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).float() for k, v in W.items()}
    C, N = cfg["d_model"], cfg["head_dim"]
    H = C // N
    T = len(tokens)
    x = t["model.embeddings.weight"][torch.as_tensor(np.asarray(tokens, np.int64))].unsqueeze(0)  # [1,T,C]
    v_first = None
    zero = torch.zeros(1, C)
    for l in range(cfg["n_layers"]):
        p = f"model.layers.{l}."
        a_ = p + "attn."
        if l == 0:
            x = F.layer_norm(x, (C,), t[p + "pre_norm.weight"], t[p + "pre_norm.bias"], cfg["norm_eps"])
        h = F.layer_norm(x, (C,), t[p + "attn_norm.weight"], t[p + "attn_norm.bias"], cfg["norm_eps"])
        delta = torch.cat((zero.unsqueeze(1), h[:, :-1]), dim=1) - h
        xr, xw, xk, xv, xa, xg = torch_addcmul_rwkv7(h, delta, *(t[a_ + n].view(1, 1, C) for n in ("x_r", "x_w", "x_k", "x_v", "x_a", "x_g")))
        r = xr @ t[a_ + "r_proj.weight"].T
        w_l = torch.tanh(xw @ t[a_ + "w_lora.lora.0.weight"].T) @ t[a_ + "w_lora.lora.2.weight"].T + t[a_ + "w_lora.lora.2.bias"]
        w = -0.6065306597126334 * torch.sigmoid(w_l)  # log decay
        k = xk @ t[a_ + "k_proj.weight"].T
        v = xv @ t[a_ + "v_proj.weight"].T
        if l == 0:
            v_first = v
        else:
            nu = torch.sigmoid((xv @ t[a_ + "v_lora.lora.0.weight"].T) @ t[a_ + "v_lora.lora.2.weight"].T + t[a_ + "v_lora.lora.2.bias"])
            v = torch.lerp(v, v_first, nu)
        a = torch.sigmoid((xa @ t[a_ + "a_lora.lora.0.weight"].T) @ t[a_ + "a_lora.lora.2.weight"].T + t[a_ + "a_lora.lora.2.bias"])
        g = torch.sigmoid(xg @ t[a_ + "g_lora.lora.0.weight"].T) @ t[a_ + "g_lora.lora.2.weight"].T
        kk = F.normalize((k * t[a_ + "k_k"]).view(1, T, H, N), dim=-1, p=2.0)
        k = k_update_ref(k, a, t[a_ + "k_a"])
        r4, w4, k4, a4, v4 = (z.view(1, T, H, N) for z in (r, w, k, a, v))
        tr = lambda z: z.transpose(1, 2).contiguous()  # [B,H,T,N]
        # dplr_recurrence scales q by d_k^-0.5 (= 1/8, exact): pre-multiply by 8 for scale = 1
        o, _ = dplr_recurrence(tr(r4) * 8.0, tr(k4), tr(v4), tr(-kk), tr(kk * a4), tr(w4))
        o = o.transpose(1, 2).reshape(T, C)
        o = F.group_norm(o, H, t[a_ + "g_norm.weight"], t[a_ + "g_norm.bias"], eps=N * cfg["norm_eps"]).view(1, T, C)
        o = gate_output_correction_ref(o, r4, k4, t[a_ + "r_k"].view(H, N), v4, g)
        x = x + o @ t[a_ + "o_proj.weight"].T
        h = F.layer_norm(x, (C,), t[p + "ffn_norm.weight"], t[p + "ffn_norm.bias"], cfg["norm_eps"])
        kmix = rwkv_mix_torch(h, zero, t[p + "ffn.x_k"].view(1, 1, C))
        x = x + rwkv_relu_and_square_torch(kmix @ t[p + "ffn.key.weight"].T) @ t[p + "ffn.value.weight"].T
    x = F.layer_norm(x, (C,), t["model.norm.weight"], t["model.norm.bias"], cfg["norm_eps"])
    return (x[0] @ t["lm_head.weight"].T).numpy().astype(np.float32)


def main():
    cfg = RWKV7_TINY
    W = make_weights(cfg, seed=7)
    rng = np.random.default_rng(11)
    tokens = rng.integers(0, cfg["vocab"], 24).astype(np.uint32)
    logits = reference_logits(cfg, W, tokens)
    np.savez_compressed(OUT, tokens=tokens, logits=logits, seed=np.array(7))
    print("wrote", OUT, logits.shape, float(logits.std()))


if __name__ == "__main__":
    main()
