#!/usr/bin/env python3
"""Golden logits for the oracle's LLaMA restatement, produced by an INDEPENDENT implementation:
transformers.LlamaForCausalLM (f32, eager attention) on the product's seeded random-init weights (SMOLLM_TINY).
candle-transformers' llama.rs is not under /root/reference, so this is the available cross-check.
    python tests/golden/make_llama_golden.py      # writes tests/golden/llama_tiny_golden.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import candlezip_b200 as cz  # noqa: E402
from transformers import LlamaConfig, LlamaForCausalLM  # noqa: E402

SEED, EMBED_STD = 11, 0.2


def main():
    c = cz.SMOLLM_TINY
    host = cz.Context(-1)
    m = cz.Model(host, c).random_init(SEED, 0.05, EMBED_STD)
    tensors = m.tensors()
    hf_cfg = LlamaConfig(vocab_size=c["vocab"], hidden_size=c["d_model"], intermediate_size=c["d_ffn"], num_hidden_layers=c["n_layers"],
                         num_attention_heads=c["n_heads"], num_key_value_heads=c["n_kv_heads"], rms_norm_eps=c["norm_eps"],
                         rope_theta=c["rope_theta"], max_position_embeddings=2048, tie_word_embeddings=True, hidden_act="silu",
                         attention_bias=False, mlp_bias=False, attn_implementation="eager")
    model = LlamaForCausalLM(hf_cfg).to(torch.float32).eval()
    sd = model.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    new = {}
    for name, arr in tensors.items():
        new[name] = torch.from_numpy(arr.reshape(shapes[name]).copy())
    new["lm_head.weight"] = new["model.embed_tokens.weight"]
    missing = model.load_state_dict(new, strict=False)
    assert not [k for k in missing.missing_keys if "rotary" not in k], missing
    rng = np.random.default_rng(123)
    toks = rng.integers(0, c["vocab"], 48)
    nxt = int(rng.integers(0, c["vocab"]))
    with torch.no_grad():
        out = model(torch.tensor(toks[None, :]))
        last = out.logits[0, -1].numpy()
        out2 = model(torch.tensor(np.concatenate([toks, [nxt]])[None, :]))
        nxt_logits = out2.logits[0, -1].numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "llama_tiny_golden.npz"), seed=SEED, embed_std=EMBED_STD,
                        tokens=toks.astype(np.uint32), next_token=nxt, last_logits=last.astype(np.float32),
                        next_logits=nxt_logits.astype(np.float32))
    print("wrote llama_tiny_golden.npz; logits range", last.min(), last.max())


if __name__ == "__main__":
    main()
