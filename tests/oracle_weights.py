"""numpy restatement of cz_model_random_init (candlezip_b200/csrc/model_core.cu) for the SmolLM architecture: the seeded
counter-hash weights the product generates, produced WITHOUT loading the product library -- bench.py's reference arm must
run the CPU oracle alone.  tests/test_cpu_oracle.py asserts bit-equality with the product's generator."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
SMOLLM_135M = dict(vocab=49152, d_model=576, n_layers=30, n_heads=9, n_kv_heads=3, head_dim=64, d_ffn=1536, rms_eps=1e-5, rope_theta=1e5)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def _fnv1a64(s: str) -> int:
    h = 0xCBF29CE484222325
    for c in s.encode():
        h = ((h ^ c) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def _bf16_round(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32)
    u = (u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)
    return u.view(np.float32)


def tensor(name: str, n: int, seed: int, sd: float) -> np.ndarray:
    k = np.float32(np.float32(sd) * np.float32(1.7320508075688772) / np.float32(65536.0))
    ts = _splitmix64(np.uint64(seed ^ _fnv1a64(name)))
    with np.errstate(over="ignore"):
        h = _splitmix64((ts + np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) & _M)
    s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)) + ((h >> np.uint64(32)) & np.uint64(0xFFFF)) +
         ((h >> np.uint64(48)) & np.uint64(0xFFFF))).astype(np.int64)
    return _bf16_round((s - 131070).astype(np.float32) * k)


def smollm_names(cfg):
    D, F, V, kvd = cfg["d_model"], cfg["d_ffn"], cfg["vocab"], cfg["n_kv_heads"] * cfg["head_dim"]
    out = [("model.embed_tokens.weight", V * D)]
    for l in range(cfg["n_layers"]):
        p = f"model.layers.{l}."
        out += [(p + "input_layernorm.weight", D), (p + "self_attn.q_proj.weight", D * D), (p + "self_attn.k_proj.weight", kvd * D),
                (p + "self_attn.v_proj.weight", kvd * D), (p + "self_attn.o_proj.weight", D * D), (p + "post_attention_layernorm.weight", D),
                (p + "mlp.gate_proj.weight", F * D), (p + "mlp.up_proj.weight", F * D), (p + "mlp.down_proj.weight", D * F)]
    out.append(("model.norm.weight", D))
    return out


def smollm_random_init(cfg, seed=0, std=0.02, embed_std=None):
    """{tensor name: float32 array holding exactly the bf16 values} -- what Model.random_init(seed, std, embed_std).tensors() returns"""
    embed_std = std if embed_std is None else embed_std
    w = {}
    for name, n in smollm_names(cfg):
        if "norm" in name:
            w[name] = np.ones(n, np.float32)
        else:
            w[name] = tensor(name, n, seed, embed_std if "embed" in name else std)
    return w
