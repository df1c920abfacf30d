"""Per-kernel-family device time of the stepwise decoders (eager launches, deferred event pairs).  GPU box only."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import candlezip_b200 as cz  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "smollm"
n, segs = int(sys.argv[2]) if len(sys.argv) > 2 else 16384, int(sys.argv[3]) if len(sys.argv) > 3 else 256
ctx = cz.Context(0)
cfg = cz.SMOLLM_135M if arch == "smollm" else cz.RWKV7_0P1B
model = cz.Model(ctx, cfg).random_init(0, 0.02, 0.02)
rng = np.random.default_rng(0)
if len(sys.argv) > 4 and sys.argv[4] == "corpus":  # real bytes, ids spread over the vocabulary like real token ids (the CDF search walks further)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import corpus

    d = corpus.load("enwik8_3mib")
    ids = corpus.byte_ids((d * (n // len(d) + 1))[:n], cfg["vocab"], True)
else:
    ids = rng.integers(97, 123, n).astype(np.uint32)
pays, seg = model.encode(ids, n_segments=segs)
model.decode(pays, seg)  # warm (allocations)
res = {"arch": arch, "tokens": n, "segments": segs, "ids": sys.argv[4] if len(sys.argv) > 4 else "a-z"}
for graph in (1, 0):
    os.environ.pop("CZ_DECODE_NO_GRAPH", None)
    if not graph:
        os.environ["CZ_DECODE_NO_GRAPH"] = "1"
        ctx.profile(2)
        ctx.profile_read(reset=True)
    t0 = time.perf_counter()
    out = model.decode(pays, seg)
    dt = time.perf_counter() - t0
    assert np.array_equal(out, ids)
    res["graph" if graph else "eager"] = {"s": round(dt, 4), "tok_per_s": round(n / dt), "ms_per_step": round(1e3 * dt / (n / segs), 3)}
    if not graph:
        fam = ctx.profile_read(reset=True)
        steps = n / segs
        res["family_ms_per_step"] = {k: round(v[0] / steps, 4) for k, v in fam.items()}
        res["family_launches_per_step"] = {k: round(v[1] / steps, 1) for k, v in fam.items()}
print(json.dumps(res))
