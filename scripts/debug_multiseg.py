import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import candlezip_b200 as cz
ctx = cz.Context(0)
model = cz.Model(ctx, cz.SMOLLM_TINY).random_init(5, 0.05, 0.05)
rng = np.random.default_rng(2303)
ids = rng.integers(0, 1024, 2300).astype(np.uint32)
pays, seg = model.encode(ids, n_segments=3)
for g in range(3):
    a, b = int(seg[g]), int(seg[g + 1])
    solo, s1 = model.encode(ids[a:b], n_segments=1)
    print("seg", g, "len", b - a, "encode multi==solo:", solo[0] == pays[g], len(solo[0]), len(pays[g]))
    out1 = model.decode([pays[g]], s1)
    ok1 = np.array_equal(out1, ids[a:b])
    print("   solo decode ok:", ok1, "" if ok1 else int(np.argmax(out1 != ids[a:b])))
out = model.decode(pays, seg)
for g in range(3):
    a, b = int(seg[g]), int(seg[g + 1])
    ok = np.array_equal(out[a:b], ids[a:b])
    print("multi decode seg", g, ok, "" if ok else int(np.argmax(out[a:b] != ids[a:b])))
