import sys, os, numpy as np
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
import candlezip_b200 as cz
ctx=cz.Context(0)
m=cz.Model(ctx, cz.SMOLLM_TINY).random_init(5,0.05,0.05)
rng=np.random.default_rng(41)
ids = rng.integers(0, 1024, 1500).astype(np.uint32)
ev = [(200, rng.integers(0, 1024, 300).astype(np.uint32), 200 + 512), (900, rng.integers(0, 1024, 63).astype(np.uint32), 900 + 512)]
def tryenc(name, **kw):
    try:
        p,s=m.encode(ids, n_segments=1, **kw); print(name, "ok", len(p[0])); return p
    except Exception as e: print(name, "FAIL", str(e)[:140])
seq=np.concatenate([[0],ids]).astype(np.uint32)
def chunk1(): return m.chunk_logits(seq[2:513], ids[512:1024])
if len(sys.argv)>1 and sys.argv[1]=="a":
    l0=chunk1()
    tryenc("both", events=ev)
    l1=chunk1(); print("chunk1 logits equal after events-encode:", np.array_equal(l0.view(np.uint32), l1.view(np.uint32)), np.isnan(l1).any())
    tryenc("plain")
    tryenc("plain again")
else:
    tryenc("both", events=ev)
    tryenc("plain")
    tryenc("plain again")
    l1=chunk1(); print("nan", np.isnan(l1).any())
