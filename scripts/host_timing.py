"""diagnostic: host-side time of cz_encode_dev with and without the nvidia-smi sampler running"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import candlezip_b200 as cz
from candlezip_b200 import _lib
import bench
ctx = cz.Context(0)
model = cz.Model(ctx, cz.SMOLLM_135M).random_init(0, 0.02, 0.02)
n = 262144
for segs in (8, 32):
    ids = torch.from_numpy(bench.synth_tokens(n, 1).astype(np.int32)).cuda()
    seg = cz.split_segments(n, segs)
    sched, keep = model._schedule(n, seg, 0, 512, 512, None, 0)
    cap = 4 * n + 8 * segs + 16
    out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    off = np.zeros(segs + 1, np.uint64)
    def step():
        _lib.check(_lib.lib.cz_encode_dev(model._h, C.c_void_p(ids.data_ptr()), n, C.byref(sched), C.c_void_p(out.data_ptr()), cap, off.ctypes.data_as(_lib.u64p)))
    for _ in range(2): step()
    for sampler in (False, True):
        s = None
        if sampler:
            s = bench.ClockSampler(0); s.start()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): step()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        if s: s.stop()
        print(f"segments={segs} sampler={sampler}: {dt*1e3:.1f} ms/step", flush=True)
