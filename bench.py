#!/usr/bin/env python3
"""bench.py -- the hot path's headline metric on B200: encode MB/s, SmolLM-135M, ctx 512 / reprime 512.

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port) on host cores

A "step" is one pass of the hot path over one batch of synthetic input per GPU: 262,144 byte-level tokens
(id = byte, 1 B/token -- no tokenizer.json exists offline) coded as 32 independent segments = 512 reprime-chunks
(BASELINE.json configs[1]: "SmolLM-135M batched 512 chunks ... on 1xB200"), seeded random-init weights of the
SmolLM2-135M architecture.  `value` is timed with CUDA events on the library's stream with the token ids already
in HBM (cz_encode_dev); `e2e` goes through the public host-buffer call (cz_encode: pinned host ids in, payload out).
Chunks are independent, so N GPUs each take their own batch with no collective on the data path (weak scaling).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "encode_MB_per_s_smollm135m_ctx512"
# per layer-wave launch (253k rows), dram read + written under ncu --set full (profiles/ncu_summary_r01g.md): qkv+rope 769 MB,
# o-proj (+ fused norm operands) 1969 MB, gate/up 1061 MB, down (+ fused norm operands) 2335 MB; x60 launches each, plus the
# LM head's 51.5 GB of logits written + 0.4 GB read
NCU_GEMM_TRAFFIC_BYTES_PER_STEP = int(60 * (759 + 1969 + 1061 + 2335) * 1e6 + 51.9e9)
MFLOP_PER_TOKEN = 551.0  # trunk 423.84 + head 56.62 + attention 70.57 MFLOP per coded token at ctx 512 / reprime 512
TRUNK_PARAMS = 106_168_320  # matmul params per token position, SURVEY 8d
HEAD_PARAMS = 28_311_552


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def synth_tokens(n, seed):
    """byte-level 'text': printable ASCII with word structure (content does not affect throughput)"""
    rng = np.random.default_rng(seed)
    b = rng.integers(97, 123, n).astype(np.uint32)
    b[rng.random(n) < 0.17] = 32
    return b


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1403.8), d.get("hbm_gbs", 6553.9), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons DURING the timed region"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def oracle_chunk_sample(coded, seed=0):
    """Times the oracle (CPU restatement of the reference) on ONE steady-state reprime chunk: a (coded-1)-token prefill
    followed by `coded` coded tokens (src/main.rs:2280-2350).  Returns (seconds, tokens, cores)."""
    import candlezip_b200 as cz
    import oracle

    host = cz.Context(-1)
    m = cz.Model(host, cz.SMOLLM_135M).random_init(0, 0.02, 0.02)
    cfg = dict(cz.SMOLLM_135M)
    cfg["rms_eps"] = cfg.pop("norm_eps")
    sess = oracle.Session.llama(cfg, m.tensors(), round_bf16=0, max_pos=1100)
    m.close()
    toks = synth_tokens(2 * coded, seed)
    prime, targets = toks[: coded - 1], toks[coded - 1 : 2 * coded - 1]
    v = cfg["vocab"]
    cdf = np.empty(v + 1, np.uint32)
    cdfp = cdf.ctypes.data_as(oracle.u32p)

    def one_chunk():
        t0 = time.perf_counter()
        enc = oracle.lib.czo_encoder_new()
        pa = np.ascontiguousarray(prime, np.uint32)
        logits = oracle.lib.czo_session_reprime(sess._h, pa.ctypes.data_as(oracle.u32p), len(pa))
        for s in targets:
            oracle.lib.czo_logits_to_cdf(logits, v, 0, cdfp)
            oracle.lib.czo_encoder_encode_counts(enc, int(cdf[s]), int(cdf[s + 1]), 1 << 30)
            logits = oracle.lib.czo_session_step_logits(sess._h, int(s))
        n = C.c_size_t()
        oracle.lib.czo_encoder_finish(enc, C.byref(n))
        oracle.lib.czo_encoder_free(enc)
        return time.perf_counter() - t0

    return one_chunk, len(targets), os.cpu_count()


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    coded = args.ref_chunk
    one_chunk, n_tok, cores = oracle_chunk_sample(coded)
    for _ in range(args.warmup if args.warmup < 2 else 1):  # CPU path has no warm-up effects worth minutes of wall clock
        one_chunk()
    times = [one_chunk() for _ in range(args.steps)]
    t = sum(times)
    mbps = n_tok * args.steps / t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": mbps, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SmolLM-135M (random-init), ctx 512 / reprime 512, byte-level tokens", "sample_tokens_per_step": n_tok},
        "cpu_baseline": {"value": mbps, "unit": "MB/s", "cores": cores, "kind": "port",
                         "sample": f"one reprime chunk per step: {coded - 1}-token prefill + {n_tok} coded tokens, oracle (C restatement, OpenMP)"},
        "e2e": {"value": mbps, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "tokens_per_s": n_tok * args.steps / t,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import candlezip_b200 as cz
    from candlezip_b200 import _lib

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on STDOUT when the first communicator is created; stdout must carry exactly one JSON
        # line, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = cz.Context(local)
    model = cz.Model(ctx, cz.SMOLLM_135M).random_init(0, 0.02, 0.02)
    n = args.tokens
    ids_host = torch.empty(n, dtype=torch.int32).pin_memory()
    ids_np = ids_host.numpy().view(np.uint32)
    ids_np[:] = synth_tokens(n, 1234 + rank)
    seg_start = cz.split_segments(n, args.segments)
    sched, keep = model._schedule(n, seg_start, 0, 512, 512, None, 0)
    S = int(sched.n_segments)
    cap = 4 * n + 8 * S + 16
    # schedule-derived work (for FLOP accounting)
    first = np.zeros(4096, np.uint64); nc = np.zeros(4096, np.uint32); ps = np.zeros(4096, np.uint64); pl = np.zeros(4096, np.uint32)
    rows = 0
    n_chunks = 0
    for g in range(S):
        k = _lib.lib.cz_schedule_chunks(int(seg_start[g + 1] - seg_start[g]), 512, 512, first.ctypes.data_as(_lib.u64p),
                                        nc.ctypes.data_as(_lib.u32p), ps.ctypes.data_as(_lib.u64p), pl.ctypes.data_as(_lib.u32p), 4096)
        rows += int((pl[:k].astype(np.int64) + nc[:k].astype(np.int64) - 1).sum())
        n_chunks += int(k)
    gemm_flops = 2.0 * TRUNK_PARAMS * rows + 2.0 * HEAD_PARAMS * n

    # ---- device-resident arm (`value`) ----
    ids_dev = torch.empty(n, dtype=torch.int32, device="cuda")
    ids_dev.copy_(ids_host)
    out_dev = torch.empty(cap, dtype=torch.uint8, device="cuda")
    seg_off = np.zeros(S + 1, np.uint64)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr())

    def step_dev():
        _lib.check(_lib.lib.cz_encode_dev(model._h, C.c_void_p(ids_dev.data_ptr()), n, C.byref(sched), C.c_void_p(out_dev.data_ptr()), cap,
                                          seg_off.ctypes.data_as(_lib.u64p)))

    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.profile(2)
    ctx.profile_read(reset=True)
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_dev()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - l0
    fam = ctx.profile_read(reset=True)
    ctx.profile(0)
    clocks = sampler.stop()
    payload_bytes = int(seg_off[S])
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * args.steps / (ms_max / 1e3) / 1e6

    # ---- end-to-end arm through the public host-buffer API ----
    out_host = torch.empty(cap, dtype=torch.uint8).pin_memory()
    seg_off2 = np.zeros(S + 1, np.uint64)
    bs = _lib.Bitstreams(C.cast(out_host.data_ptr(), _lib.u8p), cap, seg_off2.ctypes.data_as(_lib.u64p))
    idp = C.cast(ids_host.data_ptr(), _lib.u32p)

    def step_e2e():
        _lib.check(_lib.lib.cz_encode(model._h, idp, n, C.byref(sched), C.byref(bs)))

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / float(t.item()) / 1e6
    same = bool(np.array_equal(seg_off, seg_off2)) and bool(
        torch.equal(out_dev[:payload_bytes].cpu(), out_host[:payload_bytes]))

    # ---- decode (stepwise, lock-step streams): extra figure, one pass ----
    decode = None
    if args.decode_tokens > 0:
        nd, sd = args.decode_tokens, args.decode_segments
        ids_d = synth_tokens(nd, 99 + rank)
        pays, seg_d = model.encode(ids_d, n_segments=sd)
        barrier()
        t0 = time.perf_counter()
        out = model.decode(pays, seg_d)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        decode = {"value": world * nd / float(t.item()) / 1e6, "unit": "MB/s", "tokens": nd, "segments": sd,
                  "roundtrip_ok": bool(np.array_equal(out, ids_d)), "timing": "wall clock around cz_decode (host payload in, ids out)"}

    # ---- RWKV-7 0.1B (BASELINE config 3), extra figures: slab-fused encode + lock-step decode, random-init weights ----
    rwkv = None
    if args.rwkv_tokens > 0:
        model.close()
        rmodel = cz.Model(ctx, cz.RWKV7_0P1B).random_init(0, 0.02, 0.02)
        nr = args.rwkv_tokens
        ids_r = synth_tokens(nr, 7 + rank)
        rmodel.encode(ids_r, n_segments=args.rwkv_segments)  # warm-up at full size (grow-only buffers, function attributes)
        barrier()
        ctx.profile(2)
        ctx.profile_read(reset=True)
        t0 = time.perf_counter()
        pays_r, seg_r = rmodel.encode(ids_r, n_segments=args.rwkv_segments)
        barrier()
        enc_s = time.perf_counter() - t0
        rfam = ctx.profile_read(reset=True)
        ctx.profile(0)
        nd = min(nr, args.rwkv_decode_tokens)
        pays_d, seg_d = rmodel.encode(ids_r[:nd], n_segments=args.rwkv_decode_segments)
        barrier()
        t0 = time.perf_counter()
        out_r = rmodel.decode(pays_d, seg_d)
        barrier()
        dec_s = time.perf_counter() - t0
        tt = torch.tensor([enc_s, dec_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        rwkv = {"model": "rwkv7-g1-0.1b shape (random-init)", "encode_MB_per_s": world * nr / float(tt[0].item()) / 1e6,
                "encode_tokens": nr, "encode_segments": args.rwkv_segments,
                "decode_MB_per_s": world * nd / float(tt[1].item()) / 1e6, "decode_tokens": nd, "decode_segments": args.rwkv_decode_segments,
                "roundtrip_ok": bool(np.array_equal(out_r, ids_r[:nd])), "compressed_bytes": int(sum(len(p) for p in pays_r)),
                "encode_kernel_ms": {k: round(v[0], 2) for k, v in rfam.items()},
                "encode_kernel_launches": {k: v[1] for k, v in rfam.items()},
                "timing": "wall clock around cz_encode / cz_decode (host buffers in and out)"}
        rmodel.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_hbm, peak_src = measured_peaks()
    gemm_ms = fam["gemm"][0] / args.steps
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None

    # per-kernel rooflines: algorithmic flops or bytes of ONE step over the family's live device time
    cfgm = cz.SMOLLM_135M
    D, F, L, V = cfgm["d_model"], cfgm["d_ffn"], cfgm["n_layers"], cfgm["vocab"]
    qkv_n = D + 2 * cfgm["n_kv_heads"] * 64

    def kern(name, fam_key, bound, work):
        ms = fam[fam_key][0] / args.steps
        if ms <= 0:
            return None
        if bound == "tensor":
            a, pk, unit = work / (ms / 1e3) / 1e12, peak_tf, "TFLOP/s"
        else:
            a, pk, unit = work / (ms / 1e3) / 1e9, peak_hbm, "GB/s"
        return {"kernel": name, "bound": bound, "achieved": a, "peak": pk, "unit": unit, "frac": a / pk, "ms_per_step": ms}

    kernels = [k for k in [
        kern("gemm_tc_kernel<256,SWIGLU,pair> gate/up", "gemm_gu", "tensor", 2.0 * rows * 2 * F * D * L),
        kern("gemm_tc_kernel<192,ADD_NORM_TMA,pair> down_proj (+ next norm's operands)", "gemm_down", "tensor", 2.0 * rows * D * F * L),
        # o_proj is bound by the fp32 residual: per row A 1152 B + residual read and written 4608 B + bf16 norm operand 1152 B
        kern("gemm_tc_kernel<192,ADD_NORM_TMA,pair> o_proj (+ next norm's operands)", "gemm_o", "hbm", float(rows) * (2 * D + 8 * D + 2 * D) * L),
        kern("gemm_tc_kernel<192,QKV_ROPE> qkv + RoPE", "gemm_qkv", "tensor", 2.0 * rows * qkv_n * D * L),
        kern("gemm_tc_kernel<256,COLMAX> LM head", "gemm_head", "tensor", 2.0 * HEAD_PARAMS * n),
        kern("attn_tc_kernel", "attn", "tensor", 70.57e6 * n),
        kern("cdf_cols_kernel", "cdf", "hbm", 4.0 * V * n),
    ] if k]
    line = {
        "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "SmolLM-135M (seeded random-init, bf16 weights, fp32 accumulate), ctx 512 / reprime 512: "
                               f"{n} byte-level tokens per GPU per step as {S} segments = {n_chunks} reprime-chunks ({rows} teacher-forced rows)",
                   "tokens_per_step_per_gpu": n, "segments": S, "chunks": n_chunks, "bytes_per_token": 1,
                   "l2": "per-step working set (about 3 GB of activations per wave + a 51.5 GB logits batch) exceeds the 126 MB L2 by orders of magnitude; no explicit flush",
                   "engine": "tcgen05"},
        "tokens_per_s": value * 1e6, "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "MB/s", "h2d_bytes_per_step": int(4 * n + 16 * rows + 4 * n),
                "d2h_bytes_per_step": int(payload_bytes + 8 * S + 16), "bitstream_equal_to_device_arm": same},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if achieved else None,
                     # dram__bytes_read + dram__bytes_write per launch from the ncu --set full captures of this build
                     # (profiles/ncu_summary_r01g.md), summed over the family's launches of one step
                     "traffic": NCU_GEMM_TRAFFIC_BYTES_PER_STEP, "traffic_source": "profiles/ncu_summary_r01g.md (ncu --set full, per launch x launches per step)",
                     "kernel": "gemm_tc_kernel (tcgen05 GEMM family: qkv+rope/o/gate-up/down/lm_head; the RMSNorm passes live in the o/down epilogues)",
                     "flops_per_step": gemm_flops, "kernel_ms_per_step": gemm_ms, "peak_source": peak_src},
        # whole-path tensor roofline exactly as SURVEY 8d defines it: tokens/s x 551.0 MFLOP / measured sustained bf16 peak
        "roofline_path": {"bound": "tensor", "achieved": value * 1e6 / max(1, world) * MFLOP_PER_TOKEN * 1e6 / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                          "frac": value * 1e6 / max(1, world) * MFLOP_PER_TOKEN * 1e6 / 1e12 / peak_tf, "per_gpu": True},
        "roofline_kernels": kernels,
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in fam.items()},
        "kernel_launches_per_step": {k: v[1] // max(1, args.steps) for k, v in fam.items()},
        "compressed_bytes_per_step": payload_bytes, "decode": decode, "rwkv7": rwkv,
    }
    if world == 1 and not args.no_cpu_baseline:
        one_chunk, n_tok, cores = oracle_chunk_sample(args.ref_chunk)
        dt = one_chunk()
        line["cpu_baseline"] = {"value": n_tok / dt / 1e6, "unit": "MB/s", "cores": cores, "kind": "port",
                                "sample": f"one reprime chunk: {args.ref_chunk - 1}-token prefill + {n_tok} coded tokens ({dt:.1f} s), oracle C restatement, OpenMP"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tokens", type=int, default=262144)
    ap.add_argument("--segments", type=int, default=32)
    ap.add_argument("--decode-tokens", type=int, default=262144)
    ap.add_argument("--decode-segments", type=int, default=2048,
                    help="lock-step streams of the stepwise decoder (its throughput scales with the stream count)")
    ap.add_argument("--ref-chunk", type=int, default=256, help="coded tokens per CPU-reference sample chunk (512 = the full reprime chunk)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rwkv-tokens", type=int, default=131072, help="RWKV-7 extra figures (0 = skip)")
    ap.add_argument("--rwkv-segments", type=int, default=256)
    ap.add_argument("--rwkv-decode-tokens", type=int, default=65536)
    ap.add_argument("--rwkv-decode-segments", type=int, default=1024)
    args = ap.parse_args()
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use every host core it can
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
