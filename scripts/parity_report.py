#!/usr/bin/env python3
"""Measures the SmolLM logits tolerance on the GPU box and writes the table kept in profiles/parity_rNN.md.

    python scripts/parity_report.py > gpurun_out/parity_r02.md

For every sequence position of two chunk shapes (a BOS-started run, and the steady-state reprime chunk: 511-token prime + 512
coded tokens) it reports max|logit_gpu - logit_oracle| / std(oracle logits of that position) against the oracle run with the same
bf16 rounding points (round_bf16=1: what remains is accumulation order and the attention kernel's exp2 / bf16 P) and against the
pure-f32 oracle (the reference's CPU semantics).  TEST INFRASTRUCTURE: uses the oracle as the checker."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import candlezip_b200 as cz  # noqa: E402
import parity_util as pu  # noqa: E402
from candlezip_b200 import _lib  # noqa: E402

KEY = (0, 1, 63, 64, 127, 128, 129, 255, 256, 383, 384, 511, 512, 639, 640, 767, 768, 895, 896, 1021)


def run(name, model, shapes, seed):
    o16, o32 = pu.oracle_llama(model, 1), pu.oracle_llama(model, 0)
    V = model.cfg["vocab"]
    rng = np.random.default_rng(seed)
    print(f"\n## {name}\n")
    for n_prime, n_targets in shapes:
        prime = rng.integers(0, V, n_prime).astype(np.uint32)
        targets = rng.integers(0, V, n_targets).astype(np.uint32)
        got = model.chunk_logits(prime, targets)
        w16 = pu.oracle_chunk_logits(o16, prime, targets)
        w32 = pu.oracle_chunk_logits(o32, prime, targets)
        r16, r32 = pu.logits_parity(got, w16, w32), pu.logits_parity(got, w32)
        rr = pu.logits_parity(w16, w32)  # the rounding points alone (oracle vs oracle)
        p0 = n_prime - 1
        print(f"shape: {n_prime}-token prime + {n_targets} coded tokens = sequence positions {p0}..{p0 + n_targets - 1}; "
              f"std(logits) {w32.std(axis=1).mean():.4f}\n")
        print("| | vs oracle, bf16 rounding points | vs oracle, pure f32 | (oracle bf16 vs oracle f32) |")
        print("|---|---|---|---|")
        print(f"| **max over all {n_targets} positions** | **{r16.max():.5f}** (pos {p0 + int(r16.argmax())}) | **{r32.max():.5f}** (pos {p0 + int(r32.argmax())}) | {rr.max():.5f} |")
        print(f"| mean | {r16.mean():.5f} | {r32.mean():.5f} | {rr.mean():.5f} |")
        for p in KEY:
            j = p - p0
            if 0 <= j < n_targets:
                print(f"| position {p} | {r16[j]:.5f} | {r32[j]:.5f} | {rr[j]:.5f} |")
        for lo in range(0, 1024, 128):
            js = [j for j in range(n_targets) if lo <= p0 + j < lo + 128]
            if js:
                print(f"| key block {lo // 128} (positions {lo}..{lo + 127}), max | {r16[js].max():.5f} | {r32[js].max():.5f} | {rr[js].max():.5f} |")
        print()


def main():
    ctx = cz.Context(0)
    print("# SmolLM logits parity, measured on B200 (scripts/parity_report.py)\n")
    print("metric: max over the vocabulary of |logit_gpu - logit_oracle| divided by the standard deviation of the oracle's logits at that "
          "position.  Random-init weights (seeded), uniformly random tokens.")
    for eng, nm in ((_lib.CZ_ENGINE_TCGEN05, "tcgen05"), (_lib.CZ_ENGINE_SIMT, "simt")):
        m = cz.Model(ctx, cz.SMOLLM_TINY, engine=eng).random_init(5, 0.05, 0.2)
        run(f"SMOLLM_TINY (2 layers, d 192, V 1024), engine {nm}", m, [(1, 1022), (511, 512)], 11)
        m.close()
    m = cz.Model(ctx, cz.SMOLLM_135M).random_init(0, 0.02, 0.05)
    run("SmolLM-135M shape (30 layers, d 576, V 49152), engine tcgen05", m, [(1, 600), (511, 512)], 3)


if __name__ == "__main__":
    main()
