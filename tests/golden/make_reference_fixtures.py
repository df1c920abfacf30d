#!/usr/bin/env python3
"""Generate tests/golden/reference_fixtures.npz from the reference's SHIPPED artefacts.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_reference_fixtures.py
What it extracts:
  * container headers: the bytes of the four shipped .canz files up to the start of the AC payload
    (final_bench/cantrbry/{asyoulik,fields,alice29}.canz, final_bench/enwik8_samples/enwik8_128kb_0.canz),
    plus file sizes and BLAKE3-128 of the matching source file (orig_hash16 check).
  * per shipped self-test run (results_300s_nomem/*): the decoded token ids (`sym` per decode_step), the
    positions of every `context_reprime` event, and the per-chunk gate/candidate/budget columns of proof.csv.
    These pin the reprime schedule of src/main.rs:2275-2290 / 2530-2541 incl. the hint hold-off (2149, 2614).
Nothing here is reference SOURCE code; only data the reference ships as results.
"""
import csv
import glob
import json
import os
import struct

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixtures.npz")


def read_var(b, n):
    shift = 0
    v = 0
    while True:
        byte = b[n]
        n += 1
        v |= (byte & 0x7F) << shift
        if not byte & 0x80:
            return v, n
        shift += 7


def header_len(b):
    """offset of the AC payload: header v2 (+ AGT2/AGTB section when flag bit 2 is set)."""
    n = 4 + 2 + 4
    _, n = read_var(b, n)
    _, n = read_var(b, n)
    n += 48
    flags, ctx, vocab, repr_len, reprime = struct.unpack_from("<5I", b, n)
    n += 20 + repr_len
    if flags & 4:
        magic = b[n : n + 4]
        n += 4
        cnt, n = read_var(b, n)
        n += cnt if magic == b"AGT2" else (cnt + 7) // 8
    return n


def main():
    out = {}
    try:
        import blake3
    except ImportError:
        blake3 = None
    canz = {
        "asyoulik": ("final_bench/cantrbry/asyoulik.canz", "final_bench/cantrbry/asyoulik.txt"),
        "fields": ("final_bench/cantrbry/fields.canz", "final_bench/cantrbry/fields.c"),
        "alice29": ("final_bench/cantrbry/alice29.canz", "final_bench/cantrbry/alice29.txt"),
        "enwik8_128kb_0": ("final_bench/enwik8_samples/enwik8_128kb_0.canz", "final_bench/enwik8_samples/enwik8_128kb_0"),
    }
    for name, (cz, src) in canz.items():
        b = open(os.path.join(REF, cz), "rb").read()
        hl = header_len(b)
        out[f"canz_{name}_header"] = np.frombuffer(b[:hl], dtype=np.uint8)
        out[f"canz_{name}_file_size"] = np.int64(len(b))
        data = open(os.path.join(REF, src), "rb").read()
        out[f"canz_{name}_src_size"] = np.int64(len(data))
        if blake3 is not None:
            out[f"canz_{name}_src_blake3_16"] = np.frombuffer(blake3.blake3(data).digest()[:16], dtype=np.uint8)
    runs = {
        "asyoulik": "results_300s_nomem/asyoulik_selftest_smollm_20251003_121454",
        "alice29": "results_300s_nomem/alice29_selftest_smollm_20251003_043405",
        "enwik8_128kb_0": "results_300s_nomem/results_360s_nomem/enwik8_128kb_0_selftest_smollm_20251003_215646",
    }
    for name, d in runs.items():
        syms, reprimes = [], []
        with open(os.path.join(REF, d, "watchdog_decode_steps.jsonl")) as f:
            for line in f:
                r = json.loads(line)
                if r.get("phase") == "decode_step":
                    assert r["i"] == len(syms)
                    syms.append(r["sym"])
                elif r.get("phase") == "context_reprime":
                    assert r["window"] == 511
                    reprimes.append(r["i"])
        gates = []
        with open(os.path.join(REF, d, "proof.csv")) as f:
            for row in csv.DictReader(f):
                gates.append((int(row["chunk_index"]), int(row["gate"]), int(row["candidate_id"]), int(row["budget_id"])))
        out[f"run_{name}_syms"] = np.asarray(syms, dtype=np.uint16)
        out[f"run_{name}_reprimes"] = np.asarray(reprimes, dtype=np.int64)
        out[f"run_{name}_gates"] = np.asarray(gates, dtype=np.int32)
        mm = os.path.join(REF, d, "watchdog_mismatch.json")
        if os.path.exists(mm):
            out[f"run_{name}_first_mismatch_i"] = np.int64(json.load(open(mm))["i"])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
