#!/usr/bin/env python3
"""Generate tests/golden/data/* from the corpora and replay caches the reference SHIPS (data, not source).

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_data_fixtures.py
Writes (xz-compressed; `candlezip_b200.corpus` / tests read them back with the stdlib `lzma`):
  enwik8_3mib.xz      final_bench/enwik8_samples/enwik8_128kb_{0..23} concatenated = the first 3,145,728 bytes of enwik8
                      (BASELINE config 2/3 stand-in: final_bench/enwik8.zst is not mounted offline)
  alice29.txt.xz      final_bench/cantrbry/alice29.txt (148,481 B LF copy; BASELINE config 1)
  asyoulik.txt.xz     final_bench/cantrbry/asyoulik.txt (125,179 B)
  alphabet.txt.xz     final_bench/synthetic/alphabet.txt (the a-z cycle of BASELINE config 5's low-entropy stream)
  replay_<run>.json.xz  results_300s_nomem/<run>/{agent_cache.jsonl, proof.csv}: the agent texts keyed by chunk index and the
                      ledger rows (all 26 columns, as strings) -- the replay input of BASELINE config 4 (src/main.rs:1152-1195)
  MANIFEST.json       sizes + SHA-1 + BLAKE3-128 of every uncompressed payload
"""
import csv
import hashlib
import io
import json
import lzma
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def put(manifest, name, data: bytes):
    import blake3

    with open(os.path.join(OUT, name + ".xz"), "wb") as f:
        f.write(lzma.compress(data, preset=9 | lzma.PRESET_EXTREME))
    manifest[name] = {"bytes": len(data), "sha1": hashlib.sha1(data).hexdigest(), "blake3_16": blake3.blake3(data).digest()[:16].hex()}


def main():
    os.makedirs(OUT, exist_ok=True)
    man = {}
    enwik = b"".join(open(os.path.join(REF, "final_bench/enwik8_samples", f"enwik8_128kb_{i}"), "rb").read() for i in range(24))
    assert len(enwik) == 3145728 and hashlib.sha1(enwik).hexdigest().startswith("51f7761c")
    put(man, "enwik8_3mib", enwik)
    put(man, "alice29.txt", open(os.path.join(REF, "final_bench/cantrbry/alice29.txt"), "rb").read())
    put(man, "asyoulik.txt", open(os.path.join(REF, "final_bench/cantrbry/asyoulik.txt"), "rb").read())
    put(man, "alphabet.txt", open(os.path.join(REF, "final_bench/synthetic/alphabet.txt"), "rb").read())
    runs = {
        "alice29": "results_300s_nomem/alice29_selftest_smollm_20251003_043405",
        "asyoulik": "results_300s_nomem/asyoulik_selftest_smollm_20251003_121454",
    }
    for name, d in runs.items():
        texts = {}
        for line in open(os.path.join(REF, d, "agent_cache.jsonl"), encoding="utf-8"):
            if line.strip():
                v = json.loads(line)
                texts[str(v["chunk_index"])] = {"agent_text": v["agent_text"], "agent_calls": v["agent_calls"]}
        raw = open(os.path.join(REF, d, "proof.csv"), encoding="utf-8", newline="").read()
        rows = list(csv.reader(io.StringIO(raw)))
        meta = json.load(open(os.path.join(REF, d, "meta.json"))) if os.path.exists(os.path.join(REF, d, "meta.json")) else {}
        blob = json.dumps({"run": d, "agent_cache": texts, "proof_header": rows[0], "proof_rows": rows[1:], "meta": meta}, ensure_ascii=False)
        put(man, f"replay_{name}.json", blob.encode("utf-8"))
    json.dump(man, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    for k, v in man.items():
        print(k, v["bytes"], os.path.getsize(os.path.join(OUT, k + ".xz")))


if __name__ == "__main__":
    main()
