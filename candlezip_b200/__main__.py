"""`python -m candlezip_b200 {compress,decompress,self-test}`: a thin twin of the reference CLI's hot-path surface
(src/main.rs:34-221: compress | decompress | self-test, --backend smollm|rwkv7, --context, --reprime-interval) over the
batched B200 path.  Out of scope here, as in DESIGN.md: hf-hub download, agents, watchdog logging.

Weights: --weights *.safetensors (HF LLaMA / candle_rwkv7 tensor names) or --random-init SEED (no checkpoints exist offline).
Tokenizer: --tokenizer tokenizer.json (HF `tokenizers`, SmolLM) or a RWKV vocab json; default = byte-level ids (lossless for
every input).  --segments N > 1 writes the SEG1 container extension (parallel decode); N = 1 is the reference's v2 layout.
"""
import argparse
import sys
import time

import numpy as np

from . import RWKV7_0P1B, SMOLLM_135M, Context, Model, codec


class HfTokenizer:
    """tok.encode(from_utf8_lossy(data), add_special_tokens=false) / tok.decode(ids, skip_special_tokens=true): the reference's
    SmolLM tokenisation (src/main.rs:1850-1857, 2654).  Lossy for non-UTF-8 input exactly like the reference; decompress()
    verifies the BLAKE3 of the decoded bytes and reports a mismatch instead of silently returning different bytes."""

    def __init__(self, path):
        from tokenizers import Tokenizer

        self.tok = Tokenizer.from_file(path)

    def encode_bytes(self, data: bytes):
        return np.asarray(self.tok.encode(data.decode("utf-8", errors="replace"), add_special_tokens=False).ids, np.uint32)

    def decode_bytes(self, ids):
        return self.tok.decode([int(t) for t in ids], skip_special_tokens=True).encode("utf-8")


def _model(args):
    ctx = Context(args.device)
    cfg = SMOLLM_135M if args.backend == "smollm" else RWKV7_0P1B
    m = Model(ctx, cfg)
    if args.weights:
        m.load_safetensors(args.weights)
    else:
        m.random_init(args.random_init, 0.02, 0.02)
    tok = None
    if args.tokenizer:
        tok = HfTokenizer(args.tokenizer) if args.backend == "smollm" else codec.RwkvTokenizer.from_json(args.tokenizer)
    return m, tok


def main(argv=None):
    ap = argparse.ArgumentParser(prog="candlezip_b200")
    ap.add_argument("command", choices=["compress", "decompress", "self-test"])
    ap.add_argument("input")
    ap.add_argument("output", nargs="?")
    ap.add_argument("--backend", choices=["smollm", "rwkv7"], default="smollm")
    ap.add_argument("--context", type=int, default=512)
    ap.add_argument("--reprime-interval", type=int, default=512)
    ap.add_argument("--segments", type=int, default=1)
    ap.add_argument("--weights", nargs="*")
    ap.add_argument("--random-init", type=int, default=0)
    ap.add_argument("--tokenizer")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    model, tok = _model(args)
    data = open(args.input, "rb").read()
    if args.command == "compress":
        blob = codec.compress(model, data, tok, args.segments, context=args.context, reprime_interval=args.reprime_interval)
        open(args.output or args.input + ".canz", "wb").write(blob)
        print(f"{len(data)} -> {len(blob)} bytes ({8 * len(blob) / max(1, len(data)):.4f} bits/byte)")
    elif args.command == "decompress":
        out = codec.decompress(model, data, tok)
        open(args.output or args.input + ".out", "wb").write(out)
        print(f"{len(data)} -> {len(out)} bytes")
    else:  # self-test: encode + decode round trip with timings (src/main.rs:156-221)
        t0 = time.perf_counter()
        blob = codec.compress(model, data, tok, args.segments, context=args.context, reprime_interval=args.reprime_interval)
        t1 = time.perf_counter()
        out = codec.decompress(model, blob, tok)
        t2 = time.perf_counter()
        ok = out == data
        print(f"Compression: {t1 - t0:.3f} s  Decompression: {t2 - t1:.3f} s  {8 * len(blob) / max(1, len(data)):.4f} bits/byte  "
              f"roundtrip {'OK' if ok else 'MISMATCH'}")
        return 0 if ok else 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
