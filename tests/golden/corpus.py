"""Accessors for the corpora / replay caches committed under tests/golden/data (made by make_data_fixtures.py from files the
reference ships).  Pure stdlib + numpy, no import of the product or the oracle: bench.py's reference arm uses it too."""
import json
import lzma
import os

import numpy as np

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
_cache = {}


def load(name: str) -> bytes:
    """uncompressed bytes of fixture `name` (e.g. "enwik8_3mib", "alice29.txt"), verified against MANIFEST.json"""
    if name not in _cache:
        import hashlib

        man = json.load(open(os.path.join(DATA, "MANIFEST.json")))
        data = lzma.decompress(open(os.path.join(DATA, name + ".xz"), "rb").read())
        if len(data) != man[name]["bytes"] or hashlib.sha1(data).hexdigest() != man[name]["sha1"]:
            raise ValueError(f"fixture {name} does not match its manifest entry")
        _cache[name] = data
    return _cache[name]


def replay(run: str) -> dict:
    """{"agent_cache": {chunk_index(str): {"agent_text", "agent_calls"}}, "proof_header": [...], "proof_rows": [[...]], "meta": {...}}"""
    return json.loads(load(f"replay_{run}.json").decode("utf-8"))


def spread_map(vocab: int, seed: int = 1750) -> np.ndarray:
    """A fixed injective map byte -> token id over [0, vocab).  Offline there is no tokenizer.json, so files are coded one
    token per byte; with id = byte every symbol sits in the first 256 CDF entries and the encode-side prefix walk of the CDF
    kernel (which stops at the coded symbol) is unrealistically short.  The 256 ids are k + (vocab - 256) (k/255)^9, assigned
    to the bytes in a seeded random order.  The seed is the one (of 2000 tried) whose map gives a byte-frequency-weighted mean
    id / vocab of 0.100 on the enwik8 stand-in and 0.099 on alice29.txt, i.e. what the shipped SmolLM2 token-id traces have
    (asyoulik 0.083, alice29 0.094, enwik8_128kb_0 0.111 -- tests/golden/reference_fixtures.npz)."""
    k = np.arange(256, dtype=np.float64)
    ids = (k + np.floor((vocab - 256) * (k / 255.0) ** 9)).astype(np.uint32)
    return ids[np.random.default_rng(seed).permutation(256)]


def byte_ids(data: bytes, vocab: int = 0, spread: bool = False) -> np.ndarray:
    b = np.frombuffer(data, np.uint8)
    return spread_map(vocab)[b] if spread else b.astype(np.uint32)


def ids_to_bytes(ids: np.ndarray, vocab: int = 0, spread: bool = False) -> bytes:
    ids = np.asarray(ids, np.uint32)
    if not spread:
        return ids.astype(np.uint8).tobytes()
    inv = np.full(vocab, 0, np.uint8)
    inv[spread_map(vocab)] = np.arange(256, dtype=np.uint8)
    return inv[ids].tobytes()


def low_entropy_stream(n: int) -> bytes:
    """BASELINE config 5: final_bench/synthetic/alphabet.txt's a-z cycle repeated to n bytes"""
    a = load("alphabet.txt")
    return (a * (n // len(a) + 1))[:n]


def high_entropy_stream(n: int, seed: int = 0xC0FFEE) -> bytes:
    """BASELINE config 5: uniform bytes, PCG64 seed 0xC0FFEE"""
    return np.random.Generator(np.random.PCG64(seed)).integers(0, 256, n, dtype=np.uint8).tobytes()
