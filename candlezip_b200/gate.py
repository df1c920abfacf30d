"""Agentic gate ("scan") over the batched hot path: SURVEY 8 a-6 / a-7, BASELINE config 4.

Mirrors the boundary block of the reference's encode loop (src/main.rs:1981-2270) without the agent subprocess: the agent
texts are an input (replayed from `agent_cache.jsonl`, src/main.rs:1152-1195, or any other source).

  reference (per boundary, sequential, on a second session)      here
  -------------------------------------------------------------  ------------------------------------------------------------
  candidates = [normalized, first 2000 bytes, digit lines,       build_candidates()
                Capitalised words]            (main.rs:2018-2032)
  budgets = [max_ctx/8, /4, /2, 3/4]          (main.rs:2034-2035)  BUDGETS (max_ctx = 511 for SmolLM)
  baseline XE + up to 16 conditioned XE passes (2043-2054)       ONE cz_xe_bits call for ALL boundaries: every (boundary,
                                                                  candidate, budget) is an independent paired stream
  best = argmax saved; gate = saved > 0 (+ thresholds) (2068-72)  same selection order (first maximum wins)
  gate: prime = tail(ids[..=i], 511 - budget) ++ hint[..budget],  cz_prime_event{i, hint, hold_until = i + lookahead,
        hold_until = i + lookahead               (2123-2149)       hist_take}: the library builds the tail itself, so the
                                                                  decoder needs only the hint tokens and the gate records
  GateRecordV2 {gate, cand, budget} -> AGT2   (main.rs:2122)      gate_records (one byte each: gate | cand<<1 | budget<<3)
"""
import numpy as np

MAX_CTX_SMOLLM = 511  # max_context_length() - 1, src/models.rs:91, main.rs:2034


def build_candidates(agent_text: str):
    """src/main.rs:2018-2032 (the Rust slices the first 2000 BYTES of the normalized string)"""
    normalized = "".join(" " if (ord(c) < 32 or 127 <= ord(c) < 160) else c for c in agent_text)
    raw = normalized.encode("utf-8")
    head = raw[:2000].decode("utf-8", errors="ignore") if len(raw) > 2000 else normalized
    nums = "\n".join(l for l in normalized.split("\n") if any(c.isascii() and c.isdigit() for c in l))
    caps = " ".join(w for w in normalized.split() if w[:1].isascii() and w[:1].isupper())
    return [normalized, head, nums, caps]


def budgets(max_ctx=MAX_CTX_SMOLLM):
    return [max_ctx // 8, max_ctx // 4, max_ctx // 2, (max_ctx * 3) // 4]


def plan_scan(ids, agent_texts, tokenize_hint, agent_chunk, scan_lookahead=512, scan_max_hint_tokens=512, bos=0,
              max_ctx=MAX_CTX_SMOLLM):
    """Builds the XE jobs of every boundary.  ids: coded tokens (no BOS).  agent_texts: {chunk_index: text} (1-based chunk
    index = boundary / agent_chunk, main.rs:1985).  Returns (jobs, plan) with plan[k] = dict(i, chunk_index, job0, hints)
    where jobs[job0] is the baseline and jobs[job0 + 1 + cid*4 + bid] the conditioned stream (or None when it equals the
    baseline: empty hint, main.rs:2050)."""
    seq = np.concatenate([[bos], np.asarray(ids, np.uint32)]).astype(np.uint32)  # the reference's `ids` (BOS first)
    n = len(seq) - 1
    bud = budgets(max_ctx)
    jobs, plan = [], []
    boundary = agent_chunk
    for i in range(n):
        if i + 1 != boundary:
            continue
        boundary += agent_chunk
        chunk_index = (i + 1) // agent_chunk
        end = min(i + scan_lookahead, len(seq))
        targets = seq[i:end]  # main.rs:2036-2038: starts AT ids[chunk_end], as the reference does
        if len(targets) == 0:
            continue
        hist = seq[max(0, i - min(max_ctx, i)):i]  # main.rs:2043-2044
        cands = build_candidates(agent_texts.get(chunk_index, ""))
        hints = [np.asarray(tokenize_hint(c, scan_max_hint_tokens), np.uint32) for c in cands]
        entry = dict(i=i, chunk_index=chunk_index, job0=len(jobs), hints=hints, slots=[])
        jobs.append((xe_prime(hist, None, max_ctx), targets))
        for cid in range(4):
            for bid in range(4):
                h = hints[cid][: bud[bid]]
                if len(h) == 0:
                    entry["slots"].append(None)
                else:
                    entry["slots"].append(len(jobs))
                    jobs.append((xe_prime(hist, h, max_ctx), targets))
        plan.append(entry)
    return jobs, plan


def xe_prime(history, hint, max_ctx=MAX_CTX_SMOLLM):
    """prime = tail(history, max_ctx - |hint|) ++ hint[..max_ctx]   (src/main.rs:1727-1739)"""
    hint = np.zeros(0, np.uint32) if hint is None else np.asarray(hint, np.uint32)
    hb = min(len(hint), max_ctx)
    take = min(max_ctx - hb, len(history))
    return np.concatenate([history[len(history) - take:], hint[:hb]]).astype(np.uint32)


def decide(plan, bits, scan_lookahead=512, thr_abs_bits=0.0, thr_pct=0.0, max_ctx=MAX_CTX_SMOLLM):
    """Selection + gate (main.rs:2046-2072) and the resulting prime events (2123-2149).  Returns (records, events, rows):
    records: one byte per boundary (gate | cand<<1 | budget<<3, main.rs:658-670); events: cz_prime_event tuples
    (i, hint tokens, hold_until, hist_take); rows: proof.csv-like dicts."""
    bud = budgets(max_ctx)
    records, events, rows = [], [], []
    for e in plan:
        base = float(bits[e["job0"]])
        best_saved, best_cond, best_cid, best_bid = -np.inf, 0.0, 0, 2
        for cid in range(4):
            for bid in range(4):
                slot = e["slots"][cid * 4 + bid]
                cond = base if slot is None else float(bits[slot])
                saved = base - cond
                if saved > best_saved:
                    best_saved, best_cond, best_cid, best_bid = saved, cond, cid, bid
        pct = best_saved / base if base > 0 else 0.0
        abs_ok = best_saved >= thr_abs_bits if thr_abs_bits > 0 else True
        pct_ok = pct * 100.0 >= thr_pct if thr_pct > 0 else True
        gate = 1 if (best_saved > 0 and abs_ok and pct_ok) else 0
        records.append(gate | (best_cid << 1) | (best_bid << 3))
        rows.append(dict(chunk_index=e["chunk_index"], i=e["i"], baseline_bits=base, conditioned_bits=best_cond, bits_saved=best_saved,
                         percent_saved=pct, gate=gate, candidate_id=best_cid, budget_id=best_bid))
        if gate:
            i, b = e["i"], bud[best_bid]
            hint = e["hints"][best_cid][:b]
            hist_take = min(max(max_ctx - b, 0), i + 1)  # history = ids[..=i] (BOS first): main.rs:2132-2134
            events.append((i, hint, i + scan_lookahead, hist_take))
    return records, events, rows


def events_from_records(records, hints_by_boundary, agent_chunk, n_tokens, scan_lookahead=512, max_ctx=MAX_CTX_SMOLLM):
    """Decode side (main.rs:2543-2614): the gate records from the container + the same agent texts give the same events.
    hints_by_boundary[k]: the four candidate hint token arrays of the k-th boundary (in boundary order)."""
    bud = budgets(max_ctx)
    events, k = [], 0
    boundary = agent_chunk
    for i in range(n_tokens):
        if i + 1 != boundary:
            continue
        boundary += agent_chunk
        if k >= len(records):
            break
        r = records[k]
        gate, cid, bid = r & 1, (r >> 1) & 3, (r >> 3) & 3
        if gate:
            b = bud[bid]
            events.append((i, np.asarray(hints_by_boundary[k][cid][:b], np.uint32), i + scan_lookahead, min(max(max_ctx - b, 0), i + 1)))
        k += 1
    return events


def scan_encode(model, ids, agent_texts, tokenize_hint, agent_chunk, n_segments=1, **kw):
    """compress with the agentic gate: one batched XE scan, then one batched encode with the gated hint primes.
    Returns (payloads, seg_start, records, rows, events)."""
    if n_segments != 1:
        raise ValueError("the gate works on a single AC stream (the reference's container has one)")
    look = kw.get("scan_lookahead", 512)
    jobs, plan = plan_scan(ids, agent_texts, tokenize_hint, agent_chunk, look, kw.get("scan_max_hint_tokens", 512), kw.get("bos", 0))
    bits = model.xe_bits(jobs) if jobs else np.zeros(0)
    records, events, rows = decide(plan, bits, look, kw.get("thr_abs_bits", 0.0), kw.get("thr_pct", 0.0))
    pays, seg = model.encode(ids, n_segments=1, bos=kw.get("bos", 0), events=events or None)
    return pays, seg, records, rows, events
