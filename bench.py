#!/usr/bin/env python
"""Headline benchmark: PageRank GTEPS/iter (BASELINE.json configs[1]) and batched
query scoring queries/s (configs[2]) on N B200s, plus the CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one
batch of synthetic input:
  pagerank  one ss_pagerank run to convergence (eps 1e-9) over the resident graph
  scoring   one ss_score_batch over the resident index
`value` is timed on the engine's CUDA stream with inputs resident in HBM; `e2e`
goes through the same C ABI with pinned HOST buffers and includes the copies
(for PageRank: graph export H2D + device-side transpose + ranks D2H, i.e. what
ranking.UpdateTopicSensitivePagerank costs minus Badger).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

T_TOPICS = 16
DAMPING = 0.75   # cmd/crawl/start_crawl.go:175
EPS = 1e-9       # BASELINE.json configs[1]
TOP_K = 10       # BASELINE.json configs[2]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(key)
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi SM clock / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        # "under load" = samples at or above the median (idle samples around the region pull it down)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def dist_env(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    return rank, local, world


def pin(a):
    """numpy -> pinned torch tensor sharing no memory with the source."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
    return t


def pin_view(t, dtype):
    return t.numpy().view(dtype)


# --------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    from spaghettisearch_b200 import capi, sharding, synth

    rank, local, world = dist_env(args)
    # stdout carries exactly one JSON line: NCCL's own banner/debug lines go to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cores = os.cpu_count() or 1
    threads = max(1, cores // world)
    peak, peak_src = measured_peak_gbs()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    eng = capi.Engine(device=local, timing=True)
    if world > 1:
        eng.comm_init(sharding.share_unique_id(capi.comm_unique_id), rank, world)
    ext = torch.cuda.ExternalStream(eng.stream_handle(), device=torch.device("cuda", local))
    out = {}

    # ------------------------------------------------------------------ PageRank
    if args.workload in ("both", "pagerank"):
        n_nodes, n_edges_target = args.nodes * world, args.edges * world  # weak scaling
        t0 = time.time()
        if world == 1:
            g = synth.graph(n_nodes, n_edges_target, seed=42, n_threads=threads)
            row_ptr, col_idx = g.row_ptr, g.col_idx
        else:
            # each rank generates a slice of the rows, slices are exchanged over NCCL
            lo, hi = sharding.row_slice(rank, world, n_nodes)
            part = synth.graph_rows(n_nodes, n_edges_target, lo, hi, seed=42, n_threads=threads)
            row_ptr, col_idx = sharding.assemble_graph(n_nodes, part.row_ptr, part.col_idx, device="cuda")
            del part
            torch.cuda.empty_cache()
        E = int(row_ptr[-1])
        gen_s = time.time() - t0
        npg = synth.topics(T_TOPICS)
        h_row_ptr, h_col = pin(row_ptr), pin(col_idx)
        p_row_ptr, p_col = pin_view(h_row_ptr, np.uint64), pin_view(h_col, np.uint32)
        log(f"[rank {rank}] graph N={n_nodes} E={E} generated in {gen_s:.1f}s")

        eng.graph_load_csr(p_row_ptr, p_col)
        st0 = eng.pagerank_stats()
        R, E_loc = int(st0.local_rows), int(st0.local_edges)
        for _ in range(args.warmup):
            eng.pagerank(DAMPING, EPS, npg, want_rank=False)
        sweeps_total, launches, sweep_ms, gather_ms, exch_ms = 0, 0, 0.0, 0.0, 0.0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with ClockSampler(local) as clk:
            ev0.record(ext)
            for _ in range(args.steps):
                _, iters, status = eng.pagerank(DAMPING, EPS, npg, want_rank=False)
                s = eng.pagerank_stats()
                sweeps_total += s.sweeps
                launches += s.launches
                sweep_ms += s.sweep_ms_total
                gather_ms += s.gather_ms_total
                exch_ms += s.exchange_ms_total
            ev1.record(ext)
            barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        clocks = clk.summary()
        gteps = E * T_TOPICS * sweeps_total / (ms * 1e-3) / 1e9
        # roofline of the sweep (dominant kernels: k_sweep_short + k_sweep_long), this rank's rows
        b_pr = 4 * E_loc + 8 * (R + 1) + 8 * R + 16 * T_TOPICS * R
        avg_sweep_s = sweep_ms * 1e-3 / max(1, sweeps_total)
        achieved = b_pr / avg_sweep_s / 1e9
        # gather-traffic model beside it: every edge moves one 8T-byte row out of L2/HBM
        gather_bytes = E_loc * 8 * T_TOPICS

        # e2e through the C ABI with host buffers
        h_rank = torch.empty((R if world > 1 else n_nodes) * T_TOPICS, dtype=torch.float64).pin_memory()
        e2e_steps = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        e2e_sweeps = 0
        for _ in range(e2e_steps):
            eng.graph_load_csr(p_row_ptr, p_col)
            if world == 1:
                eng.pagerank(DAMPING, EPS, npg, out=h_rank.numpy().reshape(n_nodes, T_TOPICS))
            else:  # every rank copies out its own row block
                eng.pagerank(DAMPING, EPS, npg, want_rank=False)
                eng.pagerank_fetch(int(st0.row_lo), int(st0.row_lo) + R, out=h_rank.numpy().reshape(R, T_TOPICS))
            e2e_sweeps += eng.pagerank_stats().sweeps
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e_gteps = E * T_TOPICS * e2e_sweeps / e2e_s / 1e9
        h2d = (n_nodes + 1) * 8 + E * 4
        d2h = (n_nodes if world == 1 else R) * T_TOPICS * 8 + 3 * 2 * T_TOPICS * 8 * (e2e_sweeps // e2e_steps)

        out.update({
            "metric": "pagerank_gteps_per_iter", "value": gteps, "unit": "GTEPS (topic-edges/s, E*T per sweep)",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[1]: topic-sensitive PageRank, 16 ODP topics, "
                                   f"{args.nodes} nodes / {args.edges} edges power-law graph per GPU, fp64, "
                                   "eps 1e-9, d 0.75", "nodes": n_nodes, "edges": E, "topics": T_TOPICS,
                       "sweeps_per_step": sweeps_total / args.steps,
                       "cache": "inputs larger than L2 (state 2x%.2f GB + graph %.2f GB vs 126 MB L2)" %
                                (n_nodes * T_TOPICS * 8 / 1e9, (4 * E_loc + 8 * R) / 1e9),
                       "partition": "rows, edge-balanced; NCCL exchange per sweep" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e_gteps, "unit": "GTEPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3 / e2e_steps,
                    "what": "ss_graph_load_csr (pinned host CSR, device transpose) + ss_pagerank with host output"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("pagerank_sweep_dram_bytes"), "peak_source": peak_src,
                         "kernel": "k_sweep_short32 + k_sweep_long (one sweep)",
                         "algorithmic_bytes_per_sweep": b_pr, "avg_sweep_ms": avg_sweep_s * 1e3,
                         "gather_model_bytes_per_sweep": b_pr + gather_bytes,
                         "gather_model_GBps": (b_pr + gather_bytes) / avg_sweep_s / 1e9,
                         "exchange_ms_per_sweep": exch_ms / max(1, sweeps_total)},
            "edges_per_s": E * sweeps_total / (ms * 1e-3),
        })
        if rank == 0 and world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_pagerank(row_ptr, col_idx, npg, cores)
        del h_rank
    # ------------------------------------------------------------------- scoring
    if args.workload in ("both", "scoring"):
        sc = run_scoring(args, eng, ext, rank, world, local, threads, cores, peak, peak_src, barrier, max_over_ranks,
                         sum_over_ranks)
        if "metric" in out:
            out["scoring"] = sc
        else:
            out.update(sc)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


def run_scoring(args, eng, ext, rank, world, local, threads, cores, peak, peak_src, barrier, max_over_ranks,
                sum_over_ranks):
    import torch
    from spaghettisearch_b200 import capi, synth
    D, V, Q = args.docs * world, args.terms, args.queries
    from spaghettisearch_b200 import sharding
    lo, hi = sharding.doc_shard(rank, world, D)
    t0 = time.time()
    title = synth.index_table(V, D, 0, doc_lo=lo, doc_hi=hi, n_threads=threads)
    body = synth.index_table(V, D, 1, doc_lo=lo, doc_hi=hi, n_threads=threads)
    q = synth.queries(Q, V, phrase_fraction=0.0, seed=44)
    log(f"[rank {rank}] index D={D} V={V} postings title={title.n_postings} body={body.n_postings} "
        f"generated in {time.time() - t0:.1f}s")
    eng.index_clear()
    t0 = time.time()
    eng.index_load(capi.SS_TITLE, D, title.term_ptr, title.doc_ids, title.norm_tf)
    eng.index_load(capi.SS_BODY, D, body.term_ptr, body.doc_ids, body.norm_tf)
    eng.term_weights(capi.SS_TITLE, float(D), title.n_postings, D, df_global=title.df_global, want=False)
    eng.term_weights(capi.SS_BODY, float(D), body.n_postings, D, df_global=body.df_global, want=False)
    load_s = time.time() - t0
    # forw[3] rows for the blend: synthetic ranks around 1/D (the reference's ranks are near uniform)
    rng = np.random.default_rng(7)
    pr = (rng.random((D, T_TOPICS)) + 0.5) / D
    eng.set_pagerank(pr)
    del pr
    probs = np.full(T_TOPICS, 1.0 / T_TOPICS)
    h_kw_ptr, h_kw = pin(q.kw_ptr), pin(q.kw_terms)
    kw_ptr, kw = pin_view(h_kw_ptr, np.uint64), pin_view(h_kw, np.uint32)
    bufs = (torch.empty(Q * TOP_K, dtype=torch.int32).pin_memory(), torch.empty(Q * TOP_K, dtype=torch.float64).pin_memory(),
            torch.empty(Q * TOP_K, dtype=torch.float64).pin_memory(), torch.empty(Q, dtype=torch.int32).pin_memory())
    outs = (bufs[0].numpy().view(np.uint32).reshape(Q, TOP_K), bufs[1].numpy().reshape(Q, TOP_K),
            bufs[2].numpy().reshape(Q, TOP_K), bufs[3].numpy().view(np.uint32))
    for _ in range(args.warmup):
        eng.score_batch(kw_ptr, kw, topic_probs=probs, k=TOP_K, out=outs)
    kernel_ms, score_ms, launches, alg_bytes, postings = 0.0, 0.0, 0, 0, 0
    barrier()
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.score_batch(kw_ptr, kw, topic_probs=probs, k=TOP_K, out=outs)
            s = eng.score_stats()
            kernel_ms += s.kernel_ms
            score_ms += s.score_kernel_ms
            launches += s.launches
            alg_bytes += s.algorithmic_bytes
            postings += s.postings_scanned
        barrier()
        wall_s = time.perf_counter() - t0
    kernel_ms = max_over_ranks(kernel_ms)
    wall_s = max_over_ranks(wall_s)
    achieved = alg_bytes / (score_ms * 1e-3) / 1e9
    # The same batch with the impact-vector path and the cross-slab bound switched off: every posting of
    # every query list is walked (the "batched sparse gather" taken literally).  Reported beside the
    # default so that the effect of the screening structures is visible; results are identical.
    os.environ["SS_SCORE_DENSE"] = "0"
    os.environ["SS_SCORE_QTHR"] = "0"
    eng.score_batch(kw_ptr, kw, topic_probs=probs, k=TOP_K, out=outs)
    s_walk = eng.score_stats()
    del os.environ["SS_SCORE_DENSE"], os.environ["SS_SCORE_QTHR"]
    walk_ms = max_over_ranks(s_walk.kernel_ms)
    walk = {"value": Q / (walk_ms * 1e-3), "unit": "queries/s", "ms_per_step": walk_ms, "steps": 1,
            "roofline_frac": s_walk.algorithmic_bytes / (s_walk.score_kernel_ms * 1e-3) / 1e9 / peak,
            "what": "SS_SCORE_DENSE=0 SS_SCORE_QTHR=0: posting lists walked for every query, same results"}
    sc = {
        "metric": "scoring_queries_per_s", "value": Q * args.steps / (kernel_ms * 1e-3), "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 sums of f32 weights",
        "data": "synthetic",
        "config": {"workload": f"BASELINE.json configs[2]: batched cosine scoring, {Q} keyword queries over "
                               f"{args.docs}-doc/GPU / {V}-term synthetic Zipf index, PageRank blend + top-{TOP_K}",
                   "docs": D, "terms": V, "queries": Q, "k": TOP_K,
                   "postings": int(sum_over_ranks(title.n_postings + body.n_postings)),
                   "cache": "index %.1f GB larger than L2" % ((title.n_postings + body.n_postings) * 8 / 1e9),
                   "shard": "docs" if world > 1 else "single GPU", "index_load_s": load_s},
        "clocks": clk.summary(),
        "e2e": {"value": Q * args.steps / wall_s, "unit": "queries/s",
                "h2d_bytes_per_step": int(kw_ptr.nbytes + kw.nbytes + probs.nbytes),
                "d2h_bytes_per_step": int(sum(o.nbytes for o in outs)), "ms_per_step": wall_s * 1e3 / args.steps,
                "what": "ss_score_batch with pinned host query/result buffers, index resident"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("score_dram_bytes"), "peak_source": peak_src, "kernel": "k_score",
                     "algorithmic_bytes_per_batch": alg_bytes // max(1, args.steps),
                     "postings_per_batch": postings // max(1, args.steps),
                     "avg_kernel_ms": score_ms / max(1, args.steps),
                     "note": "algorithmic bytes = 8 B per posting of every query list (+ per matched doc, per "
                             "result), re-reads across queries counted (SURVEY.md 8(d)); the impact-vector path "
                             "reads 2 B per doc and dense term instead of the lists, so frac is a rate against "
                             "the nominal bytes, not the bytes moved"},
        "walk_every_posting": walk,
    }
    if world > 1:
        sc["config"]["note"] = "per-shard top-k lists; cross-shard merge (ss_merge_topk) not in the timed region"
    if rank == 0 and world == 1 and not args.no_cpu:
        sc["cpu_baseline"] = cpu_scoring(eng, title, body, D, q, probs, cores)
    return sc


# ------------------------------------------------------------------- CPU arms
def cpu_pagerank(row_ptr, col_idx, npg, cores, sweeps=1):
    """Oracle timed on the host cores: 'fair' (CSC pull, OpenMP) on the full graph and
    'faithful' (string-keyed maps, one thread, one topic) on a 1/20 subgraph."""
    from oracle import loader as O
    from spaghettisearch_b200 import synth
    E = int(row_ptr[-1])
    t0 = time.time()
    in_ptr, in_src = O.csc_build(row_ptr, col_idx)
    build_s = time.time() - t0
    _, _, secs = O.pagerank_fair_csc(row_ptr, in_ptr, in_src, DAMPING, EPS, npg, fixed_iters=sweeps, n_threads=cores,
                                     want_rank=False)
    fair = E * len(npg) * sweeps / secs / 1e9
    n_small = max(1000, (len(row_ptr) - 1) // 20)
    gs = synth.graph(n_small, n_small * 15, seed=42)
    _, _, fsecs = O.pagerank_faithful(gs.row_ptr, gs.col_idx, DAMPING, EPS, int(npg[0]), fixed_iters=1, want_rank=False)
    faithful = gs.n_edges / fsecs / 1e9
    return {"value": fair, "unit": "GTEPS", "cores": cores, "kind": "port",
            "sample": f"{sweeps} sweep(s), all 16 topics, full graph, dense-id CSC pull with OpenMP "
                      f"(transpose {build_s:.1f}s not counted)",
            "faithful": {"value": faithful, "unit": "GTEPS", "cores": 1,
                         "sample": f"1 sweep, 1 topic, {n_small}-node/{gs.n_edges}-edge graph, string-keyed hash maps "
                                   "as in ranking/pagerank.go (single goroutine)"}}


def cpu_scoring(eng, title, body, D, q, probs, cores, n_sample=None):
    """Oracle timed on the host cores: 'fair' (dense accumulators, shared blend term, bounded selection, OpenMP
    over queries) on a few thousand queries and 'faithful' (a hash map per query as in retrieval/main_retrieve.go)
    on 64; both return identical results (tests/test_oracle.py)."""
    from oracle import loader as O
    n_fair = min(len(q.kw_ptr) - 1, n_sample or 4096)
    n_faithful = min(len(q.kw_ptr) - 1, max(cores, 64))
    wt, mt = O.term_weights(title.term_ptr, title.doc_ids, title.norm_tf, D, float(D))
    wb, mb = O.term_weights(body.term_ptr, body.doc_ids, body.norm_tf, D, float(D))
    rng = np.random.default_rng(7)
    pr = (rng.random((D, T_TOPICS)) + 0.5) / D
    ot, ob = O.Table(title.term_ptr, title.doc_ids, wt), O.Table(body.term_ptr, body.doc_ids, wb)

    def timed(n, fair):
        kw_ptr = q.kw_ptr[: n + 1]
        kw = q.kw_terms[: int(kw_ptr[-1])]
        t0 = time.time()
        O.score_batch(ot, ob, D, mt, mb, pr, kw_ptr, kw, topic_probs=probs, k=TOP_K, n_threads=cores, fair=fair)
        return n / (time.time() - t0)

    fair = timed(n_fair, True)
    faithful = timed(n_faithful, False)
    return {"value": fair, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"first {n_fair} queries of the batch, dense-id accumulators per thread, blend term computed "
                      "once, bounded top-k selection, queries spread over all cores with OpenMP",
            "faithful": {"value": faithful, "unit": "queries/s", "cores": cores,
                         "sample": f"first {n_faithful} queries, a hash map per query as in retrieval/main_retrieve.go, "
                                   "queries spread over all cores"}}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; no Go toolchain in this image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import loader as O
    from spaghettisearch_b200 import synth
    cores = os.cpu_count() or 1
    world = args.gpus
    out = {"impl": "reference"}
    if args.workload in ("both", "pagerank"):
        g = synth.graph(args.nodes, args.edges, seed=42)
        npg = synth.topics(T_TOPICS)
        E = g.n_edges
        in_ptr, in_src = O.csc_build(g.row_ptr, g.col_idx)
        for _ in range(min(args.warmup, 1)):
            O.pagerank_fair_csc(g.row_ptr, in_ptr, in_src, DAMPING, EPS, npg, fixed_iters=1, n_threads=cores, want_rank=False)
        secs = 0.0
        for _ in range(args.steps):
            _, _, s = O.pagerank_fair_csc(g.row_ptr, in_ptr, in_src, DAMPING, EPS, npg, fixed_iters=1, n_threads=cores,
                                          want_rank=False)
            secs += s
        val = E * T_TOPICS * args.steps / secs / 1e9
        sample = (f"each step = 1 sweep of all 16 topics over the {args.nodes}-node/{E}-edge graph "
                  f"(the 1-GPU size; CPU rate is size independent), OpenMP on {cores} cores")
        out.update({"metric": "pagerank_gteps_per_iter", "value": val, "unit": "GTEPS (topic-edges/s, E*T per sweep)",
                    "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3 / args.steps,
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                    "data": "synthetic",
                    "config": {"workload": "BASELINE.json configs[1] on the host CPU (oracle port of ranking/pagerank.go)",
                               "nodes": args.nodes, "edges": E, "topics": T_TOPICS},
                    "cpu_baseline": {"value": val, "unit": "GTEPS", "cores": cores, "kind": "port", "sample": sample},
                    "e2e": {"value": val, "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0})
    if args.workload in ("both", "scoring"):
        D, V = args.docs, args.terms
        title = synth.index_table(V, D, 0)
        body = synth.index_table(V, D, 1)
        q = synth.queries(args.queries, V, seed=44)
        probs = np.full(T_TOPICS, 1.0 / T_TOPICS)
        n_sample = min(args.queries, 4096)
        res = None
        secs_total = 0.0
        for _ in range(args.steps):
            res = cpu_scoring(None, title, body, D, q, probs, cores, n_sample)
            secs_total += n_sample / res["value"]
        val = n_sample * args.steps / secs_total
        sc = {"metric": "scoring_queries_per_s", "value": val, "unit": "queries/s", "n_gpus": world,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs_total * 1e3 / args.steps,
              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 sums of f32 weights",
              "data": "synthetic",
              "config": {"workload": "BASELINE.json configs[2] on the host CPU (oracle port of retrieval.Retrieve core)",
                         "docs": D, "terms": V, "queries": args.queries, "k": TOP_K},
              "cpu_baseline": dict(res, value=val),
              "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
              "gpu_launches": 0, "impl": "reference"}
        if "metric" in out:
            out["scoring"] = sc
        else:
            out.update(sc)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["both", "pagerank", "scoring"], default="both")
    ap.add_argument("--nodes", type=int, default=10_000_000, help="graph nodes per GPU")
    ap.add_argument("--edges", type=int, default=150_000_000, help="graph edges per GPU")
    ap.add_argument("--docs", type=int, default=10_000_000, help="index docs per GPU")
    ap.add_argument("--terms", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=100_000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: contract asks for >= 3 warm-up steps")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
