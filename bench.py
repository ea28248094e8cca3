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


def pagerank_workload(args, strong=False):
    return ("BASELINE.json configs[1]: topic-sensitive PageRank, 16 ODP topics, "
            f"{args.nodes} nodes / {args.edges} edges power-law graph {'in total' if strong else 'per GPU'}, "
            "fp64, eps 1e-9, d 0.75")


def scoring_workload(args):
    return (f"BASELINE.json configs[2]: batched cosine scoring, {args.queries} keyword queries over "
            f"{args.docs}-doc/GPU / {args.terms}-term synthetic Zipf index, PageRank blend + top-{TOP_K}")


CACHE_NOTE = ("inputs larger than L2, no flush between steps: per GPU the PageRank state and graph exceed 1 GB and the "
              "index 8 GB at the default sizes, L2 is 126 MB (exact sizes in `details`)")


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
def pr_grid(world, spec):
    """(row groups, topic groups) of the PageRank engine grid.  A GPU gathers ~36-45 G source rows/s whatever
    the row width (scripts/bench_gather2.cu), so splitting the 16 topics across GPUs halves a GPU's
    topic-edge rate per split; splitting rows costs an exchange of the replicated state that grows with the
    row group and moves at ~460 GB/s per GPU and direction when every GPU sends and receives at once.
    Measured (profiles/r02_multigpu.txt): 4 GPUs 4x1 924 / 2x2 1027 GTEPS; 8 GPUs 8x1 683 / 4x2 1586 / 2x4 1189."""
    if spec:
        rg, tg = (int(x) for x in spec.lower().split("x"))
        assert rg * tg == world and T_TOPICS % tg == 0, spec
        return rg, tg
    return {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}.get(world, (world, 1))


def run_c1(args):
    """BASELINE.json configs[0]: 10k pages / 50k terms / 100 queries written in the reference's JSON table
    encodings and run through the C++ host mirror of the three Go-API calls (csrc/host/), checked against
    the oracle.  Reports the wall-clock seconds per call (JSON decode/encode included) next to the oracle's."""
    import tempfile
    from tests import c1_workload as c1
    from spaghettisearch_b200 import capi
    eng = capi.Engine(device=0, timing=True)
    with tempfile.TemporaryDirectory() as d:
        tmp = Path(d)
        w = c1.make_tables(tmp)
        c1.run(eng, tmp, w)  # warm-up: allocations, the impact vectors of the first batch
        results, times = c1.run(eng, tmp, w)
        info = c1.check(tmp, w, results)
    eng.close()
    total = sum(v for k, v in times.items() if k not in ("load_snapshots", "score_batch_calls") and "cold" not in k)
    cpu_total = sum(info["cpu_seconds"].values())
    print(json.dumps({
        "metric": "c1_go_api_seconds", "value": total, "unit": "s", "n_gpus": 1, "steps": 1, "warmup": 1,
        "ms_per_step": total * 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 / f32 as the reference", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[0]: 10k-page link graph + 50k-term index in the reference's "
                               "JSON table encodings, UpdateTopicSensitivePagerank + UpdateTermWeights x2 + "
                               "Retrieve x100 through the host mirror of the Go API"},
        "seconds_per_call": times,
        "parity": dict(info, ok=True, against="oracle on the equivalent dense arrays: forw[3] within 1e-9 L1, "
                       "forw[4] bit-exact, result lists identical up to ties, scores within 1e-6"),
        "cpu_baseline": {"value": cpu_total, "unit": "s", "cores": 1, "kind": "port",
                         "sample": "the oracle on the dense arrays (no JSON): " + json.dumps(info["cpu_seconds"])},
        "e2e": {"value": total, "unit": "s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                "what": "table snapshots in host memory -> JSON decode -> C ABI -> JSON encode"},
        "gpu_launches": None}), flush=True)


def run_ours(args):
    if args.workload == "c1":
        return run_c1(args)
    import torch
    import torch.distributed as dist
    from spaghettisearch_b200 import capi, sharding, synth

    rank, local, world = dist_env(args)
    # stdout carries exactly one JSON line.  NCCL prints its version banner on the stdout of whichever process is
    # rank 0 of a communicator (with a row x topic grid that is several processes), so for the duration of the run
    # file descriptor 1 points at stderr; it is restored for the one line rank 0 prints at the end.
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    saved_stdout = None
    if world > 1:  # a single process creates no communicator: nothing to keep off stdout
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
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

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MAX) if world > 1 else x

    def min_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MIN) if world > 1 else x

    def sum_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.SUM) if world > 1 else x

    ctx = dict(torch=torch, dist=dist, capi=capi, sharding=sharding, synth=synth, rank=rank, local=local, world=world,
               cores=cores, threads=threads, peak=peak, peak_src=peak_src, barrier=barrier,
               max_over_ranks=max_over_ranks, min_over_ranks=min_over_ranks, sum_over_ranks=sum_over_ranks)
    out = {}
    if args.workload in ("both", "pagerank"):
        out.update(run_pagerank(args, ctx))
    if args.workload in ("both", "scoring"):
        sc = run_scoring(args, ctx)
        if "metric" in out:
            out["scoring"] = sc
        else:
            out.update(sc)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(out), flush=True)


def make_group_engine(ctx, group_rank, group_size, group_id):
    """One engine on this GPU whose communicator spans the `group_size` ranks of group `group_id`."""
    dist, capi, rank = ctx["dist"], ctx["capi"], ctx["rank"]
    eng = capi.Engine(device=ctx["local"], timing=True)
    if group_size > 1:
        eng.comm_init(ctx["sharding"].share_group_unique_id(capi.comm_unique_id, group_rank, group_size), group_rank,
                      group_size)
    return eng


def pagerank_parity(ctx, rg, tg, g_rank, t_g, t_lo):
    """Correctness evidence carried by every line: a small graph (1M nodes / 15M edges) goes through exactly
    the path that is timed (same grid, sharded load, chunked overlapped exchange) and every row this rank
    owns is compared with the oracle (ranking/pagerank.go restated, CSC pull) for this rank's topics."""
    from oracle import loader as O
    capi, sharding, synth = ctx["capi"], ctx["sharding"], ctx["synth"]
    n, e_target = 1_000_000, 15_000_000
    npg = synth.topics(T_TOPICS)[t_lo:t_lo + t_g]
    eng = make_group_engine(ctx, g_rank, rg, ctx["rank"] // rg)
    try:
        lo, hi = sharding.row_slice(g_rank, rg, n)
        part = synth.graph_rows(n, e_target, lo, hi, seed=42, n_threads=ctx["threads"])
        eng.graph_load_csr_rows(n, lo, hi, part.row_ptr, part.col_idx)
        _, iters, status = eng.pagerank(DAMPING, EPS, npg, want_rank=False)
        st = eng.pagerank_stats()
        own = eng.pagerank_fetch(int(st.row_lo), int(st.row_lo + st.local_rows))
        full = synth.graph(n, e_target, seed=42, n_threads=ctx["threads"])
        ref, it_ref, _ = O.pagerank_fair(full.row_ptr, full.col_idx, DAMPING, EPS, npg, n_threads=ctx["threads"])
        ref_own = ref[int(st.row_lo): int(st.row_lo + st.local_rows)]
        l1_own = float(np.abs(own - ref_own).sum(axis=0).max()) if len(ref_own) else 0.0
        same_iters = iters.tolist() == it_ref.tolist()
    finally:
        eng.close()
    l1 = ctx["sum_over_ranks"](l1_own) if rg > 1 else l1_own   # per-topic L1 over all rows <= sum of the blocks' maxima
    l1 = ctx["max_over_ranks"](l1)
    ok = ctx["min_over_ranks"](1.0 if (status == 0 and same_iters and l1 <= 1e-9) else 0.0) == 1.0
    return {"ok": bool(ok), "graph": f"{n} nodes / {full.n_edges} edges through the timed path (grid {rg}x{tg})",
            "max_l1_per_topic": l1, "bar": 1e-9, "equal_sweep_counts": bool(same_iters),
            "against": "oracle (ranking/pagerank.go:85-145 restated), every row of every rank"}


def run_pagerank(args, ctx):
    torch, capi, sharding, synth = ctx["torch"], ctx["capi"], ctx["sharding"], ctx["synth"]
    rank, local, world, threads, cores = ctx["rank"], ctx["local"], ctx["world"], ctx["threads"], ctx["cores"]
    peak, peak_src, barrier = ctx["peak"], ctx["peak_src"], ctx["barrier"]
    rg, tg = pr_grid(world, args.pr_grid)
    g_rank, t_group, t_lo, t_g = sharding.engine_grid(rank, rg, tg, T_TOPICS)
    strong = args.scaling == "strong"
    n_nodes = args.nodes if strong else args.nodes * world
    n_edges_target = args.edges if strong else args.edges * world
    parity = None if args.no_parity else pagerank_parity(ctx, rg, tg, g_rank, t_g, t_lo)

    eng = make_group_engine(ctx, g_rank, rg, t_group)
    ext = torch.cuda.ExternalStream(eng.stream_handle(), device=torch.device("cuda", local))
    t0 = time.time()
    lo, hi = sharding.row_slice(g_rank, rg, n_nodes)  # the slice of the out-edge CSR this rank exports
    if rg == 1:
        g = synth.graph(n_nodes, n_edges_target, seed=42, n_threads=threads)
    else:
        g = synth.graph_rows(n_nodes, n_edges_target, lo, hi, seed=42, n_threads=threads)
    row_ptr, col_idx = g.row_ptr, g.col_idx
    gen_s = time.time() - t0
    npg_all = synth.topics(T_TOPICS)
    npg = npg_all[t_lo:t_lo + t_g]
    h_row_ptr, h_col = pin(row_ptr), pin(col_idx)
    p_row_ptr, p_col = pin_view(h_row_ptr, np.uint64), pin_view(h_col, np.uint32)
    eng.graph_load_csr_rows(n_nodes, lo, hi, p_row_ptr, p_col)
    st0 = eng.pagerank_stats()
    E, R, E_loc = int(st0.n_edges), int(st0.local_rows), int(st0.local_edges)
    log(f"[rank {rank}] graph N={n_nodes} E={E} slice {lo}:{hi} generated in {gen_s:.1f}s; owns {R} rows / {E_loc} "
        f"in-edges, {t_g} topics (grid {rg}x{tg}), load {st0.load_ms:.1f} ms")
    for _ in range(args.warmup):
        eng.pagerank(DAMPING, EPS, npg, want_rank=False)
    sweeps_total, launches, sweep_ms, gather_ms, exch_ms, busy_ms = 0, 0, 0.0, 0.0, 0.0, 0.0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clk:
        ev0.record(ext)
        for _ in range(args.steps):
            eng.pagerank(DAMPING, EPS, npg, want_rank=False)
            s = eng.pagerank_stats()
            sweeps_total += s.sweeps
            launches += s.launches
            sweep_ms += s.sweep_ms_total
            gather_ms += s.gather_ms_total
            exch_ms += s.exchange_ms_total
            busy_ms += s.exchange_busy_ms_total
        ev1.record(ext)
        barrier()
    ms = ctx["max_over_ranks"](ev0.elapsed_time(ev1))
    clocks = clk.summary()
    # topic-edges swept by all ranks: every rank sweeps its in-edges for its topics
    work = ctx["sum_over_ranks"](float(E_loc) * t_g * sweeps_total)
    gteps = work / (ms * 1e-3) / 1e9
    # roofline of the sweep (dominant kernels: k_sweep_short32 + k_sweep_long), this rank's rows and topics
    b_pr = 4 * E_loc + 8 * (R + 1) + 8 * R + 16 * t_g * R
    avg_sweep_s = sweep_ms * 1e-3 / max(1, sweeps_total)
    achieved = b_pr / avg_sweep_s / 1e9
    gather_bytes = E_loc * 8 * t_g  # gather-traffic model: every edge moves one 8*T-byte row

    # e2e through the C ABI with pinned host buffers: export slice H2D + device transpose (+ edge exchange) +
    # run + this rank's rows x topics D2H
    h_rank = torch.empty(R * t_g, dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    e2e_work = 0.0
    for _ in range(e2e_steps):
        eng.graph_load_csr_rows(n_nodes, lo, hi, p_row_ptr, p_col)
        if world == 1:
            eng.pagerank(DAMPING, EPS, npg, out=h_rank.numpy().reshape(n_nodes, t_g))
        else:
            eng.pagerank(DAMPING, EPS, npg, want_rank=False)
            eng.pagerank_fetch(int(st0.row_lo), int(st0.row_lo) + R, out=h_rank.numpy().reshape(R, t_g))
        e2e_work += float(E_loc) * t_g * eng.pagerank_stats().sweeps
    barrier()
    e2e_s = ctx["max_over_ranks"](time.perf_counter() - t0)
    e2e_gteps = ctx["sum_over_ranks"](e2e_work) / e2e_s / 1e9
    h2d = int(p_row_ptr.nbytes + p_col.nbytes)
    d2h = R * t_g * 8 + 3 * 2 * t_g * 8 * sweeps_total // max(1, args.steps)

    out = {
        "metric": "pagerank_gteps_per_iter", "value": gteps, "unit": "GTEPS (topic-edges/s, E*T per sweep)",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # `config` is the workload only (the reference arm prints the same dict); how it was run is in `details`
        "config": {"workload": pagerank_workload(args, strong), "nodes": n_nodes, "edges": E, "topics": T_TOPICS,
                   "cache": CACHE_NOTE},
        "details": {"sweeps_per_step": sweeps_total / args.steps,
                    "cache": "inputs larger than L2 (state 2x%.2f GB + graph %.2f GB vs 126 MB L2)" %
                             (n_nodes * t_g * 8 / 1e9, (4 * E_loc + 8 * R) / 1e9),
                    "partition": (f"{rg} row groups x {tg} topic groups; rows edge-balanced; sharded export "
                                  "(ss_graph_load_csr_rows: one all-to-all of the edges at load)") if world > 1
                                 else "single GPU",
                    "collective": ("per sweep inside a row group: every rank pushes its finished row chunks into the "
                                   "peers' state with copy-engine peer copies over NVLink (CUDA IPC), overlapped with "
                                   "the sweep of the next chunk on an exchange stream (SS_PR_EXCHANGE=nccl: grouped "
                                   "ncclBroadcast instead), + ncclAllReduce of 3*T sums as the barrier" if rg > 2 else
                                   "2-rank row group: sweep epilogue pushes rows into the peer's state over NVLink "
                                   "(CUDA IPC peer memory), ncclAllReduce of 3*T sums as the barrier" if rg == 2 else
                                   "none (topics are independent)")},
        "clocks": clocks,
        "e2e": {"value": e2e_gteps, "unit": "GTEPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3 / e2e_steps, "load_ms": eng.pagerank_stats().load_ms,
                "what": "ss_graph_load_csr_rows (pinned host CSR slice, device transpose, edge exchange) + "
                        "ss_pagerank + this rank's rows to pinned host memory; bytes are per rank"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "frac_of_nominal_8TBps": achieved / 8000.0,
                     "traffic": ncu_traffic("pagerank_sweep_dram_bytes") if world == 1 else None,
                     "peak_source": peak_src, "kernel": "k_sweep_short32 + k_sweep_long (one sweep, one rank)",
                     "algorithmic_bytes_per_sweep": b_pr, "avg_sweep_ms": avg_sweep_s * 1e3,
                     "gather_model_bytes_per_sweep": b_pr + gather_bytes,
                     "gather_model_GBps": (b_pr + gather_bytes) / avg_sweep_s / 1e9,
                     "gathered_rows_per_s": E_loc / avg_sweep_s,
                     "exchange_exposed_ms_per_sweep": exch_ms / max(1, sweeps_total),
                     "exchange_busy_ms_per_sweep": busy_ms / max(1, sweeps_total),
                     "exchange_hidden_frac": (1.0 - exch_ms / busy_ms) if busy_ms > 0 else None},
        "edges_per_s": gteps * 1e9 / T_TOPICS,
    }
    if parity is not None:
        out["parity"] = parity
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_pagerank(row_ptr, col_idx, npg_all, cores)
    del h_rank
    eng.close()
    return out


def scoring_parity(ctx, eng_factory):
    """A 400K-doc index with positions, doc-sharded exactly like the timed one (shard-local ids, global df,
    in-engine all-gather + merge), 512 queries of which 25 % carry a phrase: merged top-10 of every query
    against the oracle on the unsharded index."""
    from oracle import loader as O
    capi, sharding, synth = ctx["capi"], ctx["sharding"], ctx["synth"]
    rank, world, threads = ctx["rank"], ctx["world"], ctx["threads"]
    V, D, Q = 50_000, 400_000, 512
    lo, hi = sharding.doc_shard(rank, world, D)
    rng = np.random.default_rng(7)
    pr = (rng.random((D, T_TOPICS)) + 0.5) / D
    probs = np.full(T_TOPICS, 1.0 / T_TOPICS)
    q = synth.queries(Q, V, phrase_fraction=0.25, seed=45)
    eng = eng_factory()
    try:
        eng.index_set_doc_base(lo)
        full = {}
        for tid in (capi.SS_TITLE, capi.SS_BODY):
            t = synth.index_table(V, D, tid, doc_lo=lo, doc_hi=hi, with_positions=True, n_threads=threads)
            eng.index_load(tid, hi - lo, t.term_ptr, t.doc_ids - np.uint32(lo), t.norm_tf, t.pos_ptr, t.pos)
            eng.term_weights(tid, float(D), t.n_postings, hi - lo, df_global=t.df_global, want=False)
        eng.set_pagerank(pr[lo:hi])
        got = eng.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=TOP_K, sharded=True)
    finally:
        eng.close()
    ok = True
    if rank == 0:
        for tid in (0, 1):
            t = synth.index_table(V, D, tid, with_positions=True, n_threads=threads)
            w, mag = O.term_weights(t.term_ptr, t.doc_ids, t.norm_tf, D, float(D))
            full[tid] = (O.Table(t.term_ptr, t.doc_ids, w, t.pos_ptr, t.pos), mag)
        exp = O.score_batch(full[0][0], full[1][0], D, full[0][1], full[1][1], pr, q.kw_ptr, q.kw_terms, q.ph_ptr,
                            q.ph_terms, topic_probs=probs, k=TOP_K, n_threads=threads, fair=True)
        ok = bool(np.array_equal(got[0], exp[0]) and np.array_equal(got[3], exp[3]) and
                  np.allclose(got[1], exp[1], rtol=1e-6, atol=0) and np.allclose(got[2], exp[2], rtol=1e-6, atol=0))
    ok = ctx["min_over_ranks"](1.0 if ok else 0.0) == 1.0
    return {"ok": bool(ok), "index": f"{D} docs / {V} terms with positions, {world} doc shard(s)", "queries": Q,
            "phrase_fraction": 0.25, "against": "oracle (retrieval.Retrieve core restated) on the unsharded index: "
            "identical top-10 ids, order and counts, scores within 1e-6 relative"}


def run_scoring(args, ctx):
    torch, capi, sharding, synth = ctx["torch"], ctx["capi"], ctx["sharding"], ctx["synth"]
    rank, local, world, threads, cores = ctx["rank"], ctx["local"], ctx["world"], ctx["threads"], ctx["cores"]
    peak, peak_src, barrier = ctx["peak"], ctx["peak_src"], ctx["barrier"]
    max_over_ranks, sum_over_ranks = ctx["max_over_ranks"], ctx["sum_over_ranks"]

    def engine():
        return make_group_engine(ctx, rank, world, 0)  # one communicator over all doc shards

    parity = None if args.no_parity else scoring_parity(ctx, engine)
    eng = engine()
    V, Q = args.terms, args.queries
    mixed = args.phrase_fraction > 0
    state = {}

    def load_index(docs_per_gpu, with_positions):
        """Doc shard of this rank under shard-local ids, global df, synthetic forw[3] rows."""
        D = docs_per_gpu * world
        lo, hi = sharding.doc_shard(rank, world, D)
        t0 = time.time()
        title = synth.index_table(V, D, 0, doc_lo=lo, doc_hi=hi, with_positions=with_positions, n_threads=threads)
        body = synth.index_table(V, D, 1, doc_lo=lo, doc_hi=hi, with_positions=with_positions, n_threads=threads)
        log(f"[rank {rank}] index D={D} (shard {lo}:{hi}) V={V} postings title={title.n_postings} "
            f"body={body.n_postings} positions={with_positions} generated in {time.time() - t0:.1f}s")
        eng.index_clear()
        eng.index_set_doc_base(lo)
        t0 = time.time()
        for tid, t in ((capi.SS_TITLE, title), (capi.SS_BODY, body)):
            ids = t.doc_ids - np.uint32(lo) if lo else t.doc_ids
            eng.index_load(tid, hi - lo, t.term_ptr, ids, t.norm_tf, t.pos_ptr, t.pos)
            eng.term_weights(tid, float(D), t.n_postings, hi - lo, df_global=t.df_global, want=False)
            del ids
        # forw[3] rows for the blend: synthetic ranks around 1/D (the reference's ranks are near uniform)
        rng = np.random.default_rng(7)
        eng.set_pagerank((rng.random((hi - lo, T_TOPICS)) + 0.5) / D)
        state.update(title=title, body=body, D=D, load_s=time.time() - t0)

    load_index(args.docs, mixed and args.mixed_docs == 0)
    title, body, D, load_s = state["title"], state["body"], state["D"], state["load_s"]
    q = synth.queries(Q, V, phrase_fraction=0.0, seed=44)
    probs = np.full(T_TOPICS, 1.0 / T_TOPICS)

    def timed_leg(qb, steps, warmup):
        """steps x ss_score_batch_sharded over the resident index; returns the accumulated stats."""
        has_ph = int(qb.ph_ptr[-1]) > 0
        hp = [pin(a) for a in (qb.kw_ptr, qb.kw_terms, qb.ph_ptr, qb.ph_terms if has_ph else np.zeros(1, np.uint32))]
        kw_ptr, ph_ptr = pin_view(hp[0], np.uint64), pin_view(hp[2], np.uint64)
        kw, ph = pin_view(hp[1], np.uint32), pin_view(hp[3], np.uint32)
        nq = len(kw_ptr) - 1
        bufs = (torch.empty(nq * TOP_K, dtype=torch.int32).pin_memory(),
                torch.empty(nq * TOP_K, dtype=torch.float64).pin_memory(),
                torch.empty(nq * TOP_K, dtype=torch.float64).pin_memory(), torch.empty(nq, dtype=torch.int32).pin_memory())
        outs = (bufs[0].numpy().view(np.uint32).reshape(nq, TOP_K), bufs[1].numpy().reshape(nq, TOP_K),
                bufs[2].numpy().reshape(nq, TOP_K), bufs[3].numpy().view(np.uint32))
        call = lambda: eng.score_batch(kw_ptr, kw, ph_ptr if has_ph else None, ph if has_ph else None,
                                       topic_probs=probs, k=TOP_K, out=outs, sharded=True)
        for _ in range(warmup):
            call()
        acc = {"kernel_ms": 0.0, "score_ms": 0.0, "merge_ms": 0.0, "launches": 0, "alg_bytes": 0, "model_bytes": 0,
               "postings": 0}
        barrier()
        with ClockSampler(local) as clk:
            t0 = time.perf_counter()
            for _ in range(steps):
                call()
                s = eng.score_stats()
                acc["kernel_ms"] += s.kernel_ms
                acc["score_ms"] += s.score_kernel_ms
                acc["merge_ms"] += s.shard_merge_ms
                acc["launches"] += s.launches
                acc["alg_bytes"] += s.algorithmic_bytes
                acc["model_bytes"] += s.model_bytes
                acc["postings"] += s.postings_scanned
            barrier()
            acc["wall_s"] = max_over_ranks(time.perf_counter() - t0)
        acc["kernel_ms"] = max_over_ranks(acc["kernel_ms"])
        acc["clocks"] = clk.summary()
        acc["h2d"] = int(kw_ptr.nbytes + kw.nbytes + probs.nbytes + (ph_ptr.nbytes + ph.nbytes if has_ph else 0))
        acc["d2h"] = int(sum(o.nbytes for o in outs))
        acc["nq"] = nq
        acc["call"] = call
        return acc

    a = timed_leg(q, args.steps, args.warmup)
    kernel_ms, score_ms = a["kernel_ms"], a["score_ms"]
    # The same batch with the impact-vector path and the cross-slab bound switched off: every posting of
    # every query list is walked (SURVEY.md 8(d)'s "batched sparse gather" taken literally); same results.
    os.environ["SS_SCORE_DENSE"] = "0"
    os.environ["SS_SCORE_QTHR"] = "0"
    a["call"]()
    s_walk = eng.score_stats()
    del os.environ["SS_SCORE_DENSE"], os.environ["SS_SCORE_QTHR"]
    walk_ms = max_over_ranks(s_walk.kernel_ms)
    walk = {"value": Q / (walk_ms * 1e-3), "unit": "queries/s", "ms_per_step": walk_ms, "steps": 1,
            "roofline_frac": s_walk.algorithmic_bytes / (s_walk.score_kernel_ms * 1e-3) / 1e9 / peak,
            "what": "SS_SCORE_DENSE=0 SS_SCORE_QTHR=0: posting lists walked for every query (SURVEY.md 8(d) bytes: "
                    "8 B per posting + per matched doc + per result), same results"}
    model_frac = a["model_bytes"] / (score_ms * 1e-3) / 1e9 / peak
    sc = {
        "metric": "scoring_queries_per_s", "value": Q * args.steps / (kernel_ms * 1e-3), "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 sums of f32 weights",
        "data": "synthetic",
        "config": {"workload": scoring_workload(args), "docs": D, "terms": V, "queries": Q, "k": TOP_K,
                   "cache": CACHE_NOTE},
        "details": {"postings": int(sum_over_ranks(title.n_postings + body.n_postings)),
                    "cache": "index %.1f GB larger than L2" % ((title.n_postings + body.n_postings) * 8 / 1e9),
                    "shard": ("docs, shard-local ids + doc base; every rank scores the whole batch, ncclAllGather of the "
                              "[Q][k] lists + k-way merge on the engine stream inside the timed region "
                              "(ss_score_batch_sharded)") if world > 1 else "single GPU",
                    "index_load_s": load_s},
        "clocks": a["clocks"],
        "e2e": {"value": Q * args.steps / a["wall_s"], "unit": "queries/s", "h2d_bytes_per_step": a["h2d"],
                "d2h_bytes_per_step": a["d2h"], "ms_per_step": a["wall_s"] * 1e3 / args.steps,
                "what": "ss_score_batch_sharded with pinned host query/result buffers, index resident"},
        "gpu_launches": a["launches"],
        "shard_merge_ms_per_step": a["merge_ms"] / args.steps,
        "roofline": {"bound": "hbm", "achieved": a["model_bytes"] / (score_ms * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": model_frac, "frac_of_nominal_8TBps": model_frac * peak / 8000.0,
                     "traffic": ncu_traffic("score_dram_bytes") if world == 1 else None, "peak_source": peak_src,
                     "kernel": "k_score",
                     "bytes_model": "what the default path has to move: a query with a dense keyword streams 2 B per "
                                    "doc for each of its dense tokens and reads 8 B per "
                                    "posting of its other tokens; any other query reads 8 B per posting of every "
                                    "list; + 12 B per result.  Survivor lookups are not counted.",
                     "model_bytes_per_batch": a["model_bytes"] // max(1, args.steps),
                     "nominal_bytes_per_batch": a["alg_bytes"] // max(1, args.steps),
                     "nominal_frac": a["alg_bytes"] / (score_ms * 1e-3) / 1e9 / peak,
                     "postings_per_batch": a["postings"] // max(1, args.steps),
                     "avg_kernel_ms": score_ms / max(1, args.steps),
                     "note": "frac is against the bytes of the model above; it is far below 1 because the kernel is "
                             "bound by instruction issue and per-CTA latency, not by HBM (DESIGN.md 5).  "
                             "nominal_frac charges SURVEY.md 8(d)'s 8 B for every posting of every query list, which "
                             "the impact-vector path does not read: a speed-up statement, not a bandwidth figure; "
                             "walk_every_posting is the run that does read them."},
        "walk_every_posting": walk,
    }
    if parity is not None:
        sc["parity"] = parity
    # single-query latency through the same entry point (retrieval.Retrieve is called once per HTTP request,
    # cmd/server/server.go:47): wall clock of 200 one-query calls, index resident
    lat = []
    for i in range(200):
        a0, a1 = int(q.kw_ptr[i]), int(q.kw_ptr[i + 1])
        one_ptr, one_kw = np.array([0, a1 - a0], np.uint64), np.ascontiguousarray(q.kw_terms[a0:a1])
        t0 = time.perf_counter()
        eng.score_batch(one_ptr, one_kw, topic_probs=probs, k=TOP_K, sharded=True)
        lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    sc["single_query_latency_ms"] = {"p50": lat[len(lat) // 2], "p90": lat[int(len(lat) * 0.9)], "max": lat[-1],
                                     "what": "one ss_score_batch_sharded call per query, first 200 queries of the batch"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_scoring(eng, title, body, D, q, probs, cores)
    if mixed:
        # BASELINE.json configs[4] flavour: a fresh batch of which a fifth carries a 2-3-term phrase, over a
        # doc-sharded index with positions (--mixed-docs 0: the main index itself, loaded with positions)
        if args.mixed_docs:
            del title, body
            state.clear()
            load_index(args.mixed_docs, True)
        qm = synth.queries(args.mixed_queries, V, phrase_fraction=args.phrase_fraction, seed=46)
        steps_m = max(1, min(args.steps, 3))
        m = timed_leg(qm, steps_m, 1)
        sc["mixed"] = {"value": m["nq"] * steps_m / (m["kernel_ms"] * 1e-3), "unit": "queries/s",
                       "queries": m["nq"], "phrase_fraction": args.phrase_fraction, "steps": steps_m,
                       "docs": state["D"], "ms_per_step": m["kernel_ms"] / steps_m,
                       "e2e": {"value": m["nq"] * steps_m / m["wall_s"], "unit": "queries/s",
                               "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"]},
                       "shard_merge_ms_per_step": m["merge_ms"] / steps_m, "gpu_launches": m["launches"],
                       "what": "BASELINE.json configs[4] flavour: phrase + keyword batch over the doc-sharded index "
                               "with positions, PageRank blend, top-10, cross-shard merge in the timed region"}
    if cpu is not None:
        sc["cpu_baseline"] = cpu
    eng.close()
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
                    "config": {"workload": pagerank_workload(args), "nodes": args.nodes, "edges": E, "topics": T_TOPICS,
                               "cache": CACHE_NOTE},
                    "details": {"what": "the oracle port of ranking/pagerank.go on the host CPU; at N > 1 the "
                                        "1-GPU size is timed (the CPU rate does not depend on the size)"},
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
              "config": {"workload": scoring_workload(args), "docs": D, "terms": V, "queries": args.queries, "k": TOP_K,
                         "cache": CACHE_NOTE},
              "details": {"what": "the oracle port of the retrieval.Retrieve core on the host CPU"},
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
    ap.add_argument("--workload", choices=["both", "pagerank", "scoring", "c1"], default="both")
    ap.add_argument("--nodes", type=int, default=10_000_000, help="graph nodes per GPU")
    ap.add_argument("--edges", type=int, default=150_000_000, help="graph edges per GPU")
    ap.add_argument("--docs", type=int, default=10_000_000, help="index docs per GPU")
    ap.add_argument("--terms", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=100_000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the small-problem parity passes")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: --nodes/--edges per GPU; strong: in total (configs[3]: --nodes 100000000 --edges 1500000000)")
    ap.add_argument("--pr-grid", default="", help="RGxTG: row groups x topic groups of the PageRank grid")
    ap.add_argument("--phrase-fraction", type=float, default=0.2,
                    help="phrase share of the mixed scoring leg (0: no mixed leg, index without positions)")
    ap.add_argument("--mixed-queries", type=int, default=100_000)
    ap.add_argument("--mixed-docs", type=int, default=2_000_000,
                    help="docs per GPU of the positions-carrying index of the mixed leg (0: the main index itself)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: contract asks for >= 3 warm-up steps")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
