"""bench.py's reference arm runs on the host cores only, so its JSON contract can be checked
without a GPU: one line on stdout, the keys the driver reads, the cpu_baseline / e2e objects."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line(built):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--nodes", "20000", "--edges", "300000", "--docs", "20000", "--terms", "5000", "--queries", "200"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    out = json.loads(lines[0])
    for obj in (out, out["scoring"]):
        assert obj["impl"] == "reference" or obj is out["scoring"]
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                    "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
            assert key in obj, key
        assert obj["value"] > 0 and obj["higher_is_better"] is True
        assert obj["cpu_baseline"]["kind"] == "port" and obj["cpu_baseline"]["cores"] >= 1
        assert obj["e2e"]["h2d_bytes_per_step"] == 0 and obj["e2e"]["d2h_bytes_per_step"] == 0
        assert "workload" in obj["config"]
    assert out["metric"] == "pagerank_gteps_per_iter" and out["scoring"]["metric"] == "scoring_queries_per_s"


def test_reference_arm_other_ranks_do_nothing(built):
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_pagerank_grid_policy():
    """bench.py's row x topic engine grid: covers the world, topic groups divide the 16 topics."""
    sys.path.insert(0, str(ROOT))
    import bench
    for world in (1, 2, 4, 8):
        rg, tg = bench.pr_grid(world, "")
        assert rg * tg == world and bench.T_TOPICS % tg == 0
    assert bench.pr_grid(8, "2x4") == (2, 4) and bench.pr_grid(1, "") == (1, 1)


def test_go_shim_files_reference_only_declared_entry_points():
    """integration/go/gpuengine/engine.go is the only cgo file; every C.ss_* it calls is declared in spaghetti.h."""
    import re
    header = (ROOT / "include" / "spaghetti.h").read_text()
    declared = set(re.findall(r"SS_API\s+[\w\s\*]+?\b(ss_\w+)\s*\(", header))
    go_dir = ROOT / "integration" / "go"
    used = set()
    for f in go_dir.rglob("*.go"):
        text = f.read_text()
        calls = set(re.findall(r"C\.(ss_\w+)\(", text))
        if f.name != "engine.go":
            assert 'import "C"' not in text and not calls, f
        used |= calls
    assert used and used <= declared, used - declared
    for name in ("ss_create", "ss_graph_load_csr", "ss_pagerank", "ss_index_load", "ss_term_weights", "ss_set_doc_norms",
                 "ss_set_pagerank", "ss_score_batch"):
        assert name in used, name
