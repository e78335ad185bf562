"""bench.py's reference arm (`--impl reference`) on the host cores: the JSON line contract the driver parses — same
metric / unit / config keys as the GPU arm, `impl: reference`, a `cpu_baseline` describing the run and an `e2e` that moves
no bytes.  Runs the reference's own Model::forward (oracle/_ref) on the smallest workload with a bounded step count."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_the_contract_line():
    if not (REPO / "oracle" / "_ref" / "libref.so").exists():
        pytest.skip("oracle/_ref is not built (needs /root/reference; __graft_entry__.build() compiles it)")
    r = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--workload", "gemma-3-1b-q4_0",
                        "--steps", "2", "--warmup", "1"], cwd=REPO, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, r.stdout[-1000:] + r.stderr[-1000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decode tok/s" and d["unit"] == "tok/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2
    assert d["value"] > 0 and d["ms_per_step"] > 0 and abs(d["value"] * d["ms_per_step"] / 1e3 - 1.0) < 0.05
    assert d["config"]["workload"] == "gemma-3-1b-q4_0" and d["vs_baseline"] is None and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == "tok/s" and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
