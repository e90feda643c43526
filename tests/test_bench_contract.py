"""CPU: bench.py's reference arm runs here and prints one JSON line with the contract's keys; the work model is sane."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "smt_inclusion_proofs_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["workload"] == "smt_inclusion_dense_160_levels"


def test_work_model_matches_kernel_structure():
    sys.path.insert(0, str(ROOT))
    import bench

    # x^5 = 2 squarings (100 wide) + 1 multiply (128); a lazy dot row = t*64 + 64; rank-1 updates = 128 each
    # t = 3 (poseidon.cuh, partial rounds in pairs): 28 pairs of 11 products + 4 reductions, the first of the 57 rounds as
    # a round B with x0 = 0 (8 products + 3 reductions); one round at a time it was 57 * (328 + 256 + 2 * 128): 61 384
    assert bench.wide_per_hash(3, 57) == 8 * 3 * 328 + 7 * 3 * 256 + 256 + 57 * 328 + 28 * 15 * 64 + 11 * 64 == 59784
    assert bench.wide_per_hash(4, 56) == 8 * 4 * 328 + 7 * 4 * 320 + 320 + 56 * (328 + 320 + 3 * 128) == 77568
