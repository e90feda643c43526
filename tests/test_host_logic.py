"""CPU: the host-side logic under the C ABI that needs no GPU - chunk schedule of the host-buffer pipelines
(csrc/chunkplan.h) and the process-wide copy pool (csrc/hostcopy.h) - as a plain C++ program."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_chunk_plan_and_copy_pool(tmp_path):
    exe = tmp_path / "test_host_logic"
    csrc = ROOT / "gnark_crypto_primitives_b200" / "csrc"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-pthread", f"-I{csrc}",
                    str(ROOT / "tests" / "cpp" / "test_host_logic.cpp"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "host logic ok" in out.stdout
