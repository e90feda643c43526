"""GPU: a C++ program binds the C ABI through include/gcp_b200.hpp (the reference-named mirror) and checks a KAT."""
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_cpp_program_links_and_runs(tmp_path):
    pkg = ROOT / "gnark_crypto_primitives_b200"
    exe = tmp_path / "test_hpp"
    subprocess.run(["g++", "-std=c++17", "-O1", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "cpp" / "test_hpp.cpp"),
                    f"-L{pkg}", "-lgcp_b200", f"-Wl,-rpath,{pkg}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "cpp mirror ok" in out.stdout
