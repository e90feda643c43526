"""CPU: the C-ABI library loads and exports every symbol include/gcp_b200.h declares (no compute calls)."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "gcp_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gnark_crypto_primitives_b200 import _lib, build

    build.build_library()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, f"{name} missing from the ctypes signature table"
    assert sorted(_lib.SIGNATURES) == names


def test_no_device_is_a_loud_error():
    import gnark_crypto_primitives_b200 as g
    from gnark_crypto_primitives_b200 import _lib

    lib = _lib.load()
    if lib.gcp_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(g.EngineError) as e:
        g.Engine(0)
    assert e.value.code == _lib.GCP_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The product path may not import, link or execute anything under oracle/ (comments may mention it)."""
    pkg = ROOT / "gnark_crypto_primitives_b200"
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle)|liboracle|oracle_[a-z_]+\s*\(|cport", re.M)
    for path in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        if path.suffix in (".py", ".cu", ".cuh", ".h"):
            assert not pat.search(path.read_text()), path


def test_header_is_plain_c(tmp_path):
    """include/gcp_b200.h must compile as C99 on its own (the cgo preamble includes nothing else) and as C++."""
    import subprocess

    src = tmp_path / "probe.c"
    src.write_text('#include "gcp_b200.h"\nint probe(void) { return GCP_OK + GCP_FMT_MONTGOMERY + GCP_STATUS_MALFORMED; }\n')
    for compiler, std in (("gcc", "-std=c99"), ("g++", "-std=c++17")):
        subprocess.run([compiler, std, "-Wall", "-Wextra", "-Werror", "-pedantic", "-x", "c" if compiler == "gcc" else "c++",
                        f"-I{ROOT / 'include'}", "-c", str(src), "-o", str(tmp_path / "probe.o")], check=True)


def test_group_and_host_alloc_are_loud_without_a_device():
    import gnark_crypto_primitives_b200 as g
    from gnark_crypto_primitives_b200 import _lib

    lib = _lib.load()
    if lib.gcp_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(g.EngineError) as e:
        g.Group([0])
    assert e.value.code == _lib.GCP_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(g.EngineError):
        g.PinnedBuffer(64)
    with pytest.raises(g.EngineError) as e:
        g.Group([])
    assert e.value.code == _lib.GCP_ERR_BAD_ARG


def _c_prototypes():
    """name -> number of parameters, parsed from the public header."""
    text = (ROOT / "include" / "gcp_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(gcp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = m.group(2).strip()
        protos[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return protos


def _go_calls(src):
    """(name, argument count) of every C.gcp_* call in the Go shim (balanced-parenthesis scan, top-level commas)."""
    calls = []
    for m in re.finditer(r"\bC\.(gcp_[a-z0-9_]+)\(", src):
        depth, i, commas, empty = 1, m.end(), 0, True
        while depth:
            ch = src[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                commas += 1
            if depth and not ch.isspace():
                empty = False
            i += 1
        calls.append((m.group(1), 0 if empty else commas + 1))
    return calls


def test_go_shim_matches_the_header():
    """The cgo shim cannot be compiled here (no Go toolchain): at least every C function it calls must be declared in
    include/gcp_b200.h with the number of arguments the shim passes, and every C constant it names must exist."""
    src = (ROOT / "go" / "gcpb200" / "gcpb200.go").read_text()
    protos = _c_prototypes()
    calls = _go_calls(src)
    assert len(calls) >= 20
    for name, nargs in calls:
        assert name in protos, f"{name} is not declared in gcp_b200.h"
        assert nargs == protos[name], f"{name}: shim passes {nargs} arguments, header declares {protos[name]}"
    header = (ROOT / "include" / "gcp_b200.h").read_text()
    for const in set(re.findall(r"\bC\.(GCP_[A-Z0-9_]+)\b", src)):
        assert re.search(rf"\b{const}\b", header), const
    # the C++ mirror calls the same ABI: same check on its gcp_* calls
    hpp = (ROOT / "include" / "gcp_b200.hpp").read_text()
    for name in set(re.findall(r"\b(gcp_[a-z0-9_]+)\s*\(", re.sub(r"//.*", "", hpp))):
        assert name in protos, name
