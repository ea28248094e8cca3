"""The C-ABI library loads and exports every symbol include/spaghetti.h declares;
the product fails loudly without a GPU (no CPU fallback).  CPU only."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def gpu_lib():
    from spaghettisearch_b200 import _build
    path = _build.build_gpu()
    return ctypes.CDLL(str(path))


def test_exports_match_header(gpu_lib):
    header = (ROOT / "include" / "spaghetti.h").read_text()
    declared = re.findall(r"SS_API\s+[\w\s\*]+?\b(ss_\w+)\s*\(", header)
    assert len(declared) >= 18
    from spaghettisearch_b200 import capi
    assert sorted(set(declared)) == sorted(capi.EXPORTS)
    for name in declared:
        assert hasattr(gpu_lib, name), f"{name} declared in spaghetti.h but not exported"


def test_no_oracle_in_product():
    """Nothing under the package may import, link or call oracle/."""
    for p in (ROOT / "spaghettisearch_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h") and p.is_file():
            text = p.read_text()
            assert "liboracle" not in text and "oracle." not in text.replace("oracle/", ""), p
            assert "from oracle" not in text and "import oracle" not in text, p


def test_fails_loudly_without_gpu(gpu_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from spaghettisearch_b200 import capi
    with pytest.raises(capi.SSError) as ei:
        capi.Engine()
    assert ei.value.code == -6 and "no CPU path" in str(ei.value)


def test_header_is_c99_and_links_from_c(gpu_lib, tmp_path):
    """cgo compiles include/spaghetti.h as C: a C99 translation unit that takes the address of every entry point
    must compile with -pedantic -Werror, link against the library and behave (no device here => loud failure)."""
    import subprocess
    from spaghettisearch_b200 import _build
    exe = tmp_path / "c_abi_check"
    src = ROOT / "tests" / "c_abi_check.c"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I",
                    str(ROOT / "include"), "-o", str(exe), str(src), "-L", str(_build.PKG), "-lspaghetti_gpu",
                    f"-Wl,-rpath,{_build.PKG}"], check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "entry points" in r.stdout
    # every SS_API declaration of the header is referenced by the C file
    header = (ROOT / "include" / "spaghetti.h").read_text()
    declared = set(re.findall(r"SS_API\s+[\w\s\*]+?\b(ss_\w+)\s*\(", header))
    used = set(re.findall(r"\(fn_t\)(ss_\w+)", src.read_text()))
    assert declared == used, declared ^ used
