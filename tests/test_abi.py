"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol that
include/b200_hnsw.h declares, and refuses (loudly) to work without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import has_gpu
from hnsw_b200 import _lib


def _header_functions():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bh_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libb200hnsw.so not built (run __graft_entry__.build())"
    L = C.CDLL(_lib.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in b200_hnsw.h but not exported"
    # and the Python binding covers the whole header
    assert sorted(_lib.EXPORTED) == names


def test_no_torch_types_in_the_abi():
    src = open(_lib.HEADER_PATH).read()
    assert "torch" not in src.lower() and "Tensor" not in src and "#include <stdint.h>" in src


def test_version_and_error_channel():
    L = _lib.lib()
    assert b"sm_100a" in L.bh_version()
    h = C.c_void_p()
    assert L.bh_index_create(C.byref(h), 4096, 32, 1, 0) != 0       # d out of range
    assert "[1, 2048]" in _lib.last_error()
    assert L.bh_index_create(C.byref(h), 128, 100, 1, 0) != 0       # M too large
    assert L.bh_index_create(C.byref(h), 128, 32, 5, 0) != 0        # unknown metric
    assert L.bh_index_ntotal(None) == -1


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_a_gpu():
    import hnsw_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hnsw_b200.IndexHNSWFlat(128, 32)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under hnsw_b200/ may import, load or link it."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "hnsw_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(liboracle)|(orc_[a-z_]+\s*\()|(#include\s*[\"<].*oracle)", re.M)
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or fn == "Makefile":
                txt = open(os.path.join(dp, fn)).read()
                assert not bad.search(txt), f"{fn} reaches into oracle/"


def test_header_is_plain_c_and_the_cpp_example_links(tmp_path):
    """include/b200_hnsw.h must compile as C99 (it is a C-ABI), and examples/index_shards_b200.cpp — the
    faiss::IndexShards replacement written against the C-ABI only — must compile and link against the
    built library (running it needs GPUs)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PATH="/usr/bin:/bin:" + os.environ.get("PATH", ""))
    r = subprocess.run(["gcc", "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", _lib.HEADER_PATH],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    cuda = "/usr/local/cuda"
    if not (shutil.which("g++", path=env["PATH"]) and os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h"))):
        pytest.skip("no g++ / CUDA headers")
    _lib.lib()
    out = str(tmp_path / "index_shards_b200")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"),
                        os.path.join(root, "examples", "index_shards_b200.cpp"), "-L", os.path.join(root, "hnsw_b200"),
                        "-lb200hnsw", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lpthread", "-o", out],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
