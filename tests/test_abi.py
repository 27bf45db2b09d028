"""CPU: the C-ABI library builds, loads, and exports every symbol include/prisma_b200.h declares; the host-side
logic (net-size arithmetic, argument checks) works without a GPU.  No compute call is made here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from prisma_b200.build import build
    path = build()
    assert os.path.exists(path)
    return path


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "prisma_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(prisma_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(built_lib):
    l = ctypes.CDLL(built_lib)
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(l, s), f"{s} declared in include/prisma_b200.h but not exported"


def test_binding_covers_header(built_lib):
    from prisma_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    assert _lib.lib().prisma_version().decode().startswith("prisma_b200")


def test_no_gpu_fails_loudly(built_lib):
    """Without a GPU the product must raise, never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from prisma_b200.depth import DepthAnythingEngine
    from prisma_b200._lib import PrismaError
    with pytest.raises(PrismaError):
        DepthAnythingEngine("vits")


def test_net_size_host_logic(golden_dir):
    from prisma_b200.depth import da_net_size
    rows = np.load(os.path.join(golden_dir, "da_sizes.npz"))["rows"]
    for w, h, rw, rh in rows:
        assert da_net_size(int(w), int(h)) == (int(rw), int(rh))


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under prisma_b200/ or bands/ may reference it."""
    for base in ("prisma_b200", "bands"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dp, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


def test_bench_b200_arm_never_imports_oracle():
    """Only bench.py's cpu_baseline / --impl reference legs may touch oracle/ (the measured arm must be the CUDA product)."""
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"cpu_reference_step", "run_reference"}
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        mods = []
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom):
                mods.append(node.module or "")
            elif isinstance(node, ast.Import):
                mods.extend(a.name for a in node.names)
        bad = [m for m in mods if m.split(".")[0] == "oracle"]
        assert not bad or fn.name in allowed, (fn.name, bad)
