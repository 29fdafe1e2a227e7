"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(snpg_[a-z0-9_]+)\s*\(", text))
    return names


def test_every_declared_symbol_is_exported_and_bound():
    from snpgenie_b200 import _abi
    decl = declared_symbols()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(_abi.lib, name), f"{name} declared in include/ but not exported"
    assert decl == set(_abi.SIGNATURES), (decl ^ set(_abi.SIGNATURES))


def test_python_constants_follow_the_header():
    """Option numbers and values in the ctypes mirror are the header's #defines."""
    from snpgenie_b200 import _abi
    text = open(os.path.join(ROOT, "include", "snpgenie_cuda.h")).read()
    defs = {k: int(v, 0) for k, v in re.findall(r"^#define\s+(SNPG_[A-Z0-9_]+)\s+(-?(?:0x)?[0-9A-Fa-f]+)\s*$", text, flags=re.M)}
    pairs = {"SNPG_OPT_KEEP_CODES": _abi.OPT_KEEP_CODES, "SNPG_OPT_PROFILE": _abi.OPT_PROFILE, "SNPG_OPT_SITE_COUNTS": _abi.OPT_SITE_COUNTS,
             "SNPG_OPT_FUSED_KERNEL": _abi.OPT_FUSED_KERNEL, "SNPG_OPT_BOOTSTRAP_KERNEL": _abi.OPT_BOOTSTRAP_KERNEL,
             "SNPG_FUSED_TMA_TILES": _abi.FUSED_TMA_TILES, "SNPG_FUSED_LDG_SPANS": _abi.FUSED_LDG_SPANS,
             "SNPG_FUSED_TMA_SPANS": _abi.FUSED_TMA_SPANS,
             "SNPG_BOOT_GATHER": _abi.BOOT_GATHER, "SNPG_BOOT_WEIGHTS": _abi.BOOT_WEIGHTS, "SNPG_BOOT_CLASSES": _abi.BOOT_CLASSES, "SNPG_OPT_BOOT_SLOTS": _abi.OPT_BOOT_SLOTS,
             "SNPG_OPT_INGEST_ASYNC": _abi.OPT_INGEST_ASYNC, "SNPG_OPT_FIRST_DEFINED": _abi.OPT_FIRST_DEFINED, "SNPG_RWIN_COLS": _abi.RWIN_COLS}
    for name, value in pairs.items():
        assert defs[name] == value, name
    assert len({defs[k] for k in defs if k.startswith("SNPG_OPT_")}) == len([k for k in defs if k.startswith("SNPG_OPT_")])


def test_host_only_entry_points_work_without_a_gpu():
    from oracle import oracle as orc
    from snpgenie_b200 import _abi, engine
    assert b"sm_100a" in _abi.lib.snpg_version()
    sites, diffs = engine.ng_tables()                      # the product's own table builder ...
    osites, odiffs = orc.tables()                          # ... against the oracle's (independent code)
    assert np.array_equal(sites, osites) and np.array_equal(diffs, odiffs)
    f, c = ctypes.c_int64(), ctypes.c_int64()
    assert _abi.lib.snpg_shard_plan(9677, 3, 8, ctypes.byref(f), ctypes.byref(c)) == 0
    assert (f.value, c.value) == (9677 * 3 // 8, 9677 * 4 // 8 - 9677 * 3 // 8)
    assert _abi.lib.snpg_shard_plan(10, 2, 2, ctypes.byref(f), ctypes.byref(c)) == _abi.E_ARG


def test_no_silent_cpu_fallback():
    """Without a device the product must fail loudly, never compute on the host."""
    import torch
    from snpgenie_b200 import _abi, engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_abi.SnpgError) as e:
        engine.Engine()
    assert e.value.code == _abi.E_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "snpgenie_b200", "**", "*"), recursive=True) + glob.glob(os.path.join(ROOT, "perl", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".h", ".cpp", ".pl", ".pm", ".xs")):
            text = open(path, errors="ignore").read()
            assert "oracle" not in text.replace("the oracle's", ""), path
