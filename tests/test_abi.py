"""CPU-side checks of the shipped library: it loads, exports every symbol include/kanzi_b200.h declares, and refuses to
run without a GPU (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(kz):
    hdr = open(os.path.join(ROOT, "include", "kanzi_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(kz_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    lib = kz.load_library()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(kz.ABI_SYMBOLS) == declared


def test_no_cpu_fallback(kz):
    lib = kz.load_library()
    if lib.kz_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(kz.KanziError):
        kz.Context(0)


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (only tests/, smoke() and bench.py's cpu legs may)."""
    pkg = os.path.join(ROOT, "kanzi-go_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in text and "kzo_" not in text and "oracle/" not in text, os.path.join(dirpath, f)


def test_type_name_parsing(kz):
    assert kz.transform_type("bwt") == 1 << 42
    assert kz.transform_type("LZX") == 16 << 42
    assert kz.entropy_type("ans0") == 5
    with pytest.raises(kz.KanziError):
        kz.transform_type("FOO")
    with pytest.raises(kz.KanziError):
        kz.Writer(block_size=1000)
    with pytest.raises(kz.KanziError):
        kz.Writer(jobs=65)
