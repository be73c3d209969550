"""CPU: the C-ABI library builds, loads and exports every symbol include/isa_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from isa_b200 import _lib
    return _lib.load()


def _declared():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            src = open(os.path.join(inc, f)).read()
            names |= set(re.findall(r"\b(isa_[a-z0-9_]+)\s*\(", src))
    return names


def test_header_symbols_exported(lib):
    from isa_b200 import _lib
    declared = _declared()
    assert declared, "include/isa_b200.h declares nothing?"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(raw, n)]
    assert not missing, "declared in include/ but not exported: %s" % missing
    unbound = [n for n in sorted(declared) if n not in _lib.SIGNATURES]
    assert not unbound, "declared in include/ but not bound in _lib.SIGNATURES: %s" % unbound
    undeclared = [n for n in _lib.SIGNATURES if n not in declared]
    assert not undeclared, "bound in _lib.py but missing from the header: %s" % undeclared


def test_no_gpu_calls_fail_loudly(lib):
    import torch
    from isa_b200 import _lib
    from isa_b200.losses import DiscriminativeLoss
    assert lib.isa_version() >= 100
    if not torch.cuda.is_available():
        with pytest.raises(_lib.IsaError):
            DiscriminativeLoss(0.5, 1.5, 2)(torch.zeros(1, 4, 4, 4), torch.zeros(1, 2, 4, 4), torch.tensor([1]), 2)
    with pytest.raises(_lib.IsaError):
        DiscriminativeLoss(0.5, 1.5, 2, usegpu=False)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "instance-segmentation-attention_b200")
    bad = []
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                s = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", s, re.M) or "/root/reference" in s.replace("/root/reference/code/", "REFCITE"):
                    bad.append(os.path.join(dp, f))
    assert not bad, "product files touching oracle/ or the reference tree: %s" % bad
