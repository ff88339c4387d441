"""CPU: the C ABI library loads and exports every symbol include/bh_b200.h declares; host-only entry points."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(pkg):
    hdr = open(os.path.join(ROOT, "include", "bh_b200.h")).read()
    declared = set(re.findall(r"\b(bh_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("bh_ctx")
    assert len(declared) >= 25
    L = pkg.capi.load()
    for s in sorted(declared):
        assert hasattr(L, s), f"{s} is declared in include/bh_b200.h but not exported"
    assert declared == set(pkg.capi.SYMBOLS), declared ^ set(pkg.capi.SYMBOLS)


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, "include", "bh_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)   # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code and 'extern "C"' in code


def test_host_side_entry_points(pkg):
    c = pkg.capi
    for m, n in [(8, 8), (10, 10), (12, 12), (14, 14), (5, 9)]:
        assert c.dimension(m, n) == O.dimension(m, n)
    assert c.dimension(14, 14) == 20058300
    for m in (2, 3, 5, 12):
        for closed in (True, False):
            p, i = c.neighbours_chain(m, closed)
            op, oi = O.chain(m, closed)
            assert (p == op).all() and (i == oi).all()
    p, i = c.neighbours_rect(4, 3)
    op, oi = O.rect(4, 3)
    assert (p == op).all() and (i == oi).all()
    p, i = c.neighbours_rect(3, 3, 3)
    assert p[-1] == 27 * 6
    e = np.array([0.3, -1.0, 2.0, 2.0, 5.5, 0.31])
    assert np.allclose(c.gap_ratios(e), O.gap_ratios(e))
    rng = np.random.default_rng(0)
    a = rng.normal(size=(6, 6))
    rho = a @ a.T
    assert abs(c.condensate_fraction(rho) - O.condensate_fraction(rho)) < 1e-12
    assert abs(c.coherence(rho) - O.coherence(rho)) < 1e-14
    assert abs(c.condensate_fraction(rho) - np.abs(np.linalg.eigvalsh(rho)).max() / np.trace(rho)) < 1e-12


def test_fails_loudly_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.BhError) as ei:
        pkg.Context(0)
    assert "no CPU path" in str(ei.value)


def test_product_does_not_import_oracle():
    # the product path must never route through oracle/ (checked textually over the package sources)
    pk = os.path.join(ROOT, "bose-hubbard-phase-transition_b200")
    for dp, _, files in os.walk(pk):
        if "build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                for token in ("oracle_lib", "liboracle", "bh_oracle", "bho_", "import oracle", "ref_lib", "_ref/"):
                    assert token not in txt, (os.path.join(dp, f), token)
