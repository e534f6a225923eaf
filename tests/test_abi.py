"""The drop-in boundary: libpt_cuda.so loads without a GPU and exports every symbol include/pt_abi.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'pt_abi.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pt_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported(ptlib):
    lib = ctypes.CDLL(ptlib.library_path())
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_sizes():
    text = open(os.path.join(ROOT, 'include', 'pt_abi.h')).read()
    assert '16388' in text and '88 bytes' in text
    assert 7 + 1024 + 768 + 783 + 128 + 64 + 1323 == 4097


def test_no_cpu_fallback(ptlib):
    """Without a CUDA device the context refuses to exist (this test only asserts that on a GPU-less box)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(ptlib.PtError) as e:
        ptlib.Renderer()
    assert e.value.code == -5


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under pathtracer_b200/ or include/ may import, include or link it."""
    bad = []
    for base in ('pathtracer_b200', 'include', 'examples'):
        for dp, dn, fn in os.walk(os.path.join(ROOT, base)):
            if 'lib' in dp.split(os.sep):
                continue
            for f in fn:
                if f.endswith(('.py', '.cpp', '.cu', '.cuh', '.h', 'Makefile')):
                    t = open(os.path.join(dp, f), errors='replace').read()
                    if re.search(r'(import\s+oracle|from\s+oracle|oracle/|liboracle)', t) and 'TEST' not in f:
                        # comments that merely mention the oracle as the parity gate are fine
                        for line in t.split('\n'):
                            if re.search(r'(import\s+oracle|from\s+oracle|#include\s*"[^"]*oracle|liboracle)', line):
                                bad.append((f, line))
    assert not bad, bad
