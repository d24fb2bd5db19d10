"""CPU: the C-ABI library builds/loads here and exports every symbol include/ynet_b200.h declares."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, 'include', 'ynet_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ynet_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported():
    from motion_style_transfer_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/ynet_b200.h but not exported'
    # and the ctypes prototypes cover the header exactly
    assert sorted(_lib.exported_names()) == names
    assert lib.ynet_version() >= 100


def test_no_cpu_fallback():
    import torch
    from motion_style_transfer_b200 import ops
    with pytest.raises(RuntimeError, match='CUDA only'):
        ops.softargmax2d(torch.zeros(1, 1, 4, 4))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'motion_style_transfer_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f'{f} imports oracle/'
