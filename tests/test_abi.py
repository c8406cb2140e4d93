"""The C-ABI library loads on a CPU-only box and exports every symbol that include/lgp_b200.h declares;
the ctypes table in lsqfitgp_b200/_lib.py lists exactly those symbols.  No compute calls here."""
import ctypes
import pathlib
import re

import pytest

from lsqfitgp_b200 import _lib

ROOT = pathlib.Path(__file__).resolve().parent.parent
HEADER = ROOT / 'include' / 'lgp_b200.h'


def declared_functions():
    text = HEADER.read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(lgp_[a-z0-9_]+)\s*\(', text)) - {'lgp_factor'})


def test_library_built():
    assert _lib.LIB_PATH.exists(), 'run __graft_entry__.build() first'


def test_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f'{name} declared in lgp_b200.h but not exported'


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_load_and_version():
    lib = _lib.load()
    assert lib.lgp_abi_version() == 1
    assert b'sm_100a' in lib.lgp_build_info()
    assert lib.lgp_chol_npad(1) == 128 and lib.lgp_chol_npad(20000) == 20096
    assert lib.lgp_chol_aux_doubles(128) == 3 * 128 + 16 + 128 * 128


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import numpy as np
    import lsqfitgp_b200 as lgp
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        lgp.ExpQuad()(np.arange(3.)[:, None], np.arange(3.)[None, :])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        lgp._linalg.Chol(np.eye(3))


def test_product_does_not_import_oracle():
    """ the oracle is test infrastructure: nothing under lsqfitgp_b200/ may import or reference it """
    for f in (ROOT / 'lsqfitgp_b200').rglob('*.py'):
        src = f.read_text()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
    for f in (ROOT / 'lsqfitgp_b200' / 'csrc').glob('*.cu*'):
        assert 'oracle' not in f.read_text(), f


def test_xla_ffi_shim_binds_only_declared_entry_points():
    """ the (untested, jax-less) XLA-FFI shim forwards to C-ABI entry points: every lgp_* function it calls must be
    declared in the header and exported by the library, and the JAX-side module must name exactly the handlers the shim
    defines """
    from lsqfitgp_b200 import _lib
    shim = (ROOT / 'lsqfitgp_b200' / 'csrc' / 'xla_ffi_shim.cc').read_text()
    called = set(re.findall(r'\b(lgp_(?!xla_)\w+)\s*\(', shim))
    assert called and called <= set(_lib.SIGNATURES), called - set(_lib.SIGNATURES)
    handlers = set(re.findall(r'XLA_FFI_DEFINE_HANDLER_SYMBOL\((lgp_xla_\w+),', shim))
    py = (ROOT / 'lsqfitgp_b200' / '_jaxffi.py').read_text()
    named = set(re.findall(r"'(lgp_xla_\w+)'", py))
    assert handlers == named, handlers ^ named
