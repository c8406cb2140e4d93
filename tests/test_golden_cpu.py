"""The committed golden fixtures are reproduced by the oracle (guards both against drift)."""
import pathlib
import sys

import numpy as np

GOLD = pathlib.Path(__file__).resolve().parent / 'golden'
sys.path.insert(0, str(GOLD))
import make_golden  # noqa: E402


def _check(name, fn):
    ref = np.load(GOLD / f'{name}.npz')
    new = fn()
    assert sorted(ref.files) == sorted(new)
    for k in ref.files:
        np.testing.assert_allclose(new[k], ref[k], rtol=1e-12, atol=1e-14, err_msg=f'{name}:{k}')


def test_c1():
    _check('c1_expquad', make_golden.c1)


def test_c2():
    _check('c2_matern', make_golden.c2)


def test_c4():
    _check('c4_bart', make_golden.c4)
