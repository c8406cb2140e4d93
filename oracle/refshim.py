"""TEST INFRASTRUCTURE (never imported by the product): run the REFERENCE'S OWN SOURCE FILES without JAX.

`import lsqfitgp` fails in this image (jax, jaxlib and gvar are not installable).  The arithmetic
of the hot path, however, is plain array code written against `jax.numpy` / `jax.scipy`; with float64 enabled those
functions are drop-in equivalents of numpy / scipy.  This module installs a minimal stand-in for the `jax` package
(numpy + scipy behind the jax names, `.at[...]` functional updates, decorators as no-ops) and loads individual modules of
`/root/reference/src/lsqfitgp` under it, so that golden vectors can be produced by executing the reference's code
itself (tests/golden/gen_reference_vectors.py) instead of a restatement of it.

What this is NOT: the JAX runtime.  XLA's own rounding (fusions, reduction order) is not reproduced; LAPACK comes from
scipy (the routine family jaxlib's CPU backend calls).  Only eager evaluation is supported: `jit`, `custom_jvp`,
`ensure_compile_time_eval` are pass-throughs and the autodiff transforms raise.

Only usable where /root/reference exists (this container); nothing on the GPU box imports it.
"""

import contextlib
import functools
import importlib
import importlib.util
import pathlib
import sys
import types

import numpy
import scipy.linalg
import scipy.special

REF_SRC = pathlib.Path('/root/reference/src/lsqfitgp')


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self._arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def _apply(self, fn):
        out = numpy.array(self._arr, copy=True)
        fn(out)
        return out.view(Array)

    def _inbounds(self):
        """ jax drops out-of-bounds scatter updates; numpy raises: filter plain integer indices along axis 0 """
        idx = self._idx
        if isinstance(idx, tuple) or isinstance(idx, slice) or idx is Ellipsis:
            return idx, None
        ia = numpy.asarray(idx)
        if ia.dtype.kind not in 'iu':
            return idx, None
        ok = (ia >= -self._arr.shape[0]) & (ia < self._arr.shape[0])
        return ia, ok

    def set(self, v):
        ia, ok = self._inbounds()
        if ok is None:
            return self._apply(lambda a: a.__setitem__(self._idx, v))
        if ia.ndim == 0:
            return self._apply(lambda a: a.__setitem__(int(ia), v)) if bool(ok) else numpy.array(self._arr).view(Array)
        vv = numpy.broadcast_to(numpy.asarray(v), ia.shape + self._arr.shape[1:])
        return self._apply(lambda a: a.__setitem__(ia[ok], vv[ok]))

    def add(self, v):
        return self._apply(lambda a: numpy.add.at(a, self._idx, v))

    def multiply(self, v):
        return self._apply(lambda a: numpy.multiply.at(a, self._idx, v))

    def get(self):
        return self._arr[self._idx]


class Array(numpy.ndarray):
    """ numpy array with jax's functional-update accessor """

    @property
    def at(self):
        return _At(self)


def _wrap(x):
    if isinstance(x, numpy.ndarray) and not isinstance(x, Array):
        return x.view(Array)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _wrapfun(f):
    @functools.wraps(f)
    def g(*a, **k):
        return _wrap(f(*a, **k))
    return g


class _NumpyProxy(types.ModuleType):
    """ jax.numpy: numpy functions returning `Array` """

    def __init__(self, name, backing):
        super().__init__(name)
        self._backing = backing

    def __getattr__(self, name):
        v = getattr(self._backing, name)
        if callable(v) and not isinstance(v, type):
            v = _wrapfun(v)
        setattr(self, name, v)
        return v


def _passthrough_decorator(*dargs, **dkw):
    """ jax.jit & co.: usable as @jit, @jit(static_argnums=...) or functools.partial(jit, ...)(f) """
    if len(dargs) == 1 and callable(dargs[0]) and not dkw:
        return dargs[0]
    if dargs and callable(dargs[0]):
        return dargs[0]
    return lambda f: f


class _CustomJVP:
    def __init__(self, fun, nondiff_argnums=()):
        self.fun = fun
        functools.update_wrapper(self, fun)

    def __call__(self, *a, **k):
        return self.fun(*a, **k)

    def defjvp(self, f, **kw):
        return f

    defjvps = defjvp


def _not_available(name):
    def f(*a, **k):
        raise NotImplementedError(f'jax.{name} is not available under oracle/refshim.py (eager numpy stand-in)')
    return f


def _cholesky(a, lower=False, **kw):
    """ jax.scipy.linalg.cholesky: NaNs instead of an exception when the matrix is not positive definite """
    a = numpy.asarray(a)
    try:
        return scipy.linalg.cholesky(a, lower=lower, check_finite=False)
    except scipy.linalg.LinAlgError:
        return numpy.full_like(a, numpy.nan, dtype=float)


def _solve_triangular(a, b, trans=0, lower=False, unit_diagonal=False, **kw):
    return scipy.linalg.solve_triangular(numpy.asarray(a), numpy.asarray(b), trans=trans, lower=lower,
                                         unit_diagonal=unit_diagonal, check_finite=False)


def _lax_triangular_solve(a, b, *, left_side=False, lower=False, transpose_a=False, conjugate_a=False,
                          unit_diagonal=False):
    a, b = numpy.asarray(a), numpy.asarray(b)
    if left_side:
        return scipy.linalg.solve_triangular(a, b, lower=lower, trans=1 if transpose_a else 0,
                                             unit_diagonal=unit_diagonal, check_finite=False)
    # x a = b  <=>  a^T x^T = b^T
    return scipy.linalg.solve_triangular(a, b.T, lower=lower, trans=0 if transpose_a else 1,
                                         unit_diagonal=unit_diagonal, check_finite=False).T


def install():
    """ put the stand-in `jax` package into sys.modules (idempotent); returns the module """
    if 'jax' in sys.modules and getattr(sys.modules['jax'], '_lgp_refshim', False):
        return sys.modules['jax']
    if 'jax' in sys.modules:
        raise RuntimeError('a real jax is importable: use it instead of the stand-in')
    jax = types.ModuleType('jax')
    jax._lgp_refshim = True
    jnp = _NumpyProxy('jax.numpy', numpy)
    jnp.ndarray = Array
    jnp.asarray = lambda x, dtype=None, **k: numpy.asarray(x, dtype=dtype).view(Array)
    jnp.array = lambda x, dtype=None, **k: numpy.array(x, dtype=dtype).view(Array)
    jnp.vectorize = lambda pyfunc=None, **kw: (lambda f: _wrapfun(numpy.vectorize(f, **kw))) if pyfunc is None \
        else _wrapfun(numpy.vectorize(pyfunc, **kw))
    def _unique(x, size=None, fill_value=None, **kw):
        u = numpy.unique(numpy.asarray(x), **kw)
        if size is None or isinstance(u, tuple):
            return _wrap(u)
        out = numpy.full(size, u[0] if fill_value is None else fill_value, dtype=u.dtype)
        out[:min(size, u.size)] = u[:size]
        return out.view(Array)
    jnp.unique = _unique
    jnp.linalg = _NumpyProxy('jax.numpy.linalg', numpy.linalg)
    jax.numpy = jnp
    jax.Array = Array
    jax.jit = _passthrough_decorator
    jax.custom_jvp = lambda fun=None, nondiff_argnums=(): _CustomJVP(fun, nondiff_argnums) if fun is not None \
        else (lambda f: _CustomJVP(f, nondiff_argnums))
    jax.custom_vjp = jax.custom_jvp
    jax.ensure_compile_time_eval = contextlib.nullcontext
    for name in ('vjp', 'jvp', 'jacfwd', 'jacrev', 'grad', 'value_and_grad', 'linearize'):
        setattr(jax, name, _not_available(name))

    def vmap(f, in_axes=0, out_axes=0):
        """ jax.vmap by an explicit loop over the mapped axis (array arguments and a single array output) """
        def g(*args):
            axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
            n = next(numpy.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
            outs = [f(*[a if ax is None else numpy.take(numpy.asarray(a), i, axis=ax) for a, ax in zip(args, axes)])
                    for i in range(n)]
            return numpy.stack([numpy.asarray(o) for o in outs], axis=out_axes).view(Array)
        return g
    jax.vmap = vmap
    jax.pure_callback = lambda callback, result_shape, *args, **kw: _wrap(callback(*args))
    jax.ShapeDtypeStruct = lambda shape, dtype: types.SimpleNamespace(shape=shape, dtype=dtype)
    def eval_shape(f, *a, **k):
        """ abstract evaluation: inputs may be shape/dtype mock-ups; only .shape / .dtype of the result are used """
        conc = [numpy.zeros(x.shape, x.dtype).view(Array)
                if (hasattr(x, 'shape') and hasattr(x, 'dtype') and not isinstance(x, numpy.ndarray)) else x for x in a]
        return _wrap(f(*conc, **k))
    jax.eval_shape = eval_shape
    errors = types.ModuleType('jax.errors')

    class ConcretizationTypeError(Exception):
        pass

    class TracerArrayConversionError(Exception):
        pass
    errors.ConcretizationTypeError = ConcretizationTypeError
    errors.TracerArrayConversionError = TracerArrayConversionError
    jax.errors = errors
    core = types.ModuleType('jax.core')
    core.Tracer = type('Tracer', (), {})
    jax.core = core
    config = types.SimpleNamespace(update=lambda *a, **k: None, jax_enable_x64=True)
    jax.config = config
    tree_util = types.ModuleType('jax.tree_util')
    registry = set()

    def register_pytree_node_class(cls):
        registry.add(cls)
        return cls

    def tree_map(f, tree, *rest):
        """ jax.tree_util.tree_map for registered classes (tree_flatten / tree_unflatten), dict, list, tuple, None """
        if type(tree) in registry:
            children, aux = tree.tree_flatten()
            others = [r.tree_flatten()[0] for r in rest]
            new = [tree_map(f, c, *[o[i] for o in others]) for i, c in enumerate(children)]
            return type(tree).tree_unflatten(aux, new)
        if isinstance(tree, dict):
            return {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
        if isinstance(tree, (list, tuple)):
            return type(tree)(tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
        if tree is None:
            return None
        return f(tree, *rest)
    tree_util.register_pytree_node_class = register_pytree_node_class
    tree_util.tree_map = tree_map
    jax.tree_util = tree_util
    lax = types.ModuleType('jax.lax')
    lax.linalg = types.SimpleNamespace(triangular_solve=_wrapfun(_lax_triangular_solve),
                                       cholesky=_wrapfun(lambda a, **k: _cholesky(a, lower=True)))
    lax.stop_gradient = lambda x: x

    def _tree_index(tree, i):
        if isinstance(tree, (tuple, list)):
            return type(tree)(_tree_index(t, i) for t in tree)
        if tree is None:
            return None
        return tree[i]

    def _tree_len(tree):
        if isinstance(tree, (tuple, list)):
            for t in tree:
                n = _tree_len(t)
                if n is not None:
                    return n
            return None
        return None if tree is None else len(tree)

    def _tree_stack(items):
        first = items[0]
        if isinstance(first, (tuple, list)):
            return type(first)(_tree_stack([it[j] for it in items]) for j in range(len(first)))
        if first is None:
            return None
        return numpy.stack([numpy.asarray(it) for it in items]).view(Array)

    def scan(f, init, xs=None, length=None, reverse=False, unroll=1):
        n = _tree_len(xs) if xs is not None else length
        order = range(n - 1, -1, -1) if reverse else range(n)
        carry, ys = init, [None] * n
        for i in order:
            carry, y = f(carry, _tree_index(xs, i) if xs is not None else None)
            ys[i] = y
        return carry, (_tree_stack(ys) if n else None)

    def fori_loop(lower, upper, body, val, **kw):
        for i in range(int(lower), int(upper)):
            val = body(i, val)
        return val
    lax.scan = scan
    lax.fori_loop = fori_loop
    lax.cond = lambda pred, tf, ff, *ops: tf(*ops) if pred else ff(*ops)
    lax.select = _wrapfun(lambda c, a, b: numpy.where(c, a, b))
    jax.lax = lax
    jscipy = types.ModuleType('jax.scipy')
    jlinalg = types.ModuleType('jax.scipy.linalg')
    jlinalg.cholesky = _wrapfun(_cholesky)
    jlinalg.solve_triangular = _wrapfun(_solve_triangular)
    jlinalg.solve = _wrapfun(lambda a, b, assume_a='gen', **k: scipy.linalg.solve(numpy.asarray(a), numpy.asarray(b),
                                                                                 assume_a=assume_a, check_finite=False))
    jspecial = _NumpyProxy('jax.scipy.special', scipy.special)
    jscipy.linalg, jscipy.special = jlinalg, jspecial
    jax.scipy = jscipy
    jrandom = types.ModuleType('jax.random')
    jax.random = jrandom
    for name, mod in [('jax', jax), ('jax.numpy', jnp), ('jax.numpy.linalg', jnp.linalg), ('jax.errors', errors),
                      ('jax.core', core), ('jax.tree_util', tree_util), ('jax.lax', lax), ('jax.scipy', jscipy),
                      ('jax.scipy.linalg', jlinalg), ('jax.scipy.special', jspecial), ('jax.random', jrandom)]:
        sys.modules[name] = mod
    return jax


class _AnyModule(types.ModuleType):
    """ module whose unknown attributes are empty placeholder classes (enough for isinstance checks and imports) """

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        v = type(name, (), {})
        setattr(self, name, v)
        return v


def install_gvar_stub():
    """ `gvar` is only needed by the reference for inputs/outputs that carry uncertainties; the raw=True paths used for the
    golden vectors touch gvar.mean (identity on plain arrays) and nothing else.  Everything else is a placeholder. """
    if 'gvar' in sys.modules:
        return sys.modules['gvar']
    g = _AnyModule('gvar')
    g.mean = lambda y: numpy.asarray(y)
    g._lgp_refshim = True
    sys.modules['gvar'] = g
    return g


def load_reference(*submodules):
    """ import `lsqfitgp.<submodule>` from the reference tree under the stand-in jax, without running the package's
    own __init__ (which needs gvar): a bare namespace package named `lsqfitgp` holds the submodules.
    Example: decomp = load_reference('_linalg._decomp')  ->  the reference's _linalg/_decomp.py module """
    if not REF_SRC.exists():
        raise FileNotFoundError(f'{REF_SRC} not present (the reference tree exists only in the build container)')
    install()
    install_gvar_stub()
    pkg = sys.modules.get('lsqfitgp')
    if pkg is None:
        pkg = types.ModuleType('lsqfitgp')
        pkg.__path__ = [str(REF_SRC)]
        pkg.__package__ = 'lsqfitgp'
        sys.modules['lsqfitgp'] = pkg
    out = [importlib.import_module('lsqfitgp.' + name) for name in submodules]
    return out[0] if len(out) == 1 else out
