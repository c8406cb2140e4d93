"""Columnar structured arrays: the input-layout contract of the kernels.

Mirrors the part of lsqfitgp's StructuredArray that the GP-fitting path relies on
(reference: src/lsqfitgp/_array.py:29-411 StructuredArray, :488-504 unstructured_to_structured,
:555-564 _nd): one array per field, every leaf field (and every element of a shaped field) is
one covariate dimension.  On the device the points are stored field-major, x[d, i], which is what
the Gram kernels read (coalesced along i).
"""

import numpy

__all__ = ['StructuredArray', 'unstructured_to_structured', 'asarray']


class StructuredArray:
    """Read-only columnar version of a numpy structured array.

    Parameters
    ----------
    array : numpy structured array, StructuredArray, or dict name -> array
    """

    def __init__(self, array):
        if isinstance(array, StructuredArray):
            self._fields = dict(array._fields)
            self.shape = array.shape
            self.dtype = array.dtype
            return
        if isinstance(array, dict):
            self._init_from_dict(array)
            return
        array = numpy.asarray(array)
        if array.dtype.names is None:
            raise ValueError('StructuredArray needs a structured dtype; use unstructured_to_structured')
        self._fields = {}
        for name in array.dtype.names:
            col = array[name]
            if col.dtype.names is not None:
                col = StructuredArray(col)
            self._fields[name] = col
        self.shape = array.shape
        self.dtype = array.dtype

    def _init_from_dict(self, d):
        if not d:
            raise ValueError('empty dictionary')
        arrays = {k: (v if isinstance(v, StructuredArray) else numpy.asarray(v)) for k, v in d.items()}
        shape = None
        for k, v in arrays.items():
            s = v.shape
            shape = s if shape is None else _common_prefix(shape, s)
        descr = []
        for k, v in arrays.items():
            sub = v.shape[len(shape):]
            descr.append((k, v.dtype, sub) if sub else (k, v.dtype))
        self._fields = arrays
        self.shape = shape
        self.dtype = numpy.dtype(descr)

    @classmethod
    def from_dict(cls, mapping):
        return cls(dict(mapping))

    @classmethod
    def from_dataframe(cls, df):
        """ pandas / polars DataFrame -> 1-d StructuredArray, one field per column """
        return cls({str(c): numpy.asarray(df[c]) for c in df.columns})

    @property
    def size(self):
        return int(numpy.prod(self.shape, dtype=int))

    @property
    def ndim(self):
        return len(self.shape)

    def __len__(self):
        if not self.shape:
            raise TypeError('len() of unsized object')
        return self.shape[0]

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._fields[key]
        if isinstance(key, list) and key and all(isinstance(k, str) for k in key):
            return StructuredArray({k: self._fields[k] for k in key})
        new = object.__new__(StructuredArray)
        new._fields = {}
        probe = numpy.empty(self.shape, dtype=bool)[key]
        for name, col in self._fields.items():
            if isinstance(col, StructuredArray):
                new._fields[name] = col[key]
            else:
                sub = col.ndim - len(self.shape)
                idx = key if isinstance(key, tuple) else (key,)
                new._fields[name] = col[idx + (Ellipsis,)] if sub else col[key]
        new.shape = probe.shape
        new.dtype = self.dtype
        return new

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        probe = numpy.empty(self.shape, dtype=bool).reshape(shape)
        new = object.__new__(StructuredArray)
        new._fields = {}
        for name, col in self._fields.items():
            if isinstance(col, StructuredArray):
                new._fields[name] = col.reshape(probe.shape)
            else:
                sub = col.shape[len(self.shape):]
                new._fields[name] = col.reshape(probe.shape + sub)
        new.shape = probe.shape
        new.dtype = self.dtype
        return new

    @property
    def T(self):
        if self.ndim < 2:
            return self
        raise NotImplementedError('transpose of n-d StructuredArray')

    def leaf_columns(self):
        """ list of (label, flat 1-d float array) for every covariate dimension, in dtype order """
        out = []
        n = self.size
        for name, col in self._fields.items():
            if isinstance(col, StructuredArray):
                out += [((name,) + lab if isinstance(lab, tuple) else (name, lab), c) for lab, c in col.leaf_columns()]
            else:
                flat = numpy.asarray(col).reshape(n, -1)
                if flat.shape[1] == 1 and col.ndim == len(self.shape):
                    out.append((name, flat[:, 0]))
                else:
                    out += [((name, j), flat[:, j]) for j in range(flat.shape[1])]
        return out

    def __repr__(self):
        return f'StructuredArray(shape={self.shape}, fields={list(self._fields)})'


def _common_prefix(a, b):
    out = []
    for x, y in zip(a, b):
        if x != y:
            break
        out.append(x)
    return tuple(out)


def unstructured_to_structured(arr, dtype=None, names=None, **_):
    """ like numpy.lib.recfunctions.unstructured_to_structured: last axis -> fields """
    arr = numpy.asarray(arr)
    nf = arr.shape[-1]
    if dtype is not None:
        names = numpy.dtype(dtype).names
    if names is None:
        names = [f'f{i}' for i in range(nf)]
    if len(names) != nf:
        raise ValueError('number of names does not match the last axis')
    return StructuredArray({n: arr[..., i] for i, n in enumerate(names)})


def asarray(x):
    """ numpy array, or StructuredArray for structured inputs """
    if isinstance(x, StructuredArray):
        return x
    if isinstance(x, dict):
        return StructuredArray(x)
    try:
        import pandas
        if isinstance(x, pandas.DataFrame):
            return StructuredArray.from_dataframe(x)
    except ImportError:  # pragma: no cover
        pass
    x = numpy.asarray(x)
    if x.dtype.names is not None:
        return StructuredArray(x)
    return x


def columns_of(x):
    """ (labels, (ndim, n) float64 numpy array, shape) of an array-like of points """
    x = asarray(x)
    if isinstance(x, StructuredArray):
        cols = x.leaf_columns()
        labels = [c[0] for c in cols]
        data = numpy.stack([numpy.asarray(c[1], dtype=numpy.float64) for c in cols]) if cols else numpy.empty((0, x.size))
        return labels, data, x.shape
    if not (numpy.issubdtype(x.dtype, numpy.number) or x.dtype == bool):
        raise TypeError(f'points have non-numerical dtype {x.dtype!r}')
    return [None], x.reshape(1, -1).astype(numpy.float64), x.shape
