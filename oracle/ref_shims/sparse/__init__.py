"""TEST INFRASTRUCTURE ONLY -- stand-in for the pydata `sparse` package.

The reference (ayosprakob/grassmanntn) imports `sparse` (COO arrays) but the
package is not installed in the build container and there is no network.  This
module gives the reference just enough of `sparse.COO` to run: plain
(coords, data) attributes like the real class, with arithmetic done densely.
It is used ONLY by oracle/ref_harness.py when golden fixtures are generated from
the real reference; nothing in grassmanntn_b200/ imports it.
"""
import numpy as np


class COO:
    __array_priority__ = 100

    def __init__(self, coords, data=None, shape=None):
        if data is None:
            arr = np.array(coords)
            self._from_dense(arr)
            return
        coords = np.asarray(coords).reshape(len(shape), -1).astype(np.int64)
        data = np.asarray(data)
        arr = np.zeros(tuple(shape), dtype=data.dtype)
        if data.size:
            np.add.at(arr, tuple(coords[a] for a in range(coords.shape[0])), data)
        self._from_dense(arr)

    def _from_dense(self, arr):
        arr = np.asarray(arr)
        self.shape = tuple(arr.shape)
        nz = np.nonzero(arr)
        self.coords = np.array(nz, dtype=np.int64).reshape(arr.ndim, -1)
        self.data = np.array(arr[nz])
        self._dtype = arr.dtype

    @classmethod
    def from_numpy(cls, arr):
        out = cls.__new__(cls)
        out._from_dense(np.asarray(arr))
        return out

    def todense(self):
        arr = np.zeros(self.shape, dtype=np.result_type(self._dtype, np.asarray(self.data).dtype))
        if np.asarray(self.data).size:
            c = np.asarray(self.coords)
            np.add.at(arr, tuple(c[a] for a in range(c.shape[0])), np.asarray(self.data))
        return arr

    def copy(self):
        out = COO.__new__(COO)
        out.shape = self.shape
        out.coords = np.array(self.coords)
        out.data = np.array(self.data)
        out._dtype = self._dtype
        return out

    @property
    def nnz(self):
        return int(np.asarray(self.data).shape[0])

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64)) if len(self.shape) else 1

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def dtype(self):
        return np.asarray(self.data).dtype

    def reshape(self, shape):
        return COO.from_numpy(self.todense().reshape(shape))

    def take(self, indices, axis=None):
        return COO.from_numpy(self.todense().take(indices, axis=axis))

    def __array_function__(self, func, types, args, kwargs):
        # numpy functions (np.reshape, np.conjugate, ...) applied to a COO return a COO,
        # as in the real package
        def conv(x):
            if isinstance(x, COO):
                return x.todense()
            if isinstance(x, (list, tuple)):
                return type(x)(conv(y) for y in x)
            return x
        res = func(*conv(args), **{k: conv(v) for k, v in kwargs.items()})
        if isinstance(res, np.ndarray) and res.ndim > 0:
            return COO.from_numpy(res)
        return res

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        ins = [x.todense() if isinstance(x, COO) else x for x in inputs]
        res = getattr(ufunc, method)(*ins, **kwargs)
        if isinstance(res, np.ndarray) and res.ndim > 0:
            return COO.from_numpy(res)
        return res

    def __array__(self, dtype=None, copy=None):
        d = self.todense()
        return d if dtype is None else d.astype(dtype)

    @staticmethod
    def _val(x):
        return x.todense() if isinstance(x, COO) else x

    def __add__(self, o):
        return COO.from_numpy(self.todense() + COO._val(o))

    __radd__ = __add__

    def __sub__(self, o):
        return COO.from_numpy(self.todense() - COO._val(o))

    def __mul__(self, o):
        if np.isscalar(o):  # keep explicit entries, like the real package
            out = self.copy()
            out.data = out.data * o
            return out
        return COO.from_numpy(self.todense() * COO._val(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        if np.isscalar(o):
            out = self.copy()
            out.data = out.data / o
            return out
        return COO.from_numpy(self.todense() / COO._val(o))

    def __neg__(self):
        return self * (-1)

    def __repr__(self):
        return "<COO-shim shape=%s nnz=%d>" % (self.shape, self.nnz)
