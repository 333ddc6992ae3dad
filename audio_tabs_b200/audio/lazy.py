"""Lazy array base: a front-end stage that is only computed (on the GPU) when its values are needed.

madmom's stages are eager ndarray subclasses; computing each one would force five round trips
through HBM.  Here every stage records its recipe, and materialising the *last* stage of an intact
chain launches the single fused kernel.  ``np.asarray(stage)`` / any numpy function / arithmetic
triggers materialisation, so downstream madmom code (NeuralNetworkEnsemble, np.hstack) sees a
plain C-contiguous ndarray.
"""
from __future__ import annotations

import numpy as np
from numpy.lib.mixins import NDArrayOperatorsMixin


class LazyArray(NDArrayOperatorsMixin):
    _cache = None
    _tensor = None      # device-side result (torch tensor), kept when available

    # subclasses implement ---------------------------------------------------------------------
    def _compute_tensor(self):
        """Run the kernels; return a torch CUDA tensor holding this stage's values."""
        raise NotImplementedError

    def _result_shape(self):
        raise NotImplementedError

    _result_dtype = np.dtype(np.float32)

    # materialisation ---------------------------------------------------------------------------
    def tensor(self):
        """Device-resident result (no device->host copy)."""
        if self._tensor is None:
            self._tensor = self._compute_tensor()
        return self._tensor

    def materialize(self):
        if self._cache is None:
            self._cache = np.ascontiguousarray(self.tensor().cpu().numpy())
        return self._cache

    def __array__(self, dtype=None, copy=None):
        arr = self.materialize()
        if dtype is not None and np.dtype(dtype) != arr.dtype:
            return arr.astype(dtype)
        return arr.copy() if copy else arr

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        inputs = tuple(np.asarray(x) if isinstance(x, LazyArray) else x for x in inputs)
        if "out" in kwargs:
            kwargs["out"] = tuple(np.asarray(x) if isinstance(x, LazyArray) else x for x in kwargs["out"])
        return getattr(ufunc, method)(*inputs, **kwargs)

    # ndarray-like surface ------------------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self._result_shape())

    @property
    def dtype(self):
        return self._result_dtype

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def T(self):
        return self.materialize().T

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, index):
        return self.materialize()[index]

    def __iter__(self):
        return iter(self.materialize())

    def __getattr__(self, name):
        # anything else an ndarray offers (sum, max, reshape, astype, ...) acts on the materialised values
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)
