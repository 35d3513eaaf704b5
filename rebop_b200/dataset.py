"""Result container.

The reference returns an `xarray.Dataset` with a `time` dimension
(python/rebop/gillespie.py:153-160).  xarray is used when it is importable; otherwise a small
stand-in offers the access patterns the reference's tests and examples rely on:
``ds.S``, ``ds["S"]``, ``ds.time``, ``"S" in ds``, iteration over variable names,
``ds.data_vars`` and ``ds.sizes``.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    import xarray as _xr
except Exception:  # noqa: BLE001
    _xr = None


class Dataset:
    """Minimal stand-in for xarray.Dataset: named integer arrays sharing a `time` coordinate."""

    def __init__(self, data_vars, time, dims):
        self.data_vars = dict(data_vars)
        self.time = np.asarray(time)
        self.dims = tuple(dims)
        self.coords = {"time": self.time}
        first = next(iter(self.data_vars.values()), None)
        self.sizes = {"time": len(self.time)}
        if "trajectory" in self.dims and first is not None:
            self.sizes["trajectory"] = first.shape[1]

    def __getitem__(self, key):
        if isinstance(key, (list, tuple)):
            return Dataset({k: self.data_vars[k] for k in key}, self.time, self.dims)
        return self.data_vars[key]

    def __getattr__(self, name):
        try:
            return self.__dict__["data_vars"][name]
        except KeyError:
            raise AttributeError(name) from None

    def __contains__(self, key):
        return key in self.data_vars

    def __iter__(self):
        return iter(self.data_vars)

    def __len__(self):
        return len(self.data_vars)

    def keys(self):
        return self.data_vars.keys()

    def __eq__(self, other):
        return (isinstance(other, Dataset) and set(self.data_vars) == set(other.data_vars)
                and np.array_equal(self.time, other.time)
                and all(np.array_equal(v, other.data_vars[k]) for k, v in self.data_vars.items()))

    def mean(self, dim="trajectory"):
        axis = self.dims.index(dim)
        return {k: v.mean(axis=axis) for k, v in self.data_vars.items()}

    def __repr__(self):
        shape = ", ".join(f"{k}: {v}" for k, v in self.sizes.items())
        return f"<rebop_b200.Dataset ({shape}) vars: {', '.join(self.data_vars)}>"


def make_dataset(values, times, batched):
    dims = ("time", "trajectory") if batched else ("time",)
    if _xr is not None:
        coords = {"time": times}
        return _xr.Dataset(
            data_vars={name: _xr.DataArray(v, dims=dims, coords=coords) for name, v in values.items()},
            coords=coords,
        )
    return Dataset(values, times, dims)
