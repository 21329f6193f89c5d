"""shared column bookkeeping of the two buffer classes"""
import ctypes as C

import numpy as np

from .. import _carrier as K
from .. import _ffi


class ColumnSpec:
    def __init__(self, shapes, types):
        if len(shapes) != len(types):
            raise ValueError("shapes and types must have the same length")
        if not 0 < len(shapes) <= _ffi.MAX_COLS:
            raise ValueError(f"between 1 and {_ffi.MAX_COLS} columns are supported")
        self.shapes = [tuple(int(d) for d in (s if isinstance(s, (tuple, list)) else (s,)))
                       for s in shapes]
        self.dtypes = [K.np_dtype(t) for t in types]
        self.row_bytes = [int(np.prod(s, dtype=np.int64)) * d.itemsize
                          for s, d in zip(self.shapes, self.dtypes)]
        self.ncols = len(self.shapes)
        self.row_bytes_arr = np.asarray(self.row_bytes, dtype=np.int64)

    def rows_in(self, cols):
        """number of rows carried by a list of per-column arrays (>=1; 'Support batch input',
        README.md:428): leading dims beyond the column shape are the batch."""
        if len(cols) != self.ncols:
            raise ValueError(f"expected {self.ncols} columns, got {len(cols)}")
        n = None
        for c, shape in zip(cols, self.shapes):
            if K.is_torch(c):
                numel = c.numel()
            elif isinstance(c, K.DeviceArray):
                numel = int(np.prod(c.shape, dtype=np.int64))
            else:
                numel = np.asarray(c).size
            per_row = int(np.prod(shape, dtype=np.int64))
            if numel % per_row:
                raise ValueError("column size is not a multiple of its row size")
            k = numel // per_row
            if n is None:
                n = k
            elif k != n:
                raise ValueError("columns disagree on the number of rows")
        if not n:
            raise ValueError("empty insert")
        return n

    def prepare_inputs(self, cols):
        """-> (kind, arrays, pointer table); kind 'device' only if every column is on the device"""
        kinds = {K.classify(c) for c in cols}
        if kinds == {"device"}:
            arrs = [K.as_device(c, d) for c, d in zip(cols, self.dtypes)]
            return "device", arrs, _ffi.ptr_table([a.data_ptr() for a in arrs])
        arrs = [K.as_host(c, d) for c, d in zip(cols, self.dtypes)]
        return "host", arrs, _ffi.ptr_table([a.ctypes.data for a in arrs])

    def outputs(self, n, kind):
        lead = () if n is None else (n,)
        if kind == "device":
            arrs = [K.device_empty(lead + s, d) for s, d in zip(self.shapes, self.dtypes)]
            return arrs, _ffi.ptr_table([a.data_ptr() for a in arrs])
        arrs = [np.empty(lead + s, dtype=d) for s, d in zip(self.shapes, self.dtypes)]
        return arrs, _ffi.ptr_table([a.ctypes.data for a in arrs])


def default_carrier(carrier):
    if carrier in (None, "auto"):
        return "device"
    if carrier in ("torch", "device"):
        return "device"
    if carrier in ("numpy", "host"):
        return "host"
    raise ValueError(f"unknown carrier {carrier!r}")
