"""Neighbourhood construction and nearest-point search on the GPU (extension).

The reference leaves these steps to the caller and to ``scipy.spatial.cKDTree`` on the host:

    tree  = scipy.spatial.cKDTree(x)
    hoods = tree.query(x, k + 1)[1][:, 1:]           # examples/expertsolver_example.py:59-66
    xk, fk = x[hoods], f[hoods]

At a million points the kd-tree query costs seconds while the fit costs milliseconds, so the same searches are
offered on the device (uniform grid, exact k nearest neighbours ordered by (distance, index); identical to
cKDTree wherever distances are distinct):

    grid  = wlsqm_b200.PointGrid(x)                  # x: numpy array or CUDA tensor, (n, dim) or (n,)
    hoods = grid.knn(k)                              # (n, k) int32, self excluded
    d, I  = grid.query(xq, k=1)                      # like cKDTree.query

``ExpertSolver.prepare_hoods(x, hoods)`` / ``solve_hoods(f, fi)`` take the index lists directly.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

__all__ = ["PointGrid", "knn_hoods", "gather"]


def _points(x, name="x"):
    """(n, dim) or (n,) float64, numpy or CUDA tensor -> (Arr, n, dim, row stride)"""
    nd = x.dim() if _lib._is_torch_tensor(x) else np.ndim(x)
    if nd == 1:
        a = _lib.as_arr(x, np.float64, 1, name, last_contig=False)
        return a, a.shape[0], 1, (a.strides[0] if a.shape[0] > 1 else 1)
    a = _lib.as_arr(x, np.float64, 2, name)
    if a.shape[1] not in (1, 2, 3):
        raise ValueError(f"{name}: points must have 1, 2 or 3 coordinates, got {a.shape[1]}")
    return a, a.shape[0], a.shape[1], a.strides[0]


class PointGrid:
    """Uniform search grid over a point cloud, resident on one GPU."""

    def __init__(self, x, device=None):
        a, n, dim, s0 = _points(x)
        if device is None:
            device = a.device if a.is_cuda else _lib.default_device()
        self.device, self.n, self.dimension = int(device), n, dim
        self._on_device = a.is_cuda
        self._torch_device = x.device if a.is_cuda else None
        h = C.c_void_p()
        _lib.announce_stream(self.device)
        _lib.check(_lib.lib().wlsqm_grid_create(dim, n, a.ptr, s0, self.device, C.byref(h)))
        self._handle = h

    def close(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None and h.value:
            _lib.lib().wlsqm_grid_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        """(number of cells, cell edge length, device bytes)"""
        nc, hh, nb = C.c_int64(), C.c_double(), C.c_int64()
        _lib.check(_lib.lib().wlsqm_grid_info(self._handle, C.byref(nc), C.byref(hh), C.byref(nb)))
        return nc.value, hh.value, nb.value

    def _out(self, shape, np_dtype, on_device):
        if on_device:
            import torch
            t = torch.empty(shape, dtype=getattr(torch, np.dtype(np_dtype).name), device=self._torch_device or f"cuda:{self.device}")
            return t, int(t.data_ptr())
        a = np.empty(shape, dtype=np_dtype)
        return a, a.ctypes.data

    def knn(self, k, exclude_self=True, return_distance=False):
        """The k nearest neighbours of every point of the cloud: (n, k) int32, nearest first; with
        ``exclude_self`` the point itself is left out (``tree.query(x, k+1)[1][:, 1:]``)."""
        k = int(k)
        idx, ip = self._out((self.n, k), np.int32, self._on_device)
        d2, dp = (self._out((self.n, k), np.float64, self._on_device) if return_distance else (None, None))
        _lib.announce_stream(self.device)
        _lib.check(_lib.lib().wlsqm_grid_knn(self._handle, None, 0, self.n, k, 1 if exclude_self else 0, ip, None, dp))
        if return_distance:
            return (d2.sqrt() if self._on_device else np.sqrt(d2)), idx
        return idx

    def query(self, xq, k=1):
        """``cKDTree.query(xq, k)``: (distances, indices); indices are int64 (``np.int_``), missing neighbours n."""
        a, nq, dim, s0 = _points(xq, "xq")
        if dim != self.dimension:
            raise ValueError("xq has %d coordinates, the grid has %d" % (dim, self.dimension))
        k = int(k)
        idx, ip = self._out((nq, k), np.int64, a.is_cuda)
        d2, dp = self._out((nq, k), np.float64, a.is_cuda)
        if nq:
            _lib.announce_stream(self.device)
            _lib.check(_lib.lib().wlsqm_grid_knn(self._handle, a.ptr, s0, nq, k, 0, None, ip, dp))
        d = d2.sqrt() if a.is_cuda else np.sqrt(d2)
        if k == 1:
            return d[:, 0], idx[:, 0]
        return d, idx


def knn_hoods(x, k, device=None):
    """hoods = cKDTree(x).query(x, k+1)[1][:, 1:] on the GPU: (n, k) int32."""
    g = PointGrid(x, device=device)
    try:
        return g.knn(k)
    finally:
        g.close()


def gather(src, hoods):
    """``src[hoods]`` on the device: src (npoints,) or (npoints, w) float64 CUDA tensor, hoods (n, k) int32 CUDA
    tensor -> (n, k) or (n, k, w)."""
    import torch
    if not (_lib._is_torch_tensor(src) and src.is_cuda and _lib._is_torch_tensor(hoods) and hoods.is_cuda):
        raise ValueError("gather() works on CUDA tensors (numpy arrays: use src[hoods])")
    if hoods.dtype != torch.int32 or hoods.dim() != 2 or hoods.stride(1) != 1:
        raise ValueError("hoods must be a (n, k) int32 tensor with contiguous rows")
    if src.dtype != torch.float64:
        raise ValueError("src must be float64")
    w = 1 if src.dim() == 1 else src.shape[1]
    if src.dim() == 2 and src.stride(1) != 1:
        src = src.contiguous()
    n, k = hoods.shape
    out = torch.empty((n, k) if src.dim() == 1 else (n, k, w), dtype=torch.float64, device=src.device)
    sp = _lib.current_stream_ptr(src.device.index)
    _lib.check(_lib.lib().wlsqm_gather_hoods(int(src.data_ptr()), src.stride(0), w, src.shape[0], int(hoods.data_ptr()),
                                             hoods.stride(0), n, k, int(out.data_ptr()), src.device.index, sp))
    return out
