"""SART inversion on the device: host-side mirror of the reference's inversion entry points.

Same names, argument meaning and return values as ``cherab.tools.inversions.invert_sart`` /
``invert_constrained_sart`` (cherab/tools/inversions/sart.pyx:26-155, :161-302) and the same life cycle as the
reference's GPU solver ``SartOpencl`` (cherab/tools/inversions/opencl/sart_opencl.py:33-318): the matrices go to the
device once, then any number of measurement vectors are inverted against them.  Differences, all deliberate:

* arithmetic is float64 like the CPU reference (SartOpencl is float32 and normalises the measurements by their maximum,
  sart_opencl.py:240-241); ``value_dtype=np.float32`` only halves the stored matrix, the sums stay float64;
* the geometry matrix may be given as the CSR triplet the ray-transfer kernel produces, on the host or still on the device
  (``SartSolver.from_csr`` / ``from_device_csr``), besides the dense (N_d, N_s) array of the reference;
* a 2-D ``measurement_vector`` (frames, N_d) inverts all frames in one pass over the matrix per iteration.

There is no CPU fallback: everything runs in libcherab_b200.so (core_b200/csrc/cb2_sart.cu).
"""
import ctypes as C

import numpy as np

from . import _abi


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class SartSolver:
    """Device-resident SART solver: ``solution, convergence = SartSolver(geometry_matrix)(measurement_vector)``."""

    def __init__(self, geometry_matrix=None, laplacian_matrix=None, device=0, value_dtype=np.float64, csr=None, _device_csr=None):
        self._lib = _abi.load_library()
        self._h = None
        d = _abi.SartDesc()
        d.abi_version = _abi.ABI_VERSION
        d.value_f64 = int(np.dtype(value_dtype) == np.float64)
        keep = []
        if geometry_matrix is not None:
            g = np.asarray(geometry_matrix)
            if g.ndim != 2:
                raise ValueError("geometry_matrix must be an array with shape (N_d, N_s)")
            if g.dtype != np.float32:
                g = g.astype(np.float64, copy=False)
            g = np.ascontiguousarray(g)
            keep.append(g)
            d.n_detectors, d.n_sources = g.shape
            d.dense, d.dense_f64 = _ptr(g), int(g.dtype == np.float64)
        elif csr is not None:
            row_offset, columns, values, n_sources = csr
            row_offset = np.ascontiguousarray(row_offset, dtype=np.int64)
            columns = np.ascontiguousarray(columns, dtype=np.int32)
            values = np.ascontiguousarray(values, dtype=np.float64)
            if columns.shape != values.shape or row_offset[-1] != columns.size:
                raise ValueError("inconsistent CSR arrays")
            keep += [row_offset, columns, values]
            d.n_detectors, d.n_sources = row_offset.size - 1, int(n_sources)
            d.row_offset, d.columns, d.values = _ptr(row_offset), _ptr(columns), _ptr(values)
        elif _device_csr is not None:
            row_offset, columns, values, n_sources = _device_csr       # torch CUDA tensors: int64, int32, float64
            import torch
            for t, dt, name in ((row_offset, torch.int64, "row_offset"), (columns, torch.int32, "columns"), (values, torch.float64, "values")):
                if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dt and t.is_contiguous() and t.dim() == 1):
                    raise TypeError("%s must be a contiguous 1-D CUDA tensor of dtype %s" % (name, dt))
            if len({t.device for t in (row_offset, columns, values)}) != 1:
                raise ValueError("the CSR tensors must live on one device")
            if columns.numel() != values.numel() or row_offset.numel() < 1 or int(row_offset[-1].item()) != columns.numel():
                raise ValueError("inconsistent CSR arrays")
            if columns.numel() and (int(columns.min().item()) < 0 or int(columns.max().item()) >= int(n_sources)):
                raise ValueError("column index out of range")
            keep += [row_offset, columns, values]
            d.n_detectors, d.n_sources = row_offset.numel() - 1, int(n_sources)
            d.memory = 1
            d.row_offset, d.columns, d.values = (C.c_void_p(t.data_ptr()) for t in (row_offset, columns, values))
        else:
            raise ValueError("geometry matrix missing")
        lap = self._laplacian_arrays(laplacian_matrix, int(d.n_sources))
        keep += [a for a in lap if a is not None]
        d.laplacian_dense, d.lap_row_offset, d.lap_columns, d.lap_values = (_ptr(a) for a in lap)
        self.n_detectors, self.n_sources = int(d.n_detectors), int(d.n_sources)
        h = C.c_void_p()
        _abi.check(self._lib, self._lib.cb2_sart_create(C.byref(d), int(device), C.byref(h)))
        self._h = h

    @classmethod
    def from_csr(cls, row_offset, columns, values, n_sources, **kw):
        """Geometry matrix as host CSR arrays (what RayTransferScene.render_csr returns)."""
        return cls(csr=(row_offset, columns, values, n_sources), **kw)

    @classmethod
    def from_device_csr(cls, row_offset, columns, values, n_sources, **kw):
        """Geometry matrix as torch CUDA tensors (what RayTransferScene.render_csr_device returns): no host round trip."""
        kw.setdefault("device", row_offset.device.index or 0)
        return cls(_device_csr=(row_offset, columns, values, n_sources), **kw)

    @staticmethod
    def _laplacian_arrays(laplacian_matrix, n_sources):
        if laplacian_matrix is None:
            return None, None, None, None
        if isinstance(laplacian_matrix, tuple):                        # (row_offset, columns, values) CSR
            ro, co, va = laplacian_matrix
            ro = np.ascontiguousarray(ro, dtype=np.int64)
            if ro.size != n_sources + 1:
                raise ValueError("laplacian_matrix must have shape (N_s, N_s)")
            return None, ro, np.ascontiguousarray(co, dtype=np.int32), np.ascontiguousarray(va, dtype=np.float64)
        lap = np.ascontiguousarray(laplacian_matrix, dtype=np.float64)
        if lap.shape != (n_sources, n_sources):
            raise ValueError("laplacian_matrix must have shape (N_s, N_s)")
        return lap, None, None, None

    def update_laplacian_matrix(self, laplacian_matrix):
        """Replaces the Laplacian held on the device (SartOpencl.update_laplacian_matrix, sart_opencl.py:186-196)."""
        lap = self._laplacian_arrays(laplacian_matrix, self.n_sources)
        _abi.check(self._lib, self._lib.cb2_sart_set_laplacian(self._h, *(_ptr(a) for a in lap)))

    def clean(self):
        """Releases the device buffers (SartOpencl.clean, sart_opencl.py:166-178)."""
        if getattr(self, "_h", None):
            self._lib.cb2_sart_destroy(self._h)
            self._h = None

    close = clean

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.clean()

    def __del__(self):
        try:
            self.clean()
        except Exception:
            pass

    def info(self):
        keys = ("nnz", "laplacian_nnz", "bytes_per_iteration", "solve_ms", "iterations_launched")
        return {k: self._lib.cb2_sart_info(self._h, i) for i, k in enumerate(keys)}

    def __call__(self, measurement_vector, initial_guess=None, max_iterations=250, relaxation=1.0, beta_laplace=0.01, conv_tol=1.0e-4):
        """Returns (solution, convergence list) like the reference (sart.pyx:155); for a 2-D (frames, N_d) measurement
        array, (solutions[frames, N_s], list of per-frame convergence lists)."""
        m = np.ascontiguousarray(measurement_vector, dtype=np.float64)
        single = m.ndim == 1
        m = m.reshape(-1, m.shape[-1])
        if m.shape[1] != self.n_detectors:
            raise ValueError("measurement_vector must have shape (N_d)")
        n_frames = m.shape[0]
        guess, value = None, float(np.exp(-1))
        if initial_guess is None:
            pass
        elif isinstance(initial_guess, (float, int)):
            value = float(initial_guess)
        else:
            guess = np.ascontiguousarray(np.broadcast_to(np.asarray(initial_guess, dtype=np.float64).reshape(-1, self.n_sources),
                                                         (n_frames, self.n_sources)))
        max_iterations = int(max_iterations)
        if max_iterations < 1:
            # the reference's loop does not run: the seed comes back with an empty convergence list (sart.pyx:103-152)
            seed = guess.copy() if guess is not None else np.full((n_frames, self.n_sources), value)
            return (seed[0], []) if single else (seed, [[] for _ in range(n_frames)])
        solution = np.zeros((n_frames, self.n_sources))
        conv = np.zeros((n_frames, max(max_iterations, 1)))
        n_it = np.zeros(n_frames, dtype=np.int32)
        _abi.check(self._lib, self._lib.cb2_sart_solve(self._h, m.ctypes.data_as(_abi.c_double_p), n_frames,
                                                       guess.ctypes.data_as(_abi.c_double_p) if guess is not None else None, value,
                                                       max_iterations, float(relaxation), float(beta_laplace), float(conv_tol),
                                                       solution.ctypes.data_as(_abi.c_double_p), conv.ctypes.data_as(_abi.c_double_p),
                                                       n_it.ctypes.data_as(_abi.c_int32_p)))
        lists = [list(conv[f, :n_it[f]]) for f in range(n_frames)]
        return (solution[0], lists[0]) if single else (solution, lists)


def invert_sart(geometry_matrix, measurement_vector, initial_guess=None, max_iterations=250, relaxation=1.0, conv_tol=1.0e-4, device=0):
    """cherab.tools.inversions.invert_sart (sart.pyx:26-155) on the device."""
    with SartSolver(geometry_matrix, device=device) as solver:
        return solver(measurement_vector, initial_guess=initial_guess, max_iterations=max_iterations, relaxation=relaxation,
                      conv_tol=conv_tol)


def invert_constrained_sart(geometry_matrix, laplacian_matrix, measurement_vector, initial_guess=None, max_iterations=250,
                            relaxation=1.0, beta_laplace=0.01, conv_tol=1.0e-4, device=0):
    """cherab.tools.inversions.invert_constrained_sart (sart.pyx:161-302) on the device."""
    with SartSolver(geometry_matrix, laplacian_matrix=laplacian_matrix, device=device) as solver:
        return solver(measurement_vector, initial_guess=initial_guess, max_iterations=max_iterations, relaxation=relaxation,
                      beta_laplace=beta_laplace, conv_tol=conv_tol)
