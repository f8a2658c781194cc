"""Device-side scene handles: thin Python wrappers over the C ABI of libcherab_b200.so.

``EmissionScene.render`` is the reference-facing call with HOST (numpy) buffers: ray segments in, spectra out,
host<->device copies inside.  ``render_device`` takes torch CUDA tensors (PyTorch is plumbing for device memory and
streams only) and launches on the current stream.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _abi


class EmissionScene:
    """Device-resident flattened plasma scene (cb2_scene_create / cb2_scene_destroy)."""

    def __init__(self, flat, device=0):
        self._lib = _abi.load_library()
        self.flat = flat
        self.device = int(device)
        self.bins = int(flat.desc.grid.bins)
        self._h = C.c_void_p()
        _abi.check(self._lib, self._lib.cb2_scene_create(C.byref(flat.desc), self.device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.cb2_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, rays, out=None, scale=1.0, accumulate=False, dtype=np.float64, out_of_domain="raise", rows=None):
        """spectra[n_rays, bins] (+)= scale * integral of the emission along every ray.  Returns (out, stats dict).

        Samples that leave a table built without extrapolation (rates, the psi grid) are clamped to the table edge and counted on
        the device; the reference raises ValueError from the interpolator there ('none' extrapolation), so does this
        call once the frame is back — ``out_of_domain="count"`` returns the clamped result with ``stats["out_of_domain"]`` instead."""
        st = _abi.Stats()
        rs = rays.as_struct()
        if rows is not None:
            # ``rows[i]``: the row of the frame ``out`` [n_rows, bins] that receives ray i (cb2_emission_render_rows): a rank's tiles
            # land on their pixels of the image-ordered frame; not accumulated
            rows = np.ascontiguousarray(rows, dtype=np.int64)
            if out is None or accumulate:
                raise ValueError("rows needs a frame to write into (out=...) and does not accumulate")
            if (out.dtype not in (np.float64, np.float32) or not out.flags.c_contiguous or out.ndim != 2 or out.shape[1] != self.bins
                    or rows.shape != (rays.n_rays,) or (rows.size and (rows.min() < 0 or rows.max() >= out.shape[0]))):
                raise ValueError("out must be a C-contiguous float32/float64 frame [n_rows, bins] and rows one valid row index per ray")
            _abi.check(self._lib, self._lib.cb2_emission_render_rows(self._h, C.byref(rs), rows.ctypes.data_as(_abi.c_int64_p),
                                                                     out.ctypes.data_as(C.c_void_p), int(out.dtype == np.float64),
                                                                     float(scale), C.byref(st)))
        else:
            if out is None:
                out = np.zeros((rays.n_rays, self.bins), dtype=dtype)
                accumulate = False
            if out.dtype not in (np.float64, np.float32) or not out.flags.c_contiguous or out.shape != (rays.n_rays, self.bins):
                raise ValueError("out must be a C-contiguous float32/float64 array of shape (n_rays, bins)")
            _abi.check(self._lib, self._lib.cb2_emission_render(self._h, C.byref(rs), out.ctypes.data_as(C.c_void_p),
                                                                int(out.dtype == np.float64), float(scale), int(accumulate), C.byref(st)))
        stats = st.as_dict()
        if out_of_domain == "raise" and stats["out_of_domain"] > 0:
            raise ValueError("The specified value is outside of the range of the supplied data and/or extrapolation range: %d "
                             "table lookups of this render left their tables (pass out_of_domain='count' for the clamped result)."
                             % stats["out_of_domain"])
        return out, stats

    def render_device(self, dev_rays, out, scale=1.0, accumulate=False, stats=None):
        """Device-resident render: ``dev_rays`` is a DeviceRays, ``out`` a torch CUDA tensor [n_rays, bins] (fp32/fp64).
        Launches on the current stream; the library synchronises that stream once per ray batch (it reads the batch's group
        count), so the call returns with at most the last batch in flight.  A scene handle owns its scratch buffers: one
        stream at a time per handle.  Out-of-domain lookups are only counted (``stats[5]``); the caller decides."""
        import torch
        rs = dev_rays.as_struct()
        is64 = out.dtype == torch.float64
        stream = torch.cuda.current_stream(out.device).cuda_stream
        _abi.check(self._lib, self._lib.cb2_emission_render_device(self._h, C.byref(rs), C.c_void_p(out.data_ptr()), int(is64),
                                                                   float(scale), int(accumulate),
                                                                   C.c_void_p(stats.data_ptr()) if stats is not None else None,
                                                                   C.c_void_p(stream)))
        return out

    def info(self):
        """Launch plan of this scene: CTA shape and the Bremsstrahlung formulation in use (cb2_scene_info)."""
        keys = ("warps_per_cta", "bins_per_lane", "brems_mode", "moment_row", "temperature_nodes", "distinct_charges", "batch_rays",
                "two_kernel_line_path", "contraction_on_tensor_cores", "state_table_intervals", "state_table_error_1e9")
        d = {k: int(self._lib.cb2_scene_info(self._h, i)) for i, k in enumerate(keys)}
        d["brems_mode"] = {0: "none", 1: "direct", 3: "moments"}[d["brems_mode"]]
        return d

    def profile(self, enable):
        """Per-kernel device timing (cb2_scene_profile).  profile(True) starts; profile(False) stops and returns
        {kernel: (milliseconds, launches)} accumulated since the start."""
        if enable:
            _abi.check(self._lib, self._lib.cb2_scene_profile(self._h, 1, None, None))
            return None
        ms = (C.c_double * 4)()
        n = (C.c_int64 * 4)()
        _abi.check(self._lib, self._lib.cb2_scene_profile(self._h, 0, ms, n))
        names = ("state_kernel", "bin_kernel", "contract_kernel", "helpers")
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(names)}

    def sample_state(self, points):
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        w = self._lib.cb2_state_width(self._h)
        out = np.zeros((pts.shape[0], w), dtype=np.float64)
        _abi.check(self._lib, self._lib.cb2_sample_state(self._h, pts.ctypes.data_as(_abi.c_double_p), pts.shape[0],
                                                         out.ctypes.data_as(_abi.c_double_p)))
        return out


class PlasmaRenderer:
    """The scene-plumbing seam: keeps the device scene of a Plasma for one spectral grid and rebuilds it
    only after the plasma told its notifier that something changed (Plasma._modified, cherab/core/plasma/node.pyx:545-554 —
    composition, models, distributions, geometry, integrator, atomic data).  The flattening that the reference spreads
    over every model's lazy ``_populate_cache`` happens once per change, on the first render after it."""

    def __init__(self, plasma, min_wavelength, max_wavelength, bins, device=0, **flatten_kwargs):
        self.plasma = plasma
        self.grid = (float(min_wavelength), float(max_wavelength), int(bins))
        self.device = int(device)
        self.flatten_kwargs = flatten_kwargs
        self._scene = None
        self.rebuilds = 0
        plasma.notifier.add(self._invalidate)

    def _invalidate(self):
        if self._scene is not None:
            self._scene.close()
        self._scene = None

    @property
    def scene(self):
        if self._scene is None:
            from .flatten import flatten_scene
            self._scene = EmissionScene(flatten_scene(self.plasma, *self.grid, **self.flatten_kwargs), device=self.device)
            self.rebuilds += 1
        return self._scene

    def render(self, rays, **kw):
        return self.scene.render(rays, **kw)

    def render_device(self, dev_rays, out, **kw):
        return self.scene.render_device(dev_rays, out, **kw)

    def close(self):
        self.plasma.notifier.remove(self._invalidate)
        self._invalidate()


class DeviceRays:
    """A RayBatch resident in device memory (torch tensors)."""

    def __init__(self, rays, device="cuda:0", pin=False):
        import torch
        self.n_rays, self.n_segments = rays.n_rays, rays.n_segments
        self.device = torch.device(device)

        def up(a):
            t = torch.from_numpy(a)
            if pin:
                t = t.pin_memory()
            return t.to(self.device, non_blocking=pin)
        self.origin, self.direction = up(rays.origin), up(rays.direction)
        self.seg_offset, self.seg_t0, self.seg_t1 = up(rays.seg_offset), up(rays.seg_t0), up(rays.seg_t1)
        self.nbytes = sum(t.numel() * t.element_size() for t in (self.origin, self.direction, self.seg_offset, self.seg_t0, self.seg_t1))

    def as_struct(self):
        r = _abi.Rays()
        r.n_rays, r.n_segments = self.n_rays, self.n_segments
        r.origin = C.cast(C.c_void_p(self.origin.data_ptr()), _abi.c_double_p)
        r.direction = C.cast(C.c_void_p(self.direction.data_ptr()), _abi.c_double_p)
        r.seg_offset = C.cast(C.c_void_p(self.seg_offset.data_ptr()), _abi.c_int64_p)
        r.seg_t0 = C.cast(C.c_void_p(self.seg_t0.data_ptr()), _abi.c_double_p)
        r.seg_t1 = C.cast(C.c_void_p(self.seg_t1.data_ptr()), _abi.c_double_p)
        return r


class RayTransferScene:
    """Device-resident ray-transfer grid (cb2_rt_create / cb2_rt_destroy) for a RayTransferCylinder / RayTransferBox."""

    def __init__(self, rt_object, device=0):
        self._lib = _abi.load_library()
        self.rt = rt_object
        self.bins = rt_object.bins
        self.desc, self._keep = rt_object.descriptor()
        self.device = int(device)
        self._h = C.c_void_p()
        _abi.check(self._lib, self._lib.cb2_rt_create(C.byref(self.desc), self.device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.cb2_rt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render_dense(self, rays):
        """matrix[n_rays, bins] of path lengths (m) — what the reference integrator leaves in Spectrum.samples."""
        out = np.zeros((rays.n_rays, self.bins), dtype=np.float64)
        st = _abi.Stats()
        rs = rays.as_struct()
        _abi.check(self._lib, self._lib.cb2_rt_render_dense(self._h, C.byref(rs), out.ctypes.data_as(_abi.c_double_p), 0, C.byref(st)))
        return out, st.as_dict()

    def render_csr(self, rays, capacity=None):
        """CSR geometry matrix: (row_offset[n_rays+1], columns, lengths, stats).  Grows the buffers once if needed."""
        n = rays.n_rays
        cap = int(capacity) if capacity else max(1024, 64 * n)
        rs = rays.as_struct()
        for _ in range(2):
            row_offset = np.zeros(n + 1, dtype=np.int64)
            columns = np.zeros(cap, dtype=np.int32)
            lengths = np.zeros(cap, dtype=np.float64)
            st = _abi.Stats()
            rc = self._lib.cb2_rt_render_csr(self._h, C.byref(rs), row_offset.ctypes.data_as(_abi.c_int64_p),
                                             columns.ctypes.data_as(_abi.c_int32_p), lengths.ctypes.data_as(_abi.c_double_p),
                                             cap, C.byref(st))
            if rc == -7:  # CB2_ERR_OVERFLOW: required capacity is in row_offset[n]
                cap = int(row_offset[n])
                continue
            _abi.check(self._lib, rc)
            nnz = int(row_offset[n])
            return row_offset, columns[:nnz], lengths[:nnz], st.as_dict()
        raise OverflowError("CSR capacity negotiation failed")

    def render_sparse(self, rays, capacity=None):
        """The geometry matrix as a ``scipy.sparse.csr_matrix`` of shape (n_rays, bins) — what RayTransferPipeline*.matrix holds
        densely in the reference (pipelines.py:198); SURVEY H8: 400 x 800 cells x 512^2 rays do not fit a dense ndarray."""
        import scipy.sparse
        row_offset, columns, lengths, stats = self.render_csr(rays, capacity)
        return scipy.sparse.csr_matrix((lengths, columns, row_offset), shape=(rays.n_rays, self.bins)), stats

    def render_csr_device(self, dev_rays, capacity):
        """Device-resident CSR build: returns torch tensors (row_offset, columns, lengths) on the rays' device."""
        import torch
        dev = dev_rays.device
        row_offset = torch.zeros(dev_rays.n_rays + 1, dtype=torch.int64, device=dev)
        columns = torch.empty(capacity, dtype=torch.int32, device=dev)
        lengths = torch.empty(capacity, dtype=torch.float64, device=dev)
        nnz = C.c_int64(0)
        rs = dev_rays.as_struct()
        stream = torch.cuda.current_stream(dev).cuda_stream
        _abi.check(self._lib, self._lib.cb2_rt_render_csr_device(self._h, C.byref(rs), C.c_void_p(row_offset.data_ptr()),
                                                                 C.c_void_p(columns.data_ptr()), C.c_void_p(lengths.data_ptr()),
                                                                 int(capacity), C.byref(nnz), None, C.c_void_p(stream)))
        return row_offset, columns[:nnz.value], lengths[:nnz.value]


def measure_peaks(device=0):
    """Live FP32-FMA, MUFU.EX2 and FP64-FMA issue-rate microbenchmarks -> dict(fp32_tflops, sfu_tops, sm_clock_mhz, fp64_tflops)."""
    lib = _abi.load_library()
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    _abi.check(lib, lib.cb2_measure_peaks(int(device), C.byref(a), C.byref(b), C.byref(c)))
    d = C.c_double(0)
    _abi.check(lib, lib.cb2_measure_peak_fp64(int(device), C.byref(d)))
    return {"fp32_tflops": a.value, "sfu_tops": b.value, "sm_clock_mhz": c.value, "fp64_tflops": d.value}
