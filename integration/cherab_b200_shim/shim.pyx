# cython: language_level=3
# distutils: libraries = cherab_b200
"""The Cython binding of libcherab_b200.so a Cherab maintainer would add (INTEGRATION.md).

  B200Integrator   seam S1: a raysect VolumeIntegrator; Plasma.integrator / Beam.integrator accept any instance
                   (cherab/core/plasma/node.pyx:482-489, cherab/core/beam/node.pyx:464-471).  Raysect calls integrate() once per
                   [entry, exit] pair of a ray in the plasma primitive (NumericalIntegrator.integrate is what it replaces); the
                   segment goes through cb2_emission_render with accumulate = 1, because the contract is spectrum += integral.
  render_segments  seam S4: the batch call an observe()-level driver makes, GIL released.

The scene handle is created from the address of a flat cb2_scene_desc (core_b200.flatten.FlatScene.desc, or the output of the
Cython flattener for genuine cherab objects).  Errors come back as the reference's exception types.
"""
from libc.stdint cimport int32_t, int64_t, uintptr_t
from raysect.optical cimport World, Ray, Primitive, Point3D, Spectrum, AffineMatrix3D
from raysect.optical.material.emitter.inhomogeneous cimport VolumeIntegrator, InhomogeneousVolumeEmitter

cdef extern from "cherab_b200.h" nogil:
    ctypedef struct cb2_scene:
        pass
    ctypedef struct cb2_scene_desc:
        pass
    ctypedef struct cb2_rays:
        int64_t n_rays
        int64_t n_segments
        const double *origin
        const double *direction
        const int64_t *seg_offset
        const double *seg_t0
        const double *seg_t1
    ctypedef struct cb2_stats:
        int64_t samples
    int cb2_abi_version()
    int cb2_scene_create(const cb2_scene_desc *desc, int device, cb2_scene **out)
    int cb2_scene_destroy(cb2_scene *scene)
    int cb2_emission_render(cb2_scene *scene, const cb2_rays *rays, void *out, int out_f64, double scale, int accumulate,
                            cb2_stats *stats)
    const char *cb2_last_error()


cdef object _raise(int rc):
    # status codes of include/cherab_b200.h -> the exception types the reference raises on this path
    msg = cb2_last_error().decode("utf-8", "replace")
    if rc == -1:
        return ValueError(msg)
    if rc == -3:
        return TypeError(msg)
    if rc == -4:
        return NotImplementedError(msg)
    if rc == -6:
        return MemoryError(msg)
    if rc == -7:
        return OverflowError(msg)
    return RuntimeError(msg)                 # CB2_ERR_RUNTIME, CB2_ERR_CUDA


cdef class B200Integrator(VolumeIntegrator):
    """VolumeIntegrator over a flattened scene on one GPU."""
    cdef cb2_scene *_scene
    cdef object _flat            # keeps the descriptor (and the arrays it points at) alive

    def __cinit__(self):
        self._scene = NULL

    def __init__(self, object flat_scene, uintptr_t desc_address, int device=0):
        cdef int rc
        self._flat = flat_scene
        rc = cb2_scene_create(<const cb2_scene_desc *> desc_address, device, &self._scene)
        if rc != 0:
            raise _raise(rc)

    def __dealloc__(self):
        if self._scene != NULL:
            cb2_scene_destroy(self._scene)
            self._scene = NULL

    cpdef Spectrum integrate(self, Spectrum spectrum, World world, Ray ray, Primitive primitive,
                             InhomogeneousVolumeEmitter material, Point3D start_point, Point3D end_point,
                             AffineMatrix3D world_to_primitive, AffineMatrix3D primitive_to_world):
        cdef double o[3]
        cdef double d[3]
        cdef double t0 = 0.0, t1
        cdef int64_t off[2]
        cdef cb2_rays r
        cdef int rc
        # raysect hands start_point = far end, end_point = near end of the segment (world space); the library marches from the far
        # end exactly as NumericalIntegrator does, given origin = near end and the unit direction towards the far end
        t1 = end_point.distance_to(start_point)
        if t1 == 0.0:
            return spectrum
        o[0] = end_point.x; o[1] = end_point.y; o[2] = end_point.z
        d[0] = (start_point.x - end_point.x) / t1
        d[1] = (start_point.y - end_point.y) / t1
        d[2] = (start_point.z - end_point.z) / t1
        off[0] = 0; off[1] = 1
        r.n_rays = 1; r.n_segments = 1
        r.origin = o; r.direction = d; r.seg_offset = off; r.seg_t0 = &t0; r.seg_t1 = &t1
        with nogil:
            rc = cb2_emission_render(self._scene, &r, &spectrum.samples_mv[0], 1, 1.0, 1, NULL)
        if rc != 0:
            raise _raise(rc)
        return spectrum

    def render_segments(self, double[:, ::1] origin, double[:, ::1] direction, int64_t[::1] seg_offset, double[::1] seg_t0,
                        double[::1] seg_t1, double[:, ::1] out, double scale=1.0, bint accumulate=False):
        """Seam S4: every ray segment of a frame in one call; out[n_rays, bins] (+)= scale * radiance.  Returns the sample count."""
        cdef cb2_rays r
        cdef cb2_stats st
        cdef int rc
        if origin.shape[0] != direction.shape[0] or origin.shape[1] != 3 or direction.shape[1] != 3 or seg_offset.shape[0] != origin.shape[0] + 1 \
                or seg_t0.shape[0] != seg_t1.shape[0] or out.shape[0] != origin.shape[0]:
            raise ValueError("inconsistent ray arrays")
        r.n_rays = origin.shape[0]; r.n_segments = seg_t0.shape[0]
        if r.n_rays == 0:
            return 0
        r.origin = &origin[0, 0]; r.direction = &direction[0, 0]; r.seg_offset = &seg_offset[0]
        r.seg_t0 = &seg_t0[0] if r.n_segments else NULL
        r.seg_t1 = &seg_t1[0] if r.n_segments else NULL
        with nogil:
            rc = cb2_emission_render(self._scene, &r, &out[0, 0], 1, scale, 1 if accumulate else 0, &st)
        if rc != 0:
            raise _raise(rc)
        return st.samples


def abi_version():
    return cb2_abi_version()
