# the declarations of raysect.optical the shim cimports (subset; same names and member types as Raysect 0.8.1)
cimport numpy as np


cdef class Point3D:
    cdef public double x, y, z
    cpdef double distance_to(self, Point3D p)


cdef class AffineMatrix3D:
    cdef double m[4][4]


cdef class Spectrum:
    cdef:
        readonly double min_wavelength, max_wavelength, delta_wavelength
        readonly int bins
        public np.ndarray samples
        double[::1] samples_mv


cdef class World:
    pass


cdef class Primitive:
    pass


cdef class Ray:
    cdef public Point3D origin
