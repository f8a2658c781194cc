# cython: language_level=3
import numpy as np
cimport numpy as np
from libc.math cimport sqrt


cdef class Point3D:
    def __init__(self, double x=0.0, double y=0.0, double z=0.0):
        self.x, self.y, self.z = x, y, z

    cpdef double distance_to(self, Point3D p):
        return sqrt((p.x - self.x) ** 2 + (p.y - self.y) ** 2 + (p.z - self.z) ** 2)


cdef class AffineMatrix3D:
    def __init__(self):
        cdef int i, j
        for i in range(4):
            for j in range(4):
                self.m[i][j] = 1.0 if i == j else 0.0


cdef class Spectrum:
    def __init__(self, double min_wavelength, double max_wavelength, int bins):
        self.min_wavelength, self.max_wavelength, self.bins = min_wavelength, max_wavelength, bins
        self.delta_wavelength = (max_wavelength - min_wavelength) / bins
        self.samples = np.zeros(bins, dtype=np.float64)
        self.samples_mv = self.samples


cdef class World:
    pass


cdef class Primitive:
    pass


cdef class Ray:
    def __init__(self, Point3D origin=None):
        self.origin = origin if origin is not None else Point3D()
