from raysect.optical cimport World, Ray, Primitive, Point3D, Spectrum, AffineMatrix3D


cdef class InhomogeneousVolumeEmitter:
    cdef public VolumeIntegrator integrator


cdef class VolumeIntegrator:
    cpdef Spectrum integrate(self, Spectrum spectrum, World world, Ray ray, Primitive primitive,
                             InhomogeneousVolumeEmitter material, Point3D start_point, Point3D end_point,
                             AffineMatrix3D world_to_primitive, AffineMatrix3D primitive_to_world)
