from raysect.optical._stub cimport Point3D, AffineMatrix3D, Spectrum, World, Primitive, Ray
