"""Stand-in for the handful of raysect.optical declarations integration/cherab_b200_shim uses."""
from ._stub import Point3D, AffineMatrix3D, Spectrum, World, Primitive, Ray
