"""core_b200 — B200-native implementation of Cherab's per-ray emission-integration and ray-transfer hot path.

CUDA sm_100a kernels behind the C ABI of include/cherab_b200.h (core_b200/csrc/libcherab_b200.so), with a host-side
mirror of the reference's operator interface for this path (Plasma, Species, Maxwellian, ExcitationLine,
RecombinationLine, Bremsstrahlung, line shapes, RayTransferCylinder/Box and the RayTransferPipelines).
No CPU fallback: calls fail loudly if the CUDA library or a GPU is missing.
"""
from .atomic import (AtomicData, ConstantRate, Element, Line, RateTable, RateTable3D, SyntheticADAS, carbon, deuterium, helium,
                     hydrogen, neon, nitrogen, tritium)
from .beam import (Beam, BeamCXLine, BeamEmissionLine, BeamCXTable, BeamStoppingTable, ConstantBeamCXPEC, SingleRayAttenuator, beam_ray_segments,
                   flatten_beam_scene)
from .first_wall import FirstWall, load_first_wall
from .flatten import FlatScene, RayBatch, flatten_scene
from .geometry import Box, HollowCylinder, PinholeCamera, Sphere, look_at, ray_segments, stratified_offsets, translate
from .inversions import SartSolver, invert_constrained_sart, invert_sart
from .models import (Bremsstrahlung, ExcitationLine, GaussianLine, MultipletLineShape, ParametrisedZeemanTriplet,
                     RecombinationLine, StarkBroadenedLine, ThermalCXLine, TotalRadiatedPower, ZeemanMultiplet, ZeemanStructure, ZeemanTriplet)
from .notify import Notifier
from .observers import (DevicePinhole, DeviceRayBuffer, FibreOptic, FibreOpticGroup, PowerPipeline0D, RadiancePipeline0D, RadiancePipeline2D,
                        SightLine, SightLineGroup, SpectralPowerPipeline0D, SpectralRadiancePipeline0D, SpectralRadiancePipeline2D, observe)
from .openadas import OpenADAS
from .plasma import (AxisymBlend, AxisymBlendVector, AxisymContext, Constant3D, ConstantVector3D, EFITEquilibrium,
                     EFITMagneticField, GaussianVolume, Maxwellian, ModelManager, NumericalIntegrator, Plasma, SlabIonFunction,
                     SlabNeutralFunction, Species)

__version__ = "0.1.0"
