"""ctypes mirror of include/cherab_b200.h (the C ABI of the hot path) and the loader of the CUDA library.

The product library is ``core_b200/csrc/libcherab_b200.so`` (built in-tree by ``__graft_entry__.build()``).
There is NO CPU fallback: if the library is missing or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

ABI_VERSION = 7

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)

# status codes -> the exception type the reference raises at the same point (include/cherab_b200.h cb2_status)
STATUS_EXC = {-1: ValueError, -2: RuntimeError, -3: TypeError, -4: NotImplementedError,
              -5: RuntimeError, -6: MemoryError, -7: OverflowError}

FIELD_CONSTANT, FIELD_GAUSSIAN_VOLUME, FIELD_AXISYM_BLEND, FIELD_SLAB_ION, FIELD_SLAB_NEUTRAL = range(5)
SHAPE_GAUSSIAN, SHAPE_MULTIPLET, SHAPE_ZEEMAN_TRIPLET, SHAPE_PARAM_ZEEMAN, SHAPE_ZEEMAN_MULTIPLET, SHAPE_STARK = range(6)
POL_PI, POL_SIGMA, POL_NO = range(3)
MODEL_EXCITATION_LINE, MODEL_RECOMBINATION_LINE, MODEL_BREMSSTRAHLUNG, MODEL_THERMAL_CX_LINE, MODEL_TOTAL_RADIATED_POWER, MODEL_BEAM_CX_LINE, MODEL_BEAM_EMISSION_LINE = range(7)
RT_CYLINDRICAL, RT_CARTESIAN = range(2)


class SpectralGrid(C.Structure):
    _fields_ = [("min_wavelength", C.c_double), ("max_wavelength", C.c_double), ("bins", C.c_int32), ("_pad", C.c_int32)]


class ScalarField(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("c", C.c_double * 8), ("edge", c_double_p), ("core", c_double_p)]


class VectorField(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("c", C.c_double * 8),
                ("core_vtor", c_double_p), ("core_vpol", c_double_p), ("core_vnorm", c_double_p)]


class Equilibrium(C.Structure):
    _fields_ = [("nr", C.c_int32), ("nz", C.c_int32), ("r", c_double_p), ("z", c_double_p), ("psi", c_double_p),
                ("psi_axis", C.c_double), ("psi_lcfs", C.c_double), ("n_f", C.c_int32), ("n_lcfs", C.c_int32),
                ("f_psin", c_double_p), ("f_value", c_double_p), ("lcfs_polygon", c_double_p),
                ("b_vacuum_radius", C.c_double), ("b_vacuum_magnitude", C.c_double)]


class Axisym(C.Structure):
    _fields_ = [("eq", Equilibrium), ("n_vertices", C.c_int32), ("n_triangles", C.c_int32),
                ("vertices", c_double_p), ("triangles", c_int32_p), ("n_core", C.c_int32), ("n_mask", C.c_int32),
                ("core_psin", c_double_p), ("mask_x", c_double_p), ("mask_y", c_double_p)]


class SpeciesDesc(C.Structure):
    _fields_ = [("charge", C.c_int32), ("_pad", C.c_int32), ("atomic_weight", C.c_double),
                ("density", ScalarField), ("temperature", ScalarField), ("velocity", VectorField)]


class Rate2D(C.Structure):
    _fields_ = [("n_ne", C.c_int32), ("n_te", C.c_int32), ("ne", c_double_p), ("te", c_double_p), ("rate", c_double_p),
                ("constant", C.c_double), ("extrapolate", C.c_int32), ("_pad", C.c_int32)]


class Gaunt(C.Structure):
    _fields_ = [("n_u", C.c_int32), ("n_gamma2", C.c_int32), ("u", c_double_p), ("gamma2", c_double_p), ("gaunt", c_double_p)]


class LineShape(C.Structure):
    _fields_ = [("kind", C.c_int32), ("polarisation", C.c_int32), ("param", C.c_double * 3),
                ("n_components", C.c_int32), ("n_b", C.c_int32), ("multiplet", c_double_p),
                ("n_pi", C.c_int32), ("n_sigma_plus", C.c_int32), ("n_sigma_minus", C.c_int32), ("_pad", C.c_int32),
                ("b_grid", c_double_p), ("zeeman_wavelength", c_double_p), ("zeeman_ratio", c_double_p)]


class Rate3D(C.Structure):
    _fields_ = [("n_ne", C.c_int32), ("n_te", C.c_int32), ("n_td", C.c_int32), ("_pad", C.c_int32), ("ne", c_double_p), ("te", c_double_p),
                ("td", c_double_p), ("rate", c_double_p), ("constant", C.c_double), ("extrapolate", C.c_int32), ("_pad2", C.c_int32)]


class BeamRate(C.Structure):
    _fields_ = [("n_e", C.c_int32), ("n_n", C.c_int32), ("n_t", C.c_int32), ("extrapolate", C.c_int32), ("e", c_double_p), ("n", c_double_p),
                ("t", c_double_p), ("sen", c_double_p), ("st", c_double_p), ("sref", C.c_double), ("constant", C.c_double)]


class CXRate(C.Structure):
    _fields_ = [("n_eb", C.c_int32), ("n_ti", C.c_int32), ("n_ni", C.c_int32), ("n_z", C.c_int32), ("n_b", C.c_int32), ("extrapolate", C.c_int32),
                ("eb", c_double_p), ("ti", c_double_p), ("ni", c_double_p), ("z", c_double_p), ("b", c_double_p),
                ("qeb", c_double_p), ("qti", c_double_p), ("qni", c_double_p), ("qz", c_double_p), ("qb", c_double_p),
                ("qref", C.c_double), ("constant", C.c_double)]


class BeamDesc(C.Structure):
    _fields_ = [("beam_to_plasma", C.c_double * 12), ("energy", C.c_double), ("power", C.c_double), ("temperature", C.c_double),
                ("atomic_weight", C.c_double), ("sigma", C.c_double), ("divergence_x", C.c_double), ("divergence_y", C.c_double),
                ("length", C.c_double), ("attenuator_step", C.c_double), ("clamp_sigma", C.c_double), ("clamp_to_zero", C.c_int32),
                ("n_stopping", C.c_int32), ("stopping_species", c_int32_p), ("stopping_rates", C.POINTER(BeamRate))]


class ModelExt(C.Structure):
    _fields_ = [("n_donors", C.c_int32), ("_pad", C.c_int32), ("donor_species", c_int32_p), ("donor_rates", C.POINTER(Rate3D)),
                ("line_rad_species", C.c_int32), ("recom_species", C.c_int32), ("n_hydrogen", C.c_int32), ("has_plt", C.c_int32),
                ("has_prb", C.c_int32), ("has_prc", C.c_int32), ("hydrogen_species", c_int32_p),
                ("plt", Rate2D), ("prb", Rate2D), ("prc", Rate2D), ("n_cx", C.c_int32), ("_pad3", C.c_int32), ("cx", C.POINTER(CXRate)),
                ("cx_population", C.POINTER(BeamRate)),
                ("n_bes", C.c_int32), ("_pad4", C.c_int32), ("bes_species", c_int32_p), ("bes_rates", C.POINTER(BeamRate)),
                ("mse_ratios", C.c_double * 4), ("n_mse", C.c_int32), ("_pad5", C.c_int32), ("mse_lne0", C.c_double), ("mse_dlne", C.c_double),
                ("mse_ratio_tab", c_double_p)]


class ModelDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("species", C.c_int32), ("wavelength", C.c_double), ("atomic_weight", C.c_double),
                ("pec", Rate2D), ("shape", LineShape), ("ext", C.POINTER(ModelExt))]


class SceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_species", C.c_int32), ("n_models", C.c_int32), ("min_samples", C.c_int32),
                ("step", C.c_double), ("grid", SpectralGrid), ("world_to_plasma", C.c_double * 12),
                ("electron_density", ScalarField), ("electron_temperature", ScalarField),
                ("species", C.POINTER(SpeciesDesc)), ("models", C.POINTER(ModelDesc)), ("axisym", C.POINTER(Axisym)),
                ("b_field_kind", C.c_int32), ("brems_quadrature", C.c_int32), ("b_field", C.c_double * 3),
                ("gaunt", Gaunt), ("quad_rtol", C.c_double), ("quad_max_order", C.c_int32), ("quad_min_order", C.c_int32),
                ("beam", C.POINTER(BeamDesc))]


class Rays(C.Structure):
    _fields_ = [("n_rays", C.c_int64), ("n_segments", C.c_int64), ("origin", c_double_p), ("direction", c_double_p),
                ("seg_offset", c_int64_p), ("seg_t0", c_double_p), ("seg_t1", c_double_p)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_int64), ("gaussian_bin_evals", C.c_int64), ("lorentzian_bin_evals", C.c_int64),
                ("brems_bin_evals", C.c_int64), ("rt_steps", C.c_int64), ("out_of_domain", C.c_int64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class RTDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("kind", C.c_int32), ("grid_shape", C.c_int32 * 3), ("min_samples", C.c_int32),
                ("grid_steps", C.c_double * 3), ("rmin", C.c_double), ("period", C.c_double), ("step", C.c_double),
                ("world_to_local", C.c_double * 12), ("voxel_map", c_int32_p), ("bins", C.c_int32), ("integrator", C.c_int32)]


class PrimitiveDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("p", C.c_double * 6), ("world_to_local", C.c_double * 12)]


class PinholeDesc(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("width", C.c_double), ("to_world", C.c_double * 12)]


class Observer0DDesc(C.Structure):
    _fields_ = [("to_world", C.c_double * 12), ("radius", C.c_double), ("acceptance_angle", C.c_double), ("samples", C.c_int32), ("_pad", C.c_int32)]


PRIM_HOLLOW_CYLINDER, PRIM_SPHERE, PRIM_BOX = range(3)


class SartDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("value_f64", C.c_int32), ("n_detectors", C.c_int64), ("n_sources", C.c_int64),
                ("dense", C.c_void_p), ("dense_f64", C.c_int32), ("memory", C.c_int32),
                ("row_offset", C.c_void_p), ("columns", C.c_void_p), ("values", C.c_void_p),
                ("laplacian_dense", C.c_void_p), ("lap_row_offset", C.c_void_p), ("lap_columns", C.c_void_p), ("lap_values", C.c_void_p)]


class WallDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("_pad", C.c_int32), ("n_triangles", C.c_int64), ("vertices", c_double_p)]


# every symbol include/cherab_b200.h declares for the product library
PRODUCT_SYMBOLS = [
    "cb2_abi_version", "cb2_last_error", "cb2_device_count", "cb2_measure_peaks", "cb2_measure_peak_fp64", "cb2_scene_create", "cb2_scene_destroy",
    "cb2_emission_render", "cb2_emission_render_rows", "cb2_rows_plan", "cb2_emission_render_device", "cb2_sample_state", "cb2_state_width", "cb2_scene_info", "cb2_scene_profile", "cb2_beam_sample",
    "cb2_rt_create", "cb2_rt_destroy", "cb2_rt_render_dense", "cb2_rt_render_csr", "cb2_rt_render_csr_device",
    "cb2_pinhole_rays_device", "cb2_observer0d_rays_device", "cb2_observer0d_reduce_device",
    "cb2_wall_create", "cb2_wall_destroy", "cb2_wall_hit", "cb2_wall_clip_device",
    "cb2_sart_create", "cb2_sart_destroy", "cb2_sart_set_laplacian", "cb2_sart_solve", "cb2_sart_info",
]

# CB2_LIB selects an alternative in-tree build of the same library (kernel-tuning experiments)
LIB_PATH = os.environ.get("CB2_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libcherab_b200.so")
_lib = None


def load_library():
    """Load libcherab_b200.so; fail loudly (no fallback) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("core_b200: CUDA library %s is missing — run `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.cb2_abi_version.restype = C.c_int
    lib.cb2_last_error.restype = C.c_char_p
    lib.cb2_device_count.restype = C.c_int
    lib.cb2_measure_peaks.argtypes = [C.c_int, c_double_p, c_double_p, c_double_p]
    lib.cb2_scene_create.argtypes = [C.POINTER(SceneDesc), C.c_int, C.POINTER(vp)]
    lib.cb2_scene_destroy.argtypes = [vp]
    lib.cb2_emission_render.argtypes = [vp, C.POINTER(Rays), vp, C.c_int, C.c_double, C.c_int, C.POINTER(Stats)]
    lib.cb2_emission_render_rows.argtypes = [vp, C.POINTER(Rays), c_int64_p, vp, C.c_int, C.c_double, C.POINTER(Stats)]
    lib.cb2_rows_plan.argtypes = [c_int64_p, C.c_int64, c_int64_p, C.c_int64]
    lib.cb2_rows_plan.restype = C.c_int64
    lib.cb2_emission_render_device.argtypes = [vp, C.POINTER(Rays), vp, C.c_int, C.c_double, C.c_int, vp, vp]
    lib.cb2_sample_state.argtypes = [vp, c_double_p, C.c_int64, c_double_p]
    lib.cb2_state_width.argtypes = [vp]
    lib.cb2_scene_info.argtypes = [vp, C.c_int]
    lib.cb2_scene_info.restype = C.c_int64
    lib.cb2_scene_profile.argtypes = [vp, C.c_int, c_double_p, c_int64_p]
    lib.cb2_beam_sample.argtypes = [vp, c_double_p, C.c_int64, c_double_p]
    lib.cb2_rt_create.argtypes = [C.POINTER(RTDesc), C.c_int, C.POINTER(vp)]
    lib.cb2_rt_destroy.argtypes = [vp]
    lib.cb2_rt_render_dense.argtypes = [vp, C.POINTER(Rays), c_double_p, C.c_int, C.POINTER(Stats)]
    lib.cb2_rt_render_csr.argtypes = [vp, C.POINTER(Rays), c_int64_p, c_int32_p, c_double_p, C.c_int64, C.POINTER(Stats)]
    lib.cb2_rt_render_csr_device.argtypes = [vp, C.POINTER(Rays), vp, vp, vp, C.c_int64, c_int64_p, vp, vp]
    lib.cb2_pinhole_rays_device.argtypes = [C.POINTER(PinholeDesc), C.POINTER(PrimitiveDesc), vp, C.c_int64, C.c_double, C.c_double,
                                            C.POINTER(Rays), vp]
    lib.cb2_observer0d_rays_device.argtypes = [C.POINTER(Observer0DDesc), C.c_int64, C.POINTER(PrimitiveDesc), C.POINTER(Rays), vp, vp]
    lib.cb2_observer0d_reduce_device.argtypes = [vp, C.c_int, vp, c_int64_p, c_double_p, C.c_int64, C.c_int32, vp, vp, vp]
    lib.cb2_measure_peak_fp64.argtypes = [C.c_int, c_double_p]
    lib.cb2_wall_create.argtypes = [C.POINTER(WallDesc), C.c_int, C.POINTER(vp)]
    lib.cb2_wall_destroy.argtypes = [vp]
    lib.cb2_wall_hit.argtypes = [vp, c_double_p, c_double_p, C.c_int64, c_double_p]
    lib.cb2_wall_clip_device.argtypes = [vp, C.POINTER(Rays), vp, vp]
    lib.cb2_sart_create.argtypes = [C.POINTER(SartDesc), C.c_int, C.POINTER(vp)]
    lib.cb2_sart_destroy.argtypes = [vp]
    lib.cb2_sart_set_laplacian.argtypes = [vp, vp, vp, vp, vp]
    lib.cb2_sart_solve.argtypes = [vp, c_double_p, C.c_int64, c_double_p, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                                   c_double_p, c_double_p, c_int32_p]
    lib.cb2_sart_info.argtypes = [vp, C.c_int]
    lib.cb2_sart_info.restype = C.c_double
    for name in PRODUCT_SYMBOLS:
        getattr(lib, name)  # AttributeError if the .so does not export what the header declares
    if lib.cb2_abi_version() != ABI_VERSION:
        raise RuntimeError("core_b200: ABI version mismatch between Python and libcherab_b200.so")
    _lib = lib
    return lib


def check(lib, rc, errfn="cb2_last_error"):
    if rc != 0:
        msg = getattr(lib, errfn)()
        msg = msg.decode() if msg else "error %d" % rc
        raise STATUS_EXC.get(rc, RuntimeError)(msg)
