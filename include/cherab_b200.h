/*
 * cherab_b200.h — C ABI of the B200-native Cherab hot path.
 *
 * One data-parallel path: the per-ray line-of-sight integration of plasma
 * emission (Raysect NumericalIntegrator -> PlasmaMaterial.emission_function ->
 * PlasmaModel.emission -> LineShapeModel.add_line) and the sibling ray-transfer
 * (geometry-matrix) path-length sampler.  There is no FFI in the reference: the
 * seams these entry points replace are the Cython virtual interfaces listed in
 * SURVEY.md section 8(b); every declaration below cites the reference interface
 * (file:line, relative to the cherab/core checkout) whose work it takes over.
 *
 * The SAME descriptor structs are consumed by two independent implementations:
 *   - libcherab_b200.so   (core_b200/csrc, CUDA sm_100a, fp32 math / fp64 accumulation) — the product
 *   - libcb2_oracle.so    (oracle/, plain C, fp64, scalar restatement of the reference) — test infrastructure
 * Entry points of the product are prefixed cb2_, those of the oracle cb2o_.
 *
 * Conventions
 *   - plain pointers and sizes only; all arrays C-contiguous, caller-owned, read-only for the call;
 *   - descriptor data is HOST memory and is copied at cb2_scene_create();
 *   - every function returns CB2_OK (0) or a negative cb2_status; cb2_last_error() gives the message
 *     (mirrors the reference's `except -1` / `except? -1e999` Cython conventions and the Python
 *     exception type it would raise: see cb2_status);
 *   - a scene handle is not thread-safe (one CUDA stream per call); one handle per GPU.
 */
#ifndef CHERAB_B200_H
#define CHERAB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB2_ABI_VERSION 7

/* ------------------------------------------------------------------------------------------------
 * status codes.  Python shim maps them to the exception the reference raises at the same point.
 * ---------------------------------------------------------------------------------------------- */
typedef enum cb2_status {
    CB2_OK = 0,
    CB2_ERR_VALUE = -1,        /* ValueError: bad argument / out-of-range interpolation (interpolators 'none' extrapolation) */
    CB2_ERR_RUNTIME = -2,      /* RuntimeError: missing species / rates (impact_excitation.pyx:105-118) */
    CB2_ERR_TYPE = -3,         /* TypeError: unsupported model / lineshape / field kind (impact_excitation.pyx:60-61) */
    CB2_ERR_NOT_IMPLEMENTED = -4,
    CB2_ERR_CUDA = -5,         /* CUDA runtime failure (no CPU fallback exists) */
    CB2_ERR_MEMORY = -6,
    CB2_ERR_OVERFLOW = -7      /* a caller-provided output buffer is too small */
} cb2_status;

/* ------------------------------------------------------------------------------------------------
 * Spectral grid — raysect Spectrum(min_wavelength, max_wavelength, bins); delta=(max-min)/bins.
 * Used as in cherab/core/model/lineshape/gaussian.pyx:65-85.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cb2_spectral_grid {
    double  min_wavelength;   /* nm */
    double  max_wavelength;   /* nm */
    int32_t bins;
    int32_t _pad;
} cb2_spectral_grid;

/* ------------------------------------------------------------------------------------------------
 * Scalar / vector fields: the flattened Function3D trees behind
 * Maxwellian.density / effective_temperature / bulk_velocity (cherab/core/distribution.pyx:254-301).
 * ---------------------------------------------------------------------------------------------- */
typedef enum cb2_field_kind {
    CB2_FIELD_CONSTANT        = 0, /* c[0]                                   (Constant3D; build_constant_slab_plasma, tools/plasmas/slab.pyx:198-260) */
    CB2_FIELD_GAUSSIAN_VOLUME = 1, /* c[0] + c[1]*exp(-|p-(c[3],c[4],c[5])|^2/(2 c[2]^2))  (tools/plasmas/gaussian_volume.pyx:24-59, demos/balmer_series.py:55-60) */
    CB2_FIELD_AXISYM_BLEND    = 2, /* AxisymmetricMapper(Blend2D(edge mesh, equilibrium.map2d(core 1-D cubic), mask)) — generomak/plasma/plasma.py:580-638 */
    CB2_FIELD_SLAB_ION        = 3, /* IonFunction along +x (tools/plasmas/slab.pyx:63-110): c = t_core, t_lcfs, p, q, pedestal_top */
    CB2_FIELD_SLAB_NEUTRAL    = 4  /* NeutralFunction along +x (tools/plasmas/slab.pyx:20-60): c = peak, sigma */
} cb2_field_kind;

typedef struct cb2_scalar_field {
    int32_t       kind;      /* cb2_field_kind */
    int32_t       _pad;
    double        c[8];
    const double* edge;      /* AXISYM_BLEND: per-triangle values [n_triangles] (Discrete2DMesh data), NULL -> 0 */
    const double* core;      /* AXISYM_BLEND: values on the core psi_n grid [n_core] (Interpolator1DArray 'cubic','nearest'), NULL -> 0 */
} cb2_scalar_field;

typedef struct cb2_vector_field {
    int32_t       kind;      /* CONSTANT: cartesian (c[0],c[1],c[2]); AXISYM_BLEND: edge vector (R,phi,Z)=(c[0],c[1],c[2]) */
    int32_t       _pad;
    double        c[8];
    const double* core_vtor; /* AXISYM_BLEND: flux-function velocities on the core psi_n grid [n_core] (efit.pyx:521-546) */
    const double* core_vpol;
    const double* core_vnorm;
} cb2_vector_field;

/* EFITEquilibrium inputs (cherab/tools/equilibrium/efit.pyx:92-142) */
typedef struct cb2_equilibrium {
    int32_t       nr, nz;
    const double* r;               /* [nr] */
    const double* z;               /* [nz] */
    const double* psi;             /* [nr][nz] poloidal flux */
    double        psi_axis, psi_lcfs;
    int32_t       n_f;             /* f_profile points */
    int32_t       n_lcfs;          /* LCFS polygon vertices */
    const double* f_psin;          /* [n_f] */
    const double* f_value;         /* [n_f] */
    const double* lcfs_polygon;    /* [n_lcfs][2] (R,Z) */
    double        b_vacuum_radius, b_vacuum_magnitude;
} cb2_equilibrium;

/* Shared context of all AXISYM_BLEND fields of a scene (generomak/plasma/plasma.py:96-129,233-272,610) */
typedef struct cb2_axisym {
    cb2_equilibrium eq;
    int32_t         n_vertices, n_triangles;
    const double*   vertices;     /* [n_vertices][2] (R,Z) */
    const int32_t*  triangles;    /* [n_triangles][3] */
    int32_t         n_core;       /* core psi_n grid */
    int32_t         n_mask;       /* blend-mask table (Interpolator1DArray 'linear' of psi_n) */
    const double*   core_psin;    /* [n_core] */
    const double*   mask_x;       /* [n_mask] */
    const double*   mask_y;       /* [n_mask] */
} cb2_axisym;

/* Species(element, charge, Maxwellian(n, T, v, mass)) — cherab/core/species.pyx:26-80 */
typedef struct cb2_species {
    int32_t          charge;
    int32_t          _pad;
    double           atomic_weight;   /* element.atomic_weight (amu) */
    cb2_scalar_field density;
    cb2_scalar_field temperature;
    cb2_vector_field velocity;
} cb2_species;

/* ------------------------------------------------------------------------------------------------
 * Rates in ADAS/OpenADAS repository shape.
 * ---------------------------------------------------------------------------------------------- */
/* ImpactExcitationPEC / RecombinationPEC data dict {'ne','te','rate'} (cherab/openadas/rates/pec.pyx:48-77,111-140).
 * n_ne == 0 -> constant rate `constant` in W m^3 (mock AtomicData of core/tests/test_line_emission.py:32-88). */
typedef struct cb2_rate2d {
    int32_t       n_ne, n_te;
    const double* ne;          /* [n_ne] m^-3 */
    const double* te;          /* [n_te] eV */
    const double* rate;        /* [n_ne][n_te] photon m^3 s^-1 (converted with PhotonToJ at scene build, conversion.py:44-52) */
    double        constant;
    int32_t       extrapolate; /* 1: 'nearest' (permit_extrapolation=True), 0: 'none' -> out-of-domain samples are counted and clamped */
    int32_t       _pad;
} cb2_rate2d;

/* ThermalCXPEC data dict {'ne','te','td','rate'} (cherab/openadas/rates/pec.pyx:153-192): cubic in
 * (log10 ne, log10 te, log10 td).  n_ne == 0 -> constant rate in W m^3 (mock AtomicData of
 * core/tests/test_line_emission.py:58-84).  Tabulated rates need at least 2 knots per axis; outside the table the value is
 * clamped to the edge ('nearest') and, when extrapolate == 0, the sample is counted in stats.out_of_domain. */
typedef struct cb2_rate3d {
    int32_t       n_ne, n_te, n_td, _pad;
    const double* ne;
    const double* te;
    const double* td;          /* [n_td] donor temperature eV */
    const double* rate;        /* [n_ne][n_te][n_td] photon m^3 s^-1 */
    double        constant;
    int32_t       extrapolate;
    int32_t       _pad2;
} cb2_rate3d;

/* BeamStoppingRate data dict {'e','n','t','sen','st','sref'} (cherab/openadas/rates/beam.pyx:40-103):
 * rate = 10 ** (cubic2d[log10 E, log10 n](log10 sen) + cubic1d[log10 T](log10(st/sref))).  n_e == 0 -> constant rate
 * (the mock AtomicData of core/tests/test_beam.py:33-52).  Outside the tabulated range: with `extrapolate` the reference's
 * 'linear' / 'quadratic' extrapolation; without it ('none': the reference raises ValueError) the CUDA path and the oracle clamp
 * to the edge and count the lookup as out of domain. */
typedef struct cb2_beam_rate {
    int32_t       n_e, n_n, n_t;
    int32_t       extrapolate;  /* 1: 'linear' (2-D part) / 'quadratic' (1-D parts) extrapolation, beam.pyx:73-84; 0: 'none' */
    const double* e;           /* [n_e] interaction energy eV/amu */
    const double* n;           /* [n_n] target equivalent electron density m^-3 */
    const double* t;           /* [n_t] target temperature eV */
    const double* sen;         /* [n_e][n_n] m^3 s^-1 */
    const double* st;          /* [n_t] m^3 s^-1 */
    double        sref;
    double        constant;
} cb2_beam_rate;

/* BeamCXPEC data dict {'eb','ti','ni','z','b','qeb','qti','qni','qz','qb','qref'} (cherab/openadas/rates/cx.pyx:66-142):
 * rate = 10**cubic[log10 E](log10 PhotonToJ(qeb)) * cubic[Ti](qti/qref) * cubic[ni](qni/qref) * cubic[Zeff](qz/qref) *
 * cubic[B](qb/qref), zero as soon as a partial product is <= 0.  n_eb == 0 -> constant rate in W m^3
 * (core/tests/test_beamcxline.py:34-46).  A grid with a single point is a constant factor (Constant1D). */
typedef struct cb2_cx_rate {
    int32_t       n_eb, n_ti, n_ni, n_z, n_b;
    int32_t       extrapolate;  /* 1: 'quadratic' in log10 E, 'nearest' for the four factors (cx.pyx:96-102); 0: 'none' */
    const double *eb, *ti, *ni, *z, *b;
    const double *qeb, *qti, *qni, *qz, *qb;
    double        qref;
    double        constant;
} cb2_cx_rate;

/* Beam node + SingleRayAttenuator (cherab/core/beam/node.pyx:100-583, model/attenuator/singleray.pyx:36-346).
 * A scene with a beam renders BEAM models only; its ray segments are the chords through the beam's bounding primitive and
 * its integrator settings (step, min_samples) are the beam's (beam/node.pyx:212). */
typedef struct cb2_beam_desc {
    double  beam_to_plasma[12];  /* row-major 3x4 affine: beam frame (z along the beam, origin at the source) -> plasma space
                                    (beam.to(plasma), beam/material.pyx:62).  In a beam scene cb2_scene_desc.world_to_plasma
                                    holds world -> BEAM frame: the integrator marches in the beam primitive's local space */
    double  energy;              /* eV/amu */
    double  power;               /* W */
    double  temperature;         /* eV (line width of beam emission; unused by BeamCXLine) */
    double  atomic_weight;       /* beam.element.atomic_weight */
    double  sigma;               /* Gaussian width at the origin, m */
    double  divergence_x, divergence_y;   /* degrees */
    double  length;              /* m */
    double  attenuator_step;     /* SingleRayAttenuator.step (default 0.01 m) */
    double  clamp_sigma;         /* default 5 */
    int32_t clamp_to_zero;
    int32_t n_stopping;          /* species the beam is stopped by: the whole composition (singleray.pyx:337-340) */
    const int32_t*       stopping_species;   /* [n_stopping] */
    const cb2_beam_rate* stopping_rates;     /* [n_stopping] beam_stopping_rate(beam.element, species.element, species.charge) */
} cb2_beam_desc;

/* Free-free Gaunt factor table (cherab/core/atomic/gaunt.pyx:87-140; data/maxwellian_free_free_gaunt_factor.json) */
typedef struct cb2_gaunt {
    int32_t       n_u, n_gamma2;
    const double* u;           /* [n_u] */
    const double* gamma2;      /* [n_gamma2] */
    const double* gaunt;       /* [n_u][n_gamma2] */
} cb2_gaunt;

/* ------------------------------------------------------------------------------------------------
 * Line shapes (cherab/core/model/lineshape/ *.pyx)
 * ---------------------------------------------------------------------------------------------- */
typedef enum cb2_lineshape_kind {
    CB2_SHAPE_GAUSSIAN          = 0, /* GaussianLine            gaussian.pyx:122-139 */
    CB2_SHAPE_MULTIPLET         = 1, /* MultipletLineShape      multiplet.pyx:93-117 */
    CB2_SHAPE_ZEEMAN_TRIPLET    = 2, /* ZeemanTriplet           zeeman.pyx:113-162 */
    CB2_SHAPE_PARAM_ZEEMAN      = 3, /* ParametrisedZeemanTriplet zeeman.pyx:219-268 */
    CB2_SHAPE_ZEEMAN_MULTIPLET  = 4, /* ZeemanMultiplet         zeeman.pyx:308-365 */
    CB2_SHAPE_STARK             = 5  /* StarkBroadenedLine      stark.pyx:251-348 */
} cb2_lineshape_kind;

typedef enum cb2_polarisation { CB2_POL_PI = 0, CB2_POL_SIGMA = 1, CB2_POL_NO = 2 } cb2_polarisation; /* zeeman.pyx:35-39 */

typedef struct cb2_lineshape {
    int32_t       kind;           /* cb2_lineshape_kind */
    int32_t       polarisation;   /* cb2_polarisation (Zeeman family, Stark) */
    double        param[3];       /* PARAM_ZEEMAN: alpha,beta,gamma; STARK: c_ij,a_ij,b_ij (interface.pyx:154-196) */
    int32_t       n_components;   /* MULTIPLET: number of lines */
    int32_t       n_b;            /* ZEEMAN_MULTIPLET: points of the |B| grid the component functions are tabulated on */
    const double* multiplet;      /* MULTIPLET: [2][n_components] wavelengths then ratios (multiplet.pyx:75-88) */
    /* ZEEMAN_MULTIPLET (atomic/zeeman.pyx:87-129): n_pi + n_sigma_plus + n_sigma_minus component functions, each
       tabulated (linear, clamped) on b_grid: zeeman_wavelength[comp][n_b], zeeman_ratio[comp][n_b] */
    int32_t       n_pi, n_sigma_plus, n_sigma_minus, _pad;
    const double* b_grid;
    const double* zeeman_wavelength;
    const double* zeeman_ratio;
} cb2_lineshape;

/* ------------------------------------------------------------------------------------------------
 * Emission models (cherab/core/model/plasma/ *.pyx)
 * ---------------------------------------------------------------------------------------------- */
typedef enum cb2_model_kind {
    CB2_MODEL_EXCITATION_LINE    = 0, /* ExcitationLine.emission     impact_excitation.pyx:78-100 */
    CB2_MODEL_RECOMBINATION_LINE = 1, /* RecombinationLine.emission  recombination.pyx:78-100 */
    CB2_MODEL_BREMSSTRAHLUNG     = 2, /* Bremsstrahlung.emission     bremsstrahlung.pyx:169-208 */
    CB2_MODEL_THERMAL_CX_LINE    = 3, /* ThermalCXLine.emission      thermal_cx.pyx:79-112 */
    CB2_MODEL_TOTAL_RADIATED_POWER = 4, /* TotalRadiatedPower.emission total_radiated_power.pyx:70-118 */
    CB2_MODEL_BEAM_CX_LINE       = 5, /* BeamCXLine.emission         model/beam/charge_exchange.pyx:117-167 (needs cb2_scene_desc.beam) */
    CB2_MODEL_BEAM_EMISSION_LINE = 6  /* BeamEmissionLine.emission   model/beam/beam_emission.pyx:100-176 + BeamEmissionMultiplet.add_line
                                         (model/lineshape/beam/mse.pyx:62-135); cb2_model.shape is ignored */
} cb2_model_kind;

/* What ThermalCXLine._populate_cache (thermal_cx.pyx:114-155) and TotalRadiatedPower._populate_cache
 * (total_radiated_power.pyx:120-163) resolve. */
typedef struct cb2_model_ext {
    /* THERMAL_CX_LINE: every species except the receiver that is not fully ionised donates (thermal_cx.pyx:142-148) */
    int32_t           n_donors, _pad;
    const int32_t*    donor_species;     /* [n_donors] indices into the scene species */
    const cb2_rate3d* donor_rates;       /* [n_donors] thermal_cx_pec(donor, receiver, transition) */
    /* TOTAL_RADIATED_POWER: species (element, charge) radiates lines, (element, charge+1) recombines / exchanges charge
     * with the neutral hydrogen isotopes present; a missing rate is flagged by has_* = 0 */
    int32_t           line_rad_species, recom_species;
    int32_t           n_hydrogen, has_plt, has_prb, has_prc;
    const int32_t*    hydrogen_species;  /* [n_hydrogen] */
    cb2_rate2d        plt, prb, prc;     /* W m^3 on (ne, te), log-log cubic (openadas/rates/radiated_power.pyx:48-76); NOT photon rates */
    /* BEAM_CX_LINE: beam_cx_pec(beam.element, line.element, line.charge + 1, transition) — one rate per donor metastable,
     * cx[0] the ground state (metastable 1), cx[1..] the excited states (charge_exchange.pyx:311-349; at most 4 in all).
     * cx_population[(k-1) * n_species + s] = beam_population_rate(beam.element, metastable of cx[k], species s): the
     * BeamPopulationRate of every plasma species, composition order (same table shape as the stopping rate, dimensionless:
     * openadas/rates/beam.pyx:105-189); NULL when n_cx == 1.  Neutral species are skipped in the population average (the
     * reference divides by their zero charge there, charge_exchange.pyx:286). */
    int32_t            n_cx, _pad3;
    const cb2_cx_rate* cx;
    const cb2_beam_rate* cx_population;
    /* BEAM_EMISSION_LINE: beam_emission_pec(beam.element, species.element, species.charge, transition) for every plasma
     * species (beam_emission.pyx:207-212; same table shape as the stopping rate, 'sen' in photon m^3 s^-1, a constant rate in
     * W m^3) and the MSE intensity ratios sigma_to_pi, sigma1_to_sigma0, pi2_to_pi3, pi4_to_pi3 (beam_emission.pyx:47-50): constants
     * in mse_ratios, or — the reference also takes functions of the electron density (and, for sigma_to_pi, of the beam energy, which
     * is one number per scene; mse.pyx:103-121) — tabulated by the host on n_mse knots uniform in log10(ne [m^-3]) from mse_lne0 in
     * steps of mse_dlne, mse_ratio_tab[4][n_mse], interpolated linearly in log10(ne) and clamped at the ends (n_mse == 0: constants) */
    int32_t              n_bes, _pad4;
    const int32_t*       bes_species;
    const cb2_beam_rate* bes_rates;
    double               mse_ratios[4];
    int32_t              n_mse, _pad5;
    double               mse_lne0, mse_dlne;
    const double*        mse_ratio_tab;
} cb2_model_ext;

typedef struct cb2_model {
    int32_t       kind;            /* cb2_model_kind */
    int32_t       species;         /* index into scene species: density n_target AND lineshape target species
                                      (excitation: (element,charge); recombination and thermal CX: (element,charge+1) —
                                      recombination.pyx:113-121, thermal_cx.pyx:129-136); -1 for the continuum models */
    double        wavelength;      /* rest wavelength nm (atomic_data.wavelength) */
    double        atomic_weight;   /* line.element.atomic_weight (gaussian.pyx:137 uses the LINE's element) */
    cb2_rate2d    pec;
    cb2_lineshape shape;
    const cb2_model_ext* ext;      /* THERMAL_CX_LINE / TOTAL_RADIATED_POWER, else NULL */
} cb2_model;

/* ------------------------------------------------------------------------------------------------
 * Scene = Plasma node + models + integrator (cherab/core/plasma/node.pyx:201-554, material.pyx:25-63)
 * ---------------------------------------------------------------------------------------------- */
typedef struct cb2_scene_desc {
    int32_t            abi_version;         /* CB2_ABI_VERSION */
    int32_t            n_species;
    int32_t            n_models;
    int32_t            min_samples;         /* NumericalIntegrator(step, min_samples=5) [raysect]; plasma/node.pyx:318 */
    double             step;                /* m; default 0.001 */
    cb2_spectral_grid  grid;
    double             world_to_plasma[12]; /* row-major 3x4 affine: primitive/world -> plasma space (material.pyx:55-57) */
    cb2_scalar_field   electron_density;
    cb2_scalar_field   electron_temperature;
    const cb2_species* species;             /* [n_species] plasma.composition, in composition order */
    const cb2_model*   models;              /* [n_models]  plasma.models, in order */
    const cb2_axisym*  axisym;              /* NULL unless some field is AXISYM_BLEND */
    int32_t            b_field_kind;        /* 0: constant b_field[3]; 1: EFIT MagneticField of axisym->eq (efit.pyx:437-461) */
    int32_t            brems_quadrature;    /* CUDA path: Gauss-Legendre points per bin for Bremsstrahlung (0 = choose from error bound) */
    double             b_field[3];
    cb2_gaunt          gaunt;               /* needed iff a BREMSSTRAHLUNG model is present */
    /* oracle-only knobs of GaussianQuadrature (integrators1d.pyx:73-92): relative_tolerance, max_order, min_order */
    double             quad_rtol;
    int32_t            quad_max_order, quad_min_order;
    const cb2_beam_desc* beam;              /* NULL unless the models are BEAM models */
} cb2_scene_desc;

/* Ray segments: what Raysect's tracer hands to VolumeIntegrator.integrate(start_point, end_point)
 * (SURVEY 8(b) seam S1).  Ray r owns segments [seg_offset[r], seg_offset[r+1]); a segment is the
 * interval [t0, t1] along origin + t*direction (direction normalised), in world space. */
typedef struct cb2_rays {
    int64_t        n_rays;
    int64_t        n_segments;
    const double*  origin;      /* [n_rays][3] */
    const double*  direction;   /* [n_rays][3] unit vectors, observer -> scene */
    const int64_t* seg_offset;  /* [n_rays+1] */
    const double*  seg_t0;      /* [n_segments] */
    const double*  seg_t1;      /* [n_segments] */
} cb2_rays;

/* work counters: exactly the units of SURVEY 8(d) */
typedef struct cb2_stats {
    int64_t samples;              /* emission_function evaluations: sum over segments of intervals+1 */
    int64_t gaussian_bin_evals;   /* E: sum over (sample, Gaussian component) of (end-start)+1 */
    int64_t lorentzian_bin_evals; /* L */
    int64_t brems_bin_evals;      /* samples with ne,te>0 times bins */
    int64_t rt_steps;             /* ray-transfer midpoint steps */
    int64_t out_of_domain;        /* samples clamped where the reference would raise ValueError ('none' extrapolation) */
} cb2_stats;

typedef struct cb2_scene cb2_scene; /* opaque, owns device tables */

/* Library identification / error channel */
int         cb2_abi_version(void);
const char* cb2_last_error(void);
/* number of CUDA devices visible, <0 on error (never falls back to the CPU) */
int         cb2_device_count(void);

/* Roofline denominators measured live (FFMA and MUFU.EX2 issue-rate microbenchmarks): TFLOP/s, Tops/s, nominal SM MHz. */
int         cb2_measure_peaks(int device, double* fp32_tflops, double* sfu_tops, double* sm_clock_mhz);
/* FP64 FMA issue rate, TFLOP/s (denominator of the ray-transfer kernel's float64 index arithmetic). */
int         cb2_measure_peak_fp64(int device, double* fp64_tflops);

/* Build device-resident tables from a flattened scene.  Replaces PlasmaMaterial.__init__ + the lazy
 * _populate_cache of every model (plasma/material.pyx:37-46; impact_excitation.pyx:102-128). */
int cb2_scene_create(const cb2_scene_desc* desc, int device, cb2_scene** out);
int cb2_scene_destroy(cb2_scene* scene);

/* Emission render, HOST buffers (the reference-facing call): for every ray, trapezium-integrate the emission of all
 * models over each segment exactly as NumericalIntegrator.integrate [raysect] does
 * (intervals = max(min_samples-1, ceil(L/step)), samples at t0 + k*L/intervals), and write
 *     out[r][bin] = (accumulate ? out[r][bin] : 0) + scale * spectrum_r[bin]     W/m^2/sr/nm
 * out is double[n_rays][bins] if out_f64 else float[n_rays][bins].  Host<->device copies happen inside. */
int cb2_emission_render(cb2_scene* scene, const cb2_rays* rays, void* out, int out_f64,
                        double scale, int accumulate, cb2_stats* stats);

/* Same, with a destination row per ray: ray i's spectrum goes to row dest_row[i] of the HOST frame `out` (not accumulated).
 * This is how the tiles of an image reach their pixels: the rays of a rank are listed tile by tile (16 x 16 pixels), the frame is
 * pixel-ordered, so 16 consecutive rays are 16 consecutive rows — the library turns dest_row into runs and issues one strided
 * device -> host copy per group of equally spaced runs, overlapped with the next batch's kernels.  With N ranks every rank passes
 * the same frame (shared host memory) and its own rows; nothing has to be permuted afterwards. */
int cb2_emission_render_rows(cb2_scene* scene, const cb2_rays* rays, const int64_t* dest_row, void* out, int out_f64,
                             double scale, cb2_stats* stats);

/* The copy plan cb2_emission_render_rows follows for a destination-row list (no device needed; introspection / tests):
 * plan[4k..4k+3] = (first ray, run length, repeats, pitch in rows) of copy k, up to `capacity` copies; returns the number of copies. */
int64_t cb2_rows_plan(const int64_t* dest_row, int64_t n, int64_t* plan, int64_t capacity);

/* Same, DEVICE buffers (torch tensors): every pointer in `rays` and `out` is device memory on the scene's device;
 * launches on `stream` (a cudaStream_t passed as void*), does not synchronise. stats may be NULL;
 * if not NULL it must be device memory (filled asynchronously). */
int cb2_emission_render_device(cb2_scene* scene, const cb2_rays* rays, void* out, int out_f64,
                               double scale, int accumulate, cb2_stats* stats_dev, void* stream);

/* Per-sample plasma state, for parity tests of the flattened function tree:
 * points[n][3] (world space) -> out[n][n_quantities] with quantity order
 * ne, te, then per species (density, temperature, vx, vy, vz), then Bx, By, Bz.  HOST buffers. */
int cb2_sample_state(cb2_scene* scene, const double* points, int64_t n, double* out);
int cb2_state_width(const cb2_scene* scene);
/* Beam.density / Beam.direction at n points given in BEAM coordinates (beam/node.pyx:214-279): out[n][4] =
 * (density m^-3, direction x, y, z in the beam frame).  Parity probe for the reference's test_beam.py. */
int cb2_beam_sample(cb2_scene* scene, const double* beam_points, int64_t n, double* out);

/* Launch-plan introspection (no reference counterpart; used by the benchmark and the parity tests to report which
 * formulation a scene runs).  key: 0 warps per CTA, 1 bins per lane, 2 Bremsstrahlung formulation (0 none, 1 direct
 * per-(sample, bin) evaluation, 3 per-ray temperature moments + contraction), 3 moment row length k_pad,
 * 4 temperature nodes, 5 distinct ion charges, 6 rays per batch of the two-kernel line path, 7 line path (1 two-kernel
 * state/bin path, 0 CTA-phased kernel with the direct Bremsstrahlung evaluation), 8 contraction of the moments (1 tensor
 * cores, error-compensated 3xTF32 GEMMs; 0 FFMA tile kernel).  Returns -1 for an unknown key. */
int64_t cb2_scene_info(const cb2_scene* scene, int key);

/* Per-kernel device timing for the benchmark's roofline section (no reference counterpart).  enable != 0: reset the
 * accumulators and bracket every kernel of the following emission renders with CUDA events on the launching stream
 * (adds one stream synchronisation per ray batch — not for production use); enable == 0: stop and, if ms_out is not
 * NULL, return the accumulated milliseconds: [0] state_kernel, [1] bin_kernel, [2] contract_kernel,
 * [3] count/scan helpers, and launches_out[0..3] the launch counts. */
int cb2_scene_profile(cb2_scene* scene, int enable, double* ms_out, int64_t* launches_out);

/* ------------------------------------------------------------------------------------------------
 * Ray transfer (cherab/tools/raytransfer/emitters.pyx:88-224, raytransfer.py:183-268)
 * ---------------------------------------------------------------------------------------------- */
typedef enum cb2_rt_kind { CB2_RT_CYLINDRICAL = 0, CB2_RT_CARTESIAN = 1 } cb2_rt_kind;

typedef struct cb2_rt_desc {
    int32_t        abi_version;
    int32_t        kind;              /* cb2_rt_kind */
    int32_t        grid_shape[3];     /* (n_r, n_phi, n_z) or (nx, ny, nz) — RayTransferEmitter.grid_shape emitters.pyx:261-266 */
    int32_t        min_samples;       /* RayTransferIntegrator(step, min_samples=2) emitters.pyx:48 */
    double         grid_steps[3];     /* (dr, dphi [deg], dz) or (dx, dy, dz) */
    double         rmin;              /* cylindrical: radius_inner */
    double         period;            /* cylindrical: degrees */
    double         step;              /* integration step m (raytransfer.py:195,262) */
    double         world_to_local[12];/* row-major 3x4 affine world -> primitive-local (z in [0,height]) */
    const int32_t* voxel_map;         /* [shape0][shape1][shape2], -1 = unmapped */
    int32_t        bins;              /* voxel_map.max()+1 */
    int32_t        integrator;        /* 0: the emitter's own RayTransferIntegrator (midpoint samples, emitters.pyx:88-224);
                                         1: a foreign NumericalIntegrator(step, min_samples) [raysect] over the emitter's
                                            emission_function (unit emissivity in the sample's cell, emitters.pyx:452-473,557-571):
                                            trapezium rule, intervals = max(min_samples - 1, ceil(L / step)) */
} cb2_rt_desc;

typedef struct cb2_rt_scene cb2_rt_scene;

int cb2_rt_create(const cb2_rt_desc* desc, int device, cb2_rt_scene** out);
int cb2_rt_destroy(cb2_rt_scene* scene);

/* Dense geometry-matrix rows, HOST buffers: out[r][source] (+)= path length (m) of ray r in light source `source`
 * (the Spectrum the reference integrator fills, emitters.pyx:137-150).  Only for small `bins`. */
int cb2_rt_render_dense(cb2_rt_scene* scene, const cb2_rays* rays, double* out, int accumulate, cb2_stats* stats);

/* Sparse (CSR) rows, HOST buffers.  row_offset[n_rays+1]; columns/lengths capacity `capacity` entries.
 * Within a row every source appears once, in order of first visit.  Returns CB2_ERR_OVERFLOW (and the required
 * capacity in row_offset[n_rays]) if capacity is too small. */
int cb2_rt_render_csr(cb2_rt_scene* scene, const cb2_rays* rays, int64_t* row_offset,
                      int32_t* columns, double* lengths, int64_t capacity, cb2_stats* stats);

/* DEVICE buffers variant: rays/out pointers are device memory; row_offset[n_rays+1] int64, columns int32,
 * lengths float64 on device; launches on `stream`; *nnz_host receives the total after an internal sync on the stream. */
int cb2_rt_render_csr_device(cb2_rt_scene* scene, const cb2_rays* rays, int64_t* row_offset,
                             int32_t* columns, double* lengths, int64_t capacity, int64_t* nnz_host,
                             cb2_stats* stats_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Observer front-end on the device: the rays of a pinhole camera and their chords through the primitive that
 * bounds the plasma / beam / ray-transfer grid, generated straight into DEVICE memory in the cb2_rays layout.  Replaces the
 * per-pixel Python loop of raysect's PinholeCamera / Observer2D.observe() [raysect, SURVEY Appendix B.9] in front of
 * cb2_emission_render_device and cb2_rt_render_csr_device.
 * ---------------------------------------------------------------------------------------------- */
typedef enum cb2_primitive_kind {
    CB2_PRIM_HOLLOW_CYLINDER = 0,   /* Subtract(Cylinder(r_outer), Cylinder(r_inner)): p = r_inner (0: solid), r_outer, z_min, z_max
                                       (generomak/plasma/plasma.py:673-681, raytransfer.py:198) */
    CB2_PRIM_SPHERE = 1,            /* p = radius */
    CB2_PRIM_BOX = 2                /* p = lower x, y, z, upper x, y, z (tools/plasmas/slab.pyx:255, raytransfer.py:266) */
} cb2_primitive_kind;

typedef struct cb2_primitive {
    int32_t kind, _pad;
    double  p[6];
    double  world_to_local[12];     /* row-major 3x4 affine: world -> primitive-local */
} cb2_primitive;

typedef struct cb2_pinhole {
    int32_t nx, ny;                 /* pixels; pixel p = ix * ny + iy */
    double  width;                  /* image-plane width at unit distance: 2 tan(fov / 2) */
    double  to_world[12];           /* row-major 3x4 affine: camera space (looking along +z) -> world */
} cb2_pinhole;

/* One ray per listed pixel (pixel_index_dev: DEVICE int64[n], or NULL for all nx * ny pixels in order) through the
 * sub-pixel position (sub_x, sub_y) in [0, 1)^2 (0.5 = pixel centre).  out_dev: caller-allocated DEVICE arrays
 * origin[n][3], direction[n][3], seg_offset[n+1], seg_t0[2n], seg_t1[2n]; n_rays and n_segments are filled in (one stream
 * synchronisation for the segment total). */
int cb2_pinhole_rays_device(const cb2_pinhole* camera, const cb2_primitive* primitive, const int64_t* pixel_index_dev, int64_t n,
                            double sub_x, double sub_y, cb2_rays* out_dev, void* stream);

/* 0-D observers (raysect SightLine / FibreOptic as cherab/tools/observers/group/{sightline,fibreoptic}.py:73-95,98-160 place them):
 * a sight line is one ray from the observer's origin along its +z axis; a fibre's `samples` rays start on the tip disc of `radius`
 * (sunflower points, uniform in area) and leave within `acceptance_angle` degrees of +z (Fibonacci points, uniform in solid angle
 * on the cone cap) — the deterministic stand-in for raysect's random ConeUniformSampler / DiskSampler that
 * core_b200/observers.py::FibreOptic.rays defines.  weight = cos(theta): the projected tip area a direction sees. */
typedef struct cb2_observer0d {
    double  to_world[12];           /* row-major 3x4 affine: observer space (looking along +z) -> world */
    double  radius;                 /* fibre tip radius, m; 0 = sight line */
    double  acceptance_angle;       /* degrees, (0, 90]; ignored by a sight line */
    int32_t samples;                /* rays of this observer (1 for a sight line) */
    int32_t _pad;
} cb2_observer0d;

/* Rays of n_observers 0-D observers (HOST array) and their chords through `primitive`, observer after observer, into
 * caller-allocated DEVICE arrays sized for n = sum(samples) rays (origin[n][3], direction[n][3], seg_offset[n+1], seg_t0[2n],
 * seg_t1[2n]); weight_dev[n] (DEVICE double) receives cos(theta) of every ray.  One stream synchronisation (segment total). */
int cb2_observer0d_rays_device(const cb2_observer0d* observers, int64_t n_observers, const cb2_primitive* primitive,
                               cb2_rays* out_dev, double* weight_dev, void* stream);

/* Per-observer reduction of per-ray spectra (what raysect's pixel loop + SpectralRadiancePipeline0D / SpectralPowerPipeline0D do,
 * cherab/tools/observers/group/base.py:130-135,399-436): observer i owns rays [ray_offset[i], ray_offset[i+1]) (HOST int64[n+1]);
 * radiance_dev[i][bin] = sum_r w_r S[r][bin] / sum_r w_r, power_dev[i][bin] = etendue[i] * sum_r w_r S[r][bin] / rays_i
 * (etendue: HOST double[n_observers]; either output may be NULL).  spectra_dev: [n_rays][bins] float32 / float64. */
int cb2_observer0d_reduce_device(const void* spectra_dev, int spectra_f64, const double* weight_dev, const int64_t* ray_offset,
                                 const double* etendue, int64_t n_observers, int32_t bins, double* radiance_dev, double* power_dev,
                                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * First-wall occlusion (replaces the role the wall meshes of cherab/generomak/machine/first_wall.py:120-184
 * play in Raysect's tracer: a ray ends at its first opaque hit, so the volume integral of the plasma stops there).
 * The wall is a soup of world-space triangles; a bounding-volume hierarchy is built on the host at create time and the
 * first-hit search (float32 boxes grown by their rounding, float64 Moeller-Trumbore triangle test) runs on the device.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cb2_wall_desc {
    int32_t       abi_version;
    int32_t       _pad;
    int64_t       n_triangles;
    const double* vertices;         /* HOST [n_triangles][3 vertices][3] world coordinates, m */
} cb2_wall_desc;

typedef struct cb2_wall cb2_wall;

int cb2_wall_create(const cb2_wall_desc* desc, int device, cb2_wall** out);
int cb2_wall_destroy(cb2_wall* wall);
/* t_hit[n] (HOST): distance, in units of |direction|, from origin[i] to the first triangle hit along direction[i] (t > 0),
 * +inf for a miss.  origin / direction: HOST [n][3]. */
int cb2_wall_hit(cb2_wall* wall, const double* origin, const double* direction, int64_t n, double* t_hit);
/* Clips DEVICE ray segments in place on `stream`: seg_t1 = max(seg_t0, min(seg_t1, t_hit(ray))) — a segment behind the hit
 * becomes empty and the marcher skips it (segment counts and offsets do not change).  t_hit_dev: optional DEVICE double[n_rays]. */
int cb2_wall_clip_device(cb2_wall* wall, const cb2_rays* rays_dev, double* t_hit_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SART inversion on the device-resident geometry matrix (replaces
 * cherab/tools/inversions/sart.pyx:26-155 invert_sart, :161-302 invert_constrained_sart and the OpenCL solver
 * cherab/tools/inversions/opencl/sart_opencl.py:33-318 + sart_kernels.cl:28-148).
 * The matrix is held as CSR and CSC on the device (what copy_column_major=True keeps, sart_opencl.py:120-125);
 * arithmetic is float64 (the CPU reference's), the stored matrix values float32 or float64.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cb2_sart_desc {
    int32_t        abi_version;
    int32_t        value_f64;         /* device storage of the matrix values: 1 = float64 (sart.pyx), 0 = float32 (sart_opencl.py:99-100) */
    int64_t        n_detectors;       /* N_d rows */
    int64_t        n_sources;         /* N_s columns */
    /* geometry matrix, ONE of: dense row-major [n_detectors][n_sources] ... */
    const void*    dense;             /* float32 (dense_f64 == 0) or float64 (dense_f64 == 1); NULL if CSR is given */
    int32_t        dense_f64;
    int32_t        memory;            /* 0: all pointers below and `dense` are HOST memory; 1: the CSR arrays are DEVICE memory
                                         (the output of cb2_rt_render_csr_device) and are copied device-to-device */
    /* ... or CSR as cb2_rt_render_csr produces it */
    const int64_t* row_offset;        /* [n_detectors + 1] */
    const int32_t* columns;           /* [nnz] */
    const double*  values;            /* [nnz] */
    /* optional Laplacian regularisation operator [n_sources][n_sources]: dense float64 row-major, or CSR; HOST memory */
    const double*  laplacian_dense;
    const int64_t* lap_row_offset;
    const int32_t* lap_columns;
    const double*  lap_values;
} cb2_sart_desc;

typedef struct cb2_sart cb2_sart;

int cb2_sart_create(const cb2_sart_desc* desc, int device, cb2_sart** out);
int cb2_sart_destroy(cb2_sart* solver);
/* Replaces the Laplacian (SartOpencl.update_laplacian_matrix, sart_opencl.py:186-196); same pointer rules as at create. */
int cb2_sart_set_laplacian(cb2_sart* solver, const double* dense, const int64_t* row_offset, const int32_t* columns, const double* values);

/* Inverts n_frames measurement vectors at once (HOST buffers): measurements[n_frames][n_detectors],
 * initial_guess[n_frames][n_sources] or NULL (then every cell starts at initial_value; the reference default is exp(-1),
 * sart.pyx:88-89), solution[n_frames][n_sources], convergence[n_frames][max_iterations] (entries past a frame's last
 * iteration are left untouched), n_iterations[n_frames].  Per frame this is exactly the reference loop: update every
 * cell from the previous estimate, clamp at zero, forward-project, record (|m|^2 - |y_hat|^2)/|m|^2 and stop once two
 * successive records differ by less than conv_tol (sart.pyx:116-153); beta_laplace is ignored without a Laplacian.
 * The matrix is streamed from HBM twice per iteration whatever n_frames is. */
int cb2_sart_solve(cb2_sart* solver, const double* measurements, int64_t n_frames, const double* initial_guess,
                   double initial_value, int max_iterations, double relaxation, double beta_laplace, double conv_tol,
                   double* solution, double* convergence, int32_t* n_iterations);
/* [0] nnz of the geometry matrix, [1] nnz of the Laplacian, [2] bytes of matrix values+indices streamed per iteration
 * (CSR pass + CSC pass), [3] device milliseconds spent in the iteration kernels of the last solve, [4] iterations
 * launched by the last solve (>= the largest n_iterations: the stop test is read back every few iterations) */
double cb2_sart_info(const cb2_sart* solver, int what);

#ifdef __cplusplus
}
#endif
#endif /* CHERAB_B200_H */
