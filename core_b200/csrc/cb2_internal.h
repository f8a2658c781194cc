// cb2_internal.h — device-resident scene layout shared by the host builder and the kernels.
//
// Data layout in HBM (all tables fp32, built once at cb2_scene_create from the fp64 descriptor):
//   * 2-D cubic tables (psi_n, dpsi/dR, dpsi/dZ, every PEC, Gaunt): per-cell 16 polynomial coefficients as 4 float4
//     (row i = coefficients of t^i * u^0..3), knots + reciprocal cell widths for the (rare) non-uniform search;
//   * 1-D cubic tables (core profiles on the psi_n grid, f-profile): per-interval float4 (a0..a3 in t in [0,1]);
//   * edge mesh: per-triangle vertex coordinates (6 floats) + uniform bucket grid (cell_start / cell_tris);
//   * densities are stored scaled by 1e-19 and PEC tables as log10(W m^3) + 38 so that every product stays in
//     fp32 range (ne*ni reaches 1e42 m^-6 on Generomak).
// A scene is a few MB: it lives in L2 (126 MB) for the whole frame; the hot per-sample records are in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cherab_b200.h"

#define CB2_MAX_SPECIES 24
#define CB2_MAX_MODELS 32
#define CB2_MAX_COMP 96
#define CB2_MAX_BREMS_Z 8
#define CB2_DENSITY_SCALE 1e-19
#define CB2_PEC_LOG_OFFSET 38.0

struct DevTable2D {
    int nx, ny;             // knots
    int uniform;            // 1: both axes uniform -> direct index
    float x0, inv_dx, y0, inv_dy;
    float xmin, xmax, ymin, ymax;
    const float* x;         // [nx]
    const float* y;         // [ny]
    const float* inv_wx;    // [nx-1] reciprocal cell widths
    const float* inv_wy;    // [ny-1]
    const float4* coef;     // [(nx-1)*(ny-1)*4]
};

struct DevTable3D {          // tricubic on (log10 ne, log10 te, log10 td): ThermalCXPEC (openadas/rates/pec.pyx:153-194)
    int nx, ny, nz;
    float xmin, xmax, ymin, ymax, zmin, zmax;
    const float* x;          // knots
    const float* y;
    const float* z;
    const float* inv_wx;     // reciprocal cell widths
    const float* inv_wy;
    const float* inv_wz;
    const float4* coef;      // [(nx-1)(ny-1)(nz-1)][16]: entry 4p+q holds the powers of w multiplying t^p u^q
};

struct DevTable1D {          // knots shared, coefficients per quantity
    int n;
    int uniform;             // 1: uniform knots -> direct index, no dependent loads
    float x0, inv_dx;
    float xmin, xmax;
    const float* x;          // [n]
    const float* inv_w;      // [n-1]
};

struct DevScalar {
    int kind;
    float c[6];
    const float* edge;       // [n_triangles]
    const float4* core;      // [n_core-1]
};

struct DevVector {
    int kind;
    float c[3];
    const float4* vtor;
    const float4* vpol;
    const float4* vnorm;
};

struct DevSpecies {
    int charge;
    float z2;
    DevScalar density, temperature;
    DevVector velocity;
};

struct DevAxisym {
    int present;
    DevTable2D psin, dpsi_dr, dpsi_dz;
    DevTable1D core;                 // core psi_n knots
    DevTable1D fprof;                // f-profile knots
    const float4* fprof_coef;
    float b_vac;                     // b_vacuum_magnitude * b_vacuum_radius
    int n_poly;
    const float4* poly;              // per edge: (xi, yi, yj, slope = (xj-xi)/(yj-yi))
    float poly_xmin, poly_xmax, poly_ymin, poly_ymax;
    // coarse classification grid over the polygon bounding box: 0 outside, 1 inside, 2 boundary cell (run the edge loop)
    int pgx, pgy;
    float p_icx, p_icy;
    const unsigned char* poly_cls;
    const int* poly_row_start;       // [pgy + 1] per grid row: the edges that can cross a horizontal ray starting in that row
    const float4* poly_row_edges;
    const double2* poly_row_edges_d; // the same edges in float64, (xi, yi), (xj, yj): exact test for samples within rounding of the boundary
    int n_mask;
    float mask_x[8], mask_y[8];
    // mesh
    int n_tri, gx, gy;
    float mx0, my0, inv_cx, inv_cy;
    const int* cell_start;
    const int* cell_tris;
    const double2* tri;              // [n_tri*3] fp64 vertices: the containment test must decide exactly like the fp64 reference
    const float4* trif;              // [n_tri*2] the same vertices rounded to fp32 (sign filter in front of the exact test)
    double mx0_d, my0_d, inv_cx_d, inv_cy_d;
};

// static part of one line component slot (Gaussian or Lorentzian)
struct DevComp {
    int c0_int;       // integer part of (lambda_component - lambda_min)/delta
    float c0_frac;    // fractional part
    float dlambda;    // static wavelength offset from the model's rest wavelength already included in c0 (0)
    int type;         // 0 gaussian, 1 lorentzian
};

#define CB2_MAX_META 4
// BeamCXPEC: rate = 10^c0(log10 E) c1(Ti) c2(n_ion) c3(Zeff) c4(|B|)   (openadas/rates/cx.pyx:104-142)
struct DevCXRate {
    int extrapolate;                       // 1: 'quadratic' in log10 E, 'nearest' for the factors (cx.pyx:96-102); 0: clamp and count
    int is_const;                          // 1: constant rate
    float lconst;                          // log10(rate [W m^3]) + 38
    int n[5];                              // knots per factor (1: constant factor `single`)
    float single[5];
    DevTable1D t[5];                       // knots: log10 E[eV/amu], Ti[eV], n_ion[1e19 m^-3], Zeff, |B|[T]
    const float4* c[5];
};
// BeamPopulationRate: 10^(A(log10 E, log10 n_eq) + B(log10 T)), dimensionless   (openadas/rates/beam.pyx:105-189)
struct DevPopRate {
    int extrapolate;                       // 1: 'linear' (2-D) / 'quadratic' (1-D) extrapolation (beam.pyx:73-84); 0: clamp and count
    int is_const;
    float lconst;                          // log10(value), -inf for a null rate
    DevTable2D a;                          // (log10 E[eV/amu], log10 n_eq[m^-3]) -> log10 sen
    DevTable1D tk;                         // log10 T[eV] knots
    const float4* tc;                      // -> log10(st / sref)
};

// ThermalCXLine donors / TotalRadiatedPower species and rates (device memory, referenced from DevModel::ext)
struct DevModelExt {
    int n_donors;
    int donor_species[CB2_MAX_SPECIES];
    float donor_lrate[CB2_MAX_SPECIES];   // log10(rate [W m^3]) + 38, constant rates
    int donor_tab[CB2_MAX_SPECIES];       // 1: tabulated rate in donor_t3 (log10 + 38), 0: constant
    int donor_extrapolate[CB2_MAX_SPECIES];
    DevTable3D donor_t3[CB2_MAX_SPECIES];
    int line_rad, recom, n_hyd, hyd[3];
    // BEAM_CX_LINE: one effective emission coefficient per donor metastable, cx[0] the ground state, and the population of
    // every excited metastable relative to the ground state for each plasma species (charge_exchange.pyx:204-292)
    int n_cx;
    DevCXRate cx[CB2_MAX_META];
    const DevPopRate* pop;                 // [(n_cx - 1) * n_species]
    // BEAM_EMISSION_LINE: rate = sum_s (n_s Z_s) 10^(A_s(log10 E, log10 n_eq) + B_s(log10 T_s))   (beam_emission.pyx:131-176)
    int n_bes;
    int bes_species[CB2_MAX_SPECIES];
    int bes_charge[CB2_MAX_SPECIES];
    int bes_const[CB2_MAX_SPECIES];
    int bes_extrapolate[CB2_MAX_SPECIES];
    float bes_lconst[CB2_MAX_SPECIES];     // log10(rate [W m^3]) + 38
    DevTable2D bes_a[CB2_MAX_SPECIES];     // (log10 E[eV/amu], log10 n_eq[m^-3]) -> log10(sen [W m^3]) + 38
    DevTable1D bes_tk[CB2_MAX_SPECIES];    // log10 T[eV] knots
    const float4* bes_tc[CB2_MAX_SPECIES]; // -> log10(st / sref)
    float mse_amp[9];                      // relative intensities of the 9 multiplet components (mse.pyx:105-133), constant ratios
    int mse_n;                             // > 1: ratios tabulated on mse_n knots uniform in log10(ne), mse_tab[knot] = the four ratios
    float mse_lne0, mse_inv_dlne;
    const float4* mse_tab;
    float mse_sigma_b;                     // sigma in bins from the beam temperature and mass
    int has[3];                            // plt, prb, prc present
    int is_const[3];
    float lconst[3];                       // log10(rate) + 38
    int extrapolate[3];
    DevTable2D tab[3];                     // log10 ne[m^-3], log10 te -> log10(W m^3) + 38
};

struct DevModel {
    int kind, species, shape, polarisation;
    float wavelength;     // rest wavelength (nm)
    float sigma_coef;     // sigma[bins] = sigma_coef * sqrt(T[eV])
    float inv_delta;      // 1/delta_wavelength
    float inv_c;          // 1/SPEED_OF_LIGHT
    float param[3];
    int comp0, ncomp;     // component slots
    int pec_const;        // 1: constant rate
    float pec_value;      // log10(rate) + 38 for constant rates
    int pec_extrapolate;
    int pec_grid;         // models with the same id share the (ne, te) knots: the cell search is done once per sample
    DevTable2D pec;       // log10 ne[m^-3], log10 te -> log10(W m^3) + 38
    // multiplet
    int n_mult;
    const float* mult_ratio;      // [n_mult]
    const float* mult_lambda;     // [n_mult] component rest wavelengths
    // zeeman multiplet tables
    int n_b, n_pi, n_sp, n_sm;
    float b0, inv_db;             // uniform |B| grid
    const float* zee_dlambda;     // [ncomp][n_b]  lambda_j(B) - lambda0
    const float* zee_ratio;       // [ncomp][n_b]
    const DevModelExt* ext;       // THERMAL_CX_LINE / TOTAL_RADIATED_POWER
};

struct DevBrems {
    int present;
    int nq;                       // Gauss-Legendre points per bin
    const float4* bin_tab;        // [bins_padded][nq]: (1/lambda, 2*log2(1/lambda), log10(lambda) - lref, weight)
    const float4* bin_tab1;       // [bins_padded] one-point (bin centre) version of the same
    float mid_c1, mid_c0;         // one-point rule error model: h = |mid_c1/Te - mid_c0|, relative error ~ h^2/6
    float lref;                   // log10 of the window centre
    float log_hc;                 // log10(HC_EV_NM)
    float exp_coef;               // EXP_FACTOR * log2(e)
    float pref;                   // BREMS_CONST * 1e38
    float rho_min, rho_max;       // 1/lambda_max, 1/lambda_min
    float lp_min, lp_max;         // log10(lambda_min) - lref, log10(lambda_max) - lref
    // gaunt table in (log10 u, log10 gamma2)
    DevTable2D gaunt;
    float lu_min, lu_max, lg_min, lg_max;
    int n_charged;
    int charged[CB2_MAX_SPECIES];
    // ---- moment formulation (mode 3): the per-(sample, bin) work is replaced by per-ray moments on a
    // temperature-node grid followed by one dense contraction with a per-scene table.
    //   bin average of eps_s = W_s sum_z N_{z,s} Phi_z(Te_s; bin),  Phi_z(T; j) = Z^2 <g_ff(Z,T,l) l^-2 exp(-(hc/l - x_ref)/T)>_j
    //   Phi_z(Te_s; j) ~= sum_m L_m(s(Te_s)) Phi_z(T_m; j)   (4-point Lagrange on nodes uniform in s = ln(tau) + tau/tau_c, tau = 1/T)
    //   mom[z][m] = sum_s W_s N_{z,s} L_m(s_s),  W_s = weight pref ne/sqrt(Te) exp(-x_ref/Te);   spectrum = mom . Phi
    int mode;                     // 1/2: direct per-(sample, bin) path, 3: moments
    int n_z;                      // distinct charges (<= CB2_MAX_BREMS_Z)
    int zidx[CB2_MAX_SPECIES];    // charged[s] -> index into the distinct-charge list
    int zlist[CB2_MAX_SPECIES];   // scene species indices grouped by distinct charge: group z is zlist[zstart[z] .. zstart[z+1])
    int zstart[CB2_MAX_BREMS_Z + 1];
    int n_nodes;                  // temperature nodes M
    int k_pad;                    // n_z * M rounded up to a multiple of 16 (row length of the moment matrix)
    int n_pad;                    // bins rounded up to a multiple of 128 (row length of phi)
    float s0, inv_ds, inv_tau_c;  // node k sits at s = s0 + k ds
    float x_ref;                  // reference photon energy (eV) factored out of the table
    float te_lo, te_hi;           // temperatures covered by the node grid (samples outside are clamped and counted)
    const float* phi;             // [k_pad][n_pad]
};

// Beam + SingleRayAttenuator (beam/node.pyx, attenuator/singleray.pyx): in a beam scene the integrator marches in the beam
// frame (DevScene::w2p holds world -> beam) and l2p takes the sample to plasma space
struct DevBeam {
    int present;
    double l2p[12];                // beam -> plasma
    float speed;                   // sqrt(2 E e / m_u), m/s
    float length, sigma2, tanx, tany;
    int clamp_to_zero;
    float clamp2;                  // clamp_sigma^2
    int n_axis;                    // attenuation table on z = linspace(0, length, n_axis)
    float inv_dz;
    const float* axis_density;     // [n_axis] in units of 1e19 m^-3
};

struct alignas(16) DevScene {
    int n_species, n_models, n_comp;
    int bins, bins_padded;
    float min_wavelength, delta;
    double min_wavelength_d, delta_d;
    double step;
    int min_samples;
    double w2p[12];
    int need_b, need_pol;          // which expensive shared quantities any model needs
    int b_kind;
    float b_const[3];
    DevScalar ne, te;
    DevSpecies species[CB2_MAX_SPECIES];
    DevModel models[CB2_MAX_MODELS];
    DevComp comps[CB2_MAX_COMP];
    DevAxisym ax;
    DevBrems brems;
    DevBeam beam;
    // modified-Lorentzian (Stark) cumulative profile, universal in u = (x - centre)/FWHM (stark.pyx:52-81):
    // knots u_k = k/512 on [0, 4]: (Phi(u_k), dPhi/du(u_k)); beyond u = 4 an asymptotic tail series is used
    int has_flat;                  // some model adds a wavelength-independent radiance (TotalRadiatedPower)
    float inv_range;               // 1 / (max_wavelength - min_wavelength)
    int has_lorentz;
    const double2* lorentz_tab;
    double lorentz_phi_inf;
};

static_assert(sizeof(DevScene) % 16 == 0, "DevScene is copied to shared memory as uint4");

// ---- state tables ("memo"): in a scene whose scalar fields are all AXISYM_BLEND the plasma state at a sample is a
// function of the mesh triangle alone where the blend weight is 0 (edge) and of psi_n alone where it is 1 (core).  Everything the
// line records and the Bremsstrahlung moments need from the state is tabulated once per scene by the generic evaluation code:
// one row per triangle (exact) and one row per knot of a fine uniform psi_n grid (linear interpolation, refined until the
// mid-interval error against the generic evaluation is below 4e-6).  Samples in the blend zone take the generic kernel.
// Row layout (float4 units): [n_sp x (sqrt(Ts), vtor, vpol, vnorm)] [ceil(n_lines / 4) x amplitudes per unit weight]
// [(f, ood, U0, U1) (U2..U5) (U6, U7, -, -)]; on edge rows the velocity slots hold (v_phi, v_R, v_Z) of the constant edge vector.
#define CB2_MEMO_MAX_LINES 16
#define CB2_MEMO_MAX_SP 4
struct MemoLine {
    int model, slot, comp0, ncomp, shape;
    int rec_off;                             // comp0 * floats per component record
    float c0_frac;                           // fractional bin position of the (first) component's rest wavelength
    float shift_coef;                        // wavelength / (c delta): Doppler shift in bins per m/s of v.d
    float sigma_coef, wavelength, inv_c, inv_delta;
    const float* mult_ratio;
    const float* mult_lambda;
};
struct alignas(16) DevMemo {
    int enabled;
    int n_lines, n_sp, n_z, has_brems;
    int row_f4, off_amp, off_brems;          // float4 per row, float4 offsets of the amplitude and Bremsstrahlung blocks
    int sp_species[CB2_MEMO_MAX_SP];
    int sp_const[CB2_MEMO_MAX_SP];           // 1: constant cartesian velocity (sp_v), nothing in the row
    float sp_v[CB2_MEMO_MAX_SP][3];
    int core_n;                              // intervals of the psi_n grid (core_n + 1 rows); 0: no core table
    float core_scale;                        // core_n / psi_max
    float psi_max;
    const float4* core;
    const float4* edge;                      // [n_tri] rows
    MemoLine lines[CB2_MEMO_MAX_LINES];
};

struct DevRays {
    int64_t n_rays;
    const double* origin;
    const double* direction;
    const int64_t* seg_offset;
    const double* seg_t0;
    const double* seg_t1;
};

// ---- ray transfer ----
struct DevRT {
    int kind;
    int n0, n1, n2;
    int min_samples;
    int bins;
    double s0, s1, s2;      // grid steps
    double rmin, period, step;
    int trapezium;          // 1: NumericalIntegrator sampling over the emitter's emission_function (cb2_rt_desc.integrator)
    double w2l[12];
    const int32_t* voxel_map;
};

struct cb2_scene {
    int device;
    DevScene host;           // host copy (pointers are device pointers)
    DevScene* dev;           // device copy
    void** allocs;           // device allocations to free
    int n_allocs, cap_allocs;
    // launch configuration
    int nw, bpl, smem_bytes;   // CTA-phased kernel (direct Bremsstrahlung)
    int bin_nw;                // bin_kernel warps per ray
    int fused;                 // table-driven scenes: fused_fast_kernel instead of state_fast_kernel + bin_kernel (CB2_FUSED=1)
    int warp_kernel;         // 1: warp-autonomous kernel (cb2_emission_warp.cu), 0: CTA-phased kernel with the direct Bremsstrahlung path
    int feat;                // the scene needs the general state kernel (beam, ThermalCXLine, TotalRadiatedPower)
    int ax_only;             // every scalar field of the scene is an AXISYM_BLEND: branch-free field evaluation
    int acc_f64;             // warp kernel: private accumulators in fp64 (else fp32)
    // staging buffers for the host-buffer entry point
    void* stage[8];
    size_t stage_bytes[8];
    unsigned long long* stats_dev;
    // Bremsstrahlung moment matrix [rays][k_pad] fp32 (grow-only)
    float* mom;
    size_t mom_bytes;
    // tensor-core contraction (cb2_contract_tc.cu): tf32 hi/lo parts of phi (transposed, K-major) and of the moments
    int contract_tc;
    int tc_pairs;              // CTA pairs of the persistent tcgen05 kernel (SMs / 2)
    float *phi_hi, *phi_lo, *mom_split, *tmp32;
    size_t mom_split_bytes, tmp32_bytes;
    // two-kernel line path (cb2_emission_warp.cu), grow-only: per-ray group offsets [batch+1], per-group live masks, and
    // the per-(group, component) line records [3][32] fp32 (centre, width, amplitude)
    int64_t* gbase;
    size_t gbase_bytes;
    unsigned* gmask;
    size_t gmask_bytes;
    unsigned* gblend;          // per-group masks of the samples the state tables do not cover (blend zone): generic kernel
    size_t gblend_bytes;
    DevMemo memo;              // state tables (memo.enabled: the table-driven state kernel runs)
    float memo_err;            // largest mid-interval error of the accepted core table
    int64_t memo_err_at;       // (entry << 24 | row) where it occurs
    float* rec;
    size_t rec_bytes;
    double* flat;              // per-ray wavelength-independent radiance (TotalRadiatedPower)
    size_t flat_bytes;
    int64_t* total_host;       // pinned, device-mapped word: the batch's group total reaches the host without a copy-engine transfer
    int64_t* total_dev;
    // host-buffer entry point: rows of finished ray batches are copied to the caller's buffer on a second stream while
    // the next batch computes (set for the duration of cb2_emission_render only)
    void* d2h_host;
    const int64_t* d2h_rows;   // optional destination row of every ray of the call (cb2_emission_render_rows)
    cudaStream_t copy_stream;
    cudaStream_t copy_stream2;  // strided row copies alternate between two streams (two copy engines)
    cudaEvent_t copy_ev;
    // optional per-kernel timing (cb2_scene_profile)
    int prof_on;
    double prof_ms[4];
    double prof_fixup_ms;      // part of prof_ms[0] spent in the generic kernel's fix-up pass behind the table-driven one
    int64_t prof_launches[4];
    cudaEvent_t prof_ev[10];
};

struct cb2_rt_scene {
    int device;
    DevRT rt;
    int32_t* voxel_map_dev;
    double* scratch;         // [n_warps][bins] per-warp dense accumulators
    int32_t* touched;        // [n_warps][touch_cap]
    int n_warps, touch_cap;
    int32_t* row_cols;       // single-traversal CSR: strided scratch rows [n_rays][touch_cap], grow-only
    double* row_len;
    size_t row_cap_rays;
    void* stage[8];
    size_t stage_bytes[8];
    unsigned long long* stats_dev;
};

// error channel (cb2_api.cu)
int cb2_fail(int code, const char* fmt, ...);
int cb2_cuda_check(cudaError_t e, const char* what);
#define CB2_CUDA(call)                                         \
    do {                                                       \
        int _rc = cb2_cuda_check((call), #call);               \
        if (_rc != CB2_OK) return _rc;                         \
    } while (0)

// kernels (cb2_emission.cu / cb2_raytransfer.cu)
int cb2_launch_emission(cb2_scene* sc, const DevRays& rays, void* out, int out_f64, double scale, int accumulate,
                        unsigned long long* stats_dev, cudaStream_t stream);
int cb2_emission_config(cb2_scene* sc);
size_t cb2_warp_smem_bytes(int nw, int acc_f64, int bins);
int cb2_launch_rt_compact(int64_t n_rays, int64_t row_stride, const int64_t* row_offset, const int32_t* scratch_cols, const double* scratch_len,
                          int32_t* columns, double* lengths, cudaStream_t st);
int cb2_d2h(void* dst, const void* src_dev, size_t bytes, cudaStream_t st);
int cb2_d2h_rows(void* dst_base, const int64_t* dest_row, int64_t n, const void* src_dev, size_t row_bytes, cudaStream_t st, cudaStream_t st2 = nullptr);
int cb2_launch_emission_warp(cb2_scene* sc, const DevRays& rays, void* out, int out_f64, double scale, int accumulate,
                             unsigned long long* stats, int count_samples, cudaStream_t stream);
int64_t cb2_warp_batch_rays(const cb2_scene* sc);
int cb2_memo_build(cb2_scene* sc);   // state tables of an eligible scene (after the device scene is complete); frees nothing on failure
int cb2_launch_sample_state(const cb2_scene* sc, const double* points_dev, int64_t n, double* out_dev, int points_in_plasma_space,
                            cudaStream_t stream);
int cb2_launch_beam_sample(const cb2_scene* sc, const double* beam_points_dev, int64_t n, double* out_dev, cudaStream_t stream);
int cb2_launch_rt(const cb2_rt_scene* sc, const DevRays& rays, int mode, double* dense_out, int accumulate,
                  int64_t* row_offset, int32_t* columns, double* lengths, unsigned long long* stats_dev, cudaStream_t stream);
int cb2_launch_scan(int64_t* counts_inout, int64_t n, cudaStream_t stream);
// rays per moment batch: bounds the moment matrix [batch][k_pad] fp32 to ~1.5 GB
static inline int64_t cb2_moment_batch(int k_pad) {
    int64_t batch = ((int64_t)3 << 29) / ((int64_t)k_pad * (int64_t)sizeof(float));
    batch = batch / 128 * 128;
    return batch < 128 ? 128 : batch;
}
int cb2_contract_tc_init(cb2_scene* sc, const float* phi, int k_pad, int n_pad);
void cb2_contract_tc_destroy(cb2_scene* sc);
int cb2_launch_contract_tc(cb2_scene* sc, const float* mom, int64_t n_rays, int k_pad, int n_pad, int bins, void* out, int out_f64,
                           double scale, cudaStream_t stream);
// out[ray][bin] += scale * sum_k mom[ray][k] phi[k][bin]   (cb2_contract.cu)
int cb2_launch_contract(const float* mom, const float* phi, int64_t n_rays, int k_pad, int n_pad, int bins, void* out, int out_f64,
                        double scale, cudaStream_t stream);
