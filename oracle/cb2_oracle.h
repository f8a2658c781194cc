/* cb2_oracle.h — entry points of libcb2_oracle.so, the fp64 scalar CPU restatement of the reference for the hot path.
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * It consumes the same descriptor structs as the product library (include/cherab_b200.h) and is not part of that ABI. */
#ifndef CB2_ORACLE_H
#define CB2_ORACLE_H
#include "../include/cherab_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

int         cb2o_abi_version(void);
const char* cb2o_last_error(void);
int cb2o_emission_render(const cb2_scene_desc* desc, const cb2_rays* rays, double* out,
                         double scale, int accumulate, int n_threads, cb2_stats* stats);
int cb2o_sample_state(const cb2_scene_desc* desc, const double* points, int64_t n, double* out);
int cb2o_state_width(const cb2_scene_desc* desc);
int cb2o_beam_sample(const cb2_scene_desc* desc, const double* beam_points, int64_t n, double* out);   /* beam/node.pyx:214-279 */
int cb2o_rt_render_dense(const cb2_rt_desc* desc, const cb2_rays* rays, double* out, int accumulate,
                         int n_threads, cb2_stats* stats);
/* building blocks exposed so the reference's own unit tests can be replayed against the oracle */
int cb2o_add_gaussian_line(double radiance, double wavelength, double sigma,
                           const cb2_spectral_grid* grid, double* samples);              /* gaussian.pyx:40-90 */
int cb2o_add_lorentzian_line(double radiance, double wavelength, double lambda_1_2,
                             const cb2_spectral_grid* grid, double* samples,
                             double rtol, int min_order, int max_order);                  /* stark.pyx:88-147 */
double cb2o_interp1d_cubic(const double* x, const double* f, int n, double px, int extrapolate);       /* raysect Interpolator1DArray 'cubic' */
double cb2o_interp2d_cubic(const double* x, const double* y, const double* f, int nx, int ny,
                           double px, double py, int extrapolate);                                       /* raysect Interpolator2DArray 'cubic' */
double cb2o_gauss_legendre(double (*fn)(double, void*), void* ctx, double a, double b,
                           double rtol, int min_order, int max_order);                                   /* integrators1d.pyx:189-224 */
double cb2o_gaunt_factor(const cb2_gaunt* g, double z, double te, double wavelength);                   /* gaunt.pyx:109-140 */
double cb2o_pec_evaluate(const cb2_rate2d* pec, double wavelength, double ne, double te);               /* pec.pyx:70-77 */
double cb2o_interp3d_cubic(const double* x, const double* y, const double* z, const double* f, int nx, int ny, int nz,
                           double px, double py, double pz);                                             /* raysect Interpolator3DArray 'cubic' */
double cb2o_thermal_cx_pec_evaluate(const cb2_rate3d* pec, double wavelength, double ne, double te, double td); /* pec.pyx:186-194 */

/* first wall hit of n rays against n_triangles world-space triangles [n_triangles][3][3], brute force (first_wall.py:120-184) */
int cb2o_wall_hit(const double* vertices, int64_t n_triangles, const double* origin, const double* direction, int64_t n, double* t_hit);
#ifdef __cplusplus
}
#endif
#endif /* CB2_ORACLE_H */
