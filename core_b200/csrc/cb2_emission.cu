// cb2_emission.cu — the emission hot path on sm_100a.
//
// One CTA integrates one ray (all of its segments).  Work is organised in chunks of NT = blockDim.x samples:
//
//   STATE phase   (one thread per sample)  position -> (R, Z) -> psi_n / LCFS mask / blend weight / mesh triangle,
//                 each evaluated ONCE per sample (the reference re-walks the function tree for every quantity,
//                 SURVEY 0.5) -> species profiles -> PEC bicubics -> per line component a compact record
//                 (centre, width, amplitude, series coefficients, bin range) in shared memory; for Bremsstrahlung
//                 the per-sample cubic of sum_i n_i Z_i^2 g_ff(Z_i, Te, lambda) in log10(lambda).
//   BIN phase     (one warp per spectral tile)  warp w owns bins [w*32*BPL, (w+1)*32*BPL), lane l owns bins
//                 tile + 32 j + l (interleaved so narrow lines still fill the warp); accumulators live in
//                 registers (fp32 per chunk, flushed into fp64 per-ray accumulators after every chunk: "fp32 math,
//                 fp64 per-ray accumulation").  Every warp sweeps the chunk's records and adds the bin integrals
//                 of the components that touch its tile.  No atomics, no cross-warp reduction.
//
// Gaussian bin integrals I = 1/2 [erf(x_hi) - erf(x_lo)] are evaluated in fp32 with RELATIVE accuracy
// (tools/proto_gauss_fp32.py: <= 1e-5 everywhere the acceptance floor allows):
//   h = half bin width in units of sqrt(2) sigma
//   h <  1/16 : I = kb/sqrt(pi) exp(-m^2) [1 + h^2 H2(m)/6 + h^4 H4(m)/120], one MUFU.EX2 per bin, no cancellation;
//   h >= 1/16 : differences of 1/2 erfc(|x|) = s Q(s) exp(-x^2), s = 1/(1 + |x|/2) (degree-9 minimax Q), the lower
//               edge value taken from the neighbouring lane by warp shuffle (the reference's lower=upper recurrence).
// Follows: cherab/core/plasma/material.pyx:48-63, model/plasma/impact_excitation.pyx:78-100, recombination.pyx:78-100,
// bremsstrahlung.pyx:70-90,169-208, model/lineshape/gaussian.pyx:40-139, doppler.pyx:29-59, multiplet.pyx:93-117,
// zeeman.pyx:113-365, atomic/gaunt.pyx:109-140, tools/equilibrium/efit.pyx:219-546, generomak/plasma/plasma.py:580-638,
// and Raysect's NumericalIntegrator.
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "cb2_device.cuh"

// ------------------------------------------------------------------------------------------------------------------
// STATE phase helpers
// ------------------------------------------------------------------------------------------------------------------
struct RecWriter {
    float4* rec;      // [ncomp][2][NT]
    int* rng;         // [ncomp][NG][2] (lo, hi) bounding bin ranges per group of 32 samples, relative bins
    int nt, tid, ng, group;
    int bins;
    unsigned long long gauss_evals, lorentz_evals;
};


__device__ __forceinline__ float pack_range(int lo, int hi) { return __int_as_float((lo & 0xffff) | (hi << 16)); }
__device__ __forceinline__ void unpack_range(float p, int& lo, int& hi) {
    const int v = __float_as_int(p);
    lo = (int)(short)(v & 0xffff);
    hi = v >> 16;
}

// Write one Gaussian component: centre cf (bins, relative to the slot's integer origin c0_int), width sigma_b (bins),
// amplitude amp = weight * radiance / delta_wavelength.   (add_gaussian_line, gaussian.pyx:40-90)
// record a = (kx, xoff, s0, s1), b = (s2, s3, s4, packed [lo, hi)):
//   series path (s1 >= 0): x' = rel*kx + xoff = sqrt(log2 e) * (bin centre - line centre)/(sqrt2 sigma), m = x'^2,
//                          value = exp2(-m) * (s0 + m (s1 + m (s2 + m (s3 + m s4))))   [amplitude and 1/log2 e folded in]
//                          = amp * kb/sqrt(pi) exp(-x^2) sum_{n<=4} H_2n(x) h^2n/(2n+1)!,  h = kb/2 <= 0.36
//   erfc path   (s1 <  0): x = rel*kx + xoff at the bin's UPPER edge, s0 = amplitude  (sub-bin lines, sigma < 1 bin)
__device__ __forceinline__ void put_gaussian(RecWriter& W, const DevComp& cs, int slot, float cf, float sigma_b, float amp) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (amp > 0.f && sigma_b > 0.f) {
        const float cut = 10.0f * sigma_b;                       // GAUSSIAN_CUTOFF_SIGMA
        const float win_lo = (float)(-cs.c0_int), win_hi = (float)(W.bins - cs.c0_int);
        const float flo = floorf(cf - cut), fhi = ceilf(cf + cut);
        if (fhi > win_lo && flo < win_hi) {
            const int lo = (int)fmaxf(flo, win_lo), hi = (int)fminf(fhi, win_hi);
            if (hi > lo) {
                const float kb = 0.70710678f / sigma_b;          // delta / (sqrt(2) sigma), per bin
                const float h = 0.5f * kb;
                if (h <= H_SERIES_MAX) {
                    const float h2 = h * h;
                    const float t1 = h2 * (1.0f / 6.0f), t2 = h2 * h2 * (1.0f / 120.0f), t3 = h2 * h2 * h2 * (1.0f / 5040.0f),
                                t4 = h2 * h2 * h2 * h2 * (1.0f / 362880.0f);
                    const float A = amp * kb * INV_SQRT_PI;
                    a.x = kb * SQRT_L2E;
                    a.y = (0.5f - cf) * a.x;
                    a.z = A * (1.0f - 2.0f * t1 + 12.0f * t2 - 120.0f * t3 + 1680.0f * t4);
                    a.w = A * (4.0f * t1 - 48.0f * t2 + 720.0f * t3 - 13440.0f * t4) * INV_L2E;
                    b.x = A * (16.0f * t2 - 480.0f * t3 + 13440.0f * t4) * (INV_L2E * INV_L2E);
                    b.y = A * (64.0f * t3 - 3584.0f * t4) * (INV_L2E * INV_L2E * INV_L2E);
                    b.z = A * (256.0f * t4) * (INV_L2E * INV_L2E * INV_L2E * INV_L2E);
                } else {
                    a.x = kb;
                    a.y = (1.0f - cf) * kb;                      // x at the bin's upper edge
                    a.z = amp;
                    a.w = -1.0f;                                 // erfc-difference path
                }
                if (a.z > 0.f) {
                    b.w = pack_range(lo, hi);
                    atomicMin(&W.rng[2 * (slot * W.ng + W.group)], lo);
                    atomicMax(&W.rng[2 * (slot * W.ng + W.group) + 1], hi);
                    W.gauss_evals += (unsigned long long)(hi - lo) + 1ull;
                } else {
                    a = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
    W.rec[(2 * slot) * W.nt + W.tid] = a;
    W.rec[(2 * slot + 1) * W.nt + W.tid] = b;
}

// Write one modified-Lorentzian component (add_lorentzian_line, stark.pyx:88-147): centre cf (relative bins), FWHM
// lam_b (bins), amplitude amp.  record a = (1/lam_b, -cf/lam_b, amp, 0), b.w = packed [lo, hi) of the +-50 FWHM cut-off.
__device__ __forceinline__ void put_lorentzian(RecWriter& W, const DevComp& cs, int slot, float cf, float lam_b, float amp) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (amp > 0.f && lam_b > 0.f) {
        const float cut = 50.0f * lam_b;                         // LORENTZIAN_CUTOFF_GAMMA
        const float win_lo = (float)(-cs.c0_int), win_hi = (float)(W.bins - cs.c0_int);
        const float flo = floorf(cf - cut), fhi = ceilf(cf + cut);
        if (fhi > win_lo && flo < win_hi) {
            const int lo = (int)fmaxf(flo, win_lo), hi = (int)fminf(fhi, win_hi);
            if (hi > lo) {
                a.x = 1.0f / lam_b;
                a.y = -cf * a.x;
                a.z = amp;
                b.w = pack_range(lo, hi);
                atomicMin(&W.rng[2 * (slot * W.ng + W.group)], lo);
                atomicMax(&W.rng[2 * (slot * W.ng + W.group) + 1], hi);
                W.lorentz_evals += (unsigned long long)(hi - lo) + 1ull;
            }
        }
    }
    W.rec[(2 * slot) * W.nt + W.tid] = a;
    W.rec[(2 * slot + 1) * W.nt + W.tid] = b;
}

__device__ __forceinline__ void put_empty(RecWriter& W, int slot) {
    W.rec[(2 * slot) * W.nt + W.tid] = make_float4(0.f, 0.f, 0.f, 0.f);
}


// All line models at one sample (PlasmaMaterial.emission_function loop, material.pyx:59-61).
static __device__ void sample_lines(const DevScene& S, const SampleIn& in, const AxCtx& ctx, float ne, float te, RecWriter& W, unsigned& ood) {
    const bool live = ne > 0.f && te > 0.f && in.weight > 0.f;
    float lne = 0.f, lte = 0.f;
    if (live) {
        lne = log10f(ne) + 19.0f;     // densities are stored in units of 1e19 m^-3
        lte = log10f(te);
    }
    int cur = -1;
    float ni = 0.f, ts = 0.f, vd = 0.f;
    float3 bf = make_float3(0.f, 0.f, 0.f);
    float bm = 0.f, cos_sqr = 0.f;
    bool have_b = false;
    for (int m = 0; m < S.n_models; m++) {
        const DevModel& M = S.models[m];
        if (M.kind == CB2_MODEL_BREMSSTRAHLUNG) continue;
        bool on = live;
        if (on && M.species != cur) {
            cur = M.species;
            const DevSpecies& sp = S.species[cur];
            ni = eval_scalar(sp.density, ctx, in.x, in.y, in.z);
            ts = eval_scalar(sp.temperature, ctx, in.x, in.y, in.z);
            const float3 v = eval_vector(sp.velocity, ctx);
            vd = v.x * in.dx + v.y * in.dy + v.z * in.dz;   // velocity projected on the (unit) ray direction
        }
        on = on && ni > 0.f;
        float radiance = 0.f;
        if (on) {
            // radiance = 1/(4 pi) PEC ne ni  (impact_excitation.pyx:99); exp10 of (log PEC + 38) * (ne ni * 1e-38)
            radiance = RECIP_4_PI * exp10f(eval_pec_log(M, lne, lte, ood)) * ne * ni;
        }
        const float amp = radiance * in.weight * M.inv_delta;
        if (M.shape == CB2_SHAPE_STARK) {
            // StarkBroadenedLine.add_line (stark.pyx:251-348): pseudo-Voigt = (1-eta) Gaussian + eta modified Lorentzian per
            // Zeeman component; does NOT return early on ts <= 0
            bool done = !(on && amp > 0.f);
            float sigma_b = 0.f, lam_b = 0.f, eta = 0.f;
            if (!done) {
                const float SIGMA2FWHM = 2.3548200450309493f;
                const float fl = M.param[0] * powf(ne, M.param[1]) / powf(te, M.param[2]);      // nm (ne in 1e19 m^-3 folded in)
                const float fg = ts > 0.f ? SIGMA2FWHM * M.sigma_coef * sqrtf(ts) / M.inv_delta : 0.f;   // nm
                if (fl == 0.f && fg == 0.f) done = true;
                else {
                    float full;
                    if (fg <= fl) {
                        const float r = fg / fl;
                        full = fl * (1.f + r * r * (0.57575f + r * (0.37902f + r * (-0.42519f + r * (-0.31525f + r * 0.31718f)))));
                    } else {
                        const float r = fl / fg;
                        full = fg * (1.f + r * (0.15882f + r * (1.04388f + r * (-1.38281f + r * (0.46251f + r * (0.82325f + r * -0.58026f))))));
                    }
                    float sigma = full / SIGMA2FWHM;
                    const float l2t = fl / full;
                    if (l2t < 0.01f) { eta = 0.f; full = 0.f; }
                    else if (l2t > 0.999f) { eta = 1.f; sigma = 0.f; }
                    else {
                        const float lg = logf(l2t);
                        eta = expf(5.14820e-04f + lg * (1.38821e+00f + lg * (-9.60424e-02f + lg * (-3.83995e-02f + lg * (-7.40042e-03f + lg * -5.47626e-04f)))));
                    }
                    sigma_b = sigma * M.inv_delta;
                    lam_b = full * M.inv_delta;
                }
            }
            if (done) {
                for (int k = 0; k < M.ncomp; k++) put_empty(W, M.comp0 + k);
                continue;
            }
            if (!have_b) {
                bf = eval_b_field(S, ctx);
                bm = sqrtf(bf.x * bf.x + bf.y * bf.y + bf.z * bf.z);
                const float c = bm > 0.f ? (bf.x * in.dx + bf.y * in.dy + bf.z * in.dz) / bm : 0.f;
                cos_sqr = c * c;
                have_b = true;
            }
            const DevComp& c0 = S.comps[M.comp0];
            const float dop = vd * M.inv_c;
            const float shift0 = M.wavelength * dop * M.inv_delta;
            const float wg = 1.f - eta;
            if (bm == 0.f) {
                const float r = M.polarisation == CB2_POL_NO ? amp : 0.5f * amp;
                put_gaussian(W, c0, M.comp0, c0.c0_frac + shift0, sigma_b, wg * r);
                put_lorentzian(W, c0, M.comp0 + 3, c0.c0_frac + shift0, lam_b, eta * r);
                put_empty(W, M.comp0 + 1); put_empty(W, M.comp0 + 2); put_empty(W, M.comp0 + 4); put_empty(W, M.comp0 + 5);
                continue;
            }
            const float sin_sqr = 1.0f - cos_sqr;
            if (M.polarisation != CB2_POL_SIGMA) {
                const float r = 0.5f * sin_sqr * amp;
                put_gaussian(W, c0, M.comp0, c0.c0_frac + shift0, sigma_b, wg * r);
                put_lorentzian(W, c0, M.comp0 + 3, c0.c0_frac + shift0, lam_b, eta * r);
            } else { put_empty(W, M.comp0); put_empty(W, M.comp0 + 3); }
            if (M.polarisation != CB2_POL_PI) {
                const float r = (0.25f * sin_sqr + 0.5f * cos_sqr) * amp;
                const float e = BOHR_MAGNETON * bm * M.wavelength * (1.0f / HC_EV_NM_F);
                const float dlp = M.wavelength * e / (1.0f - e), dlm = -M.wavelength * e / (1.0f + e);
                const float cp = c0.c0_frac + (dlp + (M.wavelength + dlp) * dop) * M.inv_delta;
                const float cm = c0.c0_frac + (dlm + (M.wavelength + dlm) * dop) * M.inv_delta;
                put_gaussian(W, c0, M.comp0 + 1, cp, sigma_b, wg * r);
                put_lorentzian(W, c0, M.comp0 + 4, cp, lam_b, eta * r);
                put_gaussian(W, c0, M.comp0 + 2, cm, sigma_b, wg * r);
                put_lorentzian(W, c0, M.comp0 + 5, cm, lam_b, eta * r);
            } else { put_empty(W, M.comp0 + 1); put_empty(W, M.comp0 + 2); put_empty(W, M.comp0 + 4); put_empty(W, M.comp0 + 5); }
            continue;
        }
        // all Gaussian-family shapes return before touching the spectrum if ts <= 0 (gaussian.pyx:127-129)
        const bool shape_on = on && ts > 0.f && amp > 0.f;
        if (!shape_on) {
            for (int k = 0; k < M.ncomp; k++) put_empty(W, M.comp0 + k);
            continue;
        }
        float sigma_b = M.sigma_coef * sqrtf(ts);                      // thermal_broadening, doppler.pyx:48-59, in bins
        const float dop = vd * M.inv_c;                                 // doppler_shift: lambda (1 + v.d/c)
        const float shift0 = M.wavelength * dop * M.inv_delta;          // Doppler shift of the rest wavelength, in bins
        if (M.shape == CB2_SHAPE_GAUSSIAN) {
            const DevComp& cs = S.comps[M.comp0];
            put_gaussian(W, cs, M.comp0, cs.c0_frac + shift0, sigma_b, amp);
            continue;
        }
        if (M.shape == CB2_SHAPE_MULTIPLET) {                           // multiplet.pyx:108-115
            for (int k = 0; k < M.ncomp; k++) {
                const DevComp& cs = S.comps[M.comp0 + k];
                const float sh = __ldg(M.mult_lambda + k) * dop * M.inv_delta;
                put_gaussian(W, cs, M.comp0 + k, cs.c0_frac + sh, sigma_b, amp * __ldg(M.mult_ratio + k));
            }
            continue;
        }
        // Zeeman family (zeeman.pyx)
        if (!have_b) {
            bf = eval_b_field(S, ctx);
            bm = sqrtf(bf.x * bf.x + bf.y * bf.y + bf.z * bf.z);
            const float c = bm > 0.f ? (bf.x * in.dx + bf.y * in.dy + bf.z * in.dz) / bm : 0.f;
            cos_sqr = c * c;
            have_b = true;
        }
        if (M.shape == CB2_SHAPE_PARAM_ZEEMAN) sigma_b *= sqrtf(1.0f + M.param[1] * M.param[1] * powf(ts, 2.0f * M.param[2]));
        const float sin_sqr = 1.0f - cos_sqr;
        const bool pol_pi = M.polarisation != CB2_POL_SIGMA, pol_sigma = M.polarisation != CB2_POL_PI;
        if (bm == 0.f) {
            // no splitting: single Gaussian, halved if a polarisation filter is set (zeeman.pyx:132-136)
            const DevComp& cs = S.comps[M.comp0];
            put_gaussian(W, cs, M.comp0, cs.c0_frac + shift0, sigma_b, M.polarisation == CB2_POL_NO ? amp : 0.5f * amp);
            for (int k = 1; k < M.ncomp; k++) put_empty(W, M.comp0 + k);
            continue;
        }
        const float a_pi = 0.5f * sin_sqr * amp, a_sigma = (0.25f * sin_sqr + 0.5f * cos_sqr) * amp;
        if (M.shape == CB2_SHAPE_ZEEMAN_TRIPLET || M.shape == CB2_SHAPE_PARAM_ZEEMAN) {
            float dl_plus, dl_minus;  // wavelength offsets of the sigma components from the rest wavelength
            if (M.shape == CB2_SHAPE_ZEEMAN_TRIPLET) {
                // hc/(hc/l0 -+ muB B) - l0 = +- l0 e/(1 -+ e), e = muB B l0 / hc   (zeeman.pyx:152-158)
                const float e = BOHR_MAGNETON * bm * M.wavelength * (1.0f / HC_EV_NM_F);
                dl_plus = M.wavelength * e / (1.0f - e);
                dl_minus = -M.wavelength * e / (1.0f + e);
            } else {
                dl_plus = 0.5f * M.param[0] * bm;                      // zeeman.pyx:260-264
                dl_minus = -dl_plus;
            }
            const DevComp& c0 = S.comps[M.comp0];
            if (pol_pi) put_gaussian(W, c0, M.comp0, c0.c0_frac + shift0, sigma_b, a_pi); else put_empty(W, M.comp0);
            if (pol_sigma) {
                put_gaussian(W, c0, M.comp0 + 1, c0.c0_frac + (dl_plus + (M.wavelength + dl_plus) * dop) * M.inv_delta, sigma_b, a_sigma);
                put_gaussian(W, c0, M.comp0 + 2, c0.c0_frac + (dl_minus + (M.wavelength + dl_minus) * dop) * M.inv_delta, sigma_b, a_sigma);
            } else {
                put_empty(W, M.comp0 + 1);
                put_empty(W, M.comp0 + 2);
            }
            continue;
        }
        if (M.shape == CB2_SHAPE_ZEEMAN_MULTIPLET) {                    // zeeman.pyx:340-363, atomic/zeeman.pyx:87-129
            float fb = (bm - M.b0) * M.inv_db;
            fb = fminf(fmaxf(fb, 0.f), (float)(M.n_b - 1));
            const int ib = min((int)fb, M.n_b - 2);
            const float tb = fb - (float)ib;
            const int offs[4] = {0, M.n_pi, M.n_pi + M.n_sp, M.n_pi + M.n_sp + M.n_sm};
            for (int g = 0; g < 3; g++) {
                const bool gon = g == 0 ? pol_pi : pol_sigma;
                float rsum = 0.f;
                for (int k = offs[g]; k < offs[g + 1]; k++) {
                    const float* r = M.zee_ratio + (size_t)k * M.n_b + ib;
                    rsum += fmaf(tb, __ldg(r + 1) - __ldg(r), __ldg(r));
                }
                const float rnorm = rsum > 0.f ? 1.0f / rsum : 1.0f;
                for (int k = offs[g]; k < offs[g + 1]; k++) {
                    const DevComp& cs = S.comps[M.comp0 + k];
                    if (!gon) { put_empty(W, M.comp0 + k); continue; }
                    const float* r = M.zee_ratio + (size_t)k * M.n_b + ib;
                    const float* l = M.zee_dlambda + (size_t)k * M.n_b + ib;
                    const float ratio = fmaf(tb, __ldg(r + 1) - __ldg(r), __ldg(r)) * rnorm;
                    const float dl = fmaf(tb, __ldg(l + 1) - __ldg(l), __ldg(l));
                    put_gaussian(W, cs, M.comp0 + k, cs.c0_frac + (dl + (M.wavelength + dl) * dop) * M.inv_delta, sigma_b,
                                 (g == 0 ? a_pi : a_sigma) * ratio);
                }
            }
            continue;
        }
        for (int k = 0; k < M.ncomp; k++) put_empty(W, M.comp0 + k);
    }
}

// Bremsstrahlung at one sample: cubic(s) in l' = log10(lambda) - lref of  A * sum_i n_i Z_i^2 g_ff(Z_i, Te, lambda),
// A = weight * BREMS_CONST * ne / sqrt(Te); up to three pieces when the window crosses knots of the Gaunt table's u grid.
// record: r0 = (a2, n_pieces (+8 if the one-point rule is not accurate enough for this sample), bc1, bc2) where piece 0
//         (lowest u, longest wavelengths) covers bins [bc1, bins), piece 1 [bc2, bc1), piece 2 [0, bc2),
//         r1..r3 = piece coefficients (c0..c3).
static __device__ void sample_brems(const DevScene& S, const SampleIn& in, const AxCtx& ctx, float ne, float te, float4* brec, int nt, int tid,
                             unsigned long long& brems_evals, unsigned& ood) {
    const DevBrems& B = S.brems;
    float4 r0 = make_float4(0.f, 0.f, __int_as_float(-1), __int_as_float(-1));
    float4 pc0 = make_float4(0, 0, 0, 0), pc1 = pc0, pc2 = pc0;
    if (ne > 0.f && te > 0.f && in.weight > 0.f) {
        const float lte = log10f(te);
        const float L0 = B.log_hc - lte - B.lref;            // log10 u = L0 - l'
        const float lu_lo = L0 - B.lp_max, lu_hi = L0 - B.lp_min;
        const DevTable2D& G = B.gaunt;
        // interval index: -1 Born (u < u_min), nx-1 classical (u >= u_max), else table cell
        auto interval = [&](float lu) -> int {
            if (lu < B.lu_min) return -1;
            if (lu >= B.lu_max) return G.nx - 1;
            return search_knots(G.x, G.nx, lu);
        };
        const int i_lo = interval(lu_lo);
        int i_hi = interval(lu_hi);
        if (i_hi > i_lo + 2) { i_hi = i_lo + 2; ood++; }
        const int np = i_hi - i_lo + 1;
        // wavelength (1/lambda) thresholds where the piece changes: u = knot  <=>  rho = knot * Te / hc
        // bins whose centre wavelength is <= lambda_knot (1/lambda >= knot Te/hc) belong to the next piece:
        // store the first bin of the lower piece, bc = floor(xc - 1/2) + 1 with xc = (lambda_knot - lambda_min)/delta
        // (integers are stored bit-cast: float<->int conversions run on the same XU pipe as MUFU.EX2, the binding pipe of the BIN phase)
        if (np > 1) r0.z = __int_as_float((int)fminf(fmaxf(floorf((exp10f(B.log_hc - lte - __ldg(G.x + i_lo + 1)) - S.min_wavelength) / S.delta - 0.5f) + 1.0f, -1.0f), 1.0e6f));
        if (np > 2) r0.w = __int_as_float((int)fminf(fmaxf(floorf((exp10f(B.log_hc - lte - __ldg(G.x + i_lo + 2)) - S.min_wavelength) / S.delta - 0.5f) + 1.0f, -1.0f), 1.0e6f));
        const float k1 = 0.5513288954217921f * 2.302585092994046f;            // sqrt(3)/pi * ln 10
        const float k0 = 0.5513288954217921f * (1.3862943611198906f - 0.5772156649015329f);  // sqrt(3)/pi (ln 4 - gamma_E)
        for (int s = 0; s < B.n_charged; s++) {
            const DevSpecies& sp = S.species[B.charged[s]];
            const float ni = eval_scalar(sp.density, ctx, in.x, in.y, in.z);
            if (!(ni > 0.f)) continue;
            const float w = ni * sp.z2;
            const float lg = log10f(sp.z2 * RYDBERG_EV / te);                  // log10 gamma^2
            int jg = 0; float ug = 0.f;
            const bool g_hi = lg >= B.lg_max, g_lo = lg < B.lg_min;
            if (!g_hi && !g_lo) {
                jg = search_knots(G.y, G.ny, lg);
                ug = (lg - __ldg(G.y + jg)) * __ldg(G.inv_wy + jg);
            }
#pragma unroll
            for (int p = 0; p < 3; p++) {
                if (p < np) {
                    const int iu = i_lo + p;
                    float c0, c1 = 0.f, c2 = 0.f, c3 = 0.f;
                    if (g_hi || iu >= G.nx - 1) c0 = 1.0f;                          // classical limit (gaunt.pyx:129-130)
                    else if (g_lo || iu < 0) { c0 = k0 - k1 * L0; c1 = k1; }        // Born approximation (gaunt.pyx:133-134)
                    else {
                        const float4 e = eval2d_rows(G, iu, jg, ug);                // g = sum_p e_p t^p, t = alpha - beta l'
                        const float beta = __ldg(G.inv_wx + iu), alpha = (L0 - __ldg(G.x + iu)) * beta;
                        c0 = fmaf(fmaf(fmaf(e.w, alpha, e.z), alpha, e.y), alpha, e.x);
                        c1 = -beta * fmaf(fmaf(3.0f * e.w, alpha, 2.0f * e.z), alpha, e.y);
                        c2 = beta * beta * fmaf(3.0f * e.w, alpha, e.z);
                        c3 = -beta * beta * beta * e.w;
                    }
                    float4& pc = p == 0 ? pc0 : (p == 1 ? pc1 : pc2);
                    pc.x = fmaf(w, c0, pc.x); pc.y = fmaf(w, c1, pc.y);
                    pc.z = fmaf(w, c2, pc.z); pc.w = fmaf(w, c3, pc.w);
                }
            }
        }
        const float A = in.weight * B.pref * ne * rsqrtf(te);
        pc0.x *= A; pc0.y *= A; pc0.z *= A; pc0.w *= A;
        pc1.x *= A; pc1.y *= A; pc1.z *= A; pc1.w *= A;
        pc2.x *= A; pc2.y *= A; pc2.z *= A; pc2.w *= A;
        // one-point (bin centre) rule: relative error ~ h^2/6 with h = delta/2 * |d ln eps / d lambda| at lambda_min
        const float hh = fabsf(B.mid_c1 / te - B.mid_c0) + 0.1f * B.mid_c0;
        const bool multi = (B.nq > 1) && (hh * hh * (1.0f / 6.0f) > 2e-6f);
        r0.x = B.exp_coef / te;
        r0.y = __int_as_float(np + (multi ? 8 : 0));
        brems_evals += (unsigned long long)S.bins;
    }
    brec[tid] = r0;
    brec[nt + tid] = pc0;
    brec[2 * nt + tid] = pc1;
    brec[3 * nt + tid] = pc2;
}



// ------------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------------
// WINDOW PASS — bin integrals of one line component over one 32-bin window, lanes = samples.  Every lane evaluates its
// own sample's Gaussian at the window's bins (series form: no cross-bin dependence; erfc form: the reference's
// lower = upper recurrence along the window, gaussian.pyx:78-88) and adds into its private part[w].  After all groups
// of the chunk have been accumulated the 32 partial vectors are summed across lanes with a transpose-reduce
// (31 shuffles) so that lane w holds window bin w, which is added to the fp64 per-ray accumulator.  Lane utilisation is
// 100% however narrow the line is, and the per-sample bookkeeping is paid once per 32 bins.
template <bool FULL_WINDOW>
__device__ __forceinline__ void window_accumulate(float (&part)[32], const float4* __restrict__ ra, const float4* __restrict__ rb, int n,
                                                  int wbase, int wcount, int lane) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bq = make_float4(0.f, 0.f, 0.f, 0.f);
    int lo = 0, hi = 0;
    if (lane < n) {
        a = ra[lane];
        if (a.z != 0.f) { bq = rb[lane]; unpack_range(bq.w, lo, hi); lo -= wbase; hi -= wbase; }
    }
    const bool live = hi > 0 && lo < wcount && a.z != 0.f;           // this sample touches the window
    const bool series = live && a.w >= 0.f, erfc = live && a.w < 0.f;
    if (__any_sync(FULL, series)) {
        // exp2(-m) (s0 + m (s1 + m (s2 + m (s3 + m s4)))), m = x'^2, x' = (wbase + w) kx + xoff  (bin centres)
        const float x0 = fmaf((float)wbase, a.x, a.y);
        const float s0 = series ? a.z : 0.f, s1 = series ? a.w : 0.f, s2 = series ? bq.x : 0.f, s3 = series ? bq.y : 0.f,
                    s4 = series ? bq.z : 0.f;
#pragma unroll
        for (int w = 0; w < 32; w++) {
            if (FULL_WINDOW || w < wcount) {                           // warp-uniform
                const float x = fmaf((float)w, a.x, x0);
                const float m2 = x * x;
                part[w] = fmaf(ex2_approx(-m2), fmaf(fmaf(fmaf(fmaf(s4, m2, s3), m2, s2), m2, s1), m2, s0), part[w]);
            }
        }
    }
    if (__any_sync(FULL, erfc)) {
        // sub-bin lines: 1/2 erfc differences; the record holds x at a bin's UPPER edge: x(rel) = rel * kx + xoff
        const float xbase = fmaf((float)(wbase - 1), a.x, a.y);       // x at the lower edge of window bin 0
        const float amp = erfc ? a.z : 0.f;
        float xl = xbase;
        float tl = half_erfc(fabsf(xl));
#pragma unroll
        for (int w = 0; w < 32; w++) {
            if (FULL_WINDOW || w < wcount) {
                const float xu = fmaf((float)(w + 1), a.x, xbase);
                const float tu = half_erfc(fabsf(xu));
                const float dd = (xl >= 0.f) ? (tl - tu) : ((xu <= 0.f) ? (tu - tl) : (1.0f - tl - tu));
                if (w >= lo && w < hi) part[w] = fmaf(amp, dd, part[w]);
                xl = xu; tl = tu;
            }
        }
    }
}

// window pass for a modified-Lorentzian component (lanes = samples): float64 CDF differences along the window's edges
template <bool FULL_WINDOW>
__device__ __forceinline__ void window_accumulate_lorentz(float (&part)[32], const float4* __restrict__ ra, const float4* __restrict__ rb,
                                                          int n, int wbase, int wcount, int lane, const double2* __restrict__ tab,
                                                          double phi_inf) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    int lo = 0, hi = 0;
    if (lane < n) {
        a = ra[lane];
        if (a.z != 0.f) { const float4 bq = rb[lane]; unpack_range(bq.w, lo, hi); lo -= wbase; hi -= wbase; }
    }
    const bool live = hi > 0 && lo < wcount && a.z != 0.f;
    if (!__any_sync(FULL, live)) return;
    // u at bin edge e (relative bins): u = e / lam_b - cf / lam_b
    const double inv_l = (double)a.x, off = (double)a.y;
    double pl = live ? lorentz_cdf(tab, phi_inf, (double)wbase * inv_l + off) : 0.0;
#pragma unroll
    for (int w = 0; w < 32; w++) {
        if (FULL_WINDOW || w < wcount) {
            if (live && w >= lo - 0 && w < hi) {
                if (w == lo && lo > 0) pl = lorentz_cdf(tab, phi_inf, (double)(wbase + w) * inv_l + off);
                const double pu = lorentz_cdf(tab, phi_inf, (double)(wbase + w + 1) * inv_l + off);
                part[w] = fmaf(a.z, (float)(pu - pl), part[w]);
                pl = pu;
            }
        }
    }
}

// all windows of component c that belong to this warp, for one chunk
template <int NW, int NG, int NT>
__device__ __forceinline__ void line_windows(const float4* __restrict__ ra, const float4* __restrict__ rb, const int* __restrict__ rng,
                                             int nact, int c, int c0_int, int bins, double* __restrict__ racc, int warp, int lane,
                                             int type, const double2* __restrict__ ltab, double phi_inf) {
    int Rlo = INT_MAX, Rhi = INT_MIN;                                  // union over the chunk
#pragma unroll
    for (int g = 0; g < NG; g++) { Rlo = min(Rlo, rng[2 * g]); Rhi = max(Rhi, rng[2 * g + 1]); }
    if (Rhi <= Rlo) return;
    int k = (warp - c) % NW;
    if (k < 0) k += NW;
    for (int wbase = Rlo + 32 * k; wbase < Rhi; wbase += 32 * NW) {
        const int wcount = min(32, Rhi - wbase);
        float part[32];
#pragma unroll
        for (int w = 0; w < 32; w++) part[w] = 0.f;
        for (int g = 0; g * 32 < nact; g++) {
            if (rng[2 * g + 1] <= wbase || rng[2 * g] >= wbase + wcount) continue;   // group does not touch the window
            const int n = min(32, nact - g * 32);
            if (type == 1) window_accumulate_lorentz<false>(part, ra + g * 32, rb + g * 32, n, wbase, wcount, lane, ltab, phi_inf);
            else if (wcount == 32) window_accumulate<true>(part, ra + g * 32, rb + g * 32, n, wbase, 32, lane);
            else window_accumulate<false>(part, ra + g * 32, rb + g * 32, n, wbase, wcount, lane);
        }
        // transpose-reduce: after the step with offset o a lane keeps the half of its vector selected by (lane & o)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < o; i++) {
                const float send = upper ? part[i] : part[i + o];
                const float keep = upper ? part[i + o] : part[i];
                part[i] = keep + __shfl_xor_sync(FULL, send, o);
            }
        }
        const int bin = c0_int + wbase + lane;
        if (lane < wcount && part[0] != 0.f && bin >= 0 && bin < bins) atomicAdd(&racc[bin], (double)part[0]);
    }
}


#ifndef CB2_TARGET_THREADS
#define CB2_TARGET_THREADS 512
#endif
#ifndef CB2_MIN_BLOCKS_NW4
#define CB2_MIN_BLOCKS_NW4 4
#endif
#ifndef CB2_MIN_BLOCKS_NW8
#define CB2_MIN_BLOCKS_NW8 2
#endif

template <int NW, int BPL, int BREMS>
__global__ void __launch_bounds__(NW * 32, (NW <= 4 ? 768 : 512) / (NW * 32))
emission_kernel(const DevScene* __restrict__ Sp, DevRays rays, void* __restrict__ out, int out_f64, double scale, int accumulate,
                unsigned long long* __restrict__ stats, float* __restrict__ mom_out) {
    constexpr int NT = NW * 32;
    constexpr int TB = 32 * BPL;
    constexpr int NG = NW;   // groups of 32 samples per chunk
    extern __shared__ double smem_d[];
    const DevScene& S = *Sp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncomp = 0;   // this kernel now serves the direct Bremsstrahlung path only; line models run in cb2_emission_warp.cu
    // dynamic shared memory: fp64 per-ray accumulators [NT*BPL], records, Bremsstrahlung records, per-group bin ranges
    double* racc = smem_d;
    double* mom = smem_d + (size_t)NT * BPL;                            // [k_pad] (BREMS == 3 only)
    const int k_pad = BREMS == 3 ? S.brems.k_pad : 0;
    float4* rec = reinterpret_cast<float4*>(mom + k_pad);
    float4* brec = rec + (size_t)2 * ncomp * NT;
    int* s_rng = reinterpret_cast<int*>(brec + ((BREMS == 1 || BREMS == 2) ? 4 * NT : 0));   // [2][ncomp][NG][2]
    const int rng_stride = 2 * ncomp * NG;

    for (int i = tid; i < 2 * rng_stride; i += NT) s_rng[i] = (i & 1) ? INT_MIN : INT_MAX;
#pragma unroll
    for (int j = 0; j < BPL; j++) racc[warp * TB + 32 * j + lane] = 0.0;
    if (BREMS == 3)
        for (int i = tid; i < k_pad; i += NT) mom[i] = 0.0;

    const int64_t ray = blockIdx.x;
    const double ox = rays.origin[3 * ray], oy = rays.origin[3 * ray + 1], oz = rays.origin[3 * ray + 2];
    const double dwx = rays.direction[3 * ray], dwy = rays.direction[3 * ray + 1], dwz = rays.direction[3 * ray + 2];
    SampleIn in;
    {
        // ray direction in plasma space (direction.transform(local_to_plasma), normalised inside doppler_shift)
        const double d0 = xform_row(S.w2p, dwx, dwy, dwz, false), d1 = xform_row(S.w2p + 4, dwx, dwy, dwz, false),
                     d2 = xform_row(S.w2p + 8, dwx, dwy, dwz, false);
        const double dl = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        in.dx = (float)(d0 / dl); in.dy = (float)(d1 / dl); in.dz = (float)(d2 / dl);
    }

    float acc[BPL];
#pragma unroll
    for (int j = 0; j < BPL; j++) acc[j] = 0.f;

    // Bremsstrahlung per-bin constants (1/lambda, 2 log2(1/lambda), log10(lambda) - lref) for the one-point rule
    float b_rho[BREMS == 1 ? BPL : 1], b_cb[BREMS == 1 ? BPL : 1], b_lp[BREMS == 1 ? BPL : 1];
    if (BREMS == 1) {
#pragma unroll
        for (int j = 0; j < BPL; j++) {
            const float4 t = __ldg(S.brems.bin_tab1 + (warp * TB + 32 * j + lane));
            b_rho[j] = t.x; b_cb[j] = t.y; b_lp[j] = t.z;
        }
    }

    unsigned long long n_samples = 0, n_gauss = 0, n_brems = 0, n_lorentz = 0;
    unsigned ood = 0;
    int parity = 0;
    __syncthreads();

    const int64_t s_begin = rays.seg_offset[ray], s_end = rays.seg_offset[ray + 1];
    for (int64_t sg = s_begin; sg < s_end; sg++) {
        // NumericalIntegrator.integrate [raysect]: start_point = far end of the segment, end_point = near end, both taken to
        // plasma space; float64 with the reference's operation order so that step positions are bit-identical
        const double t0 = rays.seg_t0[sg], t1 = rays.seg_t1[sg];
        const double swx = __dadd_rn(ox, __dmul_rn(t1, dwx)), swy = __dadd_rn(oy, __dmul_rn(t1, dwy)), swz = __dadd_rn(oz, __dmul_rn(t1, dwz));
        const double ewx = __dadd_rn(ox, __dmul_rn(t0, dwx)), ewy = __dadd_rn(oy, __dmul_rn(t0, dwy)), ewz = __dadd_rn(oz, __dmul_rn(t0, dwz));
        const double sx = xform_row(S.w2p, swx, swy, swz, true), sy = xform_row(S.w2p + 4, swx, swy, swz, true), sz = xform_row(S.w2p + 8, swx, swy, swz, true);
        double ivx = __dsub_rn(xform_row(S.w2p, ewx, ewy, ewz, true), sx), ivy = __dsub_rn(xform_row(S.w2p + 4, ewx, ewy, ewz, true), sy),
               ivz = __dsub_rn(xform_row(S.w2p + 8, ewx, ewy, ewz, true), sz);
        const double length = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(ivx, ivx), __dmul_rn(ivy, ivy)), __dmul_rn(ivz, ivz)));
        if (!(length > 0.0)) continue;
        ivx = __ddiv_rn(ivx, length); ivy = __ddiv_rn(ivy, length); ivz = __ddiv_rn(ivz, length);
        int iv = (int)ceil(__ddiv_rn(length, S.step));             // intervals = max(min_samples - 1, ceil(L / step))
        iv = max(iv, max(S.min_samples - 1, 1));
        const double h = __ddiv_rn(length, (double)iv);
        const float hf = (float)h;
        if (tid == 0) n_samples += (unsigned long long)iv + 1ull;

        for (int k0 = 0; k0 <= iv; k0 += NT) {
            // ---------------- STATE phase ----------------
            const int k = k0 + tid;
            const bool active = k <= iv;
            const double tk = __dmul_rn((double)k, h);
            const double pxd = __dadd_rn(sx, __dmul_rn(tk, ivx)), pyd = __dadd_rn(sy, __dmul_rn(tk, ivy)), pzd = __dadd_rn(sz, __dmul_rn(tk, ivz));
            in.x = (float)pxd; in.y = (float)pyd; in.z = (float)pzd;
            in.weight = active ? ((k == 0 || k == iv) ? 0.5f * hf : hf) : 0.f;
            {
                AxCtx ctx;
                float ne = 0.f, te = 0.f;
                if (active) {
                    ax_setup(S, pxd, pyd, pzd, ctx, ood);
                    ne = eval_scalar(S.ne, ctx, in.x, in.y, in.z);
                    te = eval_scalar(S.te, ctx, in.x, in.y, in.z);
                } else {
                    ctx.m = 0.f; ctx.tri = -1; ctx.in_lcfs = false;
                }
                RecWriter W;
                W.rec = rec; W.rng = s_rng + parity * rng_stride; W.nt = NT; W.tid = tid; W.ng = NG; W.group = warp;
                W.bins = S.bins; W.gauss_evals = 0; W.lorentz_evals = 0;
                if (ncomp > 0) sample_lines(S, in, ctx, ne, te, W, ood);
                n_gauss += W.gauss_evals;
                n_lorentz += W.lorentz_evals;
                if (BREMS == 1 || BREMS == 2) sample_brems(S, in, ctx, ne, te, brec, NT, tid, n_brems, ood);
                if (BREMS == 3) sample_brems_moments(S, in, ctx, ne, te, mom, lane, n_brems, ood);
            }
            __syncthreads();
            // ---------------- BIN phase ----------------
            for (int i = tid; i < rng_stride; i += NT) s_rng[(parity ^ 1) * rng_stride + i] = (i & 1) ? INT_MIN : INT_MAX;
            const int nact = min(NT, iv - k0 + 1);
            const int* rng = s_rng + parity * rng_stride;
            // LINES: lanes = samples.  The union of a component's bin ranges over the chunk is cut into 32-bin windows,
            // dealt round-robin to the warps.
            for (int c = 0; c < ncomp; c++)
                line_windows<NW, NG, NT>(rec + (size_t)(2 * c) * NT, rec + (size_t)(2 * c + 1) * NT, rng + 2 * c * NG, nact, c,
                                         S.comps[c].c0_int, S.bins, racc, warp, lane, S.comps[c].type, S.lorentz_tab, S.lorentz_phi_inf);
            if (BREMS == 1 || BREMS == 2) {
                const float4* btab = S.brems.bin_tab;
                const float4* btab1 = S.brems.bin_tab1;
                const int nq = S.brems.nq;
                for (int s = 0; s < nact; s++) {
                    const float4 r0 = brec[s];
                    if (r0.x == 0.f) continue;
                    const float na2 = -r0.x;
                    const int flags = __float_as_int(r0.y);
                    if (flags < 8) {
                        // one-point rule.  Pieces (window crossing a knot of the Gaunt table's u grid) are contiguous in bin
                        // index: piece 0 = bins [bc1, end), piece 1 = [bc2, bc1), piece 2 = [0, bc2).  A warp tile lies in
                        // one piece unless the knot falls inside it.
                        const int np = flags;
                        const int bc1 = __float_as_int(r0.z), bc2 = __float_as_int(r0.w);
                        int pf = 0, pl = 0;
                        if (np > 1) {
                            const int t_first = warp * TB, t_last = warp * TB + TB - 1;
                            pf = (t_first < bc1) ? ((np > 2 && t_first < bc2) ? 2 : 1) : 0;
                            pl = (t_last < bc1) ? ((np > 2 && t_last < bc2) ? 2 : 1) : 0;
                        }
                        if (pf == pl) {
                            const float4 pc = brec[(1 + pf) * NT + s];
#pragma unroll
                            for (int j = 0; j < BPL; j++) {
                                float rho, cb, lp;
                                if (BREMS == 1) { rho = b_rho[j]; cb = b_cb[j]; lp = b_lp[j]; }
                                else { const float4 t = __ldg(btab1 + (warp * TB + 32 * j + lane)); rho = t.x; cb = t.y; lp = t.z; }
                                acc[j] = fmaf(horner4(pc, lp), ex2_approx(fmaf(na2, rho, cb)), acc[j]);
                            }
                        } else {
                            const float4 p0 = brec[NT + s], p1 = brec[2 * NT + s], p2 = brec[3 * NT + s];
#pragma unroll
                            for (int j = 0; j < BPL; j++) {
                                float rho, cb, lp;
                                if (BREMS == 1) { rho = b_rho[j]; cb = b_cb[j]; lp = b_lp[j]; }
                                else { const float4 t = __ldg(btab1 + (warp * TB + 32 * j + lane)); rho = t.x; cb = t.y; lp = t.z; }
                                const int bin = warp * TB + 32 * j + lane;
                                float4 pc = p0;
                                if (bin < bc1) pc = p1;
                                if (np > 2 && bin < bc2) pc = p2;
                                acc[j] = fmaf(horner4(pc, lp), ex2_approx(fmaf(na2, rho, cb)), acc[j]);
                            }
                        }
                    } else {
                        // cold sample: nq-point Gauss-Legendre rule from the table, piece chosen per quadrature point
                        const int np = flags - 8;
                        const float4 p0 = brec[NT + s], p1 = brec[2 * NT + s], p2 = brec[3 * NT + s];
                        const int xc1 = np > 1 ? __float_as_int(r0.z) : -1, xc2 = np > 2 ? __float_as_int(r0.w) : -1;
#pragma unroll
                        for (int j = 0; j < BPL; j++) {
                            const int bin = warp * TB + 32 * j + lane;
                            const float4* tb = btab + (size_t)bin * nq;
                            float4 pc = p0;
                            if (bin < xc1) pc = p1;
                            if (bin < xc2) pc = p2;
                            float v = 0.f;
                            for (int q = 0; q < nq; q++) {
                                const float4 t = __ldg(tb + q);
                                v = fmaf(t.w * horner4(pc, t.z), ex2_approx(fmaf(na2, t.x, t.y)), v);
                            }
                            acc[j] += v;
                        }
                    }
                }
            }
            // fp32 chunk sums -> fp64 per-ray accumulators (shared memory, each bin owned by exactly one lane)
#pragma unroll
            for (int j = 0; j < BPL; j++) {
                if (acc[j] != 0.f) atomicAdd(&racc[warp * TB + 32 * j + lane], (double)acc[j]);   // narrow path adds concurrently
                acc[j] = 0.f;
            }
            parity ^= 1;
            __syncthreads();
        }
    }

    // write the ray's spectrum: lane-consecutive bins -> coalesced 128-byte rows
#pragma unroll
    for (int j = 0; j < BPL; j++) {
        const int bin = warp * TB + 32 * j + lane;
        if (bin < S.bins) {
            const size_t idx = (size_t)ray * S.bins + bin;
            const double v = scale * racc[bin];
            if (out_f64) {
                double* p = (double*)out + idx;
                *p = (accumulate ? *p : 0.0) + v;
            } else {
                float* p = (float*)out + idx;
                *p = (float)((accumulate ? (double)*p : 0.0) + v);
            }
        }
    }

    if (BREMS == 3) {
        // the ray's moment row (fp32) for the contraction kernel; the last __syncthreads of the chunk loop made it final
        float* row = mom_out + (size_t)ray * k_pad;
        for (int i = tid; i < k_pad; i += NT) row[i] = (float)mom[i];
    }

    if (stats) {
        // warp-reduce the work counters, one atomic per warp
        unsigned long long oodl = ood;
        for (int off = 16; off > 0; off >>= 1) {
            n_gauss += __shfl_down_sync(FULL, n_gauss, off);
            n_brems += __shfl_down_sync(FULL, n_brems, off);
            n_lorentz += __shfl_down_sync(FULL, n_lorentz, off);
            oodl += __shfl_down_sync(FULL, oodl, off);
        }
        if (lane == 0) {
            if (n_gauss) atomicAdd(stats + 1, n_gauss);
            if (n_brems) atomicAdd(stats + 3, n_brems);
            if (n_lorentz) atomicAdd(stats + 2, n_lorentz);
            if (oodl) atomicAdd(stats + 5, oodl);
        }
        if (tid == 0 && n_samples) atomicAdd(stats + 0, n_samples);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// launch configuration
// ------------------------------------------------------------------------------------------------------------------
static size_t emission_smem_bytes(int nw, int bpl, int n_comp, const DevBrems& b) {
    const size_t nt = (size_t)nw * 32;
    const bool direct = b.present && b.mode != 3;
    const size_t mom = (b.present && b.mode == 3) ? (size_t)b.k_pad * sizeof(double) : 0;
    return nt * bpl * sizeof(double) + mom + ((size_t)2 * n_comp * nt + (direct ? 4 * nt : 0)) * sizeof(float4) +
           (size_t)2 * 2 * n_comp * nw * sizeof(int);
}

// The (NW, BPL) instances are compiled in CB2_N_GROUPS translation units from this one source (-DCB2_GROUP=g) so that the
// build parallelises; group 0 also holds the dispatcher and the helper kernels.
#ifndef CB2_GROUP
#define CB2_GROUP -1   // single translation unit: everything
#endif
#define CB2_IN_GROUP(g) (CB2_GROUP < 0 || CB2_GROUP == (g))

template <int NW, int BPL>
int launch_cfg(const cb2_scene* sc, const DevRays& rays, void* out, int out_f64, double scale, int accumulate,
               unsigned long long* stats, float* mom, cudaStream_t st) {
    const DevScene& S = sc->host;
    const int NT = NW * 32;
    const size_t smem = emission_smem_bytes(NW, BPL, 0, S.brems);
    const int mode = !S.brems.present ? 0 : (S.brems.mode == 3 ? 3 : (BPL <= 8 ? 1 : 2));
    if (rays.n_rays > 0x7fffffffLL) return cb2_fail(CB2_ERR_VALUE, "too many rays for one launch");
    dim3 grid((unsigned)rays.n_rays), block(NT);
#define CB2_LAUNCH(MODE)                                                                                                   \
    do {                                                                                                                   \
        auto kern = emission_kernel<NW, BPL, MODE>;                                                                        \
        if (smem > 48 * 1024) CB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, block, smem, st>>>(sc->dev, rays, out, out_f64, scale, accumulate, stats, mom);                       \
    } while (0)
    if (smem > 200 * 1024) return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "too many line components for shared memory (%zu bytes)", smem);
    if (mode == 1) CB2_LAUNCH(1);
    else if (mode == 2) CB2_LAUNCH(2);
    else return cb2_fail(CB2_ERR_RUNTIME, "internal error: the CTA-phased kernel only serves the direct Bremsstrahlung path");
#undef CB2_LAUNCH
    return cb2_cuda_check(cudaGetLastError(), "emission_kernel launch");
}

#define CB2_INSTANCES(X) X(4, 1, 0) X(4, 2, 0) X(4, 4, 0) X(4, 8, 0) X(8, 4, 1) X(8, 8, 1) X(8, 16, 2) X(2, 8, 2) X(1, 16, 3) X(4, 16, 3) X(2, 16, 3)
#define CB2_SIG(NW, BPL) \
    int launch_cfg<NW, BPL>(const cb2_scene*, const DevRays&, void*, int, double, int, unsigned long long*, float*, cudaStream_t);
#if CB2_GROUP >= 0
#define CB2_DECL(NW, BPL, G) extern template CB2_SIG(NW, BPL)
CB2_INSTANCES(CB2_DECL)
#endif
#define CB2_DEFN(NW, BPL) template CB2_SIG(NW, BPL)
#define CB2_SKIP(NW, BPL)
#if CB2_GROUP < 0 || CB2_GROUP == 0
#define CB2_INST_0 CB2_DEFN
#else
#define CB2_INST_0 CB2_SKIP
#endif
#if CB2_GROUP < 0 || CB2_GROUP == 1
#define CB2_INST_1 CB2_DEFN
#else
#define CB2_INST_1 CB2_SKIP
#endif
#if CB2_GROUP < 0 || CB2_GROUP == 2
#define CB2_INST_2 CB2_DEFN
#else
#define CB2_INST_2 CB2_SKIP
#endif
#if CB2_GROUP < 0 || CB2_GROUP == 3
#define CB2_INST_3 CB2_DEFN
#else
#define CB2_INST_3 CB2_SKIP
#endif
#define CB2_INST(NW, BPL, G) CB2_INST_##G(NW, BPL)
CB2_INSTANCES(CB2_INST)

#if CB2_IN_GROUP(0)
int cb2_emission_config(cb2_scene* sc) {
    const DevBrems& B = sc->host.brems;
    {
        // two-kernel line path (every scene's line models run through it): 4 warps per ray in the bin kernel; the warps'
        // private accumulators are fp64 while six CTAs still fit an SM (<= 1024 bins), else fp32 (a warp adds <= ~40 group
        // sums per bin and component; the sum over the warps and the frame stay fp64)
        int nw = 4, f64 = cb2_warp_smem_bytes(4, 1, sc->host.bins) <= 36 * 1024;
        if (cb2_warp_smem_bytes(nw, f64, sc->host.bins) > 200 * 1024) nw = 2;
        if (cb2_warp_smem_bytes(nw, f64, sc->host.bins) > 200 * 1024)
            return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "spectral_bins = %d is too large for the per-ray accumulators", sc->host.bins);
        if (const char* e = getenv("CB2_NW")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 8) nw = v; }
        if (const char* e = getenv("CB2_ACC")) f64 = !strcmp(e, "f64");
        sc->bin_nw = nw;
        sc->acc_f64 = f64;
        sc->host.bins_padded = (sc->host.bins + 31) & ~31;
    }
    if (!B.present || B.mode == 3) {
        sc->warp_kernel = 1;
        sc->nw = sc->bin_nw;
        sc->bpl = 0;
        return CB2_OK;
    }
    // direct Bremsstrahlung: the CTA-phased kernel evaluates the continuum only (lines_off), the lines take the two-kernel path
    sc->warp_kernel = 0;
    // (warps per CTA, bins per lane) instances in order of preference per spectral size; the per-sample records live in
    // shared memory (32 B x components x samples per chunk), so scenes with many components fall back to fewer warps
    static const int cand[][2] = {{4, 1}, {4, 2}, {4, 4}, {2, 8}, {1, 16}, {8, 4}, {4, 8}, {2, 16}, {8, 8}, {4, 16}, {8, 16}};
    const int bins = sc->host.bins;
    const size_t budget = 100 * 1024;
    int nw = 0, bpl = 0;
    size_t best = (size_t)-1;
    for (auto& c : cand) {
        if (c[0] * 32 * c[1] < bins) continue;
        const size_t need = emission_smem_bytes(c[0], c[1], 0, sc->host.brems);
        if (need <= budget) { nw = c[0]; bpl = c[1]; break; }          // first (preferred) instance that fits
        if (need < best) { best = need; nw = c[0]; bpl = c[1]; }        // otherwise the leanest one
    }
    if (!nw) return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "spectral_bins > 4096 per launch is not supported yet (got %d)", bins);
    if (emission_smem_bytes(nw, bpl, 0, sc->host.brems) > 200 * 1024)
        return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "%d spectral bins do not fit the continuum kernel's shared memory", bins);
    // tuning override (experiments): CB2_NW x CB2_BPL must cover the bins and be an instantiated pair
    const char *env_nw = getenv("CB2_NW"), *env_bpl = getenv("CB2_BPL");
    if (env_nw && env_bpl) {
        const int enw = atoi(env_nw), ebpl = atoi(env_bpl);
        if (enw * 32 * ebpl >= bins) { nw = enw; bpl = ebpl; }
    }
    sc->nw = nw;
    sc->bpl = bpl;
    return CB2_OK;
}

static int emission_batch(cb2_scene* sc, const DevRays& rays, void* out, int out_f64, double scale, int accumulate,
                          unsigned long long* stats, float* mom, cudaStream_t st);

int cb2_launch_emission(cb2_scene* sc, const DevRays& rays, void* out, int out_f64, double scale, int accumulate,
                        unsigned long long* stats, cudaStream_t st) {
    if (sc->warp_kernel) return cb2_launch_emission_warp(sc, rays, out, out_f64, scale, accumulate, stats, 1, st);
    const DevBrems& B = sc->host.brems;
    const bool moments = B.present && B.mode == 3;
    if (!moments) {
        // direct Bremsstrahlung: lines (and the flat TotalRadiatedPower term) through the two-kernel path, then the continuum
        // kernel adds to the same rows; the continuum kernel counts the samples
        int acc = accumulate;
        if (sc->host.n_comp > 0 || sc->host.has_flat) {
            const int rc = cb2_launch_emission_warp(sc, rays, out, out_f64, scale, accumulate, stats, 0, st);
            if (rc != CB2_OK) return rc;
            acc = 1;
        }
        return emission_batch(sc, rays, out, out_f64, scale, acc, stats, nullptr, st);
    }
    // moment matrix [rays][k_pad] fp32: grow-only, owned by the scene handle (one stream per handle); rays are processed in
    // batches that bound it to ~1.5 GB
    const int64_t batch = std::min(cb2_moment_batch(B.k_pad), rays.n_rays);
    const size_t need = (size_t)batch * B.k_pad * sizeof(float);
    if (need > sc->mom_bytes) {
        CB2_CUDA(cudaStreamSynchronize(st));
        if (sc->mom) cudaFree(sc->mom);
        sc->mom = nullptr;
        sc->mom_bytes = 0;
        CB2_CUDA(cudaMalloc((void**)&sc->mom, need));
        sc->mom_bytes = need;
    }
    const size_t esz = out_f64 ? sizeof(double) : sizeof(float);
    for (int64_t r0 = 0; r0 < rays.n_rays; r0 += batch) {
        DevRays sub = rays;
        sub.n_rays = std::min(batch, rays.n_rays - r0);
        sub.origin = rays.origin + 3 * r0;
        sub.direction = rays.direction + 3 * r0;
        sub.seg_offset = rays.seg_offset + r0;      // entries are absolute segment indices
        void* o = (char*)out + (size_t)r0 * sc->host.bins * esz;
        int rc = emission_batch(sc, sub, o, out_f64, scale, accumulate, stats, sc->mom, st);
        if (rc != CB2_OK) return rc;
        rc = cb2_launch_contract(sc->mom, B.phi, sub.n_rays, B.k_pad, B.n_pad, sc->host.bins, o, out_f64, scale, st);
        if (rc != CB2_OK) return rc;
    }
    return CB2_OK;
}

static int emission_batch(cb2_scene* sc, const DevRays& rays, void* out, int out_f64, double scale, int accumulate,
                          unsigned long long* stats, float* mom, cudaStream_t st) {
#define CB2_CASE(NW, BPL, G) \
    if (sc->nw == NW && sc->bpl == BPL) return launch_cfg<NW, BPL>(sc, rays, out, out_f64, scale, accumulate, stats, mom, st);
    CB2_INSTANCES(CB2_CASE)
#undef CB2_CASE
    return cb2_fail(CB2_ERR_NOT_IMPLEMENTED, "no kernel instance for nw=%d bpl=%d", sc->nw, sc->bpl);
}

// ------------------------------------------------------------------------------------------------------------------
// per-point plasma state (parity tests of the flattened function tree)
// ------------------------------------------------------------------------------------------------------------------
__global__ void sample_state_kernel(const DevScene* __restrict__ Sp, const double* __restrict__ pts, int64_t n, double* __restrict__ out,
                                    int in_plasma_space) {
    const DevScene& S = *Sp;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
    double pd[3] = {px, py, pz};
    if (!in_plasma_space)
        for (int k = 0; k < 3; k++) pd[k] = xform_row(S.w2p + 4 * k, px, py, pz, true);
    const float p[3] = {(float)pd[0], (float)pd[1], (float)pd[2]};
    AxCtx ctx;
    unsigned ood = 0;
    ax_setup(S, pd[0], pd[1], pd[2], ctx, ood);
    const int w = 2 + 5 * S.n_species + 3;
    double* o = out + i * w;
    const double inv_scale = 1.0 / CB2_DENSITY_SCALE;
    o[0] = (double)eval_scalar(S.ne, ctx, p[0], p[1], p[2]) * inv_scale;
    o[1] = (double)eval_scalar(S.te, ctx, p[0], p[1], p[2]);
    for (int s = 0; s < S.n_species; s++) {
        o[2 + 5 * s] = (double)eval_scalar(S.species[s].density, ctx, p[0], p[1], p[2]) * inv_scale;
        o[3 + 5 * s] = (double)eval_scalar(S.species[s].temperature, ctx, p[0], p[1], p[2]);
        const float3 v = eval_vector(S.species[s].velocity, ctx);
        o[4 + 5 * s] = v.x; o[5 + 5 * s] = v.y; o[6 + 5 * s] = v.z;
    }
    const float3 b = eval_b_field(S, ctx);
    o[2 + 5 * S.n_species] = b.x; o[3 + 5 * S.n_species] = b.y; o[4 + 5 * S.n_species] = b.z;
}

// Beam.density / Beam.direction probe: out[n][4]
__global__ void beam_sample_kernel(const DevScene* __restrict__ Sp, const double* __restrict__ pts, int64_t n, double* __restrict__ out) {
    const DevScene& S = *Sp;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = (float)pts[3 * i], y = (float)pts[3 * i + 1], z = (float)pts[3 * i + 2];
    const float3 dv = beam_direction(S.beam, x, y, z);
    out[4 * i] = (double)beam_density(S.beam, x, y, z) * (1.0 / CB2_DENSITY_SCALE);
    out[4 * i + 1] = dv.x; out[4 * i + 2] = dv.y; out[4 * i + 3] = dv.z;
}

int cb2_launch_beam_sample(const cb2_scene* sc, const double* beam_points_dev, int64_t n, double* out_dev, cudaStream_t st) {
    const int bs = 128;
    beam_sample_kernel<<<(unsigned)((n + bs - 1) / bs), bs, 0, st>>>(sc->dev, beam_points_dev, n, out_dev);
    int rc = cb2_cuda_check(cudaGetLastError(), "beam_sample_kernel launch");
    if (rc == CB2_OK) rc = cb2_cuda_check(cudaStreamSynchronize(st), "beam_sample sync");
    return rc;
}

int cb2_launch_sample_state(const cb2_scene* sc, const double* points_dev, int64_t n, double* out_dev, int in_plasma_space, cudaStream_t st) {
    // the state probe always wants B and the poloidal direction: use a scene copy with the flags forced on
    DevScene tmp = sc->host;
    tmp.need_b = 1;
    tmp.need_pol = 1;
    DevScene* dtmp = nullptr;
    CB2_CUDA(cudaMalloc((void**)&dtmp, sizeof(DevScene)));
    int rc = cb2_cuda_check(cudaMemcpyAsync(dtmp, &tmp, sizeof(DevScene), cudaMemcpyHostToDevice, st), "cudaMemcpy(scene)");
    if (rc == CB2_OK) {
        const int bs = 128;
        sample_state_kernel<<<(unsigned)((n + bs - 1) / bs), bs, 0, st>>>(dtmp, points_dev, n, out_dev, in_plasma_space);
        rc = cb2_cuda_check(cudaGetLastError(), "sample_state_kernel launch");
        if (rc == CB2_OK) rc = cb2_cuda_check(cudaStreamSynchronize(st), "sample_state sync");
    }
    cudaFree(dtmp);
    return rc;
}
#endif  // CB2_IN_GROUP(0)
