// cb2_emission_warp.cu — K1: the line-emission hot path on sm_100a as two kernels per ray batch.
//
//   state_kernel  (K1a, latency-bound: table gathers)   one CTA per ray, warps take the ray's 32-sample groups round-robin,
//                 a thread owns one sample: position (fp64, reference operation order) -> (R, Z) -> psi_n / LCFS mask /
//                 blend weight / mesh triangle ONCE per sample (the reference re-walks the function tree for every
//                 quantity) -> species profiles -> PEC bicubics (cell search shared between models on the
//                 same knots) -> per line component (centre, width, amplitude) written as coalesced 128-byte rows;
//                 Bremsstrahlung goes to the ray's temperature-node moments (cb2_device.cuh; contraction cb2_contract.cu).
//                 The flattened scene travels as a __grid_constant__ kernel parameter (constant bank, no dependent loads).
//   bin_kernel    (K1b, FP32/SFU-bound)                 one CTA per ray, warps take the groups round-robin, lanes = samples:
//                 the union of the group's bin ranges is cut into 32-bin windows; lane l evaluates ITS sample at the
//                 window's bins in the lane-permuted order bin = r XOR l for register r, so that the 31-shuffle
//                 butterfly part[i] += shfl_xor(part[i + o], o) needs no selects and leaves window bin l on lane l, which
//                 adds it to the warp's private per-ray accumulator in shared memory (plain load/add/store, no atomics);
//                 tails use 16/8-bin windows, sub-bin lines a rolled 4-bin loop with the reference's lower = upper
//                 recurrence.  One barrier at the end, fp64 sum of the private accumulators, coalesced rows out.
//
// Why two kernels: fused, the hot instruction footprint (~100 KB) thrashes the 32 KB L1.5 instruction cache and the
// kernel is fetch-bound (ncu: no_instruction 4.1 stall cycles per issue, profiles/r1b_*); split, each half fits, the
// state half gets the whole L1 for table data and the bin half a small register footprint.  The records cost 384 B per
// live (group, component) through L2/HBM.
//
// Gaussian bin integrals (gaussian.pyx:40-90) in fp32 with RELATIVE accuracy (tools/proto_gauss_fp32.py):
//   sigma >= 0.98 bin: kb/sqrt(pi) e^{-m} sum_{n<=4} H_2n h^2n/(2n+1)! as exp2(-m') (s0 + m'(s1 + m'(s2 + m'(s3 + m' s4)))),
//                      one MUFU.EX2 + 7 FMA-pipe instructions per bin, no cancellation;
//   sub-bin lines:     differences of 1/2 erfc(|x|) along the bin edges.
// Bins beyond 7 sigma are not evaluated: e^{-24.5} = 2e-11 of the sample's peak, two decades under the acceptance floor
// (1e-9 of the ray's largest bin); the reference's own cut-off is 10 sigma and the work counter E still uses it.
// Follows: cherab/core/plasma/material.pyx:48-63, model/plasma/impact_excitation.pyx:78-100, recombination.pyx:78-100,
// model/lineshape/gaussian.pyx:40-139, doppler.pyx:29-59, multiplet.pyx:93-117, zeeman.pyx:113-365, stark.pyx:88-348,
// tools/equilibrium/efit.pyx:219-546, generomak/plasma/plasma.py:580-638, Raysect's NumericalIntegrator.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "cb2_device.cuh"

#define EVAL_CUTOFF_SIGMA 7.0f
#ifndef CB2_FUSED_MINB
#define CB2_FUSED_MINB 4
#endif

// per-sample state shared by the line models of one sample
struct LineCache {
    float lne, lte;
    int cur;                 // species whose (ni, ts, vd) are cached
    float ni, ts, vd;
    bool have_b;
    float bm, cos_sqr;
    float bfx, bfy, bfz;     // B in plasma space (the MSE multiplet needs v x B)
    int grid;                // PEC knot set whose cell is cached
    Cell2 cell;
};

// per-(sample, model) quantities every component of the model's line shape needs
struct ModelCtx {
    bool on;                 // the model emits at this sample and its shape has something to add
    float amp;               // radiance * trapezium weight / delta_wavelength
    float sigma_b;           // Gaussian sigma in bins
    float lam_b, eta;        // Stark: pseudo-Voigt FWHM in bins, Lorentzian fraction
    float dop;               // v.d / c
    float shift0;            // Doppler shift of the rest wavelength in bins
    bool bzero;              // |B| == 0: no splitting
    float a_pi, a_sigma;     // polarisation-weighted amplitudes
    float dl_plus, dl_minus; // sigma+- wavelength offsets (triplets, Stark)
    int ib; float tb;        // ZeemanStructure |B| interval
    float rnorm[3];          // ZeemanStructure ratio normalisation per polarisation group
};

__device__ __forceinline__ void need_b(const DevScene& S, const SampleIn& in, const AxCtx& ctx, LineCache& lc, unsigned& ood) {
    if (lc.have_b) return;
    if (S.b_kind != 0 && ctx.b_outside) ood++;
    const float3 bf = eval_b_field(S, ctx);
    lc.bfx = bf.x; lc.bfy = bf.y; lc.bfz = bf.z;
    lc.bm = sqrtf(bf.x * bf.x + bf.y * bf.y + bf.z * bf.z);
    const float c = lc.bm > 0.f ? (bf.x * in.dx + bf.y * in.dy + bf.z * in.dz) / lc.bm : 0.f;
    lc.cos_sqr = c * c;
    lc.have_b = true;
}

// BeamCXPEC.evaluate (openadas/rates/cx.pyx:104-142): 10^c0(log10 E) times the four linear factors, zero as soon as a partial
// product is <= 0; returns the rate in W m^3 scaled by 1e38.  args = log10 E (-inf: E <= 0), Ti, n_ion[1e19], Zeff, |B|.
__device__ __forceinline__ float cx_rate_eval(const DevCXRate& R, const float (&args)[5], unsigned& ood) {
    if (R.is_const) return exp10f(R.lconst);
    if (!(args[0] > -INFINITY)) return 0.f;
    float lq = 0.f, factor = 1.0f;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        float v;
        if (R.n[k] == 1) v = R.single[k];
        else if (R.extrapolate && k == 0) v = eval1d_quadratic(R.t[0], R.c[0], args[0]);      // 'quadratic' in log10 E, cx.pyx:96,98
        else {
            // 'none': clamped and counted; with extrapolation the four factors are 'nearest' (cx.pyx:97)
            // (a few fp32 ulps of slack: Zeff of a Z = 1 plasma must not count as below a table that starts at 1)
            if (!R.extrapolate && (args[k] < R.t[k].xmin - 4e-6f * fabsf(R.t[k].xmin) || args[k] > R.t[k].xmax + 4e-6f * fabsf(R.t[k].xmax))) ood++;
            int ci; float ct;
            locate1d(R.t[k], args[k], ci, ct);
            v = horner4(__ldg(R.c[k] + ci), ct);
        }
        if (k == 0) lq = v;
        else {
            factor *= v;
            if (!(factor > 0.f)) return 0.f;
        }
    }
    return exp10f(lq) * factor;
}

// BeamCXLine._beam_population (charge_exchange.pyx:241-292): charge-density weighted mean of the species' population
// coefficients of one excited metastable; neutral species are skipped
template <int AXONLY>
__device__ __forceinline__ float beam_population(const DevScene& S, const DevPopRate* P, const SampleIn& in, const AxCtx& ctx,
                                                 float density_sum, unsigned& ood) {
    float pop = 0.f, total_ne = 0.f;
    for (int sidx = 0; sidx < S.n_species; sidx++) {
        const DevSpecies& sp = S.species[sidx];
        if (sp.charge == 0) continue;
        const float target_ne = eval_scalar_t<AXONLY>(sp.density, ctx, in.x, in.y, in.z) * (float)sp.charge;
        const float ti = eval_scalar_t<AXONLY>(sp.temperature, ctx, in.x, in.y, in.z);
        const DevPopRate& R = P[sidx];
        float val;
        if (R.is_const) val = exp10f(R.lconst);
        else {
            const float3 v = eval_vector(sp.velocity, ctx);
            const float ivx = in.bvx - v.x, ivy = in.bvy - v.y, ivz = in.bvz - v.z;
            const float energy = (ivx * ivx + ivy * ivy + ivz * ivz) * 5.18213506e-9f;
            const float n_eq = density_sum / (float)sp.charge;
            val = 0.f;
            if (energy > 0.f && n_eq > 0.f && ti > 0.f) {
                const float le = log10f(energy), ln = log10f(n_eq) + 19.0f, lt = log10f(ti);
                const bool out_a = le < R.a.xmin || le > R.a.xmax || ln < R.a.ymin || ln > R.a.ymax, out_t = lt < R.tk.xmin || lt > R.tk.xmax;
                if (R.extrapolate && (out_a || out_t)) {
                    val = exp10f(eval2d_linear(R.a, le, ln) + eval1d_quadratic(R.tk, R.tc, lt));   // beam.pyx:73-84
                } else {
                    // the oracle counts every axis that leaves its table
                    if (le < R.a.xmin || le > R.a.xmax) ood++;
                    if (ln < R.a.ymin || ln > R.a.ymax) ood++;
                    if (out_t) ood++;
                    const Cell2 c = locate2d(R.a, le, ln);
                    int ci; float ct;
                    locate1d(R.tk, lt, ci, ct);
                    val = exp10f(eval2d(R.a, c) + horner4(__ldg(R.tc + ci), ct));
                }
            }
        }
        pop = fmaf(target_ne, val, pop);
        total_ne += target_ne;
    }
    return total_ne > 0.f ? pop / total_ne : 0.f;
}

// ExcitationLine / RecombinationLine .emission up to the add_line call (impact_excitation.pyx:86-100) plus the
// component-independent part of LineShapeModel.add_line
template <int AXONLY, int FEAT>
__device__ __forceinline__ void model_setup(const DevScene& S, const DevModel& M, const SampleIn& in, const AxCtx& ctx, float ne, float te,
                                            bool live, LineCache& lc, ModelCtx& mc, unsigned& ood) {
    mc.on = false;
    bool on = live;
    if (FEAT && M.kind == CB2_MODEL_BEAM_CX_LINE) on = in.weight > 0.f && in.donor > 0.f;    // no ne / te condition on the beam path
    if (on && M.species != lc.cur) {
        lc.cur = M.species;
        const DevSpecies& sp = S.species[lc.cur];
        lc.ni = eval_scalar_t<AXONLY>(sp.density, ctx, in.x, in.y, in.z);
        lc.ts = eval_scalar_t<AXONLY>(sp.temperature, ctx, in.x, in.y, in.z);
        const float3 v = eval_vector(sp.velocity, ctx);
        lc.vd = v.x * in.dx + v.y * in.dy + v.z * in.dz;   // velocity projected on the (unit) ray direction
    }
    on = on && (lc.ni > 0.f);
    if (!on) return;
    float radiance;
    if (FEAT && M.kind == CB2_MODEL_BEAM_CX_LINE) {
        // BeamCXLine.emission (charge_exchange.pyx:117-167), ground-state donor: radiance = 1/(4 pi) n_beam n_rec q_eff
        if (!(in.donor > 0.f) || lc.ni == 0.f || lc.ts == 0.f) return;
        const DevModelExt& X = *M.ext;
        float args[5] = {0.f, lc.ts, 0.f, 0.f, 0.f};
        float density_sum = 0.f;                              // sum Z^2 n over the plasma species (charge_exchange.pyx:266-268)
        bool tabulated = false;
        for (int k = 0; k < X.n_cx; k++) tabulated = tabulated || !X.cx[k].is_const;
        if (tabulated || X.n_cx > 1) {
            const float3 vr = eval_vector(S.species[M.species].velocity, ctx);
            const float ivx = in.bvx - vr.x, ivy = in.bvy - vr.y, ivz = in.bvz - vr.z;
            const float energy = (ivx * ivx + ivy * ivy + ivz * ivz) * 5.18213506e-9f;    // m_u / (2 e): (m/s)^2 -> eV/amu
            // Plasma.ion_density / z_effective (plasma/node.pyx:396-462)
            float n_ion = 0.f, snz = 0.f, snz2 = 0.f;
            for (int sidx = 0; sidx < S.n_species; sidx++) {
                const float n = eval_scalar_t<AXONLY>(S.species[sidx].density, ctx, in.x, in.y, in.z);
                const float zc = (float)S.species[sidx].charge;
                n_ion += n;
                snz = fmaf(n, zc, snz);
                snz2 = fmaf(n * zc, zc, snz2);
            }
            density_sum = snz2;
            if (tabulated) need_b(S, in, ctx, lc, ood);
            args[0] = energy > 0.f ? log10f(energy) : -INFINITY;                           // -inf: the rate is zero (cx.pyx:118-119)
            args[2] = n_ion;
            args[3] = snz > 0.f ? snz2 / snz : 0.f;
            args[4] = lc.bm;
        }
        float rate = cx_rate_eval(X.cx[0], args, ood);                                   // W m^3 * 1e38
        if (X.n_cx > 1) {
            float total_population = 1.0f;
            for (int k = 1; k < X.n_cx; k++) {
                const float population = beam_population<AXONLY>(S, X.pop + (size_t)(k - 1) * S.n_species, in, ctx, density_sum, ood);
                rate = fmaf(population, cx_rate_eval(X.cx[k], args, ood), rate);
                total_population += population;
            }
            rate /= total_population;
        }
        radiance = RECIP_4_PI * rate * in.donor * lc.ni;
    } else if (FEAT && M.kind == CB2_MODEL_THERMAL_CX_LINE) {
        // radiance = 1/(4 pi) n_receiver sum_donors n_donor q_donor(ne, te, T_donor)   (thermal_cx.pyx:103-111)
        const DevModelExt& X = *M.ext;
        float weighted = 0.f;
        for (int k = 0; k < X.n_donors; k++) {
            const float nd = eval_scalar_t<AXONLY>(S.species[X.donor_species[k]].density, ctx, in.x, in.y, in.z);
            float lq = X.donor_lrate[k];
            if (X.donor_tab[k]) {
                // ThermalCXPEC.evaluate (pec.pyx:186-194): zero for a non-positive donor temperature
                const float td = eval_scalar_t<AXONLY>(S.species[X.donor_species[k]].temperature, ctx, in.x, in.y, in.z);
                if (td > 0.f) {
                    bool inside;
                    lq = eval3d(X.donor_t3[k], lc.lne, lc.lte, log10f(td), inside);
                    if (!inside && !X.donor_extrapolate[k]) ood++;
                } else lq = -INFINITY;
            }
            weighted = fmaf(nd, exp10f(lq), weighted);
        }
        radiance = RECIP_4_PI * weighted * lc.ni;
    } else {
        // radiance = 1/(4 pi) PEC ne ni (impact_excitation.pyx:99): exp10 of (log PEC + 38) times (ne ni 1e-38)
        float lp;
        if (M.pec_const) lp = M.pec_value;
        else {
            if (M.pec_grid != lc.grid) { lc.cell = locate2d(M.pec, lc.lne, lc.lte); lc.grid = M.pec_grid; }
            if (!lc.cell.inside && !M.pec_extrapolate) ood++;
            lp = eval2d(M.pec, lc.cell);
        }
        radiance = RECIP_4_PI * exp10f(lp) * ne * lc.ni;
    }
    mc.amp = radiance * in.weight * M.inv_delta;
    if (!(mc.amp > 0.f)) return;
    mc.dop = lc.vd * M.inv_c;                                  // doppler_shift: lambda (1 + v.d/c), doppler.pyx:29-44
    mc.shift0 = M.wavelength * mc.dop * M.inv_delta;
    const float ts = lc.ts;
    if (M.shape == CB2_SHAPE_STARK) {
        // StarkBroadenedLine.add_line (stark.pyx:251-348): does NOT return early on ts <= 0
        const float SIGMA2FWHM = 2.3548200450309493f;
        const float fl = M.param[0] * powf(ne, M.param[1]) / powf(te, M.param[2]);          // nm (ne in 1e19 m^-3 folded in)
        const float fg = ts > 0.f ? SIGMA2FWHM * M.sigma_coef * sqrtf(ts) / M.inv_delta : 0.f; // nm
        if (fl == 0.f && fg == 0.f) return;
        float full;
        if (fg <= fl) {
            const float r = fg / fl;
            full = fl * (1.f + r * r * (0.57575f + r * (0.37902f + r * (-0.42519f + r * (-0.31525f + r * 0.31718f)))));
        } else {
            const float r = fl / fg;
            full = fg * (1.f + r * (0.15882f + r * (1.04388f + r * (-1.38281f + r * (0.46251f + r * (0.82325f + r * -0.58026f))))));
        }
        float sigma = full / SIGMA2FWHM, eta;
        const float l2t = fl / full;
        if (l2t < 0.01f) { eta = 0.f; full = 0.f; }
        else if (l2t > 0.999f) { eta = 1.f; sigma = 0.f; }
        else {
            const float lg = logf(l2t);
            eta = expf(5.14820e-04f + lg * (1.38821e+00f + lg * (-9.60424e-02f + lg * (-3.83995e-02f + lg * (-7.40042e-03f + lg * -5.47626e-04f)))));
        }
        mc.sigma_b = sigma * M.inv_delta;
        mc.lam_b = full * M.inv_delta;
        mc.eta = eta;
    } else {
        // all Gaussian-family shapes return before touching the spectrum if ts <= 0 (gaussian.pyx:127-129)
        if (!(ts > 0.f)) return;
        mc.sigma_b = M.sigma_coef * sqrtf(ts);                 // thermal_broadening, doppler.pyx:48-59, in bins
        if (M.shape == CB2_SHAPE_GAUSSIAN || M.shape == CB2_SHAPE_MULTIPLET) { mc.on = true; return; }
    }
    // Zeeman family and Stark: field strength and angle to the line of sight (zeeman.pyx:125-131)
    need_b(S, in, ctx, lc, ood);
    mc.bzero = lc.bm == 0.f;
    const float cos_sqr = lc.cos_sqr, sin_sqr = 1.0f - cos_sqr;
    mc.a_pi = 0.5f * sin_sqr * mc.amp;
    mc.a_sigma = (0.25f * sin_sqr + 0.5f * cos_sqr) * mc.amp;
    if (M.shape == CB2_SHAPE_PARAM_ZEEMAN) {
        mc.sigma_b *= sqrtf(1.0f + M.param[1] * M.param[1] * powf(ts, 2.0f * M.param[2]));
        mc.dl_plus = 0.5f * M.param[0] * lc.bm;                // zeeman.pyx:260-264
        mc.dl_minus = -mc.dl_plus;
    } else if (M.shape == CB2_SHAPE_ZEEMAN_TRIPLET || M.shape == CB2_SHAPE_STARK) {
        // hc/(hc/l0 -+ muB B) - l0 = +- l0 e/(1 -+ e), e = muB B l0 / hc   (zeeman.pyx:152-158)
        const float e = BOHR_MAGNETON * lc.bm * M.wavelength * (1.0f / HC_EV_NM_F);
        mc.dl_plus = M.wavelength * e / (1.0f - e);
        mc.dl_minus = -M.wavelength * e / (1.0f + e);
    } else if (M.shape == CB2_SHAPE_ZEEMAN_MULTIPLET && !mc.bzero) {  // zeeman.pyx:340-363, atomic/zeeman.pyx:87-129
        float fb = (lc.bm - M.b0) * M.inv_db;
        fb = fminf(fmaxf(fb, 0.f), (float)(M.n_b - 1));
        mc.ib = min((int)fb, M.n_b - 2);
        mc.tb = fb - (float)mc.ib;
        const int offs[4] = {0, M.n_pi, M.n_pi + M.n_sp, M.n_pi + M.n_sp + M.n_sm};
#pragma unroll
        for (int g = 0; g < 3; g++) {
            float rsum = 0.f;
            for (int k = offs[g]; k < offs[g + 1]; k++) {
                const float* r = M.zee_ratio + (size_t)k * M.n_b + mc.ib;
                rsum += fmaf(mc.tb, __ldg(r + 1) - __ldg(r), __ldg(r));
            }
            mc.rnorm[g] = rsum > 0.f ? 1.0f / rsum : 1.0f;
        }
    }
    mc.on = true;
}

// TotalRadiatedPower.emission (total_radiated_power.pyx:70-118): wavelength-independent radiance of one sample
template <int AXONLY>
__device__ __forceinline__ float total_radiated_power(const DevScene& S, const DevModel& M, const SampleIn& in, const AxCtx& ctx, float ne,
                                                      float lne, float lte, unsigned& ood) {
    const DevModelExt& X = *M.ext;
    const float ni = eval_scalar_t<AXONLY>(S.species[X.line_rad].density, ctx, in.x, in.y, in.z);
    const float ni_upper = eval_scalar_t<AXONLY>(S.species[X.recom].density, ctx, in.x, in.y, in.z);
    float nhyd = 0.f;
    for (int k = 0; k < X.n_hyd; k++) nhyd += eval_scalar_t<AXONLY>(S.species[X.hyd[k]].density, ctx, in.x, in.y, in.z);
    const float dens[3] = {ne * ni, ne * ni_upper, nhyd * ni_upper};
    const bool use[3] = {ni > 0.f, ni_upper > 0.f, ni_upper > 0.f && nhyd > 0.f};
    float power = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (!X.has[k] || !use[k]) continue;
        float lp;
        if (X.is_const[k]) lp = X.lconst[k];
        else {
            const Cell2 c = locate2d(X.tab[k], lne, lte);
            if (!c.inside && !X.extrapolate[k]) ood++;
            lp = eval2d(X.tab[k], c);
        }
        power = fmaf(exp10f(lp), dens[k], power);        // rates carry +38 in the exponent, densities 1e-19 each
    }
    return RECIP_4_PI * power * S.inv_range * in.weight;
}

// BeamEmissionLine.emission (beam_emission.pyx:100-176) + BeamEmissionMultiplet.add_line (mse.pyx:62-135) at one sample:
// amplitude (radiance * weight / delta), Doppler shift of the beam atoms and Stark splitting, both in bins
template <int AXONLY>
__device__ __forceinline__ void beam_emission_setup(const DevScene& S, const DevModel& M, const SampleIn& in, const AxCtx& ctx, float ne,
                                                    float te, LineCache& lc, float& amp, float& shift_b, float& split_b, unsigned& ood) {
    amp = 0.f; shift_b = 0.f; split_b = 0.f;
    if (!(in.weight > 0.f) || !(in.donor > 0.f) || !(ne > 0.f) || !(te > 0.f)) return;
    const DevModelExt& X = *M.ext;
    float density_sum = 0.f;
    for (int k = 0; k < X.n_bes; k++) {
        const float zc = (float)X.bes_charge[k];
        density_sum = fmaf(zc * zc, eval_scalar_t<AXONLY>(S.species[X.bes_species[k]].density, ctx, in.x, in.y, in.z), density_sum);
    }
    float rate = 0.f;
    for (int k = 0; k < X.n_bes; k++) {
        const int zc = X.bes_charge[k];
        if (zc == 0) continue;                                       // no beam emission data for neutrals
        const DevSpecies& sp = S.species[X.bes_species[k]];
        const float target_ne = eval_scalar_t<AXONLY>(sp.density, ctx, in.x, in.y, in.z) * (float)zc;
        if (!(target_ne > 0.f)) continue;
        float lq;
        if (X.bes_const[k]) lq = X.bes_lconst[k];
        else {
            const float ti = eval_scalar_t<AXONLY>(sp.temperature, ctx, in.x, in.y, in.z);
            const float3 v = eval_vector(sp.velocity, ctx);
            const float ivx = in.bvx - v.x, ivy = in.bvy - v.y, ivz = in.bvz - v.z;
            const float energy = (ivx * ivx + ivy * ivy + ivz * ivz) * 5.18213506e-9f;
            const float n_eq = density_sum / (float)zc;
            if (!(energy > 0.f) || !(n_eq > 0.f) || !(ti > 0.f)) continue;
            const float le = log10f(energy), ln = log10f(n_eq) + 19.0f, lt = log10f(ti);
            const Cell2 c = locate2d(X.bes_a[k], le, ln);
            const bool out_t = lt < X.bes_tk[k].xmin || lt > X.bes_tk[k].xmax;
            if (X.bes_extrapolate[k] && (!c.inside || out_t)) {
                lq = eval2d_linear(X.bes_a[k], le, ln) + eval1d_quadratic(X.bes_tk[k], X.bes_tc[k], lt);   // beam.pyx:209-221
            } else {
                if (!c.inside) ood++;
                if (out_t) ood++;
                int ci; float ct;
                locate1d(X.bes_tk[k], lt, ci, ct);
                lq = eval2d(X.bes_a[k], c) + horner4(__ldg(X.bes_tc[k] + ci), ct);
            }
        }
        rate = fmaf(target_ne, exp10f(lq), rate);
    }
    amp = RECIP_4_PI * in.donor * rate * in.weight * M.inv_delta;
    if (!(amp > 0.f)) { amp = 0.f; return; }
    need_b(S, in, ctx, lc, ood);
    const float cx = in.bvy * lc.bfz - in.bvz * lc.bfy, cy = in.bvz * lc.bfx - in.bvx * lc.bfz, cz = in.bvx * lc.bfy - in.bvy * lc.bfx;
    split_b = fabsf(2.77e-8f * sqrtf(cx * cx + cy * cy + cz * cz)) * M.inv_delta;            // STARK_SPLITTING_FACTOR, mse.pyx:33
    shift_b = M.wavelength * (in.bvx * in.dx + in.bvy * in.dy + in.bvz * in.dz) * M.inv_c * M.inv_delta;
}

// component k of model M at this sample: type (0 Gaussian, 1 modified Lorentzian), centre cf in bins relative to the
// component slot's integer origin, width (sigma or FWHM) in bins, amplitude; amp == 0 means "nothing to add"
__device__ __forceinline__ void model_component(const DevScene& S, const DevModel& M, int k, const ModelCtx& mc, int& type, float& cf,
                                                float& width, float& amp) {
    const DevComp& cs = S.comps[M.comp0 + k];
    type = cs.type;
    cf = cs.c0_frac + mc.shift0;
    width = mc.sigma_b;
    amp = 0.f;
    if (!mc.on) return;
    const bool pol_pi = M.polarisation != CB2_POL_SIGMA, pol_sigma = M.polarisation != CB2_POL_PI;
    switch (M.shape) {
    case CB2_SHAPE_GAUSSIAN: amp = mc.amp; return;
    case CB2_SHAPE_MULTIPLET:                                   // multiplet.pyx:108-115
        cf = cs.c0_frac + __ldg(M.mult_lambda + k) * mc.dop * M.inv_delta;
        amp = mc.amp * __ldg(M.mult_ratio + k);
        return;
    case CB2_SHAPE_ZEEMAN_MULTIPLET:
        if (!mc.bzero) {
            const int g = k < M.n_pi ? 0 : (k < M.n_pi + M.n_sp ? 1 : 2);
            if (!(g == 0 ? pol_pi : pol_sigma)) return;
            const float* r = M.zee_ratio + (size_t)k * M.n_b + mc.ib;
            const float* l = M.zee_dlambda + (size_t)k * M.n_b + mc.ib;
            const float ratio = fmaf(mc.tb, __ldg(r + 1) - __ldg(r), __ldg(r)) * mc.rnorm[g];
            const float dl = fmaf(mc.tb, __ldg(l + 1) - __ldg(l), __ldg(l));
            cf = cs.c0_frac + (dl + (M.wavelength + dl) * mc.dop) * M.inv_delta;
            amp = (g == 0 ? mc.a_pi : mc.a_sigma) * ratio;
            return;
        }
        // no splitting: single Gaussian, halved if a polarisation filter is set (zeeman.pyx:132-136)
        if (k == 0) amp = M.polarisation == CB2_POL_NO ? mc.amp : 0.5f * mc.amp;
        return;
    case CB2_SHAPE_ZEEMAN_TRIPLET:
    case CB2_SHAPE_PARAM_ZEEMAN:
    case CB2_SHAPE_STARK: {
        const int kk = (M.shape == CB2_SHAPE_STARK && k >= 3) ? k - 3 : k;   // 0 pi, 1 sigma+, 2 sigma-
        float a;
        if (mc.bzero) {
            if (kk != 0) return;
            a = M.polarisation == CB2_POL_NO ? mc.amp : 0.5f * mc.amp;
        } else if (kk == 0) {
            if (!pol_pi) return;
            a = mc.a_pi;
        } else {
            if (!pol_sigma) return;
            a = mc.a_sigma;
            const float dl = kk == 1 ? mc.dl_plus : mc.dl_minus;
            cf = cs.c0_frac + (dl + (M.wavelength + dl) * mc.dop) * M.inv_delta;
        }
        if (M.shape == CB2_SHAPE_STARK) {
            // pseudo-Voigt: (1 - eta) Gaussian + eta modified Lorentzian per Zeeman component (stark.pyx:305-346)
            if (k >= 3) { a *= mc.eta; width = mc.lam_b; }
            else a *= 1.0f - mc.eta;
        }
        amp = a;
        return;
    }
    }
}

static __device__ __noinline__ double lorentz_cdf_call(const double2* __restrict__ tab, double phi_inf, double u) {
    return lorentz_cdf(tab, phi_inf, u);
}

// what a lane knows about its sample's line component while the windows are evaluated
struct LineRec {
    float kx, xoff;          // series: x'(rel) = rel kx + xoff at the bin CENTRE; erfc: x at the bin's UPPER edge; Lorentzian: u at edge rel
    float s0, s1, s2, s3, s4;// series coefficients, or s0 = amplitude
    int lo, hi;              // evaluated bins [lo, hi), relative to the slot origin
    int kind;                // 0 nothing, 1 series, 2 erfc differences, 3 modified Lorentzian
};

// series coefficients of a lane's component, each duplicated into both halves of a pair
struct SeriesPair {
    f32x2 c0, c1, c2, c3, c4;
};
__device__ __forceinline__ SeriesPair series_pair(const LineRec& R, bool mine) {
    const float c0 = mine ? R.s0 : 0.f, c1 = mine ? R.s1 : 0.f, c2 = mine ? R.s2 : 0.f, c3 = mine ? R.s3 : 0.f, c4 = mine ? R.s4 : 0.f;
    SeriesPair P;
    P.c0 = pack2(c0, c0); P.c1 = pack2(c1, c1); P.c2 = pack2(c2, c2); P.c3 = pack2(c3, c3); P.c4 = pack2(c4, c4);
    return P;
}
// bin integrals of two bins at x' = (X.lo, X.hi): 2^(-x'^2) * poly4(x'^2)
__device__ __forceinline__ f32x2 series_eval2(const SeriesPair& P, f32x2 X) {
    const f32x2 m2 = mul2(X, X);
    float a, b;
    unpack2(m2, a, b);
    const f32x2 e = pack2(ex2_approx(-a), ex2_approx(-b));
    return mul2(e, fma2(fma2(fma2(fma2(P.c4, m2, P.c3), m2, P.c2), m2, P.c1), m2, P.c0));
}

// 32-bin window, series lanes, lane-permuted bin order (register r of lane l holds bin r ^ l) so that the butterfly needs
// no selects; lane l ends up with window bin l.  Registers are evaluated as the pairs (r, r + 1), r even.
template <typename AccT>
__device__ __forceinline__ void window_series_xor(const LineRec& R, int wbase, int c0_int, int bins, AccT* __restrict__ wacc, int lane) {
    const bool mine = R.kind == 1 && R.hi > wbase && R.lo < wbase + 32;
    const SeriesPair P = series_pair(R, mine);
    // per-lane steps of the XOR-ordered walk: r ^ l = l + sum_{b in r} (+-)2^b
    float db0;
    f32x2 db[5];
#pragma unroll
    for (int b = 0; b < 5; b++) {
        const float d = ((lane >> b) & 1) ? -(float)(1 << b) * R.kx : (float)(1 << b) * R.kx;
        if (b == 0) db0 = d;
        db[b] = pack2(d, d);
    }
    const float xb = fmaf((float)(wbase + lane), R.kx, R.xoff);
    f32x2 X[16], part[16];
    X[0] = pack2(xb, xb + db0);
#pragma unroll
    for (int j = 1; j < 16; j++) {
        // pair j holds registers (2j, 2j + 1): x = x of the pair without j's top bit + the step of that bit
        int top = 3;
        while (!(j & (1 << top))) top--;
        X[j] = add2(X[j & ~(1 << top)], db[top + 1]);
    }
#pragma unroll
    for (int j = 0; j < 16; j++) part[j] = series_eval2(P, X[j]);
    // butterfly part[i] += shfl_xor(part[i + o], o) on the register index; pairs move as two 32-bit shuffles, add as one FADD2
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j < o; j++) {
            float a, b;
            unpack2(part[j + o], a, b);
            part[j] = add2(part[j], pack2(__shfl_xor_sync(FULL, a, 2 * o), __shfl_xor_sync(FULL, b, 2 * o)));
        }
    float p0, p1;
    unpack2(part[0], p0, p1);
    p0 += __shfl_xor_sync(FULL, p1, 1);
    const int bin = c0_int + wbase + lane;
    __syncwarp();        // the warp's accumulator is updated by whichever lane a window assigns to a bin: order the windows' updates
    if (bin >= 0 && bin < bins && p0 != 0.f) wacc[bin] += (AccT)p0;
}

// transpose-reduce WB registers over the warp: inside each group of WB lanes a lane keeps, after the step with offset o,
// the half selected by (lane & o); the 32 / WB groups are then folded, so lane l < WB holds window bin l
template <int WB>
__device__ __forceinline__ float reduce_window(float (&part)[WB], int lane) {
#pragma unroll
    for (int o = WB / 2; o > 0; o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; i++) {
            const float send = upper ? part[i] : part[i + o];
            const float keep = upper ? part[i + o] : part[i];
            part[i] = keep + __shfl_xor_sync(FULL, send, o);
        }
    }
#pragma unroll
    for (int o = WB; o < 32; o <<= 1) part[0] += __shfl_xor_sync(FULL, part[0], o);
    return part[0];
}

// WB-bin tail window (WB = 8, 16) for series lanes, in the same lane-permuted order: register r of lane l holds bin r ^ (l % WB), the
// butterfly over the register index needs no selects, the 32 / WB lane groups are folded at the end; lane l < WB holds window bin l.
template <int WB, typename AccT>
__device__ __forceinline__ void window_series_tail(const LineRec& R, int wbase, int c0_int, int bins, AccT* __restrict__ wacc, int lane) {
    constexpr int BITS = WB == 16 ? 4 : 3;
    const bool mine = R.kind == 1 && R.hi > wbase && R.lo < wbase + WB;
    const SeriesPair P = series_pair(R, mine);
    const int lw = lane & (WB - 1);
    float db0;
    f32x2 db[BITS];
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        const float d = ((lw >> b) & 1) ? -(float)(1 << b) * R.kx : (float)(1 << b) * R.kx;
        if (b == 0) db0 = d;
        db[b] = pack2(d, d);
    }
    const float xb = fmaf((float)(wbase + lw), R.kx, R.xoff);
    f32x2 X[WB / 2], part[WB / 2];
    X[0] = pack2(xb, xb + db0);
#pragma unroll
    for (int j = 1; j < WB / 2; j++) {
        int top = BITS - 2;
        while (!(j & (1 << top))) top--;
        X[j] = add2(X[j & ~(1 << top)], db[top + 1]);
    }
#pragma unroll
    for (int j = 0; j < WB / 2; j++) part[j] = series_eval2(P, X[j]);
#pragma unroll
    for (int o = WB / 4; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j < o; j++) {
            float a, b;
            unpack2(part[j + o], a, b);
            part[j] = add2(part[j], pack2(__shfl_xor_sync(FULL, a, 2 * o), __shfl_xor_sync(FULL, b, 2 * o)));
        }
    float v, p1;
    unpack2(part[0], v, p1);
    v += __shfl_xor_sync(FULL, p1, 1);
#pragma unroll
    for (int o = WB; o < 32; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
    const int bin = c0_int + wbase + lane;
    __syncwarp();
    if (lane < WB && bin >= 0 && bin < bins && v != 0.f) wacc[bin] += (AccT)v;
}

// erfc-difference (sub-bin Gaussian) and modified-Lorentzian lanes: rolled loop over 4-bin steps with the reference's
// lower = upper recurrence along the bin edges (gaussian.pyx:78-88, stark.pyx:139-146).  Small code, any range length.
template <typename AccT, int LOR>
__device__ __forceinline__ void edges_pass(const LineRec& R, int c0_int, int bins, AccT* __restrict__ wacc, int lane,
                                           const double2* __restrict__ ltab, double phi_inf) {
    const bool is_e = R.kind == 2, is_l = LOR && R.kind == 3;
    const int lo = (is_e || is_l) ? R.lo : INT_MAX, hi = (is_e || is_l) ? R.hi : INT_MIN;
    const int Elo = __reduce_min_sync(FULL, lo), Ehi = __reduce_max_sync(FULL, hi);
    if (Ehi <= Elo) return;
    const float amp = (is_e || is_l) ? R.s0 : 0.f;
    // value of the cumulative profile at the lower edge of bin Elo
    float xl = fmaf((float)(Elo - 1), R.kx, R.xoff), tl = is_e ? half_erfc(fabsf(xl)) : 0.f;
    double pl = 0.0;
    const double inv_l = (double)R.kx, off = (double)R.xoff;
    bool have_pl = false;
    for (int wbase = Elo; wbase < Ehi; wbase += 4) {
        float part[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const int rel = wbase + w;
            float v = 0.f;
            if (is_e) {
                const float xu = fmaf((float)rel, R.kx, R.xoff);                  // x at the upper edge of bin rel
                const float tu = half_erfc(fabsf(xu));
                const float dd = (xl >= 0.f) ? (tl - tu) : ((xu <= 0.f) ? (tu - tl) : (1.0f - tl - tu));
                if (rel >= lo && rel < hi) v = amp * dd;
                xl = xu; tl = tu;
            }
            if (LOR && is_l && rel >= lo && rel < hi) {
                if (!have_pl) { pl = lorentz_cdf_call(ltab, phi_inf, (double)rel * inv_l + off); have_pl = true; }
                const double pu = lorentz_cdf_call(ltab, phi_inf, (double)(rel + 1) * inv_l + off);
                v = amp * (float)(pu - pl);
                pl = pu;
            }
            part[w] = v;
        }
        const float v = reduce_window<4>(part, lane);
        const int bin = c0_int + wbase + lane;
        __syncwarp();
        if (lane < 4 && bin >= 0 && bin < bins && v != 0.f) wacc[bin] += (AccT)v;
    }
}

// One line component of one 32-sample group (lanes = samples): bin integrals into the warp's private accumulator.
template <typename AccT, int LOR>
__device__ __forceinline__ void component_pass(int type, float cf, float width, float amp, int c0_int, int bins, AccT* __restrict__ wacc,
                                               int lane, const double2* __restrict__ ltab, double phi_inf,
                                               unsigned& gauss_evals, unsigned& lorentz_evals) {
    LineRec R;
    R.kx = R.xoff = R.s0 = R.s1 = R.s2 = R.s3 = R.s4 = 0.f;
    R.lo = INT_MAX; R.hi = INT_MIN; R.kind = 0;
    if (amp > 0.f && width > 0.f) {
        const float win_lo = (float)(-c0_int), win_hi = (float)(bins - c0_int);
        // the reference's range (GAUSSIAN_CUTOFF_SIGMA = 10, LORENTZIAN_CUTOFF_GAMMA = 50) defines the work counters
        const float cut = (type == 1 ? 50.0f : 10.0f) * width;
        const float flo = floorf(cf - cut), fhi = ceilf(cf + cut);
        if (fhi > win_lo && flo < win_hi) {
            const int l = (int)fmaxf(flo, win_lo), h = (int)fminf(fhi, win_hi);
            if (h > l) {
                if (type == 1) {
                    lorentz_evals += (unsigned)(h - l) + 1u;
                    if (LOR) { R.kind = 3; R.lo = l; R.hi = h; R.kx = 1.0f / width; R.xoff = -cf * R.kx; R.s0 = amp; }  // u at edge e: e kx + xoff
                } else {
                    gauss_evals += (unsigned)(h - l) + 1u;
                    // Evaluated range: 7 sigma when the component's own core (centre +- 2.5 sigma) overlaps the spectral window —
                    // the ray's largest bin is then >= e^-3.1 of this sample's peak and what lies beyond 7 sigma (e^-24.5) is
                    // under the 1e-9 acceptance floor.  A window that only sees the far wing has no such floor: the reference's
                    // full 10 sigma range is evaluated, with the series only where it still converges there (m h <= 0.25), else
                    // with erfc differences (accurate in relative terms in the wing).
                    const bool core_in = (cf + 2.5f * width > win_lo) && (cf - 2.5f * width < win_hi);
                    const float ecut = (core_in ? EVAL_CUTOFF_SIGMA : 10.0f) * width;
                    R.lo = max(l, (int)fmaxf(floorf(cf - ecut), win_lo));
                    R.hi = min(h, (int)fminf(ceilf(cf + ecut), win_hi));
                    const float kb = 0.70710678f * rcp_approx(width);         // delta / (sqrt(2) sigma), per bin
                    const float hh = 0.5f * kb;
                    if (hh <= (core_in ? H_SERIES_MAX : 0.035f)) {
                        const float h2 = hh * hh;
                        const float t1 = h2 * (1.0f / 6.0f), t2 = h2 * h2 * (1.0f / 120.0f), t3 = h2 * h2 * h2 * (1.0f / 5040.0f),
                                    t4 = h2 * h2 * h2 * h2 * (1.0f / 362880.0f);
                        const float A = amp * kb * INV_SQRT_PI;
                        R.kind = 1;
                        R.kx = kb * SQRT_L2E;
                        R.xoff = (0.5f - cf) * R.kx;                          // x' at the centre of relative bin 0
                        R.s0 = A * (1.0f - 2.0f * t1 + 12.0f * t2 - 120.0f * t3 + 1680.0f * t4);
                        R.s1 = A * (4.0f * t1 - 48.0f * t2 + 720.0f * t3 - 13440.0f * t4) * INV_L2E;
                        R.s2 = A * (16.0f * t2 - 480.0f * t3 + 13440.0f * t4) * (INV_L2E * INV_L2E);
                        R.s3 = A * (64.0f * t3 - 3584.0f * t4) * (INV_L2E * INV_L2E * INV_L2E);
                        R.s4 = A * (256.0f * t4) * (INV_L2E * INV_L2E * INV_L2E * INV_L2E);
                        if (!(R.s0 > 0.f)) R.kind = 0;
                    } else {
                        R.kind = 2;
                        R.kx = kb;
                        R.xoff = (1.0f - cf) * kb;                            // x at the UPPER edge of relative bin 0
                        R.s0 = amp;
                    }
                    if (R.hi <= R.lo || R.kind == 0) { R.kind = 0; R.lo = INT_MAX; R.hi = INT_MIN; }
                }
            }
        }
    }
    const unsigned any_series = __ballot_sync(FULL, R.kind == 1), any_edges = __ballot_sync(FULL, R.kind >= 2);
    if (any_series) {
        // windows start at the union's first bin; the last (or only) one uses the narrowest width class that covers it
        const int Rlo = __reduce_min_sync(FULL, R.kind == 1 ? R.lo : INT_MAX), Rhi = __reduce_max_sync(FULL, R.kind == 1 ? R.hi : INT_MIN);
        for (int wbase = Rlo; wbase < Rhi; wbase += 32) {
            const int left = Rhi - wbase;
            if (left <= 8) { window_series_tail<8, AccT>(R, wbase, c0_int, bins, wacc, lane); break; }
            if (left <= 16) { window_series_tail<16, AccT>(R, wbase, c0_int, bins, wacc, lane); break; }
            window_series_xor<AccT>(R, wbase, c0_int, bins, wacc, lane);
        }
    }
    if (any_edges) edges_pass<AccT, LOR>(R, c0_int, bins, wacc, lane, ltab, phi_inf);
}

// ------------------------------------------------------------------------------------------------------------------
// ray / segment geometry shared by the kernels: NumericalIntegrator.integrate [raysect] takes start_point = far end of the
// segment, end_point = near end, both to plasma space; float64 with the reference's operation order so that step positions
// are bit-identical
// ------------------------------------------------------------------------------------------------------------------
struct SegGeom {
    double sx, sy, sz, ivx, ivy, ivz, h;
    int iv;                                           // intervals; samples k = 0..iv; 0 for a degenerate segment (skipped)
};

__device__ __forceinline__ SegGeom segment_geometry(const double* __restrict__ w2p, double step, int min_samples, double ox, double oy,
                                                    double oz, double dwx, double dwy, double dwz, double t0, double t1) {
    SegGeom g;
    const double swx = __dadd_rn(ox, __dmul_rn(t1, dwx)), swy = __dadd_rn(oy, __dmul_rn(t1, dwy)), swz = __dadd_rn(oz, __dmul_rn(t1, dwz));
    const double ewx = __dadd_rn(ox, __dmul_rn(t0, dwx)), ewy = __dadd_rn(oy, __dmul_rn(t0, dwy)), ewz = __dadd_rn(oz, __dmul_rn(t0, dwz));
    g.sx = xform_row(w2p, swx, swy, swz, true); g.sy = xform_row(w2p + 4, swx, swy, swz, true); g.sz = xform_row(w2p + 8, swx, swy, swz, true);
    g.ivx = __dsub_rn(xform_row(w2p, ewx, ewy, ewz, true), g.sx);
    g.ivy = __dsub_rn(xform_row(w2p + 4, ewx, ewy, ewz, true), g.sy);
    g.ivz = __dsub_rn(xform_row(w2p + 8, ewx, ewy, ewz, true), g.sz);
    const double length = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(g.ivx, g.ivx), __dmul_rn(g.ivy, g.ivy)), __dmul_rn(g.ivz, g.ivz)));
    g.iv = 0; g.h = 0.0;
    if (!(length > 0.0)) return g;
    g.ivx = __ddiv_rn(g.ivx, length); g.ivy = __ddiv_rn(g.ivy, length); g.ivz = __ddiv_rn(g.ivz, length);
    int iv = (int)ceil(__ddiv_rn(length, step));      // intervals = max(min_samples - 1, ceil(L / step))
    iv = max(iv, max(min_samples - 1, 1));
    g.iv = iv;
    g.h = __ddiv_rn(length, (double)iv);
    return g;
}

// groups of 32 samples per ray (each segment starts a new group) -> counts[ray]; counts[n_rays] = 0 for the scan
__global__ void count_groups_kernel(const __grid_constant__ DevScene S, DevRays rays, int64_t* __restrict__ counts) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray > rays.n_rays) return;
    if (ray == rays.n_rays) { counts[ray] = 0; return; }
    const double ox = rays.origin[3 * ray], oy = rays.origin[3 * ray + 1], oz = rays.origin[3 * ray + 2];
    const double dwx = rays.direction[3 * ray], dwy = rays.direction[3 * ray + 1], dwz = rays.direction[3 * ray + 2];
    int64_t n = 0;
    for (int64_t sg = rays.seg_offset[ray]; sg < rays.seg_offset[ray + 1]; sg++) {
        const SegGeom g = segment_geometry(S.w2p, S.step, S.min_samples, ox, oy, oz, dwx, dwy, dwz, rays.seg_t0[sg], rays.seg_t1[sg]);
        if (g.iv > 0) n += g.iv / 32 + 1;
    }
    counts[ray] = n;
}

// the scan's total goes to a host-mapped word: a device->host memcpy of 8 bytes would queue on the copy engine behind a
// frame read-back that may be in flight on another stream and stall the next batch
__global__ void publish_total_kernel(const int64_t* __restrict__ src, volatile int64_t* __restrict__ dst) { *dst = *src; }

// record layout: rec[((group * n_comp + comp) * 3 + field) * 32 + lane], field 0 centre, 1 width, 2 amplitude
#define REC_FLOATS_PER_COMP 96

// ------------------------------------------------------------------------------------------------------------------
// K1a: per-sample plasma state -> line records + Bremsstrahlung moments
// ------------------------------------------------------------------------------------------------------------------
#ifndef CB2_FIX_MINB
// the fix-up instance (one warp per ray over the flagged samples): 5 CTAs per SM (102 registers, no spills) — 2.70 ms per 65 536 C3
// rays against 2.79 at 6, 2.90 at 4, 3.26 at 8 (64 registers, spills)
#define CB2_FIX_MINB 5
#endif
#ifndef CB2_STATE_MINB
// 6 CTAs of 4 warps per SM (85 registers, ~200 B of spills): measured 3 % faster than 5 (102 registers), 4 is 6 % slower, 8 slower again
#define CB2_STATE_MINB 6
#endif
// FEAT = 0: plasma line models and Bremsstrahlung only (the benchmark's scene); FEAT = 1 adds the beam frame / BeamCXLine,
// ThermalCXLine and TotalRadiatedPower branches (kept out of the common instance: they cost registers and instruction cache)
#define CB2_FIX_WINDOW 32                              // groups per list round of the fix-up pass (at most 1024 flagged samples)
template <int NW, int MOM, int AXONLY, int FEAT, int FIX>
__global__ void __launch_bounds__(NW * 32, FIX ? CB2_FIX_MINB : CB2_STATE_MINB)
state_kernel(const __grid_constant__ DevScene Sparam, DevRays rays, const int64_t* __restrict__ gbase, unsigned* __restrict__ gmask,
             float* __restrict__ rec, unsigned long long* __restrict__ stats, float* __restrict__ mom_out, double* __restrict__ flat_out,
             int count_samples, int dbg_skip, const unsigned* __restrict__ gblend, const __grid_constant__ DevMemo FM) {
    // FIX: fix-up pass behind state_fast_kernel — only the samples flagged there (blend zone, gblend) are evaluated, ONE WARP PER
    // RAY: the warp gathers the flagged samples of CB2_FIX_WINDOW groups into an ordered list and takes 32 of them at a time
    // (lanes = flagged samples of any group: no lane idles beside a partly flagged group, no block barrier, every warp of the
    // SM busy); their record entries are overwritten one by one and their moments added to the ray's row in HBM (float atomics
    // issued by one warp in program order: reproducible)
    extern __shared__ double smem_d[];
    __shared__ double flat_s;
    __shared__ int fix_list_s[FIX ? NW * CB2_FIX_WINDOW * 32 : 1];
    // the flattened scene (13 KB of model / species / table descriptors, indexed with run-time model and species numbers)
    // is staged in shared memory: indexed constant-bank loads miss the small constant cache and stall (ncu: short scoreboard)
    __shared__ __align__(16) unsigned char scene_s[sizeof(DevScene)];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint4* src = reinterpret_cast<const uint4*>(&Sparam);
        uint4* dst = reinterpret_cast<uint4*>(scene_s);
        for (int i = tid; i < (int)(sizeof(DevScene) / sizeof(uint4)); i += NW * 32) dst[i] = src[i];
    }
    const DevScene& S = *reinterpret_cast<const DevScene*>(scene_s);
    const int k_pad = MOM ? Sparam.brems.k_pad : 0;
    double* mom = smem_d;
    if (MOM && !FIX)
        for (int i = tid; i < k_pad; i += NW * 32) mom[i] = 0.0;
    if (tid == 0) flat_s = 0.0;
    float flat_acc = 0.f;
    __syncthreads();

    const int64_t ray = FIX ? (int64_t)blockIdx.x * NW + warp : blockIdx.x;
    if (FIX && ray >= rays.n_rays) return;                // (no block barrier follows in the fix-up mode)
    int* const fix_list = fix_list_s + (FIX ? warp * CB2_FIX_WINDOW * 32 : 0);
    float* const mom_row = MOM ? mom_out + (size_t)ray * k_pad : nullptr;
    const double ox = rays.origin[3 * ray], oy = rays.origin[3 * ray + 1], oz = rays.origin[3 * ray + 2];
    const double dwx = rays.direction[3 * ray], dwy = rays.direction[3 * ray + 1], dwz = rays.direction[3 * ray + 2];
    SampleIn in;
    in.donor = 0.f; in.bvx = in.bvy = in.bvz = 0.f;
    const bool has_beam = FEAT && S.beam.present != 0;
    {
        // ray direction in plasma space (direction.transform(local_to_plasma), normalised inside doppler_shift); in a beam
        // scene w2p is world -> beam and the observation direction goes on to plasma space (beam/material.pyx:62-65)
        double d0 = xform_row(S.w2p, dwx, dwy, dwz, false), d1 = xform_row(S.w2p + 4, dwx, dwy, dwz, false),
               d2 = xform_row(S.w2p + 8, dwx, dwy, dwz, false);
        if (has_beam) {
            const double e0 = xform_row(S.beam.l2p, d0, d1, d2, false), e1 = xform_row(S.beam.l2p + 4, d0, d1, d2, false),
                         e2 = xform_row(S.beam.l2p + 8, d0, d1, d2, false);
            d0 = e0; d1 = e1; d2 = e2;
        }
        const double dl = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        in.dx = (float)(d0 / dl); in.dy = (float)(d1 / dl); in.dz = (float)(d2 / dl);
    }
    unsigned long long n_samples = 0;
    unsigned n_brems = 0, ood = 0;
    const int n_comp = S.n_comp;

    int64_t G0 = gbase[ray];                          // first group of the current segment
    int g_rot = 0;                                    // keeps the round-robin going across segments
    for (int64_t sg = rays.seg_offset[ray]; sg < rays.seg_offset[ray + 1]; sg++) {
        const SegGeom sgm = segment_geometry(S.w2p, S.step, S.min_samples, ox, oy, oz, dwx, dwy, dwz, rays.seg_t0[sg], rays.seg_t1[sg]);
        const int iv = sgm.iv;
        if (iv <= 0) continue;
        const float hf = (float)sgm.h;
        if (tid == 0 && !FIX) n_samples += (unsigned long long)iv + 1ull;
        const int n_groups = iv / 32 + 1;
        int first = (warp - g_rot) % NW;
        if (first < 0) first += NW;
        g_rot = (g_rot + n_groups) % NW;

        const int n_win = FIX ? (n_groups + CB2_FIX_WINDOW - 1) / CB2_FIX_WINDOW : 1;
        for (int win = 0; win < n_win; win++) {
        int it0 = first, it_end = n_groups, n_list = 0;
        if (FIX) {
            // ordered list of the window's flagged samples: lane l owns group win * 32 + l
            __syncwarp();                                          // the previous round's list has been consumed
            unsigned bm = 0;
            const int gl = win * CB2_FIX_WINDOW + lane;
            if (gl < n_groups) bm = __ldg(gblend + G0 + gl);
            const int cnt = __popc(bm);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            n_list = __shfl_sync(FULL, incl, 31);
            if (!n_list) continue;
            int at = incl - cnt;
            while (bm) {
                const int b = __ffs(bm) - 1;
                bm &= bm - 1;
                fix_list[at++] = gl * 32 + b;
            }
            __syncwarp();
            it0 = 0; it_end = (n_list + 31) >> 5;
        }
        for (int it = it0; it < it_end; it += FIX ? 1 : NW) {
            int k = it * 32 + lane;
            bool sel = true;
            if (FIX) {
                sel = k < n_list;
                k = sel ? fix_list[k] : 0;
            }
            const int g = FIX ? (k >> 5) : it;                     // FIX: per lane
            const int rl = FIX ? (k & 31) : lane;                  // the sample's slot in its group's record rows
            const bool active = k <= iv && sel;
            const double tk = __dmul_rn((double)k, sgm.h);
            double pxd = __dadd_rn(sgm.sx, __dmul_rn(tk, sgm.ivx)), pyd = __dadd_rn(sgm.sy, __dmul_rn(tk, sgm.ivy)),
                   pzd = __dadd_rn(sgm.sz, __dmul_rn(tk, sgm.ivz));
            in.weight = active ? ((k == 0 || k == iv) ? 0.5f * hf : hf) : 0.f;
            if (has_beam) {
                // the sample is in beam coordinates: donor density and velocity there, then on to plasma space
                const float xb = (float)pxd, yb = (float)pyd, zb = (float)pzd;
                in.donor = active ? beam_density(S.beam, xb, yb, zb) : 0.f;
                const float3 bd = beam_direction(S.beam, xb, yb, zb);
                const double* L = S.beam.l2p;
                float vx = (float)L[0] * bd.x + (float)L[1] * bd.y + (float)L[2] * bd.z, vy = (float)L[4] * bd.x + (float)L[5] * bd.y + (float)L[6] * bd.z,
                      vz = (float)L[8] * bd.x + (float)L[9] * bd.y + (float)L[10] * bd.z;
                const float sc_ = S.beam.speed * rsqrtf(vx * vx + vy * vy + vz * vz);
                in.bvx = vx * sc_; in.bvy = vy * sc_; in.bvz = vz * sc_;
                const double qx = xform_row(L, pxd, pyd, pzd, true), qy = xform_row(L + 4, pxd, pyd, pzd, true), qz = xform_row(L + 8, pxd, pyd, pzd, true);
                pxd = qx; pyd = qy; pzd = qz;
            }
            in.x = (float)pxd; in.y = (float)pyd; in.z = (float)pzd;
            AxCtx ctx;
            float ne = 0.f, te = 0.f;
            if (active && (!has_beam || in.donor > 0.f)) {        // beam scenes: nothing is evaluated where the beam density is zero
                ax_setup(S, pxd, pyd, pzd, ctx, ood);
                ne = eval_scalar_t<AXONLY>(S.ne, ctx, in.x, in.y, in.z);
                te = eval_scalar_t<AXONLY>(S.te, ctx, in.x, in.y, in.z);
            } else {
                ctx.m = 0.f; ctx.tri = -1; ctx.in_lcfs = false; ctx.b_outside = false;
            }
            const bool live = has_beam ? (in.weight > 0.f && in.donor > 0.f) : (ne > 0.f && te > 0.f && in.weight > 0.f);
            const unsigned live_mask = __ballot_sync(FULL, live);
            const int64_t G = G0 + g;
            if (lane == 0 && !FIX) gmask[G] = live_mask;
            if (!live_mask) continue;                              // the whole group is in vacuum
            if (MOM && dbg_skip != 2) {
                if (FIX) sample_brems_moments<unsigned, AXONLY>(S, in, ctx, ne, te, mom_row, lane, n_brems, ood);
                else sample_brems_moments<unsigned, AXONLY>(S, in, ctx, ne, te, mom, lane, n_brems, ood);
            }
            LineCache lc;
            lc.cur = -1; lc.ni = lc.ts = lc.vd = 0.f; lc.have_b = false; lc.bm = lc.cos_sqr = 0.f; lc.grid = -2;
            lc.lne = lc.lte = 0.f;
            if (live && ne > 0.f && te > 0.f) {
                lc.lne = log10f(ne) + 19.0f;                       // densities are stored in units of 1e19 m^-3
                lc.lte = log10f(te);
            }
            float* grec = rec + (size_t)G * n_comp * REC_FLOATS_PER_COMP + rl;
            if (FIX) {
                // a scene with state tables holds Excitation / Recombination lines with Gaussian or multiplet shapes only (cb2_memo_build):
                // the model loop without its generality — per species slot (ni, sqrt(Ts), v.d) once, then per line the PEC and the rows
                float nis[CB2_MEMO_MAX_SP], sqs[CB2_MEMO_MAX_SP], vds[CB2_MEMO_MAX_SP];
#pragma unroll
                for (int sl = 0; sl < CB2_MEMO_MAX_SP; sl++) {
                    nis[sl] = sqs[sl] = vds[sl] = 0.f;
                    if (sl < FM.n_sp && live) {                       // (the context of a lane that is not live is not set up)
                        const DevSpecies& sp = S.species[FM.sp_species[sl]];
                        nis[sl] = eval_scalar_t<AXONLY>(sp.density, ctx, in.x, in.y, in.z);
                        const float ts = eval_scalar_t<AXONLY>(sp.temperature, ctx, in.x, in.y, in.z);
                        sqs[sl] = ts > 0.f ? sqrtf(ts) : 0.f;
                        const float3 v = eval_vector(sp.velocity, ctx);
                        vds[sl] = v.x * in.dx + v.y * in.dy + v.z * in.dz;
                    }
                }
                for (int l = 0; l < FM.n_lines; l++) {
                    const MemoLine& L = FM.lines[l];
                    const DevModel& M = S.models[L.model];
                    float ni = nis[0], sq = sqs[0], vd = vds[0];
#pragma unroll
                    for (int sl = 1; sl < CB2_MEMO_MAX_SP; sl++)
                        if (L.slot == sl) { ni = nis[sl]; sq = sqs[sl]; vd = vds[sl]; }
                    float amp = 0.f;
                    if (live && ni > 0.f) {
                        float lp;
                        if (M.pec_const) lp = M.pec_value;
                        else {
                            if (M.pec_grid != lc.grid) { lc.cell = locate2d(M.pec, lc.lne, lc.lte); lc.grid = M.pec_grid; }
                            if (!lc.cell.inside && !M.pec_extrapolate) ood++;
                            lp = eval2d(M.pec, lc.cell);
                        }
                        amp = RECIP_4_PI * exp10f(lp) * ne * ni * in.weight * M.inv_delta;     // impact_excitation.pyx:99
                        if (!(amp > 0.f) || !(sq > 0.f)) amp = 0.f;                          // gaussian.pyx:127-129
                    }
                    const float width = L.sigma_coef * sq;
                    float* r = grec + L.rec_off;
                    if (L.shape == CB2_SHAPE_GAUSSIAN) {
                        if (sel) { r[64] = amp; r[0] = fmaf(L.shift_coef, vd, L.c0_frac); r[32] = width; }
                    } else {
                        const float dop = vd * L.inv_c;
                        for (int kc = 0; kc < L.ncomp; kc++, r += REC_FLOATS_PER_COMP)
                            if (sel) { r[64] = amp * __ldg(L.mult_ratio + kc); r[0] = S.comps[L.comp0 + kc].c0_frac + __ldg(L.mult_lambda + kc) * dop * L.inv_delta; r[32] = width; }
                    }
                }
                continue;
            }
            for (int m = 0; m < S.n_models; m++) {                 // PlasmaMaterial.emission_function loop, material.pyx:59-61
                const DevModel& M = S.models[m];
                if (M.kind == CB2_MODEL_BREMSSTRAHLUNG) continue;
                if (FEAT && M.kind == CB2_MODEL_BEAM_EMISSION_LINE) {
                    float amp, shift_b, split_b;
                    beam_emission_setup<AXONLY>(S, M, in, ctx, ne, te, lc, amp, shift_b, split_b, ood);
                    const bool any_amp = __any_sync(FULL, amp > 0.f);
                    const DevModelExt& X = *M.ext;
                    float mamp[9];
#pragma unroll
                    for (int kc = 0; kc < 9; kc++) mamp[kc] = X.mse_amp[kc];
                    if (X.mse_n > 1 && amp > 0.f) {
                        // intensity ratios as functions of ne (mse.pyx:103-121): linear in log10(ne) on the host's knots, clamped
                        float f = fminf(fmaxf((lc.lne - X.mse_lne0) * X.mse_inv_dlne, 0.f), (float)(X.mse_n - 1));
                        const int i0 = min((int)f, X.mse_n - 2);
                        const float w = f - (float)i0;
                        const float4 a = __ldg(X.mse_tab + i0), b = __ldg(X.mse_tab + i0 + 1);
                        const float s2p = fmaf(w, b.x - a.x, a.x), s1s0 = fmaf(w, b.y - a.y, a.y), p23 = fmaf(w, b.z - a.z, a.z), p43 = fmaf(w, b.w - a.w, a.w);
                        const float dd = 1.0f / (1.0f + s2p), isig = s2p * dd, ipi = 0.5f * dd, is0 = 1.0f / (s1s0 + 1.0f), is1 = 0.5f * s1s0 * is0;
                        const float ip3 = 1.0f / (1.0f + p23 + p43), ip2 = p23 * ip3, ip4 = p43 * ip3;
                        mamp[0] = isig * is0; mamp[1] = mamp[2] = isig * is1;
                        mamp[3] = mamp[4] = ipi * ip2; mamp[5] = mamp[6] = ipi * ip3; mamp[7] = mamp[8] = ipi * ip4;
                    }
#pragma unroll
                    for (int kc = 0; kc < 9; kc++) {
                        float* r = grec + (size_t)(M.comp0 + kc) * REC_FLOATS_PER_COMP;
                        // component order: sigma0, sigma1 +-, pi2 +-, pi3 +-, pi4 +-  (mse.pyx:113-133)
                        const float off = kc == 0 ? 0.f : (float)((kc + 1) >> 1) * ((kc & 1) ? 1.f : -1.f);
                        if (!FIX || sel) r[64] = amp * mamp[kc];
                        if (FIX ? sel : any_amp) { r[0] = S.comps[M.comp0 + kc].c0_frac + shift_b + off * split_b; r[32] = X.mse_sigma_b; }
                    }
                    continue;
                }
                if (FEAT && M.kind == CB2_MODEL_TOTAL_RADIATED_POWER) {
                    if (live) flat_acc += total_radiated_power<AXONLY>(S, M, in, ctx, ne, lc.lne, lc.lte, ood);
                    continue;
                }
                ModelCtx mc;
                model_setup<AXONLY, FEAT>(S, M, in, ctx, ne, te, live, lc, mc, ood);
                const bool any_on = __any_sync(FULL, mc.on);
                for (int kc = 0; kc < M.ncomp; kc++) {
                    float* r = grec + (size_t)(M.comp0 + kc) * REC_FLOATS_PER_COMP;
                    int type; float cf = 0.f, width = 0.f, amp = 0.f;
                    if (any_on) model_component(S, M, kc, mc, type, cf, width, amp);
                    if (!FIX || sel) r[64] = amp;
                    if (FIX ? sel : __any_sync(FULL, amp > 0.f)) { r[0] = cf; r[32] = width; }
                }
            }
        }
        }
        G0 += n_groups;
    }
    if (FEAT && flat_out) {
        // per-ray wavelength-independent radiance: fp32 per thread, fp64 across the CTA
        for (int off = 16; off > 0; off >>= 1) flat_acc += __shfl_down_sync(FULL, flat_acc, off);
        if (lane == 0 && flat_acc != 0.f) atomicAdd(&flat_s, (double)flat_acc);
    }
    if (!FIX && (MOM || (FEAT && flat_out))) __syncthreads();
    if (MOM && !FIX)
        for (int i = tid; i < k_pad; i += NW * 32) mom_row[i] = (float)mom[i];
    if (FEAT && flat_out && tid == 0) flat_out[ray] = flat_s;
    if (stats) {
        unsigned long long nb = n_brems, oodl = ood;
        for (int off = 16; off > 0; off >>= 1) {
            nb += __shfl_down_sync(FULL, nb, off);
            oodl += __shfl_down_sync(FULL, oodl, off);
        }
        if (lane == 0) {
            if (nb) atomicAdd(stats + 3, nb);
            if (oodl) atomicAdd(stats + 5, oodl);
        }
        if (tid == 0 && n_samples && count_samples) atomicAdd(stats + 0, n_samples);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K1a, table-driven (DevMemo): the state of a core sample (blend weight 1) is a row interpolated on the psi_n grid, that
// of an edge sample (weight 0) the row of its triangle; what stays per sample is the geometry — position, (R, Z), psi_n,
// triangle, direction of the poloidal field, v.d — the record rows and the moment scatter.  Blend-zone samples are only
// flagged (gblend) and evaluated by state_kernel in its fix-up mode.  (Tried: field rows — ne, te, n_i, T_s, velocities, N_z per
// triangle and on a psi_n grid over the ramp — blended per sample inside this kernel, rates on the blended state.  Correct, 0.004 of
// the tolerance from the fix-up path, but the divergent block costs more than the pass it replaces: state 16.7 -> 18.8 ms per
// 65 536 rays inlined, 19.9 ms as a call.  Dropped.)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 lerp4(const float4 a, const float4 b, float t) {
    return make_float4(fmaf(t, b.x - a.x, a.x), fmaf(t, b.y - a.y, a.y), fmaf(t, b.z - a.z, a.z), fmaf(t, b.w - a.w, a.w));
}

// one table row from the generic evaluation code: which = 0 triangle `i`, 1 psi_n = (i + frac) / core_scale
__device__ __forceinline__ void memo_row(const DevScene& S, const DevMemo& FM, int which, int i, float frac, float* row) {
    AxCtx c;
    c.R = c.Z = 0.f; c.cphi = 1.f; c.sphi = 0.f; c.br = c.bt = c.bz = 0.f; c.b_outside = false;
    if (which == 0) {
        c.m = 0.f; c.tri = i; c.tri_c = i; c.we = 1.f; c.wc = 0.f; c.ci = 0; c.ct = 0.f; c.psi = 2.f; c.in_lcfs = false;
    } else {
        const float psi = fminf(((float)i + frac) / FM.core_scale, FM.psi_max);
        c.m = 1.f; c.tri = -1; c.tri_c = 0; c.we = 0.f; c.wc = 1.f; c.psi = psi; c.in_lcfs = true;
        locate1d(S.ax.core, psi, c.ci, c.ct);
    }
    SampleIn in;
    in.x = in.y = in.z = 0.f; in.dx = in.dy = 0.f; in.dz = 1.f; in.weight = 1.f; in.donor = 0.f; in.bvx = in.bvy = in.bvz = 0.f;
    for (int k = 0; k < 4 * FM.row_f4; k++) row[k] = 0.f;
    const float ne = eval_blend(S.ne, c), te = eval_blend(S.te, c);
    const bool live = ne > 0.f && te > 0.f;
    unsigned ood = 0;
    for (int s = 0; s < FM.n_sp; s++) {
        const DevSpecies& sp = S.species[FM.sp_species[s]];
        const float ts = eval_blend(sp.temperature, c);
        row[4 * s] = ts > 0.f ? sqrtf(ts) : 0.f;
        const DevVector& f = sp.velocity;
        if (FM.sp_const[s]) continue;
        if (which == 0) { row[4 * s + 1] = f.c[1]; row[4 * s + 2] = f.c[0]; row[4 * s + 3] = f.c[2]; }
        else {
            row[4 * s + 1] = f.vtor ? horner4(__ldg(f.vtor + c.ci), c.ct) : 0.f;
            row[4 * s + 2] = f.vpol ? horner4(__ldg(f.vpol + c.ci), c.ct) : 0.f;
            row[4 * s + 3] = f.vnorm ? horner4(__ldg(f.vnorm + c.ci), c.ct) : 0.f;
        }
    }
    LineCache lc;
    lc.lne = lc.lte = 0.f;
    if (live) { lc.lne = log10f(ne) + 19.0f; lc.lte = log10f(te); }
    for (int l = 0; l < FM.n_lines; l++) {
        lc.cur = -1; lc.ni = lc.ts = lc.vd = 0.f; lc.have_b = false; lc.bm = lc.cos_sqr = 0.f; lc.grid = -2;
        ModelCtx mc;
        model_setup<1, 0>(S, S.models[FM.lines[l].model], in, c, ne, te, live, lc, mc, ood);
        row[4 * FM.off_amp + l] = mc.on ? mc.amp : 0.f;
    }
    float f = 0.f, U[CB2_MAX_BREMS_Z];
#pragma unroll
    for (int z = 0; z < CB2_MAX_BREMS_Z; z++) U[z] = 0.f;
    if (FM.has_brems && live) brems_state<1>(S, in, c, ne, te, f, U, ood);
    float* b = row + 4 * FM.off_brems;
    b[0] = f; b[1] = (float)ood;
    for (int z = 0; z < FM.n_z; z++) b[2 + z] = U[z];
}

// mode 0: write rows [0, n) of table `which`; mode 2: column maxima of |entry| (float bits, atomicMax); mode 1 (core only):
// largest mid-interval deviation of the interpolated row from the generic evaluation, relative to the larger knot value — but
// not less than 1e-5 of the column's maximum: an entry that small against its own values elsewhere cannot be seen in a line
// integral — (node coordinate: 0.05 x absolute), as (error bits << 32 | entry << 24 | row) via atomicMax
__global__ void memo_rows_kernel(const __grid_constant__ DevScene S, const __grid_constant__ DevMemo FM, int which, int n, int mode,
                                 float4* __restrict__ tab, unsigned* __restrict__ colmax, unsigned long long* __restrict__ err_key) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* t = reinterpret_cast<float*>(tab) + (size_t)i * 4 * FM.row_f4;
    if (mode == 2) {
        for (int k = 0; k < 4 * FM.row_f4; k++) atomicMax(colmax + k, __float_as_uint(fabsf(t[k])));
        return;
    }
    float row[64];
    memo_row(S, FM, which, i, mode ? 0.5f : 0.f, row);
    if (!mode) {
        for (int k = 0; k < 4 * FM.row_f4; k++) t[k] = row[k];
        return;
    }
    const float* u = t + 4 * FM.row_f4;
    float worst = 0.f;
    int kw = 0;
    for (int k = 0; k < 4 * FM.row_f4; k++) {
        if (k == 4 * FM.off_brems + 1) continue;                       // out-of-domain count: taken from the lower knot
        const float a = t[k], b = u[k], v = 0.5f * (a + b), d = fabsf(v - row[k]);
        float e;
        if (k == 4 * FM.off_brems) e = (a >= 1.f && b >= 1.f) ? 0.05f * d : 0.f;   // d ln(spectrum) / d f = ds d ln(Phi) / ds < 0.05
        else e = d / (fmaxf(fmaxf(fmaxf(fabsf(a), fabsf(b)), fabsf(row[k])), 1e-5f * __uint_as_float(colmax[k])) + 1e-30f);
        if (e > worst) { worst = e; kw = k; }
    }
    atomicMax(err_key, ((unsigned long long)__float_as_uint(worst) << 32) | ((unsigned long long)(kw & 255) << 24) | (unsigned)(i & 0xffffff));
}

template <int NW, int MOM, int MINB, int NSP>
__global__ void __launch_bounds__(NW * 32, MINB)
state_fast_kernel(const __grid_constant__ DevScene S, const __grid_constant__ DevMemo FM, DevRays rays, const int64_t* __restrict__ gbase,
                  unsigned* __restrict__ gmask, unsigned* __restrict__ gblend, float* __restrict__ rec,
                  unsigned long long* __restrict__ stats, float* __restrict__ mom_out, int count_samples) {
    extern __shared__ double smem_d[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k_pad = MOM ? S.brems.k_pad : 0;
    double* mom = smem_d;
    if (MOM) {
        for (int i = tid; i < k_pad; i += NW * 32) mom[i] = 0.0;
        __syncthreads();
    }
    const DevAxisym& A = S.ax;
    const int64_t ray = blockIdx.x;
    const double ox = rays.origin[3 * ray], oy = rays.origin[3 * ray + 1], oz = rays.origin[3 * ray + 2];
    const double dwx = rays.direction[3 * ray], dwy = rays.direction[3 * ray + 1], dwz = rays.direction[3 * ray + 2];
    float dx, dy, dz;
    {
        const double d0 = xform_row(S.w2p, dwx, dwy, dwz, false), d1 = xform_row(S.w2p + 4, dwx, dwy, dwz, false),
                     d2 = xform_row(S.w2p + 8, dwx, dwy, dwz, false);
        const double dl = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        dx = (float)(d0 / dl); dy = (float)(d1 / dl); dz = (float)(d2 / dl);
    }
    float vdc[NSP];
#pragma unroll
    for (int s = 0; s < NSP; s++) vdc[s] = FM.sp_v[s][0] * dx + FM.sp_v[s][1] * dy + FM.sp_v[s][2] * dz;
    unsigned long long n_samples = 0;
    unsigned n_brems = 0, ood = 0;
    const int n_comp = S.n_comp, row_f4 = FM.row_f4;
    const float4* const row0 = FM.edge ? FM.edge : FM.core;

    int64_t G0 = gbase[ray];
    int g_rot = 0;
    for (int64_t sg = rays.seg_offset[ray]; sg < rays.seg_offset[ray + 1]; sg++) {
        const SegGeom sgm = segment_geometry(S.w2p, S.step, S.min_samples, ox, oy, oz, dwx, dwy, dwz, rays.seg_t0[sg], rays.seg_t1[sg]);
        const int iv = sgm.iv;
        if (iv <= 0) continue;
        const float hf = (float)sgm.h;
        if (tid == 0) n_samples += (unsigned long long)iv + 1ull;
        const int n_groups = iv / 32 + 1;
        int first = (warp - g_rot) % NW;
        if (first < 0) first += NW;
        g_rot = (g_rot + n_groups) % NW;

        for (int g = first; g < n_groups; g += NW) {
            const int k = g * 32 + lane;
            const bool active = k <= iv;
            const double tk = __dmul_rn((double)k, sgm.h);
            const double pxd = __dadd_rn(sgm.sx, __dmul_rn(tk, sgm.ivx)), pyd = __dadd_rn(sgm.sy, __dmul_rn(tk, sgm.ivy)),
                         pzd = __dadd_rn(sgm.sz, __dmul_rn(tk, sgm.ivz));
            float w = active ? ((k == 0 || k == iv) ? 0.5f * hf : hf) : 0.f;
            // (R, Z, phi) as ax_setup has them
            const double r64 = __dsqrt_rn(__dadd_rn(__dmul_rn(pxd, pxd), __dmul_rn(pyd, pyd)));
            const float R = (float)r64, Z = (float)pzd;
            const float inv_r = R > 0.f ? 1.0f / R : 0.f;
            const float cphi = R > 0.f ? (float)pxd * inv_r : 1.f, sphi = (float)pyd * inv_r;
            int cls = 0;                                            // 0 vacuum, 1 table row, 2 blend zone
            const float4 *pa = row0, *pb = row0;
            float t = 0.f, ebr = 0.f, ebz = 0.f;
            if (active) {
                float m = 0.f, psi = 0.f;
                Cell2 cell;
                if (polygon_contains(A, R, Z, r64, pzd)) {
                    cell = locate2d(A.psin, R, Z);
                    if (!cell.inside) ood++;
                    psi = fmaxf(eval2d(A.psin, cell), 0.f);
                    if (psi <= 1.0f) {
                        m = A.mask_y[0];
                        const float p = fminf(fmaxf(psi, A.mask_x[0]), A.mask_x[A.n_mask - 1]);
                        for (int q = 0; q + 1 < A.n_mask; q++)
                            if (p >= A.mask_x[q] && p <= A.mask_x[q + 1]) {
                                m = A.mask_y[q] + (p - A.mask_x[q]) / (A.mask_x[q + 1] - A.mask_x[q]) * (A.mask_y[q + 1] - A.mask_y[q]);
                                break;
                            }
                    }
                }
                if (m >= 1.0f && FM.core_n > 0) {
                    cls = 1;
                    const float fi = fminf(psi, FM.psi_max) * FM.core_scale;
                    const int i0 = min((int)fi, FM.core_n - 1);
                    t = fi - (float)i0;
                    pa = FM.core + (size_t)i0 * row_f4;
                    pb = pa + row_f4;
                    if (S.need_pol) {
                        const float br = -eval2d(A.dpsi_dz, cell), bz = eval2d(A.dpsi_dr, cell);
                        const float n2 = br * br + bz * bz;
                        if (n2 > 0.f) { const float inv = rsqrtf(n2); ebr = br * inv; ebz = bz * inv; }
                    }
                } else if (m > 0.f) {
                    cls = 2;
                } else {
                    const int tri = mesh_locate(A, r64, pzd);
                    if (tri >= 0) { cls = 1; pa = pb = FM.edge + (size_t)tri * row_f4; ebr = 1.f; }
                }
            }
            const unsigned live_mask = __ballot_sync(FULL, cls != 0 && w > 0.f);
            const unsigned blend_mask = __ballot_sync(FULL, cls == 2 && w > 0.f);
            const int64_t G = G0 + g;
            if (lane == 0) { gmask[G] = live_mask; gblend[G] = blend_mask; }
            if (!live_mask) continue;
            if (cls != 1) w = 0.f;
            // species block: sqrt(Ts) and v.d
            const float aa = cphi * dx + sphi * dy, bb = cphi * dy - sphi * dx;
            float sq[NSP], vd[NSP];
#pragma unroll
            for (int s = 0; s < NSP; s++) {
                sq[s] = 0.f; vd[s] = 0.f;
                if (s < FM.n_sp) {
                    const float4 q = lerp4(__ldg(pa + s), __ldg(pb + s), t);
                    sq[s] = q.x;
                    const float cr = ebr * q.z - ebz * q.w, cz = ebz * q.z + ebr * q.w;
                    vd[s] = FM.sp_const[s] ? vdc[s] : fmaf(cr, aa, fmaf(q.y, bb, cz * dz));
                }
            }
            float* grec = rec + (size_t)G * n_comp * REC_FLOATS_PER_COMP + lane;
            for (int l4 = 0; l4 < FM.n_lines; l4 += 4) {
                const float4 q = lerp4(__ldg(pa + FM.off_amp + (l4 >> 2)), __ldg(pb + FM.off_amp + (l4 >> 2)), t);
                const float av[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (l4 + j >= FM.n_lines) break;
                    const MemoLine& L = FM.lines[l4 + j];
                    float sqs = sq[0], vds = vd[0];
#pragma unroll
                    for (int s = 1; s < NSP; s++)
                        if (L.slot == s) { sqs = sq[s]; vds = vd[s]; }
                    float amp = av[j] * w;
                    const float width = L.sigma_coef * sqs;
                    if (!(amp > 0.f) || !(sqs > 0.f)) amp = 0.f;
                    const bool any_on = __any_sync(FULL, amp > 0.f);
                    float* r = grec + L.rec_off;
                    if (L.shape == CB2_SHAPE_GAUSSIAN) {
                        r[64] = amp;
                        if (any_on) { r[0] = fmaf(L.shift_coef, vds, L.c0_frac); r[32] = width; }
                    } else {
                        const float dop = vds * L.inv_c;
                        for (int kc = 0; kc < L.ncomp; kc++, r += REC_FLOATS_PER_COMP) {       // multiplet.pyx:108-115
                            r[64] = amp * __ldg(L.mult_ratio + kc);
                            if (any_on) { r[0] = S.comps[L.comp0 + kc].c0_frac + __ldg(L.mult_lambda + kc) * dop * L.inv_delta; r[32] = width; }
                        }
                    }
                }
            }
            {
                // Bremsstrahlung block (and the out-of-domain count of the row's state)
                const float4 a0 = __ldg(pa + FM.off_brems);
                if (w > 0.f) ood += (unsigned)a0.y;
                if (MOM) {
                    const float4 q0 = lerp4(a0, __ldg(pb + FM.off_brems), t);
                    const float4 q1 = lerp4(__ldg(pa + FM.off_brems + 1), __ldg(pb + FM.off_brems + 1), t);
                    float U[CB2_MAX_BREMS_Z] = {q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, 0.f, 0.f};
                    if (FM.n_z > 6) {
                        const float4 q2 = lerp4(__ldg(pa + FM.off_brems + 2), __ldg(pb + FM.off_brems + 2), t);
                        U[6] = q2.x; U[7] = q2.y;
                    }
                    const bool lb = w > 0.f && q0.x >= 1.0f;
                    if (lb) n_brems += (unsigned)S.bins;
                    brems_scatter(S.brems, lb, w, q0.x, U, mom, lane);
                }
            }
        }
        G0 += n_groups;
    }
    if (MOM) {
        __syncthreads();
        float* row = mom_out + (size_t)ray * k_pad;
        for (int i = tid; i < k_pad; i += NW * 32) row[i] = (float)mom[i];
    }
    if (stats) {
        unsigned long long nb = n_brems, oodl = ood;
        for (int off = 16; off > 0; off >>= 1) {
            nb += __shfl_down_sync(FULL, nb, off);
            oodl += __shfl_down_sync(FULL, oodl, off);
        }
        if (lane == 0) {
            if (nb) atomicAdd(stats + 3, nb);
            if (oodl) atomicAdd(stats + 5, oodl);
        }
        if (tid == 0 && n_samples && count_samples) atomicAdd(stats + 0, n_samples);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K1 fused (table-driven scenes): state_fast_kernel's per-sample state feeds component_pass of the same warp directly — lanes are
// samples in both halves, so a lane's (centre, width, amplitude) never leaves the SM.  No line records cross HBM (they were 35x the
// algorithmic bytes of the frame) and the table gathers of one warp overlap the bin arithmetic of the others.
// The hot code stays small (table state ~ 20 KB + one rolled call site of component_pass), which the generic fused kernel of round 1
// was not (instruction-fetch bound).  Blend-zone samples contribute nothing here: their groups are flagged (gblend) and get
// zero-amplitude record rows, state_kernel<FIX> fills in the flagged lanes and bin_kernel adds those groups to the frame afterwards.
// Shared memory: [moments double[k_pad]] [NW private accumulators AccT[bins_pad]] [NW x (n_lines + 2 NSP) x 32 float stage].
// ------------------------------------------------------------------------------------------------------------------
template <int NW, int MOM, int MINB, int NSP, typename AccT>
__global__ void __launch_bounds__(NW * 32, MINB)
fused_fast_kernel(const __grid_constant__ DevScene S, const __grid_constant__ DevMemo FM, DevRays rays, const int64_t* __restrict__ gbase,
                  unsigned* __restrict__ gblend, float* __restrict__ rec, unsigned long long* __restrict__ stats,
                  float* __restrict__ mom_out, void* __restrict__ out, int out_f64, double scale, int accumulate, int count_samples) {
    extern __shared__ double smem_d[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k_pad = MOM ? S.brems.k_pad : 0;
    const int bins = S.bins, bins_pad = (bins + 31) & ~31;
    const int n_lines = FM.n_lines;
    double* mom = smem_d;
    AccT* wall = reinterpret_cast<AccT*>(smem_d + k_pad);
    AccT* wacc = wall + (size_t)warp * bins_pad;
    float* stage = reinterpret_cast<float*>(wall + (size_t)NW * bins_pad) + (size_t)warp * (n_lines + 2 * NSP) * 32 + lane;
    if (MOM)
        for (int i = tid; i < k_pad; i += NW * 32) mom[i] = 0.0;
    for (int i = lane; i < bins_pad; i += 32) wacc[i] = (AccT)0;
    __syncthreads();
    const DevAxisym& A = S.ax;
    const int64_t ray = blockIdx.x;
    const double ox = rays.origin[3 * ray], oy = rays.origin[3 * ray + 1], oz = rays.origin[3 * ray + 2];
    const double dwx = rays.direction[3 * ray], dwy = rays.direction[3 * ray + 1], dwz = rays.direction[3 * ray + 2];
    float dx, dy, dz;
    {
        const double d0 = xform_row(S.w2p, dwx, dwy, dwz, false), d1 = xform_row(S.w2p + 4, dwx, dwy, dwz, false),
                     d2 = xform_row(S.w2p + 8, dwx, dwy, dwz, false);
        const double dl = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        dx = (float)(d0 / dl); dy = (float)(d1 / dl); dz = (float)(d2 / dl);
    }
    float vdc[NSP];
#pragma unroll
    for (int s = 0; s < NSP; s++) vdc[s] = FM.sp_v[s][0] * dx + FM.sp_v[s][1] * dy + FM.sp_v[s][2] * dz;
    unsigned long long n_samples = 0;
    unsigned n_brems = 0, ood = 0, n_gauss = 0, n_lorentz = 0;
    const int n_comp = S.n_comp, row_f4 = FM.row_f4;
    const float4* const row0 = FM.edge ? FM.edge : FM.core;

    int64_t G0 = gbase[ray];
    int g_rot = 0;
    for (int64_t sg = rays.seg_offset[ray]; sg < rays.seg_offset[ray + 1]; sg++) {
        const SegGeom sgm = segment_geometry(S.w2p, S.step, S.min_samples, ox, oy, oz, dwx, dwy, dwz, rays.seg_t0[sg], rays.seg_t1[sg]);
        const int iv = sgm.iv;
        if (iv <= 0) continue;
        const float hf = (float)sgm.h;
        if (tid == 0) n_samples += (unsigned long long)iv + 1ull;
        const int n_groups = iv / 32 + 1;
        int first = (warp - g_rot) % NW;
        if (first < 0) first += NW;
        g_rot = (g_rot + n_groups) % NW;

        for (int g = first; g < n_groups; g += NW) {
            const int k = g * 32 + lane;
            const bool active = k <= iv;
            const double tk = __dmul_rn((double)k, sgm.h);
            const double pxd = __dadd_rn(sgm.sx, __dmul_rn(tk, sgm.ivx)), pyd = __dadd_rn(sgm.sy, __dmul_rn(tk, sgm.ivy)),
                         pzd = __dadd_rn(sgm.sz, __dmul_rn(tk, sgm.ivz));
            float w = active ? ((k == 0 || k == iv) ? 0.5f * hf : hf) : 0.f;
            const double r64 = __dsqrt_rn(__dadd_rn(__dmul_rn(pxd, pxd), __dmul_rn(pyd, pyd)));
            const float R = (float)r64, Z = (float)pzd;
            const float inv_r = R > 0.f ? 1.0f / R : 0.f;
            const float cphi = R > 0.f ? (float)pxd * inv_r : 1.f, sphi = (float)pyd * inv_r;
            int cls = 0;                                            // 0 vacuum, 1 table row, 2 blend zone
            const float4 *pa = row0, *pb = row0;
            float t = 0.f, ebr = 0.f, ebz = 0.f;
            if (active) {
                float m = 0.f, psi = 0.f;
                Cell2 cell;
                if (polygon_contains(A, R, Z, r64, pzd)) {
                    cell = locate2d(A.psin, R, Z);
                    if (!cell.inside) ood++;
                    psi = fmaxf(eval2d(A.psin, cell), 0.f);
                    if (psi <= 1.0f) {
                        m = A.mask_y[0];
                        const float p = fminf(fmaxf(psi, A.mask_x[0]), A.mask_x[A.n_mask - 1]);
                        for (int q = 0; q + 1 < A.n_mask; q++)
                            if (p >= A.mask_x[q] && p <= A.mask_x[q + 1]) {
                                m = A.mask_y[q] + (p - A.mask_x[q]) / (A.mask_x[q + 1] - A.mask_x[q]) * (A.mask_y[q + 1] - A.mask_y[q]);
                                break;
                            }
                    }
                }
                if (m >= 1.0f && FM.core_n > 0) {
                    cls = 1;
                    const float fi = fminf(psi, FM.psi_max) * FM.core_scale;
                    const int i0 = min((int)fi, FM.core_n - 1);
                    t = fi - (float)i0;
                    pa = FM.core + (size_t)i0 * row_f4;
                    pb = pa + row_f4;
                    if (S.need_pol) {
                        const float br = -eval2d(A.dpsi_dz, cell), bz = eval2d(A.dpsi_dr, cell);
                        const float n2 = br * br + bz * bz;
                        if (n2 > 0.f) { const float inv = rsqrtf(n2); ebr = br * inv; ebz = bz * inv; }
                    }
                } else if (m > 0.f) {
                    cls = 2;
                } else {
                    const int tri = mesh_locate(A, r64, pzd);
                    if (tri >= 0) { cls = 1; pa = pb = FM.edge + (size_t)tri * row_f4; ebr = 1.f; }
                }
            }
            const unsigned live_mask = __ballot_sync(FULL, cls != 0 && w > 0.f);
            const unsigned blend_mask = __ballot_sync(FULL, cls == 2 && w > 0.f);
            const int64_t G = G0 + g;
            if (lane == 0) gblend[G] = blend_mask;
            if (!live_mask) continue;
            if (cls != 1) w = 0.f;
            // species block: sqrt(Ts) and v.d, staged per lane for the rolled component loop
            const float aa = cphi * dx + sphi * dy, bb = cphi * dy - sphi * dx;
#pragma unroll
            for (int s = 0; s < NSP; s++) {
                float sq = 0.f, vd = 0.f;
                if (s < FM.n_sp) {
                    const float4 q = lerp4(__ldg(pa + s), __ldg(pb + s), t);
                    sq = q.x;
                    const float cr = ebr * q.z - ebz * q.w, cz = ebz * q.z + ebr * q.w;
                    vd = FM.sp_const[s] ? vdc[s] : fmaf(cr, aa, fmaf(q.y, bb, cz * dz));
                }
                stage[(n_lines + s) * 32] = sq;
                stage[(n_lines + NSP + s) * 32] = vd;
            }
            for (int l4 = 0; l4 < n_lines; l4 += 4) {
                const float4 q = lerp4(__ldg(pa + FM.off_amp + (l4 >> 2)), __ldg(pb + FM.off_amp + (l4 >> 2)), t);
                stage[l4 * 32] = q.x * w;
                if (l4 + 1 < n_lines) stage[(l4 + 1) * 32] = q.y * w;
                if (l4 + 2 < n_lines) stage[(l4 + 2) * 32] = q.z * w;
                if (l4 + 3 < n_lines) stage[(l4 + 3) * 32] = q.w * w;
            }
            {
                // Bremsstrahlung block (and the out-of-domain count of the row's state)
                const float4 a0 = __ldg(pa + FM.off_brems);
                if (w > 0.f) ood += (unsigned)a0.y;
                if (MOM) {
                    const float4 q0 = lerp4(a0, __ldg(pb + FM.off_brems), t);
                    const float4 q1 = lerp4(__ldg(pa + FM.off_brems + 1), __ldg(pb + FM.off_brems + 1), t);
                    float U[CB2_MAX_BREMS_Z] = {q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, 0.f, 0.f};
                    if (FM.n_z > 6) {
                        const float4 q2 = lerp4(__ldg(pa + FM.off_brems + 2), __ldg(pb + FM.off_brems + 2), t);
                        U[6] = q2.x; U[7] = q2.y;
                    }
                    const bool lb = w > 0.f && q0.x >= 1.0f;
                    if (lb) n_brems += (unsigned)S.bins;
                    brems_scatter(S.brems, lb, w, q0.x, U, mom, lane);
                }
            }
            if (blend_mask) {
                // the flagged lanes' rows come from the fix-up pass; every other lane of the group must read as dark there
                float* z = rec + (size_t)G * n_comp * REC_FLOATS_PER_COMP + 64 + lane;
                for (int c = 0; c < n_comp; c++) z[(size_t)c * REC_FLOATS_PER_COMP] = 0.f;
            }
            // line components of this group, straight into the warp's accumulator (one call site of component_pass)
            for (int l = 0; l < n_lines; l++) {
                const MemoLine& L = FM.lines[l];
                const float sqs = stage[(n_lines + L.slot) * 32], vds = stage[(n_lines + NSP + L.slot) * 32];
                float amp = stage[l * 32];
                if (!(amp > 0.f) || !(sqs > 0.f)) amp = 0.f;
                if (!__any_sync(FULL, amp > 0.f)) continue;
                const float width = L.sigma_coef * sqs;
                const bool gauss = L.shape == CB2_SHAPE_GAUSSIAN;
                const float dop = vds * L.inv_c;
                for (int kc = 0; kc < L.ncomp; kc++) {
                    const int c = L.comp0 + kc;
                    const float cf = gauss ? fmaf(L.shift_coef, vds, L.c0_frac)
                                           : S.comps[c].c0_frac + __ldg(L.mult_lambda + kc) * dop * L.inv_delta;     // multiplet.pyx:108-115
                    const float a = gauss ? amp : amp * __ldg(L.mult_ratio + kc);
                    component_pass<AccT, 0>(0, cf, width, a, S.comps[c].c0_int, bins, wacc, lane, nullptr, 0.0, n_gauss, n_lorentz);
                }
            }
        }
        G0 += n_groups;
    }
    __syncthreads();
    if (MOM) {
        float* row = mom_out + (size_t)ray * k_pad;
        for (int i = tid; i < k_pad; i += NW * 32) row[i] = (float)mom[i];
    }
    // the ray's spectrum: fp64 sum of the NW private accumulators, lane-consecutive bins -> coalesced rows
    for (int bin = tid; bin < bins; bin += NW * 32) {
        double v = 0.0;
#pragma unroll
        for (int wq = 0; wq < NW; wq++) v += (double)wall[(size_t)wq * bins_pad + bin];
        v *= scale;
        const size_t idx = (size_t)ray * bins + bin;
        if (out_f64) {
            double* p = (double*)out + idx;
            *p = (accumulate ? *p : 0.0) + v;
        } else {
            float* p = (float*)out + idx;
            *p = (float)((accumulate ? (double)*p : 0.0) + v);
        }
    }
    if (stats) {
        unsigned long long nb = n_brems, oodl = ood, ng = n_gauss;
        for (int off = 16; off > 0; off >>= 1) {
            nb += __shfl_down_sync(FULL, nb, off);
            oodl += __shfl_down_sync(FULL, oodl, off);
            ng += __shfl_down_sync(FULL, ng, off);
        }
        if (lane == 0) {
            if (nb) atomicAdd(stats + 3, nb);
            if (oodl) atomicAdd(stats + 5, oodl);
            if (ng) atomicAdd(stats + 1, ng);
        }
        if (tid == 0 && n_samples && count_samples) atomicAdd(stats + 0, n_samples);
    }
}

// state tables of an eligible scene: plasma line models with Gaussian / multiplet shapes (+ Bremsstrahlung moments) over fields
// that are all AXISYM_BLEND.  CB2_STATE_MEMO=0 keeps the generic kernel (tests run both).
int cb2_memo_build(cb2_scene* sc) {
    DevMemo& FM = sc->memo;
    memset(&FM, 0, sizeof FM);
    const DevScene& S = sc->host;
    if (const char* e = getenv("CB2_STATE_MEMO")) if (atoi(e) == 0) return CB2_OK;
    if (!sc->warp_kernel || !sc->ax_only || sc->feat || !S.ax.present || S.beam.present || S.need_b) return CB2_OK;
    int n_lines = 0, n_sp = 0;
    for (int m = 0; m < S.n_models; m++) {
        const DevModel& M = S.models[m];
        if (M.kind == CB2_MODEL_BREMSSTRAHLUNG) continue;
        if (M.kind != CB2_MODEL_EXCITATION_LINE && M.kind != CB2_MODEL_RECOMBINATION_LINE) return CB2_OK;
        if (M.shape != CB2_SHAPE_GAUSSIAN && M.shape != CB2_SHAPE_MULTIPLET) return CB2_OK;
        if (n_lines == CB2_MEMO_MAX_LINES) return CB2_OK;
        int slot = -1;
        for (int s = 0; s < n_sp; s++) if (FM.sp_species[s] == M.species) slot = s;
        if (slot < 0) {
            if (n_sp == CB2_MEMO_MAX_SP) return CB2_OK;
            slot = n_sp++;
            FM.sp_species[slot] = M.species;
            const DevVector& v = S.species[M.species].velocity;
            FM.sp_const[slot] = v.kind == CB2_FIELD_CONSTANT;
            for (int k = 0; k < 3; k++) FM.sp_v[slot][k] = FM.sp_const[slot] ? v.c[k] : 0.f;
        }
        MemoLine& L = FM.lines[n_lines++];
        L.model = m; L.slot = slot; L.comp0 = M.comp0; L.ncomp = M.ncomp; L.shape = M.shape;
        L.sigma_coef = M.sigma_coef; L.wavelength = M.wavelength; L.inv_c = M.inv_c; L.inv_delta = M.inv_delta;
        L.rec_off = M.comp0 * REC_FLOATS_PER_COMP;
        L.c0_frac = S.comps[M.comp0].c0_frac;
        L.shift_coef = M.wavelength * M.inv_c * M.inv_delta;
        L.mult_ratio = M.mult_ratio; L.mult_lambda = M.mult_lambda;
    }
    const bool brems = S.brems.present && S.brems.mode == 3;
    if (S.brems.present && !brems) return CB2_OK;
    if (n_lines == 0 && !brems) return CB2_OK;
    FM.n_lines = n_lines; FM.n_sp = n_sp; FM.has_brems = brems; FM.n_z = brems ? S.brems.n_z : 0;
    FM.off_amp = n_sp;
    FM.off_brems = n_sp + (n_lines + 3) / 4;
    FM.row_f4 = FM.off_brems + (FM.n_z > 6 ? 3 : 2);
    // the psi_n grid ends at the last mask knot with full core weight (0.94 on Generomak, plasma.py:610): beyond it the rows would
    // never be used, and the profiles' steep fall towards the separatrix would only force a finer grid
    FM.psi_max = 0.f;
    for (int k = 0; k < S.ax.n_mask; k++)
        if (S.ax.mask_y[k] >= 1.0f) FM.psi_max = std::max(FM.psi_max, std::min(S.ax.mask_x[k], 1.0f));
    const bool want_core = FM.psi_max > 0.f;
    const size_t row_bytes = (size_t)FM.row_f4 * sizeof(float4);
    float4* edge = nullptr;
    float4* core = nullptr;
    unsigned long long* err_dev = nullptr;     // [0] error key, then the column maxima (unsigned[64])
    int rc = CB2_OK;
    do {
        if (S.ax.n_tri > 0) {
            if ((rc = cb2_cuda_check(cudaMalloc((void**)&edge, row_bytes * S.ax.n_tri), "cudaMalloc(edge state table)")) != CB2_OK) break;
            memo_rows_kernel<<<(S.ax.n_tri + 127) / 128, 128>>>(S, FM, 0, S.ax.n_tri, 0, edge, nullptr, nullptr);
        }
        if ((rc = cb2_cuda_check(cudaMalloc((void**)&err_dev, sizeof(unsigned long long) + 64 * sizeof(unsigned)), "cudaMalloc")) != CB2_OK) break;
        unsigned* colmax = reinterpret_cast<unsigned*>(err_dev + 1);
        int n = 32768;
        if (const char* e = getenv("CB2_MEMO_ROWS")) n = std::max(256, atoi(e));
        const float tol = 4e-6f;
        for (; want_core; n *= 2) {
            FM.core_n = n;
            FM.core_scale = (float)n / FM.psi_max;
            if ((rc = cb2_cuda_check(cudaMalloc((void**)&core, row_bytes * (size_t)(n + 1)), "cudaMalloc(core state table)")) != CB2_OK) break;
            memo_rows_kernel<<<(n + 1 + 127) / 128, 128>>>(S, FM, 1, n + 1, 0, core, nullptr, nullptr);
            if ((rc = cb2_cuda_check(cudaMemset(err_dev, 0, sizeof(unsigned long long) + 64 * sizeof(unsigned)), "cudaMemset")) != CB2_OK) break;
            memo_rows_kernel<<<(n + 1 + 127) / 128, 128>>>(S, FM, 1, n + 1, 2, core, colmax, nullptr);
            memo_rows_kernel<<<(n + 127) / 128, 128>>>(S, FM, 1, n, 1, core, colmax, err_dev);
            unsigned long long key = 0;
            if ((rc = cb2_cuda_check(cudaMemcpy(&key, err_dev, sizeof key, cudaMemcpyDeviceToHost), "state table check")) != CB2_OK) break;
            const unsigned bits = (unsigned)(key >> 32);
            memcpy(&sc->memo_err, &bits, sizeof bits);
            sc->memo_err_at = (int64_t)(key & 0xffffffffull);
            if (getenv("CB2_MEMO_DEBUG")) fprintf(stderr, "state table: %d intervals, worst mid-interval error %.3g (entry %d, row %d)\n", n, sc->memo_err, (int)((key >> 24) & 255), (int)(key & 0xffffff));
            if (sc->memo_err <= tol) break;
            cudaFree(core);
            core = nullptr;
            if (n >= 262144) { FM.core_n = 0; break; }          // no acceptable table: core samples take the generic kernel too
        }
        if (rc != CB2_OK) break;
        rc = cb2_cuda_check(cudaDeviceSynchronize(), "state tables");
    } while (0);
    if (err_dev) cudaFree(err_dev);
    if (rc != CB2_OK) {
        if (edge) cudaFree(edge);
        if (core) cudaFree(core);
        memset(&FM, 0, sizeof FM);
        return rc;
    }
    FM.edge = edge;
    FM.core = core;
    if (!FM.edge && !FM.core) { memset(&FM, 0, sizeof FM); return CB2_OK; }
    FM.enabled = 1;
    if (const char* e = getenv("CB2_FUSED")) sc->fused = atoi(e) != 0;
    return CB2_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// K1b: line records -> spectral bins
// ------------------------------------------------------------------------------------------------------------------
template <int NW, typename AccT, int LOR>
__global__ void __launch_bounds__(NW * 32, 768 / (NW * 32))
bin_kernel(const __grid_constant__ DevScene S, int64_t n_rays, const int64_t* __restrict__ gbase, const unsigned* __restrict__ gmask,
           const float* __restrict__ rec, const double* __restrict__ flat, void* __restrict__ out, int out_f64, double scale,
           int accumulate, unsigned long long* __restrict__ stats) {
    extern __shared__ double smem_d[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bins = S.bins, bins_pad = (bins + 31) & ~31;
    AccT* wall = reinterpret_cast<AccT*>(smem_d);
    AccT* wacc = wall + (size_t)warp * bins_pad;
    for (int i = lane; i < bins_pad; i += 32) wacc[i] = (AccT)0;
    __syncwarp();
    const int64_t ray = blockIdx.x;
    const int n_comp = S.n_comp;
    unsigned n_gauss = 0, n_lorentz = 0;
    int touched = 0;
    const int64_t G0 = gbase[ray], G1 = gbase[ray + 1];
    if constexpr (NW >= 8) {
    // 8- and 16-warp instances (calls with few rays): work items = (group, component) pairs dealt round-robin to the warps — a ray with few
    // groups but many components (C2: 2 groups x 30 Stark / Zeeman components, each a serial walk over its bin range) still occupies
    // every warp.  The next item's record is in flight while this one is binned.
    const unsigned n_items = (unsigned)(G1 - G0) * (unsigned)n_comp;
    unsigned it = warp;
    float amp_n = 0.f, cf_n = 0.f, width_n = 0.f;
    int c_n = 0;
    auto fetch = [&](unsigned item) {
        const unsigned g = item / (unsigned)n_comp;
        c_n = (int)(item - g * (unsigned)n_comp);
        amp_n = 0.f;
        if (__ldg(gmask + G0 + g)) {
            const float* r = rec + ((size_t)(G0 + g) * n_comp + c_n) * REC_FLOATS_PER_COMP + lane;
            amp_n = __ldg(r + 64); cf_n = __ldg(r); width_n = __ldg(r + 32);
        }
    };
    if (it < n_items) fetch(it);
    for (; it < n_items; it += NW) {
        const float amp = amp_n, cf = cf_n, width = width_n;
        const int c = c_n;
        if (it + NW < n_items) fetch(it + NW);
        if (!__any_sync(FULL, amp > 0.f)) continue;
        touched = 1;
        component_pass<AccT, LOR>(S.comps[c].type, cf, width, amp, S.comps[c].c0_int, bins, wacc, lane, S.lorentz_tab,
                                  S.lorentz_phi_inf, n_gauss, n_lorentz);
    }
    } else {
    for (int64_t G = G0 + warp; G < G1 && n_comp > 0; G += NW) {
        if (!__ldg(gmask + G)) continue;
        touched = 1;
        const float* grec = rec + (size_t)G * n_comp * REC_FLOATS_PER_COMP + lane;
        // software pipeline: the next component's record is in flight while this one is binned (records come from HBM/L2)
        float amp_n = __ldg(grec + 64), cf_n = __ldg(grec), width_n = __ldg(grec + 32);
        for (int c = 0; c < n_comp; c++) {
            const float amp = amp_n, cf = cf_n, width = width_n;
            if (c + 1 < n_comp) {
                const float* r = grec + (size_t)(c + 1) * REC_FLOATS_PER_COMP;
                amp_n = __ldg(r + 64); cf_n = __ldg(r); width_n = __ldg(r + 32);
            }
            if (!__any_sync(FULL, amp > 0.f)) continue;
            component_pass<AccT, LOR>(S.comps[c].type, cf, width, amp, S.comps[c].c0_int, bins, wacc, lane, S.lorentz_tab,
                                      S.lorentz_phi_inf, n_gauss, n_lorentz);
        }
    }
    }
    // (a ray without a live group adds nothing: in accumulate mode its row is left alone — the pass over the flagged groups of the
    // fused path touches a fraction of the rays)
    const int any_touched = __syncthreads_or(touched);
    const double flat_v = flat ? flat[ray] : 0.0;
    for (int bin = tid; bin < bins && (any_touched || !accumulate || flat_v != 0.0); bin += NW * 32) {
        double v = flat_v;
#pragma unroll
        for (int w = 0; w < NW; w++) v += (double)wall[(size_t)w * bins_pad + bin];
        v *= scale;
        const size_t idx = (size_t)ray * bins + bin;
        if (out_f64) {
            double* p = (double*)out + idx;
            *p = (accumulate ? *p : 0.0) + v;
        } else {
            float* p = (float*)out + idx;
            *p = (float)((accumulate ? (double)*p : 0.0) + v);
        }
    }
    if (stats) {
        unsigned long long ng = n_gauss, nl = n_lorentz;
        for (int off = 16; off > 0; off >>= 1) {
            ng += __shfl_down_sync(FULL, ng, off);
            nl += __shfl_down_sync(FULL, nl, off);
        }
        if (lane == 0) {
            if (ng) atomicAdd(stats + 1, ng);
            if (nl) atomicAdd(stats + 2, nl);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// launch: per batch of rays  count groups -> scan -> state_kernel -> bin_kernel -> contraction of the moments
// ------------------------------------------------------------------------------------------------------------------
size_t cb2_warp_smem_bytes(int nw, int acc_f64, int bins) {
    const size_t bins_pad = ((size_t)bins + 31) & ~(size_t)31;
    return (size_t)nw * bins_pad * (acc_f64 ? sizeof(double) : sizeof(float));
}

int64_t cb2_warp_batch_rays(const cb2_scene* sc) {
    // rays per batch: every kernel launch ends with a partly filled wave of CTAs (148 SMs x 6..8 resident), so bigger batches waste
    // less — C3: 48.4 ms per 65 536 rays at 16 384, 47.3 at 32 768, 46.7 at 65 536 (record buffer 3.9 GB per 16 384 rays)
    int64_t b = 65536;
    if (const char* e = getenv("CB2_BATCH_RAYS")) { const long v = atol(e); if (v >= 128) b = v / 128 * 128; }
    if (sc->host.brems.present && sc->host.brems.mode == 3) b = std::min(b, cb2_moment_batch(sc->host.brems.k_pad));
    return b;
}

template <typename T>
static int reserve(T** p, size_t* have, size_t need, cudaStream_t st) {
    if (*have >= need && *p) return CB2_OK;
    CB2_CUDA(cudaStreamSynchronize(st));
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    const size_t cap = need + need / 8 + 256;
    CB2_CUDA(cudaMalloc((void**)p, cap));
    *have = cap;
    return CB2_OK;
}

template <int NW, typename AccT>
static int launch_bin(const cb2_scene* sc, const unsigned* mask, int64_t n_rays, void* out, int out_f64, double scale, int accumulate,
                      unsigned long long* stats, cudaStream_t st) {
    const DevScene& S = sc->host;
    const size_t smem = cb2_warp_smem_bytes(NW, sizeof(AccT) == 8, S.bins);
    if (S.has_lorentz) {
        auto kern = bin_kernel<NW, AccT, 1>;
        if (smem > 48 * 1024) CB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dim3((unsigned)n_rays), dim3(NW * 32), smem, st>>>(S, n_rays, sc->gbase, mask, sc->rec, S.has_flat ? sc->flat : nullptr, out, out_f64, scale, accumulate, stats);
    } else {
        auto kern = bin_kernel<NW, AccT, 0>;
        if (smem > 48 * 1024) CB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dim3((unsigned)n_rays), dim3(NW * 32), smem, st>>>(S, n_rays, sc->gbase, mask, sc->rec, S.has_flat ? sc->flat : nullptr, out, out_f64, scale, accumulate, stats);
    }
    return cb2_cuda_check(cudaGetLastError(), "bin_kernel launch");
}

int cb2_launch_emission_warp(cb2_scene* sc, const DevRays& rays, void* out, int out_f64, double scale, int accumulate,
                             unsigned long long* stats, int count_samples, cudaStream_t st) {
    const DevScene& S = sc->host;
    const DevBrems& B = S.brems;
    const bool moments = B.present && B.mode == 3;
    const int n_comp = S.n_comp;
    static const int dbg = getenv("CB2_DBG_SKIP") ? atoi(getenv("CB2_DBG_SKIP")) : 0;
    // table-driven scenes, CB2_FUSED=1 at scene creation: state and binning in one kernel (no line records through HBM, no record
    // buffer; measured 10 % slower than state_fast_kernel -> records -> bin_kernel on C3: 16 resident warps per SM instead of 24)
    const int fused_env = sc->fused;
    const int fused_nsp = sc->memo.n_sp <= 2 ? 2 : CB2_MEMO_MAX_SP;
    const size_t fsmem = (moments ? (size_t)B.k_pad * sizeof(double) : 0) + cb2_warp_smem_bytes(4, sc->acc_f64, S.bins) +
                         4 * (size_t)(sc->memo.n_lines + 2 * fused_nsp) * 32 * sizeof(float);
    const bool fused = fused_env && sc->memo.enabled && sc->bin_nw == 4 && !S.has_lorentz && !S.has_flat && fsmem <= 200 * 1024;
    const size_t esz = out_f64 ? sizeof(double) : sizeof(float);
    const size_t rec_cap_bytes = (size_t)24 << 30;            // bound on the record buffer; batches shrink to respect it
    int64_t batch = std::min(cb2_warp_batch_rays(sc), rays.n_rays);
    // host-buffer calls copy a finished batch's rows back while the next batch computes: at least eight batches per call keep the
    // copy of the last one (which nothing hides) small
    if (sc->d2h_host && !getenv("CB2_BATCH_RAYS")) batch = std::min(batch, std::max<int64_t>(16384, (rays.n_rays / 8 + 127) / 128 * 128));
    int rc;
    void *pend_dst = nullptr, *pend_src = nullptr;          // deferred device -> host copy of the previous batch
    size_t pend_bytes = 0;
    int64_t pend_r0 = 0, pend_n = 0;
    auto flush_pending = [&]() -> int {
        int rc2 = CB2_OK;
        if (sc->d2h_rows) rc2 = cb2_d2h_rows(sc->d2h_host, sc->d2h_rows + pend_r0, pend_n, pend_src, (size_t)S.bins * esz, sc->copy_stream, sc->copy_stream2);
        else rc2 = cb2_cuda_check(cudaMemcpyAsync(pend_dst, pend_src, pend_bytes, cudaMemcpyDeviceToHost, sc->copy_stream), "frame rows device -> host");
        pend_bytes = 0;
        return rc2;
    };
    for (int64_t r0 = 0; r0 < rays.n_rays;) {
        DevRays sub = rays;
        sub.n_rays = std::min(batch, rays.n_rays - r0);
        // host-buffer calls: nothing hides the copy of the last batch's rows, so the call ends on a geometric run of small batches
        // (... 1/2, 1/4, ... down to 4096 rays: 33 MB of float32 rows at 2048 bins instead of a whole batch)
        if (sc->d2h_host && sub.n_rays == rays.n_rays - r0 && sub.n_rays >= 8192) sub.n_rays = (sub.n_rays / 2 + 127) / 128 * 128;
        sub.origin = rays.origin + 3 * r0;
        sub.direction = rays.direction + 3 * r0;
        sub.seg_offset = rays.seg_offset + r0;                  // entries are absolute segment indices
        void* o = (char*)out + (size_t)r0 * S.bins * esz;
        const bool prof = sc->prof_on != 0;
        if (prof) CB2_CUDA(cudaEventRecord(sc->prof_ev[0], st));
        // groups per ray -> offsets
        if ((rc = reserve(&sc->gbase, &sc->gbase_bytes, (size_t)(sub.n_rays + 1) * sizeof(int64_t), st)) != CB2_OK) return rc;
        count_groups_kernel<<<(unsigned)((sub.n_rays + 1 + 127) / 128), 128, 0, st>>>(S, sub, sc->gbase);
        if ((rc = cb2_cuda_check(cudaGetLastError(), "count_groups_kernel launch")) != CB2_OK) return rc;
        if ((rc = cb2_launch_scan(sc->gbase, sub.n_rays + 1, st)) != CB2_OK) return rc;
        if (!sc->total_host) {
            CB2_CUDA(cudaHostAlloc((void**)&sc->total_host, sizeof(int64_t), cudaHostAllocMapped));
            CB2_CUDA(cudaHostGetDevicePointer((void**)&sc->total_dev, sc->total_host, 0));
        }
        publish_total_kernel<<<1, 1, 0, st>>>(sc->gbase + sub.n_rays, sc->total_dev);
        CB2_CUDA(cudaStreamSynchronize(st));
        const int64_t n_groups = *(volatile int64_t*)sc->total_host;
        const size_t rec_bytes = (size_t)n_groups * std::max(n_comp, 1) * REC_FLOATS_PER_COMP * sizeof(float);
        if (rec_bytes > rec_cap_bytes && sub.n_rays > 128) {     // too many samples in this batch: halve it and retry
            batch = std::max<int64_t>(128, (sub.n_rays / 2 + 127) / 128 * 128);
            continue;
        }
        if ((rc = reserve(&sc->gmask, &sc->gmask_bytes, (size_t)std::max<int64_t>(n_groups, 1) * sizeof(unsigned), st)) != CB2_OK) return rc;
        if ((rc = reserve(&sc->rec, &sc->rec_bytes, std::max<size_t>(rec_bytes, 256), st)) != CB2_OK) return rc;
        if (moments && (rc = reserve(&sc->mom, &sc->mom_bytes, (size_t)sub.n_rays * B.k_pad * sizeof(float), st)) != CB2_OK) return rc;
        if (S.has_flat && (rc = reserve(&sc->flat, &sc->flat_bytes, (size_t)sub.n_rays * sizeof(double), st)) != CB2_OK) return rc;
        if (prof) CB2_CUDA(cudaEventRecord(sc->prof_ev[1], st));
        // K1a
        {
            const size_t smem = moments ? (size_t)B.k_pad * sizeof(double) : 0;
#define CB2_STATE(MOM, AX, FT, FX, GB)                                                                                               \
    do {                                                                                                                          \
        auto kern = state_kernel<4, MOM, AX, FT, FX>;                                                                             \
        if (smem > 32 * 1024) CB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        kern<<<dim3((unsigned)((FX) ? (sub.n_rays + 3) / 4 : sub.n_rays)), dim3(128), (FX) ? 0 : smem, st>>>(                     \
            S, sub, sc->gbase, sc->gmask, sc->rec, stats, sc->mom, S.has_flat ? sc->flat : nullptr, (GB) ? 0 : count_samples, dbg, GB, \
            sc->memo);                                                                                                            \
    } while (0)
            if (sc->memo.enabled) {
                // table-driven state kernel, then the generic kernel on the blend-zone samples it flagged
                if ((rc = reserve(&sc->gblend, &sc->gblend_bytes, (size_t)std::max<int64_t>(n_groups, 1) * sizeof(unsigned), st)) != CB2_OK) return rc;
#define CB2_FAST(MOM, MINB)                                                                                                      \
    do {                                                                                                                          \
        auto kern = sc->memo.n_sp <= 2 ? state_fast_kernel<4, MOM, MINB, 2> : state_fast_kernel<4, MOM, MINB, CB2_MEMO_MAX_SP>;    \
        if (smem > 32 * 1024) CB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        kern<<<dim3((unsigned)sub.n_rays), dim3(128), smem, st>>>(S, sc->memo, sub, sc->gbase, sc->gmask, sc->gblend, sc->rec, stats, \
                                                                  sc->mom, count_samples);                                       \
    } while (0)
                // resident CTAs per SM (register budget): CB2_FAST_MINB = 6 / 8 / 10 / 12 for experiments
                static const int minb = getenv("CB2_FAST_MINB") ? atoi(getenv("CB2_FAST_MINB")) : 8;
                if (fused) {
                    const int nsp = fused_nsp;
#define CB2_FUSED_K(MOM, NSPV, ACC)                                                                                              \
    do {                                                                                                                          \
        auto kern = fused_fast_kernel<4, MOM, CB2_FUSED_MINB, NSPV, ACC>;                                                         \
        if (fsmem > 48 * 1024) CB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));    \
        kern<<<dim3((unsigned)sub.n_rays), dim3(128), fsmem, st>>>(S, sc->memo, sub, sc->gbase, sc->gblend, sc->rec, stats, sc->mom, o,  \
                                                                   out_f64, scale, accumulate, count_samples);                    \
    } while (0)
                    const int sel = (moments ? 4 : 0) | (nsp > 2 ? 2 : 0) | (sc->acc_f64 ? 1 : 0);
                    switch (sel) {
                    case 0: CB2_FUSED_K(0, 2, float); break;
                    case 1: CB2_FUSED_K(0, 2, double); break;
                    case 2: CB2_FUSED_K(0, CB2_MEMO_MAX_SP, float); break;
                    case 3: CB2_FUSED_K(0, CB2_MEMO_MAX_SP, double); break;
                    case 4: CB2_FUSED_K(1, 2, float); break;
                    case 5: CB2_FUSED_K(1, 2, double); break;
                    case 6: CB2_FUSED_K(1, CB2_MEMO_MAX_SP, float); break;
                    default: CB2_FUSED_K(1, CB2_MEMO_MAX_SP, double); break;
                    }
#undef CB2_FUSED_K
                } else if (moments) {
                    if (minb <= 6) CB2_FAST(1, 6); else if (minb <= 8) CB2_FAST(1, 8); else if (minb <= 10) CB2_FAST(1, 10); else CB2_FAST(1, 12);
                } else {
                    if (minb <= 6) CB2_FAST(0, 6); else if (minb <= 8) CB2_FAST(0, 8); else if (minb <= 10) CB2_FAST(0, 10); else CB2_FAST(0, 12);
                }
#undef CB2_FAST
                if ((rc = cb2_cuda_check(cudaGetLastError(), "state_fast_kernel launch")) != CB2_OK) return rc;
                if (prof) CB2_CUDA(cudaEventRecord(sc->prof_ev[5], st));
                if (moments) CB2_STATE(1, 1, 0, 1, sc->gblend); else CB2_STATE(0, 1, 0, 1, sc->gblend);
            } else {
                const unsigned* none = nullptr;
                const int sel = (moments ? 4 : 0) | (sc->ax_only ? 2 : 0) | (sc->feat ? 1 : 0);
                switch (sel) {
                case 0: CB2_STATE(0, 0, 0, 0, none); break;
                case 1: CB2_STATE(0, 0, 1, 0, none); break;
                case 2: CB2_STATE(0, 1, 0, 0, none); break;
                case 3: CB2_STATE(0, 1, 1, 0, none); break;
                case 4: CB2_STATE(1, 0, 0, 0, none); break;
                case 5: CB2_STATE(1, 0, 1, 0, none); break;
                case 6: CB2_STATE(1, 1, 0, 0, none); break;
                default: CB2_STATE(1, 1, 1, 0, none); break;
                }
            }
#undef CB2_STATE
            if ((rc = cb2_cuda_check(cudaGetLastError(), "state_kernel launch")) != CB2_OK) return rc;
        }
        if (prof) CB2_CUDA(cudaEventRecord(sc->prof_ev[2], st));
        // K1b
        // (fused path: the frame rows exist already; only the groups with blend-zone samples are binned, on top of them)
        const unsigned* bmask = fused ? sc->gblend : sc->gmask;
        const int bacc = fused ? 1 : accumulate;
        // a call that cannot fill the GPU with one CTA per ray (0-D observer groups: C2 has 64 rays) takes 8 or 16 warps per ray; decided
        // on the call's ray count, not the batch's, so a ray's bits do not depend on where the batches are cut
        int bin_nw = sc->bin_nw;
        if (bin_nw == 4 && rays.n_rays < 1024 && cb2_warp_smem_bytes(8, sc->acc_f64, S.bins) <= 200 * 1024) bin_nw = 8;
        if (bin_nw == 8 && rays.n_rays < 256 && cb2_warp_smem_bytes(16, sc->acc_f64, S.bins) <= 200 * 1024) bin_nw = 16;
        if (bin_nw == 16) rc = sc->acc_f64 ? launch_bin<16, double>(sc, bmask, sub.n_rays, o, out_f64, scale, bacc, stats, st)
                                           : launch_bin<16, float>(sc, bmask, sub.n_rays, o, out_f64, scale, bacc, stats, st);
        else if (bin_nw == 8) rc = sc->acc_f64 ? launch_bin<8, double>(sc, bmask, sub.n_rays, o, out_f64, scale, bacc, stats, st)
                                          : launch_bin<8, float>(sc, bmask, sub.n_rays, o, out_f64, scale, bacc, stats, st);
        else if (bin_nw == 2) rc = sc->acc_f64 ? launch_bin<2, double>(sc, bmask, sub.n_rays, o, out_f64, scale, bacc, stats, st)
                                               : launch_bin<2, float>(sc, bmask, sub.n_rays, o, out_f64, scale, bacc, stats, st);
        else rc = sc->acc_f64 ? launch_bin<4, double>(sc, bmask, sub.n_rays, o, out_f64, scale, bacc, stats, st)
                              : launch_bin<4, float>(sc, bmask, sub.n_rays, o, out_f64, scale, bacc, stats, st);
        if (rc != CB2_OK) return rc;
        if (prof) CB2_CUDA(cudaEventRecord(sc->prof_ev[3], st));
        // K2
        if (moments) {
            rc = sc->contract_tc ? cb2_launch_contract_tc(sc, sc->mom, sub.n_rays, B.k_pad, B.n_pad, S.bins, o, out_f64, scale, st)
                                 : cb2_launch_contract(sc->mom, B.phi, sub.n_rays, B.k_pad, B.n_pad, S.bins, o, out_f64, scale, st);
            if (rc != CB2_OK) return rc;
        }
        if (prof) {
            CB2_CUDA(cudaEventRecord(sc->prof_ev[4], st));
            CB2_CUDA(cudaEventSynchronize(sc->prof_ev[4]));
            const int slot[4] = {3, 0, 1, 2};                   // helpers, state, bin, contract
            for (int i = 0; i < 4; i++) {
                float ms = 0.f;
                CB2_CUDA(cudaEventElapsedTime(&ms, sc->prof_ev[i], sc->prof_ev[i + 1]));
                sc->prof_ms[slot[i]] += ms;
            }
            if (sc->memo.enabled) {                             // the fix-up pass's share of the state time
                float ms = 0.f;
                CB2_CUDA(cudaEventElapsedTime(&ms, sc->prof_ev[5], sc->prof_ev[2]));
                sc->prof_fixup_ms += ms;
            }
            sc->prof_launches[3] += 2; sc->prof_launches[0] += 1; sc->prof_launches[1] += 1; sc->prof_launches[2] += moments ? 1 : 0;
        }
        // the previous batch's rows go to the host now, AFTER this batch's kernels are queued: issuing a batch's strided copies takes the
        // host ~ 0.6 ms (64 tiles per 16 384 rays), which the GPU would otherwise spend idle between two batches (the group count of
        // a batch is read back through mapped host memory, not through the copy engine, so nothing queues behind the copies)
        if (pend_bytes) {
            if ((rc = flush_pending()) != CB2_OK) return rc;
        }
        if (sc->d2h_host) {
            // overlap the device -> host copy of this batch's rows with the next batch's kernels
            CB2_CUDA(cudaEventRecord(sc->copy_ev, st));
            CB2_CUDA(cudaStreamWaitEvent(sc->copy_stream, sc->copy_ev, 0));
            if (sc->copy_stream2) CB2_CUDA(cudaStreamWaitEvent(sc->copy_stream2, sc->copy_ev, 0));
            pend_dst = (char*)sc->d2h_host + (size_t)r0 * S.bins * esz;
            pend_src = o;
            pend_bytes = (size_t)sub.n_rays * S.bins * esz;
            pend_r0 = r0; pend_n = sub.n_rays;
        }
        r0 += sub.n_rays;
    }
    if (pend_bytes && (rc = flush_pending()) != CB2_OK) return rc;
    return CB2_OK;
}
