// common.cuh -- device helpers shared by the chain and box kernels (sm_100a).
//
// Geometry and potentials restate, in restructured form, the reference arithmetic:
//   nearest image   src/utils.jl:15-28   (dx - round(dx/L)*L, sum of squares)
//   lennard_jones   src/models.jl:30-34  inverse_power :28   fene :36
//   potential(...)  src/models.jl:72-74, :121-123, :160-166, :207-209; bond_potential :219-226
// Positions are kept WRAPPED into [0, L] on the device (plus an integer image counter), so the
// nearest image per axis is min(|d|, L - |d|) -- the same value the reference's rint() form gives
// for wrapped inputs, without the fp64 divide and round (both multi-instruction on the GPU).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "pmc_b200.h"

namespace pmc {

// 1/x to <= 1 ulp: MUFU.RCP64H seed (~20 bits) + two Newton steps on the DFMA pipe.  The IEEE
// division nvcc emits for 1.0/x costs about twice as many issue slots plus a slow-path branch.
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}

__device__ __forceinline__ double lj_core(double r2, double eps4, double sig2) {
    double x = sig2 * fast_rcp(r2);
    double x3 = x * x * x;
    return eps4 * (x3 * x3 - x3);
}

// Non-bonded pair potential of model kind MODEL at squared distance r2 (cutoff test done by caller).
template <int MODEL>
__device__ __forceinline__ double pair_potential(const double *__restrict__ p, double r2) {
    if constexpr (MODEL == PMC_MODEL_LJ || MODEL == PMC_MODEL_KG) {
        return lj_core(r2, p[PMC_P_EPS], p[PMC_P_SIG2]) - p[PMC_P_SHIFT];
    } else if constexpr (MODEL == PMC_MODEL_SMOOTHLJ) {
        double lj = lj_core(r2, p[PMC_P_EPS], p[PMC_P_SIG2]);
        return lj + p[PMC_P_EPS] * (p[PMC_P_C0] + r2 * fma(r2, p[PMC_P_C4S4], p[PMC_P_C2S2]));
    } else {  // PMC_MODEL_SOFT: eps * (sig2/r2)^ndiv2 - shift, integer power when ndiv2 is integral
        double x = p[PMC_P_SIG2] * fast_rcp(r2);
        double nd = p[PMC_P_NDIV2];
        int n = (int)nd;
        double v;
        if ((double)n == nd) {
            v = 1.0;
            double b = x;
            while (n > 0) {
                if (n & 1) v *= b;
                b *= b;
                n >>= 1;
            }
        } else {
            v = pow(x, nd);
        }
        return p[PMC_P_EPS] * v - p[PMC_P_SHIFT];
    }
}

// GeneralKG bonded term: FENE (Inf beyond r0) + shifted LJ inside rcutbond (src/models.jl:219-226).
__device__ __forceinline__ double bond_potential(const double *__restrict__ p, double r2) {
    double u_fene = (r2 <= p[PMC_P_R02]) ? p[PMC_P_KR02] * log(1.0 - r2 * fast_rcp(p[PMC_P_R02])) : CUDART_INF;
    double u_lj = 0.0;
    if (r2 <= p[PMC_P_RCUT2B]) u_lj = lj_core(r2, p[PMC_P_EPS4B], p[PMC_P_SIG2B]) - p[PMC_P_SHIFTB];
    return u_fene + u_lj;
}

// Nearest-image squared separation along one axis for coordinates wrapped into [0, L].
__device__ __forceinline__ double mi_sq(double xi, double xj, double L) {
    double a = fabs(xi - xj);
    double r = fmin(a, L - a);
    return r * r;
}
__device__ __forceinline__ float mi_sq(float xi, float xj, float L) {
    float a = fabsf(xi - xj);
    float r = fminf(a, L - a);
    return r * r;
}

// Wrap x (already within one box length of [0, L)) back into the box; w receives the image shift.
__device__ __forceinline__ double wrap1(double x, double L, int &w) {
    w = 0;
    if (x >= L) {
        x -= L;
        w = 1;
    } else if (x < 0.0) {
        x += L;
        w = -1;
    }
    return x;
}

// Reference acceptance arithmetic: min(1, exp(-(e2 - e1) / T)) > u, with Julia's NaN-propagating min
// (a NaN energy difference is a rejection).
__device__ __forceinline__ bool accept_exact(double dE, double T, double u) {
    const double ex = exp(-dE / T);
    return !(ex != ex) && (fmin(1.0, ex) > u);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- 8-bit prefilter ---------------------------------------------------------------------------------------
// A candidate is three bytes (top 8 bits of the fixed-point fraction of L per axis) in one register.  Per byte
// VABSDIFF4.U8 gives |a - b| in [0, 255]; read as a SIGNED byte by IDP.4A that is the wrapped difference
// (d >= 128 -> d - 256), so one IDP.4A.S8.S8 returns the minimum-image squared distance in units of (L/256)^2,
// and its accumulator input subtracts the threshold: the sign bit is the verdict.  Three instructions per candidate
// (VABSDIFF4, IDP.4A, SHF funnel) instead of 3 IADD + 3 IMAD.HI + compare + select on 32-bit coordinates.
// Quantisation: both bytes are floors of exact scaled coordinates, so each component differs from the true one
// by less than 1 unit and sum q_a^2 <= (r + sqrt(3))^2 -- the threshold carries that margin, so no in-range
// candidate is ever dropped (survivors are then treated exactly, in fp64).
__device__ __forceinline__ uint32_t pack8(uint32_t u0, uint32_t u1, uint32_t u2) {
    return __byte_perm(__byte_perm(u0, u1, 0x4473), u2 >> 24, 0x5410);  // bytes: u0>>24, u1>>24, u2>>24, 0
}
// ~threshold (== -(thr + 1)): IDP.4A(t, t, ~thr) < 0  <=>  r2 <= thr
__device__ __forceinline__ uint32_t neg_thr8(double r_units) {
    const double t = r_units + 1.7320526;
    const double t2 = t * t + 1.0;
    return ~(t2 >= 60000.0 ? 60000u : (uint32_t)t2);
}

}  // namespace pmc
