// oc_vels.h -- phi -> unit velocity at one node (optimals.py:168-186), shared by the HJB epilogue kernels
// (oc_hjb.cu) and the GCFM sampler when the field is stored as phi slices (oc_gcfm.cu).  Written so that FMA
// contraction cannot change a bit (explicit fma, no a*b+c patterns): both translation units produce
// identical values, whichever compiler flags they use.
#pragma once

__device__ __forceinline__ double clamp_lim(double p, double lim) {
    // optimals.py:172: phi*(phi > lim) + lim*(phi < lim)  (== 0 when phi == lim exactly)
    return p > lim ? p : (p < lim ? lim : 0.0);
}

__device__ __forceinline__ void vels_point(double pc0, double pW, double pE, double pS, double pN, double mu,
                                           double lim, double inv_two_dx, double inv_two_dy, double &ox, double &oy) {
    // inputs already clamped once (optimals.py:172).  :174-186 compute v = grad/(mu*pc), n = |v| and return
    // v*(n>lim)/(n*(n>lim)+(n<lim)), i.e. the unit vector grad/|grad| where n > lim and 0 where n < lim.  The
    // common factor 1/(mu*pc) cancels in the direction, so it is only needed for the threshold test: one
    // division and one reciprocal instead of four divisions (results differ from the literal formula by
    // rounding only, <= 2 ulp).  n == lim exactly gives 0/0 = NaN in the reference; kept.
    double gx = (pE - pW) * inv_two_dx;
    double gy = (pN - pS) * inv_two_dy;
    double pc = clamp_lim(pc0, lim);  // :177 second clamp
    double ng = sqrt(fma(gx, gx, gy * gy));
    double nr = ng / (mu * pc);
    double inv = (nr > lim) ? 1.0 / ng : ((nr < lim) ? 0.0 : __longlong_as_double(0x7ff8000000000000LL));
    ox = gx * inv;
    oy = gy * inv;
}
