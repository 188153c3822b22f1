// orbit_poly_device.cuh -- device-side evaluation of the per-window orbit polynomials built by orbit_poly.h
// (position, velocity and acceleration of the Hermite / 9-point Lagrange orbit interpolants by Horner recurrences),
// shared by the geo2rdr and geozero Newton solves.
#pragma once

#include "geo2rdr_kernels.cuh"

namespace b2 {

struct OrbState {
    Vec3 x, v, a;
};

// Window of the epoch `time`: the first i with t[i] >= time (orbit.c:203-206; epochs ascending), found by walking from
// `hint` -- the answer of the caller's previous query.  Successive queries of one thread are the Newton iterates of one
// pixel and then of its neighbours on the same line, a fraction of a state-vector interval apart, so the walk is two loads
// instead of a scan from the first state vector (which was a quarter of the geo2rdr kernel's stall samples).
template <int METHOD>
__device__ __forceinline__ int poly_window(const OrbitPolyView &op, double time, int &hint)
{
    const double *__restrict__ t = op.t;
    const int n = op.n; // >= the window span (4 / 9): checked when the polynomials are built
    int i = hint < 1 ? 1 : (hint > n - 1 ? n - 1 : hint);
    const double below = __ldg(t + i - 1), at = __ldg(t + i); // both in flight at once; usually this is the window already
    if (!(below < time && at >= time)) {
        if (below >= time) {
            i--;
            while (i > 0 && __ldg(t + i - 1) >= time) i--;
        } else {
            i++;
            while (i < n && __ldg(t + i) < time) i++;
        }
    }
    hint = i;
    const int back = (METHOD == 0) ? 2 : 5, span = (METHOD == 0) ? 4 : 9;
    int w = i - back;
    w = w < 0 ? 0 : w;
    w = w > op.n - span ? op.n - span : w;
    return w;
}

template <int METHOD>
__device__ __forceinline__ void poly_state(const OrbitPolyView &op, double time, OrbState &S, int &hint)
{
    const int w = poly_window<METHOD>(op, time, hint);
    const double ih = __ldg(op.inv_h + w);
    const double s = (time - __ldg(op.tc + w)) * ih;
    constexpr int NC = (METHOD == 0) ? 8 : 9;
    double xo[3], vo[3], ao[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double *cp = op.cp + ((size_t)w * 3 + c) * NC;
        if (METHOD == 0) { // position, first and second derivative from one coefficient set
            double p = __ldg(cp), dp = 0.0, ddp = 0.0;
#pragma unroll
            for (int k = 1; k < NC; k++) {
                ddp = __fma_rn(ddp, s, dp);
                dp = __fma_rn(dp, s, p);
                p = __fma_rn(p, s, __ldg(cp + k));
            }
            xo[c] = p;
            vo[c] = dp * ih;
            ao[c] = 2.0 * ddp * ih * ih;
        } else { // Legendre: velocity is its own polynomial; acceleration = d/dt of it
            const double *cv = op.cv + ((size_t)w * 3 + c) * NC;
            double p = __ldg(cp), q = __ldg(cv), dq = 0.0;
#pragma unroll
            for (int k = 1; k < NC; k++) {
                dq = __fma_rn(dq, s, q);
                q = __fma_rn(q, s, __ldg(cv + k));
                p = __fma_rn(p, s, __ldg(cp + k));
            }
            xo[c] = p;
            vo[c] = q;
            ao[c] = dq * ih;
        }
    }
    S.x = Vec3{xo[0], xo[1], xo[2]};
    S.v = Vec3{vo[0], vo[1], vo[2]};
    S.a = Vec3{ao[0], ao[1], ao[2]};
}

// Horner evaluation of a polynomial of the compile-time degree N: coefficients read as constant-bank operands of the FMAs
template <int N>
__device__ __forceinline__ double horner_fixed(const Poly1dDev &p, double xv)
{
    double v = p.c[N];
#pragma unroll
    for (int i = N - 1; i >= 0; i--) v = __fma_rn(v, xv, p.c[i]);
    return v;
}

// Doppler polynomials have a handful of coefficients; the degree is uniform over the grid, so a switch on it costs a
// uniform branch and replaces the loop with its register-indexed constant loads (14 % of the stall samples of the
// native-Doppler geo2rdr kernel sat on that loop)
__device__ __forceinline__ double poly1d_fast(const Poly1dDev &p, double inv_norm, double x)
{
    if (p.order == 0) return p.c[0];
    const double xv = (x - p.mean) * inv_norm;
    switch (p.order) {
    case 1: return horner_fixed<1>(p, xv);
    case 2: return horner_fixed<2>(p, xv);
    case 3: return horner_fixed<3>(p, xv);
    case 4: return horner_fixed<4>(p, xv);
    case 5: return horner_fixed<5>(p, xv);
    case 6: return horner_fixed<6>(p, xv);
    default: break;
    }
    double v = p.c[p.order];
    for (int i = p.order - 1; i >= 0; i--) v = __fma_rn(v, xv, p.c[i]);
    return v;
}

} // namespace b2
