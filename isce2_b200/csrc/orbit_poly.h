// orbit_poly.h -- host-side construction of per-window orbit polynomials for the geo2rdr Newton solve.
//
// The reference re-interpolates the orbit at every Newton step with orbitHermite (4 state vectors, position and
// velocity, components/isceobj/Util/Library/orbit/src/orbitHermite.c:4-94: ~460 operations with ~60 divisions) or
// the 9-point Lagrange formula (orbit.c:236-314).  Both are, for a fixed window of state vectors, plain polynomials
// of time: degree 7 for the Hermite position (its velocity output is exactly the time derivative of that
// polynomial), degree 8 for the Lagrange position and, independently, velocity.  Here those polynomials are expanded
// once per window in extended precision (polynomial algebra on the very formulas of the reference), so that the
// device evaluates a state vector with a Horner recurrence (~40-80 FMAs, no division).
#pragma once

#include <cmath>
#include <vector>

namespace b2 {

struct HostOrbitPoly {
    int method = 0; // 0 Hermite, 2 Legendre
    int n = 0, nwin = 0, ncoef = 0;
    std::vector<double> tc, inv_h; // [nwin] window centre and 1/scale: s = (t - tc) * inv_h
    std::vector<double> cp, cv;    // [nwin][3][ncoef], highest power first; cv empty for Hermite
};

namespace detail {
typedef long double ld;
typedef std::vector<ld> Poly; // ascending powers

inline Poly pmul(const Poly &a, const Poly &b)
{
    Poly r(a.size() + b.size() - 1, 0.0L);
    for (size_t i = 0; i < a.size(); i++)
        for (size_t j = 0; j < b.size(); j++) r[i + j] += a[i] * b[j];
    return r;
}
inline void padd(Poly &acc, const Poly &a, ld scale)
{
    if (acc.size() < a.size()) acc.resize(a.size(), 0.0L);
    for (size_t i = 0; i < a.size(); i++) acc[i] += scale * a[i];
}
} // namespace detail

inline bool build_orbit_poly(int method, int n, const double *t, const double *pos, const double *vel, HostOrbitPoly &out)
{
    using namespace detail;
    out.method = method;
    out.n = n;
    if (method == 0) { // Hermite: windows of 4 state vectors (orbit.c:203-211)
        if (n < 4) return false;
        out.nwin = n - 3;
        out.ncoef = 8;
        out.tc.resize(out.nwin);
        out.inv_h.resize(out.nwin);
        out.cp.assign((size_t)out.nwin * 3 * 8, 0.0);
        out.cv.clear();
        for (int w = 0; w < out.nwin; w++) {
            ld tt[4];
            for (int i = 0; i < 4; i++) tt[i] = t[w + i];
            const ld D = (tt[3] - tt[0]) / 3.0L;
            const ld tcw = 0.5L * (tt[1] + tt[2]);
            out.tc[w] = (double)tcw;
            const ld tcd = (ld)out.tc[w]; // the device uses the rounded centre
            out.inv_h[w] = (double)(1.0L / D);
            const ld Dd = 1.0L / (ld)out.inv_h[w]; // ... and the rounded scale
            ld sn[4];
            for (int i = 0; i < 4; i++) sn[i] = (tt[i] - tcd) / Dd;
            for (int c = 0; c < 3; c++) {
                Poly acc(8, 0.0L);
                for (int i = 0; i < 4; i++) {
                    Poly h{1.0L};
                    ld S = 0.0L;
                    for (int k = 0; k < 4; k++) {
                        if (k == i) continue;
                        const ld den = sn[i] - sn[k];
                        h = pmul(h, Poly{-sn[k] / den, 1.0L / den});
                        S += 1.0L / den;
                    }
                    const Poly h2 = pmul(h, h);
                    // x_i * f0_i + v_i * f1_i with f0 = 1 - 2 (s - s_i) S, f1 = Dd (s - s_i)
                    const ld x = pos[3 * (w + i) + c], v = vel[3 * (w + i) + c];
                    Poly lin{x * (1.0L + 2.0L * sn[i] * S) - v * Dd * sn[i], -2.0L * x * S + v * Dd};
                    padd(acc, pmul(lin, h2), 1.0L);
                }
                for (int k = 0; k < 8; k++) out.cp[((size_t)w * 3 + c) * 8 + k] = (double)acc[7 - k];
            }
        }
        return true;
    }
    if (method == 2) { // Legendre: windows of 9 state vectors (orbit.c:260-267), uniform-grid Lagrange in trel
        if (n < 9) return false;
        out.nwin = n - 8;
        out.ncoef = 9;
        out.tc.resize(out.nwin);
        out.inv_h.resize(out.nwin);
        out.cp.assign((size_t)out.nwin * 3 * 9, 0.0);
        out.cv.assign((size_t)out.nwin * 3 * 9, 0.0);
        // basis polynomials in s = trel - 4 are the same for every window
        std::vector<Poly> L(9);
        for (int i = 0; i < 9; i++) {
            Poly b{1.0L};
            for (int j = 0; j < 9; j++) {
                if (j == i) continue;
                const ld den = (ld)(i - j);
                b = pmul(b, Poly{(4.0L - j) / den, 1.0L / den});
            }
            L[i] = b;
        }
        for (int w = 0; w < out.nwin; w++) {
            const ld t0 = t[w], t8 = t[w + 8];
            out.tc[w] = (double)(t0 + 0.5L * (t8 - t0));
            out.inv_h[w] = (double)(8.0L / (t8 - t0));
            for (int c = 0; c < 3; c++) {
                Poly ap(9, 0.0L), av(9, 0.0L);
                for (int i = 0; i < 9; i++) {
                    padd(ap, L[i], (ld)pos[3 * (w + i) + c]);
                    padd(av, L[i], (ld)vel[3 * (w + i) + c]);
                }
                for (int k = 0; k < 9; k++) {
                    out.cp[((size_t)w * 3 + c) * 9 + k] = (double)ap[8 - k];
                    out.cv[((size_t)w * 3 + c) * 9 + k] = (double)av[8 - k];
                }
            }
        }
        return true;
    }
    return false;
}

} // namespace b2
