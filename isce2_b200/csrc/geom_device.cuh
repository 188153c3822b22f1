// geom_device.cuh -- FP64 device math for the zero-Doppler geometry kernels (sm_100a).
//
// Ellipsoid geodesy, 3-vector helpers, orbit interpolation and polynomial evaluation as
// __host__ __device__ inline functions so that the same source can be compiled by g++ for
// the CPU-side emulation harness under tests/emu (a development aid, never a product path).
//
// Arithmetic contract: the translation units that include this header are compiled with
// -fmad=false.  Sums and products round separately, exactly like the reference's x86-64
// Fortran/C build, which keeps the float32 DEM-index quantisation (topozero.f90:525-536)
// on the same side of every rounding boundary as the reference wherever libm agrees.
// Fused multiply-adds are used only where written explicitly (b2_fma), in places whose
// result is rounded to float32 anyway or that are insensitive.
//
// Behavioural citations are relative to the ISCE2 tree.
#pragma once

#include <math.h>

#ifdef __CUDACC__
#define B2_HD __host__ __device__ __forceinline__
#define B2_D __device__ __forceinline__
#else
#define B2_HD inline
#define B2_D inline
#endif

#ifdef __CUDA_ARCH__
#define B2_CBRT(x) cbrt(x)
#define b2_fma(a, b, c) __fma_rn((a), (b), (c))
#else
// host emulation follows the oracle literally: latlon.F:62 uses **(1/3) == pow
#define B2_CBRT(x) pow((x), 1.0 / 3.0)
#define b2_fma(a, b, c) ((a) * (b) + (c))
#endif

namespace b2 {

struct Vec3 {
    double x, y, z;
};

B2_HD Vec3 v3(double x, double y, double z) { return Vec3{x, y, z}; }
// components/isceobj/Util/Library/linalg3/src/linalg3Module.F:37-382
B2_HD double dot(const Vec3 &a, const Vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
B2_HD Vec3 cross(const Vec3 &u, const Vec3 &v)
{
    return Vec3{u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x};
}
B2_HD double norm(const Vec3 &v) { return sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }
B2_HD Vec3 unitvec(const Vec3 &v)
{
    double n = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    if (n != 0) return Vec3{v.x / n, v.y / n, v.z / n};
    return v;
}
B2_HD Vec3 sub(const Vec3 &a, const Vec3 &b) { return Vec3{a.x - b.x, a.y - b.y, a.z - b.z}; }
B2_HD Vec3 add(const Vec3 &a, const Vec3 &b) { return Vec3{a.x + b.x, a.y + b.y, a.z + b.z}; }

struct Ellipsoid {
    double a, e2;
};

// latlon.F:44-49 (LLH_2_XYZ), llh in radians
B2_HD Vec3 llh_to_xyz(const Ellipsoid &e, double lat, double lon, double h)
{
    double sl, cl, so, co;
#ifdef __CUDA_ARCH__
    sincos(lat, &sl, &cl);
    sincos(lon, &so, &co);
#else
    sl = sin(lat); cl = cos(lat); so = sin(lon); co = cos(lon);
#endif
    double re = e.a / sqrt(1.0 - e.e2 * (sl * sl));
    Vec3 v;
    v.x = (re + h) * cl * co;
    v.y = (re + h) * cl * so;
    v.z = (re * (1.0 - e.e2) + h) * sl;
    return v;
}

// latlon.F:51-71 (XYZ_2_LLH): closed form; returns lat, lon (rad) and height
B2_HD void xyz_to_llh(const Ellipsoid &e, const Vec3 &v, double &lat, double &lon, double &h)
{
    double q2 = (v.x * v.x + v.y * v.y);
    double q3 = e.a * e.a;
    double e4 = e.e2 * e.e2;
    double p = q2 / q3;
    double q = (1.0 - e.e2) * (v.z * v.z) / q3;
    double r = (p + q - e4) / 6.0;
    double s = (e4 * p * q) / (4.0 * (r * r * r));
    double t = B2_CBRT(1.0 + s + sqrt(s * (2.0 + s)));
    double u = r * (1.0 + t + 1.0 / t);
    double rv = sqrt(u * u + e4 * q);
    double w = e.e2 * (u + rv - q) / (2.0 * rv);
    double k = sqrt(u + rv + w * w) - w;
    double d = k * sqrt(q2) / (k + e.e2);
    lat = atan2(v.z, d);
    lon = atan2(v.y, v.x);
    h = (k + e.e2 - 1.0) * sqrt(d * d + v.z * v.z) / k;
}

// curvature.F:26-64
B2_HD double reast(const Ellipsoid &e, double lat)
{
    double s = sin(lat);
    return e.a / sqrt(1.0 - e.e2 * (s * s));
}
B2_HD double rnorth(const Ellipsoid &e, double lat)
{
    double s = sin(lat);
    return (e.a * (1.0 - e.e2)) / pow(1.0 - e.e2 * (s * s), 1.5);
}
B2_HD double rdir(const Ellipsoid &e, double hdg, double lat)
{
    double re = reast(e, lat), rn = rnorth(e, lat);
    double c = cos(hdg), s = sin(hdg);
    return (re * rn) / (re * (c * c) + rn * (s * s));
}

// Per-azimuth-line geometry: everything topozero.f90:371-424 derives from the line's state vector.
struct LineState {
    Vec3 sat, vel, vhat, that, chat, nhat;
    double vmag, height, rcurv;
    double lat_sat, lon_sat;
    double minv[9]; // ptm%r_matinv, row-major
    Vec3 ov;
    double nv; // dot(nhat, vhat)
    double vt; // dot(vhat, that)
};

// tcnbasis.F:26-39 + radar_to_xyz.F:49-92 + topozero.f90:381-424
B2_HD void make_line_state(const Ellipsoid &e, const Vec3 &pos, const Vec3 &vel, double peghdg, LineState &L)
{
    L.sat = pos;
    L.vel = vel;
    L.vhat = unitvec(vel);
    L.vmag = norm(vel);
    double lat, lon, h;
    xyz_to_llh(e, pos, lat, lon, h);
    L.lat_sat = lat;
    L.lon_sat = lon;
    L.height = h;
    // tcnbasis (calls latlon again on the same input -> same lat/lon)
    double clt = cos(lat), slt = sin(lat), clo = cos(lon), slo = sin(lon);
    L.nhat = Vec3{-clt * clo, -clt * slo, -slt};
    L.chat = unitvec(cross(L.nhat, vel));
    L.that = unitvec(cross(L.chat, L.nhat));
    // radar_to_xyz with peg = (lat, lon, peghdg)
    double chg = cos(peghdg), shg = sin(peghdg);
    double m[9];
    m[0] = clt * clo;
    m[1] = -shg * slo - slt * clo * chg;
    m[2] = slo * chg - slt * clo * shg;
    m[3] = clt * slo;
    m[4] = clo * shg - slt * slo * chg;
    m[5] = -clo * chg - slt * slo * shg;
    m[6] = slt;
    m[7] = clt * chg;
    m[8] = clt * shg;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) L.minv[3 * i + j] = m[3 * j + i];
    L.rcurv = rdir(e, peghdg, lat);
    Vec3 p = llh_to_xyz(e, lat, lon, 0.0);
    L.ov = Vec3{p.x - L.rcurv * (clt * clo), p.y - L.rcurv * (clt * slo), p.z - L.rcurv * slt};
    L.nv = dot(L.nhat, L.vhat);
    L.vt = dot(L.vhat, L.that);
}

// convert_sch_to_xyz.F:63-72 (XYZ_2_SCH), height component only.  The spherical latlon call
// (a = rcurv, e2 = 0) degenerates to the few operations below; they are kept in the reference's
// order so that the SCH height matches the CPU path bit for bit.
B2_HD double sch_height(const LineState &L, const Vec3 &xyz)
{
    double tx = 1.0 * xyz.x + (-1.0) * L.ov.x, ty = 1.0 * xyz.y + (-1.0) * L.ov.y, tz = 1.0 * xyz.z + (-1.0) * L.ov.z;
    double sx = L.minv[0] * tx + L.minv[1] * ty + L.minv[2] * tz;
    double sy = L.minv[3] * tx + L.minv[4] * ty + L.minv[5] * tz;
    double sz = L.minv[6] * tx + L.minv[7] * ty + L.minv[8] * tz;
    double q2 = (sx * sx + sy * sy);
    double q3 = L.rcurv * L.rcurv;
    double p = q2 / q3;
    double q = (1.0 - 0.0) * (sz * sz) / q3;
    double r = (p + q - 0.0) / 6.0;
    // e2 = 0: s = 0, t = 1, u = 3 r, rv = sqrt(u*u) = u, w = 0, k = sqrt(2 u)
    double u = r * (1.0 + 1.0 + 1.0 / 1.0);
    double rv = sqrt(u * u);
    double k = sqrt(u + rv);
    double d = k * sqrt(q2) / k;
    return (k - 1.0) * sqrt(d * d + sz * sz) / k;
}

// ---------------------------------------------------------------------------------------------
// orbit (components/isceobj/Util/Library/orbit/src/orbit.c, orbitHermite.c)
// ---------------------------------------------------------------------------------------------
struct OrbitView {
    int n;
    const double *t;   // [n]
    const double *pos; // [n][3]
    const double *vel; // [n][3]
};

// orbitHermite.c:4-94 on the window chosen by orbit.c:203-211; returns the reference's stat
B2_HD int orbit_hermite(const OrbitView &o, double time, Vec3 &xx, Vec3 &vv)
{
    int i;
    for (i = 0; i < o.n; i++)
        if (o.t[i] >= time) break;
    i -= 2;
    if (i < 0) i = 0;
    if (i > o.n - 4) i = o.n - 4;
    double t[4], h[4], hdot[4], f0[4], f1[4], g0[4], g1[4];
    for (int j = 0; j < 4; j++) t[j] = o.t[i + j];
    for (int a = 0; a < 4; ++a) {
        f1[a] = time - t[a];
        double sum = 0.0;
        for (int j = 0; j < 4; ++j)
            if (a != j) sum += 1.0 / (t[a] - t[j]);
        f0[a] = 1.0 - 2.0 * (time - t[a]) * sum;
    }
    for (int a = 0; a < 4; ++a) {
        double product = 1.0;
        for (int k = 0; k < 4; ++k)
            if (k != a) product *= (time - t[k]) / (t[a] - t[k]);
        h[a] = product;
        double sum = 0.0;
        for (int j = 0; j < 4; ++j) {
            product = 1.0;
            for (int k = 0; k < 4; ++k)
                if ((k != a) && (k != j)) product *= (time - t[k]) / (t[a] - t[k]);
            if (j != a) sum += 1.0 / (t[a] - t[j]) * product;
        }
        hdot[a] = sum;
    }
    for (int a = 0; a < 4; ++a) {
        g1[a] = h[a] + 2.0 * (time - t[a]) * hdot[a];
        double sum = 0.0;
        for (int j = 0; j < 4; ++j)
            if (a != j) sum += 1.0 / (t[a] - t[j]);
        g0[a] = 2.0 * (f0[a] * hdot[a] - h[a] * sum);
    }
    double xo[3], vo[3];
    for (int k = 0; k < 3; ++k) {
        double sum = 0.0;
        for (int a = 0; a < 4; ++a) sum += (o.pos[3 * (i + a) + k] * f0[a] + o.vel[3 * (i + a) + k] * f1[a]) * h[a] * h[a];
        xo[k] = sum;
        sum = 0.0;
        for (int a = 0; a < 4; ++a) sum += (o.pos[3 * (i + a) + k] * g0[a] + o.vel[3 * (i + a) + k] * g1[a]) * h[a];
        vo[k] = sum;
    }
    xx = Vec3{xo[0], xo[1], xo[2]};
    vv = Vec3{vo[0], vo[1], vo[2]};
    return ((time < o.t[0]) || (time > o.t[o.n - 1])) ? 1 : 0;
}

// orbit.c:236-314
B2_HD int orbit_legendre(const OrbitView &o, double time, Vec3 &xx, Vec3 &vv)
{
    const double noemer[9] = {40320.0, -5040.0, 1440.0, -720.0, 576.0, -720.0, 1440.0, -5040.0, 40320.0};
    int i;
    for (i = 0; i < o.n; i++)
        if (o.t[i] >= time) break;
    i -= 5;
    if (i < 0) i = 0;
    if (i > o.n - 9) i = o.n - 9;
    double trel = 8.0 * (time - o.t[i]) / (o.t[i + 8] - o.t[i]);
    double teller = 1.0;
    for (int j = 0; j < 9; j++) teller *= (trel - j);
    double xo[3] = {0.0, 0.0, 0.0}, vo[3] = {0.0, 0.0, 0.0};
    if (teller == 0.0) {
        int k = (int)trel;
        for (int j = 0; j < 3; j++) {
            xo[j] = o.pos[3 * (i + k) + j];
            vo[j] = o.vel[3 * (i + k) + j];
        }
    } else {
        for (int k = 0; k < 9; k++) {
            double coeff = teller / noemer[k] / (trel - k);
            for (int j = 0; j < 3; j++) {
                xo[j] += coeff * o.pos[3 * (i + k) + j];
                vo[j] += coeff * o.vel[3 * (i + k) + j];
            }
        }
    }
    xx = Vec3{xo[0], xo[1], xo[2]};
    vv = Vec3{vo[0], vo[1], vo[2]};
    return ((time < o.t[0]) || (time > o.t[o.n - 1])) ? 1 : 0;
}

// orbit.c:119-172 (outputs untouched when stat != 0)
B2_HD int orbit_sch(const OrbitView &o, double time, Vec3 &xx, Vec3 &vv)
{
    if ((time < o.t[0]) || (time > o.t[o.n - 1])) return 1;
    double xo[3] = {0.0, 0.0, 0.0}, vo[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < o.n; i++) {
        double frac = 1.0;
        double t0 = o.t[i];
        for (int j = 0; j < o.n; j++) {
            if (i == j) continue;
            double t1 = o.t[j];
            double num = t1 - time;
            double den = t1 - t0;
            frac *= num / den;
        }
        for (int k = 0; k < 3; k++) {
            xo[k] += frac * o.pos[3 * i + k];
            vo[k] += frac * o.vel[3 * i + k];
        }
    }
    xx = Vec3{xo[0], xo[1], xo[2]};
    vv = Vec3{vo[0], vo[1], vo[2]};
    return 0;
}

B2_HD int orbit_interp(int method, const OrbitView &o, double time, Vec3 &xx, Vec3 &vv)
{
    if (method == 0) return orbit_hermite(o, time, xx, vv);
    if (method == 1) return orbit_sch(o, time, xx, vv);
    return orbit_legendre(o, time, xx, vv);
}

// ---------------------------------------------------------------------------------------------
// polynomials
// ---------------------------------------------------------------------------------------------
constexpr int kMaxPoly2dCoeffs = 64;
constexpr int kMaxPoly1dCoeffs = 32;

struct Poly2dDev {
    int range_order, azimuth_order;
    double mean_range, mean_azimuth, norm_range, norm_azimuth;
    double c[kMaxPoly2dCoeffs];
};

struct Poly1dDev {
    int order;
    double mean, norm;
    double c[kMaxPoly1dCoeffs];
};

// poly2d.c:92-111 (accumulation order kept: value += scalex*scaley*c)
B2_HD double eval_poly2d(const Poly2dDev &p, double azi, double rng)
{
    double value = 0.0;
    double xval = (rng - p.mean_range) / p.norm_range;
    double yval = (azi - p.mean_azimuth) / p.norm_azimuth;
    double scaley = 1.0;
    for (int i = 0; i <= p.azimuth_order; i++, scaley *= yval) {
        double scalex = 1.0;
        for (int j = 0; j <= p.range_order; j++, scalex *= xval) value += scalex * scaley * p.c[i * (p.range_order + 1) + j];
    }
    return value;
}

// poly1d.c:87-104
B2_HD double eval_poly1d(const Poly1dDev &p, double x)
{
    double value = 0.0, scalex = 1.0;
    double xval = (x - p.mean) / p.norm;
    for (int i = 0; i <= p.order; i++, scalex *= xval) value += scalex * p.c[i];
    return value;
}

} // namespace b2
