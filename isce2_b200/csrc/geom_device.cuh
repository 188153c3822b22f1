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
#define b2_fma(a, b, c) fma((a), (b), (c))
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

// a / b for a divisor whose correctly rounded reciprocal rb = RN(1/b) is known (hoisted per scene, per line or per
// pixel).  q0 = RN(a*rb) is within one ulp of the quotient, the residual a - q0*b is exact in an FMA, and the
// corrected quotient rounds correctly (Markstein 1990): the result is the IEEE quotient at 3 FP64 instructions
// instead of the ~30 of a full division, so the reference's divisions can be kept bit for bit.
B2_HD double div_r(double a, double b, double rb)
{
    double q = a * rb;
    double r = b2_fma(-q, b, a);
    return b2_fma(r, rb, q);
}

// ---------------------------------------------------------------------------------------------
// IEEE-exact division and square root in ~8-11 FP64 instructions.
//
// CUDA's own double-precision `/` and sqrt() expand to ~30 / ~20 executed instructions (special-case handling
// included) and drag a slow path into the instruction stream.  For normal, finite operands the sequences below
// give the same correctly rounded result: a hardware seed (MUFU.RCP64H / MUFU.RSQ64H, ~20 bits), two Newton
// steps to ~1 ulp, then one residual correction whose residual is exact in an FMA; the corrected value differs from
// the infinitely precise one by O(2^-100) before the final rounding, so it can only miss the correctly rounded
// result when the exact value lies within ~1e-30 (relative) of a rounding boundary.  The host emulation uses the
// IEEE operators themselves.
// ---------------------------------------------------------------------------------------------
B2_HD double rcp_n(double b) // ~1 ulp reciprocal
{
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    y = __fma_rn(__fma_rn(-b, y, 1.0), y, y);
    y = __fma_rn(__fma_rn(-b, y, 1.0), y, y);
    return y;
#else
    return 1.0 / b;
#endif
}

B2_HD double div_n(double a, double b) // a / b, correctly rounded (see above)
{
#ifdef __CUDA_ARCH__
    return div_r(a, b, rcp_n(b));
#else
    return a / b;
#endif
}

// sqrt(x) correctly rounded (see above) and, optionally, ~1 ulp 1/sqrt(x); x must be a positive normal number or 0
B2_HD double sqrt_n(double x, double *rs = nullptr)
{
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = __fma_rn(-h, g, 0.5);
    g = __fma_rn(g, r, g);
    h = __fma_rn(h, r, h);
    r = __fma_rn(-h, g, 0.5);
    g = __fma_rn(g, r, g);
    h = __fma_rn(h, r, h);
    double d = __fma_rn(-g, g, x);
    g = __fma_rn(d, h, g);
    if (rs) *rs = h + h;
    return x == 0.0 ? 0.0 : g;
#else
    double g = sqrt(x);
    if (rs) *rs = 1.0 / g;
    return g;
#endif
}

// same for an argument known to be a positive normal number (no zero guard: three instructions less)
B2_HD double sqrt_p(double x, double *rs = nullptr)
{
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = __fma_rn(-h, g, 0.5);
    g = __fma_rn(g, r, g);
    h = __fma_rn(h, r, h);
    r = __fma_rn(-h, g, 0.5);
    g = __fma_rn(g, r, g);
    h = __fma_rn(h, r, h);
    double d = __fma_rn(-g, g, x);
    g = __fma_rn(d, h, g);
    if (rs) *rs = h + h;
    return g;
#else
    return sqrt_n(x, rs);
#endif
}

// ---------------------------------------------------------------------------------------------
// Trigonometry about a reference angle.
//
// Inside one block of azimuth lines every latitude (longitude) handled by the kernels lies within a few hundredths
// of a radian of a reference angle a0.  a0 is chosen on the host as an exact multiple of 2^-6 rad, and sin(a0),
// cos(a0) are carried as double-double (hi + lo, evaluated in extended precision on the host).  Then
//   atan2(y, x)  = a0 + atan(t),  t = (y*c0 - x*s0) / (x*c0 + y*s0)        (|t| small: short odd series)
//   sin(a0 + d)  = s0 + (c0*sin d - s0*(1 - cos d)),  cos likewise          (|d| small: short series)
// with the cancelling products formed exactly (FMA error terms).  The results carry ~0.5 ulp total error, i.e. they
// are as close to correctly rounded as glibc's own atan2/sin/cos, at a third of the instructions of the generic
// libm routines and without their slow paths.
// ---------------------------------------------------------------------------------------------
struct RefAngle {
    double a0;             // exact multiple of 2^-6 rad
    double sh, sl, ch, cl; // sin(a0) = sh + sl, cos(a0) = ch + cl
    double dmax;           // largest |angle - a0| the series below are good for
    int narrow;            // 1: every angle of the block is within kNarrowAngle of a0 -> the short series suffice
};

// Series coefficients live in constant memory so that the FP64 instructions read them as c[bank][offset] operands
// (as literals the compiler rebuilds each one with two UMOVs per use).
#define B2_ATAN_COEF {-1.0 / 3.0, 1.0 / 5.0, -1.0 / 7.0, 1.0 / 9.0, -1.0 / 11.0, 1.0 / 13.0, -1.0 / 15.0, 1.0 / 17.0, \
                      -1.0 / 19.0, 1.0 / 21.0, -1.0 / 23.0}
#define B2_SIN_COEF {-1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0, -1.0 / 39916800.0, 1.0 / 6227020800.0}
#define B2_COS_COEF {-0.5, 1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0, -1.0 / 3628800.0, 1.0 / 479001600.0, \
                     -1.0 / 87178291200.0}
#define B2_CUBE_COEF {1.0 / 9.0, -4.0 / 243.0, 28.0 / 6561.0, -80.0 / 59049.0, 2288.0 / 4782969.0, \
                      -23296.0 / 129140163.0, 82688.0 / 1162261467.0}
#ifdef __CUDACC__
__device__ __constant__ static double kAtanCoefDev[11] = B2_ATAN_COEF;
__device__ __constant__ static double kSinCoefDev[6] = B2_SIN_COEF;
__device__ __constant__ static double kCosCoefDev[7] = B2_COS_COEF;
__device__ __constant__ static double kCubeCoefDev[7] = B2_CUBE_COEF;
#endif
static const double kAtanCoefHost[11] = B2_ATAN_COEF;
static const double kSinCoefHost[6] = B2_SIN_COEF;
static const double kCosCoefHost[7] = B2_COS_COEF;
static const double kCubeCoefHost[7] = B2_CUBE_COEF;
#ifdef __CUDA_ARCH__
#define B2_ATAN_C(i) kAtanCoefDev[i]
#define B2_SIN_C(i) kSinCoefDev[i]
#define B2_COS_C(i) kCosCoefDev[i]
#define B2_CUBE_C(i) kCubeCoefDev[i]
#else
#define B2_ATAN_C(i) kAtanCoefHost[i]
#define B2_SIN_C(i) kSinCoefHost[i]
#define B2_COS_C(i) kCosCoefHost[i]
#define B2_CUBE_C(i) kCubeCoefHost[i]
#endif

// |angle - a0| below which the truncated series are used: the first dropped terms, t^13/13 = 2.4e-19,
// d^11/11! = 3.9e-23 and d^10/10! = 9.4e-21, are below 0.3 % of an ulp of the results (ulp(0.5) = 1.1e-16)
constexpr double kNarrowAngle = 0.045;

// host only: extended-precision constants of a reference angle near `approx`
inline RefAngle make_ref_angle(double approx)
{
    RefAngle R;
    R.a0 = nearbyint(approx * 64.0) / 64.0;
    long double s = sinl((long double)R.a0), c = cosl((long double)R.a0);
    R.sh = (double)s;
    R.sl = (double)(s - (long double)R.sh);
    R.ch = (double)c;
    R.cl = (double)(c - (long double)R.ch);
    R.dmax = 0.12;
    R.narrow = 0;
    return R;
}

// atan2(y, x) for a direction within R.dmax of R.a0
B2_HD double atan2_ref(const RefAngle &R, double y, double x)
{
    // N = y*c0 - x*s0 with both products exact (hi + fma error), D = x*c0 + y*s0
    double p1 = y * R.ch, e1 = b2_fma(y, R.ch, -p1);
    double p2 = x * R.sh, e2 = b2_fma(x, R.sh, -p2);
    double N = (p1 - p2) + ((e1 - e2) + (y * R.cl - x * R.sl));
    double D = b2_fma(x, R.ch, y * R.sh);
    double t = N * rcp_n(D);
    double t2 = t * t;
    // odd Taylor series of atan through t^23 (|t| <= 0.12: remainder < 4e-25), through t^11 on narrow blocks
    double pol = B2_ATAN_C(4);
    if (!R.narrow) {
        pol = B2_ATAN_C(10);
        pol = b2_fma(pol, t2, B2_ATAN_C(9));
        pol = b2_fma(pol, t2, B2_ATAN_C(8));
        pol = b2_fma(pol, t2, B2_ATAN_C(7));
        pol = b2_fma(pol, t2, B2_ATAN_C(6));
        pol = b2_fma(pol, t2, B2_ATAN_C(5));
        pol = b2_fma(pol, t2, B2_ATAN_C(4));
    }
    pol = b2_fma(pol, t2, B2_ATAN_C(3));
    pol = b2_fma(pol, t2, B2_ATAN_C(2));
    pol = b2_fma(pol, t2, B2_ATAN_C(1));
    pol = b2_fma(pol, t2, B2_ATAN_C(0));
    double at = b2_fma(t * t2, pol, t); // t + t^3 * pol
    return R.a0 + at;
}

// sin and cos of theta = R.a0 + d, |d| <= R.dmax
B2_HD void sincos_ref(const RefAngle &R, double theta, double &sn, double &cs)
{
    double d = theta - R.a0; // exact: both are within a factor two of each other or d is tiny
    double d2 = d * d;
    // S = sin d = d + d^3 * ps ; Cm = 1 - cos d = d^2 * pc   (|d| <= 0.12: remainders < 1e-24)
    // (narrow blocks: through d^9 and d^8)
    double ps = B2_SIN_C(3), pc = B2_COS_C(3);      // 1/9!, 1/8!
    if (!R.narrow) {
        ps = B2_SIN_C(5);                           // 1/13!
        ps = b2_fma(ps, d2, B2_SIN_C(4));           // -1/11!
        ps = b2_fma(ps, d2, B2_SIN_C(3));           // 1/9!
        pc = B2_COS_C(6);                           // -1/14!
        pc = b2_fma(pc, d2, B2_COS_C(5));           // 1/12!
        pc = b2_fma(pc, d2, B2_COS_C(4));           // -1/10!
        pc = b2_fma(pc, d2, B2_COS_C(3));           // 1/8!
    }
    ps = b2_fma(ps, d2, B2_SIN_C(2));               // -1/7!
    ps = b2_fma(ps, d2, B2_SIN_C(1));               // 1/5!
    ps = b2_fma(ps, d2, B2_SIN_C(0));               // -1/3!
    double S = b2_fma(d * d2, ps, d);
    pc = b2_fma(pc, d2, B2_COS_C(2));               // -1/6!
    pc = b2_fma(pc, d2, B2_COS_C(1));               // 1/4!
    pc = b2_fma(pc, d2, B2_COS_C(0));               // -1/2!
    double Cm = -(d2 * pc);                         // 1 - cos d  (>= 0)
    // sin(a0 + d) = s0 + [c0*S - s0*Cm],  cos(a0 + d) = c0 - [s0*S + c0*Cm]; low words of s0, c0 folded in
    sn = R.sh + (b2_fma(R.ch, S, -(R.sh * Cm)) + b2_fma(R.cl, S, R.sl));
    cs = R.ch + (-(b2_fma(R.sh, S, R.ch * Cm)) + b2_fma(-R.sl, S, R.cl));
}

struct Ellipsoid {
    double a, e2;
    // derived once (host): a^2 and its reciprocal, 1 - e2, e2^2, RN(1/6)
    double q3, inv_q3, ome2, e4, inv6;
};

// reference angles of one block of lines (see RefAngle); use_ref = 0 falls back to the generic libm routines
struct GeoRef {
    RefAngle lat, lon;
    int use_ref;
};

B2_HD Ellipsoid make_ellipsoid(double a, double e2)
{
    Ellipsoid e;
    e.a = a;
    e.e2 = e2;
    e.q3 = a * a;
    e.inv_q3 = 1.0 / e.q3;
    e.ome2 = 1.0 - e2;
    e.e4 = e2 * e2;
    e.inv6 = 1.0 / 6.0;
    return e;
}

// latlon.F:44-49 (LLH_2_XYZ) from the sines / cosines of the angles
B2_HD Vec3 llh_to_xyz_sc(const Ellipsoid &e, double sl, double cl, double so, double co, double h)
{
    double re = div_n(e.a, sqrt_p(1.0 - e.e2 * (sl * sl)));
    Vec3 v;
    v.x = (re + h) * cl * co;
    v.y = (re + h) * cl * so;
    v.z = (re * (1.0 - e.e2) + h) * sl;
    return v;
}

// latlon.F:44-49 (LLH_2_XYZ), llh in radians, generic libm trigonometry (setup kernels, fallback path)
B2_HD Vec3 llh_to_xyz(const Ellipsoid &e, double lat, double lon, double h)
{
    double sl, cl, so, co;
#ifdef __CUDA_ARCH__
    sincos(lat, &sl, &cl);
    sincos(lon, &so, &co);
#else
    sl = sin(lat); cl = cos(lat); so = sin(lon); co = cos(lon);
#endif
    return llh_to_xyz_sc(e, sl, cl, so, co, h);
}

// same, trigonometry about the block's reference angles
B2_HD Vec3 llh_to_xyz_ref(const Ellipsoid &e, const GeoRef &G, double lat, double lon, double h)
{
    double sl, cl, so, co;
    sincos_ref(G.lat, lat, sl, cl);
    sincos_ref(G.lon, lon, so, co);
    return llh_to_xyz_sc(e, sl, cl, so, co, h);
}

// latlon.F:51-71 (XYZ_2_LLH): closed form; returns k and d (lat = atan2(z, d)).
// Divisions by the scene constants a^2 and 6 use div_r, the others div_n, square roots sqrt_n: all IEEE-exact.
// The reference's  t = (1 + s + sqrt(s(2+s)))**(1/3);  u = r*(1 + t + 1/t)  is evaluated through
//   t + 1/t = 2 cosh(acosh(1+s)/3) = 2 (1 + e),   9e + 12e^2 + 4e^3 = s,
// whose reverted series in s converges fast because s <= 13.5 e2^2 ~ 6e-4 on an Earth-like ellipsoid (the generic
// cube-root form is kept for s >= 4e-3).  u feeds d only through k/(k+e2), which damps its rounding noise by
// ~e2/2 = 1/300, so this re-association is far below one ulp of the latitude.
B2_HD void xyz_to_llh_core(const Ellipsoid &e, const Vec3 &v, double &k, double &d)
{
    double q2 = (v.x * v.x + v.y * v.y);
    double p = div_r(q2, e.q3, e.inv_q3);
    double q = div_r(e.ome2 * (v.z * v.z), e.q3, e.inv_q3);
    double r = div_r(p + q - e.e4, 6.0, e.inv6);
    double s = div_n(e.e4 * p * q, 4.0 * (r * r * r));
    double u;
    if (s < 4.0e-3) {
        double ee = B2_CUBE_C(6);
        ee = b2_fma(ee, s, B2_CUBE_C(5));
        ee = b2_fma(ee, s, B2_CUBE_C(4));
        ee = b2_fma(ee, s, B2_CUBE_C(3));
        ee = b2_fma(ee, s, B2_CUBE_C(2));
        ee = b2_fma(ee, s, B2_CUBE_C(1));
        ee = b2_fma(ee, s, B2_CUBE_C(0));
        ee = ee * s;
        u = r * b2_fma(2.0, ee, 3.0);
    } else {
        double t = B2_CBRT(1.0 + s + sqrt(s * (2.0 + s)));
        u = r * (1.0 + t + 1.0 / t);
    }
    double rv = sqrt_p(u * u + e.e4 * q);
    double w = div_n(e.e2 * (u + rv - q), 2.0 * rv);
    k = sqrt_p(u + rv + w * w) - w;
    d = div_n(k * sqrt_p(q2), k + e.e2);
}

// lat, lon (rad) and height; generic libm arctangent
B2_HD void xyz_to_llh(const Ellipsoid &e, const Vec3 &v, double &lat, double &lon, double &h)
{
    double k, d;
    xyz_to_llh_core(e, v, k, d);
    lat = atan2(v.z, d);
    lon = atan2(v.y, v.x);
    h = div_n((k + e.e2 - 1.0) * sqrt_n(d * d + v.z * v.z), k);
}

// same about the block's reference angles; want_h = false skips the height (the iteration never uses it)
template <bool WANT_H>
B2_HD void xyz_to_llh_ref(const Ellipsoid &e, const GeoRef &G, const Vec3 &v, double &lat, double &lon, double &h)
{
    double k, d;
    xyz_to_llh_core(e, v, k, d);
    lat = atan2_ref(G.lat, v.z, d);
    lon = atan2_ref(G.lon, v.y, v.x);
    if (WANT_H) h = div_n((k + e.e2 - 1.0) * sqrt_n(d * d + v.z * v.z), k);
}

// lat, lon only with libm (fallback path of the iteration)
B2_HD void xyz_to_latlon(const Ellipsoid &e, const Vec3 &v, double &lat, double &lon)
{
    double k, d;
    xyz_to_llh_core(e, v, k, d);
    lat = atan2(v.z, d);
    lon = atan2(v.y, v.x);
}

// curvature.F:26-64
B2_HD double reast_s(const Ellipsoid &e, double s /* sin(lat) */) { return div_n(e.a, sqrt_n(1.0 - e.e2 * (s * s))); }
B2_HD double reast(const Ellipsoid &e, double lat) { return reast_s(e, sin(lat)); }
B2_HD double rnorth_s(const Ellipsoid &e, double s /* sin(lat) */)
{
    // (...)**1.5 of curvature.F:45 as x*sqrt(x): within an ulp of pow(), and only the float32 incidence layer sees it
    const double x = 1.0 - e.e2 * (s * s);
    return div_n(e.a * (1.0 - e.e2), x * sqrt_n(x));
}
B2_HD double rnorth(const Ellipsoid &e, double lat) { return rnorth_s(e, sin(lat)); }
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
    // hoisted reciprocals (IEEE, so that div_r reproduces the reference's divisions exactly)
    double aa, inv_aa; // height + rcurv
    double inv_vt, inv_vmag;
    double rc2, inv_rc2; // rcurv^2
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
    L.aa = L.height + L.rcurv;
    L.inv_aa = 1.0 / L.aa;
    L.inv_vt = 1.0 / L.vt;
    L.inv_vmag = 1.0 / L.vmag;
    L.rc2 = L.rcurv * L.rcurv;
    L.inv_rc2 = 1.0 / L.rc2;
}

// convert_sch_to_xyz.F:63-72 (XYZ_2_SCH), height component only.  The spherical latlon call
// (a = rcurv, e2 = 0) degenerates to the few operations below; they are kept in the reference's
// order so that the SCH height matches the CPU path bit for bit.
B2_HD double sch_height(const LineState &L, const Vec3 &xyz)
{
    double tx = xyz.x - L.ov.x, ty = xyz.y - L.ov.y, tz = xyz.z - L.ov.z; // lincomb(1, xyz, -1, ov)
    double sx = L.minv[0] * tx + L.minv[1] * ty + L.minv[2] * tz;
    double sy = L.minv[3] * tx + L.minv[4] * ty + L.minv[5] * tz;
    double sz = L.minv[6] * tx + L.minv[7] * ty + L.minv[8] * tz;
    double q2 = (sx * sx + sy * sy);
    double p = div_r(q2, L.rc2, L.inv_rc2);
    double q = div_r(sz * sz, L.rc2, L.inv_rc2); // (1 - 0) * sz^2 / a^2
    double r = div_r(p + q, 6.0, 1.0 / 6.0);     // (p + q - 0) / 6
    // e2 = 0: s = 0, t = 1, u = r * (1 + 1 + 1/1), rv = sqrt(u*u) == u (exact for u > 0), w = 0, k = sqrt(u + rv)
    double u = r * 3.0;
    double rk;
    double k = sqrt_p(u + u, &rk);
    double d = div_r(k * sqrt_p(q2), k, rk);
    return div_r((k - 1.0) * sqrt_p(d * d + sz * sz), k, rk);
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
    double inv_norm_range, inv_norm_azimuth; // IEEE reciprocals (host) for div_r
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
    double xval = div_r(rng - p.mean_range, p.norm_range, p.inv_norm_range);
    if (p.azimuth_order == 0) {
        // one row (slant range, range-only Doppler): scaley stays 1, and scalex * 1.0 * c == scalex * c, 1.0 * xval == xval
        // exactly, so this is the same sequence of roundings without the bookkeeping of the general loop
        value = value + p.c[0];
        double scalex = xval;
        for (int j = 1; j <= p.range_order; j++, scalex *= xval) value += scalex * p.c[j];
        return value;
    }
    double yval = div_r(azi - p.mean_azimuth, p.norm_azimuth, p.inv_norm_azimuth);
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
