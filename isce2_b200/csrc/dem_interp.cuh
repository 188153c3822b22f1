// dem_interp.cuh -- DEM interpolators of the topozero path on a float32, lon-fastest DEM crop.
//
// Same window checks and BADVALUE (-1000) semantics as the reference wrappers
// (components/zerodop/topozero/src/topozeroMethods.f:123-247).  ix/iy are the reference's
// 1-based integer indices, fx/fy the float32-derived fractions widened to double.
#pragma once

#include "geom_device.cuh"

namespace b2 {

struct DemView {
    const float *data; // [ny][nx], lon fastest == Fortran dem(nx, ny)
    int nx, ny;
    const float *sinc; // fintp table [8192][8] of the SINC interpolator (topozeroMethods.f:57-61), else NULL
    // BIQUINTIC on the device: the same samples widened to double (exact) with one replicated column and row
    // appended, [ny + 1][stride64]: the 6x6 window needs neither index clamps nor 36 float->double conversions
    const double *d64;
    int stride64;
};

#ifdef __CUDA_ARCH__
#define B2_LDG(p) __ldg(p)
#else
#define B2_LDG(p) (*(p))
#endif

B2_HD float dem_at(const DemView &d, int ix, int iy) // 1-based
{
    return B2_LDG(d.data + (size_t)(iy - 1) * (size_t)d.nx + (size_t)(ix - 1));
}

constexpr float kBadValue = -1000.0f; // topozeroMethods.f:33

// uniform_interp.f90:13-44 called as bilinear(dy, dx, dem): evaluated in double on float32 taps
B2_HD float interp_bilinear(const DemView &d, int i_x, int i_y, double f_x, double f_y)
{
    if ((i_x < 1) || (i_x >= d.nx)) return kBadValue;
    if ((i_y < 1) || (i_y >= d.ny)) return kBadValue;
    double x = i_y + f_y, y = i_x + f_x; // x: lat index, y: lon index (argument order of the reference call)
    double x1 = floor(x), x2 = ceil(x), y1 = ceil(y), y2 = floor(y);
    double q11 = dem_at(d, (int)y1, (int)x1);
    double q12 = dem_at(d, (int)y2, (int)x1);
    double q21 = dem_at(d, (int)y1, (int)x2);
    double q22 = dem_at(d, (int)y2, (int)x2);
    double r;
    if (y1 == y2 && x1 == x2) r = q11;
    else if (y1 == y2) r = (x2 - x) / (x2 - x1) * q11 + (x - x1) / (x2 - x1) * q21;
    else if (x1 == x2) r = (y2 - y) / (y2 - y1) * q11 + (y - y1) / (y2 - y1) * q12;
    else {
        // (x2-x1)*(y2-y1) == 1 * -1: dividing by -1 is an exact sign flip
        r = -(q11 * (x2 - x) * (y2 - y)) + -(q21 * (x - x1) * (y2 - y)) + -(q12 * (x2 - x) * (y - y1)) +
            -(q22 * (x - x1) * (y - y1));
    }
    return (float)r;
}

// SINC: 8x8 taps, 8192 sub-sample shifts, everything in float32 (topozeroMethods.f:100-121 -> uniform_interp.f90:
// 407-430 sinc_eval_2d_f): products and the running sum round to float32 in the reference's order (k outer, m inner),
// so the result is bit-identical.  The coefficient table is built on the host by sinc_make_table().
constexpr int kSincSub = 8192, kSincLen = 8;
B2_HD float interp_sinc(const DemView &d, int i_x, int i_y, double f_x, double f_y)
{
    if ((i_x < 4) || (i_x > (d.nx - 3))) return kBadValue;
    if ((i_y < 4) || (i_y > (d.ny - 3))) return kBadValue;
    const int intpx = i_x + kSincLen / 2, intpy = i_y + kSincLen / 2; // 0-based into the DEM
    float acc = 0.f;
    if ((intpx >= kSincLen - 1 && intpx < d.nx) && (intpy >= kSincLen - 1 && intpy < d.ny)) {
        int ifx = (int)(f_x * kSincSub), ify = (int)(f_y * kSincSub);
        ifx = ifx < 0 ? 0 : (ifx > kSincSub - 1 ? kSincSub - 1 : ifx);
        ify = ify < 0 ? 0 : (ify > kSincSub - 1 ? kSincSub - 1 : ify);
        const float *cx = d.sinc + (size_t)ifx * kSincLen, *cy = d.sinc + (size_t)ify * kSincLen;
        float wy[kSincLen];
#pragma unroll
        for (int m = 0; m < kSincLen; m++) wy[m] = B2_LDG(cy + m);
#pragma unroll 1
        for (int k = 0; k < kSincLen; k++) {
            const float wxk = B2_LDG(cx + k);
#pragma unroll
            for (int m = 0; m < kSincLen; m++) {
                float a = B2_LDG(d.data + (size_t)(intpy - m) * (size_t)d.nx + (size_t)(intpx - k));
                float t = a * wxk;
                t = t * wy[m];
                acc = acc + t;
            }
        }
    }
    return acc;
}

// host: sinc_coef(beta=1, relfiltlen=8, decfactor=8192, pedestal=0, weight=1) (uniform_interp.f90:296-384) rearranged
// as prepareMethods does (topozeroMethods.f:57-61)
inline void sinc_make_table(float *fintp /* [kSincSub * kSincLen] */)
{
    const double pi = 4.0 * atan(1.0);
    const int nco = kSincLen * kSincSub;
    const double wgthgt = 0.5, soff = nco / 2.0;
    double *r = new double[nco];
    for (int i = 0; i < nco; i++) {
        double wa = i - soff;
        double sx = wa * 1.0 / (1.0 * kSincSub);
        double fct = (sx != 0.0) ? sin(pi * sx) / (pi * sx) : 1.0;
        double wgt = (1.0 - wgthgt) + wgthgt * cos((pi * wa) / soff);
        r[i] = fct * wgt;
    }
    for (int i = 0; i < kSincLen; i++)
        for (int j = 0; j < kSincSub; j++) fintp[i + j * kSincLen] = (float)r[j + i * kSincSub];
    delete[] r;
}


// AKIMA (components/isceobj/Util/src/akima_reg.F:54-317 through intp_akima, topozeroMethods.f:222-247), as written:
// the partial derivatives are taken at (ix+1..ix+2, iy+1..iy+2) while the corner values are those of
// (ix..ix+1, iy..iy+1) (getParDer :70-73 vs polyfitAkima :166-169), and the weights wx2/wx3/wy2/wy3 keep the value of
// the previous grid point when the "equal slopes" branch is taken (:81-86, :95-100; initialised to 0 here, which the
// guard at :113-120 turns into 1).  Sample differences are float32, everything else double, divisions IEEE.
B2_HD bool aki_almost_equal(double x, double y) { return fabs(x - y) <= 2.220446049250313e-16; }
B2_HD float interp_akima(const DemView &d, int i_x, int i_y, double f_x, double f_y)
{
    if ((i_x < 1) || (i_x >= (d.nx - 1))) return kBadValue;
    if ((i_y < 1) || (i_y >= (d.ny - 1))) return kBadValue;
#define Z(a, b) dem_at(d, (a), (b))
    double sx[4], sy[4], sxy[4]; // index (jj-1) + 2*(ii-1)
    double wx2 = 0.0, wx3 = 0.0, wy2 = 0.0, wy3 = 0.0;
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
        const int ii = (q >> 1) + 1, jj = (q & 1) + 1;
        int yy = i_y + ii;
        yy = yy < 3 ? 3 : (yy > d.ny - 2 ? d.ny - 2 : yy);
        int xx = i_x + jj;
        xx = xx < 3 ? 3 : (xx > d.nx - 2 ? d.nx - 2 : xx);
        const float zc = Z(xx, yy);
        float f;
        double m1, m2, m3, m4;
        const float zxm1 = Z(xx - 1, yy), zxp1 = Z(xx + 1, yy);
        f = zxm1 - Z(xx - 2, yy); m1 = f;
        f = zc - zxm1; m2 = f;
        f = zxp1 - zc; m3 = f;
        f = Z(xx + 2, yy) - zxp1; m4 = f;
        if (aki_almost_equal(m1, m2) && aki_almost_equal(m3, m4)) sx[q] = 0.5 * (m2 + m3);
        else {
            wx2 = fabs(m4 - m3);
            wx3 = fabs(m2 - m1);
            sx[q] = div_n(wx2 * m2 + wx3 * m3, wx2 + wx3);
        }
        f = Z(xx, yy - 1) - Z(xx, yy - 2); m1 = f;
        f = zc - Z(xx, yy - 1); m2 = f;
        f = Z(xx, yy + 1) - zc; m3 = f;
        f = Z(xx, yy + 2) - Z(xx, yy + 1); m4 = f;
        if (aki_almost_equal(m1, m2) && aki_almost_equal(m3, m4)) sy[q] = 0.5 * (m2 + m3);
        else {
            wy2 = fabs(m4 - m3);
            wy3 = fabs(m2 - m1);
            sy[q] = div_n(wy2 * m2 + wy3 * m3, wy2 + wy3);
        }
        double d22, d23, d42, d43;
        f = zxm1 - Z(xx - 1, yy - 1); d22 = f;
        f = Z(xx - 1, yy + 1) - zxm1; d23 = f;
        f = zxp1 - Z(xx + 1, yy - 1); d42 = f;
        f = Z(xx + 1, yy + 1) - zxp1; d43 = f;
        const double e22 = m2 - d22, e23 = m3 - d23, e32 = d42 - m2, e33 = d43 - m3;
        if (aki_almost_equal(wx2, 0.0) && aki_almost_equal(wx3, 0.0)) { wx2 = 1.; wx3 = 1.; }
        if (aki_almost_equal(wy2, 0.0) && aki_almost_equal(wy3, 0.0)) { wy2 = 1.; wy3 = 1.; }
        sxy[q] = div_n(wx2 * (wy2 * e22 + wy3 * e23) + wx3 * (wy2 * e32 + wy3 * e33), (wx2 + wx3) * (wy2 + wy3));
    }
    const double b1 = Z(i_x, i_y), b2 = Z(i_x + 1, i_y), b3 = Z(i_x + 1, i_y + 1), b4 = Z(i_x, i_y + 1);
#undef Z
    // sx(jj,ii): (1,1)->[0], (2,1)->[1], (1,2)->[2], (2,2)->[3]
    const double b5 = sx[0], b6 = sx[1], b7 = sx[3], b8 = sx[2];
    const double b9 = sy[0], b10 = sy[1], b11 = sy[3], b12 = sy[2];
    const double b13 = sxy[0], b14 = sxy[1], b15 = sxy[3], b16 = sxy[2];
    const double c1 = b1 - b2, c2 = b3 - b4, c3 = b5 + b6, c4 = b7 + b8, c5 = b9 - b10, c6 = b11 - b12, c7 = b13 + b14, c8 = b15 + b16;
    const double c9 = 2 * b5 + b6, c10 = b7 + 2 * b8, c11 = 2 * b13 + b14, c12 = b15 + 2 * b16, c13 = b5 - b8, c14 = b1 - b4;
    const double c15 = b13 + b16, c16 = 2 * b13 + b16, c17 = b9 + b12, c18 = 2 * b9 + b12;
    const double d1 = c1 + c2, d2 = c3 - c4, d3 = c5 - c6, d4 = c7 + c8, d5 = c9 - c10, d6 = 2 * c5 - c6, d7 = 2 * c7 + c8;
    const double d8 = c11 + c12, d9 = 2 * c11 + c12;
    const double f1 = 2 * d1 + d2, f2 = 2 * d3 + d4, f3 = 2 * d6 + d7, f4 = 3 * d1 + d5, f5 = 3 * d3 + d8, f6 = 3 * d6 + d9;
    const double v1 = 2 * f1 + f2, v2 = -(3 * f1 + f3), v3 = 2 * c5 + c7, v4 = 2 * c1 + c3;
    const double v5 = -(2 * f4 + f5), v6 = 3 * f4 + f6, v7 = -(3 * c5 + c11), v8 = -(3 * c1 + c9);
    const double v9 = 2 * c13 + c15, v10 = -(3 * c13 + c16), v11 = b13, v12 = b5;
    const double v13 = 2 * c14 + c17, v14 = -(3 * c14 + c18), v15 = b9, v16 = b1;
    const double x = f_x, y = f_y;
    const double p1 = ((v1 * y + v2) * y + v3) * y + v4;
    const double p2 = ((v5 * y + v6) * y + v7) * y + v8;
    const double p3 = ((v9 * y + v10) * y + v11) * y + v12;
    const double p4 = ((v13 * y + v14) * y + v15) * y + v16;
    return (float)(((p1 * x + p2) * x + p3) * x + p4);
}

// topozeroMethods.f:200-220
B2_HD float interp_nearest(const DemView &d, int i_x, int i_y, double f_x, double f_y)
{
    int dx = (int)lround(i_x + f_x), dy = (int)lround(i_y + f_y);
    if ((dx < 1) || (dx > d.nx)) return kBadValue;
    if ((dy < 1) || (dy > d.ny)) return kBadValue;
    return dem_at(d, dx, dy);
}

// uniform_interp.f90:123-130, DATA wt in column-major fill order: wt(i,k) = table[(k-1)*16 + (i-1)]
#define B2_BICUBIC_WT { \
    1, 0, -3, 2, 0, 0, 0, 0, -3, 0, 9, -6, 2, 0, -6, 4, \
    0, 0, 0, 0, 0, 0, 0, 0, 3, 0, -9, 6, -2, 0, 6, -4, \
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 9, -6, 0, 0, -6, 4, \
    0, 0, 3, -2, 0, 0, 0, 0, 0, 0, -9, 6, 0, 0, 6, -4, \
    0, 0, 0, 0, 1, 0, -3, 2, -2, 0, 6, -4, 1, 0, -3, 2, \
    0, 0, 0, 0, 0, 0, 0, 0, -1, 0, 3, -2, 1, 0, -3, 2, \
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -3, 2, 0, 0, 3, -2, \
    0, 0, 0, 0, 0, 0, 3, -2, 0, 0, -6, 4, 0, 0, 3, -2, \
    0, 1, -2, 1, 0, 0, 0, 0, 0, -3, 6, -3, 0, 2, -4, 2, \
    0, 0, 0, 0, 0, 0, 0, 0, 0, 3, -6, 3, 0, -2, 4, -2, \
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -3, 3, 0, 0, 2, -2, \
    0, 0, -1, 1, 0, 0, 0, 0, 0, 0, 3, -3, 0, 0, -2, 2, \
    0, 0, 0, 0, 0, 1, -2, 1, 0, -2, 4, -2, 0, 1, -2, 1, \
    0, 0, 0, 0, 0, 0, 0, 0, 0, -1, 2, -1, 0, 1, -2, 1, \
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, 0, 0, -1, 1, \
    0, 0, 0, 0, 0, 0, -1, 1, 0, 0, 2, -2, 0, 0, -1, 1}
#ifdef __CUDACC__
__device__ __constant__ static const signed char kBicubicWtDev[256] = B2_BICUBIC_WT;
#endif
static const signed char kBicubicWtHost[256] = B2_BICUBIC_WT;
#ifdef __CUDA_ARCH__
#define B2_BICUBIC_WT_AT(i) kBicubicWtDev[i]
#else
#define B2_BICUBIC_WT_AT(i) kBicubicWtHost[i]
#endif

// uniform_interp.f90:112-200 called as bicubic(dy, dx, dem).  z(a,b) == dem(lon=a, lat=b); sample
// differences are float32 (the Fortran subtracts real*4 values); the dzdy(2..4) column typo is kept.
B2_HD float interp_bicubic(const DemView &d, int i_x, int i_y, double f_x, double f_y)
{
    if ((i_x < 2) || (i_x >= (d.nx - 1))) return kBadValue;
    if ((i_y < 2) || (i_y >= (d.ny - 1))) return kBadValue;
    double x = i_y + f_y, y = i_x + f_x;
    int x1 = (int)floor(x), x2 = (int)ceil(x), y1 = (int)floor(y), y2 = (int)ceil(y);
#define Z(a, b) dem_at(d, (a), (b))
    double q[16];
    float f;
    q[0] = Z(y1, x1);
    q[3] = Z(y2, x1);
    q[1] = Z(y1, x2);
    q[2] = Z(y2, x2);
    f = Z(y1, x1 + 1) - Z(y1, x1 - 1); q[4] = f / 2.0;
    f = Z(y1, x2 + 1) - Z(y1, x2 - 1); q[5] = f / 2.0;
    f = Z(y2, x2 + 1) - Z(y2, x2 - 1); q[6] = f / 2.0;
    f = Z(y2, x1 + 1) - Z(y2, x1 - 1); q[7] = f / 2.0;
    f = Z(y1 + 1, x1) - Z(y1 - 1, x1); q[8] = f / 2.0;
    f = Z(y1 + 1, x2 + 1) - Z(y1 - 1, x2); q[9] = f / 2.0;
    f = Z(y2 + 1, x2 + 1) - Z(y2 - 1, x2); q[10] = f / 2.0;
    f = Z(y2 + 1, x1 + 1) - Z(y2 - 1, x1); q[11] = f / 2.0;
    f = Z(y1 + 1, x1 + 1) - Z(y1 - 1, x1 + 1); f = f - Z(y1 + 1, x1 - 1); f = f + Z(y1 - 1, x1 - 1); q[12] = 0.25 * f;
    f = Z(y2 + 1, x1 + 1) - Z(y2 - 1, x1 + 1); f = f - Z(y2 + 1, x1 - 1); f = f + Z(y2 - 1, x1 - 1); q[15] = 0.25 * f;
    f = Z(y1 + 1, x2 + 1) - Z(y1 - 1, x2 + 1); f = f - Z(y1 + 1, x2 - 1); f = f + Z(y1 - 1, x2 - 1); q[13] = 0.25 * f;
    f = Z(y2 + 1, x2 + 1) - Z(y2 - 1, x2 + 1); f = f - Z(y2 + 1, x2 - 1); f = f + Z(y2 - 1, x2 - 1); q[14] = 0.25 * f;
#undef Z
    double cl[16];
    for (int i = 0; i < 16; i++) {
        double qq = 0.0;
        for (int k = 0; k < 16; k++) {
            int w = B2_BICUBIC_WT_AT(k * 16 + i);
            if (w != 0) qq = qq + (double)w * q[k]; // adding 0*q(k) never changes qq
        }
        cl[i] = qq;
    }
    double t = (x - x1), u = (y - y1), r = 0.0;
    for (int i = 3; i >= 0; i--) r = t * r + ((cl[4 * i + 3] * u + cl[4 * i + 2]) * u + cl[4 * i + 1]) * u + cl[4 * i + 0];
    return (float)r;
}

// ---------------------------------------------------------------------------------------------
// "biquintic" == separable natural cubic spline over a 6x6 window (components/isceobj/Util/src/
// spline.f:15-117 as called at topozeroMethods.f:196).  The spline of spline.f is linear in the six
// samples and is always evaluated in the interval between the 2nd and 3rd node at X = 2 + frac, so it
// reduces to six weights that are cubic polynomials of frac.  The second-derivative rows R(2), R(3)
// come from the data-independent tridiagonal elimination of INITSPLINE (Q, P of spline.f:21-27):
//   R(k) = sum_j Rk[j] * Y(j).
// They are tabulated once on the host by running INITSPLINE on unit vectors (spline6_make_table).  The result
// differs from the reference's own evaluation order only by double rounding (~1e-16 relative) and is
// then rounded to float32 like the reference (sngl(temp), spline.f:115).
// ---------------------------------------------------------------------------------------------
// cubic-in-frac weight polynomials: w_j(xx) = tab[0][j] + xx*(tab[1][j] + xx*(tab[2][j] + xx*tab[3][j]))
struct Spline6Table {
    double c[4][6];
};

// Builds the table by running INITSPLINE's data-independent elimination on unit vectors (host, once):
//   Q(1)=0; P=Q(k-1)/2+2; Q(k)=-0.5/P; R(k)=(3*(Y(k+1)-2Y(k)+Y(k-1))-R(k-1)/2)/P; back substitution;
// then SPLINE at X = 2 + xx (J = 2, spline.f:50-52):
//   S = Y2 + xx*((Y3 - Y2 - R2/3 - R3/6) + xx*(R2/2 + xx*(R3 - R2)/6))
inline void spline6_make_table(Spline6Table &T)
{
    double Q[6], Rm[6][6];
    for (int j = 0; j < 6; j++) Rm[0][j] = 0.0;
    Q[0] = 0.0;
    for (int K = 1; K <= 4; K++) { // Fortran K = 2..5
        double P = Q[K - 1] / 2 + 2;
        Q[K] = -0.5 / P;
        for (int j = 0; j < 6; j++) {
            double d2 = ((j == K + 1) ? 1.0 : 0.0) - 2.0 * ((j == K) ? 1.0 : 0.0) + ((j == K - 1) ? 1.0 : 0.0);
            Rm[K][j] = (3 * d2 - Rm[K - 1][j] / 2) / P;
        }
    }
    for (int j = 0; j < 6; j++) Rm[5][j] = 0.0;
    for (int K = 4; K >= 1; K--)
        for (int j = 0; j < 6; j++) Rm[K][j] = Q[K] * Rm[K + 1][j] + Rm[K][j];
    for (int j = 0; j < 6; j++) {
        double R2 = Rm[1][j], R3 = Rm[2][j];
        T.c[0][j] = (j == 1) ? 1.0 : 0.0;
        T.c[1][j] = ((j == 2) ? 1.0 : 0.0) - ((j == 1) ? 1.0 : 0.0) - R2 / 3 - R3 / 6;
        T.c[2][j] = R2 / 2;
        T.c[3][j] = (R3 - R2) / 6;
    }
}

B2_HD float interp_biquintic(const DemView &d, const Spline6Table &T, int i_x, int i_y, double f_x, double f_y)
{
    if ((i_x < 3) || (i_x >= (d.nx - 2))) return kBadValue;
    if ((i_y < 3) || (i_y >= (d.ny - 2))) return kBadValue;
    // window floor-1 .. floor+4 on both axes, indices clamped to [1, n] (spline.f:87-104); only the last
    // column / row of the window can leave the grid once the checks above passed
    double wx[6];
#pragma unroll
    for (int j = 0; j < 6; j++) wx[j] = b2_fma(b2_fma(b2_fma(T.c[3][j], f_x, T.c[2][j]), f_x, T.c[1][j]), f_x, T.c[0][j]);
    double acc = 0.0;
#ifdef __CUDA_ARCH__
    // padded double copy: the replicated last column / row is what the clamp would have selected
    const double *row = d.d64 + (size_t)(i_y - 2) * (size_t)d.stride64 + (size_t)(i_x - 2);
#pragma unroll
    for (int J = 0; J < 6; J++, row += d.stride64) {
        double hc = wx[0] * __ldg(row);
        hc = b2_fma(wx[1], __ldg(row + 1), hc);
        hc = b2_fma(wx[2], __ldg(row + 2), hc);
        hc = b2_fma(wx[3], __ldg(row + 3), hc);
        hc = b2_fma(wx[4], __ldg(row + 4), hc);
        hc = b2_fma(wx[5], __ldg(row + 5), hc);
        double wy = b2_fma(b2_fma(b2_fma(T.c[3][J], f_y, T.c[2][J]), f_y, T.c[1][J]), f_y, T.c[0][J]);
        acc = b2_fma(wy, hc, acc);
    }
    return (float)acc;
#else
    const int x5 = (i_x + 4 > d.nx) ? d.nx : i_x + 4;
#pragma unroll
    for (int J = 0; J < 6; J++) { // latitude rows: six consecutive longitudes per row (one or two 32-byte sectors)
        int iy = i_y - 1 + J;
        iy = iy > d.ny ? d.ny : iy;
        const float *row = d.data + (size_t)(iy - 1) * (size_t)d.nx + (size_t)(i_x - 2);
        double hc = wx[0] * (double)B2_LDG(row);
        hc = b2_fma(wx[1], (double)B2_LDG(row + 1), hc);
        hc = b2_fma(wx[2], (double)B2_LDG(row + 2), hc);
        hc = b2_fma(wx[3], (double)B2_LDG(row + 3), hc);
        hc = b2_fma(wx[4], (double)B2_LDG(row + 4), hc);
        hc = b2_fma(wx[5], (double)B2_LDG(d.data + (size_t)(iy - 1) * (size_t)d.nx + (size_t)(x5 - 1)), hc);
        double wy = b2_fma(b2_fma(b2_fma(T.c[3][J], f_y, T.c[2][J]), f_y, T.c[1][J]), f_y, T.c[0][J]);
        acc = b2_fma(wy, hc, acc);
    }
    return (float)acc;
#endif
}

// The four slope probes of the incidence computation (topozero.f90:680-688): the interpolator at (i_x-1, i_y),
// (i_x+1, i_y), (i_x, i_y-1), (i_x, i_y+1) with the same fractions.  Their 6x6 windows overlap in an 8x8 block (less
// its corners); the row interpolants are shared between them, 60 samples and 12 weights instead of 4 x (36 + 12).
// Every probe goes through exactly the operation sequence of interp_biquintic, so the values are bit-identical to four
// separate calls.  Returns false (nothing computed) when a probe window touches the DEM edge; the caller then makes
// the four calls.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ bool biquintic_probes4(const DemView &d, const Spline6Table &T, int i_x, int i_y, double f_x, double f_y,
                                                  double &p0, double &p1, double &p2, double &p3)
{
    if (i_x < 4 || i_x + 1 >= d.nx - 2 || i_y < 4 || i_y + 1 >= d.ny - 2) return false;
    double wx[6], wy[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
        wx[j] = b2_fma(b2_fma(b2_fma(T.c[3][j], f_x, T.c[2][j]), f_x, T.c[1][j]), f_x, T.c[0][j]);
        wy[j] = b2_fma(b2_fma(b2_fma(T.c[3][j], f_y, T.c[2][j]), f_y, T.c[1][j]), f_y, T.c[0][j]);
    }
    // rows i_y-3 .. i_y+4, columns i_x-3 .. i_x+4 of the padded double copy (0-based = 1-based index - 1)
    const double *row = d.d64 + (size_t)(i_y - 3) * (size_t)d.stride64 + (size_t)(i_x - 3);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int r = 0; r < 8; r++, row += d.stride64) {
        const double v1 = __ldg(row + 1), v2 = __ldg(row + 2), v3 = __ldg(row + 3), v4 = __ldg(row + 4), v5 = __ldg(row + 5),
                     v6 = __ldg(row + 6);
        double h = wx[0] * v1; // centre columns: the row interpolant of the probes at i_y -+ 1
        h = b2_fma(wx[1], v2, h);
        h = b2_fma(wx[2], v3, h);
        h = b2_fma(wx[3], v4, h);
        h = b2_fma(wx[4], v5, h);
        h = b2_fma(wx[5], v6, h);
        if (r <= 5) a2 = b2_fma(wy[r <= 5 ? r : 0], h, a2);
        if (r >= 2) a3 = b2_fma(wy[r >= 2 ? r - 2 : 0], h, a3);
        if (r >= 1 && r <= 6) { // rows of the probes at i_x -+ 1
            const double v0 = __ldg(row), v7 = __ldg(row + 7);
            double hm = wx[0] * v0;
            hm = b2_fma(wx[1], v1, hm);
            hm = b2_fma(wx[2], v2, hm);
            hm = b2_fma(wx[3], v3, hm);
            hm = b2_fma(wx[4], v4, hm);
            hm = b2_fma(wx[5], v5, hm);
            double hp = wx[0] * v2;
            hp = b2_fma(wx[1], v3, hp);
            hp = b2_fma(wx[2], v4, hp);
            hp = b2_fma(wx[3], v5, hp);
            hp = b2_fma(wx[4], v6, hp);
            hp = b2_fma(wx[5], v7, hp);
            const double w = wy[(r >= 1 && r <= 6) ? r - 1 : 0];
            a0 = b2_fma(w, hm, a0);
            a1 = b2_fma(w, hp, a1);
        }
    }
    p0 = (double)(float)a0;
    p1 = (double)(float)a1;
    p2 = (double)(float)a2;
    p3 = (double)(float)a3;
    return true;
}
#endif

} // namespace b2
