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
    const int x5 = (i_x + 4 > d.nx) ? d.nx : i_x + 4;
    double acc = 0.0;
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
}

} // namespace b2
