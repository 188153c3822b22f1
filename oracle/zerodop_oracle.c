/*
 * zerodop_oracle.c -- TEST INFRASTRUCTURE ONLY (see zerodop_oracle.h).
 *
 * Line-by-line CPU restatement of the ISCE2 zero-Doppler topozero / geo2rdr path.
 * Build:  gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC
 * (no FMA contraction: the reference is Fortran/C built for generic x86-64, which
 * has no fused multiply-add, so every product and sum rounds separately).
 *
 * All citations are relative to /root/reference.
 */
#include "zerodop_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* linalg3: components/isceobj/Util/Library/linalg3/src/linalg3Module.F */
/* ------------------------------------------------------------------ */
static inline void v_cross(const double *u, const double *v, double *w) /* :37-76 */
{
    double w0 = u[1] * v[2] - u[2] * v[1];
    double w1 = u[2] * v[0] - u[0] * v[2];
    double w2 = u[0] * v[1] - u[1] * v[0];
    w[0] = w0; w[1] = w1; w[2] = w2;
}
static inline double v_dot(const double *v, const double *w) /* :78-115 */
{
    return v[0] * w[0] + v[1] * w[1] + v[2] * w[2];
}
static inline double v_norm(const double *v) /* :256-292 */
{
    return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}
static inline void v_unit(const double *v, double *u) /* :338-382: leaves u untouched if |v| == 0 */
{
    double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (n != 0) {
        double a = v[0] / n, b = v[1] / n, c = v[2] / n;
        u[0] = a; u[1] = b; u[2] = c;
    }
}
/* matvec with Fortran r_t(i,j) stored row-major m[3*i+j] : :212-254 */
static inline void m_vec(const double *m, const double *v, double *w)
{
    double w0 = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
    double w1 = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
    double w2 = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    w[0] = w0; w[1] = w1; w[2] = w2;
}

/* ------------------------------------------------------------------ */
/* geometry                                                            */
/* ------------------------------------------------------------------ */
/* components/isceobj/Util/Library/geometry/src/latlon.F:44-71 */
void orc_latlon(double r_a, double r_e2, double *r_v, double *r_llh, int i_type)
{
    if (i_type == 1) { /* LLH_2_XYZ :44-49 */
        double sl = sin(r_llh[0]);
        double r_re = r_a / sqrt(1.0 - r_e2 * (sl * sl));
        r_v[0] = (r_re + r_llh[2]) * cos(r_llh[0]) * cos(r_llh[1]);
        r_v[1] = (r_re + r_llh[2]) * cos(r_llh[0]) * sin(r_llh[1]);
        r_v[2] = (r_re * (1.0 - r_e2) + r_llh[2]) * sin(r_llh[0]);
    } else { /* XYZ_2_LLH :51-71 */
        double r_q2 = (r_v[0] * r_v[0] + r_v[1] * r_v[1]);
        double r_q3 = r_a * r_a;
        double r_e4 = r_e2 * r_e2;
        double r_p = r_q2 / r_q3;
        double r_q = (1.0 - r_e2) * (r_v[2] * r_v[2]) / r_q3;
        double r_r = (r_p + r_q - r_e4) / 6.0;
        double r_s = (r_e4 * r_p * r_q) / (4.0 * (r_r * r_r * r_r));
        double r_t = pow(1.0 + r_s + sqrt(r_s * (2.0 + r_s)), 1.0 / 3.0);
        double r_u = r_r * (1.0 + r_t + 1.0 / r_t);
        double r_rv = sqrt(r_u * r_u + r_e4 * r_q);
        double r_w = r_e2 * (r_u + r_rv - r_q) / (2.0 * r_rv);
        double r_k = sqrt(r_u + r_rv + r_w * r_w) - r_w;
        double r_d = r_k * sqrt(r_q2) / (r_k + r_e2);
        r_llh[0] = atan2(r_v[2], r_d);
        r_llh[1] = atan2(r_v[1], r_v[0]);
        r_llh[2] = (r_k + r_e2 - 1.0) * sqrt(r_d * r_d + r_v[2] * r_v[2]) / r_k;
    }
}

/* components/isceobj/Util/Library/geometry/src/curvature.F:26-64 */
double orc_reast(double a, double e2, double lat)
{
    double s = sin(lat);
    return a / sqrt(1.0 - e2 * (s * s));
}
double orc_rnorth(double a, double e2, double lat)
{
    double s = sin(lat);
    return (a * (1.0 - e2)) / pow(1.0 - e2 * (s * s), 1.5);
}
double orc_rdir(double a, double e2, double hdg, double lat)
{
    double re = orc_reast(a, e2, lat);
    double rn = orc_rnorth(a, e2, lat);
    double c = cos(hdg), s = sin(hdg);
    return (re * rn) / (re * (c * c) + rn * (s * s));
}

/* components/isceobj/Util/Library/geometry/src/tcnbasis.F:26-39 */
void orc_tcnbasis(const double *pos, const double *vel, double a, double e2, double *r_t, double *r_c, double *r_n)
{
    double llh[3], tmp[3], p[3] = {pos[0], pos[1], pos[2]};
    orc_latlon(a, e2, p, llh, 2);
    double lat = llh[0], lon = llh[1];
    r_n[0] = -cos(lat) * cos(lon);
    r_n[1] = -cos(lat) * sin(lon);
    r_n[2] = -sin(lat);
    v_cross(r_n, vel, tmp);
    v_unit(tmp, r_c);
    v_cross(r_c, r_n, tmp);
    v_unit(tmp, r_t);
}

/* components/isceobj/Util/Library/geometry/src/enubasis.F:39-60; m[3*i+j] = r_enumat(i+1,j+1) */
void orc_enubasis(double r_lat, double r_lon, double *m)
{
    double clt = cos(r_lat), slt = sin(r_lat), clo = cos(r_lon), slo = sin(r_lon);
    m[0 * 3 + 1] = -slt * clo; m[1 * 3 + 1] = -slt * slo; m[2 * 3 + 1] = clt;   /* north */
    m[0 * 3 + 0] = -slo;       m[1 * 3 + 0] = clo;        m[2 * 3 + 0] = 0.0;   /* east  */
    m[0 * 3 + 2] = clt * clo;  m[1 * 3 + 2] = clt * slo;  m[2 * 3 + 2] = slt;   /* up    */
}

/* components/isceobj/Util/Library/geometry/src/radar_to_xyz.F:49-92 */
double orc_radar_to_xyz(double a, double e2, double plat, double plon, double phdg, double *mat, double *ov)
{
    double clt = cos(plat), slt = sin(plat), clo = cos(plon), slo = sin(plon);
    double chg = cos(phdg), shg = sin(phdg);
    mat[0] = clt * clo;
    mat[1] = -shg * slo - slt * clo * chg;
    mat[2] = slo * chg - slt * clo * shg;
    mat[3] = clt * slo;
    mat[4] = clo * shg - slt * slo * chg;
    mat[5] = -clo * chg - slt * slo * shg;
    mat[6] = slt;
    mat[7] = clt * chg;
    mat[8] = clt * shg;
    double radcur = orc_rdir(a, e2, phdg, plat);
    double llh[3] = {plat, plon, 0.0}, p[3];
    orc_latlon(a, e2, p, llh, 1);
    double up[3] = {clt * clo, clt * slo, slt};
    for (int i = 0; i < 3; i++) ov[i] = p[i] - radcur * up[i];
    return radcur;
}

/* components/isceobj/Util/Library/geometry/src/convert_sch_to_xyz.F:63-72 (XYZ_2_SCH branch) */
void orc_xyz_to_sch(const double *mat, const double *ov, double radcur, const double *xyz, double *sch)
{
    double t[3], s[3], llh[3], minv[9];
    /* lincomb(1, xyz, -1, ov) : linalg3Module.F:117-159 */
    for (int i = 0; i < 3; i++) t[i] = 1.0 * xyz[i] + (-1.0) * ov[i];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) minv[3 * i + j] = mat[3 * j + i];
    m_vec(minv, t, s);
    orc_latlon(radcur, 0.0, s, llh, 2);
    sch[0] = radcur * llh[1];
    sch[1] = radcur * llh[0];
    sch[2] = llh[2];
}

/* ------------------------------------------------------------------ */
/* orbit: components/isceobj/Util/Library/orbit/src                    */
/* ------------------------------------------------------------------ */
/* orbitHermite.c:4-94 */
static void orbit_hermite(double x[][3], double v[][3], const double *t, double time, double *xx, double *vv)
{
    double h[4], hdot[4], f0[4], f1[4], g0[4], g1[4], sum, product;
    const int n1 = 4, n2 = 3;
    for (int i = 0; i < n1; ++i) {
        f1[i] = time - t[i];
        sum = 0.0;
        for (int j = 0; j < n1; ++j)
            if (i != j) sum += 1.0 / (t[i] - t[j]);
        f0[i] = 1.0 - 2.0 * (time - t[i]) * sum;
    }
    for (int i = 0; i < n1; ++i) {
        product = 1.0;
        for (int k = 0; k < n1; ++k)
            if (k != i) product *= (time - t[k]) / (t[i] - t[k]);
        h[i] = product;
        sum = 0.0;
        for (int j = 0; j < n1; ++j) {
            product = 1.0;
            for (int k = 0; k < n1; ++k)
                if ((k != i) && (k != j)) product *= (time - t[k]) / (t[i] - t[k]);
            if (j != i) sum += 1.0 / (t[i] - t[j]) * product;
        }
        hdot[i] = sum;
    }
    for (int i = 0; i < n1; ++i) {
        g1[i] = h[i] + 2.0 * (time - t[i]) * hdot[i];
        sum = 0.0;
        for (int j = 0; j < n1; ++j)
            if (i != j) sum += 1.0 / (t[i] - t[j]);
        g0[i] = 2.0 * (f0[i] * hdot[i] - h[i] * sum);
    }
    for (int k = 0; k < n2; ++k) {
        sum = 0.0;
        for (int i = 0; i < n1; ++i) sum += (x[i][k] * f0[i] + v[i][k] * f1[i]) * h[i] * h[i];
        xx[k] = sum;
        sum = 0.0;
        for (int i = 0; i < n1; ++i) sum += (x[i][k] * g0[i] + v[i][k] * g1[i]) * h[i];
        vv[k] = sum;
    }
}

/* orbit.c:175-234 interpolateWGS84Orbit */
int orc_interp_hermite(const orc_orbit *orb, double tintp, double *opos, double *ovel)
{
    double pos[4][3], vel[4][3], t[4];
    int i, j;
    if (orb->nvec < 4) return 1;
    for (i = 0; i < orb->nvec; i++)
        if (orb->t[i] >= tintp) break;
    i -= 2;
    if (i < 0) i = 0;
    if (i > orb->nvec - 4) i = orb->nvec - 4;
    for (j = 0; j < 4; j++) {
        t[j] = orb->t[i + j];
        for (int k = 0; k < 3; k++) {
            pos[j][k] = orb->pos[3 * (i + j) + k];
            vel[j][k] = orb->vel[3 * (i + j) + k];
        }
    }
    orbit_hermite(pos, vel, t, tintp, opos, ovel);
    if ((tintp < orb->t[0]) || (tintp > orb->t[orb->nvec - 1])) return 1;
    return 0;
}

/* orbit.c:236-314 interpolateLegendreOrbit */
int orc_interp_legendre(const orc_orbit *orb, double tintp, double *opos, double *ovel)
{
    int i, j;
    double pos[9][3], vel[9][3], t[9], trel, coeff, teller;
    const double noemer[] = {40320.0, -5040.0, 1440.0, -720.0, 576.0, -720.0, 1440.0, -5040.0, 40320.0};
    opos[0] = opos[1] = opos[2] = 0.0;
    ovel[0] = ovel[1] = ovel[2] = 0.0;
    if (orb->nvec < 9) return 1;
    for (i = 0; i < orb->nvec; i++)
        if (orb->t[i] >= tintp) break;
    i -= 5;
    if (i < 0) i = 0;
    if (i > orb->nvec - 9) i = orb->nvec - 9;
    for (j = 0; j < 9; j++) {
        t[j] = orb->t[i + j];
        for (int k = 0; k < 3; k++) {
            pos[j][k] = orb->pos[3 * (i + j) + k];
            vel[j][k] = orb->vel[3 * (i + j) + k];
        }
    }
    trel = 8.0 * (tintp - t[0]) / (t[8] - t[0]);
    teller = 1.0;
    for (j = 0; j < 9; j++) teller *= (trel - j);
    if (teller == 0.0) {
        i = (int)trel;
        for (j = 0; j < 3; j++) {
            opos[j] = pos[i][j];
            ovel[j] = vel[i][j];
        }
    } else {
        for (i = 0; i < 9; i++) {
            coeff = teller / noemer[i] / (trel - i);
            for (j = 0; j < 3; j++) {
                opos[j] += coeff * pos[i][j];
                ovel[j] += coeff * vel[i][j];
            }
        }
    }
    if ((tintp < orb->t[0]) || (tintp > orb->t[orb->nvec - 1])) return 1;
    return 0;
}

/* orbit.c:119-172 interpolateSCHOrbit (Lagrange over all state vectors) */
int orc_interp_sch(const orc_orbit *orb, double tintp, double *opos, double *ovel)
{
    if (orb->nvec < 2) return 1;
    if ((tintp < orb->t[0]) || (tintp > orb->t[orb->nvec - 1])) return 1;
    opos[0] = opos[1] = opos[2] = 0.0;
    ovel[0] = ovel[1] = ovel[2] = 0.0;
    for (int i = 0; i < orb->nvec; i++) {
        double frac = 1.0;
        double t0 = orb->t[i];
        for (int j = 0; j < orb->nvec; j++) {
            if (i == j) continue;
            double t1 = orb->t[j];
            double num = t1 - tintp;
            double den = t1 - t0;
            frac *= num / den;
        }
        for (int k = 0; k < 3; k++) {
            opos[k] += frac * orb->pos[3 * i + k];
            ovel[k] += frac * orb->vel[3 * i + k];
        }
    }
    return 0;
}

/* orbit.c:316-354 computeAcceleration (always Hermite) */
int orc_compute_acceleration(const orc_orbit *orb, double tintp, double *acc)
{
    double xbef[3], vbef[3], xaft[3], vaft[3], temp;
    acc[0] = acc[1] = acc[2] = 0.0;
    temp = tintp - 0.01;
    if (orc_interp_hermite(orb, temp, xbef, vbef) != 0) return 1;
    temp = tintp + 0.01;
    if (orc_interp_hermite(orb, temp, xaft, vaft) != 0) return 1;
    for (int i = 0; i < 3; i++) acc[i] = (vaft[i] - vbef[i]) / 0.02;
    return 0;
}

static int interp_orbit(int method, const orc_orbit *o, double t, double *p, double *v)
{
    if (method == ORC_HERMITE) return orc_interp_hermite(o, t, p, v);
    if (method == ORC_SCH) return orc_interp_sch(o, t, p, v);
    return orc_interp_legendre(o, t, p, v);
}

/* ------------------------------------------------------------------ */
/* polynomials                                                         */
/* ------------------------------------------------------------------ */
/* components/isceobj/Util/Library/poly2d/src/poly2d.c:92-111 */
double orc_eval_poly2d(const orc_poly2d *poly, double azi, double rng)
{
    double value = 0.0, scalex, scaley;
    double xval = (rng - poly->mean_range) / (poly->norm_range);
    double yval = (azi - poly->mean_azimuth) / (poly->norm_azimuth);
    int i, j;
    scaley = 1.0;
    for (i = 0; i <= poly->azimuth_order; i++, scaley *= yval) {
        scalex = 1.0;
        for (j = 0; j <= poly->range_order; j++, scalex *= xval)
            value += scalex * scaley * poly->coeffs[i * (poly->range_order + 1) + j];
    }
    return value;
}

/* components/isceobj/Util/Library/poly1d/src/poly1d.c:87-104 */
double orc_eval_poly1d(const orc_poly1d *poly, double xin)
{
    double value = 0.0, scalex = 1.0;
    double xval = (xin - poly->mean) / (poly->norm);
    for (int i = 0; i <= poly->order; i++, scalex *= xval) value += scalex * poly->coeffs[i];
    return value;
}

/* ------------------------------------------------------------------ */
/* DEM interpolators.  DEM(ix,iy) with 1-based Fortran indices, lon (x) fastest. */
/* ------------------------------------------------------------------ */
#define DEM(ix, iy) dem[(size_t)((iy) - 1) * (size_t)nx + (size_t)((ix) - 1)]
#define ORC_BADVALUE (-1000.0f) /* topozeroMethods.f:33 */

/* components/isceobj/Util/src/uniform_interp.f90:13-44, called as bilinear(dy,dx,dem) (topozeroMethods.f:145) */
static double bilinear(double x, double y, const float *dem, int nx)
{
    double x1 = floor(x), x2 = ceil(x), y1 = ceil(y), y2 = floor(y);
    double q11 = DEM((int)y1, (int)x1);
    double q12 = DEM((int)y2, (int)x1);
    double q21 = DEM((int)y1, (int)x2);
    double q22 = DEM((int)y2, (int)x2);
    if (y1 == y2 && x1 == x2) return q11;
    if (y1 == y2) return (x2 - x) / (x2 - x1) * q11 + (x - x1) / (x2 - x1) * q21;
    if (x1 == x2) return (y2 - y) / (y2 - y1) * q11 + (y - y1) / (y2 - y1) * q12;
    return q11 * (x2 - x) * (y2 - y) / ((x2 - x1) * (y2 - y1)) +
           q21 * (x - x1) * (y2 - y) / ((x2 - x1) * (y2 - y1)) +
           q12 * (x2 - x) * (y - y1) / ((x2 - x1) * (y2 - y1)) +
           q22 * (x - x1) * (y - y1) / ((x2 - x1) * (y2 - y1));
}

/* uniform_interp.f90:123-130 DATA wt (column-major fill): wt(i,k) = WT_FLAT[(k-1)*16 + (i-1)] */
static const double WT_FLAT[256] = {
    1, 0, -3, 2, 0, 0, 0, 0, -3, 0, 9, -6, 2, 0, -6, 4,
    0, 0, 0, 0, 0, 0, 0, 0, 3, 0, -9, 6, -2, 0, 6, -4,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 9, -6, 0, 0, -6, 4,
    0, 0, 3, -2, 0, 0, 0, 0, 0, 0, -9, 6, 0, 0, 6, -4,
    0, 0, 0, 0, 1, 0, -3, 2, -2, 0, 6, -4, 1, 0, -3, 2,
    0, 0, 0, 0, 0, 0, 0, 0, -1, 0, 3, -2, 1, 0, -3, 2,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -3, 2, 0, 0, 3, -2,
    0, 0, 0, 0, 0, 0, 3, -2, 0, 0, -6, 4, 0, 0, 3, -2,
    0, 1, -2, 1, 0, 0, 0, 0, 0, -3, 6, -3, 0, 2, -4, 2,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 3, -6, 3, 0, -2, 4, -2,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -3, 3, 0, 0, 2, -2,
    0, 0, -1, 1, 0, 0, 0, 0, 0, 0, 3, -3, 0, 0, -2, 2,
    0, 0, 0, 0, 0, 1, -2, 1, 0, -2, 4, -2, 0, 1, -2, 1,
    0, 0, 0, 0, 0, 0, 0, 0, 0, -1, 2, -1, 0, 1, -2, 1,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, 0, 0, -1, 1,
    0, 0, 0, 0, 0, 0, -1, 1, 0, 0, 2, -2, 0, 0, -1, 1};

/* uniform_interp.f90:112-200, called as bicubic(dy,dx,dem) (topozeroMethods.f:171).
 * z(a,b) in the Fortran is dem(lon=a, lat=b); differences of two real*4 samples are real*4. */
static double bicubic(double x, double y, const float *dem, int nx)
{
    int x1 = (int)floor(x), x2 = (int)ceil(x), y1 = (int)floor(y), y2 = (int)ceil(y);
    double zz[4], dzdx[4], dzdy[4], dzdxy[4], q[16], cl[16], c[4][4];
    float f;
    zz[0] = DEM(y1, x1);
    zz[3] = DEM(y2, x1);
    zz[1] = DEM(y1, x2);
    zz[2] = DEM(y2, x2);
    f = DEM(y1, x1 + 1) - DEM(y1, x1 - 1); dzdx[0] = f / 2.0;
    f = DEM(y1, x2 + 1) - DEM(y1, x2 - 1); dzdx[1] = f / 2.0;
    f = DEM(y2, x2 + 1) - DEM(y2, x2 - 1); dzdx[2] = f / 2.0;
    f = DEM(y2, x1 + 1) - DEM(y2, x1 - 1); dzdx[3] = f / 2.0;
    f = DEM(y1 + 1, x1) - DEM(y1 - 1, x1);         dzdy[0] = f / 2.0;
    f = DEM(y1 + 1, x2 + 1) - DEM(y1 - 1, x2);     dzdy[1] = f / 2.0; /* :152 typo kept */
    f = DEM(y2 + 1, x2 + 1) - DEM(y2 - 1, x2);     dzdy[2] = f / 2.0; /* :153 typo kept */
    f = DEM(y2 + 1, x1 + 1) - DEM(y2 - 1, x1);     dzdy[3] = f / 2.0; /* :154 typo kept */
    f = DEM(y1 + 1, x1 + 1) - DEM(y1 - 1, x1 + 1); f = f - DEM(y1 + 1, x1 - 1); f = f + DEM(y1 - 1, x1 - 1);
    dzdxy[0] = 0.25 * f;
    f = DEM(y2 + 1, x1 + 1) - DEM(y2 - 1, x1 + 1); f = f - DEM(y2 + 1, x1 - 1); f = f + DEM(y2 - 1, x1 - 1);
    dzdxy[3] = 0.25 * f;
    f = DEM(y1 + 1, x2 + 1) - DEM(y1 - 1, x2 + 1); f = f - DEM(y1 + 1, x2 - 1); f = f + DEM(y1 - 1, x2 - 1);
    dzdxy[1] = 0.25 * f;
    f = DEM(y2 + 1, x2 + 1) - DEM(y2 - 1, x2 + 1); f = f - DEM(y2 + 1, x2 - 1); f = f + DEM(y2 - 1, x2 - 1);
    dzdxy[2] = 0.25 * f;
    for (int i = 0; i < 4; i++) {
        q[i] = zz[i];
        q[i + 4] = dzdx[i];
        q[i + 8] = dzdy[i];
        q[i + 12] = dzdxy[i];
    }
    for (int i = 0; i < 16; i++) {
        double qq = 0.0;
        for (int k = 0; k < 16; k++) qq = qq + WT_FLAT[k * 16 + i] * q[k];
        cl[i] = qq;
    }
    int l = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) c[i][j] = cl[l++];
    double t = (x - x1), u = (y - y1), r = 0.0;
    for (int i = 3; i >= 0; i--) r = t * r + ((c[i][3] * u + c[i][2]) * u + c[i][1]) * u + c[i][0];
    return r;
}

/* components/isceobj/Util/src/spline.f:15-33 */
static void initspline(const double *Y, int N, double *R, double *Q)
{
    /* 0-based arrays hold the Fortran 1-based entries at [k-1] */
    Q[0] = 0.0;
    R[0] = 0.0;
    for (int K = 2; K <= N - 1; K++) {
        double P = Q[K - 2] / 2 + 2;
        Q[K - 1] = -0.5 / P;
        R[K - 1] = (3 * (Y[K] - 2 * Y[K - 1] + Y[K - 2]) - R[K - 2] / 2) / P;
    }
    R[N - 1] = 0.0;
    for (int K = N - 1; K >= 2; K--) R[K - 1] = Q[K - 1] * R[K] + R[K - 1];
}
/* spline.f:5-13 */
static int ifrac(double r)
{
    int i = (int)r;
    if (r >= 0) return i;
    if (r == i) return i;
    return i - 1;
}
/* spline.f:36-54 */
static double spline(double X, const double *Y, int N, const double *R)
{
    if (X < 1) return Y[0] + (X - 1) * (Y[1] - Y[0] - R[1] / 6);
    if (X > N) return Y[N - 1] + (X - N) * (Y[N - 1] - Y[N - 2] + R[N - 2] / 6);
    int J = ifrac(X);
    double XX = X - J;
    /* Fortran reads Y(J+1), R(J+1): J == N only when X == N exactly; not reachable from interp2DSpline */
    return Y[J - 1] + XX * ((Y[J] - Y[J - 1] - R[J - 1] / 3 - R[J] / 6) + XX * (R[J - 1] / 2 + XX * (R[J] - R[J - 1]) / 6));
}
/* spline.f:58-117 interp2DSpline(order, nx=ny_dem, ny=nx_dem, z, x=dy, y=dx) as called at topozeroMethods.f:276 */
static float interp2dspline(int order, const float *dem, int nx /*lon count*/, int ny /*lat count*/, double x /*lat idx*/, double y /*lon idx*/)
{
    double A[20], R[20], Q[20], HC[20];
    int I0, J0;
    int lodd = (order / 2) * 2 != order;
    if (lodd) {
        I0 = (int)(y - 0.5);
        J0 = (int)(x - 0.5);
    } else {
        I0 = (int)y;
        J0 = (int)x;
    }
    I0 = I0 - order / 2 + 1;
    J0 = J0 - order / 2 + 1;
    for (int I = 1; I <= order; I++) {
        int INDI = I0 + I;
        if (INDI < 1) INDI = 1;
        if (INDI > nx) INDI = nx; /* Fortran local "ny" == lon count */
        for (int J = 1; J <= order; J++) {
            int INDJ = J0 + J;
            if (INDJ < 1) INDJ = 1;
            if (INDJ > ny) INDJ = ny; /* Fortran local "nx" == lat count */
            A[J - 1] = DEM(INDI, INDJ);
        }
        initspline(A, order, R, Q);
        HC[I - 1] = spline(x - J0, A, order, R);
    }
    initspline(HC, order, R, Q);
    double temp = spline(y - I0, HC, order, R);
    return (float)temp;
}

/* ---- sinc: uniform_interp.f90:296-384 (sinc_coef) + topozeroMethods.f:46-63 (prepareMethods) ---- */
#define SINC_SUB 8192
#define SINC_LEN 8
static float *g_fintp = NULL;
static int g_sinc_cpp_arith;
static const float *sinc_table(void)
{
#pragma omp critical(orc_sinc_table)
    if (!g_fintp) {
        const double pi = 4.0 * atan(1.0);
        const double r_beta = 1.0, r_relfiltlen = 1.0 * SINC_LEN, r_pedestal = 0.0;
        const int i_decfactor = SINC_SUB;
        const int i_intplength = (int)lround(r_relfiltlen / r_beta);
        const int i_filtercoef = i_intplength * i_decfactor;
        const double r_wgthgt = (1.0 - r_pedestal) / 2.0;
        const double r_soff = i_filtercoef / 2.0;
        double *r_filter = malloc(sizeof(double) * (i_filtercoef + 1));
        for (int i = 0; i < i_filtercoef; i++) {
            double r_wa = i - r_soff;
            double r_s = r_wa * r_beta / (1.0 * i_decfactor);
            double r_fct = (r_s != 0.0) ? sin(pi * r_s) / (pi * r_s) : 1.0;
            double r_wgt = (1.0 - r_wgthgt) + r_wgthgt * cos((pi * r_wa) / r_soff);
            r_filter[i] = r_fct * r_wgt;
        }
        float *f = malloc(sizeof(float) * SINC_SUB * SINC_LEN);
        for (int i = 0; i < SINC_LEN; i++)
            for (int j = 0; j < SINC_SUB; j++) f[i + j * SINC_LEN] = (float)r_filter[j + i * SINC_SUB];
        free(r_filter);
        g_fintp = f;
    }
    return g_fintp;
}
void orc_sinc_table(float *out) { memcpy(out, sinc_table(), sizeof(float) * SINC_SUB * SINC_LEN); }
/* test hook (tests/test_oracle_cpp_pins.py): replace the topo / geozero sinc table -- with the one the reference's C++
 * builds by its older formula (UniformInterp.cpp:150-168), so that whole images interpolated with SINC can be compared with
 * Topo.cpp, which builds its table internally; NULL restores the table of the current formula on next use */
void orc_test_set_sinc_table(const float *table)
{
#pragma omp critical(orc_sinc_table)
    {
        free(g_fintp);
        g_fintp = NULL;
        if (table) {
            g_fintp = malloc(sizeof(float) * SINC_SUB * SINC_LEN);
            memcpy(g_fintp, table, sizeof(float) * SINC_SUB * SINC_LEN);
        }
    }
}

/* uniform_interp.f90:407-430 sinc_eval_2d_f: everything in real*4, k outer / m inner */
static float sinc_eval_2d_f(const float *dem, const float *intarr, int idec, int ilen, int intpx, int intpy, double frpx,
                            double frpy, int nx, int ny)
{
    float acc = 0.f;
    if ((intpx >= ilen - 1 && intpx < nx) && (intpy >= ilen - 1 && intpy < ny)) {
        int ifracx = (int)(frpx * idec), ifracy = (int)(frpy * idec);
        ifracx = ifracx < 0 ? 0 : (ifracx > idec - 1 ? idec - 1 : ifracx);
        ifracy = ifracy < 0 ? 0 : (ifracy > idec - 1 ? idec - 1 : ifracy);
        for (int k = 0; k < ilen; k++)
            for (int m = 0; m < ilen; m++) {
                /* arrin(intpx-k, intpy-m) is 0-based: dem element (intpx-k+1, intpy-m+1) in 1-based terms */
                float a = DEM(intpx - k + 1, intpy - m + 1);
                if (g_sinc_cpp_arith) { /* test hook, see g_akima_cpp_quirks */
                    acc = (float)(acc + (a * (double)intarr[k + ifracx * ilen] * (double)intarr[m + ifracy * ilen]));
                    continue;
                }
                float t = a * intarr[k + ifracx * ilen];
                t = t * intarr[m + ifracy * ilen];
                acc = acc + t;
            }
    }
    return acc;
}


/* ---- Akima: components/isceobj/Util/src/akima_reg.F:54-317 as called by intp_akima (topozeroMethods.f:222-247) ----
 * Kept as written, including (a) the slopes being taken at (ix+1..ix+2, iy+1..iy+2) while the values are taken at
 * (ix..ix+1, iy..iy+1) (getParDer :70-73 vs polyfitAkima :166-169) and (b) wx2/wx3/wy2/wy3 keeping their value from
 * the previous grid point when the "equal slopes" branch is taken (:81-86, :95-100; the Fortran leaves them
 * unassigned there -- they are initialised to 0 here, which the subsequent guard turns into 1). */
/* TEST HOOK (tests/test_oracle_cpp_pins.py only): when set, the two places where the reference's C++ restatement
 * (components/zerodop/GPUtopozero/src/AkimaLib.cpp) departs from akima_reg.F are reproduced, so that every OTHER line of
 * this function can be held bit for bit against that reference-authored code: (1) AkimaLib.cpp:70-76 stores the Y slope
 * into slpx and leaves slpy zero (akima_reg.F:95-100 stores slpy); (2) Constants.h:13 declares AKI_EPS as an int, i.e. 0
 * (akima_reg.F:12 epsilon(1.0d0)). */
static int g_akima_cpp_quirks = 0;
/* (3) UniformInterp.cpp:186 forms each sinc term in double and rounds the running sum to float once per tap
 * (uniform_interp.f90:424-425 multiplies and adds in real*4). */
static int g_sinc_cpp_arith = 0;
static int g_resamp_cpp_positions = 0; /* bit2, see orc_resamp_slc */
void orc_test_set_cpp_quirks(int on) { g_akima_cpp_quirks = on & 1; g_sinc_cpp_arith = (on >> 1) & 1; g_resamp_cpp_positions = (on >> 2) & 1; }
static int aki_almost_equal(double x, double y) { return fabs(x - y) <= (g_akima_cpp_quirks ? 0.0 : 2.220446049250313e-16); }
static double akima_eval(const float *dem, int nx, int ny, int ix, int iy, double fx, double fy)
{
    double sx[2][2], sy[2][2], sxy[2][2]; /* [jj][ii] */
    double wx2 = 0.0, wx3 = 0.0, wy2 = 0.0, wy3 = 0.0;
    for (int ii = 1; ii <= 2; ii++) {
        int yy = iy + ii; if (yy < 3) yy = 3; if (yy > ny - 2) yy = ny - 2;
        for (int jj = 1; jj <= 2; jj++) {
            int xx = ix + jj; if (xx < 3) xx = 3; if (xx > nx - 2) xx = nx - 2;
            float f;
            double m1, m2, m3, m4;
            f = DEM(xx - 1, yy) - DEM(xx - 2, yy); m1 = f;
            f = DEM(xx, yy) - DEM(xx - 1, yy); m2 = f;
            f = DEM(xx + 1, yy) - DEM(xx, yy); m3 = f;
            f = DEM(xx + 2, yy) - DEM(xx + 1, yy); m4 = f;
            if (aki_almost_equal(m1, m2) && aki_almost_equal(m3, m4)) sx[jj - 1][ii - 1] = 0.5 * (m2 + m3);
            else {
                wx2 = fabs(m4 - m3);
                wx3 = fabs(m2 - m1);
                sx[jj - 1][ii - 1] = (wx2 * m2 + wx3 * m3) / (wx2 + wx3);
            }
            f = DEM(xx, yy - 1) - DEM(xx, yy - 2); m1 = f;
            f = DEM(xx, yy) - DEM(xx, yy - 1); m2 = f;
            f = DEM(xx, yy + 1) - DEM(xx, yy); m3 = f;
            f = DEM(xx, yy + 2) - DEM(xx, yy + 1); m4 = f;
            if (aki_almost_equal(m1, m2) && aki_almost_equal(m3, m4)) sy[jj - 1][ii - 1] = 0.5 * (m2 + m3);
            else {
                wy2 = fabs(m4 - m3);
                wy3 = fabs(m2 - m1);
                sy[jj - 1][ii - 1] = (wy2 * m2 + wy3 * m3) / (wy2 + wy3);
            }
            if (g_akima_cpp_quirks) { sx[jj - 1][ii - 1] = sy[jj - 1][ii - 1]; sy[jj - 1][ii - 1] = 0.0; }
            /* cross derivative: m2, m3 below are the Y slopes just computed (the Fortran reuses the variables) */
            double d22, d23, d42, d43;
            f = DEM(xx - 1, yy) - DEM(xx - 1, yy - 1); d22 = f;
            f = DEM(xx - 1, yy + 1) - DEM(xx - 1, yy); d23 = f;
            f = DEM(xx + 1, yy) - DEM(xx + 1, yy - 1); d42 = f;
            f = DEM(xx + 1, yy + 1) - DEM(xx + 1, yy); d43 = f;
            double e22 = m2 - d22, e23 = m3 - d23, e32 = d42 - m2, e33 = d43 - m3;
            if (aki_almost_equal(wx2, 0.0) && aki_almost_equal(wx3, 0.0)) { wx2 = 1.; wx3 = 1.; }
            if (aki_almost_equal(wy2, 0.0) && aki_almost_equal(wy3, 0.0)) { wy2 = 1.; wy3 = 1.; }
            sxy[jj - 1][ii - 1] = (wx2 * (wy2 * e22 + wy3 * e23) + wx3 * (wy2 * e32 + wy3 * e33)) / ((wx2 + wx3) * (wy2 + wy3));
        }
    }
    double poly[17] = {0};
    double b1 = DEM(ix, iy), b2 = DEM(ix + 1, iy), b3 = DEM(ix + 1, iy + 1), b4 = DEM(ix, iy + 1);
    double b5 = sx[0][0], b6 = sx[1][0], b7 = sx[1][1], b8 = sx[0][1];
    double b9 = sy[0][0], b10 = sy[1][0], b11 = sy[1][1], b12 = sy[0][1];
    double b13 = sxy[0][0], b14 = sxy[1][0], b15 = sxy[1][1], b16 = sxy[0][1];
    poly[11] = b13; poly[12] = b5; poly[15] = b9; poly[16] = b1;
    double c1 = b1 - b2, c2 = b3 - b4, c3 = b5 + b6, c4 = b7 + b8, c5 = b9 - b10, c6 = b11 - b12, c7 = b13 + b14, c8 = b15 + b16;
    double c9 = 2 * b5 + b6, c10 = b7 + 2 * b8, c11 = 2 * b13 + b14, c12 = b15 + 2 * b16, c13 = b5 - b8, c14 = b1 - b4;
    double c15 = b13 + b16, c16 = 2 * b13 + b16, c17 = b9 + b12, c18 = 2 * b9 + b12;
    double d1 = c1 + c2, d2 = c3 - c4, d3 = c5 - c6, d4 = c7 + c8, d5 = c9 - c10, d6 = 2 * c5 - c6, d7 = 2 * c7 + c8;
    double d8 = c11 + c12, d9 = 2 * c11 + c12;
    double f1 = 2 * d1 + d2, f2 = 2 * d3 + d4, f3 = 2 * d6 + d7, f4 = 3 * d1 + d5, f5 = 3 * d3 + d8, f6 = 3 * d6 + d9;
    poly[1] = 2 * f1 + f2; poly[2] = -(3 * f1 + f3); poly[3] = 2 * c5 + c7; poly[4] = 2 * c1 + c3;
    poly[5] = -(2 * f4 + f5); poly[6] = 3 * f4 + f6; poly[7] = -(3 * c5 + c11); poly[8] = -(3 * c1 + c9);
    poly[9] = 2 * c13 + c15; poly[10] = -(3 * c13 + c16); poly[13] = 2 * c14 + c17; poly[14] = -(3 * c14 + c18);
    const double x = fx, y = fy; /* xx - ix, yy - iy */
    double p1 = ((poly[1] * y + poly[2]) * y + poly[3]) * y + poly[4];
    double p2 = ((poly[5] * y + poly[6]) * y + poly[7]) * y + poly[8];
    double p3 = ((poly[9] * y + poly[10]) * y + poly[11]) * y + poly[12];
    double p4 = ((poly[13] * y + poly[14]) * y + poly[15]) * y + poly[16];
    return ((p1 * x + p2) * x + p3) * x + p4;
}

/* topozeroMethods.f:123-247 wrappers (window checks -> BADVALUE) */
float orc_interp_dem(int method, const float *dem, int i_x, int i_y, double f_x, double f_y, int nx, int ny)
{
    double dx = i_x + f_x, dy = i_y + f_y;
    switch (method) {
    case ORC_SINC: /* :100-121 */
        if ((i_x < 4) || (i_x > (nx - 3))) return ORC_BADVALUE;
        if ((i_y < 4) || (i_y > (ny - 3))) return ORC_BADVALUE;
        return sinc_eval_2d_f(dem, sinc_table(), SINC_SUB, SINC_LEN, i_x + SINC_LEN / 2, i_y + SINC_LEN / 2, f_x, f_y, nx, ny);
    case ORC_BILINEAR: /* :123-147 */
        if ((i_x < 1) || (i_x >= nx)) return ORC_BADVALUE;
        if ((i_y < 1) || (i_y >= ny)) return ORC_BADVALUE;
        return (float)bilinear(dy, dx, dem, nx);
    case ORC_BICUBIC: /* :149-172 */
        if ((i_x < 2) || (i_x >= (nx - 1))) return ORC_BADVALUE;
        if ((i_y < 2) || (i_y >= (ny - 1))) return ORC_BADVALUE;
        return (float)bicubic(dy, dx, dem, nx);
    case ORC_BIQUINTIC: /* :174-198 */
        if ((i_x < 3) || (i_x >= (nx - 2))) return ORC_BADVALUE;
        if ((i_y < 3) || (i_y >= (ny - 2))) return ORC_BADVALUE;
        return interp2dspline(6, dem, nx, ny, dy, dx);
    case ORC_NEAREST: { /* :200-220 */
        int ix = (int)lround(i_x + f_x), iy = (int)lround(i_y + f_y);
        if ((ix < 1) || (ix > nx)) return ORC_BADVALUE;
        if ((iy < 1) || (iy > ny)) return ORC_BADVALUE;
        return DEM(ix, iy);
    }
    case ORC_AKIMA: { /* :222-247 */
        if ((i_x < 1) || (i_x >= (nx - 1))) return ORC_BADVALUE;
        if ((i_y < 1) || (i_y >= (ny - 1))) return ORC_BADVALUE;
        /* polyvalAkima(i_x, i_y, dx, dy): x = dx - i_x, y = dy - i_y */
        return (float)akima_eval(dem, nx, ny, i_x, i_y, dx - i_x, dy - i_y);
    }
    default:
        return NAN;
    }
}

/* many points at once (the pins against the reference's C++ run over > 1e6 draws) */
void orc_interp_dem_batch(int method, const float *dem, int nx, int ny, long n, const int *ix, const int *iy,
                          const double *fx, const double *fy, float *out)
{
    for (long k = 0; k < n; k++) out[k] = orc_interp_dem(method, dem, ix[k], iy[k], fx[k], fy[k], nx, ny);
}

/* ------------------------------------------------------------------ */
/* sort / search: components/zerodop/topozero/src/topozero.f90:910-963 */
/* ------------------------------------------------------------------ */
void orc_insertion_sort(double *a, double *b, double *c, int num)
{
    for (int i = 2; i <= num; i++) {
        int j = i - 1;
        double ta = a[i - 1], tb = b[i - 1], tc = c[i - 1];
        while (j >= 1 && a[j - 1] > ta) {
            a[j] = a[j - 1];
            b[j] = b[j - 1];
            c[j] = c[j - 1];
            j--;
        }
        a[j] = ta;
        b[j] = tb;
        c[j] = tc;
    }
}

/* returns the reference's 1-based index */
int orc_binarysearch(const double *array, int length, double val)
{
    int left = 1, right = length, middle;
    for (;;) {
        if (left > right) break;
        /* nint((left+right)/2.0): half rounds away from zero */
        middle = (left + right + 1) / 2;
        if (left == (right - 1)) return left;
        else if (array[middle - 1] <= val) left = middle;
        else if (array[middle - 1] > val) right = middle;
        else return left; /* NaN key: the Fortran would spin forever; bail out */
        if (left == right) return left; /* length==1 guard (Fortran would spin) */
    }
    return left;
}

/* ------------------------------------------------------------------ */
/* topo: components/zerodop/topozero/src/topozero.f90:5-907             */
/* ------------------------------------------------------------------ */
typedef struct {
    double xyzsat[3], velsat[3], vhat[3], that[3], chat[3], nhat[3];
    double llhsat[3];
    double vmag, height, rcurv;
    double mat[9], ov[3];
} line_state;

static void setup_line(const orc_topo_params *p, const orc_orbit *orb, double tline, line_state *s)
{
    /* :378-424 (stat ignored, :378) */
    interp_orbit(p->orbitmethod, orb, tline, s->xyzsat, s->velsat);
    v_unit(s->velsat, s->vhat);
    s->vmag = v_norm(s->velsat);
    orc_latlon(p->major, p->e2, s->xyzsat, s->llhsat, 2);
    s->height = s->llhsat[2];
    orc_tcnbasis(s->xyzsat, s->velsat, p->major, p->e2, s->that, s->chat, s->nhat);
    s->rcurv = orc_radar_to_xyz(p->major, p->e2, s->llhsat[0], s->llhsat[1], p->peghdg, s->mat, s->ov);
}

/* range-sphere solve :495-519 ; returns delta and xyz */
static inline void solve_xyz(const orc_topo_params *p, const line_state *s, double rng, double dopfact, double zsch,
                             double *costheta_o, double *sintheta_o, double *delta, double *xyz)
{
    double aa = s->height + s->rcurv;
    double bb = s->rcurv + zsch;
    double costheta = 0.5 * ((aa / rng) + (rng / aa) - (bb / aa) * (bb / rng));
    double sintheta = sqrt(1.0 - costheta * costheta);
    double gamm = costheta * rng;
    double alpha = (dopfact - gamm * v_dot(s->nhat, s->vhat)) / v_dot(s->vhat, s->that);
    double beta = -p->ilrl * sqrt(rng * rng * sintheta * sintheta - alpha * alpha);
    for (int i = 0; i < 3; i++) {
        delta[i] = gamm * s->nhat[i] + alpha * s->that[i] + beta * s->chat[i];
        xyz[i] = s->xyzsat[i] + delta[i];
    }
    *costheta_o = costheta;
    *sintheta_o = sintheta;
}

int orc_topo(const orc_topo_params *p, const float *dem_full, const orc_orbit *orb,
             const orc_poly2d *dopp, const orc_poly2d *slrng, const double *rho_image,
             int line0, int nlines,
             double *olat, double *olon, double *ohgt, float *olos, float *oinc, int8_t *omaskimg,
             orc_topo_result *res, int nthreads)
{
    const int width = p->width, length = p->length;
    const int owidth = 2 * width + 1; /* :134-135 */
    const double pi = 4.0 * atan(1.0); /* fortranUtils.f90:38-41 */
    const double r2d = 180.0 / pi;
    const double MIN_H = -500.0, MAX_H = 9000.0, MARGIN = 0.15; /* topozeroState.f:74-75 */
    const double hgts[2] = {MIN_H, MAX_H};
    const int method = p->method;
    double min_lat = 10000., max_lat = -10000., min_lon = 10000., max_lon = -10000.;
    long long totalconv = 0, total_iters = 0;
    int rc = 0;

    if (p->orbitmethod == ORC_LEGENDRE ? orb->nvec < 9 : orb->nvec < 4) return -2; /* :104-131 'stop' */
    if (method != ORC_BILINEAR && method != ORC_BICUBIC && method != ORC_BIQUINTIC && method != ORC_NEAREST &&
        method != ORC_SINC && method != ORC_AKIMA)
        return -3;
    if (method == ORC_SINC) sinc_table();
    if (!slrng && !rho_image) return -4;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif

    double *lat = malloc(sizeof(double) * width), *lon = malloc(sizeof(double) * width);
    double *z = malloc(sizeof(double) * width), *zsch = malloc(sizeof(double) * width);
    double *rho = malloc(sizeof(double) * width), *dopline = malloc(sizeof(double) * width);
    float *distance = malloc(sizeof(float) * width);
    float *losang = malloc(sizeof(float) * 2 * width), *incang = malloc(sizeof(float) * 2 * width);
    float *elevang = malloc(sizeof(float) * width);
    int *converge = malloc(sizeof(int) * width);
    int8_t *mask = NULL, *omask = NULL;
    double *orng = NULL, *ctrack = NULL, *oview = NULL;
    if (omaskimg) {
        omask = malloc(owidth);
        mask = malloc(width);
        orng = malloc(sizeof(double) * owidth);
        ctrack = malloc(sizeof(double) * owidth);
        oview = malloc(sizeof(double) * owidth);
    }
    float *dem = NULL;

    /* ---- bbox of interest :192-263 ---- */
    line_state st;
    for (int pixel = 1; pixel <= width; pixel++) { /* getLine(..., line=1) */
        dopline[pixel - 1] = orc_eval_poly2d(dopp, 0.0, (double)(pixel - 1));
        rho[pixel - 1] = rho_image ? rho_image[pixel - 1] : orc_eval_poly2d(slrng, 0.0, (double)(pixel - 1));
    }
    for (int line = 1; line <= 2; line++) {
        double tline = p->t0 + (line - 1) * p->nazlooks * (length - 1.0) / p->prf;
        double xs[3], vs[3];
        int stat = interp_orbit(p->orbitmethod, orb, tline, xs, vs);
        if (stat != 0) break; /* :204-207 prints and exits the loop */
        setup_line(p, orb, tline, &st);
        for (int ind = 1; ind <= 2; ind++) {
            int pixel = (ind - 1) * (width - 1) + 1;
            double rng = rho[pixel - 1];
            double dopfact = (0.5 * p->wvl * dopline[pixel - 1] / st.vmag) * rng;
            for (int iter = 0; iter < 2; iter++) {
                double llh[3];
                if (rng <= (st.llhsat[2] - hgts[iter] + 1.0)) {
                    llh[0] = st.llhsat[0]; llh[1] = st.llhsat[1]; llh[2] = st.llhsat[2];
                } else {
                    double ct, sn, delta[3], xyz[3];
                    solve_xyz(p, &st, rng, dopfact, hgts[iter], &ct, &sn, delta, xyz);
                    orc_latlon(p->major, p->e2, xyz, llh, 2);
                }
                min_lat = fmin(min_lat, llh[0] * r2d);
                max_lat = fmax(max_lat, llh[0] * r2d);
                min_lon = fmin(min_lon, llh[1] * r2d);
                max_lon = fmax(max_lon, llh[1] * r2d);
            }
        }
    }
    min_lon -= MARGIN; max_lon += MARGIN; min_lat -= MARGIN; max_lat += MARGIN;

    /* ---- usable part of the DEM :281-320 ---- */
    double umin_lon = fmax(min_lon, p->firstlon);
    double umax_lon = fmin(max_lon, p->firstlon + (p->idemwidth - 1) * p->deltalon);
    double umax_lat = fmin(max_lat, p->firstlat);
    double umin_lat = fmax(min_lat, p->firstlat + (p->idemlength - 1) * p->deltalat);
    int ustartx = (int)((umin_lon - p->firstlon) / p->deltalon) + 1;
    if (ustartx < 1) ustartx = 1;
    int uendx = (int)((umax_lon - p->firstlon) / p->deltalon + 0.5) + 1;
    if (uendx > p->idemwidth) uendx = p->idemwidth;
    int ustarty = (int)((umax_lat - p->firstlat) / p->deltalat) + 1;
    if (ustarty < 1) ustarty = 1;
    int uendy = (int)((umin_lat - p->firstlat) / p->deltalat + 0.5) + 1;
    if (uendy > p->idemlength) ustarty = p->idemlength; /* :314 typo kept */
    const double ufirstlon = p->firstlon + p->deltalon * (ustartx - 1);
    const double ufirstlat = p->firstlat + p->deltalat * (ustarty - 1);
    const int udemwidth = uendx - ustartx + 1;
    const int udemlength = uendy - ustarty + 1;
    if (udemwidth < 2 || udemlength < 2 || uendy > p->idemlength) { rc = -5; goto done; }

    /* ---- crop :333-345 ---- */
    dem = malloc(sizeof(float) * (size_t)udemwidth * (size_t)udemlength);
    float demmax = -INFINITY;
    for (int j = 1; j <= udemlength; j++) {
        const float *src = dem_full + (size_t)(j + ustarty - 2) * (size_t)p->idemwidth + (size_t)(ustartx - 1);
        float *dst = dem + (size_t)(j - 1) * (size_t)udemwidth;
        memcpy(dst, src, sizeof(float) * (size_t)udemwidth);
        for (int i = 0; i < udemwidth; i++)
            if (dst[i] > demmax) demmax = dst[i];
    }
    const int nx = udemwidth, ny = udemlength;
    const float fny1 = (float)(udemlength - 1), fnx1 = (float)(udemwidth - 1);

    min_lat = 10000.; max_lat = -10000.; min_lon = 10000.; max_lon = -10000.; /* :357-360 */

    if (line0 < 0) line0 = 0;
    if (nlines < 0 || line0 + nlines > length) nlines = length - line0;

    for (int line = line0 + 1; line <= line0 + nlines; line++) { /* :365 */
        const double tline = p->t0 + p->nazlooks * (line - 1.0) / p->prf; /* :371 */
        setup_line(p, orb, tline, &st);
        /* :406,414 sequential line reads of the polynomial accessors: row = line-1, col = pixel-1 */
        for (int pixel = 1; pixel <= width; pixel++) {
            dopline[pixel - 1] = orc_eval_poly2d(dopp, (double)(line - 1), (double)(pixel - 1));
            rho[pixel - 1] = rho_image ? rho_image[(size_t)(line - 1) * width + pixel - 1]
                                       : orc_eval_poly2d(slrng, (double)(line - 1), (double)(pixel - 1));
        }
        for (int i = 0; i < width; i++) { /* :425-436 */
            converge[i] = 0;
            z[i] = 0.;
            zsch[i] = 0.;
            lat[i] = ufirstlat + 0.5 * p->deltalat * udemlength;
            lon[i] = ufirstlon + 0.05 * p->deltalon * udemwidth;
        }

        const int niter = p->numiter + p->extraiter + 1;
        for (int iter = 1; iter <= niter; iter++) { /* :458 */
            long long conv_here = 0, iters_here = 0;
#pragma omp parallel for schedule(static) reduction(+ : conv_here, iters_here)
            for (int pixel = 1; pixel <= width; pixel++) { /* :472-596 */
                const int k = pixel - 1;
                double rng = rho[k];
                double dopfact = (0.5 * p->wvl * dopline[k] / st.vmag) * rng;
                if (converge[k] == 0) {
                    double llh_prev[3], xyz_prev[3], llh[3], xyz[3], delta[3], sch[3], ct, sn;
                    iters_here++;
                    llh_prev[0] = lat[k] / r2d;
                    llh_prev[1] = lon[k] / r2d;
                    llh_prev[2] = z[k];
                    solve_xyz(p, &st, rng, dopfact, zsch[k], &ct, &sn, delta, xyz);
                    orc_latlon(p->major, p->e2, xyz, llh, 2);
                    lat[k] = llh[0] * r2d;
                    lon[k] = llh[1] * r2d;
                    float demlat = (float)((lat[k] - ufirstlat) / p->deltalat + 1);
                    float demlon = (float)((lon[k] - ufirstlon) / p->deltalon + 1);
                    if (demlat < 1) demlat = 1;
                    if (demlat > fny1) demlat = fny1;
                    if (demlon < 1) demlon = 1;
                    if (demlon > fnx1) demlon = fnx1;
                    int idemlat = (int)demlat, idemlon = (int)demlon;
                    double fraclat = (double)(float)(demlat - (float)idemlat);
                    double fraclon = (double)(float)(demlon - (float)idemlon);
                    z[k] = orc_interp_dem(method, dem, idemlon, idemlat, fraclon, fraclat, nx, ny);
                    if (z[k] < -500.0) z[k] = -500.0;
                    llh[0] = lat[k] / r2d;
                    llh[1] = lon[k] / r2d;
                    llh[2] = z[k];
                    orc_latlon(p->major, p->e2, xyz, llh, 1);
                    orc_xyz_to_sch(st.mat, st.ov, st.rcurv, xyz, sch);
                    zsch[k] = sch[2];
                    distance[k] = (float)(sqrt((xyz[0] - st.xyzsat[0]) * (xyz[0] - st.xyzsat[0]) +
                                               (xyz[1] - st.xyzsat[1]) * (xyz[1] - st.xyzsat[1]) +
                                               (xyz[2] - st.xyzsat[2]) * (xyz[2] - st.xyzsat[2])) - rng);
                    if (fabs((double)distance[k]) <= p->thresh) {
                        zsch[k] = sch[2];
                        converge[k] = 1;
                        conv_here++;
                    } else if (iter > (p->numiter + 1)) { /* :572-593 */
                        orc_latlon(p->major, p->e2, xyz_prev, llh_prev, 1);
                        xyz[0] = 0.5 * (xyz_prev[0] + xyz[0]);
                        xyz[1] = 0.5 * (xyz_prev[1] + xyz[1]);
                        xyz[2] = 0.5 * (xyz_prev[2] + xyz[2]);
                        orc_latlon(p->major, p->e2, xyz, llh, 2);
                        lat[k] = llh[0] * r2d;
                        lon[k] = llh[1] * r2d;
                        z[k] = llh[2];
                        orc_xyz_to_sch(st.mat, st.ov, st.rcurv, xyz, sch);
                        zsch[k] = sch[2];
                        distance[k] = (float)(sqrt((xyz[0] - st.xyzsat[0]) * (xyz[0] - st.xyzsat[0]) +
                                                   (xyz[1] - st.xyzsat[1]) * (xyz[1] - st.xyzsat[1]) +
                                                   (xyz[2] - st.xyzsat[2]) * (xyz[2] - st.xyzsat[2])) - rng);
                    }
                }
            }
            totalconv += conv_here;
            total_iters += iters_here;
        }

        /* ---- final computation :607-708 ---- */
#pragma omp parallel for schedule(static)
        for (int pixel = 1; pixel <= width; pixel++) {
            const int k = pixel - 1;
            double rng = rho[k];
            double dopfact = (0.5 * p->wvl * dopline[k] / st.vmag) * rng;
            double llh[3], xyz[3], delta[3], costheta, sintheta;
            solve_xyz(p, &st, rng, dopfact, zsch[k], &costheta, &sintheta, delta, xyz);
            orc_latlon(p->major, p->e2, xyz, llh, 2);
            lat[k] = llh[0] * r2d;
            lon[k] = llh[1] * r2d;
            z[k] = llh[2];
            distance[k] = (float)(sqrt((xyz[0] - st.xyzsat[0]) * (xyz[0] - st.xyzsat[0]) +
                                       (xyz[1] - st.xyzsat[1]) * (xyz[1] - st.xyzsat[1]) +
                                       (xyz[2] - st.xyzsat[2]) * (xyz[2] - st.xyzsat[2])) - rng);
            double enumat[9], enu[3];
            orc_enubasis(llh[0], llh[1], enumat);
            /* xyz2enu = transpose(enumat); enu = matmul(xyz2enu, delta) */
            for (int i = 0; i < 3; i++)
                enu[i] = enumat[0 * 3 + i] * delta[0] + enumat[1 * 3 + i] * delta[1] + enumat[2 * 3 + i] * delta[2];
            double cosalpha = fabs(enu[2]) / v_norm(enu);
            losang[2 * k] = (float)(acos(cosalpha) * r2d);
            losang[2 * k + 1] = (float)((atan2(-enu[1], -enu[0]) - 0.5 * pi) * r2d);
            elevang[k] = (float)(acos(costheta) * r2d);
            zsch[k] = rng * sintheta; /* ctrack stored in zsch :662 */

            float demlat = (float)((lat[k] - ufirstlat) / p->deltalat + 1);
            float demlon = (float)((lon[k] - ufirstlon) / p->deltalon + 1);
            if (demlat < 2) demlat = 2;
            if (demlat > fny1) demlat = fny1;
            if (demlon < 2) demlon = 2;
            if (demlon > fnx1) demlon = fnx1;
            int idemlat = (int)demlat, idemlon = (int)demlon;
            double fraclat = (double)(float)(demlat - (float)idemlat);
            double fraclon = (double)(float)(demlon - (float)idemlon);
            double aa = orc_interp_dem(method, dem, idemlon - 1, idemlat, fraclon, fraclat, nx, ny);
            double bb = orc_interp_dem(method, dem, idemlon + 1, idemlat, fraclon, fraclat, nx, ny);
            double gamm = lat[k] / r2d;
            double alpha = (bb - aa) * r2d / (2.0 * orc_reast(p->major, p->e2, gamm) * p->deltalon);
            aa = orc_interp_dem(method, dem, idemlon, idemlat - 1, fraclon, fraclat, nx, ny);
            bb = orc_interp_dem(method, dem, idemlon, idemlat + 1, fraclon, fraclat, nx, ny);
            double beta = (bb - aa) * r2d / (2.0 * orc_rnorth(p->major, p->e2, gamm) * p->deltalat);
            double en = v_norm(enu);
            enu[0] = enu[0] / en; enu[1] = enu[1] / en; enu[2] = enu[2] / en;
            costheta = (enu[0] * alpha + enu[1] * beta - enu[2]) / sqrt(1.0 + alpha * alpha + beta * beta);
            incang[2 * k + 1] = (float)(acos(costheta) * r2d);
            double n_img[3], n_img_enu[3], n_trg_enu[3], tmp[3];
            v_cross(delta, st.velsat, n_img);
            v_unit(n_img, n_img);
            for (int i = 0; i < 3; i++) tmp[i] = -p->ilrl * n_img[i];
            for (int i = 0; i < 3; i++)
                n_img_enu[i] = enumat[0 * 3 + i] * tmp[0] + enumat[1 * 3 + i] * tmp[1] + enumat[2 * 3 + i] * tmp[2];
            n_trg_enu[0] = -alpha; n_trg_enu[1] = -beta; n_trg_enu[2] = 1.0;
            double cospsi = v_dot(n_trg_enu, n_img_enu) / (v_norm(n_trg_enu) * v_norm(n_img_enu));
            incang[2 * k] = (float)(acos(cospsi) * r2d);
        }

        /* :712-726 */
        for (int i = 0; i < width; i++) {
            if (lat[i] < min_lat) min_lat = lat[i];
            if (lat[i] > max_lat) max_lat = lat[i];
            if (lon[i] < min_lon) min_lon = lon[i];
            if (lon[i] > max_lon) max_lon = lon[i];
        }
        const size_t orow = (size_t)(line - 1 - line0);
        memcpy(olat + orow * width, lat, sizeof(double) * width);
        memcpy(olon + orow * width, lon, sizeof(double) * width);
        memcpy(ohgt + orow * width, z, sizeof(double) * width);
        if (olos)
            for (int i = 0; i < width; i++) { /* BIL on disk: BILAccessor.cpp:11-37 */
                olos[orow * 2 * width + i] = losang[2 * i];
                olos[orow * 2 * width + width + i] = losang[2 * i + 1];
            }
        if (oinc)
            for (int i = 0; i < width; i++) {
                oinc[orow * 2 * width + i] = incang[2 * i];
                oinc[orow * 2 * width + width + i] = incang[2 * i + 1];
            }

        /* ---- layover / shadow mask :729-880 ---- */
        if (omaskimg) {
            double cmin = zsch[0], cmax = zsch[0];
            for (int i = 1; i < width; i++) {
                if (zsch[i] < cmin) cmin = zsch[i];
                if (zsch[i] > cmax) cmax = zsch[i];
            }
            const double ctrackmin = cmin - demmax, ctrackmax = cmax + demmax;
            const double dctrack = (ctrackmax - ctrackmin) / (owidth - 1.0);
            orc_insertion_sort(zsch, lat, lon, width); /* :735 */
#pragma omp parallel for schedule(static)
            for (int pixel = 1; pixel <= owidth; pixel++) { /* :745-782 */
                double aa = ctrackmin + (pixel - 1) * dctrack;
                ctrack[pixel - 1] = aa;
                int it = orc_binarysearch(zsch, width, aa);
                if (it == width) it = width - 1;
                if (it == 0) it = 1;
                double fraclat = (aa - zsch[it - 1]) / (zsch[it] - zsch[it - 1]);
                float demlat = (float)(lat[it - 1] + fraclat * (lat[it] - lat[it - 1])); /* real*4 :755 */
                float demlon = (float)(lon[it - 1] + fraclat * (lon[it] - lon[it - 1]));
                double llh[3], xyz[3];
                llh[0] = demlat / r2d;
                llh[1] = demlon / r2d;
                demlat = (float)((demlat - ufirstlat) / p->deltalat + 1);
                demlon = (float)((demlon - ufirstlon) / p->deltalon + 1);
                if (demlat < 2) demlat = 2;
                if (demlat > fny1) demlat = fny1;
                if (demlon < 2) demlon = 2;
                if (demlon > fnx1) demlon = fnx1;
                int idemlat = (int)demlat, idemlon = (int)demlon;
                fraclat = (double)(float)(demlat - (float)idemlat);
                double fraclon = (double)(float)(demlon - (float)idemlon);
                llh[2] = orc_interp_dem(method, dem, idemlon, idemlat, fraclon, fraclat, nx, ny);
                orc_latlon(p->major, p->e2, xyz, llh, 1);
                xyz[0] = xyz[0] - st.xyzsat[0]; xyz[1] = xyz[1] - st.xyzsat[1]; xyz[2] = xyz[2] - st.xyzsat[2];
                double bb = v_norm(xyz);
                orng[pixel - 1] = bb;
                aa = fabs(st.nhat[0] * xyz[0] + st.nhat[1] * xyz[1] + st.nhat[2] * xyz[2]);
                oview[pixel - 1] = acos(aa / bb) * r2d;
            }
            orc_insertion_sort(orng, ctrack, oview, owidth); /* :787 */
            memset(mask, 0, width);
            memset(omask, 0, owidth);
            double aa = elevang[0]; /* :791-809 shadow */
            for (int pixel = 2; pixel <= width; pixel++) {
                double bb = elevang[pixel - 1];
                if (bb <= aa) mask[pixel - 1] = 1; else aa = bb;
            }
            aa = elevang[width - 1];
            for (int pixel = width - 1; pixel >= 1; pixel--) {
                double bb = elevang[pixel - 1];
                if (bb >= aa) mask[pixel - 1] = 1; else aa = bb;
            }
            aa = ctrack[0]; /* :834-852 layover; forward loop bound is width (not owidth) as in the reference */
            for (int pixel = 2; pixel <= width; pixel++) {
                double bb = ctrack[pixel - 1];
                if ((bb <= aa) && (omask[pixel - 1] < 2)) omask[pixel - 1] = omask[pixel - 1] + 2; else aa = bb;
            }
            aa = ctrack[owidth - 1];
            for (int pixel = owidth - 1; pixel >= 1; pixel--) {
                double bb = ctrack[pixel - 1];
                if ((bb >= aa) && (omask[pixel - 1] < 2)) omask[pixel - 1] = omask[pixel - 1] + 2; else aa = bb;
            }
            for (int pixel = 1; pixel <= owidth; pixel++) { /* :855-865 */
                if (omask[pixel - 1] > 0) {
                    int id = orc_binarysearch(rho, width, orng[pixel - 1]);
                    if ((id >= 1) && (id <= width))
                        if (mask[id - 1] < omask[pixel - 1]) mask[id - 1] = mask[id - 1] + omask[pixel - 1];
                }
            }
            memcpy(omaskimg + orow * width, mask, width);
        }
    }

    if (res) {
        res->min_lat = min_lat; res->max_lat = max_lat; res->min_lon = min_lon; res->max_lon = max_lon;
        res->totalconv = totalconv; res->total_iters = total_iters;
        res->ustartx = ustartx; res->ustarty = ustarty; res->udemwidth = udemwidth; res->udemlength = udemlength;
        res->ufirstlat = ufirstlat; res->ufirstlon = ufirstlon; res->demmax = demmax;
    }
done:
    free(lat); free(lon); free(z); free(zsch); free(rho); free(dopline); free(distance);
    free(losang); free(incang); free(elevang); free(converge);
    free(mask); free(omask); free(orng); free(ctrack); free(oview); free(dem);
    return rc;
}

/* ------------------------------------------------------------------ */
/* geo2rdr: components/zerodop/geo2rdr/src/geo2rdr.f90:1-423            */
/* ------------------------------------------------------------------ */
int orc_geo2rdr(const orc_geo_params *p, const double *latimg, const double *lonimg, const double *hgtimg,
                const orc_orbit *orb, const orc_poly1d *dopAcc,
                int line0, int nlines,
                double *oazt, double *orgm, double *oazoff, double *orgoff,
                orc_geo_result *res, int nthreads)
{
    const double pi = 4.0 * atan(1.0);
    const double sol = 299792458.0; /* fortranUtils.f90:43-46 */
    const double deg2rad = pi / 180.0;
    const float BAD_VALUE = -999999.0f; /* :59-60 */
    const int demwidth = p->demwidth;

    if (p->orbitmethod == ORC_LEGENDRE ? orb->nvec < 9 : orb->nvec < 4) return -2;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    /* :118-133 */
    const double tstart = p->t0;
    const double dtaz = p->nazlooks / p->prf;
    const double tend = p->t0 + (p->length - 1) * dtaz;
    const double tmid = 0.5 * (tstart + tend);
    const double rngstart = p->rho0;
    const double dmrg = p->nrnglooks * p->drho;
    const double rngend = p->rho0 + (p->width - 1) * dmrg;

    /* doppler polynomials :161-189 */
    double fd_c[64], fdd_c[64];
    if (dopAcc->order > 62) return -6;
    orc_poly1d fdvsrng = {dopAcc->order, p->rho0 + dopAcc->mean * p->drho, dopAcc->norm * p->drho, fd_c};
    for (int k = 1; k <= dopAcc->order + 1; k++) {
        double temp = dopAcc->coeffs[k - 1];
        temp = temp * p->prf;
        fd_c[k - 1] = temp;
    }
    orc_poly1d fddotvsrng;
    fddotvsrng.coeffs = fdd_c;
    if (fdvsrng.order == 0) {
        fddotvsrng.order = 0;
        fddotvsrng.mean = 0.0; /* initPoly1D leaves mean/norm unset; 0-order eval ignores xval */
        fddotvsrng.norm = 1.0;
        fdd_c[0] = 0.0;
    } else {
        fddotvsrng.order = fdvsrng.order - 1;
        fddotvsrng.mean = fdvsrng.mean;
        fddotvsrng.norm = fdvsrng.norm;
        for (int k = 1; k <= dopAcc->order; k++) {
            double temp = fd_c[k];
            temp = k * temp / fdvsrng.norm;
            fdd_c[k - 1] = temp;
        }
    }

    /* :194-208 */
    double xyz_mid[3], vel_mid[3], acc_mid[3];
    if (interp_orbit(p->orbitmethod, orb, tmid, xyz_mid, vel_mid) != 0) return -7;
    if (orc_compute_acceleration(orb, tmid, acc_mid) != 0) return -8;

    if (line0 < 0) line0 = 0;
    if (nlines < 0 || line0 + nlines > p->demlength) nlines = p->demlength - line0;

    long long numOutside = 0, cnt = 0, conv = 0, total_iters = 0;
    for (int line = line0 + 1; line <= line0 + nlines; line++) { /* :215 */
        const size_t row = (size_t)(line - 1) * demwidth;
        const size_t orow = (size_t)(line - 1 - line0) * demwidth;
#pragma omp parallel for schedule(static) reduction(+ : numOutside, cnt, conv, total_iters)
        for (int pixel = 1; pixel <= demwidth; pixel++) {
            double azt = BAD_VALUE, rgm = BAD_VALUE, rgoff = BAD_VALUE, azoff = BAD_VALUE; /* :217-221 */
            double llh[3], xyz[3], satx[3], satv[3], sata[3], dr[3];
            double tline, tprev = 0, rngpix = 0;
            int outside = 0;
            llh[0] = latimg[row + pixel - 1] * deg2rad;
            llh[1] = lonimg[row + pixel - 1] * deg2rad;
            llh[2] = hgtimg[row + pixel - 1];
            orc_latlon(p->major, p->e2, xyz, llh, 1);
            tline = tmid;
            for (int i = 0; i < 3; i++) { satx[i] = xyz_mid[i]; satv[i] = vel_mid[i]; sata[i] = acc_mid[i]; }
            for (int k = 1; k <= 51; k++) { /* :259-305 */
                total_iters++;
                tprev = tline;
                for (int i = 0; i < 3; i++) dr[i] = xyz[i] - satx[i];
                rngpix = v_norm(dr);
                double dopfact = v_dot(dr, satv);
                double fdop = 0.5 * p->wvl * orc_eval_poly1d(&fdvsrng, rngpix);
                double fdopder = 0.5 * p->wvl * orc_eval_poly1d(&fddotvsrng, rngpix);
                double fn = dopfact - fdop * rngpix;
                double c1 = (0.0 * v_dot(sata, dr) - v_dot(satv, satv));
                double c2 = (fdop / rngpix + fdopder);
                double fnprime = c1 + c2 * dopfact;
                tline = tline - fn / fnprime;
                int stat = interp_orbit(p->orbitmethod, orb, tline, satx, satv);
                if (stat != 0) {
                    tline = BAD_VALUE;
                    rngpix = BAD_VALUE;
                    break;
                }
                if (fabs(tline - tprev) < 5.0e-9) {
                    conv++;
                    break;
                }
            }
            if (tline < tstart) outside = 1;
            else if (tline > tend) outside = 1;
            else {
                for (int i = 0; i < 3; i++) dr[i] = xyz[i] - satx[i];
                rngpix = v_norm(dr);
                if (rngpix < rngstart) outside = 1;
                else if (rngpix > rngend) outside = 1;
                else if (p->bistatic) { /* :331-368 */
                    tline = tline + 2.0 * rngpix / sol;
                    if (tline < tstart) outside = 1;
                    else if (tline > tend) outside = 1;
                    else {
                        int stat = interp_orbit(p->orbitmethod, orb, tline, satx, satv);
                        if (stat != 0) { tline = BAD_VALUE; rngpix = BAD_VALUE; }
                        if (tline == BAD_VALUE) outside = 1;
                        else {
                            for (int i = 0; i < 3; i++) dr[i] = xyz[i] - satx[i];
                            rngpix = v_norm(dr);
                            if (rngpix < rngstart) outside = 1;
                            else if (rngpix > rngend) outside = 1;
                        }
                    }
                }
            }
            if (outside) numOutside++;
            else { /* :370-376 */
                cnt++;
                rgm = rngpix;
                azt = tline;
                rgoff = ((rngpix - rngstart) / dmrg) - 1.0 * (pixel - 1);
                azoff = ((tline - tstart) / dtaz) - 1.0 * (line - 1);
            }
            if (oazt) oazt[orow + pixel - 1] = azt;
            if (orgm) orgm[orow + pixel - 1] = rgm;
            if (oazoff) oazoff[orow + pixel - 1] = azoff;
            if (orgoff) orgoff[orow + pixel - 1] = rgoff;
        }
    }
    if (res) {
        res->num_outside = numOutside; res->num_valid = cnt; res->num_conv = conv; res->total_iters = total_iters;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* geozero: components/zerodop/geozero/src/geozero.f90:1-435            */
/* ------------------------------------------------------------------ */
/* Fortran default COMPLEX (2 x real*4) and its promotion to COMPLEX*16 in mixed expressions with real*8.
 * complex * real and complex / real scale both components (what gfortran emits for a real operand). */
typedef struct { float re, im; } cx4;
typedef struct { double re, im; } cx8;
static inline cx8 c8(cx4 a) { cx8 r = {a.re, a.im}; return r; }
static inline cx4 c4(cx8 a) { cx4 r = {(float)a.re, (float)a.im}; return r; }
static inline cx8 c8_scale(cx8 a, double s) { cx8 r = {a.re * s, a.im * s}; return r; }
static inline cx8 c8_div(cx8 a, double s) { cx8 r = {a.re / s, a.im / s}; return r; }
static inline cx8 c8_add(cx8 a, cx8 b) { cx8 r = {a.re + b.re, a.im + b.im}; return r; }
static inline cx4 c4_sub(cx4 a, cx4 b) { cx4 r = {a.re - b.re, a.im - b.im}; return r; }
static inline cx4 c4_add(cx4 a, cx4 b) { cx4 r = {a.re + b.re, a.im + b.im}; return r; }

/* ifg(width,length): IFG(r, a) = sample r (1-based, range) of line a (1-based, azimuth) */
#define IFG(r, a) ifg[(size_t)((a) - 1) * (size_t)width + (size_t)((r) - 1)]

/* uniform_interp.f90:46-77 as called from geozeroMethods.F:103-117: bilinear_cx(dy, dx, ifg) */
static cx4 gz_bilinear_cx(double x, double y, const cx4 *ifg, int width)
{
    double x1 = floor(x), x2 = ceil(x), y1 = ceil(y), y2 = floor(y);
    cx4 q11 = IFG((int)y1, (int)x1), q12 = IFG((int)y2, (int)x1), q21 = IFG((int)y1, (int)x2), q22 = IFG((int)y2, (int)x2);
    if (y1 == y2 && x1 == x2) return q11;
    if (y1 == y2)
        return c4(c8_add(c8_scale(c8(q11), (x2 - x) / (x2 - x1)), c8_scale(c8(q21), (x - x1) / (x2 - x1))));
    if (x1 == x2)
        return c4(c8_add(c8_scale(c8(q11), (y2 - y) / (y2 - y1)), c8_scale(c8(q12), (y - y1) / (y2 - y1))));
    const double den = (x2 - x1) * (y2 - y1);
    cx8 s = c8_div(c8_scale(c8_scale(c8(q11), (x2 - x)), (y2 - y)), den);
    s = c8_add(s, c8_div(c8_scale(c8_scale(c8(q21), (x - x1)), (y2 - y)), den));
    s = c8_add(s, c8_div(c8_scale(c8_scale(c8(q12), (x2 - x)), (y - y1)), den));
    s = c8_add(s, c8_div(c8_scale(c8_scale(c8(q22), (x - x1)), (y - y1)), den));
    return c4(s);
}

/* uniform_interp.f90:203-292 as called from geozeroMethods.F:119-130: bicubic_cx(dy, dx, ifg); all locals are default
 * COMPLEX, so every assignment rounds to real*4 components; the dzdy column typo (:243-245) is kept */
static cx4 gz_bicubic_cx(double x, double y, const cx4 *ifg, int width)
{
    const int x1 = (int)floor(x), x2 = (int)ceil(x), y1 = (int)floor(y), y2 = (int)ceil(y);
    cx4 zz[4], dzdx[4], dzdy[4], dzdxy[4], q[16], cl[16], c[4][4];
    zz[0] = IFG(y1, x1);
    zz[3] = IFG(y2, x1);
    zz[1] = IFG(y1, x2);
    zz[2] = IFG(y2, x2);
#define HALF(a) c4(c8_div(c8(a), 2.0))
    dzdx[0] = HALF(c4_sub(IFG(y1, x1 + 1), IFG(y1, x1 - 1)));
    dzdx[1] = HALF(c4_sub(IFG(y1, x2 + 1), IFG(y1, x2 - 1)));
    dzdx[2] = HALF(c4_sub(IFG(y2, x2 + 1), IFG(y2, x2 - 1)));
    dzdx[3] = HALF(c4_sub(IFG(y2, x1 + 1), IFG(y2, x1 - 1)));
    dzdy[0] = HALF(c4_sub(IFG(y1 + 1, x1), IFG(y1 - 1, x1)));
    dzdy[1] = HALF(c4_sub(IFG(y1 + 1, x2 + 1), IFG(y1 - 1, x2)));
    dzdy[2] = HALF(c4_sub(IFG(y2 + 1, x2 + 1), IFG(y2 - 1, x2)));
    dzdy[3] = HALF(c4_sub(IFG(y2 + 1, x1 + 1), IFG(y2 - 1, x1)));
#undef HALF
#define CROSS(yy, xx) c4(c8_scale(c8(c4_add(c4_sub(c4_sub(IFG((yy) + 1, (xx) + 1), IFG((yy) - 1, (xx) + 1)), IFG((yy) + 1, (xx) - 1)), \
                                            IFG((yy) - 1, (xx) - 1))), 0.25))
    dzdxy[0] = CROSS(y1, x1);
    dzdxy[3] = CROSS(y2, x1);
    dzdxy[1] = CROSS(y1, x2);
    dzdxy[2] = CROSS(y2, x2);
#undef CROSS
    for (int i = 0; i < 4; i++) {
        q[i] = zz[i];
        q[i + 4] = dzdx[i];
        q[i + 8] = dzdy[i];
        q[i + 12] = dzdxy[i];
    }
    for (int i = 0; i < 16; i++) {
        cx4 qq = {0.f, 0.f};
        for (int k = 0; k < 16; k++) qq = c4(c8_add(c8(qq), c8_scale(c8(q[k]), WT_FLAT[k * 16 + i])));
        cl[i] = qq;
    }
    int l = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) c[i][j] = cl[l++];
    const double t = (x - x1), u = (y - y1);
    cx4 r = {0.f, 0.f};
    for (int i = 3; i >= 0; i--) {
        /* bicubic_cx = t*bicubic_cx+((c(i,4)*u+c(i,3))*u+c(i,2))*u+c(i,1): (t*r + (...)*u) + c(i,1) */
        cx8 h = c8_add(c8_scale(c8(c[i][3]), u), c8(c[i][2]));
        h = c8_scale(c8_add(c8_scale(h, u), c8(c[i][1])), u);
        r = c4(c8_add(c8_add(c8_scale(c8(r), t), h), c8(c[i][0])));
    }
    return r;
}

/* uniform_interp.f90:456-484 through geozeroMethods.F:91-101 (intp_sinc: i_xx = i_x - 1, i_yy = i_y - 1) */
static cx4 gz_sinc_cx(const cx4 *ifg, const float *intarr, int intpx, int intpy, double frpx, double frpy, int width, int length)
{
    const int idec = SINC_SUB, ilen = SINC_LEN;
    cx4 acc = {0.f, 0.f};
    if ((intpx >= ilen - 1 && intpx < width) && (intpy >= ilen - 1 && intpy < length)) {
        int ifracx = (int)(frpx * idec), ifracy = (int)(frpy * idec);
        ifracx = ifracx < 0 ? 0 : (ifracx > idec - 1 ? idec - 1 : ifracx);
        ifracy = ifracy < 0 ? 0 : (ifracy > idec - 1 ? idec - 1 : ifracy);
        double fweightsum = 0.0;
        for (int k = 0; k < ilen; k++)
            for (int m = 0; m < ilen; m++) {
                const float fw4 = intarr[k + ifracx * ilen] * intarr[m + ifracy * ilen];
                const double fweight = fw4;
                /* arrin is 0-based: arrin(a,b) = ifg(a+1,b+1) */
                acc = c4(c8_add(c8(acc), c8_scale(c8(IFG(intpx - k + 1, intpy - m + 1)), fweight)));
                fweightsum = fweightsum + fweight;
            }
        acc = c4(c8_div(c8(acc), fweightsum));
    }
    return acc;
}

void orc_geozero_grid(const orc_geozero_params *p, int *geo_len, int *geo_wid, int *min_lat_idx, int *max_lat_idx,
                      int *min_lon_idx, int *max_lon_idx)
{
    const double pi = 4.0 * atan(1.0);
    const double deg2rad = pi / 180.0;
    /* :146-149, :163-170: real*8 -> integer assignment truncates */
    const double dlonr = p->dlon * deg2rad, dlatr = p->dlat * deg2rad;
    const double lon_firstr = p->lon_first * deg2rad, lat_firstr = p->lat_first * deg2rad;
    const double min_latr = p->min_lat * deg2rad, max_latr = p->max_lat * deg2rad;
    const double min_lonr = p->min_lon * deg2rad, max_lonr = p->max_lon * deg2rad;
    *min_lat_idx = (int)((min_latr - lat_firstr) / dlatr + 1);
    *min_lon_idx = (int)((min_lonr - lon_firstr) / dlonr);
    *max_lat_idx = (int)((max_latr - lat_firstr) / dlatr);
    *max_lon_idx = (int)((max_lonr - lon_firstr) / dlonr + 1);
    *geo_len = *min_lat_idx - *max_lat_idx;
    *geo_wid = *max_lon_idx - *min_lon_idx;
}

int orc_geozero(const orc_geozero_params *p, const float *dem_full, const orc_orbit *orb, const orc_poly1d *dopAcc,
                const float *in, int iscomplex, int method, int lookSide, float *out, int16_t *dem_crop,
                double *oaz, double *orng, orc_geozero_result *res, int nthreads)
{
    const double pi = 4.0 * atan(1.0);
    const double deg2rad = pi / 180.0;
    const double BAD_VALUE = -10000.0; /* :70 */
    const int width = p->width, length = p->length, demwidth = p->demwidth, demlength = p->demlength;
    float f_delay; /* geozeroMethods.F:66-79 */
    if (method == ORC_SINC) f_delay = SINC_LEN / 2.0f;
    else if (method == ORC_BILINEAR) f_delay = 2.0f;
    else if (method == ORC_BICUBIC) f_delay = 3.0f;
    else if (method == ORC_NEAREST) f_delay = 2.0f;
    else return -1;
    if (orb->nvec < 4) return -2;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const float *fintp = method == ORC_SINC ? sinc_table() : NULL;
    /* :126-139 */
    const double tstart = p->t0;
    const double dtaz = p->nazlooks / p->prf;
    const double tend = p->t0 + (length - 1) * dtaz;
    const double tmid = 0.5 * (tstart + tend);
    const double rngstart = p->rho0;
    const double dmrg = p->nrnglooks * p->drho;
    const double dlonr = p->dlon * deg2rad, dlatr = p->dlat * deg2rad;
    const double lon_firstr = p->lon_first * deg2rad, lat_firstr = p->lat_first * deg2rad;
    int geo_len, geo_wid, min_lat_idx, max_lat_idx, min_lon_idx, max_lon_idx;
    orc_geozero_grid(p, &geo_len, &geo_wid, &min_lat_idx, &max_lat_idx, &min_lon_idx, &max_lon_idx);
    if (geo_len <= 0 || geo_wid <= 0) return -3;

    /* the band as COMPLEX ifg(width,length) (geozeroReadWrite.F:50-68) */
    cx4 *ifg = malloc(sizeof(cx4) * (size_t)width * (size_t)length);
    if (!ifg) return -4;
    for (size_t i = 0; i < (size_t)width * (size_t)length; i++) {
        ifg[i].re = iscomplex ? in[2 * i] : in[i];
        ifg[i].im = iscomplex ? in[2 * i + 1] : 0.f;
    }

    /* doppler polynomials :196-224 (same construction as geo2rdr) */
    double fd_c[64], fdd_c[64];
    if (dopAcc->order > 62) { free(ifg); return -6; }
    orc_poly1d fdvsrng = {dopAcc->order, p->rho0 + dopAcc->mean * p->drho, dopAcc->norm * p->drho, fd_c};
    for (int k = 1; k <= dopAcc->order + 1; k++) {
        double temp = dopAcc->coeffs[k - 1];
        temp = temp * p->prf;
        fd_c[k - 1] = temp;
    }
    orc_poly1d fddotvsrng;
    fddotvsrng.coeffs = fdd_c;
    if (fdvsrng.order == 0) {
        fddotvsrng.order = 0;
        fddotvsrng.mean = 0.0;
        fddotvsrng.norm = 1.0;
        fdd_c[0] = 0.0;
    } else {
        fddotvsrng.order = fdvsrng.order - 1;
        fddotvsrng.mean = fdvsrng.mean;
        fddotvsrng.norm = fdvsrng.norm;
        for (int k = 1; k <= dopAcc->order; k++) {
            double temp = fd_c[k];
            temp = k * temp / fdvsrng.norm;
            fdd_c[k - 1] = temp;
        }
    }
    /* :229-236 (geozero always interpolates with the Hermite scheme: interpolateWGS84Orbit_f) */
    double xyz_mid[3], vel_mid[3];
    if (orc_interp_hermite(orb, tmid, xyz_mid, vel_mid) != 0) { free(ifg); return -7; }

    long long numOutsideDEM = 0, numOutsideImage = 0, cnt = 0, total_iters = 0;
    const int nout = iscomplex ? 2 : 1;
    for (int line = 1; line <= geo_len; line++) { /* :244 */
        float *orow = out + (size_t)(line - 1) * geo_wid * nout;
        int16_t *drow = dem_crop ? dem_crop + (size_t)(line - 1) * geo_wid : NULL;
        double *azrow = oaz ? oaz + (size_t)(line - 1) * geo_wid : NULL;
        double *rgrow = orng ? orng + (size_t)(line - 1) * geo_wid : NULL;
        for (int i = 0; i < geo_wid * nout; i++) orow[i] = 0.f;
        if (drow) for (int i = 0; i < geo_wid; i++) drow[i] = 0;
        if (azrow) for (int i = 0; i < geo_wid; i++) azrow[i] = NAN;
        if (rgrow) for (int i = 0; i < geo_wid; i++) rgrow[i] = NAN;
        const int idxlat = max_lat_idx + (line - 1);
        if (idxlat < 0 || idxlat > (demlength - 1)) { /* :250-253 */
            numOutsideDEM += demwidth;
            continue;
        }
        const float *dem = dem_full + (size_t)idxlat * demwidth; /* getLine(demAccessor, dem, idxlat+1) */
#pragma omp parallel for schedule(static) reduction(+ : numOutsideImage, cnt, total_iters)
        for (int pixel = 1; pixel <= geo_wid; pixel++) {
            cx4 z = {0.f, 0.f};
            double llh[3], xyz[3], satx[3], satv[3], dr[3], look_side_vec[3];
            double tline, tprev, rngpix = 0.0;
            int skip = 0;
            llh[2] = 0.0;
            llh[0] = lat_firstr + idxlat * dlatr;
            const int idxlon = min_lon_idx + (pixel - 1);
            llh[1] = lon_firstr + idxlon * dlonr;
            if (!(idxlon < 0 || idxlon > (demwidth - 1))) {
                llh[2] = dem[idxlon];
                if (llh[2] < -1500) skip = 1; /* bad SRTM pixels :290-292 */
            }
            if (!skip) {
                orc_latlon(p->major, p->e2, xyz, llh, 1);
                tline = tmid;
                for (int i = 0; i < 3; i++) { satx[i] = xyz_mid[i]; satv[i] = vel_mid[i]; }
                /* look side test :306-320 */
                for (int i = 0; i < 3; i++) dr[i] = xyz[i] - satx[i];
                v_cross(dr, satv, look_side_vec);
                const double look_side_sign = v_dot(look_side_vec, satx);
                const int pixel_side = look_side_sign > 0 ? -1 : 1;
                if (pixel_side != lookSide) skip = 1;
            }
            if (!skip) {
                for (int k = 1; k <= 21; k++) { /* :322-356 */
                    total_iters++;
                    tprev = tline;
                    for (int i = 0; i < 3; i++) dr[i] = xyz[i] - satx[i];
                    rngpix = v_norm(dr);
                    double dopfact = v_dot(dr, satv) / rngpix;
                    double fdop = 0.5 * p->wvl * orc_eval_poly1d(&fdvsrng, rngpix);
                    double fdopder = 0.5 * p->wvl * orc_eval_poly1d(&fddotvsrng, rngpix);
                    double c1 = dopfact - fdop;
                    double c2 = v_dot(satv, satv) / rngpix;
                    double c3 = dopfact * (fdop / rngpix + fdopder);
                    tline = tline + c1 / (c2 - c3);
                    int stat = orc_interp_hermite(orb, tline, satx, satv);
                    if (stat != 0) {
                        tline = BAD_VALUE;
                        rngpix = BAD_VALUE;
                        break;
                    }
                    if (fabs(tline - tprev) < 5.0e-7) break;
                }
                const double az_idx = ((tline - tstart) / dtaz) + 1;
                const double rng_idx = ((rngpix - rngstart) / dmrg) + 1;
                if (rng_idx <= f_delay || rng_idx >= width - f_delay || az_idx <= f_delay || az_idx >= length - f_delay) {
                    numOutsideImage++;
                } else {
                    cnt++;
                    const int int_rdx = (int)(rng_idx + f_delay);
                    const double fr_rdx = rng_idx + f_delay - int_rdx;
                    const int int_rdy = (int)(az_idx + f_delay);
                    const double fr_rdy = az_idx + f_delay - int_rdy;
                    if (azrow) azrow[pixel - 1] = az_idx;
                    if (rgrow) rgrow[pixel - 1] = rng_idx;
                    if (method == ORC_SINC) {
                        z = gz_sinc_cx(ifg, fintp, int_rdx - 1, int_rdy - 1, fr_rdx, fr_rdy, width, length);
                    } else if (method == ORC_BILINEAR) {
                        const double dx = int_rdx + fr_rdx - f_delay, dy = int_rdy + fr_rdy - f_delay;
                        z = gz_bilinear_cx(dy, dx, ifg, width);
                    } else if (method == ORC_BICUBIC) {
                        const double dx = int_rdx + fr_rdx - f_delay, dy = int_rdy + fr_rdy - f_delay;
                        z = gz_bicubic_cx(dy, dx, ifg, width);
                    } else {
                        const int dx = (int)lround(int_rdx + fr_rdx - f_delay), dy = (int)lround(int_rdy + fr_rdy - f_delay);
                        z = IFG(dx, dy);
                    }
                }
            }
            if (iscomplex) {
                orow[2 * (pixel - 1)] = z.re;
                orow[2 * (pixel - 1) + 1] = z.im;
            } else {
                orow[pixel - 1] = z.re; /* writeRealLine: real(carr) */
            }
            if (drow) drow[pixel - 1] = (int16_t)llh[2]; /* dem_crop is integer*2 (:22) */
        }
    }
    free(ifg);
    if (res) {
        res->geowidth = geo_wid;
        res->geolength = geo_len;
        res->geomin_lat = (p->lat_first + min_lat_idx * p->dlat); /* :419-422 */
        res->geomax_lat = (p->lat_first + max_lat_idx * p->dlat);
        res->geomin_lon = (p->lon_first + min_lon_idx * p->dlon);
        res->geomax_lon = (p->lon_first + max_lon_idx * p->dlon);
        res->num_outside_dem = numOutsideDEM;
        res->num_outside_image = numOutsideImage;
        res->num_valid = cnt;
        res->total_iters = total_iters;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* resamp_slc: components/stdproc/stdproc/resamp_slc/src/resamp_slc.f90  */
/* ------------------------------------------------------------------ */
/* resamp_slcMethods.f:57-83: sinc_coef as in topozero, then every sub-sample phase normalised to unit sum before the
 * real*4 table is filled */
static float *g_fintp_resamp = NULL;
static const float *resamp_sinc_table(void)
{
#pragma omp critical(orc_resamp_sinc_table)
    if (!g_fintp_resamp) {
        const double pi = 4.0 * atan(1.0);
        const double r_beta = 1.0, r_relfiltlen = 1.0 * SINC_LEN, r_pedestal = 0.0;
        const int i_decfactor = SINC_SUB;
        const int i_intplength = (int)lround(r_relfiltlen / r_beta);
        const int i_filtercoef = i_intplength * i_decfactor;
        const double r_wgthgt = (1.0 - r_pedestal) / 2.0;
        const double r_soff = i_filtercoef / 2.0;
        double *r_filter = calloc((size_t)i_filtercoef + 1, sizeof(double));
        for (int i = 0; i < i_filtercoef; i++) {
            double r_wa = i - r_soff;
            double r_s = r_wa * r_beta / (1.0 * i_decfactor);
            double r_fct = (r_s != 0.0) ? sin(pi * r_s) / (pi * r_s) : 1.0;
            double r_wgt = (1.0 - r_wgthgt) + r_wgthgt * cos((pi * r_wa) / r_soff);
            r_filter[i] = r_fct * r_wgt;
        }
        for (int i = 0; i < SINC_SUB; i++) {
            double ssum = 0.0;
            for (int j = 0; j < SINC_LEN; j++) ssum = ssum + r_filter[i + j * SINC_SUB];
            for (int j = 0; j < SINC_LEN; j++) r_filter[i + j * SINC_SUB] = r_filter[i + j * SINC_SUB] / ssum;
        }
        float *f = malloc(sizeof(float) * SINC_SUB * SINC_LEN);
        for (int i = 0; i < SINC_LEN; i++)
            for (int j = 0; j < SINC_SUB; j++) f[i + j * SINC_LEN] = (float)r_filter[j + i * SINC_SUB];
        free(r_filter);
        g_fintp_resamp = f;
    }
    return g_fintp_resamp;
}
void orc_resamp_sinc_table(float *out) { memcpy(out, resamp_sinc_table(), sizeof(float) * SINC_SUB * SINC_LEN); }

/* default COMPLEX product (real*4 components, each product and the sum rounded separately) */
static inline cx4 c4_mul(cx4 a, cx4 b)
{
    cx4 r;
    float t1 = a.re * b.re, t2 = a.im * b.im, t3 = a.re * b.im, t4 = a.im * b.re;
    r.re = t1 - t2;
    r.im = t3 + t4;
    return r;
}
/* MODULO(a, p) for reals as gfortran expands it: fmod, then shifted into the sign of p */
static inline double f_modulo(double a, double p)
{
    double r = fmod(a, p);
    if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
    return r;
}
static double eval2d_or_zero(const orc_poly2d *poly, double azi, double rng)
{
    return poly ? orc_eval_poly2d(poly, azi, rng) : 0.0;
}

int orc_resamp_slc(const orc_resamp_params *p, const orc_poly2d *rgCarrier, const orc_poly2d *azCarrier,
                   const orc_poly2d *rgOffsetsPoly, const orc_poly2d *azOffsetsPoly, const orc_poly2d *dopplerPoly,
                   const float *in, const double *residaz_img, const double *residrg_img, float *out, int nthreads)
{
    const double PI = 4.0 * atan(1.0);
    const int inwidth = p->inwidth, inlength = p->inlength, outwidth = p->outwidth, outlength = p->outlength;
    const int sinchalf = SINC_LEN / 2, sincone = SINC_LEN + 1;
    if (inwidth < 1 || inlength < 1 || outwidth < 1 || outlength < 1) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const float *fintp = resamp_sinc_table();
    cx4 *cin = malloc(sizeof(cx4) * (size_t)inwidth * (size_t)inlength);
    if (!cin) return -4;
#define CIN(i, j) cin[(size_t)((j) - 1) * (size_t)inwidth + (size_t)((i) - 1)]
    /* all carriers are removed from the data up front (:122-141) */
    for (int j = 1; j <= inlength; j++) {
        const double r_at = j;
#pragma omp parallel for schedule(static)
        for (int i = 1; i <= inwidth; i++) {
            const double r_rt = i;
            double r_ph = eval2d_or_zero(rgCarrier, r_at, r_rt) + eval2d_or_zero(azCarrier, r_at, r_rt);
            r_ph = f_modulo(r_ph, 2.0 * PI);
            cx4 cl = {in[2 * ((size_t)(j - 1) * inwidth + (i - 1))], in[2 * ((size_t)(j - 1) * inwidth + (i - 1)) + 1]};
            cx4 rot = {(float)cos(r_ph), (float)(-sin(r_ph))};
            CIN(i, j) = c4_mul(cl, rot);
        }
    }
    for (int j = 1; j <= outlength; j++) { /* :166 */
        const double *residaz = residaz_img ? residaz_img + (size_t)(j - 1) * outwidth : NULL;
        const double *residrg = residrg_img ? residrg_img + (size_t)(j - 1) * outwidth : NULL;
        float *cout = out + 2 * (size_t)(j - 1) * outwidth;
#pragma omp parallel for schedule(static)
        for (int i = 1; i <= outwidth; i++) {
            cx4 chip[9][9]; /* chip(ii,jj) -> chip[jj-1][ii-1] */
            cout[2 * (i - 1)] = 0.f;
            cout[2 * (i - 1) + 1] = 0.f;
            double r_rt = i, r_at = j;
            const double r_ro = eval2d_or_zero(rgOffsetsPoly, r_at, r_rt) + (residrg ? residrg[i - 1] : 0.0);
            const double r_ao = eval2d_or_zero(azOffsetsPoly, r_at, r_rt) + (residaz ? residaz[i - 1] : 0.0);
            const int k = (int)floor(i + r_ro);
            const double fracr = i + r_ro - k;
            if ((k <= sinchalf) || (k >= (inwidth - sinchalf))) continue;
            const int kk = (int)floor(j + r_ao);
            const double fraca = j + r_ao - kk;
            if ((kk <= sinchalf) || (kk >= (inlength - sinchalf))) continue;
            /* :211.  Test hook (orc_test_set_cpp_quirks bit 2): where the reference's C++ restatement evaluates its
             * polynomials -- the Doppler at the output pixel (ResampSlc.cpp:293, the form resamp_slc.f90 had before
             * 12-AUG-2020) and the carriers that are added back one sample / line earlier (ResampSlc.cpp:311: 0-based
             * i+ao, j+ro against the 1-based :233-235 here) -- so that the two can be compared with polynomials that vary */
            const double r_dop = g_resamp_cpp_positions ? eval2d_or_zero(dopplerPoly, r_at, r_rt)
                                                        : eval2d_or_zero(dopplerPoly, r_at + r_ao, r_rt + r_ro);
            for (int jj = 1; jj <= sincone; jj++) {
                const int chipj = kk + jj - 1 - sinchalf;
                cx4 cval = {(float)cos((jj - 5.0) * r_dop), (float)(-sin((jj - 5.0) * r_dop))};
                for (int ii = 1; ii <= sincone; ii++) {
                    const int chipi = k + ii - 1 - sinchalf;
                    chip[jj - 1][ii - 1] = c4_mul(CIN(chipi, chipj), cval);
                }
            }
            double r_ph = r_dop * fraca;
            r_rt = i + r_ro;
            r_at = j + r_ao;
            if (g_resamp_cpp_positions) {
                r_rt = (i - 1) + r_ro;
                r_at = (j - 1) + r_ao;
            }
            r_ph = r_ph + eval2d_or_zero(rgCarrier, r_at, r_rt) + eval2d_or_zero(azCarrier, r_at, r_rt);
            if (p->flatten != 0)
                r_ph = r_ph + (4.0 * PI / p->wvl) * ((p->r0 - p->refr0) + (i - 1.0) * (p->slr - p->refslr) + r_ro * p->slr) +
                       (4.0 * PI * (p->refr0 + (i - 1.0) * p->refslr)) * (1.0 / p->refwvl - 1.0 / p->wvl);
            r_ph = f_modulo(r_ph, 2.0 * PI);
            /* intp_sinc_cx(chip, 5, 5, fracr, fraca, 9, 9) -> sinc_eval_2d_cx(chip, fintp, 8192, 8, 8, 8, ...) with
             * arrin(a, b) = chip(a+1, b+1): uniform_interp.f90:456-484 */
            cx4 acc = {0.f, 0.f};
            {
                const int idec = SINC_SUB, ilen = SINC_LEN, intpx = 8, intpy = 8;
                int ifracx = (int)(fracr * idec), ifracy = (int)(fraca * idec);
                ifracx = ifracx < 0 ? 0 : (ifracx > idec - 1 ? idec - 1 : ifracx);
                ifracy = ifracy < 0 ? 0 : (ifracy > idec - 1 ? idec - 1 : ifracy);
                double fweightsum = 0.0;
                for (int kq = 0; kq < ilen; kq++)
                    for (int m = 0; m < ilen; m++) {
                        const float fw4 = fintp[kq + ifracx * ilen] * fintp[m + ifracy * ilen];
                        const double fweight = fw4;
                        acc = c4(c8_add(c8(acc), c8_scale(c8(chip[intpy - m][intpx - kq]), fweight)));
                        fweightsum = fweightsum + fweight;
                    }
                acc = c4(c8_div(c8(acc), fweightsum));
            }
            cx4 rot = {(float)cos(r_ph), (float)sin(r_ph)};
            cx4 o = c4_mul(acc, rot);
            cout[2 * (i - 1)] = o.re;
            cout[2 * (i - 1) + 1] = o.im;
        }
    }
#undef CIN
    free(cin);
    return 0;
}
