// topo_pixel.cuh -- the per-radar-pixel body of topozero (rdr2geo): iterative height solve against
// the DEM followed by the final geolocation / LOS / incidence pass.
//
// Behaviour follows components/zerodop/topozero/src/topozero.f90:458-708 including its mixed
// precision: float32 DEM indices (demlat/demlon, :54,525-536), float32 interpolated heights
// (topozeroMethods.f:116), float32 convergence distance (:20,565-567) and float32 angle layers (:19).
//
// What is NOT the reference's instruction stream (see DESIGN.md "arithmetic"):
//   * divisions by scene / line / pixel constants go through div_r with a hoisted reciprocal, the remaining
//     divisions and square roots through div_n / sqrt_n: all give the IEEE result, at a fraction of the cost;
//   * latitude / longitude arctangents and the sines / cosines of LLH->XYZ are evaluated about the block's
//     reference angles (atan2_ref / sincos_ref, ~0.5 ulp) when REF is set, with libm otherwise;
//   * the iteration skips quantities the reference computes but never reads (height of the intermediate
//     point, llh_prev outside the secondary iterations).
#pragma once

#include "dem_interp.cuh"
#include "geom_device.cuh"

namespace b2 {

struct TopoConst {
    Ellipsoid elp;
    GeoRef ref;
    double wvl, thresh;
    int ilrl, numiter, extraiter;
    double ufirstlat, ufirstlon, deltalat, deltalon; // cropped DEM origin (topozero.f90:316-317) and posting
    DemView dem;
    int method;
    int width, length;
    int nazlooks;
    double t0, prf, peghdg;
    double pi, r2d;
    double inv_r2d, inv_dlat, inv_dlon; // IEEE reciprocals for div_r
    int orbit_method;
    Poly2dDev dop, slr;
    const double *rho_image; // optional [length][width] slant-range image (slantRangeFilename case), else NULL
    Spline6Table spl;
};

struct PixelResult {
    double lat, lon, hgt; // degrees, degrees, metres (full double, topozero.f90:642-644)
    float los0, los1;     // :657-658
    float inc0, inc1;     // psi (:700) and local incidence (:692)
    float elev;           // elevang (:659), used by the shadow test
    double ctrack;        // rng*sintheta (:662), used by the layover test
    int converged;
    int iters;
};

template <int METHOD>
B2_HD float interp_dem(const TopoConst &C, int ix, int iy, double fx, double fy)
{
    if (METHOD == 0) return interp_sinc(C.dem, ix, iy, fx, fy);
    if (METHOD == 1) return interp_bilinear(C.dem, ix, iy, fx, fy);
    if (METHOD == 2) return interp_bicubic(C.dem, ix, iy, fx, fy);
    if (METHOD == 3) return interp_nearest(C.dem, ix, iy, fx, fy);
    if (METHOD == 4) return interp_akima(C.dem, ix, iy, fx, fy);
    return interp_biquintic(C.dem, C.spl, ix, iy, fx, fy);
}

// float32 DEM index of a latitude/longitude in degrees, clamped to [lo, n-1] (:525-536 / :666-677)
B2_HD void dem_index(const TopoConst &C, double lat_deg, double lon_deg, float lo, int &idemlat, int &idemlon,
                     double &fraclat, double &fraclon)
{
    float demlat = (float)(div_r(lat_deg - C.ufirstlat, C.deltalat, C.inv_dlat) + 1);
    float demlon = (float)(div_r(lon_deg - C.ufirstlon, C.deltalon, C.inv_dlon) + 1);
    const float hy = (float)(C.dem.ny - 1), hx = (float)(C.dem.nx - 1);
    if (demlat < lo) demlat = lo;
    if (demlat > hy) demlat = hy;
    if (demlon < lo) demlon = lo;
    if (demlon > hx) demlon = hx;
    idemlat = (int)demlat;
    idemlon = (int)demlon;
    fraclat = (double)(float)(demlat - (float)idemlat);
    fraclon = (double)(float)(demlon - (float)idemlon);
}

// per-pixel constants of the range-sphere solve (hoisted out of the iteration; all bit-identical to the
// reference's in-loop expressions)
struct PixelConst {
    double rng, inv_rng, rng2, dopfact;
    double a12; // (aa/rng) + (rng/aa)
};

B2_HD PixelConst make_pixel_const(const TopoConst &C, const LineState &L, double rng, double dopline)
{
    PixelConst P;
    P.rng = rng;
    P.inv_rng = rcp_n(rng);
    P.rng2 = rng * rng;
    P.dopfact = div_r(0.5 * C.wvl * dopline, L.vmag, L.inv_vmag) * rng;
    P.a12 = div_r(L.aa, rng, P.inv_rng) + div_r(rng, L.aa, L.inv_aa);
    return P;
}

// range-sphere / Doppler-cone intersection at SCH height zsch (:495-516)
B2_HD void range_sphere(const TopoConst &C, const LineState &L, const PixelConst &P, double zsch, double &costheta,
                        double &sintheta, Vec3 &delta, Vec3 &xyz)
{
    double bb = L.rcurv + zsch;
    costheta = 0.5 * (P.a12 - div_r(bb, L.aa, L.inv_aa) * div_r(bb, P.rng, P.inv_rng));
    sintheta = sqrt_n(1.0 - costheta * costheta);
    double gamm = costheta * P.rng;
    double alpha = div_r(P.dopfact - gamm * L.nv, L.vt, L.inv_vt);
    double beta = -C.ilrl * sqrt_n(P.rng2 * sintheta * sintheta - alpha * alpha);
    delta.x = gamm * L.nhat.x + alpha * L.that.x + beta * L.chat.x;
    delta.y = gamm * L.nhat.y + alpha * L.that.y + beta * L.chat.y;
    delta.z = gamm * L.nhat.z + alpha * L.that.z + beta * L.chat.z;
    xyz = add(L.sat, delta);
}

B2_HD float range_distance(const LineState &L, const Vec3 &xyz, double rng)
{
    double dx = xyz.x - L.sat.x, dy = xyz.y - L.sat.y, dz = xyz.z - L.sat.z;
    return (float)(sqrt_p(dx * dx + dy * dy + dz * dz) - rng);
}

template <bool REF>
B2_HD Vec3 geodetic_to_xyz(const TopoConst &C, double lat_deg, double lon_deg, double h)
{
    double la = div_r(lat_deg, C.r2d, C.inv_r2d), lo = div_r(lon_deg, C.r2d, C.inv_r2d); // lat(pixel)/r2d (:550-551)
    return REF ? llh_to_xyz_ref(C.elp, C.ref, la, lo, h) : llh_to_xyz(C.elp, la, lo, h);
}

// One primary iteration (:486-570).  Returns true when the pixel converged.
template <int METHOD, bool REF>
B2_HD bool topo_iterate(const TopoConst &C, const LineState &L, const PixelConst &P, double &lat, double &lon, double &z,
                        double &zsch, Vec3 &xyz)
{
    double ct, st, la, lo, h;
    Vec3 delta;
    range_sphere(C, L, P, zsch, ct, st, delta, xyz);
    if (REF) xyz_to_llh_ref<false>(C.elp, C.ref, xyz, la, lo, h);
    else xyz_to_latlon(C.elp, xyz, la, lo);
    lat = la * C.r2d;
    lon = lo * C.r2d;
    int idemlat, idemlon;
    double fraclat, fraclon;
    dem_index(C, lat, lon, 1.0f, idemlat, idemlon, fraclat, fraclon);
    z = interp_dem<METHOD>(C, idemlon, idemlat, fraclon, fraclat);
    if (z < -500.0) z = -500.0;
    xyz = geodetic_to_xyz<REF>(C, lat, lon, z);
    zsch = sch_height(L, xyz);
    float distance = range_distance(L, xyz, P.rng);
    return fabs((double)distance) <= C.thresh;
}

// Secondary iterations (:572-593): iterations numiter+2 .. numiter+extraiter+1 average each new point with the
// previous one.  Reached only by pixels that did not converge in the primary iterations (layover, DEM edges), so it
// is kept out of line and out of the hot loop's register budget.
template <int METHOD, bool REF>
#ifdef __CUDACC__
__device__ __noinline__
#else
inline
#endif
void topo_secondary(const TopoConst &C, const LineState &L, const PixelConst &P, double &lat, double &lon, double &z, double &zsch,
                    int &converged, int &iters)
{
    for (int iter = C.numiter + 2; iter <= C.numiter + C.extraiter + 1; iter++) {
        if (converged) break;
        iters++;
        const double lat_prev = lat, lon_prev = lon, z_prev = z;
        Vec3 xyz;
        if (topo_iterate<METHOD, REF>(C, L, P, lat, lon, z, zsch, xyz)) {
            converged = 1;
        } else {
            Vec3 xyz_prev = geodetic_to_xyz<REF>(C, lat_prev, lon_prev, z_prev);
            xyz.x = 0.5 * (xyz_prev.x + xyz.x);
            xyz.y = 0.5 * (xyz_prev.y + xyz.y);
            xyz.z = 0.5 * (xyz_prev.z + xyz.z);
            double la, lo, h;
            if (REF) xyz_to_llh_ref<true>(C.elp, C.ref, xyz, la, lo, h);
            else xyz_to_llh(C.elp, xyz, la, lo, h);
            lat = la * C.r2d;
            lon = lo * C.r2d;
            z = h;
            zsch = sch_height(L, xyz);
        }
    }
}

// The iterative height solve of one pixel (:425-599): returns the SCH height the final pass starts from.
template <int METHOD, bool REF>
B2_HD double topo_solve(const TopoConst &C, const LineState &L, double rng, double dopline, int &converged, int &iters)
{
    const PixelConst P = make_pixel_const(C, L, rng, dopline);
    // :425-436 (the initial lat/lon only feed llh_prev of iteration 1, which nothing reads before the secondary phase)
    double lat = C.ufirstlat + 0.5 * C.deltalat * C.dem.ny;
    double lon = C.ufirstlon + 0.05 * C.deltalon * C.dem.nx;
    double z = 0.0, zsch = 0.0;
    converged = 0;
    iters = 0;
    const int nprimary = C.numiter + 1 < C.numiter + C.extraiter + 1 ? C.numiter + 1 : C.numiter + C.extraiter + 1;
#pragma unroll 1
    for (int iter = 1; iter <= nprimary; iter++) { // :458-570
        Vec3 xyz;
        iters++;
        if (topo_iterate<METHOD, REF>(C, L, P, lat, lon, z, zsch, xyz)) {
            converged = 1;
            break;
        }
    }
    if (!converged && C.extraiter > 0) topo_secondary<METHOD, REF>(C, L, P, lat, lon, z, zsch, converged, iters);
    return zsch;
}

// The final computation of one pixel (:618-707) from the converged SCH height.
template <int METHOD, bool REF>
B2_HD void topo_final(const TopoConst &C, const LineState &L, double rng, double dopline, double zsch, bool want_inc,
                      PixelResult &R)
{
    const double r2d = C.r2d;
    const PixelConst P = make_pixel_const(C, L, rng, dopline);
    double lat, lon;
    // ---- final computation :618-707 ----
    double costheta, sintheta, la, lo, h;
    Vec3 delta, xyz;
    range_sphere(C, L, P, zsch, costheta, sintheta, delta, xyz);
    if (REF) xyz_to_llh_ref<true>(C.elp, C.ref, xyz, la, lo, h);
    else xyz_to_llh(C.elp, xyz, la, lo, h);
    lat = la * r2d;
    lon = lo * r2d;
    R.lat = lat;
    R.lon = lon;
    R.hgt = h;
    // enubasis.F:39-60, xyz2enu = transpose(enumat)
    double clt, slt, clo, slo;
    if (REF) {
        sincos_ref(C.ref.lat, la, slt, clt);
        sincos_ref(C.ref.lon, lo, slo, clo);
    } else {
        clt = cos(la); slt = sin(la); clo = cos(lo); slo = sin(lo);
    }
    Vec3 e_east = Vec3{-slo, clo, 0.0};
    Vec3 e_north = Vec3{-slt * clo, -slt * slo, clt};
    Vec3 e_up = Vec3{clt * clo, clt * slo, slt};
    Vec3 enu = Vec3{dot(e_east, delta), dot(e_north, delta), dot(e_up, delta)};
    double en = sqrt_p(enu.x * enu.x + enu.y * enu.y + enu.z * enu.z);
    double cosalpha = div_n(fabs(enu.z), en);
    R.los0 = (float)(acos(cosalpha) * r2d);
    R.los1 = (float)((atan2(-enu.y, -enu.x) - 0.5 * C.pi) * r2d);
    R.elev = (float)(acos(costheta) * r2d);
    R.ctrack = rng * sintheta;
    R.inc0 = 0.f;
    R.inc1 = 0.f;
    if (want_inc) {
        int idemlat, idemlon;
        double fraclat, fraclon;
        dem_index(C, lat, lon, 2.0f, idemlat, idemlon, fraclat, fraclon);
        // slope probes (:680-688): same fractions at ix-1, ix+1, iy-1, iy+1
        double pr0 = 0.0, pr1 = 0.0, pr2 = 0.0, pr3 = 0.0;
        bool probed = false;
#ifdef __CUDA_ARCH__
        if (METHOD == 5) probed = biquintic_probes4(C.dem, C.spl, idemlon, idemlat, fraclon, fraclat, pr0, pr1, pr2, pr3);
#endif
#pragma unroll 1
        for (int j = probed ? 4 : 0; j < 4; j++) { // one copy of the interpolator in the instruction stream
            int dx = (j == 0) ? -1 : (j == 1 ? 1 : 0);
            int dy = (j == 2) ? -1 : (j == 3 ? 1 : 0);
            double v = interp_dem<METHOD>(C, idemlon + dx, idemlat + dy, fraclon, fraclat);
            pr0 = (j == 0) ? v : pr0;
            pr1 = (j == 1) ? v : pr1;
            pr2 = (j == 2) ? v : pr2;
            pr3 = (j == 3) ? v : pr3;
        }
        double gamm = div_r(lat, r2d, C.inv_r2d);
        double sg;
        if (REF) {
            double cg;
            sincos_ref(C.ref.lat, gamm, sg, cg);
        } else {
            sg = sin(gamm);
        }
        double alpha = div_n((pr1 - pr0) * r2d, 2.0 * reast_s(C.elp, sg) * C.deltalon);
        double beta = div_n((pr3 - pr2) * r2d, 2.0 * rnorth_s(C.elp, sg) * C.deltalat);
        enu.x = div_n(enu.x, en);
        enu.y = div_n(enu.y, en);
        enu.z = div_n(enu.z, en);
        double cinc = div_n(enu.x * alpha + enu.y * beta - enu.z, sqrt_p(1.0 + alpha * alpha + beta * beta));
        R.inc1 = (float)(acos(cinc) * r2d);
        // psi: angle between the image plane normal and the local slope normal (:694-700)
        Vec3 n_img = cross(delta, L.vel);
        double nn = sqrt_p(n_img.x * n_img.x + n_img.y * n_img.y + n_img.z * n_img.z);
        if (nn != 0) n_img = Vec3{div_n(n_img.x, nn), div_n(n_img.y, nn), div_n(n_img.z, nn)};
        Vec3 tmp = Vec3{-C.ilrl * n_img.x, -C.ilrl * n_img.y, -C.ilrl * n_img.z};
        Vec3 n_img_enu = Vec3{dot(e_east, tmp), dot(e_north, tmp), dot(e_up, tmp)};
        Vec3 n_trg = Vec3{-alpha, -beta, 1.0};
        double n1 = sqrt_p(n_trg.x * n_trg.x + n_trg.y * n_trg.y + n_trg.z * n_trg.z);
        double n2 = sqrt_p(n_img_enu.x * n_img_enu.x + n_img_enu.y * n_img_enu.y + n_img_enu.z * n_img_enu.z);
        double cospsi = div_n(dot(n_trg, n_img_enu), n1 * n2);
        R.inc0 = (float)(acos(cospsi) * r2d);
    }
}

// solve + final in one call (host emulation harness)
template <int METHOD, bool REF>
B2_HD void topo_pixel(const TopoConst &C, const LineState &L, double rng, double dopline, bool want_inc, PixelResult &R)
{
    int conv, iters;
    const double zsch = topo_solve<METHOD, REF>(C, L, rng, dopline, conv, iters);
    topo_final<METHOD, REF>(C, L, rng, dopline, zsch, want_inc, R);
    R.converged = conv;
    R.iters = iters;
}

// One sample of the regular cross-track grid used by the layover test (:745-782): returns the slant
// range of the DEM surface point under cross-track position aa, given the line's pixels sorted by ctrack.
template <int METHOD, bool REF>
B2_HD double mask_resample(const TopoConst &C, const LineState &L, const double *cs, const double *lats, const double *lons,
                           int it /*1-based, in [1, width-1]*/, double aa)
{
    double fraclat = div_n(aa - cs[it - 1], cs[it] - cs[it - 1]);
    float demlat = (float)(lats[it - 1] + fraclat * (lats[it] - lats[it - 1])); // real*4 in the reference (:755)
    float demlon = (float)(lons[it - 1] + fraclat * (lons[it] - lons[it - 1]));
    int idemlat, idemlon;
    double fraclon;
    dem_index(C, (double)demlat, (double)demlon, 2.0f, idemlat, idemlon, fraclat, fraclon);
    double hh = interp_dem<METHOD>(C, idemlon, idemlat, fraclon, fraclat);
    Vec3 xyz = geodetic_to_xyz<REF>(C, (double)demlat, (double)demlon, hh);
    xyz = sub(xyz, L.sat);
    return sqrt_p(xyz.x * xyz.x + xyz.y * xyz.y + xyz.z * xyz.z);
}

} // namespace b2
