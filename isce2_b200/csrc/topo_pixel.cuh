// topo_pixel.cuh -- the per-radar-pixel body of topozero (rdr2geo): iterative height solve against
// the DEM followed by the final geolocation / LOS / incidence pass.
//
// Behaviour follows components/zerodop/topozero/src/topozero.f90:458-708 including its mixed
// precision: float32 DEM indices (demlat/demlon, :54,525-536), float32 interpolated heights
// (topozeroMethods.f:116), float32 convergence distance (:20,565-567) and float32 angle layers (:19).
#pragma once

#include "dem_interp.cuh"
#include "geom_device.cuh"

namespace b2 {

struct TopoConst {
    Ellipsoid elp;
    double wvl, thresh;
    int ilrl, numiter, extraiter;
    double ufirstlat, ufirstlon, deltalat, deltalon; // cropped DEM origin (topozero.f90:316-317) and posting
    DemView dem;
    int method;
    int width, length;
    int nazlooks;
    double t0, prf, peghdg;
    double pi, r2d;
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
    if (METHOD == 1) return interp_bilinear(C.dem, ix, iy, fx, fy);
    if (METHOD == 2) return interp_bicubic(C.dem, ix, iy, fx, fy);
    if (METHOD == 3) return interp_nearest(C.dem, ix, iy, fx, fy);
    return interp_biquintic(C.dem, C.spl, ix, iy, fx, fy);
}

// float32 DEM index of a latitude/longitude in degrees, clamped to [lo, n-1] (:525-536 / :666-677)
B2_HD void dem_index(const TopoConst &C, double lat_deg, double lon_deg, float lo, int &idemlat, int &idemlon,
                     double &fraclat, double &fraclon)
{
    float demlat = (float)((lat_deg - C.ufirstlat) / C.deltalat + 1);
    float demlon = (float)((lon_deg - C.ufirstlon) / C.deltalon + 1);
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

// range-sphere / Doppler-cone intersection at SCH height zsch (:495-516)
B2_HD void range_sphere(const TopoConst &C, const LineState &L, double rng, double dopfact, double zsch, double &costheta,
                        double &sintheta, Vec3 &delta, Vec3 &xyz)
{
    double aa = L.height + L.rcurv;
    double bb = L.rcurv + zsch;
    costheta = 0.5 * ((aa / rng) + (rng / aa) - (bb / aa) * (bb / rng));
    sintheta = sqrt(1.0 - costheta * costheta);
    double gamm = costheta * rng;
    double alpha = (dopfact - gamm * L.nv) / L.vt;
    double beta = -C.ilrl * sqrt(rng * rng * sintheta * sintheta - alpha * alpha);
    delta.x = gamm * L.nhat.x + alpha * L.that.x + beta * L.chat.x;
    delta.y = gamm * L.nhat.y + alpha * L.that.y + beta * L.chat.y;
    delta.z = gamm * L.nhat.z + alpha * L.that.z + beta * L.chat.z;
    xyz = add(L.sat, delta);
}

B2_HD float range_distance(const LineState &L, const Vec3 &xyz, double rng)
{
    return (float)(sqrt((xyz.x - L.sat.x) * (xyz.x - L.sat.x) + (xyz.y - L.sat.y) * (xyz.y - L.sat.y) +
                        (xyz.z - L.sat.z) * (xyz.z - L.sat.z)) - rng);
}

template <int METHOD>
B2_HD void topo_pixel(const TopoConst &C, const LineState &L, double rng, double dopline, bool want_inc, PixelResult &R)
{
    const double r2d = C.r2d;
    const double dopfact = (0.5 * C.wvl * dopline / L.vmag) * rng;
    // :425-436
    double lat = C.ufirstlat + 0.5 * C.deltalat * C.dem.ny;
    double lon = C.ufirstlon + 0.05 * C.deltalon * C.dem.nx;
    double z = 0.0, zsch = 0.0;
    int converged = 0, iters = 0;
    const int niter = C.numiter + C.extraiter + 1;
    for (int iter = 1; iter <= niter; iter++) { // :458-599
        if (converged) break;
        iters++;
        double llh_prev0 = lat / r2d, llh_prev1 = lon / r2d, llh_prev2 = z;
        double ct, st, la, lo, h;
        Vec3 delta, xyz;
        range_sphere(C, L, rng, dopfact, zsch, ct, st, delta, xyz);
        xyz_to_llh(C.elp, xyz, la, lo, h);
        lat = la * r2d;
        lon = lo * r2d;
        int idemlat, idemlon;
        double fraclat, fraclon;
        dem_index(C, lat, lon, 1.0f, idemlat, idemlon, fraclat, fraclon);
        z = interp_dem<METHOD>(C, idemlon, idemlat, fraclon, fraclat);
        if (z < -500.0) z = -500.0;
        xyz = llh_to_xyz(C.elp, lat / r2d, lon / r2d, z);
        zsch = sch_height(L, xyz);
        float distance = range_distance(L, xyz, rng);
        if (fabs((double)distance) <= C.thresh) {
            converged = 1;
        } else if (iter > (C.numiter + 1)) { // :572-593
            Vec3 xyz_prev = llh_to_xyz(C.elp, llh_prev0, llh_prev1, llh_prev2);
            xyz.x = 0.5 * (xyz_prev.x + xyz.x);
            xyz.y = 0.5 * (xyz_prev.y + xyz.y);
            xyz.z = 0.5 * (xyz_prev.z + xyz.z);
            xyz_to_llh(C.elp, xyz, la, lo, h);
            lat = la * r2d;
            lon = lo * r2d;
            z = h;
            zsch = sch_height(L, xyz);
        }
    }
    R.converged = converged;
    R.iters = iters;

    // ---- final computation :618-707 ----
    double costheta, sintheta, la, lo, h;
    Vec3 delta, xyz;
    range_sphere(C, L, rng, dopfact, zsch, costheta, sintheta, delta, xyz);
    xyz_to_llh(C.elp, xyz, la, lo, h);
    lat = la * r2d;
    lon = lo * r2d;
    R.lat = lat;
    R.lon = lon;
    R.hgt = h;
    // enubasis.F:39-60, xyz2enu = transpose(enumat)
    double clt = cos(la), slt = sin(la), clo = cos(lo), slo = sin(lo);
    Vec3 e_east = Vec3{-slo, clo, 0.0};
    Vec3 e_north = Vec3{-slt * clo, -slt * slo, clt};
    Vec3 e_up = Vec3{clt * clo, clt * slo, slt};
    Vec3 enu = Vec3{dot(e_east, delta), dot(e_north, delta), dot(e_up, delta)};
    double cosalpha = fabs(enu.z) / norm(enu);
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
        double aa = interp_dem<METHOD>(C, idemlon - 1, idemlat, fraclon, fraclat);
        double bb = interp_dem<METHOD>(C, idemlon + 1, idemlat, fraclon, fraclat);
        double gamm = lat / r2d;
        double alpha = (bb - aa) * r2d / (2.0 * reast(C.elp, gamm) * C.deltalon);
        aa = interp_dem<METHOD>(C, idemlon, idemlat - 1, fraclon, fraclat);
        bb = interp_dem<METHOD>(C, idemlon, idemlat + 1, fraclon, fraclat);
        double beta = (bb - aa) * r2d / (2.0 * rnorth(C.elp, gamm) * C.deltalat);
        double en = norm(enu);
        enu.x = enu.x / en;
        enu.y = enu.y / en;
        enu.z = enu.z / en;
        double cinc = (enu.x * alpha + enu.y * beta - enu.z) / sqrt(1.0 + alpha * alpha + beta * beta);
        R.inc1 = (float)(acos(cinc) * r2d);
        // psi: angle between the image plane normal and the local slope normal (:694-700)
        Vec3 n_img = unitvec(cross(delta, L.vel));
        Vec3 tmp = Vec3{-C.ilrl * n_img.x, -C.ilrl * n_img.y, -C.ilrl * n_img.z};
        Vec3 n_img_enu = Vec3{dot(e_east, tmp), dot(e_north, tmp), dot(e_up, tmp)};
        Vec3 n_trg = Vec3{-alpha, -beta, 1.0};
        double cospsi = dot(n_trg, n_img_enu) / (norm(n_trg) * norm(n_img_enu));
        R.inc0 = (float)(acos(cospsi) * r2d);
    }
}

// One sample of the regular cross-track grid used by the layover test (:745-782): returns the slant
// range of the DEM surface point under cross-track position aa, given the line's pixels sorted by ctrack.
template <int METHOD>
B2_HD double mask_resample(const TopoConst &C, const LineState &L, const double *cs, const double *lats, const double *lons,
                           int it /*1-based, in [1, width-1]*/, double aa)
{
    const double r2d = C.r2d;
    double fraclat = (aa - cs[it - 1]) / (cs[it] - cs[it - 1]);
    float demlat = (float)(lats[it - 1] + fraclat * (lats[it] - lats[it - 1])); // real*4 in the reference (:755)
    float demlon = (float)(lons[it - 1] + fraclat * (lons[it] - lons[it - 1]));
    double llh0 = demlat / r2d, llh1 = demlon / r2d;
    int idemlat, idemlon;
    double fraclon;
    dem_index(C, (double)demlat, (double)demlon, 2.0f, idemlat, idemlon, fraclat, fraclon);
    double hh = interp_dem<METHOD>(C, idemlon, idemlat, fraclon, fraclat);
    Vec3 xyz = llh_to_xyz(C.elp, llh0, llh1, hh);
    xyz = sub(xyz, L.sat);
    return norm(xyz);
}

} // namespace b2
