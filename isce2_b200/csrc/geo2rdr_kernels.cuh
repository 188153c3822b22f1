// geo2rdr_kernels.cuh -- device-side layout and launchers of the geo2rdr kernels.
#pragma once

#include <cuda_runtime.h>

#include "geom_device.cuh"

namespace b2 {

constexpr int kGeoBlock = 128;
constexpr int kGeoMaxSmemVectors = 512; // state vectors staged in shared memory (7 doubles each)

// Scalars of geo2rdr.f90:118-208 (image limits, doppler-vs-range polynomials, mid-scene state).
struct GeoConst {
    Ellipsoid elp;
    double wvl;
    double tstart, tend, tmid, dtaz;
    double rngstart, rngend, dmrg;
    Poly1dDev fd, fdd; // fdvsrng, fddotvsrng (:161-189)
    Vec3 xyz_mid, vel_mid, acc_mid;
    int orbit_method, bistatic;
    int demwidth;
    double deg2rad, sol;
    int xyz_in; // 1: the lat / lon / hgt layers hold ECEF x / y / z (frozen stack geometry, launch_llh_to_xyz)
};

struct GeoLayers {
    const double *lat, *lon, *hgt; // [nlines][demwidth] rows of the block, degrees / metres
    void *azt, *rgm, *azoff, *rgoff; // [nlines][demwidth] float or double, may be null
};

// ECEF coordinates of a lat / lon / hgt geometry, exactly as the solve kernels form them per pixel (geo2rdr.f90:247-250)
void launch_llh_to_xyz(const GeoConst &C, const double *lat, const double *lon, const double *hgt, double *x, double *y, double *z,
                       size_t n, cudaStream_t s);

struct GeoStats {
    unsigned long long outside, valid, converged, iterations;
};

// Per-window orbit polynomials (host: orbit_poly.h).  s = (t - tc[w]) * inv_h[w]; coefficients highest power first.
struct OrbitPolyView {
    int method; // 0 Hermite (velocity = d/dt of the position polynomial), 2 Legendre (separate velocity polynomial)
    int n, nwin, ncoef;
    const double *t;     // [n] state-vector epochs (window selection + span test, orbit.c:203-233)
    const double *tc;    // [nwin]
    const double *inv_h; // [nwin]
    const double *cp;    // [nwin][3][ncoef]
    const double *cv;    // [nwin][3][ncoef] (Legendre only)
};

struct GeoMid { // result of k_geo_setup
    double xyz[3], vel[3], acc[3];
    int stat_mid, stat_acc;
};

void launch_geo_setup(int orbit_method, const OrbitView &orb, double tmid, GeoMid *d_out, cudaStream_t s);
int launch_geo2rdr(const GeoConst &C, const OrbitView &orb, int line0, int nlines, const GeoLayers &L, int out_f32,
                   GeoStats *stats, cudaStream_t s);
// Newton solve on the per-window orbit polynomials (Hermite / Legendre); same outputs as launch_geo2rdr
int launch_geo2rdr_poly(const GeoConst &C, const OrbitPolyView &op, int line0, int nlines, const GeoLayers &L, int out_f32,
                        GeoStats *stats, cudaStream_t s);

} // namespace b2
