// geozero_kernels.cuh -- device-side layout and launchers of the geozero (geocoding on the zero-Doppler geometry)
// kernels.  Reference: components/zerodop/geozero/src/geozero.f90, geozeroMethods.F.
#pragma once

#include <cuda_runtime.h>

#include "geo2rdr_kernels.cuh"

namespace b2 {

constexpr int kGeozeroBlock = 128;

// Scalars of geozero.f90:120-236
struct GeozeroConst {
    Ellipsoid elp;
    double wvl;
    double tstart, tmid, dtaz;
    double rngstart, dmrg;
    Poly1dDev fd, fdd; // fdvsrng, fddotvsrng (:196-224)
    Vec3 xyz_mid, vel_mid;
    int length, width; // radar image being geocoded
    int look_side;
    double lat_firstr, lon_firstr, dlatr, dlonr; // DEM origin / posting in radians (:146-149)
    int max_lat_idx, min_lon_idx;               // first DEM row / column of the output grid (:163-168)
    int geo_len, geo_wid;
    int demwidth, demlength;
    // the part of the DEM resident on the device: rows [dem_row0, dem_row0 + dem_rows), columns likewise
    const float *dem;
    int dem_row0, dem_col0, dem_rows, dem_cols;
};

struct GeozeroStats {
    unsigned long long outside_image, valid, iterations;
};

// Geometry of the output grid, solved once per plan and shared by every band / image geocoded with it:
// fractional 1-based image coordinates (az_idx, rng_idx of geozero.f90:359-360), NaN where the reference jumps to
// label 100 without interpolating (bad DEM sample, wrong look side, orbit interpolation failure).
struct GeozeroGeometry {
    double *az_idx, *rng_idx; // [geo_len][geo_wid]
    short *dem_crop;          // [geo_len][geo_wid] integer*2 (:22, :399)
    double *row_sc;           // [geo_len][2] sin / cos of the row latitude
    double *col_sc;           // [geo_wid][2] sin / cos of the column longitude
};

// one band of the image: element strides (float for real images, float2 for complex)
struct BandView {
    size_t offset, line_stride, pix_stride;
};

void launch_geozero_axes(const GeozeroConst &C, const GeozeroGeometry &G, cudaStream_t s);
int launch_geozero_solve(const GeozeroConst &C, const OrbitPolyView &op, const GeozeroGeometry &G, GeozeroStats *stats,
                         cudaStream_t s);
// method: 0 sinc, 1 bilinear, 2 bicubic, 3 nearest (geozeroMethods.F:29-31); sinc = fintp table [8192][8] or NULL
int launch_geozero_interp(const GeozeroConst &C, const GeozeroGeometry &G, int method, int is_complex, const void *image,
                          BandView in, void *out, BandView ov, const float *sinc, GeozeroStats *stats, cudaStream_t s);

} // namespace b2
