// topo_kernels.cuh -- device-side data layout and launchers of the topozero kernels.
#pragma once

#include <cuda_runtime.h>

#include "topo_pixel.cuh"

namespace b2 {

constexpr int kTopoBlock = 128; // threads per CTA of the per-pixel kernel (one pixel per thread)
#ifndef B2_MASK_BLOCK
#define B2_MASK_BLOCK 1024
#endif
constexpr int kMaskBlock = B2_MASK_BLOCK;

// Resident output layers of one block of azimuth lines (device pointers; optional ones may be null).
struct TopoLayers {
    double *lat, *lon, *hgt; // [nlines][width]
    float *los, *inc;        // [nlines][2][width]  BIL
    signed char *mask;       // [nlines][width]
    double *ctrack;          // [nlines][width] scratch for the mask pass (rng*sintheta)
    float *elev;             // [nlines][width] scratch for the mask pass (elevang)
};

struct TopoStats {
    long long min_lat, max_lat, min_lon, max_lon; // order-preserving integer images of the doubles
    unsigned long long converged, iterations;
};

struct MaskScratch { // per persistent CTA
    double *cs, *lats, *lons;                       // [grid][width]
    double *orng, *ctr_sorted, *orng_sorted;        // [grid][2*width+1]
    double *pm, *sm;                                // [grid][2*width+1] prefix max / suffix min
    int *rank;                                      // [grid][2*width+1]
    unsigned char *oflag;                           // [grid][2*width+1]
};

double stats_decode(long long k);
float dem_max_decode(int key);

void launch_topo_bbox(const TopoConst &C, const OrbitView &orb, double *d_out, cudaStream_t s);
void launch_dem_prepare(const void *raw, int dtype, float *dem, size_t n, int *maxkey, cudaStream_t s);
void launch_dem_pad64(const float *dem, int nx, int ny, double *out, int stride, cudaStream_t s);
void launch_line_setup(const TopoConst &C, const OrbitView &orb, int line0, int nlines, LineState *states, cudaStream_t s);
// ev_mid (optional) is recorded between the solve and the final kernel when they are separate launches
int launch_topo_pixels(const TopoConst &C, const LineState *states, int line0, int nlines, const TopoLayers &out,
                       TopoStats *stats, cudaStream_t s, cudaEvent_t ev_mid = nullptr);
int topo_pixel_launches(int method);
int mask_grid_size(int nlines);
int launch_topo_mask(const TopoConst &C, const LineState *states, int line0, int nlines, const TopoLayers &out, float demmax,
                     const MaskScratch &scr, int grid, cudaStream_t s);

} // namespace b2
