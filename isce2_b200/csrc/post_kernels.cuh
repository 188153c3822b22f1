// post_kernels.cuh -- launchers of the two remaining consumers of the geometry layers:
//   multilooking (contrib/stack/stripmapStack/topo.py:365-441 runMultilook; mroipac/looks/bindings/looksmodule.cpp:130-200)
//   projection of a geocoded (water) mask into radar coordinates through lat.rdr / lon.rdr
//     (contrib/demUtils/swbdstitcher/SWBDStitcher.py:107-131 toRadar, called by stripmapStack/createWaterMask.py:66-71)
// Both are HBM-bound streaming passes with no arithmetic to speak of.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>

namespace b2 {

// element types of looksmodule.cpp:80-127 (char / short / int / long / float / double and complex<float>)
enum { kTypeByte = 0, kTypeShort = 1, kTypeInt = 2, kTypeLong = 3, kTypeFloat = 4, kTypeDouble = 5, kTypeCFloat = 6 };
enum { kSchemeBIL = 0, kSchemeBIP = 1, kSchemeBSQ = 2 };
constexpr int kLooksMaxTile = 4096; // column sums of one tile held in shared memory (doubles; two per complex sample)

size_t type_size(int dtype);

struct LooksGeom {
    int length, width, bands, scheme; // input image
    int ld, la;                       // looks down (lines) / across (samples)
    int out_length, out_width;
    int line0, nlines;                // OUTPUT lines [line0, line0 + nlines) computed by this launch
};

// box mean (method 0) or nearest-neighbour decimation (method 1); `in` / `out` are the whole images on the device
int launch_looks(const LooksGeom &G, int dtype, int method, const void *in, void *out, cudaStream_t s);

struct MaskProj {
    int mask_length, mask_width;
    double start_lat, delta_lat, start_lon, delta_lon;
};
// out[p] = mask[clip(int((lat[p] - start_lat) / delta_lat))][clip(int((lon[p] - start_lon) / delta_lon))] + 1
// coord_f32: lat / lon are float32 (the arithmetic is then single precision, as numpy does it)
int launch_mask_to_radar(const MaskProj &M, int dtype, const void *mask, const void *lat, const void *lon, int coord_f32,
                         size_t npix, void *out, cudaStream_t s);

} // namespace b2
