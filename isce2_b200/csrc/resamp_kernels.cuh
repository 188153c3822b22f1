// resamp_kernels.cuh -- device-side layout and launchers of the resamp_slc kernels (the consumer of geo2rdr's offsets).
// Reference: components/stdproc/stdproc/resamp_slc/src/resamp_slc.f90, resamp_slcMethods.f.
#pragma once

#include <cuda_runtime.h>

#include "geom_device.cuh"

namespace b2 {

constexpr int kResampBlock = 128;

// module resamp_slcState + the five polynomials handed over by Resamp_slc.py:75-80
struct ResampConst {
    int inwidth, inlength, outwidth, outlength;
    double wvl, slr, r0, refwvl, refr0, refslr;
    int flatten;
    int has_carrier; // 0: both carrier polynomials are identically zero (the up-front pass is the identity)
    double pi;
    Poly2dDev rg_carrier, az_carrier, rg_off, az_off, dop;
};

struct ResampStats {
    unsigned long long valid; // output pixels that reached the interpolator
};

// carrier removal of the whole input image (resamp_slc.f90:122-141): out may alias in
void launch_resamp_carrier(const ResampConst &C, const float2 *in, float2 *out, cudaStream_t s);
// resid_*: [outlength][outwidth] of double (resid_f32 == 0) or float32 (== 1), or NULL; sinc: the normalised table
// of resamp_slcMethods.f:57-83 ([8192][8] float)
int launch_resamp_slc(const ResampConst &C, const float2 *cin, const void *resid_az, const void *resid_rg, int resid_f32,
                      const float *sinc, float2 *out, ResampStats *stats, cudaStream_t s);

} // namespace b2
