// TEST INFRASTRUCTURE ONLY.  extern "C" doors onto the reference's own C++ restatement of resamp_slc and its
// interpolation helpers -- components/zerodop/GPUresampslc/src/{Interpolator,ResampMethods,Poly2d,ResampSlc}.cpp,
// compiled UNCHANGED where they lie (oracle/Makefile, target ref; GPU_ACC_ENABLED undefined = its CPU branch).
#include <complex>
#include <vector>

#include "Interpolator.h"
#include "Poly2d.h"
#include "ResampSlc.h"
#include "ref_mem_accessor.h"

using std::vector;

// ResampSlc.cpp compiles its GPU entry unconditionally but sees the declaration of the CUDA launcher only under
// GPU_ACC_ENABLED; the Makefile force-includes the reference's own GPUresamp.h for the declaration, and the launcher
// (never reached: usr_enable_gpu is false without GPU_ACC_ENABLED) is this stub.
void runGPUResamp(double *, int *, void *, void *, float *, float *, double *, double *, double *, double *, double *, float *) {}

// coefficient block of one polynomial as the oracle's Python side passes it: [azimuthOrder, rangeOrder, azimuthMean,
// rangeMean, azimuthNorm, rangeNorm, coeffs (azimuth-major) ...]; NULL = the zero polynomial Resamp_slc.py substitutes
static Poly2d *make_poly(const double *a)
{
    if (!a) {
        Poly2d *z = new Poly2d(0, 0, 0., 0., 1., 1.);
        z->setCoeff(0, 0, 0.);
        return z;
    }
    const int ao = (int)a[0], ro = (int)a[1];
    Poly2d *q = new Poly2d(ro, ao, a[3], a[2], a[5], a[4]); // ctor order: (rangeOrder, azimuthOrder, rangeMean, azimuthMean, ...)
    for (int i = 0; i <= ao; i++)
        for (int j = 0; j <= ro; j++) q->setCoeff(i, j, a[6 + i * (ro + 1) + j]);
    return q;
}

extern "C" {
// Interpolator::sinc_coef (Interpolator.cpp:119-136) -- the post-2021 form, identical in formula to
// uniform_interp.f90:356-384.  filter: [relfiltlen/beta * decfactor] doubles.
void ref_cpp_sinc_coef(double beta, double relfiltlen, int decfactor, double pedestal, int weight, double *filter)
{
    Interpolator it;
    int intplength = 0, filtercoef = 0;
    vector<double> f((size_t)(relfiltlen / beta + 0.5) * decfactor + 1);
    it.sinc_coef(beta, relfiltlen, decfactor, pedestal, weight, intplength, filtercoef, f);
    for (int i = 0; i < filtercoef; i++) filter[i] = f[i];
}

// ResampSlc::_resamp_cpu (ResampSlc.cpp:164-383): the whole image through the reference's CPU branch.  in / out:
// complex float32 [length][width]; residaz / residrg: [outlength][outwidth] double or NULL.
int ref_cpp_resamp_slc(int inwidth, int inlength, int outwidth, int outlength, double wvl, double slr, double r0, double refwvl,
                       double refslr, double refr0, int flatten, const double *rgCarrier, const double *azCarrier,
                       const double *rgOffsets, const double *azOffsets, const double *doppler, const float *in,
                       const double *residaz, const double *residrg, float *out)
{
    ResampSlc R;
    R.wvl = wvl; R.slr = slr; R.r0 = r0; R.refwvl = refwvl; R.refslr = refslr; R.refr0 = refr0;
    R.inWidth = inwidth; R.inLength = inlength; R.outWidth = outwidth; R.outLength = outlength;
    R.isComplex = true;
    R.flatten = flatten != 0;
    R.usr_enable_gpu = false;
    R.setRgCarrier(make_poly(rgCarrier));
    R.setAzCarrier(make_poly(azCarrier));
    R.setRgOffsets(make_poly(rgOffsets));
    R.setAzOffsets(make_poly(azOffsets));
    R.setDoppler(make_poly(doppler));
    MemAccessor a_in((void *)in, inlength, inwidth, 1, 8), a_out((void *)out, outlength, outwidth, 1, 8);
    MemAccessor a_raz((void *)residaz, outlength, outwidth, 1, 8), a_rrg((void *)residrg, outlength, outwidth, 1, 8);
    R.slcInAccessor = (uint64_t)(DataAccessor *)&a_in;
    R.slcOutAccessor = (uint64_t)(DataAccessor *)&a_out;
    R.residAzAccessor = residaz ? (uint64_t)(DataAccessor *)&a_raz : 0;
    R.residRgAccessor = residrg ? (uint64_t)(DataAccessor *)&a_rrg : 0;
    R.resamp();
    return 0;
}
}
