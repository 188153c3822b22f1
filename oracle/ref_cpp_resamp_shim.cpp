// TEST INFRASTRUCTURE ONLY.  extern "C" doors onto the reference's own C++ restatement of resamp_slc and its
// interpolation helpers -- components/zerodop/GPUresampslc/src/{Interpolator,ResampMethods,Poly2d,ResampSlc}.cpp,
// compiled UNCHANGED where they lie (oracle/Makefile, target ref; GPU_ACC_ENABLED undefined = its CPU branch).
#include <vector>

#include "Interpolator.h"

using std::vector;

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
}
