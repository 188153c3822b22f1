// TEST INFRASTRUCTURE ONLY.  extern "C" door onto the CPU branch of the reference's own C++ restatement of geo2rdr --
// components/zerodop/GPUgeo2rdr/src/{Geo2rdr,Orbit,Poly1d,Ellipsoid,LinAlg}.cpp, compiled UNCHANGED where they lie
// (oracle/Makefile, target ref; GPU_ACC_ENABLED undefined).  Marshals plain buffers, computes nothing itself.
#include <cstdint>
#include <vector>

#include "ref_mem_accessor.h"

#include "Constants.h"
#include "Geo2rdr.h"

using std::vector;

extern "C" {
// Geo2rdr::geo2rdr (Geo2rdr.cpp:84-499, CPU branch :395-480).  lat / lon / hgt: [demlength][demwidth] double;
// outputs [demlength][demwidth] double, any may be NULL.  dop: Poly1d (order, mean, norm, coeffs) vs range pixel.
int ref_cpp_geo2rdr(double major, double e2, double drho, double rngstart, double wvl, double tstart, double prf,
                    int length, int width, int demlength, int demwidth, int nrnglooks, int nazlooks, int bistatic,
                    int orbit_method, int nvec, const double *ot, const double *opos, const double *ovel,
                    int dop_order, double dop_mean, double dop_norm, const double *dop_coeffs,
                    double *lat, double *lon, double *hgt, double *azt, double *rgm, double *azoff, double *rgoff)
{
    Geo2rdr g;
    g.major = major; g.eccentricitySquared = e2; g.drho = drho; g.rngstart = rngstart; g.wvl = wvl;
    g.tstart = tstart; g.prf = prf; g.imgLength = length; g.imgWidth = width; g.demLength = demlength;
    g.demWidth = demwidth; g.nRngLooks = nrnglooks; g.nAzLooks = nazlooks; g.bistatic = bistatic != 0;
    g.orbitMethod = orbit_method; g.usr_enable_gpu = false;
    g.orbit_nvecs = nvec; g.orbit_basis = WGS84_ORBIT;
    g.orb.setOrbit(nvec, WGS84_ORBIT);
    for (int i = 0; i < nvec; i++) {
        double p[3] = {opos[3 * i], opos[3 * i + 1], opos[3 * i + 2]}, v[3] = {ovel[3 * i], ovel[3 * i + 1], ovel[3 * i + 2]};
        g.orb.setStateVector(i, ot[i], p, v);
    }
    g.dop.setPoly(dop_order, dop_mean, dop_norm);
    for (int i = 0; i <= dop_order; i++) g.dop.setCoeff(i, dop_coeffs[i]);
    MemAccessor a_lat(lat, demlength, demwidth, 1, 8), a_lon(lon, demlength, demwidth, 1, 8), a_hgt(hgt, demlength, demwidth, 1, 8);
    MemAccessor a_az(azt, demlength, demwidth, 1, 8), a_rg(rgm, demlength, demwidth, 1, 8);
    MemAccessor a_azo(azoff, demlength, demwidth, 1, 8), a_rgo(rgoff, demlength, demwidth, 1, 8);
    g.latAccessor = (uint64_t)(DataAccessor *)&a_lat;
    g.lonAccessor = (uint64_t)(DataAccessor *)&a_lon;
    g.hgtAccessor = (uint64_t)(DataAccessor *)&a_hgt;
    g.azAccessor = azt ? (uint64_t)(DataAccessor *)&a_az : 0;
    g.rgAccessor = rgm ? (uint64_t)(DataAccessor *)&a_rg : 0;
    g.azOffAccessor = azoff ? (uint64_t)(DataAccessor *)&a_azo : 0;
    g.rgOffAccessor = rgoff ? (uint64_t)(DataAccessor *)&a_rgo : 0;
    g.geo2rdr();
    return 0;
}
}
