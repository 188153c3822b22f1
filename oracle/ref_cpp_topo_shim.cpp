// TEST INFRASTRUCTURE ONLY.  extern "C" doors onto the reference's own C++ restatement of the topozero path --
// components/zerodop/GPUtopozero/src/{TopoMethods,UniformInterp,AkimaLib,Ellipsoid,LinAlg,Peg,PegTrans,Orbit,Poly2d,Topo}.cpp,
// compiled UNCHANGED where they lie (oracle/Makefile, target ref; GPU_ACC_ENABLED undefined = its CPU branch) --
// so that tests/ can hold oracle/zerodop_oracle.c against reference-authored code.  Nothing here computes anything
// itself: every function marshals plain buffers into the reference's std::vector types and calls it.
#include <cstddef>
#include <cstdint>
#include <vector>

#include "ref_mem_accessor.h"

#include "Constants.h"
#include "Ellipsoid.h"
#include "LinAlg.h"
#include "Orbit.h"
#include "Peg.h"
#include "PegTrans.h"
#include "Poly2d.h"
#include "Topo.h"
#include "TopoMethods.h"
#include "UniformInterp.h"

// Topo.cpp references its CUDA branch unconditionally (if (RUN_GPU_TOPO) with RUN_GPU_TOPO == 0): never called
size_t getDeviceFreeMem() { return 0; }
void runGPUTopo(long, long, double *, int *, float *, double *, double *, int, double *, double **) {}

using std::vector;

extern "C" {

// dem: [ny][nx] float32 row-major (lon fastest), as the oracle takes it; the reference indexes dem[ix-1][iy-1]
// n points (ix, iy 1-based; fx, fy fractions) -> out[n].  Returns 0.
// sinc_table (optional, [8192*8]): replaces the table prepareMethods built, see tests/test_oracle_cpp_pins.py (the
// reference's C++ sinc_coef predates the 2021 change of uniform_interp.f90:319-363)
int ref_cpp_interp_dem(int method, const float *dem, int nx, int ny, long n, const int *ix, const int *iy,
                       const double *fx, const double *fy, float *out, const float *sinc_table)
{
    vector<vector<float> > d(nx, vector<float>(ny));
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) d[i][j] = dem[(size_t)j * nx + i];
    TopoMethods tm;
    tm.prepareMethods(method);
    if (sinc_table && method == SINC_METHOD) tm.fintp.assign(sinc_table, sinc_table + SINC_SUB * SINC_LEN);
    for (long k = 0; k < n; k++) out[k] = tm.interpolate(d, ix[k], iy[k], fx[k], fy[k], nx, ny, method);
    return 0;
}

// the sinc table the reference builds for its DEM interpolator (TopoMethods::prepareMethods): out[8192*8]
void ref_cpp_sinc_table(float *out)
{
    TopoMethods tm;
    tm.prepareMethods(SINC_METHOD);
    for (size_t i = 0; i < tm.fintp.size(); i++) out[i] = tm.fintp[i];
}

// 1-D pieces of the spline behind the biquintic method
void ref_cpp_spline6(const double *y, double x, double *r_out, double *val)
{
    UniformInterp u;
    vector<double> Y(y, y + 6), R(6), Q(6);
    u.initSpline(Y, 6, R, Q);
    for (int i = 0; i < 6; i++) r_out[i] = R[i];
    *val = u.spline(x, Y, 6, R);
}

void ref_cpp_latlon(double a, double e2, double *xyz, double *llh, int type)
{
    Ellipsoid e(a, e2);
    vector<double> v(xyz, xyz + 3), l(llh, llh + 3);
    e.latlon(v, l, type);
    for (int i = 0; i < 3; i++) { xyz[i] = v[i]; llh[i] = l[i]; }
}
double ref_cpp_reast(double a, double e2, double lat) { Ellipsoid e(a, e2); return e.reast(lat); }
double ref_cpp_rnorth(double a, double e2, double lat) { Ellipsoid e(a, e2); return e.rnorth(lat); }
double ref_cpp_rdir(double a, double e2, double hdg, double lat) { Ellipsoid e(a, e2); return e.rdir(hdg, lat); }

void ref_cpp_tcnbasis(const double *pos, const double *vel, double a, double e2, double *t, double *c, double *n)
{
    Ellipsoid e(a, e2);
    vector<double> p(pos, pos + 3), v(vel, vel + 3), tt(3), cc(3), nn(3);
    e.tcnbasis(p, v, tt, cc, nn);
    for (int i = 0; i < 3; i++) { t[i] = tt[i]; c[i] = cc[i]; n[i] = nn[i]; }
}

void ref_cpp_enubasis(double lat, double lon, double *m)
{
    LinAlg la;
    vector<vector<double> > e(3, vector<double>(3));
    la.enubasis(lat, lon, e);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) m[3 * i + j] = e[i][j];
}

// peg transform at (lat, lon, hdg): mat[9] row-major, ov[3], returns radcur
double ref_cpp_radar_to_xyz(double a, double e2, double lat, double lon, double hdg, double *mat, double *ov)
{
    Ellipsoid e(a, e2);
    Peg peg(lat, lon, hdg);
    PegTrans pt;
    pt.radar_to_xyz(e, peg);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) mat[3 * i + j] = pt.mat[i][j];
        ov[i] = pt.ov[i];
    }
    return pt.radcur;
}

// XYZ -> SCH (type 1) or SCH -> XYZ (type 0) about the peg (lat, lon, hdg)
void ref_cpp_convert_sch(double a, double e2, double lat, double lon, double hdg, double *sch, double *xyz, int type)
{
    Ellipsoid e(a, e2);
    Peg peg(lat, lon, hdg);
    PegTrans pt;
    pt.radar_to_xyz(e, peg);
    vector<double> s(sch, sch + 3), x(xyz, xyz + 3);
    pt.convert_sch_to_xyz(s, x, type);
    for (int i = 0; i < 3; i++) { sch[i] = s[i]; xyz[i] = x[i]; }
}

static void fill_orbit(Orbit &o, int nvec, const double *t, const double *pos, const double *vel)
{
    o.setOrbit(nvec, WGS84_ORBIT);
    for (int i = 0; i < nvec; i++) {
        vector<double> p(pos + 3 * i, pos + 3 * i + 3), v(vel + 3 * i, vel + 3 * i + 3);
        o.setStateVector(i, t[i], p, v);
    }
}

int ref_cpp_interp_orbit(int nvec, const double *t, const double *pos, const double *vel, int method, double tq,
                         double *p_out, double *v_out)
{
    Orbit o;
    fill_orbit(o, nvec, t, pos, vel);
    vector<double> p(3), v(3);
    int stat = o.interpolateOrbit(tq, p, v, method);
    for (int i = 0; i < 3; i++) { p_out[i] = p[i]; v_out[i] = v[i]; }
    return stat;
}

double ref_cpp_eval_poly2d(int rorder, int aorder, double mr, double ma, double nr, double na, const double *c,
                           double azi, double rng)
{
    Poly2d p(aorder, rorder); // Poly2d.cpp:12: (azOrder, rgOrder)
    p.meanRange = mr; p.meanAzimuth = ma; p.normRange = nr; p.normAzimuth = na;
    for (int i = 0; i <= aorder; i++)
        for (int j = 0; j <= rorder; j++) p.setCoeff2d(i, j, c[i * (rorder + 1) + j]);
    return p.evalPoly2d(azi, rng);
}

void ref_cpp_insertion_sort(double *a, int n)
{
    LinAlg la;
    vector<double> v(a, a + n);
    la.insertionSort(v, n);
    for (int i = 0; i < n; i++) a[i] = v[i];
}
int ref_cpp_binary_search(const double *a, int n, double val)
{
    LinAlg la;
    vector<double> v(a, a + n);
    return la.binarySearch(v, 0, n - 1, val);
}

// The whole CPU branch of Topo::topo (Topo.cpp:127-365, 600-950).  dem: [idemlength][idemwidth] float32;
// dop, rho: [length][width] double (what the Poly2d-backed accessors deliver line by line);
// lat/lon/hgt: [length][width] double; los/inc: [length][width][2] float32 (pixel-interleaved, as handed to
// setLineSequential); mask: [length][width] double (the reference keeps it in a vector<double>) or NULL.
int ref_cpp_topo(double firstlat, double firstlon, double deltalat, double deltalon, double major, double e2,
                 double peghdg, double prf, double t0, double wvl, double thresh, int numiter, int extraiter,
                 int idemwidth, int idemlength, int ilrl, int length, int width, int nrnglooks, int nazlooks,
                 int dem_method, int orbit_method, int nvec, const double *ot, const double *opos, const double *ovel,
                 float *dem, double *dop, double *rho, double *lat, double *lon, double *hgt, float *los, float *inc,
                 double *mask)
{
    Topo tp;
    tp.firstlat = firstlat; tp.firstlon = firstlon; tp.deltalat = deltalat; tp.deltalon = deltalon;
    tp.major = major; tp.eccentricitySquared = e2; tp.rspace = 0.0; tp.r0 = rho[0];
    tp.peghdg = peghdg; tp.prf = prf; tp.t0 = t0; tp.wvl = wvl; tp.thresh = thresh;
    tp.numiter = numiter; tp.extraiter = extraiter; tp.idemwidth = idemwidth; tp.idemlength = idemlength;
    tp.ilrl = ilrl; tp.length = length; tp.width = width; tp.Nrnglooks = nrnglooks; tp.Nazlooks = nazlooks;
    tp.dem_method = dem_method; tp.orbit_method = orbit_method; tp.orbit_nvecs = nvec; tp.orbit_basis = WGS84_ORBIT;
    fill_orbit(tp.orb, nvec, ot, opos, ovel);
    MemAccessor a_dem(dem, idemlength, idemwidth, 1, 4), a_dop(dop, length, width, 1, 8), a_rho(rho, length, width, 1, 8);
    MemAccessor a_lat(lat, length, width, 1, 8), a_lon(lon, length, width, 1, 8), a_hgt(hgt, length, width, 1, 8);
    MemAccessor a_los(los, length, width, 2, 4), a_inc(inc, length, width, 2, 4), a_mask(mask, length, width, 1, 8);
    tp.demAccessor = (uint64_t)(DataAccessor *)&a_dem;
    tp.dopAccessor = (uint64_t)(DataAccessor *)&a_dop;
    tp.slrngAccessor = (uint64_t)(DataAccessor *)&a_rho;
    tp.latAccessor = (uint64_t)(DataAccessor *)&a_lat;
    tp.lonAccessor = (uint64_t)(DataAccessor *)&a_lon;
    tp.heightAccessor = (uint64_t)(DataAccessor *)&a_hgt;
    tp.losAccessor = los ? (uint64_t)(DataAccessor *)&a_los : 0;
    tp.incAccessor = inc ? (uint64_t)(DataAccessor *)&a_inc : 0;
    tp.maskAccessor = mask ? (uint64_t)(DataAccessor *)&a_mask : 0;
    tp.topo();
    return 0;
}
}
