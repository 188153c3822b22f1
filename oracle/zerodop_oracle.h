/*
 * zerodop_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99 + OpenMP, FP64 with the reference's float32 casts)
 * of ISCE2's zero-Doppler topozero (rdr2geo) and geo2rdr hot path.  It exists to
 * CHECK the CUDA product (tests/, __graft_entry__.smoke(), bench.py cpu_baseline /
 * --impl reference).  Nothing under isce2_b200/ may include, link or call it.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  Pinning status (DESIGN.md section 2 has the table): the
 * Fortran itself cannot be compiled in this image (no gfortran); what the
 * reference ships in C / C++ for the same path is compiled UNCHANGED into
 * oracle/_ref and the restatement is held to it bit for bit -- orbit.c,
 * orbitHermite.c, poly1d.c, poly2d.c, linalg3.c (tests/test_oracle_pins.py) and
 * the reference's own C++ restatement of topozero / geo2rdr, CPU branches
 * (GPUtopozero/src, GPUgeo2rdr/src: DEM interpolators, geometry primitives,
 * whole-image topo -- lat/lon/hgt/los/local incidence/shadow bit -- and
 * whole-image geo2rdr; GPUresampslc/src: whole-image resamp_slc up to float32
 * accumulation order -- tests/test_oracle_cpp_pins.py).  Restatement-only, because
 * the C++ departs from the Fortran there: the layover bit of the mask, inc
 * channel 1 (psi), the bicubic interpolator, binarysearch, and geozero (no C++
 * form of it exists); those are held
 * against the defining equations, the geometric definitions of the mask bits
 * and independent implementations (tests/test_oracle_image_properties_cpu.py)
 * and golden vectors of the reference's own Python (tests/golden/).
 */
#ifndef ZERODOP_ORACLE_H
#define ZERODOP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* DEM interpolation ids: components/zerodop/topozero/src/topozeroMethods.f:30-32 */
enum { ORC_SINC = 0, ORC_BILINEAR = 1, ORC_BICUBIC = 2, ORC_NEAREST = 3, ORC_AKIMA = 4, ORC_BIQUINTIC = 5 };
/* orbit interpolation ids: components/zerodop/topozero/src/topozeroState.f:77-78 */
enum { ORC_HERMITE = 0, ORC_SCH = 1, ORC_LEGENDRE = 2 };

/* cOrbit without the date string: components/isceobj/Util/Library/orbit/include/orbit.h:30-38 */
typedef struct {
    int nvec;
    const double *t;   /* [nvec] seconds of day */
    const double *pos; /* [nvec*3] ECEF m */
    const double *vel; /* [nvec*3] ECEF m/s */
} orc_orbit;

/* cPoly2d: components/isceobj/Util/Library/poly2d/include/poly2d.h:25-34 */
typedef struct {
    int range_order, azimuth_order;
    double mean_range, mean_azimuth, norm_range, norm_azimuth;
    const double *coeffs; /* [(azimuth_order+1)*(range_order+1)], row = azimuth */
} orc_poly2d;

/* cPoly1d: components/isceobj/Util/Library/poly1d/include/poly1d.h:25-31 */
typedef struct {
    int order;
    double mean, norm;
    const double *coeffs;
} orc_poly1d;

/* module topozeroState: components/zerodop/topozero/src/topozeroState.f:32-68 */
typedef struct {
    int numiter, extraiter;
    double thresh;
    int idemwidth, idemlength;
    double firstlat, firstlon, deltalat, deltalon;
    double major, e2;
    int length, width;
    int nrnglooks, nazlooks;
    double peghdg, prf, t0, wvl;
    int ilrl;
    int method;      /* DEM interpolation id */
    int orbitmethod; /* orbit interpolation id */
} orc_topo_params;

typedef struct {
    double min_lat, max_lat, min_lon, max_lon;
    long long totalconv;
    long long total_iters; /* sum over pixels of executed iteration bodies (for K) */
    int ustartx, ustarty, udemwidth, udemlength;
    double ufirstlat, ufirstlon;
    float demmax;
} orc_topo_result;

/* module geo2rdrState: components/zerodop/geo2rdr/src/geo2rdrState.F:28-67 */
typedef struct {
    double major, e2;
    double drho, rho0;
    double wvl, t0, prf;
    int length, width;
    int ilrl;
    int nrnglooks, nazlooks;
    int demwidth, demlength;
    int bistatic;
    int orbitmethod;
} orc_geo_params;

typedef struct {
    long long num_outside, num_valid, num_conv;
    long long total_iters; /* sum over pixels of Newton steps (for N) */
} orc_geo_result;

/* ---- primitives (exported for KATs) ---- */
void orc_latlon(double a, double e2, double *xyz, double *llh, int type); /* 1 = LLH->XYZ, 2 = XYZ->LLH */
double orc_reast(double a, double e2, double lat);
double orc_rnorth(double a, double e2, double lat);
double orc_rdir(double a, double e2, double hdg, double lat);
void orc_tcnbasis(const double *pos, const double *vel, double a, double e2, double *t, double *c, double *n);
void orc_enubasis(double lat, double lon, double *enumat /* [3][3] row-major, enumat[i][j] = Fortran r_enumat(i+1,j+1) */);
/* peg transform: fills mat[9] (row-major Fortran r_mat(i,j)), ov[3], returns radcur */
double orc_radar_to_xyz(double a, double e2, double lat, double lon, double hdg, double *mat, double *ov);
void orc_xyz_to_sch(const double *mat, const double *ov, double radcur, const double *xyz, double *sch);

int orc_interp_hermite(const orc_orbit *o, double t, double *pos, double *vel);
int orc_interp_legendre(const orc_orbit *o, double t, double *pos, double *vel);
int orc_interp_sch(const orc_orbit *o, double t, double *pos, double *vel);
int orc_compute_acceleration(const orc_orbit *o, double t, double *acc);
double orc_eval_poly2d(const orc_poly2d *p, double azi, double rng);
double orc_eval_poly1d(const orc_poly1d *p, double x);

/* dem is [ny][nx] row-major float32 (== Fortran dem(nx,ny)); ix, iy are the reference's 1-based indices */
float orc_interp_dem(int method, const float *dem, int ix, int iy, double fx, double fy, int nx, int ny);

void orc_interp_dem_batch(int method, const float *dem, int nx, int ny, long n, const int *ix, const int *iy,
                          const double *fx, const double *fy, float *out);
/* test hook of the C++ pin, see zerodop_oracle.c (never set outside tests/test_oracle_cpp_pins.py) */
void orc_test_set_cpp_quirks(int on);
void orc_test_set_sinc_table(const float *table); /* [8192][8] real*4, or NULL for the built-in one */

void orc_sinc_table(float *out /* [8192*8] fintp of topozeroMethods.f:57-61 */);
void orc_insertion_sort(double *a, double *b, double *c, int n);
int orc_binarysearch(const double *arr, int n, double val);

/* ---- the two verbs ----
 * topo: dem_full is the whole DEM [idemlength][idemwidth] float32 (what the 'read' FLOAT caster delivers).
 * Outputs (any may be NULL except lat/lon/hgt): lat/lon/hgt [length][width] double; los/inc [length][2][width]
 * float32 BIL; mask [length][width] int8.  slrng==NULL is an error; rho_image (if non-NULL, [length][width])
 * replaces the slant-range polynomial (slantRangeFilename case).
 * line0/nlines select a block of azimuth lines (bbox + DEM crop always use the whole scene).
 * Returns 0, or a negative error code.
 */
int orc_topo(const orc_topo_params *p, const float *dem_full, const orc_orbit *orb,
             const orc_poly2d *dop, const orc_poly2d *slrng, const double *rho_image,
             int line0, int nlines,
             double *lat, double *lon, double *hgt, float *los, float *inc, int8_t *mask,
             orc_topo_result *res, int nthreads);

/* geo2rdr: lat/lon (deg), hgt (m) [demlength][demwidth] double; outputs double (caller narrows), any may be NULL */
int orc_geo2rdr(const orc_geo_params *p, const double *lat, const double *lon, const double *hgt,
                const orc_orbit *orb, const orc_poly1d *dop,
                int line0, int nlines,
                double *azt, double *rgm, double *azoff, double *rgoff,
                orc_geo_result *res, int nthreads);

/* ---- geozero (geocoding on the zero-Doppler geometry): components/zerodop/geozero/src ---- */
/* module geozeroState: components/zerodop/geozero/src/geozeroState.F:36-78 */
typedef struct {
    double major, e2;
    double min_lat, min_lon, max_lat, max_lon; /* degrees */
    double drho, rho0;
    double wvl, t0, prf;
    int length, width; /* radar image to be geocoded */
    int nrnglooks, nazlooks;
    double lat_first, lon_first, dlat, dlon; /* DEM, degrees */
    int demwidth, demlength;
} orc_geozero_params;

typedef struct {
    int geowidth, geolength;
    double geomin_lat, geomax_lat, geomin_lon, geomax_lon;
    long long num_outside_dem, num_outside_image, num_valid;
    long long total_iters;
} orc_geozero_result;

/* grid of the output (geozero.f90:168-177): returns geo_len / geo_wid and the four index origins */
void orc_geozero_grid(const orc_geozero_params *p, int *geo_len, int *geo_wid, int *min_lat_idx, int *max_lat_idx,
                      int *min_lon_idx, int *max_lon_idx);

/* One band (geozero.f90:1-435).  dem: whole DEM [demlength][demwidth] float32 (the 'read' FLOAT caster);
 * in: the band [length][width] as interleaved complex float32 (re,im) when iscomplex, else float32;
 * out: [geo_len][geo_wid] of the same kind; dem_crop: [geo_len][geo_wid] int16 or NULL;
 * az_idx / rng_idx (optional, [geo_len][geo_wid] double): the fractional 1-based image coordinates of every
 * pixel that reached the interpolator, NaN elsewhere (diagnostics for the parity tests).
 * method: ORC_SINC / ORC_BILINEAR / ORC_BICUBIC / ORC_NEAREST.  lookSide: -1 right, +1 left. */
int orc_geozero(const orc_geozero_params *p, const float *dem, const orc_orbit *orb, const orc_poly1d *dop,
                const float *in, int iscomplex, int method, int lookSide, float *out, int16_t *dem_crop,
                double *az_idx, double *rng_idx, orc_geozero_result *res, int nthreads);

/* ---- resamp_slc (consumer of the geo2rdr offsets): components/stdproc/stdproc/resamp_slc/src ---- */
/* module resamp_slcState: resamp_slcState.F (sizes, WVL / SLR / R0 and their reference-image counterparts, flatten) */
typedef struct {
    int inwidth, inlength, outwidth, outlength;
    double wvl, slr, r0;
    double refwvl, refr0, refslr;
    int flatten;
} orc_resamp_params;

/* the normalised sinc table of resamp_slcMethods.f:57-83 ([8192][8] real*4; NOT the table of topozero / geozero) */
void orc_resamp_sinc_table(float *out);

/* resamp_slc.f90:1-295, complex data (the only branch the reference implements).  in / out: interleaved complex
 * float32 [length][width]; residaz / residrg: [outlength][outwidth] double (what the 'read' DOUBLE caster delivers
 * from the .off rasters) or NULL; any polynomial may be NULL (= the zero polynomial Resamp_slc.py substitutes). */
int orc_resamp_slc(const orc_resamp_params *p, const orc_poly2d *rgCarrier, const orc_poly2d *azCarrier,
                   const orc_poly2d *rgOffsetsPoly, const orc_poly2d *azOffsetsPoly, const orc_poly2d *dopplerPoly,
                   const float *in, const double *residaz, const double *residrg, float *out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
