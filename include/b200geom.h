/*
 * b200geom.h -- C ABI of libb200geom.so: the B200-native (sm_100a, hand-written FP64 CUDA)
 * implementation of ISCE2's zero-Doppler radar-geometry hot path: topozero (rdr2geo) + geo2rdr.
 *
 * This is the drop-in boundary.  It replaces the two CPython extension modules the reference
 * Component classes drive (all paths relative to the ISCE2 tree):
 *
 *   components/zerodop/topozero/bindings/topozeromodule.cpp:73-83    topo_Py(dem, dop, slrng)
 *   components/zerodop/topozero/include/topozeromodule.h:117-155     the 31 set*_Py / get*_Py functions
 *   components/zerodop/geo2rdr/bindings/geo2rdrmodule.cpp:72-88      geo2rdr_Py(lat, lon, hgt, az, rg, azoff, rgoff)
 *   components/zerodop/geo2rdr/include/geo2rdrmodule.h:43-80         the 18 set*_Py functions
 *
 * Differences by design: no module-global state (the reference keeps every parameter in Fortran
 * module variables, topozeroState.f:32-68 / geo2rdrState.F:28-67; here they travel in a params
 * struct, so the library is re-entrant and can drive several GPUs from several threads), images
 * are plain host pointers + sizes instead of uint64 DataAccessor handles, and errors come back
 * as a status code + message instead of a Fortran `stop`.
 *
 * Only plain C types cross this boundary (no torch, no numpy, no CUDA types).  Every entry point
 * returns B200_OK (0) or a negative B200_E* code and, if `err` is non-NULL, a NUL-terminated
 * message.  There is NO CPU fallback: without a CUDA device every compute call fails with
 * B200_ENODEVICE.
 */
#ifndef B200GEOM_H
#define B200GEOM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GEOM_ABI_VERSION 7

enum {
    B200_OK = 0,
    B200_EINVAL = -1,    /* bad argument (message says which) */
    B200_ENODEVICE = -2, /* no usable CUDA device / bad device ordinal */
    B200_ECUDA = -3,     /* CUDA runtime error (message carries cudaGetErrorString) */
    B200_EORBIT = -4,    /* too few state vectors for the interpolation method, or scene centre outside the orbit */
    B200_EDEM = -5,      /* DEM does not cover the scene */
    B200_ENOMEM = -6
};

/* DEM interpolation ids == Topo.interpolationMethods (Topozero.py:46-51, topozeroMethods.f:30-32) */
enum { B200_DEM_SINC = 0, B200_DEM_BILINEAR = 1, B200_DEM_BICUBIC = 2, B200_DEM_NEAREST = 3, B200_DEM_AKIMA = 4,
       B200_DEM_BIQUINTIC = 5 };
/* orbit interpolation ids == Topo.orbitInterpolationMethods (Topozero.py:53-55, topozeroState.f:77-78) */
enum { B200_ORBIT_HERMITE = 0, B200_ORBIT_SCH = 1, B200_ORBIT_LEGENDRE = 2 };
/* DEM sample types accepted (the reference reads any type through a 'FLOAT' caster, Topozero.py:380) */
enum { B200_DEM_F32 = 0, B200_DEM_I16 = 1 };

/* replaces cOrbit* given to setOrbit_Py (components/isceobj/Util/Library/orbit/include/orbit.h:30-38);
 * rows as produced by Orbit.exportToC (Orbit.py:1060-1081): t = seconds of the reference day, ECEF m, m/s */
typedef struct {
    int nvec;
    const double *t;   /* [nvec] */
    const double *pos; /* [nvec][3] */
    const double *vel; /* [nvec][3] */
} b200_orbit;

/* replaces the cPoly2d-backed accessor handles (dopAccessor / slrngAccessor of topo_Py):
 * components/isceobj/Util/Library/poly2d/include/poly2d.h:25-34; evaluated at 0-based (row, col) */
typedef struct {
    int range_order, azimuth_order;
    double mean_range, mean_azimuth, norm_range, norm_azimuth;
    const double *coeffs; /* [(azimuth_order+1)][(range_order+1)] */
} b200_poly2d;

/* replaces cPoly1d* given to setDopplerAccessor_Py (poly1d.h:25-31) */
typedef struct {
    int order;
    double mean, norm;
    const double *coeffs; /* [order+1] */
} b200_poly1d;

/* ------------------------------------------------------------------------------------------ */
/* topozero                                                                                    */
/* ------------------------------------------------------------------------------------------ */
/* One field per set*_Py of topozeromodule.h:117-155 (== module topozeroState) */
typedef struct {
    int numiter;         /* setNumberIterations_Py   (default 25, Topozero.py:140) */
    int extraiter;       /* setSecondaryIterations_Py (10) */
    double thresh;       /* setThreshold_Py (0.05 m) */
    int dem_width;       /* setDemWidth_Py  */
    int dem_length;      /* setDemLength_Py */
    double first_lat;    /* setFirstLatitude_Py  (deg, north edge)  */
    double first_lon;    /* setFirstLongitude_Py (deg, west edge)   */
    double delta_lat;    /* setDeltaLatitude_Py  (deg, < 0)         */
    double delta_lon;    /* setDeltaLongitude_Py (deg)              */
    double major;        /* setEllipsoidMajorSemiAxis_Py            */
    double e2;           /* setEllipsoidEccentricitySquared_Py      */
    int length;          /* setLength_Py (radar lines)              */
    int width;           /* setWidth_Py  (radar samples)            */
    int nrnglooks;       /* setNumberRangeLooks_Py                  */
    int nazlooks;        /* setNumberAzimuthLooks_Py                */
    double peg_heading;  /* setPegHeading_Py (rad)                  */
    double prf;          /* setPRF_Py                               */
    double t0;           /* setSensingStart_Py (seconds of day)     */
    double wvl;          /* setRadarWavelength_Py                   */
    int look_side;       /* setLookSide_Py: -1 right, +1 left       */
    int dem_method;      /* setMethod_Py       (B200_DEM_*)         */
    int orbit_method;    /* setOrbitMethod_Py  (B200_ORBIT_*)       */
    /* --- not in the reference: azimuth line block + device selection (multi-GPU sharding) --- */
    int line0;           /* first radar line computed by this call (0-based) */
    int nlines;          /* number of lines (<0: through the last line)      */
    int device;          /* CUDA device ordinal                              */
} b200_topo_params;

/* replaces set{Latitude,Longitude,Height,Los,Inc,Mask}Pointer_Py: caller-owned HOST buffers holding the
 * block's rows; any of los/inc/mask may be NULL (the reference's accessor==0).  Layout == the files the
 * reference writes: lat/lon/hgt [nlines][width] double; los/inc [nlines][2][width] float (BIL,
 * BILAccessor.cpp:11-37); mask [nlines][width] int8 (0 none, 1 shadow, 2 layover, 3 both). */
typedef struct {
    double *lat, *lon, *hgt;
    float *los, *inc;
    int8_t *mask;
} b200_topo_outputs;

typedef struct {
    double min_lat, max_lat, min_lon, max_lon; /* get{Min,Max}imum{Lat,Long}itude_Py, over this block */
    long long converged;                       /* the reference's 'Total convergence' print (topozero.f90:883) */
    long long iterations;                      /* executed iteration bodies, summed over pixels */
    int dem_x0, dem_y0, dem_nx, dem_ny;        /* 1-based crop origin + size actually used (topozero.f90:304-320) */
    float dem_max;
    float ms_setup;   /* device time: bbox + DEM upload/crop + per-line state */
    float ms_kernels; /* device time: per-pixel solve (+ mask), CUDA events on the launch stream */
    float ms_pixels;  /* ... of which the per-pixel kernels (iterative solve + final pass) */
    float ms_solve;   /* ... of which the iterative-solve kernel alone (== ms_pixels when the two are fused) */
    float ms_mask;    /* ... of which the layover/shadow kernel (0 when no mask was requested) */
    float ms_total;   /* wall time of the call, host<->device copies included */
    int gpu_launches; /* kernels launched by this call */
} b200_topo_result;

/* The verb: topo_Py(demAccessor, dopAccessor, slrngAccessor).
 * dem: whole DEM [dem_length][dem_width] of dem_dtype in host memory.  slrng: slant-range polynomial
 * (Topozero.py:337-347) or NULL when rho_image ([length][width] double, the slantRangeFilename case) is given. */
int b200_topo_run(const b200_topo_params *p, const void *dem, int dem_dtype, const b200_orbit *orbit,
                  const b200_poly2d *dop, const b200_poly2d *slrng, const double *rho_image,
                  const b200_topo_outputs *out, b200_topo_result *res, char *err, size_t errlen);

/* Device-resident form of the same path (used to time the kernels with inputs already in HBM, and by
 * callers that consume the layers on the GPU): create uploads + prepares, execute launches the kernels,
 * fetch copies the layers to the host. */
typedef struct b200_topo_plan b200_topo_plan;
int b200_topo_plan_create(const b200_topo_params *p, const void *dem, int dem_dtype, const b200_orbit *orbit,
                          const b200_poly2d *dop, const b200_poly2d *slrng, const double *rho_image,
                          int want_los, int want_inc, int want_mask, b200_topo_plan **plan, char *err, size_t errlen);
int b200_topo_plan_execute(b200_topo_plan *plan, float *ms_kernels, char *err, size_t errlen);
int b200_topo_plan_fetch(b200_topo_plan *plan, const b200_topo_outputs *out, b200_topo_result *res, char *err,
                         size_t errlen);
/* device pointers of the resident layers (NULL if not requested): for chaining into geo2rdr on the GPU */
int b200_topo_plan_device_layers(b200_topo_plan *plan, const double **lat, const double **lon, const double **hgt);
void b200_topo_plan_destroy(b200_topo_plan *plan);

/* ------------------------------------------------------------------------------------------ */
/* geo2rdr                                                                                     */
/* ------------------------------------------------------------------------------------------ */
/* One field per set*_Py of geo2rdrmodule.h:43-80 (== module geo2rdrState) */
typedef struct {
    double major, e2;        /* setEllipsoid*_Py                       */
    double drho;             /* setRangePixelSpacing_Py                */
    double rho0;             /* setRangeFirstSample_Py                 */
    double wvl;              /* setRadarWavelength_Py                  */
    double t0;               /* setSensingStart_Py (seconds of day)    */
    double prf;              /* setPRF_Py                              */
    int length, width;       /* setLength_Py / setWidth_Py (radar grid) */
    int look_side;           /* setLookSide_Py                         */
    int nrnglooks, nazlooks; /* setNumber{Range,Azimuth}Looks_Py       */
    int dem_width;           /* setDemWidth_Py  (samples of lat/lon/hgt) */
    int dem_length;          /* setDemLength_Py (lines of lat/lon/hgt)   */
    int bistatic;            /* setBistaticFlag_Py                     */
    int orbit_method;        /* setOrbitMethod_Py                      */
    /* --- not in the reference --- */
    int line0, nlines;       /* block of lat/lon/hgt lines computed by this call */
    int device;
    int out_f32;             /* 1: outputs are float32 (Geo2rdr outputPrecision 'single': FLOAT image with a
                                DOUBLE write caster, Geo2rdr.py:326-381), 0: float64 */
} b200_geo_params;

/* replaces the four output accessor handles of geo2rdr_Py (0 == NULL == not requested); host buffers
 * [nlines][dem_width] of float or double according to out_f32; invalid pixels = -999999 (geo2rdr.f90:59-60) */
typedef struct {
    void *azt, *rgm, *azoff, *rgoff;
} b200_geo_outputs;

typedef struct {
    long long num_outside, num_valid, num_converged; /* the three prints at geo2rdr.f90:407-409 */
    long long iterations;                            /* Newton steps summed over pixels */
    float ms_setup, ms_kernels, ms_total;
    int gpu_launches;
} b200_geo_result;

/* The verb: geo2rdr_Py(lat, lon, hgt, az, rg, azoff, rgoff).  lat/lon in degrees, hgt in metres,
 * [dem_length][dem_width] double in host memory (whole images; the block is selected by line0/nlines). */
int b200_geo2rdr_run(const b200_geo_params *p, const double *lat, const double *lon, const double *hgt,
                     const b200_orbit *orbit, const b200_poly1d *dop, const b200_geo_outputs *out,
                     b200_geo_result *res, char *err, size_t errlen);

/* Device-resident form: lat/lon/hgt uploaded once (or borrowed from a topo plan on the same device), then any
 * number of secondary orbits run against them (topsStack: one reference geometry x N dates). */
typedef struct b200_geo_plan b200_geo_plan;
int b200_geo_plan_create(const b200_geo_params *p, const double *lat, const double *lon, const double *hgt,
                         b200_geo_plan **plan, char *err, size_t errlen);
int b200_geo_plan_create_from_topo(const b200_geo_params *p, b200_topo_plan *topo, b200_geo_plan **plan, char *err,
                                   size_t errlen);
/* p may differ from the creation params in everything except dem_width/dem_length/line0/nlines/device */
int b200_geo_plan_execute(b200_geo_plan *plan, const b200_geo_params *p, const b200_orbit *orbit,
                          const b200_poly1d *dop, int want_azt, int want_rgm, int want_azoff, int want_rgoff,
                          float *ms_kernels, char *err, size_t errlen);
/* Stack shape only (one reference geometry x N secondary dates): declares that the plan's lat / lon / hgt will not change
 * any more and converts them ONCE to ECEF for the ellipsoid (major, e2); every later b200_geo_plan_execute with that
 * ellipsoid starts from the stored coordinates instead of redoing LLH -> XYZ per pixel and date (the reference redoes it,
 * geo2rdr.f90:247-250; the stored values are the ones the per-pixel path forms, so the outputs are bit-identical).
 * Costs 24 B/pixel of device memory.  Call again after the geometry changed (e.g. the topo plan it borrows from was
 * re-executed). */
int b200_geo_plan_freeze_geometry(b200_geo_plan *plan, double major, double e2, char *err, size_t errlen);
int b200_geo_plan_fetch(b200_geo_plan *plan, const b200_geo_outputs *out, b200_geo_result *res, char *err,
                        size_t errlen);
void b200_geo_plan_destroy(b200_geo_plan *plan);

/* Fused form of the two verbs: topo_Py followed by geo2rdr_Py on the layers it just wrote, the sequence of
 * stripmapApp (runTopo -> runGeo2rdr, components/isceobj/StripmapProc/runGeo2rdr.py:46-110), topsApp (runTopo ->
 * runCoarseOffsets / runFineOffsets' runGeo2rdrCPU, components/isceobj/TopsProc/runFineOffsets.py:16-67) and topsStack (reference geometry -> one
 * geo2rdr per secondary date, contrib/stack/topsStack/geo2rdr.py:233-302).  The reference hands lat / lon / hgt from
 * one to the other through the .rdr files; here every block of lines goes topo kernels -> geo2rdr kernel(s) on the
 * layers still resident in HBM -> one device-to-host stream carrying the topo layers and the offsets, so the 24 B/pixel
 * host-to-device trip of the standalone geo2rdr verb disappears.  Outputs are bit-identical to b200_topo_run followed
 * by b200_geo2rdr_run on its lat / lon / hgt (tests/test_gpu_fused.py).  In each job p->dem_width / dem_length must equal
 * the topo grid; p->line0 / nlines / device are taken from the topo params. */
typedef struct {
    const b200_geo_params *p;
    const b200_orbit *orbit;      /* secondary (or the same) orbit */
    const b200_poly1d *dop;
    const b200_geo_outputs *out;  /* host buffers holding the block's rows, as in b200_geo2rdr_run */
    b200_geo_result *res;         /* may be NULL */
} b200_geo_job;

int b200_topo_geo2rdr_run(const b200_topo_params *p, const void *dem, int dem_dtype, const b200_orbit *orbit,
                          const b200_poly2d *dop, const b200_poly2d *slrng, const double *rho_image,
                          const b200_topo_outputs *out, b200_topo_result *res, int njobs, const b200_geo_job *jobs,
                          char *err, size_t errlen);

/* ------------------------------------------------------------------------------------------ */
/* geozero (geocoding on the zero-Doppler geometry) -- SURVEY 8(f) row N3                      */
/* ------------------------------------------------------------------------------------------ */
/* Replaces the CPython extension components/zerodop/geozero/bindings/geozeromodule.cpp (set*_Py of
 * include/geozeromodule.h, module geozeroState of src/geozeroState.F:36-78) and its verb geozero_Py
 * (src/geozero.f90).  One field per set*_Py. */
typedef struct {
    double major, e2;                          /* setEllipsoid*_Py                            */
    double min_lat, max_lat, min_lon, max_lon; /* set{Minimum,Maximum}{Latitude,Longitude}_Py */
    double drho, rho0;                         /* setRangePixelSpacing_Py, setRangeFirstSample_Py */
    double wvl, t0, prf;                       /* setRadarWavelength_Py, setSensingStart_Py, setPRF_Py */
    int length, width;                         /* setLength_Py / setWidth_Py (image to geocode) */
    int look_side;                             /* setLookSide_Py: -1 right, +1 left           */
    int nrnglooks, nazlooks;                   /* setNumber{Range,Azimuth}Looks_Py            */
    double first_lat, first_lon, delta_lat, delta_lon; /* setFirst*_Py / setDelta*_Py (DEM, degrees) */
    int dem_width, dem_length;                 /* setDemWidth_Py / setDemLength_Py            */
    int device;                                /* not in the reference: CUDA device ordinal   */
} b200_geozero_params;

#define B200_GEOZERO_SINC 0 /* geozeroMethods.F:29-31 */
#define B200_GEOZERO_BILINEAR 1
#define B200_GEOZERO_BICUBIC 2
#define B200_GEOZERO_NEAREST 3
#define B200_SCHEME_BIL 0 /* [line][band][sample] */
#define B200_SCHEME_BIP 1 /* [line][sample][band] */
#define B200_SCHEME_BSQ 2 /* [band][line][sample] */

typedef struct {
    int geo_width, geo_length;                             /* getGeoWidth_Py / getGeoLength_Py            */
    double geo_min_lat, geo_max_lat, geo_min_lon, geo_max_lon; /* get{Min,Max}imumGeo{Lat,Long}itude_Py      */
    long long num_outside_dem, num_outside_image, num_valid;   /* the three prints at geozero.f90:407-409 (of the
                                                                   last geocoded band)                        */
    long long iterations;                                  /* fixed-point steps summed over pixels        */
    float ms_setup;   /* device time: DEM crop upload + orbit polynomials + the per-pixel solve            */
    float ms_solve;   /* ... of which the per-pixel solve kernels alone                                     */
    float ms_kernels; /* device time of the last geocode call's gather kernels                              */
    float ms_total;
    int gpu_launches;
} b200_geozero_result;

/* size of the output grid (geozero.f90:163-170), so that the caller can allocate the output image */
int b200_geozero_grid(const b200_geozero_params *p, int *geo_width, int *geo_length, char *err, size_t errlen);

/* Plan = the geometry of one output grid: uploads the needed part of the DEM ([dem_length][dem_width] of
 * dem_dtype, host memory) and solves every output pixel's (azimuth, range) image coordinate ONCE.  The reference
 * repeats that solve for every band of every product (Geozero.py:216-241). */
typedef struct b200_geozero_plan b200_geozero_plan;
int b200_geozero_plan_create(const b200_geozero_params *p, const void *dem, int dem_dtype, const b200_orbit *orbit,
                             const b200_poly1d *dop, b200_geozero_plan **plan, char *err, size_t errlen);
/* geocode one image: `image` holds nbands bands of [length][width] samples, float32 (is_complex = 0) or
 * interleaved complex64 (is_complex = 1) in the given interleaving scheme; `out` receives the same bands and
 * scheme on the [geo_length][geo_width] grid.  method = B200_GEOZERO_*. */
int b200_geozero_plan_geocode(b200_geozero_plan *plan, const void *image, int is_complex, int nbands, int scheme, int method,
                              void *out, float *ms_kernels, char *err, size_t errlen);
/* cropped DEM (setLineSequential(demCropAccessor, dem_crop), integer*2), the solved image coordinates (1-based,
 * fractional; NaN where the reference skips the pixel) and the counters; any pointer may be NULL */
int b200_geozero_plan_fetch(b200_geozero_plan *plan, int16_t *dem_crop, double *az_idx, double *rng_idx,
                            b200_geozero_result *res, char *err, size_t errlen);
void b200_geozero_plan_destroy(b200_geozero_plan *plan);

/* The verb: geozero_Py(demAccessor, inAccessor, demCropAccessor, outAccessor, inband, outband, iscomplex, method,
 * lookSide), all bands in one call. */
int b200_geozero_run(const b200_geozero_params *p, const void *dem, int dem_dtype, const b200_orbit *orbit,
                     const b200_poly1d *dop, const void *image, int is_complex, int nbands, int scheme, int method,
                     void *out, int16_t *dem_crop, b200_geozero_result *res, char *err, size_t errlen);

/* ------------------------------------------------------------------------------------------ */
/* resamp_slc (consumer of the geo2rdr offsets) -- SURVEY 8(f) row N4                          */
/* ------------------------------------------------------------------------------------------ */
/* Replaces the CPython extension components/stdproc/stdproc/resamp_slc/bindings (set*_Py of resamp_slcmodule.h, module
 * resamp_slcState of src/resamp_slcState.F) and its verb resamp_slc_Py (src/resamp_slc.f90).  Complex data, sinc
 * interpolation: the only branch the reference implements (resamp_slc.f90:69-74, :270-274). */
typedef struct {
    int in_width, in_length;   /* setInputWidth_Py / setInputLines_Py   */
    int out_width, out_length; /* setOutputWidth_Py / setOutputLines_Py */
    double wvl;                /* setRadarWavelength_Py                 */
    double slr;                /* setSlantRangePixelSpacing_Py          */
    double r0;                 /* setStartingRange_Py                   */
    double ref_wvl, ref_r0, ref_slr; /* setReference{Wavelength,StartingRange,SlantRangePixelSpacing}_Py */
    int flatten;               /* setFlatten_Py                         */
    int device;                /* not in the reference: CUDA device ordinal */
} b200_resamp_params;

typedef struct {
    long long num_valid; /* output pixels inside the input image (the others are zero, resamp_slc.f90:197-207) */
    float ms_kernels;    /* device time of the carrier + resampling kernels */
    float ms_total;      /* wall time of the call, copies included */
    int gpu_launches;
} b200_resamp_result;

#define B200_RESID_F64 0 /* residual offsets as double (what the 'read' DOUBLE caster delivers) */
#define B200_RESID_F32 1 /* residual offsets as stored by geo2rdr outputPrecision 'single' (.off FLOAT rasters) */

/* The verb: resamp_slc_Py(slcInAccessor, slcOutAccessor, residazAccessor, residrgAccessor) after
 * set{Range,Azimuth}Carrier_Py, set{Range,Azimuth}OffsetsPoly_Py, setDopplerPoly_Py.  Any polynomial may be NULL (the
 * zero polynomial Resamp_slc.py:86-140 substitutes).  slc_in: [in_length][in_width] interleaved complex float32;
 * resid_az / resid_rg: [out_length][out_width] of resid_dtype, or NULL (accessor == 0); slc_out: [out_length][out_width]. */
int b200_resamp_slc_run(const b200_resamp_params *p, const b200_poly2d *rg_carrier, const b200_poly2d *az_carrier,
                        const b200_poly2d *rg_offsets, const b200_poly2d *az_offsets, const b200_poly2d *doppler,
                        const float *slc_in, const void *resid_az, const void *resid_rg, int resid_dtype, float *slc_out,
                        b200_resamp_result *res, char *err, size_t errlen);

/* Fused form for the stack shape (contrib/stack/topsStack: geo2rdr.py then resamp_withCarrier.py per burst and date): the
 * azimuth / range offsets an executed geo2rdr plan left in HBM are the residual images, without the round trip through the
 * .off rasters (a -999999 offset resamples to zero, exactly as through the files).  The output grid is the plan's. */
int b200_resamp_slc_from_geo_plan(const b200_resamp_params *p, b200_geo_plan *geo, const b200_poly2d *rg_carrier,
                                  const b200_poly2d *az_carrier, const b200_poly2d *rg_offsets, const b200_poly2d *az_offsets,
                                  const b200_poly2d *doppler, const float *slc_in, float *slc_out, b200_resamp_result *res,
                                  char *err, size_t errlen);

/* ------------------------------------------------------------------------------------------ */
/* multilooking of the geometry layers, mask projection -- SURVEY 8(f) row N4, other consumers */
/* ------------------------------------------------------------------------------------------ */
/* element types == the cases of looks_C (components/mroipac/looks/bindings/looksmodule.cpp:80-127) */
enum { B200_T_BYTE = 0, B200_T_SHORT = 1, B200_T_INT = 2, B200_T_LONG = 3, B200_T_FLOAT = 4, B200_T_DOUBLE = 5,
       B200_T_CFLOAT = 6 };
#define B200_LOOKS_AVERAGE 0 /* mroipac.looks: box mean accumulated in double (runMultilook method='isce') */
#define B200_LOOKS_NEAREST 1 /* gdal.Translate -outsize: nearest-neighbour decimation (runMultilook method='gdal') */

typedef struct {
    int out_length, out_width; /* length / down_looks, width / across_looks (integer division, Looks.py:44-45) */
    float ms_kernels;
    float ms_total;
    int gpu_launches;
} b200_looks_result;

/* Replaces looks_Py(inPtr, outPtr, downLooks, acrossLooks, dataType) of mroipac.looks (Looks.py:79) as runMultilook
 * drives it for hgt / lat / lon / los / incLocal / shadowMask / waterMask (contrib/stack/stripmapStack/topo.py:365-441).
 * in: [length][width] x bands of dtype in the given interleaving scheme (B200_SCHEME_*), host memory; out: the same
 * scheme on the [out_length][out_width] grid.  Trailing lines / samples that do not fill a look are dropped. */
int b200_looks_run(const void *in, void *out, int dtype, int length, int width, int bands, int scheme, int down_looks,
                   int across_looks, int method, int device, b200_looks_result *res, char *err, size_t errlen);

/* Fused form for the stack shape (topo.py: runTopo then runMultilook of what it wrote): the multilooked copy of one
 * layer of an executed topo plan, taken from the full-resolution layer still resident in HBM.  layer = B200_LAYER_*;
 * out: host buffer [nlines / down_looks][width / across_looks] of the layer's type (los / inc: two BIL bands).  The
 * plan's block of lines is multilooked on its own: shard by multiples of down_looks. */
enum { B200_LAYER_LAT = 0, B200_LAYER_LON = 1, B200_LAYER_HGT = 2, B200_LAYER_LOS = 3, B200_LAYER_INC = 4, B200_LAYER_MASK = 5 };
int b200_topo_plan_looks(b200_topo_plan *plan, int layer, int down_looks, int across_looks, int method, void *out,
                         b200_looks_result *res, char *err, size_t errlen);

typedef struct {
    float ms_kernels;
    float ms_total;
    int gpu_launches;
} b200_mask_result;

/* Replaces SWBDStitcher.toRadar (contrib/demUtils/swbdstitcher/SWBDStitcher.py:107-131), the geo2radar step of
 * stripmapStack/createWaterMask.py:66-71: out[p] = mask[clip(int((lat[p] - start_lat) / delta_lat), 0, mask_length - 1)]
 * [clip(int((lon[p] - start_lon) / delta_lon), 0, mask_width - 1)] + 1 for the npix pixels of lat.rdr / lon.rdr.
 * mask / out: dtype B200_T_BYTE, SHORT, INT or FLOAT; lat / lon: double (coord_f32 = 0) or float32 (1). */
int b200_mask_to_radar_run(const void *mask, int dtype, int mask_length, int mask_width, double start_lat, double delta_lat,
                           double start_lon, double delta_lon, const void *lat, const void *lon, int coord_f32, size_t npix,
                           void *out, int device, b200_mask_result *res, char *err, size_t errlen);

/* ------------------------------------------------------------------------------------------ */
/* utilities                                                                                   */
/* ------------------------------------------------------------------------------------------ */
int b200_abi_version(void);
int b200_device_count(void);                             /* 0 when no CUDA device is visible */
int b200_device_name(int device, char *buf, size_t len); /* e.g. "NVIDIA B200" */
/* device buffers of finished calls are cached per device for the next call; this returns them to the driver */
void b200_release_cached_memory(void);
void *b200_alloc_pinned(size_t bytes);                   /* page-locked host memory (fast H2D/D2H); NULL on failure */
void b200_free_pinned(void *p);
/* Host buffers handed to any entry point may be page-locked (b200_alloc_pinned, cudaHostRegister: copied by DMA at PCIe
 * speed) or ordinary pageable memory such as a numpy.memmap over the raster being written -- what the reference's
 * callers have (Topozero.py:274-302, Geo2rdr.py:321-384).  Results bound for pageable memory are bounced through a
 * page-locked ring (B200_SINK_SLOTS slots of 32 MB, default 16) and copied out by a pool of B200_COPY_THREADS host
 * threads (environment; default the CPUs of the process, at most 16; 0 = plain cudaMemcpy), so that the page faults of
 * a file that does not exist yet are paid in parallel with the DMA instead of by one thread. */
/* File-backed destinations.  When a pageable output buffer is a shared, writable mapping of a file (numpy.memmap of the
 * raster being written: what the reference's Components hand over, Topozero.py:274-302), the caller may say so: results
 * bound for [base, base + bytes) are then written by the copier threads with pwrite(fd, ..., file_offset + (dst - base))
 * (one call per 32 MB slot) instead of stores through the mapping -- same pages of the page cache, without a page fault
 * and a zeroed page per 4 KB of a file that does not exist yet: 0.9 s instead of 1.8 s for the 16.5 GB of rasters of a
 * Sentinel-1 swath on tmpfs.  The Components do it for every raster they write (B200_FILE_WRITES=0
 * turns that off).  The library keeps its own duplicate of fd until the range is unregistered; a failed write falls back
 * to the store.  Ranges must not overlap; unregister a range only after the calls that use it have returned. */
int b200_host_file_register(const void *base, size_t bytes, int fd, long long file_offset, char *err, size_t errlen);
int b200_host_file_unregister(const void *base); /* B200_OK, or B200_EINVAL when base was not registered */
unsigned long long b200_host_file_bytes(void);   /* bytes written with pwrite since the library was loaded */
/* The same registration serves INPUTS that are mappings of files (the lat / lon / hgt rasters geo2rdr reads,
 * Geo2rdr.py:208-226): they are read with pread into the upload slots instead of through the mapping. */
unsigned long long b200_host_file_bytes_read(void);
/* device -> page-locked host copy of `bytes` (chunks of chunk_bytes, one stream), timed with CUDA events: the floor of
 * an end-to-end call that has to deliver that many bytes of results.  host must hold `bytes`. */
int b200_d2h_floor(int device, void *host, size_t bytes, size_t chunk_bytes, float *ms, char *err, size_t errlen);
/* DFMA-saturating microbenchmark: measured FP64 FMA throughput of `device` in TFLOP/s (2 flop per FMA) */
int b200_fp64_peak(int device, double *tflops, char *err, size_t errlen);
/* single-point primitives evaluated ON THE DEVICE (one thread), for known-answer tests of the device math:
 * what: 0 = LLH(rad)->XYZ, 1 = XYZ->LLH(rad), 2 = Hermite orbit (in[0]=t; out = pos,vel),
 *       3 = Legendre orbit, 4 = SCH orbit; orbit may be NULL for 0/1 */
int b200_device_primitive(int device, int what, double a, double e2, const b200_orbit *orbit, const double *in,
                          double *out, char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* B200GEOM_H */
