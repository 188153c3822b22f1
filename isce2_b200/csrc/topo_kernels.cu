// topo_kernels.cu -- sm_100a kernels of the topozero (rdr2geo) path.
//
//   k_topo_bbox        8 threads: scene-corner geolocation at h = -500 / 9000 m  (topozero.f90:194-257)
//   k_dem_prepare      DEM crop -> float32 (+ max height)                         (topozero.f90:333-345)
//   k_line_setup       one thread per azimuth line: orbit state, TCN basis, peg   (topozero.f90:371-424)
//   k_topo_solve       one thread per radar pixel: iterative height solve          (topozero.f90:458-599)
//   k_topo_final       one thread per radar pixel: final geolocation / LOS /
//                      incidence pass, coalesced layer stores                      (topozero.f90:618-726)
//   k_topo_fused       the two above in one kernel (bilinear / nearest)
//   k_topo_mask        one CTA per azimuth line: layover / shadow mask            (topozero.f90:729-880)
//
// Compiled with -fmad=false (see geom_device.cuh).
#include "topo_kernels.cuh"

#include <cfloat>
#include <cstdint>
#include <cstdlib>

namespace b2 {

// -------------------------------------------------------------------------------------------------
// small device helpers
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long order_key(double d)
{
    long long b = __double_as_longlong(d);
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
__host__ __device__ inline double order_key_inv(long long k)
{
    long long b = k >= 0 ? k : (k ^ 0x7fffffffffffffffLL);
#ifdef __CUDA_ARCH__
    return __longlong_as_double(b);
#else
    double d;
    memcpy(&d, &b, sizeof d);
    return d;
#endif
}
double stats_decode(long long k) { return order_key_inv(k); }

__device__ __forceinline__ double warp_min(double v)
{
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum(int v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Running bounding box of the block (:712-715).  Almost no pixel extends a box that the first lines have opened, so every
// thread first holds its own point against the box as it stands and a warp goes through the reduction and the four
// contended atomics only when one of its pixels pushes an edge.  The box "as it stands" is fetched by bbox_peek at the
// START of the kernel (lane q of each group of four loads edge q; the values ride in one register pair while the pixel
// is computed and are handed round by shuffles at the end): a stale copy is only ever LESS extreme than the truth, so the
// test errs on the side of updating, and the load's latency is off the kernel's tail (it was 4.5 % of the final pass's
// stall samples when it sat in bbox_update).
__device__ __forceinline__ long long bbox_peek(const TopoStats *stats)
{
    const long long *edge = &stats->min_lat + (threadIdx.x & 3); // min_lat, max_lat, min_lon, max_lon are consecutive
    return __ldcg(edge);
}
__device__ __forceinline__ void bbox_update(TopoStats *stats, long long peek, bool have, double mnlat, double mxlat, double mnlon,
                                            double mxlon)
{
    const int base = threadIdx.x & 28;
    const long long a = __shfl_sync(0xffffffffu, peek, base), b = __shfl_sync(0xffffffffu, peek, base + 1),
                    c = __shfl_sync(0xffffffffu, peek, base + 2), d = __shfl_sync(0xffffffffu, peek, base + 3);
    const bool push = have && (order_key(mnlat) < a || order_key(mxlat) > b || order_key(mnlon) < c || order_key(mxlon) > d);
    if (!__any_sync(0xffffffffu, push)) return;
    mnlat = warp_min(mnlat); mxlat = warp_max(mxlat); mnlon = warp_min(mnlon); mxlon = warp_max(mxlon);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&stats->min_lat, order_key(mnlat));
        atomicMax(&stats->max_lat, order_key(mxlat));
        atomicMin(&stats->min_lon, order_key(mnlon));
        atomicMax(&stats->max_lon, order_key(mxlon));
    }
}

__device__ __forceinline__ double line_time(const TopoConst &C, int line0based)
{
    // tline = t0 + Nazlooks*(line - 1.0d0)/prf with the reference's 1-based line (topozero.f90:371)
    return C.t0 + C.nazlooks * ((double)(line0based + 1) - 1.0) / C.prf;
}

__device__ __forceinline__ double pixel_range(const TopoConst &C, int line0based, int pix0based)
{
    if (C.rho_image) return C.rho_image[(size_t)line0based * (size_t)C.width + (size_t)pix0based];
    return eval_poly2d(C.slr, (double)line0based, (double)pix0based); // Poly2dInterpolator.cpp:5-36
}

// -------------------------------------------------------------------------------------------------
// bbox corners
// -------------------------------------------------------------------------------------------------
__global__ void k_topo_bbox(const __grid_constant__ TopoConst C, OrbitView orb, double *out /*[8][3]: lat_deg, lon_deg, ok*/)
{
    int tid = threadIdx.x;
    if (tid >= 8) return;
    int line = tid >> 2;      // 0: first, 1: last line
    int ind = (tid >> 1) & 1; // 0: near, 1: far range
    int it = tid & 1;         // 0: MIN_H, 1: MAX_H
    const double hgts[2] = {-500.0, 9000.0}; // topozeroState.f:74-75
    double tline = C.t0 + line * C.nazlooks * (C.length - 1.0) / C.prf; // :201
    Vec3 pos, vel;
    int stat = orbit_interp(C.orbit_method, orb, tline, pos, vel);
    // the reference leaves the loop at the first failing line (:204-207): line 1 failing also skips line 2
    Vec3 p0, v0;
    int stat0 = orbit_interp(C.orbit_method, orb, C.t0, p0, v0);
    if (stat != 0 || stat0 != 0) {
        out[3 * tid + 2] = 0.0;
        return;
    }
    LineState L;
    make_line_state(C.elp, pos, vel, C.peghdg, L);
    int pixel = ind * (C.width - 1);
    double rng = pixel_range(C, 0, pixel); // the bbox stage reads row 1 of both accessors (:196-197)
    double dop = eval_poly2d(C.dop, 0.0, (double)pixel);
    double la, lo, h;
    if (rng <= (L.height - hgts[it] + 1.0)) { // near-nadir: pick the nadir point (:232-235)
        la = L.lat_sat;
        lo = L.lon_sat;
    } else {
        double ct, st;
        Vec3 delta, xyz;
        PixelConst P = make_pixel_const(C, L, rng, dop);
        range_sphere(C, L, P, hgts[it], ct, st, delta, xyz);
        xyz_to_llh(C.elp, xyz, la, lo, h);
    }
    out[3 * tid + 0] = la * C.r2d;
    out[3 * tid + 1] = lo * C.r2d;
    out[3 * tid + 2] = 1.0;
}

// -------------------------------------------------------------------------------------------------
// DEM crop -> float32 and its maximum
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_float(float *addr, float v)
{
    // valid for any sign mix: order-preserving integer view
    int b = __float_as_int(v);
    int key = b >= 0 ? b : (b ^ 0x7fffffff);
    atomicMax((int *)addr, key);
}
float dem_max_decode(int key)
{
    int b = key >= 0 ? key : (key ^ 0x7fffffff);
    float f;
    memcpy(&f, &b, sizeof f);
    return f;
}

__global__ void k_dem_prepare(const void *raw, int dtype, float *dem, size_t n, int *maxkey)
{
    float m = -FLT_MAX;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = (dtype == 1) ? (float)((const short *)raw)[i] : ((const float *)raw)[i];
        dem[i] = v;
        m = fmaxf(m, v);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomic_max_float((float *)maxkey, m);
}

// padded double copy of the crop for the biquintic window (DemView::d64): out[r][c] = dem[min(r, ny-1)][min(c, nx-1)]
__global__ void k_dem_pad64(const float *dem, int nx, int ny, double *out, int stride)
{
    const size_t n = (size_t)(ny + 1) * (size_t)stride;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(i / (size_t)stride), c = (int)(i - (size_t)r * (size_t)stride);
        r = r < ny ? r : ny - 1;
        c = c < nx ? c : nx - 1;
        out[i] = (double)dem[(size_t)r * (size_t)nx + (size_t)c];
    }
}

// -------------------------------------------------------------------------------------------------
// per-line state
// -------------------------------------------------------------------------------------------------
__global__ void k_line_setup(const __grid_constant__ TopoConst C, OrbitView orb, int line0, int nlines, LineState *states)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlines) return;
    Vec3 pos = v3(0, 0, 0), vel = v3(0, 0, 0);
    orbit_interp(C.orbit_method, orb, line_time(C, line0 + i), pos, vel); // stat ignored as in :378
    LineState L;
    make_line_state(C.elp, pos, vel, C.peghdg, L);
    states[i] = L;
}

// -------------------------------------------------------------------------------------------------
// per-pixel solve
// -------------------------------------------------------------------------------------------------
#ifndef B2_TOPO_MINBLOCKS
#define B2_TOPO_MINBLOCKS 8
#endif
// the split solve kernel: 7 CTAs / SM = 72 registers; at 64 it spills 130 bytes into the iteration loop (measured
// 9.40 vs 9.58 ms per 1500 x 25000 lines with the biquintic interpolator)
#ifndef B2_SOLVE_MINBLOCKS
#define B2_SOLVE_MINBLOCKS 7
#endif
#ifndef B2_FINAL_MINBLOCKS
#define B2_FINAL_MINBLOCKS 6
#endif

// Stage the line's state in shared memory (one azimuth line per CTA row segment)
__device__ __forceinline__ void load_line_state(LineState &sL, const LineState *__restrict__ states, int row)
{
    const double *src = reinterpret_cast<const double *>(states + row);
    double *dst = reinterpret_cast<double *>(&sL);
    for (int i = threadIdx.x; i < (int)(sizeof(LineState) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
}

// The solve and the final pass are two kernels: each keeps its own (small) instruction working set resident in the
// instruction cache and its own register budget; the only state handed over is the converged SCH height (8 B/pixel,
// parked in the ctrack layer the final pass overwrites anyway).
//
// k_topo_solve: iterative height solve against the DEM (topozero.f90:458-599).
//
// Pixels converge after different numbers of iterations (3..10 on ordinary terrain, 36 in layover), so a fixed
// one-thread-per-pixel mapping leaves a quarter of the lanes idle (ncu: 24.5 of 32 active).  Instead each warp owns a
// run of consecutive pixels of one line and its lanes pull the next unsolved pixel of the run as soon as their
// current one converges (ballot + prefix popcount; no atomics, no shared counters).  The slant ranges and Doppler
// values of the upcoming pixels are evaluated ahead into shared memory so that a refill costs a few instructions:
//   * k_topo_fused: the run is a strip of kStrip pixels staged up front (the final pass needs them again);
//   * k_topo_solve: the run is a segment of up to kSegMax pixels streamed through a ring of kRing entries, which
//     makes the idle tail at the end of a run (lanes waiting for the last pixels) 8x rarer than with 128-pixel strips.
// tunables of experimental builds (tools/build_variant.sh NAME "-DB2_SEG_MAX=2048"); the defaults are the measured best
#ifndef B2_RING
#define B2_RING 64
#endif
#ifndef B2_SEG_MAX
#define B2_SEG_MAX 1024
#endif
constexpr int kStrip = 128;                                    // pixels per warp (fused kernel)
constexpr int kSolvePixelsPerCta = (kTopoBlock / 32) * kStrip; // 512
constexpr int kRing = B2_RING;                                 // staged look-ahead per warp (split solve kernel), power of two
constexpr int kStage = 32;                                     // pixels staged per refill of the ring
constexpr int kSegMax = B2_SEG_MAX;                            // longest run of pixels per warp
static_assert((kRing & (kRing - 1)) == 0 && kRing >= 2 * kStage, "the ring is indexed with a mask and refilled kStage at a time");

// Per-pixel constants of the upcoming pixels, evaluated with all lanes busy (a refill inside the solve loop typically
// has only a handful of lanes active, so everything hoisted here is paid at 1/6 of the price).
struct StagedPixels {
    double *rng, *inv_rng, *dopfact, *a12; // fused kernel (STREAM = false): only rng and the raw Doppler (in dopfact)
};

template <bool STREAM>
__device__ __forceinline__ void stage_pixels(const TopoConst &C, const LineState &L, int line, int pix0, int from, int to,
                                             const StagedPixels &S)
{
    for (int j = from + (threadIdx.x & 31); j < to; j += 32) {
        const double rng = pixel_range(C, line, pix0 + j), dop = eval_poly2d(C.dop, (double)line, (double)(pix0 + j));
        if (STREAM) {
            const int k = j & (kRing - 1);
            const PixelConst P = make_pixel_const(C, L, rng, dop);
            S.rng[k] = P.rng;
            S.inv_rng[k] = P.inv_rng;
            S.dopfact[k] = P.dopfact;
            S.a12[k] = P.a12;
        } else {
            S.rng[j] = rng;
            S.dopfact[j] = dop;
        }
    }
}

template <bool STREAM>
__device__ __forceinline__ PixelConst staged_pixel(const TopoConst &C, const LineState &L, const StagedPixels &S, int k)
{
    if (!STREAM) return make_pixel_const(C, L, S.rng[k], S.dopfact[k]);
    PixelConst P;
    P.rng = S.rng[k];
    P.inv_rng = S.inv_rng[k];
    P.rng2 = P.rng * P.rng;
    P.dopfact = S.dopfact[k];
    P.a12 = S.a12[k];
    return P;
}

// Solve the strip_n pixels pix0 .. of one warp's run: lanes pull the next unsolved pixel as soon as their current one
// converges.  zrow[j] receives the SCH height of run pixel j.  STREAM: rng_s / dop_s are rings (see above) that this
// function keeps filled; otherwise the caller staged the whole run.
template <int METHOD, bool REF, bool STREAM>
__device__ __forceinline__ void solve_strip(const TopoConst &C, const LineState &sL, int line, int pix0,
                                            const StagedPixels &S, int strip_n, double *zrow, int &conv, int &iters)
{
    const int lane = threadIdx.x & 31;
    const int nprimary = C.numiter + 1 < C.numiter + C.extraiter + 1 ? C.numiter + 1 : C.numiter + C.extraiter + 1;
    int staged = strip_n;
    if (STREAM) {
        staged = strip_n < kRing ? strip_n : kRing;
        stage_pixels<true>(C, sL, line, pix0, 0, staged, S);
        __syncwarp();
    }
    PixelConst P;
    double lat = 0.0, lon = 0.0, z = 0.0, zsch = 0.0;
    int it = 0, slot = lane, next = 32;
    conv = 0;
    iters = 0;
    bool active = slot < strip_n;
    if (active) {
        P = staged_pixel<STREAM>(C, sL, S, slot);
        lat = C.ufirstlat + 0.5 * C.deltalat * C.dem.ny; // :435-436
        lon = C.ufirstlon + 0.05 * C.deltalon * C.dem.nx;
    }
    const int nmax = C.numiter + C.extraiter + 1;
    while (__any_sync(0xffffffffu, active)) {
        bool finished = false;
        if (active) {
            // iterations numiter+2 .. are the reference's secondary iterations (:572-593): the new point is averaged
            // with the previous one.  They run in the same lock-step loop as everybody else's primary iterations
            // (lanes hold pixels in different phases here, so an out-of-line secondary loop would serialise the warp).
            const bool secondary = it >= nprimary;
            Vec3 xyz, xyz_prev = v3(0.0, 0.0, 0.0);
            if (secondary) xyz_prev = geodetic_to_xyz<REF>(C, lat, lon, z);
            it++;
            bool converged = false;
            if (topo_iterate<METHOD, REF>(C, sL, P, lat, lon, z, zsch, xyz)) {
                converged = true;
                finished = true;
            } else {
                if (secondary) {
                    xyz.x = 0.5 * (xyz_prev.x + xyz.x);
                    xyz.y = 0.5 * (xyz_prev.y + xyz.y);
                    xyz.z = 0.5 * (xyz_prev.z + xyz.z);
                    double la, lo, h;
                    if (REF) xyz_to_llh_ref<true>(C.elp, C.ref, xyz, la, lo, h);
                    else xyz_to_llh(C.elp, xyz, la, lo, h);
                    lat = la * C.r2d;
                    lon = lo * C.r2d;
                    z = h;
                    zsch = sch_height(sL, xyz);
                }
                finished = it >= nmax;
            }
            if (finished) { // the statistics are touched once per pixel, not once per iteration (they live in local memory)
                zrow[slot] = zsch;
                iters += it;
                conv += converged ? 1 : 0;
            }
        }
        // lanes without work take the next pixels of the run, in lane order
        const bool need = !active || finished;
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (need) {
            const int j = next + __popc(m & ((1u << lane) - 1u));
            active = j < strip_n;
            if (active) {
                const int k = STREAM ? (j & (kRing - 1)) : j;
                slot = j;
                P = staged_pixel<STREAM>(C, sL, S, k);
                lat = C.ufirstlat + 0.5 * C.deltalat * C.dem.ny;
                lon = C.ufirstlon + 0.05 * C.deltalon * C.dem.nx;
                z = 0.0;
                zsch = 0.0;
                it = 0;
            }
        }
        next += __popc(m);
        if (STREAM) {
            // keep 32 staged pixels ahead of `next`: entries below `next` are consumed, so the kStage new ones may
            // overwrite the ring positions of pixels [staged - kRing, staged - kRing + kStage)
            if (staged < strip_n && next + 32 > staged) {
                const int to = staged + kStage < strip_n ? staged + kStage : strip_n;
                __syncwarp();
                stage_pixels<true>(C, sL, line, pix0, staged, to, S);
                staged = to;
                __syncwarp();
            }
        }
    }
}

// One warp per segment: a line is cut into segs_per_line equal runs of at most kSegMax pixels; warps are numbered
// across lines, so a CTA's four warps may sit on different lines and each keeps its own line state.
__host__ __device__ inline int solve_segs_per_line(int width) { return (width + kSegMax - 1) / kSegMax; }

template <int METHOD, bool REF>
__global__ void __launch_bounds__(kTopoBlock, B2_SOLVE_MINBLOCKS)
k_topo_solve(const __grid_constant__ TopoConst C, const LineState *__restrict__ states, int line0, int nlines,
             double *__restrict__ zsch_out, TopoStats *stats)
{
    __shared__ LineState sL[kTopoBlock / 32];
    __shared__ double s_px[kTopoBlock / 32][4][kRing];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int spl = solve_segs_per_line(C.width);
    const int seg_len = (C.width + spl - 1) / spl;
    const long long gw = (long long)blockIdx.x * (kTopoBlock / 32) + warp;
    const int row = (int)(gw / spl);
    if (row >= nlines) return;
    const int seg = (int)(gw - (long long)row * spl);
    const int pix0 = seg * seg_len;
    const int seg_n = (C.width - pix0) < seg_len ? (C.width - pix0) : seg_len;
    {
        const double *src = reinterpret_cast<const double *>(states + row);
        double *dst = reinterpret_cast<double *>(&sL[warp]);
        for (int i = lane; i < (int)(sizeof(LineState) / sizeof(double)); i += 32) dst[i] = src[i];
    }
    __syncwarp();
    double *zrow = zsch_out + (size_t)row * (size_t)C.width + (size_t)pix0;
    int conv = 0, iters = 0;
    const StagedPixels S{s_px[warp][0], s_px[warp][1], s_px[warp][2], s_px[warp][3]};
    solve_strip<METHOD, REF, true>(C, sL[warp], line0 + row, pix0, S, seg_n, zrow, conv, iters);
    // convergence statistics (:570)
    conv = warp_sum(conv);
    iters = warp_sum(iters);
    if (lane == 0 && iters) {
        if (conv) atomicAdd(&stats->converged, (unsigned long long)conv);
        atomicAdd(&stats->iterations, (unsigned long long)iters);
    }
}

// k_topo_final: one thread per radar pixel, final geolocation / LOS / incidence (topozero.f90:618-726), coalesced
// stores of the output layers (BIL for the two-band float layers)
template <int METHOD, bool REF>
__global__ void __launch_bounds__(kTopoBlock, B2_FINAL_MINBLOCKS)
k_topo_final(const __grid_constant__ TopoConst C, const LineState *__restrict__ states, int line0, TopoLayers out, TopoStats *stats)
{
    __shared__ LineState sL;
    const int bpl = (C.width + kTopoBlock - 1) / kTopoBlock;
    const int row = blockIdx.x / bpl;
    const int seg = blockIdx.x - row * bpl;
    load_line_state(sL, states, row);
    const long long peek = bbox_peek(stats);
    __syncthreads();
    const int pix = seg * blockDim.x + threadIdx.x;
    double mnlat = 1e300, mxlat = -1e300, mnlon = 1e300, mxlon = -1e300;
    if (pix < C.width) {
        const int line = line0 + row;
        const double rng = pixel_range(C, line, pix);
        const double dop = eval_poly2d(C.dop, (double)line, (double)pix);
        const size_t w = (size_t)C.width;
        const size_t o = (size_t)row * w + (size_t)pix;
        PixelResult R;
        topo_final<METHOD, REF>(C, sL, rng, dop, out.ctrack[o], out.inc != nullptr, R);
        out.lat[o] = R.lat;
        out.lon[o] = R.lon;
        out.hgt[o] = R.hgt;
        if (out.los) { // BIL: [line][band][pixel]
            out.los[(size_t)row * 2 * w + pix] = R.los0;
            out.los[(size_t)row * 2 * w + w + pix] = R.los1;
        }
        if (out.inc) {
            out.inc[(size_t)row * 2 * w + pix] = R.inc0;
            out.inc[(size_t)row * 2 * w + w + pix] = R.inc1;
        }
        out.ctrack[o] = R.ctrack;
        if (out.elev) out.elev[o] = R.elev;
        mnlat = mxlat = R.lat;
        mnlon = mxlon = R.lon;
    }
    bbox_update(stats, peek, pix < C.width, mnlat, mxlat, mnlon, mxlon);
}

// k_topo_fused: solve + final pass in one kernel.  Used for the light interpolators (bilinear, nearest), whose whole
// instruction stream fits the instruction cache: measured 13 % faster than the two-kernel form there, while the
// heavy interpolators (biquintic, bicubic) are ~10 % faster split.
template <int METHOD, bool REF>
__global__ void __launch_bounds__(kTopoBlock, B2_TOPO_MINBLOCKS)
k_topo_fused(const __grid_constant__ TopoConst C, const LineState *__restrict__ states, int line0, TopoLayers out, TopoStats *stats)
{
    __shared__ LineState sL;
    __shared__ double s_rng[kTopoBlock / 32][kStrip];
    __shared__ double s_dop[kTopoBlock / 32][kStrip];
    __shared__ double s_z[kTopoBlock / 32][kStrip];
    const int bpl = (C.width + kSolvePixelsPerCta - 1) / kSolvePixelsPerCta;
    const int row = blockIdx.x / bpl;
    const int seg = blockIdx.x - row * bpl;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int line = line0 + row;
    const int strip0 = seg * kSolvePixelsPerCta + warp * kStrip;
    const int strip_n = (C.width - strip0) < kStrip ? (C.width - strip0) : kStrip;
    load_line_state(sL, states, row);
    __syncthreads();
    const StagedPixels S{s_rng[warp], nullptr, s_dop[warp], nullptr};
    stage_pixels<false>(C, sL, line, strip0, 0, strip_n, S);
    __syncwarp();
    int conv = 0, iters = 0;
    solve_strip<METHOD, REF, false>(C, sL, line, strip0, S, strip_n, s_z[warp], conv, iters);
    __syncwarp();
    const long long peek = bbox_peek(stats);
    // final pass over the strip, consecutive lanes on consecutive pixels (coalesced layer stores)
    double mnlat = 1e300, mxlat = -1e300, mnlon = 1e300, mxlon = -1e300;
    const size_t w = (size_t)C.width;
    for (int j = lane; j < strip_n; j += 32) {
        const int pix = strip0 + j;
        const size_t o = (size_t)row * w + (size_t)pix;
        PixelResult R;
        topo_final<METHOD, REF>(C, sL, s_rng[warp][j], s_dop[warp][j], s_z[warp][j], out.inc != nullptr, R);
        out.lat[o] = R.lat;
        out.lon[o] = R.lon;
        out.hgt[o] = R.hgt;
        if (out.los) {
            out.los[(size_t)row * 2 * w + pix] = R.los0;
            out.los[(size_t)row * 2 * w + w + pix] = R.los1;
        }
        if (out.inc) {
            out.inc[(size_t)row * 2 * w + pix] = R.inc0;
            out.inc[(size_t)row * 2 * w + w + pix] = R.inc1;
        }
        if (out.elev) {
            out.ctrack[o] = R.ctrack;
            out.elev[o] = R.elev;
        }
        mnlat = fmin(mnlat, R.lat); mxlat = fmax(mxlat, R.lat);
        mnlon = fmin(mnlon, R.lon); mxlon = fmax(mxlon, R.lon);
    }
    bbox_update(stats, peek, lane < strip_n, mnlat, mxlat, mnlon, mxlon);
    conv = warp_sum(conv);
    iters = warp_sum(iters);
    if (lane == 0 && strip_n > 0) {
        if (conv) atomicAdd(&stats->converged, (unsigned long long)conv);
        atomicAdd(&stats->iterations, (unsigned long long)iters);
    }
}

// -------------------------------------------------------------------------------------------------
// layover / shadow mask: one CTA per azimuth line (persistent over lines)
// -------------------------------------------------------------------------------------------------
// The reference's binarysearch (topozero.f90:933-963) on an ascending array returns
//   clamp(#{ i : arr(i) <= val }, 1, n-1)                                   (1-based `left`)
// (invariant of its loop: arr(left) <= val or left == 1, arr(right) > val or right == n).  The arrays searched here
// are near-uniform (cross-track samples, slant ranges), so the count is found by galloping from a linear guess and
// bisecting the bracket: two or three probes instead of sixteen dependent loads.
template <typename F>
__device__ __forceinline__ int search_count_le(F at0 /*0-based accessor*/, int n, double val, int guess)
{
    int g = guess < 0 ? 0 : (guess > n - 1 ? n - 1 : guess);
    int lo, hi; // at0(lo) <= val < at0(hi) with sentinels lo = -1, hi = n
    if (at0(g) <= val) {
        lo = g;
        hi = g + 1;
        int step = 1;
        while (hi < n && at0(hi) <= val) {
            lo = hi;
            step <<= 1;
            hi = hi + step > n ? n : hi + step;
        }
    } else {
        hi = g;
        lo = g - 1;
        int step = 1;
        while (lo >= 0 && !(at0(lo) <= val)) {
            hi = lo;
            step <<= 1;
            lo = lo - step < -1 ? -1 : lo - step;
        }
    }
    while (hi - lo > 1) {
        int m = (lo + hi) >> 1;
        if (at0(m) <= val) lo = m;
        else hi = m;
    }
    return hi; // number of elements <= val
}
__device__ __forceinline__ int ref_search_result(int count_le, int n)
{
    return count_le < 1 ? 1 : (count_le > n - 1 ? n - 1 : count_le);
}

// ---- CTA-wide exclusive scan of one value per thread (thread order), any associative operator ----
template <typename S, typename Op>
__device__ __forceinline__ S block_excl_scan(S mine, Op op, S identity, S *s_warp /*[32]*/)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    S v = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        S t = v.shfl_up(o);
        if (lane >= o) v = op(t, v);
    }
    if (lane == 31) s_warp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        S w = lane < nw ? s_warp[lane] : identity;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            S t = w.shfl_up(o);
            if (lane >= o) w = op(t, w);
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    S prev = v.shfl_up(1);
    S excl = lane > 0 ? prev : identity;
    if (wid > 0) excl = op(s_warp[wid - 1], excl);
    __syncthreads();
    return excl;
}

struct SD { // a double
    double v;
    __device__ SD shfl_up(int o) const { return SD{__shfl_up_sync(0xffffffffu, v, o)}; }
};
struct SF { // a float
    float v;
    __device__ SF shfl_up(int o) const { return SF{__shfl_up_sync(0xffffffffu, v, o)}; }
};
struct SR { // running minimum with resets: state update x -> (f ? v : min(x, v))
    double v;
    int f;
    __device__ SR shfl_up(int o) const { return SR{__shfl_up_sync(0xffffffffu, v, o), __shfl_up_sync(0xffffffffu, f, o)}; }
};
struct OpMaxD { __device__ SD operator()(SD a, SD b) const { return SD{b.v > a.v ? b.v : a.v}; } };
struct OpMinD { __device__ SD operator()(SD a, SD b) const { return SD{b.v < a.v ? b.v : a.v}; } };
struct OpMaxF { __device__ SF operator()(SF a, SF b) const { return SF{b.v > a.v ? b.v : a.v}; } };
struct OpMinF { __device__ SF operator()(SF a, SF b) const { return SF{b.v < a.v ? b.v : a.v}; } };
struct OpReset { // a applied first, then b
    __device__ SR operator()(SR a, SR b) const { return b.f ? b : SR{b.v < a.v ? b.v : a.v, a.f}; }
};

// Each thread owns a contiguous chunk of the array; forward scans give thread t chunk t, backward scans chunk nt-1-t,
// so that "earlier threads" always means "already processed by the reference's sequential loop".
__device__ __forceinline__ void chunk_bounds(int n, bool reverse, int &b, int &e)
{
    const int nt = blockDim.x;
    const int chunk = (n + nt - 1) / nt;
    const int c = reverse ? nt - 1 - (int)threadIdx.x : (int)threadIdx.x;
    b = c * chunk < n ? c * chunk : n;
    e = b + chunk < n ? b + chunk : n;
}

// Inclusive prefix maximum pm[i] = max(v[0..i]) and inclusive suffix minimum sm[i] = min(v[i..n-1])
__device__ void block_prefix_max_suffix_min(const double *v, int n, double *pm, double *sm, SD *s_warp)
{
    int b, e;
    chunk_bounds(n, false, b, e);
    double m = -INFINITY;
    for (int i = b; i < e; i++) m = v[i] > m ? v[i] : m;
    double run = block_excl_scan(SD{m}, OpMaxD(), SD{-INFINITY}, s_warp).v;
    for (int i = b; i < e; i++) {
        run = v[i] > run ? v[i] : run;
        pm[i] = run;
    }
    chunk_bounds(n, true, b, e);
    m = INFINITY;
    for (int i = b; i < e; i++) m = v[i] < m ? v[i] : m;
    run = block_excl_scan(SD{m}, OpMinD(), SD{INFINITY}, s_warp).v;
    for (int i = e - 1; i >= b; i--) {
        run = v[i] < run ? v[i] : run;
        sm[i] = run;
    }
    __syncthreads();
}

// Stable-sort rank of every element of a nearly sorted array (what the reference's InsertionSort, :910-930, produces):
//   rank(p) = p + #{q > p : v[q] < v[p]} - #{q < p : v[q] > v[p]}
// A later element can only be smaller while the suffix minimum is, an earlier one only larger while the prefix
// maximum is, so each count stops at the edge of the element's own disorder window: O(n) on sorted input, O(n x
// window) in layover, never a full O(n log n) sort.  Returns true when the array was already sorted.
#ifndef B2_RANK_TILE
#define B2_RANK_TILE 4
#endif
#ifndef B2_RANK_TILE_MIN
#define B2_RANK_TILE_MIN 384
#endif
// The counts are formed per group of 32 x R consecutive samples (R = B2_RANK_TILE; lane l of the group's warp holds
// samples g0 + 32 j + l, j < R).  Where the disorder around a group is short -- almost everywhere -- every sample walks its
// own window (two probes and out on sorted stretches).  Where it is long (the candidates the group's extreme keys admit
// reach B2_RANK_TILE_MIN / 2 samples or more beyond the group on either side: layover on steep terrain, windows of
// thousands of samples), the warp
// walks the candidates ONCE for all 32 x R samples: one broadcast load of v[q] feeds 32 x R comparisons held in registers
// where the per-sample walks spend two loads and ten instructions on every pair (measured on fold-over-heavy terrain:
// 19.7 -> 16.7 ms per 1500 lines with the tiled form everywhere, which however cost the bench terrain 0.6 ms -- hence the
// switch).  Beyond a sample's own window the comparisons are false by the monotony of the bounds, so both forms give the
// counts of the definition above.
__device__ bool block_stable_ranks(const double *v, int n, const double *pm, const double *sm, int *rank, int *s_flag)
{
    constexpr int R = B2_RANK_TILE;
    constexpr int G = 32 * R;
    if (threadIdx.x == 0) *s_flag = 1;
    __syncthreads();
    int moved = 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int g0 = wid * G; g0 < n; g0 += nwarps * G) {
        double key[R];
        int r[R];
        double kmax = -INFINITY, kmin = INFINITY;
#pragma unroll
        for (int j = 0; j < R; j++) {
            const int p = g0 + j * 32 + lane;
            key[j] = p < n ? v[p] : __longlong_as_double(0x7ff8000000000000LL); // NaN: compares false everywhere
            r[j] = p;
            kmax = fmax(kmax, key[j]); // fmax / fmin skip NaNs
            kmin = fmin(kmin, key[j]);
        }
        kmax = warp_max(kmax);
        kmin = warp_min(kmin);
        const int gend = g0 + G < n ? g0 + G : n;
        // candidates beyond the group: sm and pm are non-decreasing, so both sets are intervals next to the group.  Two
        // probes tell whether either interval is long; only then are its ends located (bisection) and the tiled walk taken
        constexpr int kFar = B2_RANK_TILE_MIN / 2;
        const bool far_fwd = gend + kFar < n && sm[gend + kFar] < kmax;
        const bool far_bwd = g0 - kFar >= 0 && pm[g0 - kFar] > kmin;
        if (far_fwd || far_bwd) {
            int lo = gend, hi = n; // first q >= gend with !(sm[q] < kmax)
            while (lo < hi) {
                const int m = (lo + hi) >> 1;
                if (sm[m] < kmax) lo = m + 1;
                else hi = m;
            }
            const int fend = lo;
            lo = 0;
            hi = g0; // first q < g0 with pm[q] > kmin
            while (lo < hi) {
                const int m = (lo + hi) >> 1;
                if (pm[m] > kmin) hi = m;
                else lo = m + 1;
            }
            const int bbeg = lo;
            for (int q = g0; q < gend; q++) { // the group against itself
                const double x = v[q];
#pragma unroll
                for (int j = 0; j < R; j++) {
                    const int p = g0 + j * 32 + lane;
                    r[j] += (int)(q > p && x < key[j]) - (int)(q < p && x > key[j]);
                }
            }
            for (int q = gend; q < fend; q++) { // later samples
                const double x = v[q];
#pragma unroll
                for (int j = 0; j < R; j++) r[j] += (int)(x < key[j]);
            }
            for (int q = g0 - 1; q >= bbeg; q--) { // earlier samples
                const double x = v[q];
#pragma unroll
                for (int j = 0; j < R; j++) r[j] -= (int)(x > key[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < R; j++) {
                const int p = g0 + j * 32 + lane;
                if (p < n) {
                    const double k = key[j];
                    int rr = p;
                    for (int q = p + 1; q < n && sm[q] < k; q++) rr += (v[q] < k);
                    for (int q = p - 1; q >= 0 && pm[q] > k; q--) rr -= (v[q] > k);
                    r[j] = rr;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < R; j++) {
            const int p = g0 + j * 32 + lane;
            if (p < n) {
                const int rr = r[j] < 0 ? 0 : (r[j] > n - 1 ? n - 1 : r[j]);
                rank[p] = rr;
                moved |= (rr != p);
            }
        }
    }
    if (moved) *s_flag = 0; // benign race: everybody writes the same value
    __syncthreads();
    const bool sorted = *s_flag != 0;
    __syncthreads();
    return sorted;
}

// Forward scan of the reference (:791-799, :834-842):  aa = v(1); for i = 2..nflag: if v(i) <= aa flag else aa = v(i).
// Flagged samples never exceed aa, so aa is the plain prefix maximum: flag[i] |= bit if v[i] <= max(v[0..i-1]).
template <typename T, typename S, typename Op>
__device__ void block_prefix_max_flags(const T *v, int n, int nflag, unsigned char *flag, unsigned char bit, S *s_warp, Op op)
{
    int b, e;
    chunk_bounds(n, false, b, e);
    T m = -INFINITY;
    for (int i = b; i < e; i++) m = v[i] > m ? v[i] : m;
    T run = block_excl_scan(S{m}, op, S{(T)-INFINITY}, s_warp).v;
    for (int i = b; i < e; i++) {
        const T x = v[i];
        if (i >= 1 && i < nflag && x <= run) flag[i] |= bit;
        run = x > run ? x : run;
    }
    __syncthreads();
}

// Backward scan of the reference (:801-809, :844-852):
//   aa = v(n); for i = n-1..1: if (v(i) >= aa .and. .not. reset(i)) flag else aa = v(i)
// where reset(i) is "already flagged by the forward layover scan" (omask(i) >= 2, :847) and absent for the shadow
// scan.  Per sample the state update is either aa <- min(aa, v) or aa <- v (reset); such updates compose
// associatively (OpReset), which gives the chunked parallel form below.
template <typename T>
__device__ void block_suffix_min_flags(const T *v, int n, const unsigned char *reset, unsigned char resetbit,
                                       unsigned char *flag, unsigned char bit, SR *s_warp)
{
    int b, e;
    chunk_bounds(n, true, b, e);
    double val = INFINITY;
    int hr = 0;
    for (int i = e - 1; i >= b; i--) {
        const bool rs = (i == n - 1) || (reset && (reset[i] & resetbit));
        const double x = (double)v[i];
        if (rs) { hr = 1; val = x; }
        else val = x < val ? x : val;
    }
    double run = block_excl_scan(SR{val, hr}, OpReset(), SR{INFINITY, 0}, s_warp).v;
    for (int i = e - 1; i >= b; i--) {
        const bool rs = (i == n - 1) || (reset && (reset[i] & resetbit));
        const double x = (double)v[i];
        if (!rs && x >= run) flag[i] |= bit;
        else run = x;
    }
    __syncthreads();
}

#ifndef B2_MASK_MODE_DEFAULT
#define B2_MASK_MODE_DEFAULT 0
#endif
#ifndef B2_MASK_KNOTS
#define B2_MASK_KNOTS 256
#endif
constexpr int kMaskKnots = B2_MASK_KNOTS;
// sweep windows (mode 2): capacity of one staged window of the sorted line, in samples, per array
#ifndef B2_MASK_WCAP
#define B2_MASK_WCAP 1024
#endif
constexpr int kMaskWinCap = B2_MASK_WCAP;
#ifndef B2_MASK_FIRST_DEPTH
#define B2_MASK_FIRST_DEPTH 4
#endif
constexpr int kMaskFirstPassDepth = B2_MASK_FIRST_DEPTH;

// ---- 1-D bulk copies global -> shared through the TMA unit, completion on an mbarrier (sm_90+) ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(
                     smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

// One staged window: samples [start, start + n) of one of the three sorted arrays of the line.  The bulk copy needs a
// 16-byte aligned source and a multiple of 16 bytes, so the window starts on an even element of its array and a last odd
// element travels as a plain load.
struct MaskWindow {
    int i0, i1;       // samples of the sorted line the sweep can touch: [i0, i1]
    int start[3];     // first staged sample of cs / lats / lons (<= i0, 16-byte aligned address)
    int use;          // 0: the window does not fit (or does not exist): this sweep searches global memory
};

// Per line: (A) min / max / fold-over test; (B) only lines with fold-over are co-sorted; (C, D) the 2 w + 1 samples of the
// regular cross-track grid are resampled and their slant ranges tested for order AS THEY ARE PRODUCED -- they are not
// stored: on a line whose ranges ascend (no layover anywhere on it) neither layover scan can flag anything, so nothing
// else is needed; (E) shadow scans; (F) only lines with range fold-over produce the samples a second time, now into the
// scratch arrays the sort / scan / scatter stages work on.  DRAM traffic of a line without layover is the algorithmic
// 28 B in + 1 B out per pixel.
//
// How a sample finds its bracket in the sorted cross-track positions (`mode`, same result in every mode):
//   0  warp-sequential walk: every warp owns a contiguous run of the grid; the bracket of its previous 32 samples is the
//      guess for the next 32, probes and lat / lon reads hit the cache lines it has just used; no barrier in the loop;
//      (a third variant, the whole sorted line in shared memory, left ~28 KB of L1 for the DEM taps and ran the C2 line
//      pass in 34.3 ms instead of 23.8: removed);
//   2  sweep windows: the 1024 samples of a sweep only touch a contiguous window of ~600 entries of the sorted line; its
//      bounds follow from the knots, thread 0 has the TMA unit copy the NEXT sweep's window of all three arrays (cs, lat,
//      lon) into the other half of a double buffer (cp.async.bulk + mbarrier) while the CTA works on the current one, and
//      both the search and the lat / lon interpolation read shared memory only.
template <int METHOD, bool REF>
__global__ void __launch_bounds__(kMaskBlock, 1024 / kMaskBlock)
k_topo_mask(const __grid_constant__ TopoConst C, const LineState *__restrict__ states, int line0, int nlines, TopoLayers out,
            float demmax, MaskScratch scr, int mode, int stage_elev)
{
    // dynamic: [mode 2: 2 x 3 x kMaskWinCap doubles] then [width] mask bytes, then (stage_elev) [width] floats
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ SR s_warp[32];                // scan scratch (largest scan state)
    __shared__ double s_mm[2];
    __shared__ double s_knot[kMaskKnots + 1];
    __shared__ double s_first[kMaskBlock / 32 + 1]; // first sample of every warp of the current sweep
    __shared__ double s_carry;                      // last sample of the previous sweep
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ MaskWindow s_win[2];
    __shared__ int s_flag;
    __shared__ LineState sL;
    const int w = C.width, ow = 2 * w + 1; // :134-135
    double *cs_s = scr.cs + (size_t)blockIdx.x * w, *lats_s = scr.lats + (size_t)blockIdx.x * w,
           *lons_s = scr.lons + (size_t)blockIdx.x * w;
    double *orng = scr.orng + (size_t)blockIdx.x * ow;
    double *ctr_sorted = scr.ctr_sorted + (size_t)blockIdx.x * ow, *orng_sorted = scr.orng_sorted + (size_t)blockIdx.x * ow;
    double *pm = scr.pm + (size_t)blockIdx.x * ow, *sm = scr.sm + (size_t)blockIdx.x * ow;
    int *rank = scr.rank + (size_t)blockIdx.x * ow;
    unsigned char *oflag = scr.oflag + (size_t)blockIdx.x * ow;
    double *s_wbuf = reinterpret_cast<double *>(s_dyn); // mode 2: [2][3][kMaskWinCap]
    const size_t head = mode == 2 ? (size_t)6 * kMaskWinCap * sizeof(double) : 0;
    unsigned char *sbytes = s_dyn + head;
    unsigned int *smask = reinterpret_cast<unsigned int *>(sbytes);
    SD *s_warp_d = reinterpret_cast<SD *>(s_warp);
    SF *s_warp_f = reinterpret_cast<SF *>(s_warp);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int nsweeps = (ow + (int)blockDim.x - 1) / (int)blockDim.x;
    unsigned bar_phase[2] = {0u, 0u};
    if (mode == 2 && threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    for (int row = blockIdx.x; row < nlines; row += gridDim.x) {
        const int line = line0 + row;
        load_line_state(sL, states, row);
        const double *ctrack_in = out.ctrack + (size_t)row * w;
        const double *lat_in = out.lat + (size_t)row * w, *lon_in = out.lon + (size_t)row * w;
        const float *elev = out.elev + (size_t)row * w;
        // ---- (A) ctrack extent :730-732 and "is the line free of fold-over" in one pass ----
        double mn = INFINITY, mx = -INFINITY;
        int unsorted = 0;
        // kMaskFirstPassDepth strided samples (and their left neighbours) per trip, all loads issued before the first use: this pass is
        // pure load latency
        for (int i0 = threadIdx.x; i0 < w; i0 += kMaskFirstPassDepth * blockDim.x) {
            double v[kMaskFirstPassDepth], u[kMaskFirstPassDepth];
#pragma unroll
            for (int q = 0; q < kMaskFirstPassDepth; q++) {
                const int i = i0 + q * (int)blockDim.x;
                v[q] = i < w ? ctrack_in[i] : NAN;
                u[q] = (i < w && i > 0) ? ctrack_in[i - 1] : NAN;
            }
#pragma unroll
            for (int q = 0; q < kMaskFirstPassDepth; q++) {
                mn = fmin(mn, v[q]); // fmin / fmax skip the NaN padding
                mx = fmax(mx, v[q]);
                if (u[q] > v[q]) unsorted = 1; // NaNs count as ordered, like the reference's insertion sort
            }
        }
        mn = warp_min(mn);
        mx = warp_max(mx);
        __syncthreads();
        if (lane == 0) s_warp_d[wid].v = mn;
        __syncthreads();
        // two-level reduce: 32 warp minima / maxima
        if (threadIdx.x < 32) {
            double a = threadIdx.x < nwarps ? s_warp_d[threadIdx.x].v : INFINITY;
            a = warp_min(a);
            if (threadIdx.x == 0) s_mm[0] = a;
        }
        __syncthreads();
        if (lane == 0) s_warp_d[wid].v = mx;
        __syncthreads();
        if (threadIdx.x < 32) {
            double a = threadIdx.x < nwarps ? s_warp_d[threadIdx.x].v : -INFINITY;
            a = warp_max(a);
            if (threadIdx.x == 0) s_mm[1] = a;
        }
        const bool ctrack_sorted = __syncthreads_or(unsorted) == 0;
        const double ctrackmin = s_mm[0] - demmax, ctrackmax = s_mm[1] + demmax;
        const double dctrack = (ctrackmax - ctrackmin) / (ow - 1.0);

        // ---- (B) stable co-sort (ctrack; lat, lon) :735: nothing to do on a line without fold-over ----
        const double *cs = ctrack_in, *lats = lat_in, *lons = lon_in;
        if (!ctrack_sorted) {
            block_prefix_max_suffix_min(ctrack_in, w, pm, sm, s_warp_d);
            block_stable_ranks(ctrack_in, w, pm, sm, rank, &s_flag);
            for (int i = threadIdx.x; i < w; i += blockDim.x) {
                const int r = rank[i];
                cs_s[r] = ctrack_in[i];
                lats_s[r] = lat_in[i];
                lons_s[r] = lon_in[i];
            }
            cs = cs_s;
            lats = lats_s;
            lons = lons_s;
            __threadfence_block();
            __syncthreads();
        }

        // ---- (C) DEM surface on the regular cross-track grid :745-782 ----
        // The sorted cross-track positions are smooth but not uniform in the sample index (ground spacing changes across
        // the swath), so a straight line through the end points misses the bracket by hundreds of samples.  kMaskKnots + 1
        // samples of the array give a piecewise linear inverse that lands within a sample or two (modes 0 / 1) and the
        // bounds of a sweep's window (mode 2); the search itself (and therefore the result) is unchanged.
        for (int k = threadIdx.x; k <= kMaskKnots; k += blockDim.x) s_knot[k] = cs[(int)(((long long)k * (w - 1)) / kMaskKnots)];
        if (threadIdx.x == 0) s_carry = -INFINITY;
        __syncthreads();
        const double cs0 = s_knot[0], csn = s_knot[kMaskKnots];
        const double kscale = (csn > cs0) ? (double)kMaskKnots / (csn - cs0) : 0.0;
        auto knot_index = [&](int k) -> int { return (int)(((long long)k * (w - 1)) / kMaskKnots); };
        auto knot_below = [&](double aa) -> int { // largest k with s_knot[k] <= aa, -1 if none
            if (!(aa >= cs0)) return -1;
            if (aa >= csn) return kMaskKnots;
            int k = (int)((aa - cs0) * kscale);
            k = k < 0 ? 0 : (k > kMaskKnots - 1 ? kMaskKnots - 1 : k);
            while (k > 0 && s_knot[k] > aa) k--;
            while (k < kMaskKnots && s_knot[k + 1] <= aa) k++;
            return k;
        };
        auto knot_guess = [&](double aa) -> int {
            if (!(aa > cs0)) return 0;
            if (!(aa < csn)) return w - 1;
            int k = (int)((aa - cs0) * kscale);
            k = k < 0 ? 0 : (k > kMaskKnots - 1 ? kMaskKnots - 1 : k);
            while (k > 0 && s_knot[k] > aa) k--;
            while (k < kMaskKnots - 1 && s_knot[k + 1] <= aa) k++;
            const int i0 = knot_index(k), i1 = knot_index(k + 1);
            const double a = s_knot[k], b = s_knot[k + 1];
            const double f = (b > a) ? (aa - a) / (b - a) : 0.0;
            return i0 + (int)(f * (double)(i1 - i0));
        };
        auto grid_pos = [&](int p) -> double { return ctrackmin + ((p + 1) - 1) * dctrack; }; // :747
        auto sample_global = [&](int p) -> double { // slant range of grid sample p (0-based), :747-782
            const double aa = grid_pos(p);
            const int it = ref_search_result(search_count_le([&](int m) { return cs[m]; }, w, aa, knot_guess(aa)), w);
            return mask_resample<METHOD, REF>(C, sL, cs, lats, lons, it, aa);
        };
        // mode 2: window of sweep `sw` -- thread 0 only.  Everything at or before sample i0 is <= the sweep's first grid
        // position, everything at or after i1 is > its last one (or i0 / i1 is an end of the line), so the bracket of every
        // sample of the sweep lies inside [i0, i1].
        auto stage_window = [&](int sw) {
            MaskWindow W;
            const int base = sw * (int)blockDim.x;
            const int plast = base + (int)blockDim.x - 1 < ow - 1 ? base + (int)blockDim.x - 1 : ow - 1;
            const int ka = knot_below(grid_pos(base)), kb = knot_below(grid_pos(plast)) + 1;
            W.i0 = ka < 0 ? 0 : knot_index(ka < kMaskKnots ? ka : kMaskKnots);
            W.i1 = kb > kMaskKnots ? w - 1 : knot_index(kb);
            if (W.i0 > 0) W.i0 -= 1; // the bracket's lower sample may be the one just below (search result clamped to w - 1)
            if (W.i1 < 1) W.i1 = 1;  // ... and its upper sample is never below the second one (search result clamped to 1)
            const double *src[3] = {cs, lats, lons};
            W.use = 1;
            unsigned bytes = 0;
            int nfull[3];
            for (int a = 0; a < 3; a++) {
                const int off = (int)((reinterpret_cast<uintptr_t>(src[a] + W.i0) >> 3) & 1);
                W.start[a] = W.i0 - off;
                const int n = W.i1 - W.start[a] + 1;
                if (W.start[a] < 0 || n > kMaskWinCap) W.use = 0;
                nfull[a] = n & ~1;
                bytes += (unsigned)nfull[a] * 8u;
            }
            s_win[sw & 1] = W;
            if (!W.use) return;
            double *dst = s_wbuf + (size_t)(sw & 1) * 3 * kMaskWinCap;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // earlier generic accesses of the buffer are done
            mbar_arrive_expect_tx(&s_bar[sw & 1], bytes);
            for (int a = 0; a < 3; a++) {
                const int n = W.i1 - W.start[a] + 1;
                if (nfull[a]) bulk_g2s(dst + (size_t)a * kMaskWinCap, src[a] + W.start[a], (unsigned)nfull[a] * 8u, &s_bar[sw & 1]);
                if (n & 1) dst[(size_t)a * kMaskWinCap + n - 1] = src[a][W.start[a] + n - 1]; // visible after the next barrier
            }
        };
        if (mode == 2) {
            if (threadIdx.x == 0) stage_window(0);
            __syncthreads();
        }
        // ---- (D) first sweep: is the slant range ascending over the grid?  Every sample is compared with its predecessor.
        int orng_unsorted = 0;
        if (mode == 0) {
            // Warp-sequential walk: every warp owns one contiguous run of the grid and walks it 32 samples at a time, so
            // (a) the bracket of the previous group is a one-probe guess for the next one and every probe hits the cache
            // lines the warp has just used, (b) the predecessor of a sample lives in the neighbouring lane or in the
            // warp's own previous group: no barrier inside the loop (the runs' end points meet once, afterwards).
            const int chunk = (((ow + nwarps - 1) / nwarps) + 31) & ~31;
            const int pbeg = wid * chunk, pend = pbeg + chunk < ow ? pbeg + chunk : ow;
            double carry = -INFINITY, lastv = -INFINITY;
            int cnt_prev = -1;
            if (lane == 0) s_first[wid] = INFINITY;
            for (int g0 = pbeg; g0 < pend; g0 += 32) {
                const int p = g0 + lane;
                double val = INFINITY; // +inf beyond the run: never smaller than its predecessor
                int cnt = 0;
                if (p < pend) {
                    const double aa = grid_pos(p);
                    const int guess = cnt_prev >= 0 ? cnt_prev + ((lane + 1) >> 1) : knot_guess(aa);
                    cnt = search_count_le([&](int m) { return cs[m]; }, w, aa, guess);
#ifndef B2_MASK_NO_PREFETCH
                    {   // the lane's bracket in the NEXT group sits ~16 entries further: have the three lines it will read on
                        // their way into L1 while this group's DEM interpolation runs
                        const int pf = cnt + 16 < w - 1 ? cnt + 16 : w - 1;
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(cs + pf));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(lats + pf));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(lons + pf));
                    }
#endif
                    val = mask_resample<METHOD, REF>(C, sL, cs, lats, lons, ref_search_result(cnt, w), aa);
                }
                double prev = __shfl_up_sync(0xffffffffu, val, 1);
                if (lane == 0) prev = carry;
                if (prev > val) orng_unsorted = 1;
                if (g0 == pbeg && lane == 0) s_first[wid] = val;
                if (p == pend - 1) lastv = val;
                carry = __shfl_sync(0xffffffffu, val, 31);
                cnt_prev = __shfl_sync(0xffffffffu, cnt, 31);
            }
            lastv = warp_max(lastv); // one lane holds the run's last sample, the others -inf
            __syncthreads();
            if (lane == 0 && wid + 1 < nwarps && lastv > s_first[wid + 1]) orng_unsorted = 1;
        }
        for (int sw = 0; mode != 0 && sw < nsweeps; sw++) {
            const int base = sw * (int)blockDim.x;
            const int p = base + (int)threadIdx.x;
            double val = INFINITY; // +inf beyond the grid: never smaller than its predecessor
            if (mode == 2) {
                // the next sweep's window travels while this one is worked on (its buffer was last read two sweeps ago)
                if (threadIdx.x == 0 && sw + 1 < nsweeps) stage_window(sw + 1);
                const MaskWindow W = s_win[sw & 1];
                if (W.use) {
                    mbar_wait(&s_bar[sw & 1], bar_phase[sw & 1]);
                    bar_phase[sw & 1] ^= 1u;
                    if (p < ow) {
                        const double *wb = s_wbuf + (size_t)(sw & 1) * 3 * kMaskWinCap;
                        const double *wcs = wb - W.start[0], *wla = wb + kMaskWinCap - W.start[1], *wlo = wb + 2 * kMaskWinCap - W.start[2];
                        const double aa = grid_pos(p);
                        const int L = W.i1 - W.i0 + 1;
                        const double a = wcs[W.i0], b = wcs[W.i1];
                        const int guess = (b > a) ? (int)((aa - a) / (b - a) * (double)(L - 1)) : 0;
                        const int cnt = W.i0 + search_count_le([&](int m) { return wcs[W.i0 + m]; }, L, aa, guess);
                        val = mask_resample<METHOD, REF>(C, sL, wcs, wla, wlo, ref_search_result(cnt, w), aa);
                    }
                } else if (p < ow) {
                    val = sample_global(p);
                }
            }
            const double prev = __shfl_up_sync(0xffffffffu, val, 1);
            if (lane != 0 && prev > val) orng_unsorted = 1;
            if (lane == 0) s_first[wid] = val;
            __syncthreads();
            if (lane == 31 && wid + 1 < nwarps && val > s_first[wid + 1]) orng_unsorted = 1;
            if (threadIdx.x == 0 && s_carry > val) orng_unsorted = 1;
            __syncthreads();
            if (threadIdx.x == blockDim.x - 1) s_carry = val;
        }
        const bool orng_sorted_already = __syncthreads_or(orng_unsorted) == 0;

        // ---- (E) shadow (:791-809) on float32 elevang in pixel order ----
        // the scans give every thread a contiguous chunk of the line (a stride of ~25 samples between neighbouring lanes):
        // the line is brought into shared memory with coalesced loads first when it fits behind the mask bytes
        const float *elev_s = elev;
        if (stage_elev) {
            float *se = reinterpret_cast<float *>(sbytes + (((size_t)w + 15) & ~(size_t)15));
#pragma unroll 4
            for (int i = threadIdx.x; i < w; i += blockDim.x) se[i] = elev[i];
            elev_s = se;
        }
        for (int i = threadIdx.x; i < (w + 3) / 4; i += blockDim.x) smask[i] = 0u;
        __syncthreads();
        block_prefix_max_flags<float>(elev_s, w, w, sbytes, (unsigned char)1, s_warp_f, OpMaxF());
        block_suffix_min_flags<float>(elev_s, w, nullptr, 0, sbytes, (unsigned char)1, s_warp);

        // ---- (F) stable co-sort (orng; ctrack) :787 and layover (:834-852) on the range-sorted ctrack ----
        // ctrack increases with the sample index by construction, so when the slant ranges are already ascending the
        // sorted ctrack is ascending too and neither layover scan can flag anything: the line is done.
        if (!orng_sorted_already) {
            for (int p = threadIdx.x; p < ow; p += blockDim.x) orng[p] = sample_global(p); // the same values, now kept
            __syncthreads();
            block_prefix_max_suffix_min(orng, ow, pm, sm, s_warp_d);
            block_stable_ranks(orng, ow, pm, sm, rank, &s_flag);
            for (int i = threadIdx.x; i < ow; i += blockDim.x) {
                const int r = rank[i];
                orng_sorted[r] = orng[i];
                ctr_sorted[r] = ctrackmin + ((i + 1) - 1) * dctrack; // the cross-track position of sample i (:747)
                oflag[i] = 0;
            }
            __syncthreads();
            // forward layover scan is bounded by `width`, not `owidth`, exactly as in the reference (:835); the
            // backward scan treats forward-flagged samples as resets (:847)
            block_prefix_max_flags<double>(ctr_sorted, ow, w, oflag, (unsigned char)2, s_warp_d, OpMaxD());
            block_suffix_min_flags<double>(ctr_sorted, ow, oflag, (unsigned char)2, oflag, (unsigned char)4, s_warp);

            // ---- scatter to radar pixels through the slant-range line (:855-865) ----
            const double rho0 = pixel_range(C, line, 0), rhon = pixel_range(C, line, w - 1);
            const double rscale = (rhon > rho0) ? (double)(w - 1) / (rhon - rho0) : 0.0;
            for (int i = threadIdx.x; i < ow; i += blockDim.x) {
                if (oflag[i]) {
                    const double val = orng_sorted[i];
                    const int guess = (int)((val - rho0) * rscale);
                    const int j = ref_search_result(search_count_le([&](int m) { return pixel_range(C, line, m); }, w, val, guess), w);
                    // mask(j) < omask(i) => mask(j) += 2  <=>  set bit 1 (mask is 0/1 before any layover hit)
                    const int bidx = j - 1;
                    atomicOr(&smask[bidx >> 2], 2u << (8 * (bidx & 3)));
                }
            }
        }
        __syncthreads();
        signed char *mrow = out.mask + (size_t)row * w;
        for (int i = threadIdx.x; i < w; i += blockDim.x) mrow[i] = (signed char)sbytes[i];
        __syncthreads();
    }
}

// -------------------------------------------------------------------------------------------------
// host-side launchers
// -------------------------------------------------------------------------------------------------
void launch_topo_bbox(const TopoConst &C, const OrbitView &orb, double *d_out, cudaStream_t s)
{
    k_topo_bbox<<<1, 32, 0, s>>>(C, orb, d_out);
}

void launch_dem_prepare(const void *raw, int dtype, float *dem, size_t n, int *maxkey, cudaStream_t s)
{
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    k_dem_prepare<<<blocks, 256, 0, s>>>(raw, dtype, dem, n, maxkey);
}

void launch_dem_pad64(const float *dem, int nx, int ny, double *out, int stride, cudaStream_t s)
{
    const size_t n = (size_t)(ny + 1) * (size_t)stride;
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    k_dem_pad64<<<blocks, 256, 0, s>>>(dem, nx, ny, out, stride);
}

void launch_line_setup(const TopoConst &C, const OrbitView &orb, int line0, int nlines, LineState *states, cudaStream_t s)
{
    k_line_setup<<<(nlines + 63) / 64, 64, 0, s>>>(C, orb, line0, nlines, states);
}

template <int METHOD>
static void launch_pixels_m(const TopoConst &C, const LineState *states, int line0, int nlines, const TopoLayers &out,
                            TopoStats *stats, unsigned grid, unsigned grid_solve, cudaStream_t s, cudaEvent_t ev_mid)
{
    constexpr bool kSplit = (METHOD == 5 || METHOD == 2 || METHOD == 0 || METHOD == 4);
    if (kSplit) {
        // one warp per segment of a line (see k_topo_solve)
        const long long nwarps = (long long)solve_segs_per_line(C.width) * nlines;
        grid_solve = (unsigned)((nwarps + kTopoBlock / 32 - 1) / (kTopoBlock / 32));
        if (C.ref.use_ref) {
            k_topo_solve<METHOD, true><<<grid_solve, kTopoBlock, 0, s>>>(C, states, line0, nlines, out.ctrack, stats);
            if (ev_mid) cudaEventRecord(ev_mid, s);
            k_topo_final<METHOD, true><<<grid, kTopoBlock, 0, s>>>(C, states, line0, out, stats);
        } else {
            k_topo_solve<METHOD, false><<<grid_solve, kTopoBlock, 0, s>>>(C, states, line0, nlines, out.ctrack, stats);
            if (ev_mid) cudaEventRecord(ev_mid, s);
            k_topo_final<METHOD, false><<<grid, kTopoBlock, 0, s>>>(C, states, line0, out, stats);
        }
    } else {
        if (C.ref.use_ref) k_topo_fused<METHOD, true><<<grid_solve, kTopoBlock, 0, s>>>(C, states, line0, out, stats);
        else k_topo_fused<METHOD, false><<<grid_solve, kTopoBlock, 0, s>>>(C, states, line0, out, stats);
        if (ev_mid) cudaEventRecord(ev_mid, s);
    }
}

int topo_pixel_launches(int method) { return (method == 5 || method == 2 || method == 0 || method == 4) ? 2 : 1; }

// launches k_topo_solve + k_topo_final (2 kernels); out.ctrack must be allocated (it carries the SCH height between them)
int launch_topo_pixels(const TopoConst &C, const LineState *states, int line0, int nlines, const TopoLayers &out,
                       TopoStats *stats, cudaStream_t s, cudaEvent_t ev_mid)
{
    const long long nblk = (long long)((C.width + kTopoBlock - 1) / kTopoBlock) * nlines;
    if (nblk > 0x7fffffffLL || !out.ctrack) return -2;
    const unsigned gs = (unsigned)(((C.width + kSolvePixelsPerCta - 1) / kSolvePixelsPerCta) * (long long)nlines);
    switch (C.method) {
    case 0: launch_pixels_m<0>(C, states, line0, nlines, out, stats, (unsigned)nblk, gs, s, ev_mid); break;
    case 1: launch_pixels_m<1>(C, states, line0, nlines, out, stats, (unsigned)nblk, gs, s, ev_mid); break;
    case 2: launch_pixels_m<2>(C, states, line0, nlines, out, stats, (unsigned)nblk, gs, s, ev_mid); break;
    case 3: launch_pixels_m<3>(C, states, line0, nlines, out, stats, (unsigned)nblk, gs, s, ev_mid); break;
    case 4: launch_pixels_m<4>(C, states, line0, nlines, out, stats, (unsigned)nblk, gs, s, ev_mid); break;
    case 5: launch_pixels_m<5>(C, states, line0, nlines, out, stats, (unsigned)nblk, gs, s, ev_mid); break;
    default: return -1;
    }
    return 0;
}

int mask_grid_size(int nlines)
{
    int g = 148 * (2048 / kMaskBlock); // persistent CTAs: as many as can be resident at 64 registers/thread, x2
    return nlines < g ? nlines : g;
}

template <int METHOD>
static void launch_mask_m(const TopoConst &C, const LineState *states, int line0, int nlines, const TopoLayers &out,
                          float demmax, const MaskScratch &scr, int grid, size_t smem, int mode, int stage_elev, cudaStream_t s)
{
    if (C.ref.use_ref) {
        cudaFuncSetAttribute(k_topo_mask<METHOD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_topo_mask<METHOD, true><<<grid, kMaskBlock, smem, s>>>(C, states, line0, nlines, out, demmax, scr, mode, stage_elev);
    } else {
        cudaFuncSetAttribute(k_topo_mask<METHOD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_topo_mask<METHOD, false><<<grid, kMaskBlock, smem, s>>>(C, states, line0, nlines, out, demmax, scr, mode, stage_elev);
    }
}

int launch_topo_mask(const TopoConst &C, const LineState *states, int line0, int nlines, const TopoLayers &out, float demmax,
                     const MaskScratch &scr, int grid, cudaStream_t s)
{
    // how a sample finds its bracket in the sorted line (see k_topo_mask): warp-sequential walk by default; B200_MASK_MODE=2
    // selects the TMA-staged sweep windows for A/B measurements (same results in both modes; measured on C2: 23.9 vs 29.9 ms)
    static const int forced = [] {
        const char *e = getenv("B200_MASK_MODE");
        return e ? atoi(e) : -1;
    }();
    const size_t bytes_mask = (size_t)((C.width + 3) / 4) * 4;
    const size_t budget = 227u * 1024u - 12288u; // static shared memory of the kernel: knots, scan scratch, line state
    int stage = (forced == 0 || forced == 2) ? forced : B2_MASK_MODE_DEFAULT;
    if (stage == 2 && (size_t)6 * kMaskWinCap * sizeof(double) + bytes_mask > budget) stage = 0;
    size_t smem = bytes_mask + (stage == 2 ? (size_t)6 * kMaskWinCap * sizeof(double) : 0);
    // the elevation line of the shadow scans (float32) can sit behind the mask bytes (B200_MASK_ELEV_KB = largest dynamic
    // shared memory for which it does); off by default: on C2 the 100 KB it takes from the L1 cost the resampling pass
    // more (3.40 vs 3.10 ms per 1500 lines) than the coalesced scan input gains
    static const size_t elev_limit = [] {
        const char *e = getenv("B200_MASK_ELEV_KB");
        return (size_t)(e ? atoi(e) : 0) * 1024u;
    }();
    const size_t bytes_elev = (((size_t)C.width + 15) & ~(size_t)15) - bytes_mask + (size_t)C.width * sizeof(float) + 16;
    const int stage_elev = (smem + bytes_elev <= elev_limit) ? 1 : 0;
    if (stage_elev) smem += bytes_elev;
    switch (C.method) {
    case 0: launch_mask_m<0>(C, states, line0, nlines, out, demmax, scr, grid, smem, stage, stage_elev, s); break;
    case 1: launch_mask_m<1>(C, states, line0, nlines, out, demmax, scr, grid, smem, stage, stage_elev, s); break;
    case 2: launch_mask_m<2>(C, states, line0, nlines, out, demmax, scr, grid, smem, stage, stage_elev, s); break;
    case 3: launch_mask_m<3>(C, states, line0, nlines, out, demmax, scr, grid, smem, stage, stage_elev, s); break;
    case 4: launch_mask_m<4>(C, states, line0, nlines, out, demmax, scr, grid, smem, stage, stage_elev, s); break;
    case 5: launch_mask_m<5>(C, states, line0, nlines, out, demmax, scr, grid, smem, stage, stage_elev, s); break;
    default: return -1;
    }
    return 0;
}

} // namespace b2
