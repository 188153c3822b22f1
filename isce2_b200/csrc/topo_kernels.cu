// topo_kernels.cu -- sm_100a kernels of the topozero (rdr2geo) path.
//
//   k_topo_bbox        8 threads: scene-corner geolocation at h = -500 / 9000 m  (topozero.f90:194-257)
//   k_dem_prepare      DEM crop -> float32 (+ max height)                         (topozero.f90:333-345)
//   k_line_setup       one thread per azimuth line: orbit state, TCN basis, peg   (topozero.f90:371-424)
//   k_topo_pixels      one thread per radar pixel: iterative height solve + final
//                      geolocation/LOS/incidence pass, coalesced layer stores     (topozero.f90:458-726)
//   k_topo_mask        one CTA per azimuth line: layover / shadow mask            (topozero.f90:729-880)
//
// Compiled with -fmad=false (see geom_device.cuh).
#include "topo_kernels.cuh"

#include <cfloat>

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
template <int METHOD, bool REF>
__global__ void __launch_bounds__(kTopoBlock, B2_TOPO_MINBLOCKS)
k_topo_pixels(const __grid_constant__ TopoConst C, const LineState *__restrict__ states, int line0, TopoLayers out,
              TopoStats *stats)
{
    __shared__ LineState sL;
    __shared__ long long s_mm[4];          // order keys: min lat, max lat, min lon, max lon
    __shared__ unsigned int s_cnt[3];      // converged, iterations, warps done
    const int bpl = (C.width + kTopoBlock - 1) / kTopoBlock; // CTAs per azimuth line
    const int row = blockIdx.x / bpl;                        // row within the block of lines
    const int seg = blockIdx.x - row * bpl;
    {
        const double *src = reinterpret_cast<const double *>(states + row);
        double *dst = reinterpret_cast<double *>(&sL);
        for (int i = threadIdx.x; i < (int)(sizeof(LineState) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
        if (threadIdx.x == 0) {
            s_mm[0] = s_mm[2] = 0x7fffffffffffffffLL;
            s_mm[1] = s_mm[3] = (long long)0x8000000000000000ULL;
            s_cnt[0] = s_cnt[1] = s_cnt[2] = 0u;
        }
    }
    __syncthreads();
    const int pix = seg * blockDim.x + threadIdx.x;
    double mnlat = 1e300, mxlat = -1e300, mnlon = 1e300, mxlon = -1e300;
    int conv = 0, iters = 0;
    if (pix < C.width) {
        const int line = line0 + row;
        double rng = pixel_range(C, line, pix);
        double dop = eval_poly2d(C.dop, (double)line, (double)pix);
        PixelResult R;
        topo_pixel<METHOD, REF>(C, sL, rng, dop, out.inc != nullptr, R);
        const size_t w = (size_t)C.width;
        const size_t o = (size_t)row * w + (size_t)pix;
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
        if (out.ctrack) {
            out.ctrack[o] = R.ctrack;
            out.elev[o] = R.elev;
        }
        mnlat = mxlat = R.lat;
        mnlon = mxlon = R.lon;
        conv = R.converged;
        iters = R.iters;
    }
    // Scene statistics (topozero.f90:712-715, :570) without a block-wide barrier: every warp folds its values into
    // shared memory as it finishes and leaves; the last warp to arrive publishes the CTA's totals.  Early-converged
    // warps therefore never wait for the slow ones.
    mnlat = warp_min(mnlat); mxlat = warp_max(mxlat); mnlon = warp_min(mnlon); mxlon = warp_max(mxlon);
    conv = warp_sum(conv); iters = warp_sum(iters);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&s_mm[0], order_key(mnlat));
        atomicMax(&s_mm[1], order_key(mxlat));
        atomicMin(&s_mm[2], order_key(mnlon));
        atomicMax(&s_mm[3], order_key(mxlon));
        atomicAdd(&s_cnt[0], (unsigned int)conv);
        atomicAdd(&s_cnt[1], (unsigned int)iters);
        __threadfence_block();
        const unsigned int done = atomicAdd(&s_cnt[2], 1u);
        if (done == (blockDim.x >> 5) - 1) {
            __threadfence_block();
            atomicMin(&stats->min_lat, atomicMin(&s_mm[0], 0x7fffffffffffffffLL));
            atomicMax(&stats->max_lat, atomicMax(&s_mm[1], (long long)0x8000000000000000ULL));
            atomicMin(&stats->min_lon, atomicMin(&s_mm[2], 0x7fffffffffffffffLL));
            atomicMax(&stats->max_lon, atomicMax(&s_mm[3], (long long)0x8000000000000000ULL));
            atomicAdd(&stats->converged, (unsigned long long)atomicAdd(&s_cnt[0], 0u));
            atomicAdd(&stats->iterations, (unsigned long long)atomicAdd(&s_cnt[1], 0u));
        }
    }
}

// -------------------------------------------------------------------------------------------------
// layover / shadow mask: one CTA per azimuth line (persistent over lines)
// -------------------------------------------------------------------------------------------------
// reference binarysearch (topozero.f90:933-963): returns the 1-based `left` in [1, n-1]
template <typename F>
__device__ __forceinline__ int ref_binarysearch(F at /*1-based accessor*/, int n, double val)
{
    int left = 1, right = n;
    while (true) {
        if (left > right) break;
        int middle = (left + right + 1) >> 1; // nint((left+right)/2.0)
        if (left == right - 1) return left;
        double a = at(middle);
        if (a <= val) left = middle;
        else if (a > val) right = middle;
        else return left; // NaN
        if (left == right) return left;
    }
    return left;
}

// (key, idx) lexicographic order == the order a stable sort by key produces
__device__ __forceinline__ bool pair_greater(double ka, int ia, double kb, int ib)
{
    return (ka > kb) || (ka == kb && ia > ib);
}

// CTA-wide bitonic sort of n (key, idx) pairs held in global scratch padded to P (power of two) entries
__device__ void block_bitonic_sort(double *key, int *idx, int P)
{
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                // element pair (i, i^j) with i having bit j clear
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int l = i | j;
                bool up = ((i & k) == 0);
                double ka = key[i], kb = key[l];
                int ia = idx[i], ib = idx[l];
                bool gt = pair_greater(ka, ia, kb, ib);
                if (gt == up) {
                    key[i] = kb; key[l] = ka;
                    idx[i] = ib; idx[l] = ia;
                }
            }
            __syncthreads();
        }
    }
}

// block-wide "is non-decreasing" test of a[0..n)
__device__ bool block_is_sorted(const double *a, int n, int *s_flag)
{
    if (threadIdx.x == 0) *s_flag = 1;
    __syncthreads();
    int bad = 0;
    for (int i = threadIdx.x + 1; i < n; i += blockDim.x)
        if (a[i - 1] > a[i]) bad = 1;
    if (bad) *s_flag = 0; // benign race: everybody writes the same value
    __syncthreads();
    int r = *s_flag;
    __syncthreads();
    return r != 0;
}

// Forward scan of the reference (:791-799, :834-842):  aa = v(1); for i = 2..nflag: if v(i) <= aa flag else aa = v(i).
// Flagged samples never exceed aa, so aa is the plain prefix maximum: flag[i] |= bit if v[i] <= max(v[0..i-1]).
template <typename T>
__device__ void block_prefix_max_flags(const T *v, int n, int nflag, unsigned char *flag, unsigned char bit, T *s_part)
{
    const int nt = blockDim.x;
    const int chunk = (n + nt - 1) / nt;
    const int b = min(n, (int)threadIdx.x * chunk);
    const int e = min(n, b + chunk);
    T m = -INFINITY;
    for (int i = b; i < e; i++) m = v[i] > m ? v[i] : m;
    s_part[threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.x == 0) { // 1024 sequential steps: negligible next to the sort
        T run = -INFINITY;
        for (int t = 0; t < nt; t++) { T x = s_part[t]; s_part[t] = run; run = x > run ? x : run; }
    }
    __syncthreads();
    T run = s_part[threadIdx.x];
    for (int i = b; i < e; i++) {
        T x = v[i];
        if (i >= 1 && i < nflag && x <= run) flag[i] |= bit;
        run = x > run ? x : run;
    }
    __syncthreads();
}

// Backward scan of the reference (:801-809, :844-852):
//   aa = v(n); for i = n-1..1: if (v(i) >= aa .and. .not. reset(i)) flag else aa = v(i)
// where reset(i) is "already flagged by the forward layover scan" (omask(i) >= 2, :847) and absent for the
// shadow scan.  Per sample the state update is either aa <- min(aa, v) or aa <- v (reset); such updates compose
// associatively as (has_reset, value), which gives the chunked parallel form below.
template <typename T>
__device__ void block_suffix_min_flags(const T *v, int n, const unsigned char *reset, unsigned char resetbit,
                                       unsigned char *flag, unsigned char bit, T *s_part, unsigned char *s_hr)
{
    const int nt = blockDim.x;
    const int chunk = (n + nt - 1) / nt;
    const int b = min(n, (int)threadIdx.x * chunk);
    const int e = min(n, b + chunk);
    T val = INFINITY;
    bool hr = false;
    for (int i = e - 1; i >= b; i--) {
        bool rs = (i == n - 1) || (reset && (reset[i] & resetbit));
        T x = v[i];
        if (rs) { hr = true; val = x; }
        else val = x < val ? x : val;
    }
    s_part[threadIdx.x] = val;
    s_hr[threadIdx.x] = hr ? 1 : 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        T run = INFINITY;
        for (int t = nt - 1; t >= 0; t--) {
            T x = s_part[t];
            s_part[t] = run;
            run = s_hr[t] ? x : (x < run ? x : run);
        }
    }
    __syncthreads();
    T run = s_part[threadIdx.x];
    for (int i = e - 1; i >= b; i--) {
        bool rs = (i == n - 1) || (reset && (reset[i] & resetbit));
        T x = v[i];
        if (!rs && x >= run) flag[i] |= bit;
        else run = x;
    }
    __syncthreads();
}

template <int METHOD, bool REF>
__global__ void __launch_bounds__(kMaskBlock)
k_topo_mask(const __grid_constant__ TopoConst C, const LineState *__restrict__ states, int line0, int nlines, TopoLayers out,
            float demmax, MaskScratch scr)
{
    extern __shared__ unsigned char s_dyn[]; // [width] shadow/layover bytes + flag bytes live in scratch
    __shared__ double s_part[kMaskBlock];
    __shared__ unsigned char s_hr[kMaskBlock];
    __shared__ double s_mm[2];
    __shared__ int s_flag;
    __shared__ LineState sL;
    const int w = C.width, ow = 2 * w + 1; // :134-135
    const int P = scr.padded;
    double *key = scr.key + (size_t)blockIdx.x * P;
    int *idx = scr.idx + (size_t)blockIdx.x * P;
    double *cs = scr.cs + (size_t)blockIdx.x * w, *lats = scr.lats + (size_t)blockIdx.x * w, *lons = scr.lons + (size_t)blockIdx.x * w;
    double *rho = scr.rho + (size_t)blockIdx.x * w;
    double *orng = scr.orng + (size_t)blockIdx.x * ow, *ctr = scr.ctr + (size_t)blockIdx.x * ow;
    double *ctr_sorted = scr.ctr_sorted + (size_t)blockIdx.x * ow;
    unsigned char *oflag = scr.oflag + (size_t)blockIdx.x * ow;
    unsigned int *smask = reinterpret_cast<unsigned int *>(s_dyn);
    unsigned char *sbytes = s_dyn;

    for (int row = blockIdx.x; row < nlines; row += gridDim.x) {
        const int line = line0 + row;
        {
            const double *src = reinterpret_cast<const double *>(states + row);
            double *dst = reinterpret_cast<double *>(&sL);
            for (int i = threadIdx.x; i < (int)(sizeof(LineState) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
        }
        const double *ctrack_in = out.ctrack + (size_t)row * w;
        const double *lat_in = out.lat + (size_t)row * w, *lon_in = out.lon + (size_t)row * w;
        const float *elev = out.elev + (size_t)row * w;
        // ---- ctrack extent :730-732 ----
        double mn = INFINITY, mx = -INFINITY;
        for (int i = threadIdx.x; i < w; i += blockDim.x) {
            double v = ctrack_in[i];
            mn = fmin(mn, v);
            mx = fmax(mx, v);
            rho[i] = pixel_range(C, line, i);
        }
        mn = warp_min(mn);
        mx = warp_max(mx);
        if ((threadIdx.x & 31) == 0) { s_part[threadIdx.x >> 5] = mn; s_part[32 + (threadIdx.x >> 5)] = mx; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = INFINITY, b = -INFINITY;
            for (int t = 0; t < (int)(blockDim.x >> 5); t++) { a = fmin(a, s_part[t]); b = fmax(b, s_part[32 + t]); }
            s_mm[0] = a;
            s_mm[1] = b;
        }
        __syncthreads();
        const double ctrackmin = s_mm[0] - demmax, ctrackmax = s_mm[1] + demmax;
        const double dctrack = (ctrackmax - ctrackmin) / (ow - 1.0);

        // ---- stable co-sort (ctrack; lat, lon) :735 ----
        if (block_is_sorted(ctrack_in, w, &s_flag)) {
            for (int i = threadIdx.x; i < w; i += blockDim.x) { cs[i] = ctrack_in[i]; lats[i] = lat_in[i]; lons[i] = lon_in[i]; }
        } else {
            int P1 = 1;
            while (P1 < w) P1 <<= 1;
            for (int i = threadIdx.x; i < P1; i += blockDim.x) { key[i] = i < w ? ctrack_in[i] : INFINITY; idx[i] = i; }
            __syncthreads();
            block_bitonic_sort(key, idx, P1);
            for (int i = threadIdx.x; i < w; i += blockDim.x) { int s = idx[i]; cs[i] = key[i]; lats[i] = lat_in[s]; lons[i] = lon_in[s]; }
        }
        __syncthreads();

        // ---- DEM surface on the regular cross-track grid :745-782 ----
        for (int p = threadIdx.x; p < ow; p += blockDim.x) {
            double aa = ctrackmin + ((p + 1) - 1) * dctrack;
            ctr[p] = aa;
            int it = ref_binarysearch([&](int m) { return cs[m - 1]; }, w, aa);
            if (it == w) it = w - 1;
            if (it == 0) it = 1;
            orng[p] = mask_resample<METHOD, REF>(C, sL, cs, lats, lons, it, aa);
        }
        __syncthreads();

        // ---- stable co-sort (orng; ctrack) :787 ----
        const double *orng_s;
        if (block_is_sorted(orng, ow, &s_flag)) {
            for (int i = threadIdx.x; i < ow; i += blockDim.x) ctr_sorted[i] = ctr[i];
            orng_s = orng;
        } else {
            for (int i = threadIdx.x; i < P; i += blockDim.x) { key[i] = i < ow ? orng[i] : INFINITY; idx[i] = i; }
            __syncthreads();
            block_bitonic_sort(key, idx, P);
            for (int i = threadIdx.x; i < ow; i += blockDim.x) ctr_sorted[i] = ctr[idx[i]];
            orng_s = key;
        }
        __syncthreads();

        // ---- shadow (:791-809) on float32 elevang in pixel order; layover (:834-852) on range-sorted ctrack ----
        for (int i = threadIdx.x; i < (w + 3) / 4; i += blockDim.x) smask[i] = 0u;
        for (int i = threadIdx.x; i < ow; i += blockDim.x) oflag[i] = 0;
        __syncthreads();
        block_prefix_max_flags<float>(elev, w, w, sbytes, (unsigned char)1, reinterpret_cast<float *>(s_part));
        block_suffix_min_flags<float>(elev, w, nullptr, 0, sbytes, (unsigned char)1, reinterpret_cast<float *>(s_part), s_hr);
        // forward layover scan is bounded by `width`, not `owidth`, exactly as in the reference (:835); the
        // backward scan treats forward-flagged samples as resets (:847)
        block_prefix_max_flags<double>(ctr_sorted, ow, w, oflag, (unsigned char)2, s_part);
        block_suffix_min_flags<double>(ctr_sorted, ow, oflag, (unsigned char)2, oflag, (unsigned char)4, s_part, s_hr);

        // ---- scatter to radar pixels through the slant-range line (:855-865) ----
        for (int i = threadIdx.x; i < ow; i += blockDim.x) {
            if (oflag[i]) {
                int j = ref_binarysearch([&](int m) { return rho[m - 1]; }, w, orng_s[i]);
                if (j >= 1 && j <= w) {
                    // mask(j) < omask(i) => mask(j) += 2  <=>  set bit 1 (mask is 0/1 before any layover hit)
                    int b = j - 1;
                    atomicOr(&smask[b >> 2], 2u << (8 * (b & 3)));
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

void launch_line_setup(const TopoConst &C, const OrbitView &orb, int line0, int nlines, LineState *states, cudaStream_t s)
{
    k_line_setup<<<(nlines + 63) / 64, 64, 0, s>>>(C, orb, line0, nlines, states);
}

template <int METHOD>
static void launch_pixels_m(const TopoConst &C, const LineState *states, int line0, const TopoLayers &out, TopoStats *stats,
                            unsigned grid, cudaStream_t s)
{
    if (C.ref.use_ref) k_topo_pixels<METHOD, true><<<grid, kTopoBlock, 0, s>>>(C, states, line0, out, stats);
    else k_topo_pixels<METHOD, false><<<grid, kTopoBlock, 0, s>>>(C, states, line0, out, stats);
}

int launch_topo_pixels(const TopoConst &C, const LineState *states, int line0, int nlines, const TopoLayers &out,
                       TopoStats *stats, cudaStream_t s)
{
    const long long nblk = (long long)((C.width + kTopoBlock - 1) / kTopoBlock) * nlines;
    if (nblk > 0x7fffffffLL) return -2;
    switch (C.method) {
    case 1: launch_pixels_m<1>(C, states, line0, out, stats, (unsigned)nblk, s); break;
    case 2: launch_pixels_m<2>(C, states, line0, out, stats, (unsigned)nblk, s); break;
    case 3: launch_pixels_m<3>(C, states, line0, out, stats, (unsigned)nblk, s); break;
    case 5: launch_pixels_m<5>(C, states, line0, out, stats, (unsigned)nblk, s); break;
    default: return -1;
    }
    return 0;
}

int mask_grid_size(int nlines)
{
    int g = 2 * 148; // persistent CTAs, two per SM
    return nlines < g ? nlines : g;
}

template <int METHOD>
static void launch_mask_m(const TopoConst &C, const LineState *states, int line0, int nlines, const TopoLayers &out,
                          float demmax, const MaskScratch &scr, int grid, size_t smem, cudaStream_t s)
{
    if (C.ref.use_ref) {
        cudaFuncSetAttribute(k_topo_mask<METHOD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_topo_mask<METHOD, true><<<grid, kMaskBlock, smem, s>>>(C, states, line0, nlines, out, demmax, scr);
    } else {
        cudaFuncSetAttribute(k_topo_mask<METHOD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_topo_mask<METHOD, false><<<grid, kMaskBlock, smem, s>>>(C, states, line0, nlines, out, demmax, scr);
    }
}

int launch_topo_mask(const TopoConst &C, const LineState *states, int line0, int nlines, const TopoLayers &out, float demmax,
                     const MaskScratch &scr, int grid, cudaStream_t s)
{
    size_t smem = (size_t)((C.width + 3) / 4) * 4;
    switch (C.method) {
    case 1: launch_mask_m<1>(C, states, line0, nlines, out, demmax, scr, grid, smem, s); break;
    case 2: launch_mask_m<2>(C, states, line0, nlines, out, demmax, scr, grid, smem, s); break;
    case 3: launch_mask_m<3>(C, states, line0, nlines, out, demmax, scr, grid, smem, s); break;
    case 5: launch_mask_m<5>(C, states, line0, nlines, out, demmax, scr, grid, smem, s); break;
    default: return -1;
    }
    return 0;
}

} // namespace b2
