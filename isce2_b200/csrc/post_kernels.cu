// post_kernels.cu -- sm_100a kernels of the two remaining consumers of the geometry layers.
//
//   k_looks_average   box mean of down x across looks, exactly as takeLooks<T> / takeLookscpx<T> accumulate it
//                     (mroipac/looks/bindings/looksmodule.cpp:130-200, :204-275): the `down` lines are added into a
//                     double line buffer in line order, then `across` neighbours of that buffer in sample order, the
//                     sum is divided by double(down*across) and cast to T.  One CTA per (output line, tile of output
//                     samples): phase 1 streams the input lines with coalesced loads into per-column sums in shared
//                     memory, phase 2 adds the neighbours.  Every input byte is read once: HBM-bound.
//   k_looks_nearest   gdal.Translate -outsize (the default method of runMultilook, stripmapStack/topo.py:411-424):
//                     nearest-neighbour decimation, source index floor((i + 0.5) * looks)
//   k_mask_to_radar   SWBDStitcher.toRadar (contrib/demUtils/swbdstitcher/SWBDStitcher.py:107-131): per radar pixel the
//                     mask sample at the truncated, clipped grid index of its latitude / longitude, plus one
//
// Compiled with -fmad=false like the rest of the library; divisions are IEEE (nvcc's default -prec-div=true).
#include "post_kernels.cuh"

#include <cstdint>

namespace b2 {

size_t type_size(int dtype)
{
    switch (dtype) {
    case kTypeByte: return 1;
    case kTypeShort: return 2;
    case kTypeInt: return 4;
    case kTypeLong: return 8;
    case kTypeFloat: return 4;
    case kTypeDouble: return 8;
    case kTypeCFloat: return 8;
    default: return 0;
    }
}

namespace {

constexpr int kLooksBlock = 256;

// static_cast<T>(double) of the x86-64 build: truncation toward zero for the integer types (values are means of T's,
// always in range), round-to-nearest for float
template <typename T> __device__ __forceinline__ T from_double(double v) { return (T)v; }
template <> __device__ __forceinline__ signed char from_double<signed char>(double v) { return (signed char)(int)v; }
template <> __device__ __forceinline__ short from_double<short>(double v) { return (short)(int)v; }

__device__ __forceinline__ size_t row_offset(const LooksGeom &G, int length, int width, int line, int band)
{
    // element offset of the first sample of (line, band); BIP rows hold all bands interleaved (band == 0)
    if (G.scheme == kSchemeBIL) return ((size_t)line * G.bands + band) * (size_t)width;
    if (G.scheme == kSchemeBSQ) return ((size_t)band * length + line) * (size_t)width;
    return (size_t)line * (size_t)width * G.bands;
}

// T: real element type; CPLX: elements are pairs (re, im) of T accumulated separately (takeLookscpx)
template <typename T, bool CPLX>
__global__ void __launch_bounds__(kLooksBlock)
k_looks_average(const __grid_constant__ LooksGeom G, const T *__restrict__ in, T *__restrict__ out, int tile_out, int tiles)
{
    __shared__ double s_col[kLooksMaxTile];
    constexpr int NC = CPLX ? 2 : 1;
    const int inner = (G.scheme == kSchemeBIP) ? G.bands : 1;
    const int row = blockIdx.x / tiles, tile = blockIdx.x - row * tiles;
    // a row is one output line with all bands interleaved (BIP) or one band of one output line (BIL, BSQ)
    const int oline = G.line0 + ((G.scheme == kSchemeBIP) ? row : row / G.bands);
    const int band = (G.scheme == kSchemeBIP) ? 0 : row % G.bands;
    const int o0 = tile * tile_out;                                        // first output sample of the tile
    const int no = (o0 + tile_out <= G.out_width) ? tile_out : G.out_width - o0;
    const int nfull = no * G.la * inner;                                   // input elements of the tile per line
    const size_t e0 = (size_t)o0 * G.la * inner;
    // ---- phase 1: bdbl[j] += ain[j] over the `down` lines, in line order (looksmodule.cpp:160-177) ----
    for (int c = threadIdx.x; c < nfull; c += kLooksBlock) {
        double acc[NC];
#pragma unroll
        for (int q = 0; q < NC; q++) acc[q] = 0.0;
#pragma unroll 4
        for (int i = 0; i < G.ld; i++) {
            const T *p = in + (row_offset(G, G.length, G.width, oline * G.ld + i, band) + e0 + c) * NC;
#pragma unroll
            for (int q = 0; q < NC; q++) acc[q] += (double)p[q];
        }
#pragma unroll
        for (int q = 0; q < NC; q++) s_col[c * NC + q] = acc[q];
    }
    __syncthreads();
    // ---- phase 2: sum of the `across` neighbours in sample order, / norm, cast (:179-193) ----
    const double norm = (double)(G.ld * G.la);
    T *orow = out + (row_offset(G, G.out_length, G.out_width, oline, band) + (size_t)o0 * inner) * NC;
    for (int o = threadIdx.x; o < no * inner; o += kLooksBlock) {
        const int jp = o / inner, b = o - jp * inner;
        double sum[NC];
#pragma unroll
        for (int q = 0; q < NC; q++) sum[q] = 0.0;
        for (int k = 0; k < G.la; k++) {
            const int c = (jp * G.la + k) * inner + b;
#pragma unroll
            for (int q = 0; q < NC; q++) sum[q] += s_col[c * NC + q];
        }
#pragma unroll
        for (int q = 0; q < NC; q++) orow[(size_t)o * NC + q] = from_double<T>(sum[q] / norm);
    }
}

// one thread per output element; ESZ-byte elements are moved untouched
template <typename E>
__global__ void __launch_bounds__(kLooksBlock)
k_looks_nearest(const __grid_constant__ LooksGeom G, const E *__restrict__ in, E *__restrict__ out)
{
    const size_t per_line = (size_t)G.out_width * G.bands;
    const size_t n = per_line * (size_t)G.nlines;
    for (size_t t = (size_t)blockIdx.x * kLooksBlock + threadIdx.x; t < n; t += (size_t)gridDim.x * kLooksBlock) {
        const int r = (int)(t / per_line);
        const size_t u = t - (size_t)r * per_line;
        int band, col;
        if (G.scheme == kSchemeBIP) {
            col = (int)(u / G.bands);
            band = (int)(u - (size_t)col * G.bands);
        } else {
            band = (int)(u / G.out_width);
            col = (int)(u - (size_t)band * G.out_width);
        }
        const int oline = G.line0 + r;
        const int sl = oline * G.ld + G.ld / 2, sc = col * G.la + G.la / 2; // floor((i + 0.5) * looks)
        size_t si, di;
        if (G.scheme == kSchemeBIP) {
            si = ((size_t)sl * G.width + sc) * G.bands + band;
            di = ((size_t)oline * G.out_width + col) * G.bands + band;
        } else {
            si = row_offset(G, G.length, G.width, sl, band) + sc;
            di = row_offset(G, G.out_length, G.out_width, oline, band) + col;
        }
        out[di] = in[si];
    }
}

// numpy's .astype(int) of a float on x86-64 (cvttsd2si): truncation toward zero, "integer indefinite" (INT64_MIN) for
// NaN and out-of-range values
__device__ __forceinline__ long long trunc_like_x86(double v)
{
    if (!(v > -9.2233720368547758e18 && v < 9.2233720368547758e18)) return (long long)0x8000000000000000ULL;
    return (long long)v;
}

template <typename M, typename F>
__global__ void __launch_bounds__(256)
k_mask_to_radar(const __grid_constant__ MaskProj P, const M *__restrict__ mask, const F *__restrict__ lat, const F *__restrict__ lon,
                size_t npix, M *__restrict__ out)
{
    // with float32 coordinates numpy keeps the arithmetic in float32 (the Python scalars are weak)
    const F slat = (F)P.start_lat, dlat = (F)P.delta_lat, slon = (F)P.start_lon, dlon = (F)P.delta_lon;
    for (size_t p = (size_t)blockIdx.x * 256 + threadIdx.x; p < npix; p += (size_t)gridDim.x * 256) {
        long long li = trunc_like_x86((double)((lat[p] - slat) / dlat));
        long long lj = trunc_like_x86((double)((lon[p] - slon) / dlon));
        li = li < 0 ? 0 : (li > P.mask_length - 1 ? P.mask_length - 1 : li);
        lj = lj < 0 ? 0 : (lj > P.mask_width - 1 ? P.mask_width - 1 : lj);
        out[p] = (M)(mask[(size_t)li * P.mask_width + (size_t)lj] + (M)1);
    }
}

template <typename T, bool CPLX>
int looks_average(const LooksGeom &G, const void *in, void *out, cudaStream_t s)
{
    const int inner = (G.scheme == kSchemeBIP) ? G.bands : 1;
    const long long per_out = (long long)G.la * inner * (CPLX ? 2 : 1);
    if (per_out > kLooksMaxTile) return -2;
    // ~1024 column sums per CTA: enough CTAs (several waves on 148 SMs) to keep HBM busy, four columns per thread
    int tile_out = (int)((per_out >= 1024 ? kLooksMaxTile : 1024) / per_out);
    if (tile_out < 1) tile_out = 1;
    if (tile_out > G.out_width) tile_out = G.out_width;
    const int tiles = (G.out_width + tile_out - 1) / tile_out;
    const long long rows = (long long)G.nlines * (G.scheme == kSchemeBIP ? 1 : G.bands);
    const long long blocks = rows * tiles;
    if (blocks <= 0) return 0;
    if (blocks > 0x7fffffffLL) return -1;
    k_looks_average<T, CPLX><<<(unsigned)blocks, kLooksBlock, 0, s>>>(G, (const T *)in, (T *)out, tile_out, tiles);
    return 0;
}

template <typename E>
int looks_nearest(const LooksGeom &G, const void *in, void *out, cudaStream_t s)
{
    const size_t n = (size_t)G.out_width * G.bands * (size_t)G.nlines;
    if (n == 0) return 0;
    size_t blocks = (n + kLooksBlock - 1) / kLooksBlock;
    if (blocks > 148u * 32u) blocks = 148u * 32u;
    k_looks_nearest<E><<<(unsigned)blocks, kLooksBlock, 0, s>>>(G, (const E *)in, (E *)out);
    return 0;
}

} // namespace

int launch_looks(const LooksGeom &G, int dtype, int method, const void *in, void *out, cudaStream_t s)
{
    if (method == 1) {
        switch (type_size(dtype)) {
        case 1: return looks_nearest<unsigned char>(G, in, out, s);
        case 2: return looks_nearest<unsigned short>(G, in, out, s);
        case 4: return looks_nearest<unsigned int>(G, in, out, s);
        case 8: return looks_nearest<unsigned long long>(G, in, out, s);
        default: return -1;
        }
    }
    switch (dtype) {
    case kTypeByte: return looks_average<signed char, false>(G, in, out, s);
    case kTypeShort: return looks_average<short, false>(G, in, out, s);
    case kTypeInt: return looks_average<int, false>(G, in, out, s);
    case kTypeLong: return looks_average<long long, false>(G, in, out, s);
    case kTypeFloat: return looks_average<float, false>(G, in, out, s);
    case kTypeDouble: return looks_average<double, false>(G, in, out, s);
    case kTypeCFloat: return looks_average<float, true>(G, in, out, s);
    default: return -1;
    }
}

int launch_mask_to_radar(const MaskProj &M, int dtype, const void *mask, const void *lat, const void *lon, int coord_f32,
                         size_t npix, void *out, cudaStream_t s)
{
    if (npix == 0) return 0;
    size_t blocks = (npix + 255) / 256;
    if (blocks > 148u * 32u) blocks = 148u * 32u;
    const unsigned g = (unsigned)blocks;
#define B2_MASK_CASE(MT)                                                                                                     \
    if (coord_f32) k_mask_to_radar<MT, float><<<g, 256, 0, s>>>(M, (const MT *)mask, (const float *)lat, (const float *)lon, npix, (MT *)out); \
    else k_mask_to_radar<MT, double><<<g, 256, 0, s>>>(M, (const MT *)mask, (const double *)lat, (const double *)lon, npix, (MT *)out);        \
    return 0;
    switch (dtype) {
    case kTypeByte: B2_MASK_CASE(signed char)
    case kTypeShort: B2_MASK_CASE(short)
    case kTypeInt: B2_MASK_CASE(int)
    case kTypeFloat: B2_MASK_CASE(float)
    default: return -1;
    }
#undef B2_MASK_CASE
}

} // namespace b2
