// geozero_kernels.cu -- sm_100a kernels of geozero: geocoding of a radar-geometry image onto the DEM grid through the
// zero-Doppler (or native-Doppler) range-Doppler solve.
//
//   k_geozero_axes    sin / cos of every output row's latitude and column's longitude (the grid is separable:
//                     llh(1) = lat_first + idxlat*dlat, llh(2) = lon_first + idxlon*dlon, geozero.f90:276-281)
//   k_geozero_solve   one thread per output pixel: DEM height, LLH->XYZ, look-side test, the reference's fixed-point
//                     iteration on azimuth time (:322-356) on the per-window orbit polynomials, fractional image
//                     coordinates out (16 B/pixel) -- done ONCE per plan, every band / product reuses it
//   k_geozero_interp  one thread per output pixel and band: image bounds tests (:362-380) and the sinc / bilinear /
//                     bicubic / nearest gather of geozeroMethods.F in the reference's single-precision COMPLEX
//                     arithmetic, strided band in, strided band out
//
// The reference redoes the whole solve for every band of every product (Geozero.py:216-241); here the solve is
// amortised and the per-band pass is a pure gather bound by HBM / L2 (8-16 B of coordinates + 4-8 B out per pixel).
//
// Compiled with -fmad=false (see geom_device.cuh): the interpolators round exactly like the Fortran.
#include "geozero_kernels.cuh"

#include "dem_interp.cuh"
#include "orbit_poly_device.cuh"

namespace b2 {

__global__ void k_geozero_axes(const __grid_constant__ GeozeroConst C, GeozeroGeometry G)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C.geo_len) {
        const int idxlat = C.max_lat_idx + i;
        double s, c;
        sincos(C.lat_firstr + idxlat * C.dlatr, &s, &c);
        G.row_sc[2 * i] = s;
        G.row_sc[2 * i + 1] = c;
    }
    if (i < C.geo_wid) {
        const int idxlon = C.min_lon_idx + i;
        double s, c;
        sincos(C.lon_firstr + idxlon * C.dlonr, &s, &c);
        G.col_sc[2 * i] = s;
        G.col_sc[2 * i + 1] = c;
    }
}

__global__ void __launch_bounds__(kGeozeroBlock)
k_geozero_solve(const __grid_constant__ GeozeroConst C, OrbitPolyView op, GeozeroGeometry G, GeozeroStats *stats)
{
    const int bpl = (C.geo_wid + kGeozeroBlock - 1) / kGeozeroBlock;
    const int row = blockIdx.x / bpl;
    const int pix = (blockIdx.x - row * bpl) * blockDim.x + threadIdx.x;
    unsigned int n_it = 0;
    if (pix < C.geo_wid) {
        const size_t o = (size_t)row * (size_t)C.geo_wid + (size_t)pix;
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        double az_idx = qnan, rng_idx = qnan;
        double h = 0.0; // default height if the point is outside the DEM (:273)
        bool skip = false;
        const int idxlat = C.max_lat_idx + row;
        const int idxlon = C.min_lon_idx + pix;
        if (idxlat < 0 || idxlat > C.demlength - 1) {
            skip = true; // the whole line is skipped (:250-253): zeros out, zero DEM crop
        } else if (!(idxlon < 0 || idxlon > C.demwidth - 1)) {
            h = (double)C.dem[(size_t)(idxlat - C.dem_row0) * (size_t)C.dem_cols + (size_t)(idxlon - C.dem_col0)];
            if (h < -1500.0) skip = true; // bad SRTM pixels (:290-292)
        }
        if (!skip) {
            const Vec3 xyz = llh_to_xyz_sc(C.elp, G.row_sc[2 * row], G.row_sc[2 * row + 1], G.col_sc[2 * pix],
                                           G.col_sc[2 * pix + 1], h);
            // look side (:306-320)
            Vec3 dr = sub(xyz, C.xyz_mid);
            const double side = dot(cross(dr, C.vel_mid), C.xyz_mid);
            const int pixel_side = side > 0 ? -1 : 1;
            if (pixel_side == C.look_side) {
                const double t_lo = __ldg(op.t), t_hi = __ldg(op.t + op.n - 1);
                const double inv_fd = C.fd.order ? rcp_n(C.fd.norm) : 0.0, inv_fdd = C.fdd.order ? rcp_n(C.fdd.norm) : 0.0;
                double tline = C.tmid, rngpix = 0.0;
                OrbState S;
                S.x = C.xyz_mid;
                S.v = C.vel_mid;
                int hint = op.n >> 1;
#pragma unroll 1
                for (int k = 1; k <= 21; k++) { // :322-356
                    n_it++;
                    const double tprev = tline;
                    dr = sub(xyz, S.x);
                    rngpix = sqrt_p(dr.x * dr.x + dr.y * dr.y + dr.z * dr.z);
                    const double dopfact = div_n(dot(dr, S.v), rngpix);
                    const double fdop = 0.5 * C.wvl * poly1d_fast(C.fd, inv_fd, rngpix);
                    const double fdopder = 0.5 * C.wvl * poly1d_fast(C.fdd, inv_fdd, rngpix);
                    const double c1 = dopfact - fdop;
                    const double c2 = div_n(dot(S.v, S.v), rngpix);
                    const double c3 = dopfact * (div_n(fdop, rngpix) + fdopder);
                    tline = tline + div_n(c1, c2 - c3);
                    if ((tline < t_lo) || (tline > t_hi) || !(tline == tline)) { // interpolator stat != 0 (orbit.c:224-233)
                        tline = -10000.0; // BAD_VALUE (:70)
                        rngpix = -10000.0;
                        break;
                    }
                    poly_state<0>(op, tline, S, hint);
                    if (fabs(tline - tprev) < 5.0e-7) break;
                }
                az_idx = div_n(tline - C.tstart, C.dtaz) + 1; // :359-360
                rng_idx = div_n(rngpix - C.rngstart, C.dmrg) + 1;
            }
        }
        G.az_idx[o] = az_idx;
        G.rng_idx[o] = rng_idx;
        G.dem_crop[o] = (short)h; // integer*2 truncation of llh(3) (:399)
    }
    for (int sft = 16; sft > 0; sft >>= 1) n_it += __shfl_xor_sync(0xffffffffu, n_it, sft);
    if ((threadIdx.x & 31) == 0 && n_it) atomicAdd(&stats->iterations, (unsigned long long)n_it);
}

// ---- Fortran default COMPLEX (2 x real*4) and its promotion to COMPLEX*16 against real*8 operands ----
struct C4 {
    float re, im;
};
struct C8 {
    double re, im;
};
__device__ __forceinline__ C8 c8(C4 a) { return C8{(double)a.re, (double)a.im}; }
__device__ __forceinline__ C4 c4(C8 a) { return C4{(float)a.re, (float)a.im}; }
__device__ __forceinline__ C8 cscale(C8 a, double s) { return C8{a.re * s, a.im * s}; }
__device__ __forceinline__ C8 cdiv(C8 a, double s) { return C8{div_n(a.re, s), div_n(a.im, s)}; }
__device__ __forceinline__ C8 cadd(C8 a, C8 b) { return C8{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ C4 csub4(C4 a, C4 b) { return C4{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ C4 cadd4(C4 a, C4 b) { return C4{a.re + b.re, a.im + b.im}; }

// ifg(r, a): sample r (1-based, range) of line a (1-based, azimuth) of the band
template <bool CPLX>
struct Ifg {
    const void *base;
    BandView v;
    __device__ __forceinline__ C4 operator()(int r, int a) const
    {
        const size_t i = v.offset + (size_t)(a - 1) * v.line_stride + (size_t)(r - 1) * v.pix_stride;
        if (CPLX) {
            const float2 z = __ldg(reinterpret_cast<const float2 *>(base) + i);
            return C4{z.x, z.y};
        }
        return C4{__ldg(reinterpret_cast<const float *>(base) + i), 0.f}; // readRealLine: cmplx(rarr(i), 0.)
    }
};

// uniform_interp.f90:46-77 as called from geozeroMethods.F:103-117: bilinear_cx(dy, dx, ifg)
template <bool CPLX>
__device__ __forceinline__ C4 gz_bilinear(const Ifg<CPLX> &ifg, double x, double y)
{
    const double x1 = floor(x), x2 = ceil(x), y1 = ceil(y), y2 = floor(y);
    const C4 q11 = ifg((int)y1, (int)x1), q12 = ifg((int)y2, (int)x1), q21 = ifg((int)y1, (int)x2), q22 = ifg((int)y2, (int)x2);
    if (y1 == y2 && x1 == x2) return q11;
    if (y1 == y2) return c4(cadd(cscale(c8(q11), div_n(x2 - x, x2 - x1)), cscale(c8(q21), div_n(x - x1, x2 - x1))));
    if (x1 == x2) return c4(cadd(cscale(c8(q11), div_n(y2 - y, y2 - y1)), cscale(c8(q12), div_n(y - y1, y2 - y1))));
    const double den = (x2 - x1) * (y2 - y1);
    C8 s = cdiv(cscale(cscale(c8(q11), (x2 - x)), (y2 - y)), den);
    s = cadd(s, cdiv(cscale(cscale(c8(q21), (x - x1)), (y2 - y)), den));
    s = cadd(s, cdiv(cscale(cscale(c8(q12), (x2 - x)), (y - y1)), den));
    s = cadd(s, cdiv(cscale(cscale(c8(q22), (x - x1)), (y - y1)), den));
    return c4(s);
}

// uniform_interp.f90:123-130 DATA wt (column-major fill): wt(i,k) = kWt[(k-1)*16 + (i-1)]; small integers
__constant__ signed char kWt[256] = {
    1, 0, -3, 2, 0, 0, 0, 0, -3, 0, 9, -6, 2, 0, -6, 4,
    0, 0, 0, 0, 0, 0, 0, 0, 3, 0, -9, 6, -2, 0, 6, -4,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 9, -6, 0, 0, -6, 4,
    0, 0, 3, -2, 0, 0, 0, 0, 0, 0, -9, 6, 0, 0, 6, -4,
    0, 0, 0, 0, 1, 0, -3, 2, -2, 0, 6, -4, 1, 0, -3, 2,
    0, 0, 0, 0, 0, 0, 0, 0, -1, 0, 3, -2, 1, 0, -3, 2,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -3, 2, 0, 0, 3, -2,
    0, 0, 0, 0, 0, 0, 3, -2, 0, 0, -6, 4, 0, 0, 3, -2,
    0, 1, -2, 1, 0, 0, 0, 0, 0, -3, 6, -3, 0, 2, -4, 2,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 3, -6, 3, 0, -2, 4, -2,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -3, 3, 0, 0, 2, -2,
    0, 0, -1, 1, 0, 0, 0, 0, 0, 0, 3, -3, 0, 0, -2, 2,
    0, 0, 0, 0, 0, 1, -2, 1, 0, -2, 4, -2, 0, 1, -2, 1,
    0, 0, 0, 0, 0, 0, 0, 0, 0, -1, 2, -1, 0, 1, -2, 1,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, 0, 0, -1, 1,
    0, 0, 0, 0, 0, 0, -1, 1, 0, 0, 2, -2, 0, 0, -1, 1};

// uniform_interp.f90:203-292 as called from geozeroMethods.F:119-130: bicubic_cx(dy, dx, ifg).  Every local of the
// reference is default COMPLEX, so each assignment rounds to real*4 components; zero table entries add (+-)0 and are
// skipped (x + 0 == x; the accumulator starts at +0, so no signed-zero difference can surface in a non-zero sum).
template <bool CPLX>
__device__ C4 gz_bicubic(const Ifg<CPLX> &ifg, double x, double y)
{
    const int x1 = (int)floor(x), x2 = (int)ceil(x), y1 = (int)floor(y), y2 = (int)ceil(y);
    C4 q[16];
    q[0] = ifg(y1, x1);
    q[3] = ifg(y2, x1);
    q[1] = ifg(y1, x2);
    q[2] = ifg(y2, x2);
#define B2_HALF(a) c4(cscale(c8(a), 0.5))
    q[4] = B2_HALF(csub4(ifg(y1, x1 + 1), ifg(y1, x1 - 1)));
    q[5] = B2_HALF(csub4(ifg(y1, x2 + 1), ifg(y1, x2 - 1)));
    q[6] = B2_HALF(csub4(ifg(y2, x2 + 1), ifg(y2, x2 - 1)));
    q[7] = B2_HALF(csub4(ifg(y2, x1 + 1), ifg(y2, x1 - 1)));
    q[8] = B2_HALF(csub4(ifg(y1 + 1, x1), ifg(y1 - 1, x1)));
    q[9] = B2_HALF(csub4(ifg(y1 + 1, x2 + 1), ifg(y1 - 1, x2))); // :243-245 column typo kept
    q[10] = B2_HALF(csub4(ifg(y2 + 1, x2 + 1), ifg(y2 - 1, x2)));
    q[11] = B2_HALF(csub4(ifg(y2 + 1, x1 + 1), ifg(y2 - 1, x1)));
#undef B2_HALF
#define B2_CROSS(yy, xx)                                                                                             \
    c4(cscale(c8(cadd4(csub4(csub4(ifg((yy) + 1, (xx) + 1), ifg((yy)-1, (xx) + 1)), ifg((yy) + 1, (xx)-1)), ifg((yy)-1, (xx)-1))), \
              0.25))
    q[12] = B2_CROSS(y1, x1);
    q[15] = B2_CROSS(y2, x1);
    q[13] = B2_CROSS(y1, x2);
    q[14] = B2_CROSS(y2, x2);
#undef B2_CROSS
    const double t = (x - x1), u = (y - y1);
    C4 r = C4{0.f, 0.f};
#pragma unroll 1
    for (int i = 3; i >= 0; i--) { // c(i+1, j+1) = cl(4*i + j + 1)
        C4 c[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int l = 4 * i + j;
            C4 qq = C4{0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int w = kWt[k * 16 + l];
                if (w != 0) qq = c4(cadd(c8(qq), cscale(c8(q[k]), (double)w)));
            }
            c[j] = qq;
        }
        // bicubic_cx = t*bicubic_cx + ((c(i,4)*u + c(i,3))*u + c(i,2))*u + c(i,1), left to right
        C8 hh = cadd(cscale(c8(c[3]), u), c8(c[2]));
        hh = cscale(cadd(cscale(hh, u), c8(c[1])), u);
        r = c4(cadd(cadd(cscale(c8(r), t), hh), c8(c[0])));
    }
    return r;
}

// uniform_interp.f90:456-484 through geozeroMethods.F:91-101 (intp_sinc: i_xx = i_x - 1, i_yy = i_y - 1)
template <bool CPLX>
__device__ C4 gz_sinc(const Ifg<CPLX> &ifg, const float *__restrict__ intarr, int intpx, int intpy, double frpx, double frpy,
                      int width, int length)
{
    constexpr int idec = kSincSub, ilen = kSincLen;
    C4 acc = C4{0.f, 0.f};
    if ((intpx >= ilen - 1 && intpx < width) && (intpy >= ilen - 1 && intpy < length)) {
        int ifracx = (int)(frpx * idec), ifracy = (int)(frpy * idec);
        ifracx = ifracx < 0 ? 0 : (ifracx > idec - 1 ? idec - 1 : ifracx);
        ifracy = ifracy < 0 ? 0 : (ifracy > idec - 1 ? idec - 1 : ifracy);
        float wy[ilen];
#pragma unroll
        for (int m = 0; m < ilen; m++) wy[m] = __ldg(intarr + m + ifracy * ilen);
        double fweightsum = 0.0;
#pragma unroll 1
        for (int k = 0; k < ilen; k++) {
            const float wx = __ldg(intarr + k + ifracx * ilen);
#pragma unroll
            for (int m = 0; m < ilen; m++) {
                const double fweight = (double)(wx * wy[m]); // real*4 product, then real*8
                acc = c4(cadd(c8(acc), cscale(c8(ifg(intpx - k + 1, intpy - m + 1)), fweight)));
                fweightsum = fweightsum + fweight;
            }
        }
        acc = c4(cdiv(c8(acc), fweightsum));
    }
    return acc;
}

template <int METHOD, bool CPLX>
__global__ void __launch_bounds__(kGeozeroBlock)
k_geozero_interp(const __grid_constant__ GeozeroConst C, GeozeroGeometry G, const void *image, BandView in, void *out,
                 BandView ov, const float *__restrict__ sinc, GeozeroStats *stats)
{
    const int bpl = (C.geo_wid + kGeozeroBlock - 1) / kGeozeroBlock;
    const int row = blockIdx.x / bpl;
    const int pix = (blockIdx.x - row * bpl) * blockDim.x + threadIdx.x;
    unsigned int n_out = 0, n_valid = 0;
    if (pix < C.geo_wid) {
        const size_t o = (size_t)row * (size_t)C.geo_wid + (size_t)pix;
        const double az_idx = G.az_idx[o], rng_idx = G.rng_idx[o];
        C4 z = C4{0.f, 0.f};
        if (az_idx == az_idx) { // NaN: the reference jumped to label 100 before the solve
            const float f_delay = METHOD == 0 ? kSincLen / 2.0f : (METHOD == 2 ? 3.0f : 2.0f); // geozeroMethods.F:66-79
            // :362-380; width - f_delay is a real*4 expression in the reference
            if (rng_idx <= (double)f_delay || rng_idx >= (double)((float)C.width - f_delay) || az_idx <= (double)f_delay ||
                az_idx >= (double)((float)C.length - f_delay)) {
                n_out = 1;
            } else {
                n_valid = 1;
                const int int_rdx = (int)(rng_idx + f_delay);
                const double fr_rdx = rng_idx + f_delay - int_rdx;
                const int int_rdy = (int)(az_idx + f_delay);
                const double fr_rdy = az_idx + f_delay - int_rdy;
                const Ifg<CPLX> ifg{image, in};
                if (METHOD == 0) {
                    z = gz_sinc<CPLX>(ifg, sinc, int_rdx - 1, int_rdy - 1, fr_rdx, fr_rdy, C.width, C.length);
                } else {
                    const double dx = int_rdx + fr_rdx - f_delay, dy = int_rdy + fr_rdy - f_delay;
                    if (METHOD == 1) z = gz_bilinear<CPLX>(ifg, dy, dx);
                    else if (METHOD == 2) z = gz_bicubic<CPLX>(ifg, dy, dx);
                    else z = ifg((int)llround(dx), (int)llround(dy)); // nint
                }
            }
        }
        const size_t oo = ov.offset + (size_t)row * ov.line_stride + (size_t)pix * ov.pix_stride;
        if (CPLX) reinterpret_cast<float2 *>(out)[oo] = make_float2(z.re, z.im);
        else reinterpret_cast<float *>(out)[oo] = z.re; // writeRealLine: real(carr)
    }
    for (int sft = 16; sft > 0; sft >>= 1) {
        n_out += __shfl_xor_sync(0xffffffffu, n_out, sft);
        n_valid += __shfl_xor_sync(0xffffffffu, n_valid, sft);
    }
    if ((threadIdx.x & 31) == 0 && stats) {
        if (n_out) atomicAdd(&stats->outside_image, (unsigned long long)n_out);
        if (n_valid) atomicAdd(&stats->valid, (unsigned long long)n_valid);
    }
}

void launch_geozero_axes(const GeozeroConst &C, const GeozeroGeometry &G, cudaStream_t s)
{
    const int n = C.geo_len > C.geo_wid ? C.geo_len : C.geo_wid;
    k_geozero_axes<<<(n + 127) / 128, 128, 0, s>>>(C, G);
}

int launch_geozero_solve(const GeozeroConst &C, const OrbitPolyView &op, const GeozeroGeometry &G, GeozeroStats *stats,
                         cudaStream_t s)
{
    const long long nblk = (long long)((C.geo_wid + kGeozeroBlock - 1) / kGeozeroBlock) * C.geo_len;
    if (nblk > 0x7fffffffLL || nblk < 1) return -2;
    k_geozero_solve<<<(unsigned)nblk, kGeozeroBlock, 0, s>>>(C, op, G, stats);
    return 0;
}

template <int METHOD>
static void launch_interp_m(const GeozeroConst &C, const GeozeroGeometry &G, int is_complex, const void *image, BandView in,
                            void *out, BandView ov, const float *sinc, GeozeroStats *stats, unsigned g, cudaStream_t s)
{
    if (is_complex) k_geozero_interp<METHOD, true><<<g, kGeozeroBlock, 0, s>>>(C, G, image, in, out, ov, sinc, stats);
    else k_geozero_interp<METHOD, false><<<g, kGeozeroBlock, 0, s>>>(C, G, image, in, out, ov, sinc, stats);
}

int launch_geozero_interp(const GeozeroConst &C, const GeozeroGeometry &G, int method, int is_complex, const void *image,
                          BandView in, void *out, BandView ov, const float *sinc, GeozeroStats *stats, cudaStream_t s)
{
    const long long nblk = (long long)((C.geo_wid + kGeozeroBlock - 1) / kGeozeroBlock) * C.geo_len;
    if (nblk > 0x7fffffffLL || nblk < 1) return -2;
    const unsigned g = (unsigned)nblk;
    switch (method) {
    case 0:
        if (!sinc) return -3;
        launch_interp_m<0>(C, G, is_complex, image, in, out, ov, sinc, stats, g, s);
        break;
    case 1: launch_interp_m<1>(C, G, is_complex, image, in, out, ov, sinc, stats, g, s); break;
    case 2: launch_interp_m<2>(C, G, is_complex, image, in, out, ov, sinc, stats, g, s); break;
    case 3: launch_interp_m<3>(C, G, is_complex, image, in, out, ov, sinc, stats, g, s); break;
    default: return -1;
    }
    return 0;
}

} // namespace b2
