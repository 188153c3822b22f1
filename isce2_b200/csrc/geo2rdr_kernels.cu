// geo2rdr_kernels.cu -- sm_100a kernels of the zero-Doppler geo2rdr path.
//
//   k_geo_setup   1 thread: orbit state and finite-difference acceleration at mid-scene (geo2rdr.f90:194-208)
//   k_geo2rdr     one thread per lat/lon/hgt pixel: LLH->XYZ, Newton solve of the range-Doppler equation on
//                 azimuth time with the orbit re-interpolated every step (state vectors staged in shared
//                 memory), validity tests, optional bistatic correction, range/azimuth (+offset) outputs
//                 (geo2rdr.f90:215-403)
//
// Compiled with -fmad=false (see geom_device.cuh).
#include "geo2rdr_kernels.cuh"
#include "orbit_poly_device.cuh"

namespace b2 {

__global__ void k_geo_setup(int orbit_method, OrbitView orb, double tmid, GeoMid *out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Vec3 p = v3(0, 0, 0), v = v3(0, 0, 0);
    out->stat_mid = orbit_interp(orbit_method, orb, tmid, p, v);
    out->xyz[0] = p.x; out->xyz[1] = p.y; out->xyz[2] = p.z;
    out->vel[0] = v.x; out->vel[1] = v.y; out->vel[2] = v.z;
    // computeAcceleration (orbit.c:316-354): always Hermite, +-0.01 s
    Vec3 xb, vb, xa, va;
    int sb = orbit_hermite(orb, tmid - 0.01, xb, vb);
    int sa = sb == 0 ? orbit_hermite(orb, tmid + 0.01, xa, va) : 1;
    out->stat_acc = (sb != 0 || sa != 0) ? 1 : 0;
    if (out->stat_acc == 0) {
        out->acc[0] = (va.x - vb.x) / 0.02;
        out->acc[1] = (va.y - vb.y) / 0.02;
        out->acc[2] = (va.z - vb.z) / 0.02;
    } else {
        out->acc[0] = out->acc[1] = out->acc[2] = 0.0;
    }
}

template <typename T>
__device__ __forceinline__ void store_out(void *base, size_t o, double v)
{
    if (base) reinterpret_cast<T *>(base)[o] = (T)v; // 'single': DoubleToFloat write caster (Geo2rdr.py:326-381)
}

template <typename T>
__global__ void __launch_bounds__(kGeoBlock)
k_geo2rdr(const __grid_constant__ GeoConst C, OrbitView orb_g, int line0, int nlines, GeoLayers L, GeoStats *stats)
{
    extern __shared__ double s_orb[]; // t[n], pos[3n], vel[3n]
    __shared__ unsigned int s_cnt[4][kGeoBlock / 32];
    OrbitView orb = orb_g;
    if (orb_g.n <= kGeoMaxSmemVectors) {
        const int n = orb_g.n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) s_orb[i] = orb_g.t[i];
        for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) {
            s_orb[n + i] = orb_g.pos[i];
            s_orb[4 * n + i] = orb_g.vel[i];
        }
        __syncthreads();
        orb.t = s_orb;
        orb.pos = s_orb + n;
        orb.vel = s_orb + 4 * n;
    }
    const int bpl = (C.demwidth + kGeoBlock - 1) / kGeoBlock;
    const int row = blockIdx.x / bpl;
    const int pix = (blockIdx.x - row * bpl) * blockDim.x + threadIdx.x;
    unsigned int n_out = 0, n_valid = 0, n_conv = 0, n_it = 0;
    if (pix < C.demwidth) {
        const double BAD_VALUE = (double)(-999999.0f); // geo2rdr.f90:59-60
        const int line = line0 + row;                   // 0-based line of the lat/lon/hgt images
        const size_t o = (size_t)row * (size_t)C.demwidth + (size_t)pix;
        double azt = BAD_VALUE, rgm = BAD_VALUE, rgoff = BAD_VALUE, azoff = BAD_VALUE;
        Vec3 xyz = C.xyz_in ? Vec3{L.lat[o], L.lon[o], L.hgt[o]} : llh_to_xyz(C.elp, L.lat[o] * C.deg2rad, L.lon[o] * C.deg2rad, L.hgt[o]);
        double tline = C.tmid, tprev = 0.0, rngpix = 0.0;
        Vec3 satx = C.xyz_mid, satv = C.vel_mid;
        const Vec3 sata = C.acc_mid;
        for (int k = 1; k <= 51; k++) { // :259-305
            n_it++;
            tprev = tline;
            Vec3 dr = sub(xyz, satx);
            rngpix = norm(dr);
            double dopfact = dot(dr, satv);
            double fdop = 0.5 * C.wvl * eval_poly1d(C.fd, rngpix);
            double fdopder = 0.5 * C.wvl * eval_poly1d(C.fdd, rngpix);
            double fn = dopfact - fdop * rngpix;
            double c1 = (0.0 * dot(sata, dr) - dot(satv, satv));
            double c2 = (fdop / rngpix + fdopder);
            double fnprime = c1 + c2 * dopfact;
            tline = tline - fn / fnprime;
            int stat = orbit_interp(C.orbit_method, orb, tline, satx, satv);
            if (stat != 0) {
                tline = BAD_VALUE;
                rngpix = BAD_VALUE;
                break;
            }
            if (fabs(tline - tprev) < 5.0e-9) {
                n_conv = 1;
                break;
            }
        }
        bool outside = false;
        if (tline < C.tstart) outside = true;
        else if (tline > C.tend) outside = true;
        else {
            rngpix = norm(sub(xyz, satx));
            if (rngpix < C.rngstart) outside = true;
            else if (rngpix > C.rngend) outside = true;
            else if (C.bistatic) { // :331-368
                tline = tline + 2.0 * rngpix / C.sol;
                if (tline < C.tstart) outside = true;
                else if (tline > C.tend) outside = true;
                else {
                    int stat = orbit_interp(C.orbit_method, orb, tline, satx, satv);
                    if (stat != 0) outside = true;
                    else {
                        rngpix = norm(sub(xyz, satx));
                        if (rngpix < C.rngstart) outside = true;
                        else if (rngpix > C.rngend) outside = true;
                    }
                }
            }
        }
        if (outside) n_out = 1;
        else { // :370-376
            n_valid = 1;
            rgm = rngpix;
            azt = tline;
            rgoff = ((rngpix - C.rngstart) / C.dmrg) - 1.0 * ((pix + 1) - 1);
            azoff = ((tline - C.tstart) / C.dtaz) - 1.0 * ((line + 1) - 1);
        }
        store_out<T>(L.azt, o, azt);
        store_out<T>(L.rgm, o, rgm);
        store_out<T>(L.azoff, o, azoff);
        store_out<T>(L.rgoff, o, rgoff);
    }
    // block-level counters (the three prints at geo2rdr.f90:407-409)
    unsigned int v[4] = {n_out, n_valid, n_conv, n_it};
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        unsigned int x = v[q];
        for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
        if (lane == 0) s_cnt[q][wid] = x;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        unsigned long long tot = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); wv++) tot += s_cnt[threadIdx.x][wv];
        unsigned long long *dst = threadIdx.x == 0 ? &stats->outside : threadIdx.x == 1 ? &stats->valid
                                  : threadIdx.x == 2 ? &stats->converged : &stats->iterations;
        atomicAdd(dst, tot);
    }
}

// -------------------------------------------------------------------------------------------------
// Newton solve on per-window orbit polynomials
// -------------------------------------------------------------------------------------------------
// The reference iterates  t <- t - fn/fnprime  with fnprime = -v.v + (fdop/r + fdop')(dr.v): the acceleration term
// is multiplied by zero (geo2rdr.f90:271), so it converges linearly (ratio ~0.1) and stops at |dt| < 5e-9 s, i.e. up
// to ~6e-10 s (3e-7 azimuth pixels) short of the root of fn(t) = dr.v - fdop(r) r.  This kernel solves the same
// equation for the same root with the true derivative (acceleration from the orbit polynomial; the mid-scene
// finite-difference acceleration of :206 for the first step) until |dt| < 1e-10 s: 3 steps instead of the reference's
// 9-11.  The reference's own first iterate is still formed, because it is the only one that can overshoot the orbit
// span by seconds and the reference invalidates the pixel when it does (geo2rdr.f90:287-291).  Final range / validity
// tests are the reference's (:308-329), evaluated at the converged state.
// Each thread walks kGeoPixPerThread pixels of its line (stride kGeoBlock, so that a warp's loads and stores stay
// coalesced) and loads the next pixel's lat / lon / hgt before it solves the current one: with one pixel per thread
// the warps of a CTA all sat on the DRAM latency of their three input loads at the same moment (ncu: 43 % of the stall
// samples on the first use of lat / lon / hgt).
#ifndef B2_GEO_PPT
#define B2_GEO_PPT 4 // tunable (tools/build_variant.sh NAME "-DB2_GEO_PPT=8")
#endif
constexpr int kGeoPixPerThread = B2_GEO_PPT;

template <int METHOD, typename T>
__global__ void __launch_bounds__(kGeoBlock)
k_geo2rdr_poly(const __grid_constant__ GeoConst C, OrbitPolyView op, int line0, int nlines, GeoLayers L, GeoStats *stats)
{
    constexpr int kSpan = kGeoBlock * kGeoPixPerThread;
    const int bpl = (C.demwidth + kSpan - 1) / kSpan;
    const int row = blockIdx.x / bpl;
    const int pix0 = (blockIdx.x - row * bpl) * kSpan + threadIdx.x;
    const size_t rowoff = (size_t)row * (size_t)C.demwidth;
    const int line = line0 + row;
    const double BAD_VALUE = (double)(-999999.0f);
    const double t_lo = __ldg(op.t), t_hi = __ldg(op.t + op.n - 1);
    const double inv_fd = C.fd.order ? rcp_n(C.fd.norm) : 0.0, inv_fdd = C.fdd.order ? rcp_n(C.fdd.norm) : 0.0;
    unsigned int n_out = 0, n_valid = 0, n_conv = 0, n_it = 0;
    int hint = op.n >> 1; // orbit window of this thread's previous query (poly_window)
    double nlat = 0.0, nlon = 0.0, nhgt = 0.0;
    if (pix0 < C.demwidth) {
        nlat = L.lat[rowoff + pix0];
        nlon = L.lon[rowoff + pix0];
        nhgt = L.hgt[rowoff + pix0];
    }
#pragma unroll 1
    for (int j = 0; j < kGeoPixPerThread; j++) {
        const int pix = pix0 + j * kGeoBlock;
        if (pix >= C.demwidth) break;
        const double lat = nlat, lon = nlon, hgt = nhgt;
        if (j + 1 < kGeoPixPerThread && pix + kGeoBlock < C.demwidth) { // prefetch the next pixel's inputs
            nlat = L.lat[rowoff + pix + kGeoBlock];
            nlon = L.lon[rowoff + pix + kGeoBlock];
            nhgt = L.hgt[rowoff + pix + kGeoBlock];
        }
        const size_t o = rowoff + (size_t)pix;
        double azt = BAD_VALUE, rgm = BAD_VALUE, rgoff = BAD_VALUE, azoff = BAD_VALUE;
        const Vec3 xyz = C.xyz_in ? Vec3{lat, lon, hgt} : llh_to_xyz(C.elp, lat * C.deg2rad, lon * C.deg2rad, hgt);
        double tline = C.tmid, rngpix = 0.0;
        OrbState S;
        S.x = C.xyz_mid;
        S.v = C.vel_mid;
        S.a = C.acc_mid;
        bool bad = false;
#pragma unroll 1
        for (int k = 1; k <= 51; k++) {
            n_it++;
            const Vec3 dr = sub(xyz, S.x);
            rngpix = sqrt_p(dr.x * dr.x + dr.y * dr.y + dr.z * dr.z);
            const double dopfact = dot(dr, S.v);
            const double fdop = 0.5 * C.wvl * poly1d_fast(C.fd, inv_fd, rngpix);
            const double fdopder = 0.5 * C.wvl * poly1d_fast(C.fdd, inv_fdd, rngpix);
            const double fn = dopfact - fdop * rngpix;
            const double vv = dot(S.v, S.v);
            const double c2 = div_n(fdop, rngpix) + fdopder;
            if (k == 1) {
                // the reference's own first iterate (derivative without the acceleration term, geo2rdr.f90:271) is the only
                // one that can overshoot the orbit span by seconds: its out-of-span test (:287-291) is reproduced on it
                const double tref = tline - div_n(fn, (0.0 - vv) + c2 * dopfact);
                if ((tref < t_lo) || (tref > t_hi) || !(tref == tref)) {
                    bad = true;
                    break;
                }
            }
            const double tnew = tline - div_n(fn, (dot(S.a, dr) - vv) + c2 * dopfact); // true Newton step
            const double step = tnew - tline;
            tline = tnew;
            if ((tline < t_lo) || (tline > t_hi) || !(tline == tline)) { // interpolator stat != 0 (orbit.c:224-233)
                bad = true;
                break;
            }
            poly_state<METHOD>(op, tline, S, hint);
            if (fabs(step) < 1.0e-10) {
                n_conv++;
                break;
            }
        }
        bool outside = bad;
        if (!outside) {
            if (tline < C.tstart) outside = true;
            else if (tline > C.tend) outside = true;
            else {
                const Vec3 dr = sub(xyz, S.x);
                rngpix = sqrt_p(dr.x * dr.x + dr.y * dr.y + dr.z * dr.z);
                if (rngpix < C.rngstart) outside = true;
                else if (rngpix > C.rngend) outside = true;
                else if (C.bistatic) { // :331-368
                    tline = tline + 2.0 * rngpix / C.sol;
                    if (tline < C.tstart) outside = true;
                    else if (tline > C.tend) outside = true;
                    else if ((tline < t_lo) || (tline > t_hi)) outside = true;
                    else {
                        poly_state<METHOD>(op, tline, S, hint);
                        const Vec3 d2 = sub(xyz, S.x);
                        rngpix = sqrt_p(d2.x * d2.x + d2.y * d2.y + d2.z * d2.z);
                        if (rngpix < C.rngstart) outside = true;
                        else if (rngpix > C.rngend) outside = true;
                    }
                }
            }
        }
        if (outside) n_out++;
        else {
            n_valid++;
            rgm = rngpix;
            azt = tline;
            rgoff = div_n(rngpix - C.rngstart, C.dmrg) - 1.0 * ((pix + 1) - 1);
            azoff = div_n(tline - C.tstart, C.dtaz) - 1.0 * ((line + 1) - 1);
        }
        store_out<T>(L.azt, o, azt);
        store_out<T>(L.rgm, o, rgm);
        store_out<T>(L.azoff, o, azoff);
        store_out<T>(L.rgoff, o, rgoff);
    }
    // counters without a block-wide barrier (warps leave as they finish)
    unsigned int v[4] = {n_out, n_valid, n_conv, n_it};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        unsigned int x = v[q];
        for (int sft = 16; sft > 0; sft >>= 1) x += __shfl_xor_sync(0xffffffffu, x, sft);
        v[q] = x;
    }
    if ((threadIdx.x & 31) == 0) {
        unsigned long long *dst[4] = {&stats->outside, &stats->valid, &stats->converged, &stats->iterations};
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (v[q]) atomicAdd(dst[q], (unsigned long long)v[q]);
    }
}

__global__ void __launch_bounds__(256)
k_llh_to_xyz(const __grid_constant__ GeoConst C, const double *__restrict__ lat, const double *__restrict__ lon,
             const double *__restrict__ hgt, double *__restrict__ x, double *__restrict__ y, double *__restrict__ z, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const Vec3 v = llh_to_xyz(C.elp, lat[i] * C.deg2rad, lon[i] * C.deg2rad, hgt[i]);
        x[i] = v.x;
        y[i] = v.y;
        z[i] = v.z;
    }
}

void launch_llh_to_xyz(const GeoConst &C, const double *lat, const double *lon, const double *hgt, double *x, double *y, double *z,
                       size_t n, cudaStream_t s)
{
    if (n == 0) return;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    k_llh_to_xyz<<<(unsigned)blocks, 256, 0, s>>>(C, lat, lon, hgt, x, y, z, n);
}

void launch_geo_setup(int orbit_method, const OrbitView &orb, double tmid, GeoMid *d_out, cudaStream_t s)
{
    k_geo_setup<<<1, 32, 0, s>>>(orbit_method, orb, tmid, d_out);
}

int launch_geo2rdr(const GeoConst &C, const OrbitView &orb, int line0, int nlines, const GeoLayers &L, int out_f32,
                   GeoStats *stats, cudaStream_t s)
{
    const long long nblk = (long long)((C.demwidth + kGeoBlock - 1) / kGeoBlock) * nlines;
    if (nblk > 0x7fffffffLL) return -2;
    size_t smem = orb.n <= kGeoMaxSmemVectors ? (size_t)orb.n * 7 * sizeof(double) : 0;
    if (out_f32) k_geo2rdr<float><<<(unsigned)nblk, kGeoBlock, smem, s>>>(C, orb, line0, nlines, L, stats);
    else k_geo2rdr<double><<<(unsigned)nblk, kGeoBlock, smem, s>>>(C, orb, line0, nlines, L, stats);
    return 0;
}

int launch_geo2rdr_poly(const GeoConst &C, const OrbitPolyView &op, int line0, int nlines, const GeoLayers &L, int out_f32,
                        GeoStats *stats, cudaStream_t s)
{
    constexpr int kSpan = kGeoBlock * kGeoPixPerThread;
    const long long nblk = (long long)((C.demwidth + kSpan - 1) / kSpan) * nlines;
    if (nblk > 0x7fffffffLL) return -2;
    const unsigned g = (unsigned)nblk;
    if (op.method == 0) {
        if (out_f32) k_geo2rdr_poly<0, float><<<g, kGeoBlock, 0, s>>>(C, op, line0, nlines, L, stats);
        else k_geo2rdr_poly<0, double><<<g, kGeoBlock, 0, s>>>(C, op, line0, nlines, L, stats);
    } else if (op.method == 2) {
        if (out_f32) k_geo2rdr_poly<2, float><<<g, kGeoBlock, 0, s>>>(C, op, line0, nlines, L, stats);
        else k_geo2rdr_poly<2, double><<<g, kGeoBlock, 0, s>>>(C, op, line0, nlines, L, stats);
    } else {
        return -1;
    }
    return 0;
}

} // namespace b2
