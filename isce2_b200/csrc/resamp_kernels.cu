// resamp_kernels.cu -- sm_100a kernels of resamp_slc: resampling of a complex SLC with polynomial + per-pixel offsets
// (the range / azimuth .off rasters of geo2rdr), carriers and Doppler handled as in the reference.
//
//   k_resamp_carrier  one thread per INPUT pixel: cin = cline * exp(-i (rgCarrier + azCarrier)) (resamp_slc.f90:122-141)
//   k_resamp_slc      one thread per OUTPUT pixel: offsets, bounds, Doppler de-rotation of the 8 chip rows the sinc
//                     uses, 8 x 8 normalised-sinc gather in the reference's single-precision COMPLEX arithmetic, carrier
//                     / Doppler / flattening phase put back (:166-262)
//
// HBM-bound by construction: 8 B in (each input pixel is touched by ~64 output pixels, from L1 / L2), 8-16 B of
// offsets, 8 B out per pixel.  Compiled with -fmad=false: products and sums round separately like the x86-64 build.
#include "resamp_kernels.cuh"

#include "dem_interp.cuh" // kSincSub, kSincLen

namespace b2 {

__device__ __forceinline__ float2 cmul4(float2 a, float2 b)
{
    // default COMPLEX product: four real*4 products, one difference, one sum
    const float t1 = a.x * b.x, t2 = a.y * b.y, t3 = a.x * b.y, t4 = a.y * b.x;
    return make_float2(t1 - t2, t3 + t4);
}

// MODULO(a, p) for reals as gfortran expands it (fmod, then shifted into the sign of p); fmod is exact
__device__ __forceinline__ double f_modulo(double a, double p)
{
    double r = fmod(a, p);
    if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
    return r;
}

__global__ void __launch_bounds__(256)
k_resamp_carrier(const __grid_constant__ ResampConst C, const float2 *in, float2 *out) // out may alias in
{
    const int bpl = (C.inwidth + 255) / 256;
    const int row = blockIdx.x / bpl;
    const int i0 = (blockIdx.x - row * bpl) * 256 + threadIdx.x;
    if (i0 >= C.inwidth) return;
    const double r_at = (double)(row + 1), r_rt = (double)(i0 + 1);
    double r_ph = eval_poly2d(C.rg_carrier, r_at, r_rt) + eval_poly2d(C.az_carrier, r_at, r_rt);
    r_ph = f_modulo(r_ph, 2.0 * C.pi);
    double sn, cs;
    sincos(r_ph, &sn, &cs);
    const size_t o = (size_t)row * (size_t)C.inwidth + (size_t)i0;
    out[o] = cmul4(in[o], make_float2((float)cs, (float)(-sn)));
}

template <bool RESID_F32>
__global__ void __launch_bounds__(kResampBlock)
k_resamp_slc(const __grid_constant__ ResampConst C, const float2 *__restrict__ cin, const void *__restrict__ resid_az,
             const void *__restrict__ resid_rg, const float *__restrict__ sinc, float2 *__restrict__ out, ResampStats *stats)
{
    constexpr int sinchalf = kSincLen / 2;
    const int bpl = (C.outwidth + kResampBlock - 1) / kResampBlock;
    const int row = blockIdx.x / bpl;
    const int col = (blockIdx.x - row * bpl) * kResampBlock + threadIdx.x;
    unsigned int n_valid = 0;
    if (col < C.outwidth) {
        const int i = col + 1, j = row + 1; // the reference's 1-based pixel / line
        const size_t o = (size_t)row * (size_t)C.outwidth + (size_t)col;
        float2 res = make_float2(0.f, 0.f);
        double rrg = 0.0, raz = 0.0; // the 'read' DOUBLE caster of Resamp_slc.py:120-133 for float32 .off rasters
        if (resid_rg) rrg = RESID_F32 ? (double)reinterpret_cast<const float *>(resid_rg)[o] : reinterpret_cast<const double *>(resid_rg)[o];
        if (resid_az) raz = RESID_F32 ? (double)reinterpret_cast<const float *>(resid_az)[o] : reinterpret_cast<const double *>(resid_az)[o];
        double r_rt = (double)i, r_at = (double)j;
        const double r_ro = eval_poly2d(C.rg_off, r_at, r_rt) + rrg;
        const double r_ao = eval_poly2d(C.az_off, r_at, r_rt) + raz;
        const int k = (int)floor(i + r_ro);
        const double fracr = i + r_ro - k;
        const int kk = (int)floor(j + r_ao);
        const double fraca = j + r_ao - kk;
        const bool inside = !((k <= sinchalf) || (k >= (C.inwidth - sinchalf))) && !((kk <= sinchalf) || (kk >= (C.inlength - sinchalf)));
        if (inside) {
            n_valid = 1;
            const double r_dop = eval_poly2d(C.dop, r_at + r_ao, r_rt + r_ro); // :211
            // intp_sinc_cx(chip, 5, 5, fracr, fraca, 9, 9): taps chip(9-kq, 9-m), kq, m = 0..7, i.e. input column
            // k + 4 - kq and line kk + 4 - m, the line rotated by exp(-i (4 - m) r_dop); chip row / column 1 is never read
            int ifracx = (int)(fracr * kSincSub), ifracy = (int)(fraca * kSincSub);
            ifracx = ifracx < 0 ? 0 : (ifracx > kSincSub - 1 ? kSincSub - 1 : ifracx);
            ifracy = ifracy < 0 ? 0 : (ifracy > kSincSub - 1 ? kSincSub - 1 : ifracy);
            float wy[kSincLen];
            float2 rot[kSincLen];
#pragma unroll
            for (int m = 0; m < kSincLen; m++) {
                wy[m] = __ldg(sinc + m + ifracy * kSincLen);
                if (r_dop == 0.0) {
                    rot[m] = make_float2(1.0f, -0.0f); // cos(0), -sin(0)
                } else {
                    double sn, cs;
                    sincos(((9 - m) - 5.0) * r_dop, &sn, &cs);
                    rot[m] = make_float2((float)cs, (float)(-sn));
                }
            }
            float accr = 0.f, acci = 0.f;
            double fweightsum = 0.0;
            const float2 *base = cin + (size_t)(kk + 4 - 1) * (size_t)C.inwidth + (size_t)(k + 4 - 1); // (col k+4, line kk+4), 0-based
#pragma unroll 1
            for (int kq = 0; kq < kSincLen; kq++) {
                const float wx = __ldg(sinc + kq + ifracx * kSincLen);
#pragma unroll
                for (int m = 0; m < kSincLen; m++) {
                    const float2 c = cmul4(__ldg(base - (size_t)m * (size_t)C.inwidth - kq), rot[m]);
                    const double fweight = (double)(wx * wy[m]);
                    accr = (float)((double)accr + (double)c.x * fweight);
                    acci = (float)((double)acci + (double)c.y * fweight);
                    fweightsum = fweightsum + fweight;
                }
            }
            accr = (float)div_n((double)accr, fweightsum);
            acci = (float)div_n((double)acci, fweightsum);
            // phase to put back (:229-243)
            double r_ph = r_dop * fraca;
            r_rt = i + r_ro;
            r_at = j + r_ao;
            r_ph = r_ph + eval_poly2d(C.rg_carrier, r_at, r_rt) + eval_poly2d(C.az_carrier, r_at, r_rt);
            if (C.flatten != 0)
                r_ph = r_ph + (4.0 * C.pi / C.wvl) * ((C.r0 - C.refr0) + (i - 1.0) * (C.slr - C.refslr) + r_ro * C.slr) +
                       (4.0 * C.pi * (C.refr0 + (i - 1.0) * C.refslr)) * (1.0 / C.refwvl - 1.0 / C.wvl);
            r_ph = f_modulo(r_ph, 2.0 * C.pi);
            double sn, cs;
            sincos(r_ph, &sn, &cs);
            res = cmul4(make_float2(accr, acci), make_float2((float)cs, (float)sn));
        }
        out[o] = res;
    }
    for (int sft = 16; sft > 0; sft >>= 1) n_valid += __shfl_xor_sync(0xffffffffu, n_valid, sft);
    if ((threadIdx.x & 31) == 0 && n_valid) atomicAdd(&stats->valid, (unsigned long long)n_valid);
}

void launch_resamp_carrier(const ResampConst &C, const float2 *in, float2 *out, cudaStream_t s)
{
    const long long nblk = (long long)((C.inwidth + 255) / 256) * C.inlength;
    k_resamp_carrier<<<(unsigned)nblk, 256, 0, s>>>(C, in, out);
}

int launch_resamp_slc(const ResampConst &C, const float2 *cin, const void *resid_az, const void *resid_rg, int resid_f32,
                      const float *sinc, float2 *out, ResampStats *stats, cudaStream_t s)
{
    const long long nblk = (long long)((C.outwidth + kResampBlock - 1) / kResampBlock) * C.outlength;
    if (nblk > 0x7fffffffLL || nblk < 1) return -2;
    if (resid_f32) k_resamp_slc<true><<<(unsigned)nblk, kResampBlock, 0, s>>>(C, cin, resid_az, resid_rg, sinc, out, stats);
    else k_resamp_slc<false><<<(unsigned)nblk, kResampBlock, 0, s>>>(C, cin, resid_az, resid_rg, sinc, out, stats);
    return 0;
}

} // namespace b2
