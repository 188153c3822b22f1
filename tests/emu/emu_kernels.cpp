// emu_kernels.cpp -- DEVELOPMENT AID (tests only): compiles the per-pixel device functions of
// isce2_b200/csrc/*.cuh for the host with g++ and loops over pixels on the CPU, so that the kernel
// arithmetic can be compared bit-for-bit with the oracle in a container without a GPU.
// It is never loaded by the product; the GPU parity tests (-m gpu) are the real check.
#include <cstdint>
#include <cstring>

#include "../../isce2_b200/csrc/topo_pixel.cuh"

using namespace b2;

extern "C" {

struct EmuTopoArgs {
    double a, e2, wvl, thresh;
    int ilrl, numiter, extraiter;
    double ufirstlat, ufirstlon, deltalat, deltalon;
    int nx, ny, method, width, length, nazlooks;
    double t0, prf, peghdg;
    int orbit_method;
    int n_orbit;
    int dop_range_order, dop_azimuth_order;
    double r0, dr;
    int line0, nlines, want_inc, use_ref;
};

int emu_topo(const EmuTopoArgs *A, const float *dem, const double *ot, const double *opos, const double *ovel,
             const double *dop_coeffs, double *lat, double *lon, double *hgt, float *los, float *inc, double *ctrack,
             float *elev, long long *iters)
{
    TopoConst C;
    memset(&C, 0, sizeof C);
    C.elp = make_ellipsoid(A->a, A->e2);
    C.wvl = A->wvl; C.thresh = A->thresh; C.ilrl = A->ilrl; C.numiter = A->numiter; C.extraiter = A->extraiter;
    C.ufirstlat = A->ufirstlat; C.ufirstlon = A->ufirstlon; C.deltalat = A->deltalat; C.deltalon = A->deltalon;
    static float *sinc_tab = nullptr;
    if (A->method == 0 && !sinc_tab) { sinc_tab = new float[(size_t)kSincSub * kSincLen]; sinc_make_table(sinc_tab); }
    C.dem = DemView{dem, A->nx, A->ny, sinc_tab};
    C.method = A->method; C.width = A->width; C.length = A->length; C.nazlooks = A->nazlooks;
    C.t0 = A->t0; C.prf = A->prf; C.peghdg = A->peghdg;
    C.pi = 4.0 * atan(1.0); C.r2d = 180.0 / C.pi; C.orbit_method = A->orbit_method;
    C.inv_r2d = 1.0 / C.r2d; C.inv_dlat = 1.0 / C.deltalat; C.inv_dlon = 1.0 / C.deltalon;
    C.dop.range_order = A->dop_range_order; C.dop.azimuth_order = A->dop_azimuth_order;
    C.dop.norm_range = C.dop.norm_azimuth = C.dop.inv_norm_range = C.dop.inv_norm_azimuth = 1.0;
    memcpy(C.dop.c, dop_coeffs, sizeof(double) * (A->dop_range_order + 1) * (A->dop_azimuth_order + 1));
    C.slr.range_order = 1; C.slr.azimuth_order = 0; C.slr.norm_range = C.slr.norm_azimuth = C.slr.inv_norm_range = C.slr.inv_norm_azimuth = 1.0;
    C.slr.c[0] = A->r0; C.slr.c[1] = A->dr;
    spline6_make_table(C.spl);
    {
        const double d2r = C.pi / 180.0;
        C.ref.lat = make_ref_angle((C.ufirstlat + 0.5 * C.deltalat * A->ny) * d2r);
        C.ref.lon = make_ref_angle((C.ufirstlon + 0.5 * C.deltalon * A->nx) * d2r);
        C.ref.use_ref = A->use_ref;
        C.ref.lat.narrow = C.ref.lon.narrow = (A->use_ref == 2) ? 1 : 0; // 2: truncated series of narrow blocks
    }
    OrbitView orb{A->n_orbit, ot, opos, ovel};
    long long it = 0;
    for (int row = 0; row < A->nlines; row++) {
        int line = A->line0 + row;
        double tline = C.t0 + C.nazlooks * ((double)(line + 1) - 1.0) / C.prf;
        Vec3 p = v3(0, 0, 0), v = v3(0, 0, 0);
        orbit_interp(C.orbit_method, orb, tline, p, v);
        LineState L;
        make_line_state(C.elp, p, v, C.peghdg, L);
        for (int pix = 0; pix < A->width; pix++) {
            double rng = eval_poly2d(C.slr, (double)line, (double)pix);
            double dop = eval_poly2d(C.dop, (double)line, (double)pix);
            PixelResult R;
            const bool winc = A->want_inc != 0;
            if (A->use_ref) {
                switch (A->method) {
                case 0: topo_pixel<0, true>(C, L, rng, dop, winc, R); break;
                case 1: topo_pixel<1, true>(C, L, rng, dop, winc, R); break;
                case 2: topo_pixel<2, true>(C, L, rng, dop, winc, R); break;
                case 3: topo_pixel<3, true>(C, L, rng, dop, winc, R); break;
                case 4: topo_pixel<4, true>(C, L, rng, dop, winc, R); break;
                default: topo_pixel<5, true>(C, L, rng, dop, winc, R); break;
                }
            } else {
                switch (A->method) {
                case 0: topo_pixel<0, false>(C, L, rng, dop, winc, R); break;
                case 1: topo_pixel<1, false>(C, L, rng, dop, winc, R); break;
                case 2: topo_pixel<2, false>(C, L, rng, dop, winc, R); break;
                case 3: topo_pixel<3, false>(C, L, rng, dop, winc, R); break;
                case 4: topo_pixel<4, false>(C, L, rng, dop, winc, R); break;
                default: topo_pixel<5, false>(C, L, rng, dop, winc, R); break;
                }
            }
            size_t w = A->width, o = (size_t)row * w + pix;
            lat[o] = R.lat; lon[o] = R.lon; hgt[o] = R.hgt;
            los[(size_t)row * 2 * w + pix] = R.los0; los[(size_t)row * 2 * w + w + pix] = R.los1;
            inc[(size_t)row * 2 * w + pix] = R.inc0; inc[(size_t)row * 2 * w + w + pix] = R.inc1;
            ctrack[o] = R.ctrack; elev[o] = R.elev;
            it += R.iters;
        }
    }
    *iters = it;
    return 0;
}
}

// ---- orbit polynomials (host construction in orbit_poly.h, Horner evaluation as in k_geo2rdr_poly) ----
#include "../../isce2_b200/csrc/orbit_poly.h"
extern "C" int emu_orbit_poly(int method, int n, const double *t, const double *pos, const double *vel, int nq, const double *tq,
                              double *out /*[nq][9]: pos, vel, acc*/)
{
    HostOrbitPoly hp;
    if (!build_orbit_poly(method, n, t, pos, vel, hp)) return -1;
    const int back = method == 0 ? 2 : 5, span = method == 0 ? 4 : 9, NC = hp.ncoef;
    for (int q = 0; q < nq; q++) {
        int i = 0;
        while (i < n && t[i] < tq[q]) i++;
        int w = i - back;
        if (w < 0) w = 0;
        if (w > n - span) w = n - span;
        const double ih = hp.inv_h[w], s = (tq[q] - hp.tc[w]) * ih;
        for (int c = 0; c < 3; c++) {
            const double *cp = &hp.cp[((size_t)w * 3 + c) * NC];
            if (method == 0) {
                double p = cp[0], dp = 0.0, ddp = 0.0;
                for (int k = 1; k < NC; k++) { ddp = fma(ddp, s, dp); dp = fma(dp, s, p); p = fma(p, s, cp[k]); }
                out[9 * q + c] = p; out[9 * q + 3 + c] = dp * ih; out[9 * q + 6 + c] = 2.0 * ddp * ih * ih;
            } else {
                const double *cv = &hp.cv[((size_t)w * 3 + c) * NC];
                double p = cp[0], v = cv[0], dv = 0.0;
                for (int k = 1; k < NC; k++) { dv = fma(dv, s, v); v = fma(v, s, cv[k]); p = fma(p, s, cp[k]); }
                out[9 * q + c] = p; out[9 * q + 3 + c] = v; out[9 * q + 6 + c] = dv * ih;
            }
        }
    }
    return 0;
}
