// fast_host.cpp -- TEST INFRASTRUCTURE: runs the tolerance-mode formulas of the product
// (ocean_model_arch_b200/csrc/sw_fast.cuh, compiled here as plain C++) over whole arrays on the CPU, so
// that the re-associated arithmetic can be compared with the oracle without a GPU
// (tests/test_fast_formulas.py).  The device kernel (sw_kernels_march.cu) calls the same functions.
//
// Arrays are (ny, nx) row-major with the reference's global indexing of ONE block:
// interior = columns 2 .. nx-3, rows 2 .. ny-3 (0-based), i.e. nx_start = ny_start = 3.
#include <cstring>
#include <vector>

#include "../ocean_model_arch_b200/csrc/sw_fast.cuh"

using namespace swf;
using namespace swcu;

namespace {

// same entries as k_build_tables (sw_kernels_fused.cu) computes on the device from column nx_start
void base_table(int nx, int ny, const float *const m[9], std::vector<double> &tab)
{
    const int h = ny;
    tab.assign((size_t)T_COUNT * h, 0.0);
    for (int r = 0; r < h; ++r) {
        const size_t c = (size_t)r * nx + 2;
        const float dx = m[0][c], dy = m[1][c], dxt = m[2][c], dyt = m[3][c], dxh = m[4][c], dyh = m[5][c], dxb = m[6][c],
                    dyb = m[7][c], rlh = m[8][c];
        tab[T_DX * h + r] = dx; tab[T_DY * h + r] = dy; tab[T_DXT * h + r] = dxt; tab[T_DYT * h + r] = dyt;
        tab[T_DXH * h + r] = dxh; tab[T_DYH * h + r] = dyh; tab[T_DXB * h + r] = dxb; tab[T_DYB * h + r] = dyb;
        tab[T_RLH * h + r] = rlh;
        volatile float p;  // real(4) sub-expressions, rounded to real(4) before promotion
        p = dx * dy; tab[T_AREA * h + r] = p;
        p = dy * dy; tab[T_DY2 * h + r] = p;
        p = dx * dx; tab[T_DX2 * h + r] = p;
        p = dxb * dxb; tab[T_DXB2 * h + r] = p;
        p = dyb * dyb; tab[T_DYB2 * h + r] = p;
        p = dy / dx; tab[T_RYX * h + r] = p;
        p = dx / dy; tab[T_RXY * h + r] = p;
        p = dxb / dyb; tab[T_RXYB * h + r] = p;
        p = dyb / dxb; tab[T_RYXB * h + r] = p;
        tab[T_RDXT * h + r] = 1.0 / (double)dxt; tab[T_RDYT * h + r] = 1.0 / (double)dyt;
        tab[T_RDXH * h + r] = 1.0 / (double)dxh; tab[T_RDYH * h + r] = 1.0 / (double)dyh;
        tab[T_RDXB * h + r] = 1.0 / (double)dxb; tab[T_RDYB * h + r] = 1.0 / (double)dyb;
        tab[T_RAREA * h + r] = 1.0 / tab[T_AREA * h + r];
    }
}

struct Planes { std::vector<double> rhu, rhv, uh, vh, t, ss, zx, zy, fxp, fyp, fxpy, fypy; };

template <bool TRANS, bool LAT>
long steps(int nx, int ny, int nsteps, double tau, double ts, int ffs_i, const unsigned char *mk, const double *fc,
           double *ssh, double *sshp, double *u, double *up, double *v, double *vp, const double *h_r, const double *mu,
           const double *rhsx, const double *rhsy, const float *rdis, const double *ft, double *ff, double *ffp)
{
    const size_t N = (size_t)nx * ny;
    const double ffs = (double)ffs_i;
    Planes P;
    for (auto *p : {&P.rhu, &P.rhv, &P.uh, &P.vh, &P.t, &P.ss, &P.zx, &P.zy, &P.fxp, &P.fyp, &P.fxpy, &P.fypy}) p->assign(N, 0.0);
    std::vector<double> o[6];
    for (auto &x : o) x.resize(N);
    long bad = 0;
    for (int s = 0; s < nsteps; ++s) {
        // stage A on the interior grown by one cell
        for (int n = 1; n <= ny - 2; ++n) {
            const ACoef k = load_acoef(&fc[(size_t)n * FC_STRIDE]);
            for (int m = 1; m <= nx - 2; ++m) {
                const size_t c = (size_t)n * nx + m, e = c + 1, w = c - 1, no = c + nx, so = c - nx, en = no + 1;
                const int b_c = mk[c] & LU, b_e = mk[e] & LU, b_n = mk[no] & LU, b_en = mk[en] & LU;
                const double q_c = h_r[c] + ssh[c] * ffs;
                const double qm_c = b_c ? q_c : 0.0, qm_e = b_e ? h_r[e] + ssh[e] * ffs : 0.0;
                const double qm_n = b_n ? h_r[no] + ssh[no] * ffs : 0.0, qm_en = b_en ? h_r[en] + ssh[en] * ffs : 0.0;
                const double musum = (mu[c] + mu[no]) + (mu[e] + mu[en]);
                const AOut a = stage_a<TRANS, LAT>(k, mk[c], b_c + b_e, b_c + b_n, b_c + b_e + b_n + b_en, q_c, qm_c, qm_n,
                                                   qm_c + qm_e, qm_n + qm_en, u[c], u[no], v[c], v[e], up[c], up[w], up[no],
                                                   vp[c], vp[so], vp[e], mu[c], musum);
                P.rhu[c] = a.rhu; P.rhv[c] = a.rhv; P.uh[c] = a.uh; P.vh[c] = a.vh; P.t[c] = a.t; P.ss[c] = a.ss;
                P.zx[c] = a.zx; P.zy[c] = a.zy;
            }
        }
        // face fluxes where stage A is known at c, e and n
        for (int n = 1; n <= ny - 3; ++n)
            for (int m = 1; m <= nx - 3; ++m) {
                const size_t c = (size_t)n * nx + m, e = c + 1, no = c + nx;
                const Flux f = stage_flux((mk[c] & LUU) != 0, P.uh[c], P.uh[e], P.uh[no], P.vh[c], P.vh[e], P.vh[no],
                                          u[c], u[e], u[no], v[c], v[e], v[no]);
                P.fxp[c] = f.fxp; P.fyp[c] = f.fyp; P.fxpy[c] = f.fxpy; P.fypy[c] = f.fypy;
            }
        // stage B on the interior
        for (int k6 = 0; k6 < 6; ++k6) {
            const double *src[6] = {ssh, sshp, u, up, v, vp};
            std::memcpy(o[k6].data(), src[k6], N * sizeof(double));
        }
        for (int n = 2; n <= ny - 3; ++n) {
            const BCoef k = load_bcoef(&fc[(size_t)n * FC_STRIDE], tau);
            for (int m = 2; m <= nx - 3; ++m) {
                const size_t c = (size_t)n * nx + m, e = c + 1, w = c - 1, no = c + nx, so = c - nx;
                const int b_c = mk[c] & LU, b_e = mk[e] & LU, b_n = mk[no] & LU;
                const double qpm_c = b_c ? h_r[c] + sshp[c] * ffs : 0.0, qpm_e = b_e ? h_r[e] + sshp[e] * ffs : 0.0;
                const double qpm_n = b_n ? h_r[no] + sshp[no] * ffs : 0.0;
                Flux f;
                f.fxp = P.fxp[c]; f.fyp = P.fyp[c]; f.fxpy = P.fxpy[c]; f.fypy = P.fypy[c];
                const double rdx = rdis ? (double)(rdis[c] + rdis[e]) : 0.0, rdy = rdis ? (double)(rdis[c] + rdis[no]) : 0.0;
                const BOut b = stage_b<TRANS, LAT>(
                    k, mk[c], b_c + b_e, b_c + b_n, 0.5 * ts, ssh[c], ssh[e], ssh[no], sshp[c], qpm_c + qpm_e,
                    mad(qpm_c, k.area, qpm_n * k.area_n), u[c], up[c], v[c], vp[c], P.rhu[c], P.rhv[c], P.uh[c], P.uh[w],
                    P.vh[c], P.vh[so], P.t[c], P.t[e], P.t[no], P.ss[c], P.ss[so], P.ss[w], P.zx[c], P.zx[so], P.zy[c],
                    P.zy[w], f, P.fxp[w], P.fyp[so], P.fxpy[w], P.fypy[so], rhsx ? rhsx[c] : 0.0, rhsy ? rhsy[c] : 0.0, rdx,
                    rdy);
                o[0][c] = b.ssh; o[1][c] = b.sshp; o[2][c] = b.u; o[3][c] = b.up; o[4][c] = b.v; o[5][c] = b.vp;
                bad += b.bad;
            }
        }
        double *dst[6] = {ssh, sshp, u, up, v, vp};
        for (int k6 = 0; k6 < 6; ++k6) std::memcpy(dst[k6], o[k6].data(), N * sizeof(double));
        if (ff) {   // expl_tracer on the state just written (control/tracer.f90:44-61)
            std::vector<double> fx(N, 0.0), fy(N, 0.0), fn(ff, ff + N), fpn(ffp, ffp + N);
            auto qm = [&](size_t i) { return (mk[i] & LU) ? h_r[i] + ssh[i] * ffs : 0.0; };
            for (int n = 1; n <= ny - 3; ++n) {
                const TCoef k = load_tcoef(&ft[(size_t)n * FT_STRIDE]);
                for (int m = 1; m <= nx - 3; ++m) {
                    const size_t c = (size_t)n * nx + m, e = c + 1, no = c + nx;
                    const int b_c = mk[c] & LU, b_e = mk[e] & LU, b_n = mk[no] & LU;
                    const TFlux f = tracer_flux(k, mk[c], b_c + b_e, b_c + b_n, qm(c), qm(e), qm(no), u[c], v[c], mu[c], mu[e],
                                                mu[no], ff[c], ff[e], ff[no]);
                    fx[c] = f.fx; fy[c] = f.fy;
                }
            }
            for (int n = 2; n <= ny - 3; ++n) {
                const TCoef k = load_tcoef(&ft[(size_t)n * FT_STRIDE]);
                for (int m = 2; m <= nx - 3; ++m) {
                    const size_t c = (size_t)n * nx + m;
                    if (!(mk[c] & LU)) continue;
                    const TOut t = tracer_update(k, 0.5 * ts, h_r[c], h_r[c] + sshp[c] * ffs, ff[c], ffp[c], fx[c], fx[c - 1],
                                                 fy[c], fy[c - nx]);
                    fn[c] = t.ffn; fpn[c] = t.ffpf;
                }
            }
            std::memcpy(ff, fn.data(), N * sizeof(double));
            std::memcpy(ffp, fpn.data(), N * sizeof(double));
        }
    }
    return bad;
}

}  // namespace

extern "C" {

// metrics: dx dy dxt dyt dxh dyh dxb dyb rlh_s (real(4), (ny, nx)); mk: one byte of mask bits per cell.
// ff / ffp: one tracer field (NULL without tracers).
// Returns the number of K11 offenders (sea cells with |ssh| >= 1e4 or NaN), summed over the steps.
long swf_host_steps(int nx, int ny, int nsteps, double tau, double ts, int ffs, int trans, int lat,
                    const unsigned char *mk, const float *const *metrics,
                    double *ssh, double *sshp, double *u, double *up, double *v, double *vp,
                    const double *h_r, const double *mu, const double *rhsx, const double *rhsy, const float *rdis,
                    double *ff, double *ffp)
{
    std::vector<double> tab;
    base_table(nx, ny, metrics, tab);
    std::vector<double> fc((size_t)ny * FC_STRIDE);
    std::vector<double> ft((size_t)ny * FT_STRIDE);
    for (int r = 0; r < ny; ++r) {
        build_fast_row(tab.data(), ny, r, tau, &fc[(size_t)r * FC_STRIDE]);
        build_tracer_row(tab.data(), ny, r, tau, &ft[(size_t)r * FT_STRIDE]);
    }
#define GO(T, L) return steps<T, L>(nx, ny, nsteps, tau, ts, ffs, mk, fc.data(), ssh, sshp, u, up, v, vp, h_r, mu, rhsx, rhsy, rdis, ft.data(), ff, ffp)
    if (trans && lat) GO(true, true);
    if (trans) GO(true, false);
    if (lat) GO(false, true);
    GO(false, false);
#undef GO
}

}  // extern "C"
