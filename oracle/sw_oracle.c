/*
 * sw_oracle.c -- CPU ORACLE (test infrastructure only; see sw_oracle.h).
 *
 * PARITY UNPINNED: the reference has no golden vectors for this path and cannot be built here.
 * Plain-C restatement of the reference's Fortran kernels, grid construction, initial state,
 * block decomposition, halo copy and step driver.  Fortran semantics honoured:
 *   - a binary operation is evaluated in the wider kind of its two operands, equal-precedence
 *     operators left to right, so real(4)*real(4) products / quotients / sums stay real(4)
 *     (C with FLT_EVAL_METHOD==0 does the same for float*float);
 *   - successive divisions are true divisions (x/2.0d0/tau is two divisions);
 *   - masked stores: `if (mask>0.5)` cells that fail keep their previous contents;
 *   - no FMA contraction / reassociation: compile with -ffp-contract=off and no -ffast-math.
 * Arrays are column-major A(bnd_x1:bnd_x2, bnd_y1:bnd_y2) with m (x) contiguous, i.e.
 * element (m,n) at a[(n-bnd_y1)*(bnd_x2-bnd_x1+1) + (m-bnd_x1)], block arrays use GLOBAL indices
 * (core/decomposition.f90:493-503).
 */
#include "sw_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__FLT_EVAL_METHOD__) && (__FLT_EVAL_METHOD__ != 0)
#error "oracle needs FLT_EVAL_METHOD == 0 (float expressions evaluated in float)"
#endif

#define LD ((size_t)(bnd_x2 - bnd_x1 + 1))
#define IX(m, n) ((size_t)((n) - bnd_y1) * LD + (size_t)((m) - bnd_x1))
#define SEA(mask, m, n) ((mask)[IX(m, n)] > 0.5f)

/* shared/constants.f90:11-23 : real(4) parameters, dPi truncated literal */
static const float  Pi_r4 = 3.1415926f;
static const float  RadEarth = 6371000.0f;
static const float  EarthAngVel = 7.2921159e-5f;
static const float  FreeFallAcc = 9.8f;
static const double dPi = 3.14159265358979;
static const double lat_extr = 89.99999;

/* core/math_tools.f90:28-54 */
static double dcosd(double x) { return cos((x / 180.0) * dPi); }
static double dsind(double x) { return sin((x / 180.0) * dPi); }

/* ------------------------------------------------------------------------------------------ */
/* kernel/shallow_water/vel_ssh.f90:15-38 */
void swo_gaussian_elimination_kernel(SWO_DIMS, const float *lu, double *ssh, double sigma, int nx0, int ny0)
{
    for (int n = ny_start; n <= ny_end; ++n)
        for (int m = nx_start; m <= nx_end; ++m)
            if (SEA(lu, m, n)) {
                double dx = (double)(m - nx0) / (nx0 * 0.25);
                double dy = (double)(n - ny0) / (ny0 * 0.25);
                /* 1.0 is a real(4) literal promoted to 1.0d0; 2*dPi integer*real8 */
                ssh[IX(m, n)] = (1.0 / (sqrt(2 * dPi) * sigma)) * exp(-((dx * dx + dy * dy) / (2 * sigma * sigma)));
            }
}

/* kernel/shallow_water/vel_ssh.f90:40-67 ; returns the number of offending cells instead of aborting */
int swo_check_ssh_err_kernel(SWO_DIMS, const float *lu, const double *ssh)
{
    int bad = 0;
    for (int n = ny_start; n <= ny_end; ++n)
        for (int m = nx_start; m <= nx_end; ++m)
            if (SEA(lu, m, n)) {
                double s = ssh[IX(m, n)];
                if (s < 10000.0 && s > -10000.0) continue;
                ++bad;
            }
    return bad;
}

/* kernel/shallow_water/vel_ssh.f90:69-106 */
void swo_sw_update_ssh_kernel(SWO_DIMS, double tau, const float *lu, const float *dx, const float *dy,
                              const float *dxh, const float *dyh, const double *hhu, const double *hhv,
                              double *sshn, const double *sshp, const double *ubrtr, const double *vbrtr)
{
    for (int n = ny_start; n <= ny_end; ++n)
        for (int m = nx_start; m <= nx_end; ++m)
            if (SEA(lu, m, n)) {
                size_t c = IX(m, n), w = IX(m - 1, n), s = IX(m, n - 1);
                float area = dx[c] * dy[c]; /* real(4) product */
                double div = ubrtr[c] * hhu[c] * dyh[c] - ubrtr[w] * hhu[w] * dyh[w]
                           + vbrtr[c] * hhv[c] * dxh[c] - vbrtr[s] * hhv[s] * dxh[s];
                sshn[c] = sshp[c] + 2.0 * tau * (-(div / area));
            }
}

/* kernel/shallow_water/vel_ssh.f90:108-195 */
void swo_sw_update_uv(SWO_DIMS, double tau, const float *lcu, const float *lcv,
                      const float *dxt, const float *dyt, const float *dxh, const float *dyh,
                      const float *dxb, const float *dyb,
                      const double *hhu, const double *hhun, const double *hhup,
                      const double *hhv, const double *hhvn, const double *hhvp,
                      const double *hhh, const double *ssh,
                      const double *ubrtr, double *ubrtrn, const double *ubrtrp,
                      const double *vbrtr, double *vbrtrn, const double *vbrtrp,
                      const float *rdis, const float *rlh_s,
                      const double *RHSx, const double *RHSy, const double *RHSx_adv, const double *RHSy_adv,
                      const double *RHSx_dif, const double *RHSy_dif)
{
    const double g = (double)FreeFallAcc;
    for (int n = ny_start; n <= ny_end; ++n)
        for (int m = nx_start; m <= nx_end; ++m) {
            size_t c = IX(m, n), e = IX(m + 1, n), w = IX(m - 1, n), no = IX(m, n + 1), s = IX(m, n - 1);
            size_t es = IX(m + 1, n - 1), wn = IX(m - 1, n + 1);
            if (SEA(lcu, m, n)) {
                double bp = hhun[c] * dxt[c] * dyh[c] / 2.0 / tau;
                double bp0 = hhup[c] * dxt[c] * dyh[c] / 2.0 / tau;
                double slx = -g * (ssh[e] - ssh[c]) * dyh[c] * hhu[c];
                float rd = rdis[c] + rdis[e]; /* real(4) sum */
                double grx = RHSx[c] + slx + RHSx_dif[c] + RHSx_adv[c]
                           - rd / 2.0 * ubrtrp[c] * dxt[c] * dyh[c] * hhu[c]
                           + (rlh_s[c] * hhh[c] * dxb[c] * dyb[c] * (vbrtr[e] + vbrtr[c])
                            + rlh_s[s] * hhh[s] * dxb[s] * dyb[s] * (vbrtr[es] + vbrtr[s])) / 4.0;
                ubrtrn[c] = (ubrtrp[c] * bp0 + grx) / bp;
            }
            if (SEA(lcv, m, n)) {
                double bp = hhvn[c] * dyt[c] * dxh[c] / 2.0 / tau;
                double bp0 = hhvp[c] * dyt[c] * dxh[c] / 2.0 / tau;
                double sly = -g * (ssh[no] - ssh[c]) * dxh[c] * hhv[c];
                float rd = rdis[c] + rdis[no];
                double gry = RHSy[c] + sly + RHSy_dif[c] + RHSy_adv[c]
                           - rd / 2.0 * vbrtrp[c] * dxh[c] * dyt[c] * hhv[c]
                           - (rlh_s[c] * hhh[c] * dxb[c] * dyb[c] * (ubrtr[no] + ubrtr[c])
                            + rlh_s[w] * hhh[w] * dxb[w] * dyb[w] * (ubrtr[wn] + ubrtr[w])) / 4.0;
                vbrtrn[c] = (vbrtrp[c] * bp0 + gry) / bp;
            }
        }
}

/* kernel/shallow_water/vel_ssh.f90:197-245 (range grown by one: :226-227) */
void swo_sw_next_step(SWO_DIMS, double time_smooth, const float *lu, const float *lcu, const float *lcv,
                      double *ssh, double *sshn, double *sshp,
                      double *ubrtr, double *ubrtrn, double *ubrtrp,
                      double *vbrtr, double *vbrtrn, double *vbrtrp)
{
    for (int n = ny_start - 1; n <= ny_end + 1; ++n)
        for (int m = nx_start - 1; m <= nx_end + 1; ++m) {
            size_t c = IX(m, n);
            if (SEA(lu, m, n)) {
                sshp[c] = ssh[c] + time_smooth * (sshn[c] - 2.0 * ssh[c] + sshp[c]) / 2.0;
                ssh[c] = sshn[c];
            }
            if (SEA(lcu, m, n)) {
                ubrtrp[c] = ubrtr[c] + time_smooth * (ubrtrn[c] - 2.0 * ubrtr[c] + ubrtrp[c]) / 2.0;
                ubrtr[c] = ubrtrn[c];
            }
            if (SEA(lcv, m, n)) {
                vbrtrp[c] = vbrtr[c] + time_smooth * (vbrtrn[c] - 2.0 * vbrtr[c] + vbrtrp[c]) / 2.0;
                vbrtr[c] = vbrtrn[c];
            }
        }
}

/* kernel/shallow_water/vel_ssh.f90:247-281 (nlev = 1, interface/shallow_water/sw_interface.f90:216) */
void swo_uv_trans_vort_kernel(SWO_DIMS, const float *luu, const float *dxt, const float *dyt,
                              const float *dxb, const float *dyb, const double *u, const double *v, double *vort)
{
    for (int n = ny_start; n <= ny_end; ++n)
        for (int m = nx_start; m <= nx_end; ++m)
            if (SEA(luu, m, n)) {
                size_t c = IX(m, n), e = IX(m + 1, n), no = IX(m, n + 1);
                vort[c] = (v[e] * dyt[e] - v[c] * dyt[c])
                        - (u[no] * dxt[no] - u[c] * dxt[c])
                        - ((v[e] - v[c]) * dyb[c] - (u[no] - u[c]) * dxb[c]);
            }
}

/* kernel/shallow_water/vel_ssh.f90:283-373 (hq is passed but unused) */
void swo_uv_trans_kernel(SWO_DIMS, const float *lcu, const float *lcv, const float *luu,
                         const float *dxh, const float *dyh, const double *u, const double *v, const double *vort,
                         const double *hq, const double *hu, const double *hv, const double *hh,
                         double *RHSx, double *RHSy)
{
    (void)hq;
    for (int n = ny_start; n <= ny_end; ++n)
        for (int m = nx_start; m <= nx_end; ++m) {
            size_t c = IX(m, n), e = IX(m + 1, n), w = IX(m - 1, n), no = IX(m, n + 1), s = IX(m, n - 1);
            size_t es = IX(m + 1, n - 1), wn = IX(m - 1, n + 1);
            if (SEA(lcu, m, n)) {
                double fx_p = (u[c] * dyh[c] * hu[c] + u[e] * dyh[e] * hu[e]) / 2.0 * (u[c] + u[e]) / 2.0;
                double fx_m = (u[c] * dyh[c] * hu[c] + u[w] * dyh[w] * hu[w]) / 2.0 * (u[c] + u[w]) / 2.0;
                double fy_p = (v[c] * dxh[c] * hv[c] + v[e] * dxh[e] * hv[e]) / 2.0 * (u[no] + u[c]) / 2.0 * (double)luu[c];
                double fy_m = (v[s] * dxh[s] * hv[s] + v[es] * dxh[es] * hv[es]) / 2.0 * (u[s] + u[c]) / 2.0 * (double)luu[s];
                RHSx[c] = -(fx_p - fx_m + fy_p - fy_m)
                        + (vort[c] * hh[c] * (v[e] + v[c]) + vort[s] * hh[s] * (v[es] + v[s])) / 4.0;
            }
            if (SEA(lcv, m, n)) {
                double fy_p = (v[c] * dxh[c] * hv[c] + v[no] * dxh[no] * hv[no]) / 2.0 * (v[c] + v[no]) / 2.0;
                double fy_m = (v[c] * dxh[c] * hv[c] + v[s] * dxh[s] * hv[s]) / 2.0 * (v[c] + v[s]) / 2.0;
                double fx_p = (u[c] * dyh[c] * hu[c] + u[no] * dyh[no] * hu[no]) / 2.0 * (v[e] + v[c]) / 2.0;
                double fx_m = (u[w] * dyh[w] * hu[w] + u[wn] * dyh[wn] * hu[wn]) / 2.0 * (v[w] + v[c]) / 2.0;
                RHSy[c] = -(fx_p - fx_m + fy_p - fy_m)
                        - (vort[c] * hh[c] * (u[no] + u[c]) + vort[w] * hh[w] * (u[wn] + u[w])) / 4.0;
            }
        }
}

/* kernel/shallow_water/vel_ssh.f90:375-452 (hu, hv passed but unused; dy**2 etc. are real(4)) */
void swo_uv_diff2_kernel(SWO_DIMS, const float *lcu, const float *lcv,
                         const float *dx, const float *dy, const float *dxt, const float *dyt,
                         const float *dxh, const float *dyh, const float *dxb, const float *dyb,
                         const double *mu, const double *str_t, const double *str_s,
                         const double *hq, const double *hu, const double *hv, const double *hh,
                         double *RHSx, double *RHSy)
{
    (void)hu; (void)hv;
    for (int n = ny_start; n <= ny_end; ++n)
        for (int m = nx_start; m <= nx_end; ++m) {
            size_t c = IX(m, n), e = IX(m + 1, n), w = IX(m - 1, n), no = IX(m, n + 1), s = IX(m, n - 1);
            size_t en = IX(m + 1, n + 1), es = IX(m + 1, n - 1), wn = IX(m - 1, n + 1);
            if (SEA(lcu, m, n)) {
                double muh_p = (mu[c] + mu[e] + mu[no] + mu[en]) / 4.0;
                double muh_m = (mu[c] + mu[e] + mu[s] + mu[es]) / 4.0;
                float dy2e = dy[e] * dy[e], dy2c = dy[c] * dy[c];
                float dxb2c = dxb[c] * dxb[c], dxb2s = dxb[s] * dxb[s];
                RHSx[c] = (dy2e * mu[e] * hq[e] * str_t[e] - dy2c * mu[c] * hq[c] * str_t[c]) / dyh[c]
                        + (dxb2c * muh_p * hh[c] * str_s[c] - dxb2s * muh_m * hh[s] * str_s[s]) / dxt[c];
            }
            if (SEA(lcv, m, n)) {
                double muh_p = (mu[c] + mu[e] + mu[no] + mu[en]) / 4.0;
                double muh_m = (mu[c] + mu[w] + mu[no] + mu[wn]) / 4.0;
                float dx2n = dx[no] * dx[no], dx2c = dx[c] * dx[c];
                float dyb2c = dyb[c] * dyb[c], dyb2w = dyb[w] * dyb[w];
                RHSy[c] = -(dx2n * mu[no] * hq[no] * str_t[no] - dx2c * mu[c] * hq[c] * str_t[c]) / dxh[c]
                        + (dyb2c * muh_p * hh[c] * str_s[c] - dyb2w * muh_m * hh[w] * str_s[w]) / dyt[c];
            }
        }
}

/* kernel/shallow_water/mixing.f90:14-58 (dy/dx etc. are real(4) divisions) */
void swo_stress_components_kernel(SWO_DIMS, const float *lu, const float *luu,
                                  const float *dx, const float *dy, const float *dxt, const float *dyt,
                                  const float *dxh, const float *dyh, const float *dxb, const float *dyb,
                                  const double *u, const double *v, double *str_t, double *str_s)
{
    for (int n = ny_start; n <= ny_end; ++n)
        for (int m = nx_start; m <= nx_end; ++m) {
            size_t c = IX(m, n), e = IX(m + 1, n), w = IX(m - 1, n), no = IX(m, n + 1), s = IX(m, n - 1);
            if (SEA(lu, m, n)) {
                float ryx = dy[c] / dx[c], rxy = dx[c] / dy[c];
                str_t[c] = ryx * (u[c] / dyh[c] - u[w] / dyh[w]) - rxy * (v[c] / dxh[c] - v[s] / dxh[s]);
            }
            if (SEA(luu, m, n)) {
                float rxy = dxb[c] / dyb[c], ryx = dyb[c] / dxb[c];
                str_s[c] = rxy * (u[no] / dxt[no] - u[c] / dxt[c]) + ryx * (v[e] / dyt[e] - v[c] / dyt[c]);
            }
        }
}

/* the masked area-weighted interpolations shared by hh_init / hh_update (depth.f90:57-94, 136-157) */
static inline double interp_u(const double *hq, const float *dx, const float *dy, const float *lu,
                              size_t c, size_t e, float dxt, float dyh)
{
    double slu = (double)(lu[c] + lu[e]); /* real(4) sum, then dble() */
    return (hq[c] * dx[c] * dy[c] * (double)lu[c] + hq[e] * dx[e] * dy[e] * (double)lu[e]) / slu / dxt / dyh;
}
static inline double interp_h(const double *hq, const float *dx, const float *dy, const float *lu,
                              size_t c, size_t e, size_t no, size_t en, float dxb, float dyb)
{
    double slu = (double)(lu[c] + lu[e] + lu[no] + lu[en]);
    return (hq[c] * dx[c] * dy[c] * (double)lu[c] + hq[e] * dx[e] * dy[e] * (double)lu[e]
          + hq[no] * dx[no] * dy[no] * (double)lu[no] + hq[en] * dx[en] * dy[en] * (double)lu[en]) / slu / dxb / dyb;
}

/* kernel/shallow_water/depth.f90:14-99 */
void swo_hh_init_kernel(SWO_DIMS, int full_free_surface,
                        const float *lu, const float *llu, const float *llv, const float *luh,
                        const float *dx, const float *dy, const float *dxt, const float *dyt,
                        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
                        double *hq, double *hqp, double *hqn, double *hu, double *hup, double *hun,
                        double *hv, double *hvp, double *hvn, double *hh, double *hhp, double *hhn,
                        const double *sh, const double *shp, const double *h_r)
{
    const double ffs = (double)full_free_surface;
    size_t total = LD * (size_t)(bnd_y2 - bnd_y1 + 1);
    for (size_t i = 0; i < total; ++i) { /* whole-array statements, depth.f90:48-50 */
        hq[i] = h_r[i] + sh[i] * ffs;
        hqp[i] = h_r[i] + shp[i] * ffs;
        hqn[i] = h_r[i];
    }
    for (int n = ny_start - 1; n <= ny_end; ++n)
        for (int m = nx_start - 1; m <= nx_end; ++m) {
            size_t c = IX(m, n), e = IX(m + 1, n), no = IX(m, n + 1), en = IX(m + 1, n + 1);
            if (SEA(llu, m, n)) {
                hu[c] = interp_u(hq, dx, dy, lu, c, e, dxt[c], dyh[c]);
                hup[c] = interp_u(hqp, dx, dy, lu, c, e, dxt[c], dyh[c]);
                hun[c] = interp_u(hqn, dx, dy, lu, c, e, dxt[c], dyh[c]);
            }
            if (SEA(llv, m, n)) {
                hv[c] = interp_u(hq, dx, dy, lu, c, no, dxh[c], dyt[c]);
                hvp[c] = interp_u(hqp, dx, dy, lu, c, no, dxh[c], dyt[c]);
                hvn[c] = interp_u(hqn, dx, dy, lu, c, no, dxh[c], dyt[c]);
            }
            if (SEA(luh, m, n)) {
                hh[c] = interp_h(hq, dx, dy, lu, c, e, no, en, dxb[c], dyb[c]);
                hhp[c] = interp_h(hqp, dx, dy, lu, c, e, no, en, dxb[c], dyb[c]);
                hhn[c] = interp_h(hqn, dx, dy, lu, c, e, no, en, dxb[c], dyb[c]);
            }
        }
}

/* kernel/shallow_water/depth.f90:101-162 */
void swo_hh_update_kernel(SWO_DIMS, const float *lu, const float *llu, const float *llv, const float *luh,
                          const float *dx, const float *dy, const float *dxt, const float *dyt,
                          const float *dxh, const float *dyh, const float *dxb, const float *dyb,
                          double *hqn, double *hun, double *hvn, double *hhn, const double *sh, const double *h_r)
{
    size_t total = LD * (size_t)(bnd_y2 - bnd_y1 + 1);
    for (size_t i = 0; i < total; ++i) hqn[i] = h_r[i] + sh[i]; /* depth.f90:129 */
    for (int n = ny_start - 1; n <= ny_end; ++n)
        for (int m = nx_start - 1; m <= nx_end; ++m) {
            size_t c = IX(m, n), e = IX(m + 1, n), no = IX(m, n + 1), en = IX(m + 1, n + 1);
            if (SEA(llu, m, n)) hun[c] = interp_u(hqn, dx, dy, lu, c, e, dxt[c], dyh[c]);
            if (SEA(llv, m, n)) hvn[c] = interp_u(hqn, dx, dy, lu, c, no, dxh[c], dyt[c]);
            if (SEA(luh, m, n)) hhn[c] = interp_h(hqn, dx, dy, lu, c, e, no, en, dxb[c], dyb[c]);
        }
}

/* kernel/shallow_water/depth.f90:164-211 (time_smooth is a module variable there) */
void swo_hh_shift_kernel(SWO_DIMS, double time_smooth,
                         const float *lu, const float *llu, const float *llv, const float *luh,
                         double *hq, double *hqp, double *hqn, double *hu, double *hup, double *hun,
                         double *hv, double *hvp, double *hvn, double *hh, double *hhp, double *hhn)
{
    for (int n = ny_start - 1; n <= ny_end + 1; ++n)
        for (int m = nx_start - 1; m <= nx_end + 1; ++m) {
            size_t c = IX(m, n);
            if (SEA(llu, m, n)) {
                hup[c] = hu[c] + time_smooth * (hun[c] - 2.0 * hu[c] + hup[c]) / 2.0;
                hu[c] = hun[c];
            }
            if (SEA(llv, m, n)) {
                hvp[c] = hv[c] + time_smooth * (hvn[c] - 2.0 * hv[c] + hvp[c]) / 2.0;
                hv[c] = hvn[c];
            }
            if (SEA(lu, m, n)) {
                hqp[c] = hq[c] + time_smooth * (hqn[c] - 2.0 * hq[c] + hqp[c]) / 2.0;
                hq[c] = hqn[c];
            }
            if (SEA(luh, m, n)) {
                hhp[c] = hh[c] + time_smooth * (hhn[c] - 2.0 * hh[c] + hhp[c]) / 2.0;
                hh[c] = hhn[c];
            }
        }
}

/* kernel/tracer/leapfrog_tracer.f90:13-98 */
void swo_tran_diff_fluxes_kernel(SWO_DIMS, const float *lcu, const float *lcv,
                                 const float *dxt, const float *dyt, const float *dxh, const float *dyh,
                                 const double *hhu, const double *hhv, const double *ff, const double *ffp,
                                 const double *uu, const double *vv, const double *mu, double factor_mu,
                                 double *flux_x, double *flux_y)
{
    (void)ffp;
    for (int n = ny_start; n <= ny_end; ++n)
        for (int m = nx_start; m <= nx_end; ++m) {
            size_t c = IX(m, n), e = IX(m + 1, n), no = IX(m, n + 1);
            if (SEA(lcu, m, n)) {
                double dfdx = ff[e] - ff[c];
                double mu_1d = (mu[c] + mu[e]) / 2.0 * factor_mu * dyh[c] / dxt[c];
                double flux_diff = mu_1d * hhu[c] * dfdx;
                double flux_adv = -uu[c] * hhu[c] * dyh[c] * (ff[c] + ff[e]) / 2.0;
                flux_x[c] = flux_adv + flux_diff + 0.0;
            }
            if (SEA(lcv, m, n)) {
                double dfdy = ff[no] - ff[c];
                double mu_1d = (mu[c] + mu[no]) / 2.0 * factor_mu * dxh[c] / dyt[c];
                double flux_diff = mu_1d * hhv[c] * dfdy;
                double flux_adv = -vv[c] * hhv[c] * dxh[c] * (ff[c] + ff[no]) / 2.0;
                flux_y[c] = flux_adv + flux_diff + 0.0;
            }
        }
}

/* kernel/tracer/leapfrog_tracer.f90:100-141 */
void swo_tran_diff_tracer_kernel(SWO_DIMS, const float *lu, const float *dx, const float *dy, double tau,
                                 const double *hhqn, const double *hhqp, const double *flux_x, const double *flux_y,
                                 const double *ffp, double *ffn)
{
    for (int n = ny_start; n <= ny_end; ++n)
        for (int m = nx_start; m <= nx_end; ++m)
            if (SEA(lu, m, n)) {
                size_t c = IX(m, n), w = IX(m - 1, n), s = IX(m, n - 1);
                double bp = hhqn[c] * dx[c] * dy[c] / tau / 2.0;
                double bp0 = hhqp[c] * dx[c] * dy[c] / tau / 2.0;
                double rhs = flux_x[c] - flux_x[w] + flux_y[c] - flux_y[s];
                double eta = bp0 * ffp[c] + rhs;
                ffn[c] = eta / bp;
            }
}

/* kernel/tracer/leapfrog_tracer.f90:143-170 */
void swo_tracer_next_step_kernel(SWO_DIMS, double time_smooth, const float *lu,
                                 const double *ffn, double *ffp, double *ff)
{
    for (int n = ny_start - 1; n <= ny_end + 1; ++n)
        for (int m = nx_start - 1; m <= nx_end + 1; ++m)
            if (SEA(lu, m, n)) {
                size_t c = IX(m, n);
                ffp[c] = ff[c] + time_smooth * (ffn[c] - 2.0 * ff[c] + ffp[c]) / 2.0;
                ff[c] = ffn[c];
            }
}

/* kernel/service/grid_kernels.f90:18-38 ; mask is the GLOBAL (nx,ny) integer array */
void swo_lu_init_kernel(int bnd_x1, int bnd_x2, int bnd_y1, int bnd_y2, int nx, int ny,
                        const int *mask, float *lu, float *lu1)
{
    (void)ny;
    for (int n = bnd_y1; n <= bnd_y2; ++n)
        for (int m = bnd_x1; m <= bnd_x2; ++m) {
            if (mask[(size_t)(n - 1) * nx + (m - 1)] == 0) lu[IX(m, n)] = 1.0f;
            lu1[IX(m, n)] = 1.0f;
        }
}

/* kernel/service/grid_kernels.f90:40-92 */
void swo_lu_lv_init_kernel(int bnd_x1, int bnd_x2, int bnd_y1, int bnd_y2, const float *lu,
                           float *luh, float *luu, float *llu, float *llv, float *lcu, float *lcv)
{
    for (int n = bnd_y1; n <= bnd_y2 - 1; ++n)
        for (int m = bnd_x1; m <= bnd_x2 - 1; ++m) {
            size_t c = IX(m, n), e = IX(m + 1, n), no = IX(m, n + 1), en = IX(m + 1, n + 1);
            if (lu[c] + lu[e] + lu[no] + lu[en] > 0.5f) luh[c] = 1.0f;
            if (lu[c] * lu[e] * lu[no] * lu[en] > 0.5f) luu[c] = 1.0f;
            if (lu[c] + lu[e] > 0.5f) llu[c] = 1.0f;
            if (lu[c] + lu[no] > 0.5f) llv[c] = 1.0f;
            if (lu[c] * lu[e] > 0.5f) lcu[c] = 1.0f;
            if (lu[c] * lu[no] > 0.5f) lcv[c] = 1.0f;
        }
}

/* ========================================================================================== */
/*                                whole-model restatement                                     */
/* ========================================================================================== */

enum { /* real(8) fields: core/ocean.f90:14-48, core/grid.f90:36-53 */
    F_ssh, F_sshn, F_sshp, F_ubrtr, F_ubrtrn, F_ubrtrp, F_vbrtr, F_vbrtrn, F_vbrtrp,
    F_RHSx, F_RHSy, F_RHSx_adv, F_RHSy_adv, F_RHSx_dif, F_RHSy_dif, F_mu, F_str_t, F_str_s, F_vort,
    F_hhq_rest, F_hhq, F_hhq_p, F_hhq_n, F_hhu, F_hhu_p, F_hhu_n, F_hhv, F_hhv_p, F_hhv_n,
    F_hhh, F_hhh_p, F_hhh_n, F_flux_x, F_flux_y, F_ff1, F_ff1n, F_ff1p, NF8
};
static const char *const f8_names[NF8] = {
    "ssh", "sshn", "sshp", "ubrtr", "ubrtrn", "ubrtrp", "vbrtr", "vbrtrn", "vbrtrp",
    "RHSx", "RHSy", "RHSx_adv", "RHSy_adv", "RHSx_dif", "RHSy_dif", "mu", "str_t", "str_s", "vort",
    "hhq_rest", "hhq", "hhq_p", "hhq_n", "hhu", "hhu_p", "hhu_n", "hhv", "hhv_p", "hhv_n",
    "hhh", "hhh_p", "hhh_n", "flux_x", "flux_y", "ff1", "ff1n", "ff1p"};
enum { /* real(4) fields: core/grid.f90:24-31,55-66, core/ocean.f90:32 */
    G_lu, G_lu1, G_luu, G_luh, G_lcu, G_lcv, G_llu, G_llv,
    G_dx, G_dy, G_dxt, G_dyt, G_dxh, G_dyh, G_dxb, G_dyb, G_rlh_s, G_r_diss, NF4
};
static const char *const f4_names[NF4] = {
    "lu", "lu1", "luu", "luh", "lcu", "lcv", "llu", "llv",
    "dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb", "rlh_s", "r_diss"};

typedef struct {
    int nx_start, nx_end, ny_start, ny_end, bnd_x1, bnd_x2, bnd_y1, bnd_y2;
    int bm, bn;
    size_t size;
    double *f8[NF8];
    float *f4[NF4];
} swo_block;

struct swo_model {
    swo_config cfg;
    int *mask;
    int bnx, bny, bcount;
    swo_block *blk;
    double tau;
    long nstep;
};

#define BDIMS(b) (b)->nx_start, (b)->nx_end, (b)->ny_start, (b)->ny_end, (b)->bnd_x1, (b)->bnd_x2, (b)->bnd_y1, (b)->bnd_y2

/* core/decomposition.f90:427-503 (block_uniform_decomposition): sizes floor((N-done)/(blocks left)),
 * last block takes the remainder; interior starts at global index 3; arrays carry a +-2 border. */
static void uniform_split(int total, int nb, int *start, int *size)
{
    int done = 0;
    for (int i = 0; i < nb; ++i) {
        int s = (i == nb - 1) ? total - done : (int)floorf((float)(total - done) / (float)(nb - i));
        start[i] = done;
        size[i] = s;
        done += s;
    }
}

/* shared/mpp/syncborder_block2D_gen_all.fi:177-252 with core/decomposition.f90:94-154,230-290:
 * width-1 halo, 4 faces + 4 corners, copied from the neighbouring block's boundary cells. */
static void sync_block(swo_model *M, int k, int is8, int fid)
{
    static const int ddx[8] = {1, -1, 0, 0, 1, 1, -1, -1};
    static const int ddy[8] = {0, 0, 1, -1, 1, -1, 1, -1};
    swo_block *b = &M->blk[k];
    for (int d = 0; d < 8; ++d) {
        int bm = b->bm + ddx[d], bn = b->bn + ddy[d];
        if (bm < 0 || bm >= M->bnx || bn < 0 || bn >= M->bny) continue;
        swo_block *o = &M->blk[bn * M->bnx + bm];
        /* my halo cells in direction d (global indices) */
        int hx1 = ddx[d] > 0 ? b->nx_end + 1 : (ddx[d] < 0 ? b->nx_start - 1 : b->nx_start);
        int hx2 = ddx[d] > 0 ? b->nx_end + 1 : (ddx[d] < 0 ? b->nx_start - 1 : b->nx_end);
        int hy1 = ddy[d] > 0 ? b->ny_end + 1 : (ddy[d] < 0 ? b->ny_start - 1 : b->ny_start);
        int hy2 = ddy[d] > 0 ? b->ny_end + 1 : (ddy[d] < 0 ? b->ny_start - 1 : b->ny_end);
        size_t ldb = (size_t)(b->bnd_x2 - b->bnd_x1 + 1), ldo = (size_t)(o->bnd_x2 - o->bnd_x1 + 1);
        for (int n = hy1; n <= hy2; ++n)
            for (int m = hx1; m <= hx2; ++m) {
                size_t ib = (size_t)(n - b->bnd_y1) * ldb + (size_t)(m - b->bnd_x1);
                size_t io = (size_t)(n - o->bnd_y1) * ldo + (size_t)(m - o->bnd_x1);
                if (is8) b->f8[fid][ib] = o->f8[fid][io];
                else b->f4[fid][ib] = o->f4[fid][io];
            }
    }
}
static void sync8(swo_model *M, int fid)
{
    if (M->bcount == 1) return;
#pragma omp parallel for schedule(static, 1)
    for (int k = 0; k < M->bcount; ++k) sync_block(M, k, 1, fid);
}
static void sync4(swo_model *M, int fid)
{
    if (M->bcount == 1) return;
#pragma omp parallel for schedule(static, 1)
    for (int k = 0; k < M->bcount; ++k) sync_block(M, k, 0, fid);
}

/* kernel/service/grid_kernels.f90:94-204 + :206-538 (carthesian / spherical branches, uniform grid only)
 * with kernel/service/grid_parameters.f90:16-181.  Only the arrays the hot path reads are produced. */
static void grid_metrics_block(const swo_config *c, swo_block *b)
{
    /* grid_base_init_kernel, uniform grid (xgr_type = ygr_type = 0), grid_kernels.f90:164-201 */
    int bnd_x1 = b->bnd_x1, bnd_x2 = b->bnd_x2, bnd_y1 = b->bnd_y1;
    const float pip180 = Pi_r4 / 180.0f; /* shared/constants.f90:11-12 */
    float sx = (float)c->dxst * pip180 * RadEarth; /* sngl(dxst)*pip180*RadEarth, real(4) left to right */
    float sy = (float)c->dyst * pip180 * RadEarth;
    for (int n = b->ny_start - 1; n <= b->ny_end + 1; ++n)
        for (int m = b->nx_start - 1; m <= b->nx_end + 1; ++m) {
            size_t i = IX(m, n);
            b->f4[G_dxt][i] = sx; b->f4[G_dxb][i] = sx; b->f4[G_dx][i] = sx; b->f4[G_dxh][i] = sx;
            b->f4[G_dyt][i] = sy; b->f4[G_dyb][i] = sy; b->f4[G_dy][i] = sy; b->f4[G_dyh][i] = sy;
        }
    for (size_t i = 0; i < b->size; ++i) b->f4[G_rlh_s][i] = 2.0f * EarthAngVel; /* whole array, :201 */
    /* (the halo syncs of the eight metric arrays, grid_interface.f90:98-105, happen in the caller) */
}

static void grid_geo_block(const swo_config *c, swo_block *b)
{
    int bnd_x1 = b->bnd_x1, bnd_x2 = b->bnd_x2, bnd_y1 = b->bnd_y1, bnd_y2 = b->bnd_y2;
    int nxl = bnd_x2 - bnd_x1 + 1, nyl = bnd_y2 - bnd_y1 + 1;
    double *xt = calloc(nxl, sizeof(double)), *yt = calloc(nyl, sizeof(double));
    double *xu = calloc(nxl, sizeof(double)), *yv = calloc(nyl, sizeof(double));
    for (int m = bnd_x1; m <= bnd_x2; ++m) xt[m - bnd_x1] = c->rlon + (double)(m - 3) * c->dxst;
    for (int n = bnd_y1; n <= bnd_y2; ++n) yt[n - bnd_y1] = c->rlat + (double)(n - 3) * c->dyst;
    for (int m = bnd_x1; m <= bnd_x2 - 1; ++m) xu[m - bnd_x1] = (xt[m - bnd_x1] + xt[m + 1 - bnd_x1]) / 2.0;
    for (int n = bnd_y1; n <= bnd_y2 - 1; ++n) yv[n - bnd_y1] = (yt[n - bnd_y1] + yt[n + 1 - bnd_y1]) / 2.0;

    int m1 = b->nx_start - 1, m2 = b->nx_end + 1, n1 = b->ny_start - 1, n2 = b->ny_end + 1;
    if (c->curve_grid == 0) {
        /* grid_parameters_carthesian (grid_parameters.f90:16-78): metrics * 1.0; on the H-grid call
         * (key_cor = 1, grid_kernels.f90:311-329) cor_sin = cor_sin / sqrt(2.0) in real(4). */
        const float sq2 = sqrtf(2.0f);
        for (int n = n1; n <= n2; ++n)
            for (int m = m1; m <= m2; ++m) b->f4[G_rlh_s][IX(m, n)] = b->f4[G_rlh_s][IX(m, n)] / sq2;
    } else {
        /* grid_parameters_spherical (grid_parameters.f90:80-181), called for the T, U, V, H grids
         * (grid_kernels.f90:331-421): metr_x *= sngl(dcosd(lat_mod)); H-grid: cor_sin *= sngl(sin_lat). */
        const double sinlat_extr = dsind(lat_extr);
        struct { int fx; const double *ymod; const double *xmod; int key_cor; } g[4] = {
            {G_dx, yt, xt, 0}, {G_dxt, yt, xu, 0}, {G_dxh, yv, xt, 0}, {G_dxb, yv, xu, 1}};
        for (int q = 0; q < 4; ++q)
            for (int n = n1; n <= n2; ++n) {
                double y = g[q].ymod[n - bnd_y1];
                double lat_mod = fmax(fmin(y, lat_extr), -lat_extr);
                float coslat = (float)dcosd(lat_mod);
                for (int m = m1; m <= m2; ++m) {
                    size_t i = IX(m, n);
                    b->f4[g[q].fx][i] = b->f4[g[q].fx][i] * coslat;
                    if (g[q].key_cor) {
                        double x = g[q].xmod[m - bnd_x1];
                        double sin_lat = dsind(y) * dcosd(c->rotation_on_lat)
                                       + dcosd(x) * dcosd(y) * dsind(c->rotation_on_lat);
                        sin_lat = fmin(fmax(sin_lat, -sinlat_extr), sinlat_extr);
                        b->f4[G_rlh_s][i] = b->f4[G_rlh_s][i] * (float)sin_lat;
                    }
                }
            }
    }
    free(xt); free(yt); free(xu); free(yv);
}

/* one (kernel, sync) envoke: core/kernel_interface.f90:84-101 (_MPP_BLOCK_MODE_) */
#define FOR_BLOCKS(M, b) _Pragma("omp parallel for schedule(static, 1)") \
    for (int k__ = 0; k__ < (M)->bcount; ++k__) for (swo_block *b = &(M)->blk[k__]; b; b = NULL)

static void envoke_hh_init(swo_model *M)
{   /* interface/shallow_water/sw_interface.f90:42-90 */
    FOR_BLOCKS(M, b)
        swo_hh_init_kernel(BDIMS(b), M->cfg.full_free_surface,
                           b->f4[G_lu], b->f4[G_llu], b->f4[G_llv], b->f4[G_luh],
                           b->f4[G_dx], b->f4[G_dy], b->f4[G_dxt], b->f4[G_dyt],
                           b->f4[G_dxh], b->f4[G_dyh], b->f4[G_dxb], b->f4[G_dyb],
                           b->f8[F_hhq], b->f8[F_hhq_p], b->f8[F_hhq_n],
                           b->f8[F_hhu], b->f8[F_hhu_p], b->f8[F_hhu_n],
                           b->f8[F_hhv], b->f8[F_hhv_p], b->f8[F_hhv_n],
                           b->f8[F_hhh], b->f8[F_hhh_p], b->f8[F_hhh_n],
                           b->f8[F_ssh], b->f8[F_sshp], b->f8[F_hhq_rest]);
    sync8(M, F_hhu); sync8(M, F_hhv); sync8(M, F_hhh);
}

static void fill8(swo_model *M, int fid, double v)
{   /* core/data_types.f90:665-716: whole block array incl. frame */
    for (int k = 0; k < M->bcount; ++k)
        for (size_t i = 0; i < M->blk[k].size; ++i) M->blk[k].f8[fid][i] = v;
}
static void copy8(swo_model *M, int dst, int src)
{
    for (int k = 0; k < M->bcount; ++k) memcpy(M->blk[k].f8[dst], M->blk[k].f8[src], M->blk[k].size * sizeof(double));
}

swo_model *swo_create(const swo_config *cfg, const int *mask)
{
    swo_model *M = calloc(1, sizeof(*M));
    M->cfg = *cfg;
    swo_config *c = &M->cfg;
    if (c->bnx < 1) c->bnx = 1;
    if (c->bny < 1) c->bny = 1;
    if (c->hhq_rest == 0.0) c->hhq_rest = 100.0;
#ifdef _OPENMP
    if (c->nthreads > 0) omp_set_num_threads(c->nthreads);
#endif
    int nx = c->nx, ny = c->ny;
    M->tau = (double)c->time_step; /* tools/time_manager.f90:270: real(8) tau = real(4) time_step */
    M->mask = malloc((size_t)nx * ny * sizeof(int));
    if (mask) memcpy(M->mask, mask, (size_t)nx * ny * sizeof(int));
    else      /* tools/io.f90:49-59 */
        for (int n = 1; n <= ny; ++n)
            for (int m = 1; m <= nx; ++m)
                M->mask[(size_t)(n - 1) * nx + (m - 1)] = (m < 3 || m > nx - 2 || n < 3 || n > ny - 2) ? 1 : 0;

    /* decomposition */
    M->bnx = c->bnx; M->bny = c->bny; M->bcount = c->bnx * c->bny;
    M->blk = calloc(M->bcount, sizeof(swo_block));
    int *xs = malloc(sizeof(int) * M->bnx), *xn = malloc(sizeof(int) * M->bnx);
    int *ys = malloc(sizeof(int) * M->bny), *yn = malloc(sizeof(int) * M->bny);
    uniform_split(nx - 4, M->bnx, xs, xn);
    uniform_split(ny - 4, M->bny, ys, yn);
    for (int bn = 0; bn < M->bny; ++bn)
        for (int bm = 0; bm < M->bnx; ++bm) {
            swo_block *b = &M->blk[bn * M->bnx + bm];
            b->bm = bm; b->bn = bn;
            b->nx_start = 3 + xs[bm]; b->nx_end = b->nx_start + xn[bm] - 1;
            b->ny_start = 3 + ys[bn]; b->ny_end = b->ny_start + yn[bn] - 1;
            b->bnd_x1 = b->nx_start - 2; b->bnd_x2 = b->nx_end + 2;
            b->bnd_y1 = b->ny_start - 2; b->bnd_y2 = b->ny_end + 2;
            b->size = (size_t)(b->bnd_x2 - b->bnd_x1 + 1) * (size_t)(b->bnd_y2 - b->bnd_y1 + 1);
            for (int f = 0; f < NF8; ++f) {
                if (!c->use_tracers && f >= F_flux_x) { b->f8[f] = NULL; continue; }
                b->f8[f] = calloc(b->size, sizeof(double)); /* zero-initialised, core/data_types.f90:529 */
            }
            for (int f = 0; f < NF4; ++f) b->f4[f] = calloc(b->size, sizeof(float));
        }
    free(xs); free(xn); free(ys); free(yn);

    /* control/init_data.f90:96-125 init_grid_data: gridcon (service/gridcon.f90:19-40) */
    FOR_BLOCKS(M, b)
        swo_lu_init_kernel(b->bnd_x1, b->bnd_x2, b->bnd_y1, b->bnd_y2, nx, ny, M->mask, b->f4[G_lu], b->f4[G_lu1]);
    sync4(M, G_lu);
    FOR_BLOCKS(M, b)
        swo_lu_lv_init_kernel(b->bnd_x1, b->bnd_x2, b->bnd_y1, b->bnd_y2, b->f4[G_lu],
                              b->f4[G_luh], b->f4[G_luu], b->f4[G_llu], b->f4[G_llv], b->f4[G_lcu], b->f4[G_lcv]);
    sync4(M, G_luh); sync4(M, G_luu); sync4(M, G_lcu); sync4(M, G_llu); sync4(M, G_lcv); sync4(M, G_llv);
    /* basinpar (service/basinpar_construction.f90:71-72) */
    FOR_BLOCKS(M, b) grid_metrics_block(c, b);
    sync4(M, G_dxt); sync4(M, G_dxb); sync4(M, G_dx); sync4(M, G_dxh);
    sync4(M, G_dyt); sync4(M, G_dyb); sync4(M, G_dy); sync4(M, G_dyh);
    FOR_BLOCKS(M, b) grid_geo_block(c, b);
    fill8(M, F_hhq_rest, c->hhq_rest); /* init_data.f90:112-114 */

    /* control/init_data.f90:29-94 init_ocean_data */
    FOR_BLOCKS(M, b)
        swo_gaussian_elimination_kernel(BDIMS(b), b->f4[G_lu], b->f8[F_ssh], 1.0, nx / 2, ny / 2);
    sync8(M, F_ssh);
    copy8(M, F_sshn, F_ssh);
    copy8(M, F_sshp, F_ssh);
    envoke_hh_init(M);
    fill8(M, F_ubrtr, 0.0); fill8(M, F_ubrtrn, 0.0); fill8(M, F_ubrtrp, 0.0);
    fill8(M, F_vbrtr, 0.0); fill8(M, F_vbrtrn, 0.0); fill8(M, F_vbrtrp, 0.0);
    fill8(M, F_mu, c->lvisc_2);
    if (!c->keep_mu) fill8(M, F_mu, 0.0); /* init_data.f90:76-77 */
    if (c->r_diss != 0.0f)
        for (int k = 0; k < M->bcount; ++k)
            for (size_t i = 0; i < M->blk[k].size; ++i) M->blk[k].f4[G_r_diss][i] = c->r_diss;
    if (c->use_tracers > 0) { /* init_data.f90:80-90 ; one tracer field is carried (tracer_num = 1) */
        FOR_BLOCKS(M, b)
            swo_gaussian_elimination_kernel(BDIMS(b), b->f4[G_lu], b->f8[F_ff1], 0.5, nx / 2, ny / 2);
        sync8(M, F_ff1);
        copy8(M, F_ff1n, F_ff1);
        copy8(M, F_ff1p, F_ff1);
        fill8(M, F_flux_x, 0.0); fill8(M, F_flux_y, 0.0);
    }
    return M;
}

void swo_destroy(swo_model *M)
{
    if (!M) return;
    for (int k = 0; k < M->bcount; ++k) {
        for (int f = 0; f < NF8; ++f) free(M->blk[k].f8[f]);
        for (int f = 0; f < NF4; ++f) free(M->blk[k].f4[f]);
    }
    free(M->blk); free(M->mask); free(M);
}

/* control/shallow_water/shallow_water.f90:22-94 expl_shallow_water, then control/tracer.f90:33-62 */
long swo_step(swo_model *M, int nsteps)
{
    const swo_config *c = &M->cfg;
    const double tau = M->tau, ts = c->time_smooth;
    long bad_total = 0;
    for (int it = 0; it < nsteps; ++it) {
        /* K1 sw_update_ssh (sw_interface.f90:310-334) */
        FOR_BLOCKS(M, b)
            swo_sw_update_ssh_kernel(BDIMS(b), tau, b->f4[G_lu], b->f4[G_dx], b->f4[G_dy], b->f4[G_dxh], b->f4[G_dyh],
                                     b->f8[F_hhu], b->f8[F_hhv], b->f8[F_sshn], b->f8[F_sshp], b->f8[F_ubrtr], b->f8[F_vbrtr]);
        sync8(M, F_sshn);
        if (c->full_free_surface > 0) { /* K2 hh_update (:145-178), takes ssh */
            FOR_BLOCKS(M, b)
                swo_hh_update_kernel(BDIMS(b), b->f4[G_lu], b->f4[G_llu], b->f4[G_llv], b->f4[G_luh],
                                     b->f4[G_dx], b->f4[G_dy], b->f4[G_dxt], b->f4[G_dyt],
                                     b->f4[G_dxh], b->f4[G_dyh], b->f4[G_dxb], b->f4[G_dyb],
                                     b->f8[F_hhq_n], b->f8[F_hhu_n], b->f8[F_hhv_n], b->f8[F_hhh_n],
                                     b->f8[F_ssh], b->f8[F_hhq_rest]);
            sync8(M, F_hhu_n); sync8(M, F_hhv_n); sync8(M, F_hhh_n);
        }
        if (c->trans_terms > 0) { /* K3, K4 (:211-270) */
            FOR_BLOCKS(M, b)
                swo_uv_trans_vort_kernel(BDIMS(b), b->f4[G_luu], b->f4[G_dxt], b->f4[G_dyt], b->f4[G_dxb], b->f4[G_dyb],
                                         b->f8[F_ubrtr], b->f8[F_vbrtr], b->f8[F_vort]);
            sync8(M, F_vort);
            FOR_BLOCKS(M, b)
                swo_uv_trans_kernel(BDIMS(b), b->f4[G_lcu], b->f4[G_lcv], b->f4[G_luu], b->f4[G_dxh], b->f4[G_dyh],
                                    b->f8[F_ubrtr], b->f8[F_vbrtr], b->f8[F_vort],
                                    b->f8[F_hhq], b->f8[F_hhu], b->f8[F_hhv], b->f8[F_hhh],
                                    b->f8[F_RHSx_adv], b->f8[F_RHSy_adv]);
            sync8(M, F_hhu_p); sync8(M, F_hhv_p); sync8(M, F_hhh_p); /* "lazy" syncs */
        }
        if (c->ksw_lat > 0) { /* K5, K6 (:110-142, :273-307); K5 takes ubrtrp, vbrtrp */
            FOR_BLOCKS(M, b)
                swo_stress_components_kernel(BDIMS(b), b->f4[G_lu], b->f4[G_luu],
                                             b->f4[G_dx], b->f4[G_dy], b->f4[G_dxt], b->f4[G_dyt],
                                             b->f4[G_dxh], b->f4[G_dyh], b->f4[G_dxb], b->f4[G_dyb],
                                             b->f8[F_ubrtrp], b->f8[F_vbrtrp], b->f8[F_str_t], b->f8[F_str_s]);
            sync8(M, F_str_t); sync8(M, F_str_s);
            FOR_BLOCKS(M, b)
                swo_uv_diff2_kernel(BDIMS(b), b->f4[G_lcu], b->f4[G_lcv],
                                    b->f4[G_dx], b->f4[G_dy], b->f4[G_dxt], b->f4[G_dyt],
                                    b->f4[G_dxh], b->f4[G_dyh], b->f4[G_dxb], b->f4[G_dyb],
                                    b->f8[F_mu], b->f8[F_str_t], b->f8[F_str_s],
                                    b->f8[F_hhq], b->f8[F_hhu], b->f8[F_hhv], b->f8[F_hhh],
                                    b->f8[F_RHSx_dif], b->f8[F_RHSy_dif]);
        }
        /* K7 sw_update_uv (:337-381) */
        FOR_BLOCKS(M, b)
            swo_sw_update_uv(BDIMS(b), tau, b->f4[G_lcu], b->f4[G_lcv],
                             b->f4[G_dxt], b->f4[G_dyt], b->f4[G_dxh], b->f4[G_dyh], b->f4[G_dxb], b->f4[G_dyb],
                             b->f8[F_hhu], b->f8[F_hhu_n], b->f8[F_hhu_p],
                             b->f8[F_hhv], b->f8[F_hhv_n], b->f8[F_hhv_p],
                             b->f8[F_hhh], b->f8[F_ssh],
                             b->f8[F_ubrtr], b->f8[F_ubrtrn], b->f8[F_ubrtrp],
                             b->f8[F_vbrtr], b->f8[F_vbrtrn], b->f8[F_vbrtrp],
                             b->f4[G_r_diss], b->f4[G_rlh_s],
                             b->f8[F_RHSx], b->f8[F_RHSy], b->f8[F_RHSx_adv], b->f8[F_RHSy_adv],
                             b->f8[F_RHSx_dif], b->f8[F_RHSy_dif]);
        sync8(M, F_vbrtrn); sync8(M, F_ubrtrn);
        /* K8 sw_next_step (:384-408) */
        FOR_BLOCKS(M, b)
            swo_sw_next_step(BDIMS(b), ts, b->f4[G_lu], b->f4[G_lcu], b->f4[G_lcv],
                             b->f8[F_ssh], b->f8[F_sshn], b->f8[F_sshp],
                             b->f8[F_ubrtr], b->f8[F_ubrtrn], b->f8[F_ubrtrp],
                             b->f8[F_vbrtr], b->f8[F_vbrtrn], b->f8[F_vbrtrp]);
        if (c->full_free_surface > 0) { /* K9 hh_shift (:181-208), K10 hh_init (:42-90) */
            FOR_BLOCKS(M, b)
                swo_hh_shift_kernel(BDIMS(b), ts, b->f4[G_lu], b->f4[G_llu], b->f4[G_llv], b->f4[G_luh],
                                    b->f8[F_hhq], b->f8[F_hhq_p], b->f8[F_hhq_n],
                                    b->f8[F_hhu], b->f8[F_hhu_p], b->f8[F_hhu_n],
                                    b->f8[F_hhv], b->f8[F_hhv_p], b->f8[F_hhv_n],
                                    b->f8[F_hhh], b->f8[F_hhh_p], b->f8[F_hhh_n]);
            envoke_hh_init(M);
        }
        /* K11 check_ssh_err (:93-107) */
        long bad = 0;
#pragma omp parallel for schedule(static, 1) reduction(+ : bad)
        for (int k = 0; k < M->bcount; ++k) {
            swo_block *b = &M->blk[k];
            bad += swo_check_ssh_err_kernel(BDIMS(b), b->f4[G_lu], b->f8[F_ssh]);
        }
        bad_total += bad;

        if (c->use_tracers > 0) { /* control/tracer.f90:44-61, interface/tracer/tracer_interface.f90:28-102 */
            FOR_BLOCKS(M, b)
                swo_tran_diff_fluxes_kernel(BDIMS(b), b->f4[G_lcu], b->f4[G_lcv],
                                            b->f4[G_dxt], b->f4[G_dyt], b->f4[G_dxh], b->f4[G_dyh],
                                            b->f8[F_hhu], b->f8[F_hhv], b->f8[F_ff1], b->f8[F_ff1p],
                                            b->f8[F_ubrtr], b->f8[F_vbrtr], b->f8[F_mu], 1.0,
                                            b->f8[F_flux_x], b->f8[F_flux_y]);
            sync8(M, F_flux_x); sync8(M, F_flux_y);
            FOR_BLOCKS(M, b)
                swo_tran_diff_tracer_kernel(BDIMS(b), b->f4[G_lu], b->f4[G_dx], b->f4[G_dy], tau,
                                            b->f8[F_hhq_n], b->f8[F_hhq_p], b->f8[F_flux_x], b->f8[F_flux_y],
                                            b->f8[F_ff1p], b->f8[F_ff1n]);
            sync8(M, F_ff1n);
            FOR_BLOCKS(M, b)
                swo_tracer_next_step_kernel(BDIMS(b), ts, b->f4[G_lu], b->f8[F_ff1n], b->f8[F_ff1p], b->f8[F_ff1]);
        }
        M->nstep++;
    }
    return bad_total;
}

static int find_name(const char *name, int *is8)
{
    for (int f = 0; f < NF8; ++f) if (!strcmp(name, f8_names[f])) { *is8 = 1; return f; }
    for (int f = 0; f < NF4; ++f) if (!strcmp(name, f4_names[f])) { *is8 = 0; return f; }
    return -1;
}

/* cell (m,n) of the global array is owned by the block whose [start-ext .. end+ext] range holds it,
 * ext = 2 on a side that touches the global frame, 0 otherwise */
static int owns(const swo_model *M, const swo_block *b, int m, int n)
{
    int x1 = b->bm == 0 ? b->bnd_x1 : b->nx_start, x2 = b->bm == M->bnx - 1 ? b->bnd_x2 : b->nx_end;
    int y1 = b->bn == 0 ? b->bnd_y1 : b->ny_start, y2 = b->bn == M->bny - 1 ? b->bnd_y2 : b->ny_end;
    return m >= x1 && m <= x2 && n >= y1 && n <= y2;
}

int swo_get_field(const swo_model *M, const char *name, double *out8, float *out4)
{
    int is8, f = find_name(name, &is8);
    if (f < 0) return 1;
    if ((is8 && !out8) || (!is8 && !out4)) return 2;
    int nx = M->cfg.nx;
    for (int k = 0; k < M->bcount; ++k) {
        const swo_block *b = &M->blk[k];
        if (is8 && !b->f8[f]) return 1;
        size_t ld = (size_t)(b->bnd_x2 - b->bnd_x1 + 1);
        for (int n = b->bnd_y1; n <= b->bnd_y2; ++n)
            for (int m = b->bnd_x1; m <= b->bnd_x2; ++m) {
                if (!owns(M, b, m, n)) continue;
                size_t ib = (size_t)(n - b->bnd_y1) * ld + (size_t)(m - b->bnd_x1);
                size_t ig = (size_t)(n - 1) * nx + (size_t)(m - 1);
                if (is8) out8[ig] = b->f8[f][ib]; else out4[ig] = b->f4[f][ib];
            }
    }
    return 0;
}

int swo_set_field(swo_model *M, const char *name, const double *in8, const float *in4)
{
    int is8, f = find_name(name, &is8);
    if (f < 0) return 1;
    if ((is8 && !in8) || (!is8 && !in4)) return 2;
    int nx = M->cfg.nx;
    for (int k = 0; k < M->bcount; ++k) {
        swo_block *b = &M->blk[k];
        if (is8 && !b->f8[f]) return 1;
        size_t ld = (size_t)(b->bnd_x2 - b->bnd_x1 + 1);
        for (int n = b->bnd_y1; n <= b->bnd_y2; ++n)
            for (int m = b->bnd_x1; m <= b->bnd_x2; ++m) {
                size_t ib = (size_t)(n - b->bnd_y1) * ld + (size_t)(m - b->bnd_x1);
                size_t ig = (size_t)(n - 1) * nx + (size_t)(m - 1);
                if (is8) b->f8[f][ib] = in8[ig]; else b->f4[f][ib] = in4[ig];
            }
    }
    return 0;
}

int swo_block_count(const swo_model *M) { return M->bcount; }
int swo_block_dims(const swo_model *M, int k, int d[8])
{
    if (k < 0 || k >= M->bcount) return 1;
    const swo_block *b = &M->blk[k];
    d[0] = b->nx_start; d[1] = b->nx_end; d[2] = b->ny_start; d[3] = b->ny_end;
    d[4] = b->bnd_x1; d[5] = b->bnd_x2; d[6] = b->bnd_y1; d[7] = b->bnd_y2;
    return 0;
}
void *swo_block_field(swo_model *M, int k, const char *name, int *is_real8)
{
    int is8, f = find_name(name, &is8);
    if (f < 0 || k < 0 || k >= M->bcount) return NULL;
    if (is_real8) *is_real8 = is8;
    return is8 ? (void *)M->blk[k].f8[f] : (void *)M->blk[k].f4[f];
}
