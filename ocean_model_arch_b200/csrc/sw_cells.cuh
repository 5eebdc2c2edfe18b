// sw_cells.cuh -- the two per-cell stages of the fused step, written once and instantiated over
// global memory (k_prep / k_update) and over shared-memory tiles (k_step).  All arrays of one
// `Cell*` view share one index space: element (m,n) at [c], neighbours at c+-1 and c+-p.
#pragma once
#include "sw_formulas.cuh"

namespace swcu {

// bit <=> the reference's real(4) mask of the same name is > 0.5 (core/grid.f90:24-31)
enum : int { MB_LU = 1, MB_LCU = 2, MB_LCV = 4, MB_LUU = 8, MB_LUH = 16, MB_LLU = 32, MB_LLV = 64 };

__device__ __forceinline__ double md(unsigned char bits, int bit) { return (bits & bit) ? 1.0 : 0.0; }

struct PrepOut { double hu, hv, hh, vort, str_t, str_s; };

// Stage A for cell c: K10/K2 depths on U/V/H points (depth.f90:48,57-94), K3 vorticity
// (vel_ssh.f90:273-275), K5 stresses (mixing.f90:43-51).  Masked-out results are 0, the value the
// reference's zero-initialised arrays keep where the kernels never store.
template <bool TRANS, bool LAT, class M>
__device__ __forceinline__ PrepOut prep_cell(long c, int r, int p, const M &mt, double ffs,
        const unsigned char *__restrict__ mask,
        const double *__restrict__ ssh, const double *__restrict__ h_r,
        const double *__restrict__ u, const double *__restrict__ v,
        const double *__restrict__ up, const double *__restrict__ vp)
{
    const long e = c + 1, no = c + p, en = c + 1 + p;
    const unsigned char mb = mask[c];
    const double q_c = h_r[c] + ssh[c] * ffs, q_e = h_r[e] + ssh[e] * ffs;
    const double q_n = h_r[no] + ssh[no] * ffs, q_en = h_r[en] + ssh[en] * ffs;
    const int b_c = mb & MB_LU, b_e = mask[e] & MB_LU, b_n = mask[no] & MB_LU, b_en = mask[en] & MB_LU;  // 0 / 1
    const double lu_c = b_c ? 1.0 : 0.0, lu_e = b_e ? 1.0 : 0.0, lu_n = b_n ? 1.0 : 0.0, lu_en = b_en ? 1.0 : 0.0;
    const double dx_c = mt.dx(c, r), dy_c = mt.dy(c, r), dx_e = mt.dx(e, r), dy_e = mt.dy(e, r);
    const double dx_n = mt.dx(no, r + 1), dy_n = mt.dy(no, r + 1), dx_en = mt.dx(en, r + 1), dy_en = mt.dy(en, r + 1);
    PrepOut o;
    const double hu = f_interp2_b<M>(q_c, q_e, dx_c, dy_c, lu_c, dx_e, dy_e, lu_e, b_c + b_e,
                                     mt.dxt(c, r), mt.r_dxt(c, r), mt.dyh(c, r), mt.r_dyh(c, r));
    const double hv = f_interp2_b<M>(q_c, q_n, dx_c, dy_c, lu_c, dx_n, dy_n, lu_n, b_c + b_n,
                                     mt.dxh(c, r), mt.r_dxh(c, r), mt.dyt(c, r), mt.r_dyt(c, r));
    const double hh = f_interp4_b<M>(q_c, q_e, q_n, q_en, dx_c, dy_c, lu_c, dx_e, dy_e, lu_e, dx_n, dy_n, lu_n,
                                     dx_en, dy_en, lu_en, b_c + b_e + b_n + b_en,
                                     mt.dxb(c, r), mt.r_dxb(c, r), mt.dyb(c, r), mt.r_dyb(c, r));
    o.hu = (mb & MB_LLU) ? hu : 0.0;
    o.hv = (mb & MB_LLV) ? hv : 0.0;
    o.hh = (mb & MB_LUH) ? hh : 0.0;
    o.vort = 0.0; o.str_t = 0.0; o.str_s = 0.0;
    if (TRANS) {
        const double vo = f_vort(c, r, p, mt, u, v);
        o.vort = (mb & MB_LUU) ? vo : 0.0;
    }
    if (LAT) {
        const double st = f_str_t(c, r, p, mt, up, vp), ss = f_str_s(c, r, p, mt, up, vp);
        o.str_t = (mb & MB_LU) ? st : 0.0;
        o.str_s = (mb & MB_LUU) ? ss : 0.0;
    }
    return o;
}

struct UpdOut { double ssh, sshp, u, up, v, vp; int bad; };

// Stage B for cell c: K1, K4, K6, K7, K8, K11.  Everything is evaluated unconditionally and the
// masks select at the end (branch-free; masked lanes may carry Inf/NaN that are never stored).
// rhsx / rhsy / rdx / rdy: RHSx(m,n), RHSy(m,n), dble(rdis(m,n)+rdis(m+1,n)), dble(rdis(m,n)+rdis(m,n+1)).
template <bool TRANS, bool LAT, class M>
__device__ __forceinline__ UpdOut update_cell(long c, int r, int p, const M &mt, const Tau &tau, double ts, double ffs,
        const unsigned char *__restrict__ mask,
        const double *__restrict__ ssh, const double *__restrict__ sshp,
        const double *__restrict__ u, const double *__restrict__ up,
        const double *__restrict__ v, const double *__restrict__ vp,
        const double *__restrict__ h_r, const double *__restrict__ mu,
        const double *__restrict__ hu, const double *__restrict__ hv, const double *__restrict__ hh,
        const double *__restrict__ vort, const double *__restrict__ str_t, const double *__restrict__ str_s,
        double rhsx, double rhsy, double rdx, double rdy)
{
    const long e = c + 1, no = c + p;
    const unsigned char mb = mask[c];
    const double ssh_c = ssh[c], sshp_c = sshp[c];
    const double u_c = u[c], up_c = up[c], v_c = v[c], vp_c = vp[c];
    UpdOut o;

    // K1 + K8 + K11
    const double sshn = f_sshn(c, r, p, tau.tau, mt, hu, hv, sshp, u, v);
    const bool sea = mb & MB_LU;
    o.ssh = sea ? sshn : ssh_c;
    o.sshp = sea ? f_filter(ssh_c, sshn, sshp_c, ts) : sshp_c;  // vel_ssh.f90:230-231
    o.bad = sea && !(sshn < 10000.0 && sshn > -10000.0);          // vel_ssh.f90:55

    const double h_c = h_r[c];
    const double q_c = h_c + ssh_c * ffs, qp_c = h_c + sshp_c * ffs;
    const double dx_c = mt.dx(c, r), dy_c = mt.dy(c, r);
    const int b_c = mb & MB_LU;
    const double lu_c = b_c ? 1.0 : 0.0;
    {   // zonal velocity: K10 (shp), K4, K6, K7, K8
        const double h_e = h_r[e];
        const double q_e = h_e + ssh[e] * ffs, qp_e = h_e + sshp[e] * ffs;
        const double hu_c = hu[c];
        const int b_e = mask[e] & MB_LU;
        const double hup_c = f_interp2_b<M>(qp_c, qp_e, dx_c, dy_c, lu_c, mt.dx(e, r), mt.dy(e, r), b_e ? 1.0 : 0.0,
                                            b_c + b_e, mt.dxt(c, r), mt.r_dxt(c, r), mt.dyh(c, r), mt.r_dyh(c, r));  // depth.f90:62-63
        const double adv = TRANS ? f_rhsx_adv(c, r, p, mt, md(mb, MB_LUU), md(mask[c - p], MB_LUU),
                                              u, v, vort, hu, hv, hh) : 0.0;
        const double dif = LAT ? f_rhsx_dif(c, r, p, mt, q_c, q_e, mu, str_t, str_s, hh) : 0.0;
        const double un = f_un(c, r, p, tau, mt, hu_c, hu_c, hup_c, rhsx, dif, adv, rdx, hh, ssh, v, up);
        const bool w = mb & MB_LCU;
        o.u = w ? un : u_c;
        o.up = w ? f_filter(u_c, un, up_c, ts) : up_c;
    }
    {   // meridional velocity
        const double h_n = h_r[no];
        const double q_n = h_n + ssh[no] * ffs, qp_n = h_n + sshp[no] * ffs;
        const double hv_c = hv[c];
        const int b_n = mask[no] & MB_LU;
        const double hvp_c = f_interp2_b<M>(qp_c, qp_n, dx_c, dy_c, lu_c, mt.dx(no, r + 1), mt.dy(no, r + 1),
                                            b_n ? 1.0 : 0.0, b_c + b_n, mt.dxh(c, r), mt.r_dxh(c, r),
                                            mt.dyt(c, r), mt.r_dyt(c, r));  // depth.f90:73-74
        const double adv = TRANS ? f_rhsy_adv(c, r, p, mt, u, v, vort, hu, hv, hh) : 0.0;
        const double dif = LAT ? f_rhsy_dif(c, r, p, mt, q_c, q_n, mu, str_t, str_s, hh) : 0.0;
        const double vn = f_vn(c, r, p, tau, mt, hv_c, hv_c, hvp_c, rhsy, dif, adv, rdy, hh, ssh, u, vp);
        const bool w = mb & MB_LCV;
        o.v = w ? vn : v_c;
        o.vp = w ? f_filter(v_c, vn, vp_c, ts) : vp_c;
    }
    return o;
}

// ---- tracer step (control/tracer.f90:44-61) fused into one cell function ------------------------
// hhu/hhv at the NEW time level (what K10 left in grid_data%hhu / %hhv when expl_tracer runs),
// re-evaluated from the resident state like stage A does.
template <class M>
__device__ __forceinline__ double depth_u(long c, int r, const M &mt, double ffs, const unsigned char *__restrict__ mask,
                                          const double *__restrict__ ssh, const double *__restrict__ h_r)
{
    const long e = c + 1;
    const int b_c = mask[c] & MB_LU, b_e = mask[e] & MB_LU;
    return f_interp2_b<M>(h_r[c] + ssh[c] * ffs, h_r[e] + ssh[e] * ffs, mt.dx(c, r), mt.dy(c, r), b_c ? 1.0 : 0.0,
                          mt.dx(e, r), mt.dy(e, r), b_e ? 1.0 : 0.0, b_c + b_e,
                          mt.dxt(c, r), mt.r_dxt(c, r), mt.dyh(c, r), mt.r_dyh(c, r));
}
template <class M>
__device__ __forceinline__ double depth_v(long c, int r, int p, const M &mt, double ffs,
                                          const unsigned char *__restrict__ mask,
                                          const double *__restrict__ ssh, const double *__restrict__ h_r)
{
    const long no = c + p;
    const int b_c = mask[c] & MB_LU, b_n = mask[no] & MB_LU;
    return f_interp2_b<M>(h_r[c] + ssh[c] * ffs, h_r[no] + ssh[no] * ffs, mt.dx(c, r), mt.dy(c, r), b_c ? 1.0 : 0.0,
                          mt.dx(no, r + 1), mt.dy(no, r + 1), b_n ? 1.0 : 0.0, b_c + b_n,
                          mt.dxh(c, r), mt.r_dxh(c, r), mt.dyt(c, r), mt.r_dyt(c, r));
}

// kernel/tracer/leapfrog_tracer.f90:59-73: total zonal flux through the east face of cell c
// (0 where lcu is off: the reference's flux array keeps its initial zero there)
template <class M>
__device__ __forceinline__ double tracer_flux_x(long c, int r, const M &mt, double ffs, double factor_mu,
        const unsigned char *__restrict__ mask, const double *__restrict__ ssh, const double *__restrict__ h_r,
        const double *__restrict__ uu, const double *__restrict__ mu, const double *__restrict__ ff)
{
    const long e = c + 1;
    const double hhu = depth_u(c, r, mt, ffs, mask, ssh, h_r);
    const double dfdx = ff[e] - ff[c];
    const double mu_1d = dv<M>((mu[c] + mu[e]) / 2.0 * factor_mu * mt.dyh(c, r), mt.dxt(c, r), mt.r_dxt(c, r));
    const double flux_diff = mu_1d * hhu * dfdx;
    const double flux_adv = -uu[c] * hhu * mt.dyh(c, r) * (ff[c] + ff[e]) / 2.0;
    const double f = flux_adv + flux_diff + 0.0;
    return (mask[c] & MB_LCU) ? f : 0.0;
}
// kernel/tracer/leapfrog_tracer.f90:76-90
template <class M>
__device__ __forceinline__ double tracer_flux_y(long c, int r, int p, const M &mt, double ffs, double factor_mu,
        const unsigned char *__restrict__ mask, const double *__restrict__ ssh, const double *__restrict__ h_r,
        const double *__restrict__ vv, const double *__restrict__ mu, const double *__restrict__ ff)
{
    const long no = c + p;
    const double hhv = depth_v(c, r, p, mt, ffs, mask, ssh, h_r);
    const double dfdy = ff[no] - ff[c];
    const double mu_1d = dv<M>((mu[c] + mu[no]) / 2.0 * factor_mu * mt.dxh(c, r), mt.dyt(c, r), mt.r_dyt(c, r));
    const double flux_diff = mu_1d * hhv * dfdy;
    const double flux_adv = -vv[c] * hhv * mt.dxh(c, r) * (ff[c] + ff[no]) / 2.0;
    const double f = flux_adv + flux_diff + 0.0;
    return (mask[c] & MB_LCV) ? f : 0.0;
}

struct TracerOut { double ff, ffp; };

// tran_diff_fluxes + tran_diff_tracer + tracer_next_step for cell c (leapfrog_tracer.f90:13-170), with
// hhq_n = hhq_rest and hhq_p = hhq_rest + sshp*ffs as K10 leaves them (depth.f90:48-50).
// ssh, sshp, uu, vv are the state AFTER the shallow-water step of the same model step.
template <class M>
__device__ __forceinline__ TracerOut tracer_cell(long c, int r, int p, const M &mt, const Tau &tau, double ts, double ffs,
        const unsigned char *__restrict__ mask, const double *__restrict__ ssh, const double *__restrict__ sshp,
        const double *__restrict__ h_r, const double *__restrict__ uu, const double *__restrict__ vv,
        const double *__restrict__ mu, const double *__restrict__ ff, const double *__restrict__ ffp)
{
    const double fx_c = tracer_flux_x(c, r, mt, ffs, 1.0, mask, ssh, h_r, uu, mu, ff);
    const double fx_w = tracer_flux_x(c - 1, r, mt, ffs, 1.0, mask, ssh, h_r, uu, mu, ff);
    const double fy_c = tracer_flux_y(c, r, p, mt, ffs, 1.0, mask, ssh, h_r, vv, mu, ff);
    const double fy_s = tracer_flux_y(c - p, r - 1, p, mt, ffs, 1.0, mask, ssh, h_r, vv, mu, ff);
    const double dx = mt.dx(c, r), dy = mt.dy(c, r);
    const double bp = tau.div(h_r[c] * dx * dy) / 2.0;                          // leapfrog_tracer.f90:128
    const double bp0 = tau.div((h_r[c] + sshp[c] * ffs) * dx * dy) / 2.0;       // :129
    const double rhs = fx_c - fx_w + fy_c - fy_s;
    const double eta = bp0 * ffp[c] + rhs;
    const double ffn = eta / bp;
    const bool sea = mask[c] & MB_LU;
    TracerOut o;
    o.ffp = sea ? f_filter(ff[c], ffn, ffp[c], ts) : ffp[c];                    // :163-164
    o.ff = sea ? ffn : ff[c];
    return o;
}

}  // namespace swcu
