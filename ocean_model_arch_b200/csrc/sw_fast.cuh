// sw_fast.cuh -- the shallow-water step in TOLERANCE mode ("exact" = 0): the same discrete scheme as
// the reference (control/shallow_water/shallow_water.f90:22-94 and its kernels), with the arithmetic of
// every statement re-associated so that a cell costs ~140 fp64 instructions instead of ~490:
//   - every division by a metric becomes a multiplication by a per-row coefficient that already holds
//     the product of the reciprocals involved (tables built once per metric upload / time step);
//   - the two divisions by the new layer thickness (vel_ssh.f90:176,190) become one reciprocal per
//     U / V point (MUFU.RCP64H seed + a cubic Newton step);
//   - products the reference evaluates several times for neighbouring cells -- u*dyh*hhu, v*dxh*hhv
//     (K1, K4), mu*hhq*str_t, muh*hhh*str_s (K6), (vort + f)*hhh*(v_e+v_c) (K4 + the Coriolis term of
//     K7), the advective face fluxes (fx_m(c) = fx_p(c-1), fy_m(c) = fy_p(c-p)) -- are evaluated once
//     at the point that owns them;
//   - explicit fma() where a product feeds a sum.
// Masks are applied with selects at the same places as the bitwise path, so land / masked-out cells
// are bit-exact (they keep their values; masked intermediates are exactly 0).  Sea values differ from
// the reference's strict left-to-right evaluation by rounding only; the north-star bound is a relative
// L2 <= 1e-12 on ssh / u / v after 1000 steps (tests/test_fast_formulas.py on the CPU,
// tests/test_gpu_fast.py on the device).  The bitwise path (sw_formulas.cuh / sw_cells.cuh, k_step)
// stays available as swcu_set_option(ctx, "exact", 1).
//
// This header is host-compilable (plain C++): tests/fast_host.cpp runs exactly these functions over
// whole arrays, so the formulas are validated against the oracle without a GPU.  Only frcp() differs
// between host (1.0 / x) and device (approximate reciprocal + Newton), by <= 1 ulp.
#pragma once
#include "sw_tables.h"

#if defined(__CUDACC__)
#define SWF_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define SWF_HD inline
#endif

namespace swf {

using swcu::MetTab;

// mask bits, identical to MB_* of sw_cells.cuh (bit <=> the reference's real(4) mask > 0.5)
enum : unsigned { LU = 1, LCU = 2, LCV = 4, LUU = 8, LUH = 16, LLU = 32, LLV = 64 };

// Columns of the per-row coefficient table [row][FC_STRIDE] (one row = 32 doubles = 256 B, every row
// self-contained: the entries of rows r-1 / r+1 a row needs are stored again as *_S / *_N; the order
// pairs entries that are loaded together as 16-byte words).
enum FastCoef : int {
    FC_KU,      // dx*dy / (dxt*dyh)            depth on U points (depth.f90:59-63)
    FC_AREA,    // dx*dy (exact in double)
    FC_AREA_N,  //   ... of row r+1
    FC_KV,      // 1 / (dxh*dyt)                depth on V points (:70-74)
    FC_KH,      // 1 / (dxb*dyb)                depth on H points (:81-85)
    FC_DYH, FC_DXH,
    FC_C1,      // dyt - dyb                    vorticity (vel_ssh.f90:273-275), re-associated
    FC_C2,      // dxt(r+1) - dxb(r)
    FC_C3,      // dxt(r) - dxb(r)
    FC_COR,     // rlh_s*dxb*dyb                Coriolis (vel_ssh.f90:173-174)
    FC_S1,      // (dy/dx) / dyh                str_t (mixing.f90:43-44)
    FC_RXY,     // dx/dy
    FC_RDXH,    // 1/dxh
    FC_RDXH_S,  //   ... of row r-1
    FC_RXYB,    // dxb/dyb                      str_s (mixing.f90:50-51)
    FC_RDXT,    // 1/dxt
    FC_RDXT_N,  //   ... of row r+1
    FC_S2,      // (dyb/dxb) / dyt
    FC_CSSH,    // 2 tau / real4(dx*dy)         K1 (vel_ssh.f90:98-100)
    FC_CU,      // 2 tau / (dxt*dyh)            1/bp without the thickness (vel_ssh.f90:167,176)
    FC_GX,      // 2 tau g / dxt                surface slope term of u
    FC_DCX,     // dy**2 / dyh                  K6 (vel_ssh.f90:422-428)
    FC_DXB2,    // dxb**2
    FC_DXB2_S,  //   ... of row r-1
    FC_CV,      // 2 tau / (dyt*dxh)
    FC_GY,      // 2 tau g / dyt
    FC_DX2,     // dx**2
    FC_DX2_N,   //   ... of row r+1
    FC_DCY,     // dyb**2 / dyt                 (vel_ssh.f90:438-444)
    FC_RDXT_B,  // copies of FC_RDXT / FC_RDXH next to the other stage-B entries (16-byte loads)
    FC_RDXH_B,
    FC_COUNT,
    FC_STRIDE = 32
};

SWF_HD double mad(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return ::fma(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}

// 1 / x to <= 1 ulp.  Device: MUFU.RCP64H seed (>= 20 bits) and one cubic Newton step
// y (1 + e + e^2), e = 1 - x y, relative error e^3 < 2^-60.
SWF_HD double frcp(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = ::fma(-x, y, 1.0);
    const double t = ::fma(e, e, e);
    return ::fma(y, t, y);
#else
    return 1.0 / x;
#endif
}

// One row of the coefficient table from the base table tab[T_COUNT][h] (sw_tables.h).
SWF_HD void build_fast_row(const double *tab, int h, int r, double tau, double *out)
{
    using namespace swcu;
    const double g = (double)9.8f;  // FreeFallAcc is real(4) (shared/constants.f90)
    const double dx = tab[T_DX * h + r], dy = tab[T_DY * h + r];
    const double dxt = tab[T_DXT * h + r], dyt = tab[T_DYT * h + r];
    const double dxh = tab[T_DXH * h + r], dyh = tab[T_DYH * h + r];
    const double dxb = tab[T_DXB * h + r], dyb = tab[T_DYB * h + r];
    const double rdxt = tab[T_RDXT * h + r], rdyt = tab[T_RDYT * h + r];
    const double rdxh = tab[T_RDXH * h + r], rdyh = tab[T_RDYH * h + r];
    const double rdxb = tab[T_RDXB * h + r], rdyb = tab[T_RDYB * h + r];
    const int rn = r + 1 < h ? r + 1 : r, rs = r > 0 ? r - 1 : r;  // clamped: edge rows feed discarded lanes only
    const double dxt_n = tab[T_DXT * h + rn];
    for (int k = 0; k < FC_STRIDE; ++k) out[k] = 0.0;
    out[FC_AREA] = dx * dy;
    out[FC_AREA_N] = tab[T_DX * h + rn] * tab[T_DY * h + rn];
    out[FC_RDXH_S] = tab[T_RDXH * h + rs];
    out[FC_RDXT_N] = tab[T_RDXT * h + rn];
    out[FC_DXB2_S] = tab[T_DXB2 * h + rs];
    out[FC_DX2_N] = tab[T_DX2 * h + rn];
    out[FC_KU] = dx * dy * rdxt * rdyh;
    out[FC_KV] = rdxh * rdyt;
    out[FC_KH] = rdxb * rdyb;
    out[FC_DYH] = dyh;
    out[FC_DXH] = dxh;
    out[FC_C1] = dyt - dyb;
    out[FC_C2] = dxt_n - dxb;
    out[FC_C3] = dxt - dxb;
    out[FC_COR] = tab[T_RLH * h + r] * dxb * dyb;
    out[FC_S1] = tab[T_RYX * h + r] * rdyh;
    out[FC_RXY] = tab[T_RXY * h + r];
    out[FC_RDXH] = rdxh;
    out[FC_RXYB] = tab[T_RXYB * h + r];
    out[FC_RDXT] = rdxt;
    out[FC_S2] = tab[T_RYXB * h + r] * rdyt;
    out[FC_CSSH] = 2.0 * tau * tab[T_RAREA * h + r];
    out[FC_CU] = 2.0 * tau * rdxt * rdyh;
    out[FC_GX] = 2.0 * tau * g * rdxt;
    out[FC_DCX] = tab[T_DY2 * h + r] * rdyh;
    out[FC_DXB2] = tab[T_DXB2 * h + r];
    out[FC_CV] = 2.0 * tau * rdyt * rdxh;
    out[FC_GY] = 2.0 * tau * g * rdyt;
    out[FC_DX2] = tab[T_DX2 * h + r];
    out[FC_DCY] = tab[T_DYB2 * h + r] * rdyt;
    out[FC_RDXT_B] = rdxt;
    out[FC_RDXH_B] = rdxh;
}

// Coefficient row of the tracer step (kernel/tracer/leapfrog_tracer.f90:13-170), [row][FT_STRIDE]
enum FastTracerCoef : int {
    FT_KU, FT_AREA,      // as FC_KU, FC_AREA
    FT_AREA_N, FT_KV,
    FT_DYH, FT_DXH,
    FT_DYH_RDXT,         // dyh / dxt   (mu_1d of the zonal flux, :63-66)
    FT_DXH_RDYT,         // dxh / dyt   (meridional flux, :80-83)
    FT_CTR,              // 2 tau / (dx*dy)   (1 / bp without the thickness, :128)
    FT_COUNT,
    FT_STRIDE = 16
};

SWF_HD void build_tracer_row(const double *tab, int h, int r, double tau, double *out)
{
    using namespace swcu;
    const int rn = r + 1 < h ? r + 1 : r;
    const double dx = tab[T_DX * h + r], dy = tab[T_DY * h + r];
    for (int k = 0; k < FT_STRIDE; ++k) out[k] = 0.0;
    out[FT_KU] = dx * dy * tab[T_RDXT * h + r] * tab[T_RDYH * h + r];
    out[FT_AREA] = dx * dy;
    out[FT_AREA_N] = tab[T_DX * h + rn] * tab[T_DY * h + rn];
    out[FT_KV] = tab[T_RDXH * h + r] * tab[T_RDYT * h + r];
    out[FT_DYH] = tab[T_DYH * h + r];
    out[FT_DXH] = tab[T_DXH * h + r];
    out[FT_DYH_RDXT] = tab[T_DYH * h + r] * tab[T_RDXT * h + r];
    out[FT_DXH_RDYT] = tab[T_DXH * h + r] * tab[T_RDYT * h + r];
    out[FT_CTR] = 2.0 * tau / (dx * dy);
}

// x / dble(lu + lu [+ lu + lu]) with 0/1 masks: a multiplication by 1, 1/2, 1/3, 1/4 (table look-up, no branch)
#if defined(__CUDACC__)
static __device__ __constant__ double c_inv_nsea[5] = {1.0, 1.0, 0.5, 1.0 / 3.0, 0.25};
#endif
SWF_HD double inv4(int nsea)
{
#if defined(__CUDA_ARCH__)
    return c_inv_nsea[nsea];
#else
    const double t[5] = {1.0, 1.0, 0.5, 1.0 / 3.0, 0.25};
    return t[nsea];
#endif
}
SWF_HD double inv2(int nsea) { return nsea == 2 ? 0.5 : 1.0; }

// ---- stage A: everything a U / V / H / T point owns, from the time-level-n state ------------------
// Row a of the block; "n" = row a+1, "s" = row a-1, "e" / "w" = column +-1.
struct ACoef {
    double ku, area, area_n, kv, kh, dyh, dxh, c1, c2, c3, cor, s1, rxy, rdxh, rdxh_s, rxyb, rdxt, rdxt_n, s2;
};

struct AOut {
    double rhu, rhv;  // 1 / hhu, 1 / hhv       (K10 / K2: depth.f90:57-74)
    double uh, vh;    // u*dyh*hhu, v*dxh*hhv   volume fluxes through the east / north face (K1, K4)
    double t;         // mu*hhq*str_t           (K5 + K6)
    double ss;        // muh*hhh*str_s
    double zx, zy;    // (vort + f)*hhh*(v_e + v_c), (vort + f)*hhh*(u_n + u_c)   (K3, K4, Coriolis of K7)
};

// qm_* = lu ? hhq_rest + ssh*ffs : 0 (masked thickness at T points), q_c the unmasked value at c;
// s_a = qm_c + qm_e (row a), s_n = qm_n + qm_en (row a+1); musum = mu_c + mu_e + mu_n + mu_en.
template <bool TRANS, bool LAT>
SWF_HD AOut stage_a(const ACoef &k, unsigned mb, int nsea_u, int nsea_v, int nsea_h,
                    double q_c, double qm_c, double qm_n, double s_a, double s_n,
                    double u_c, double u_n, double v_c, double v_e,
                    double up_c, double up_w, double up_n, double vp_c, double vp_s, double vp_e,
                    double mu_c, double musum)
{
    AOut o;
    // depth.f90:59-63, 70-74, 81-85 with dx, dy constant along the row
    const double hu = (mb & LLU) ? s_a * (k.ku * inv2(nsea_u)) : 0.0;
    const double hv = (mb & LLV) ? mad(qm_c, k.area, qm_n * k.area_n) * (k.kv * inv2(nsea_v)) : 0.0;
    const double hh = (mb & LUH) ? mad(s_a, k.area, s_n * k.area_n) * (k.kh * inv4(nsea_h)) : 0.0;
    o.rhu = frcp(hu);
    o.rhv = frcp(hv);
    o.uh = u_c * k.dyh * hu;
    o.vh = v_c * k.dxh * hv;
    // vel_ssh.f90:273-275: (v_e dyt - v_c dyt) - (u_n dxt_n - u_c dxt) - ((v_e - v_c) dyb - (u_n - u_c) dxb)
    double vort = 0.0;
    if (TRANS) {
        const double vo = mad(u_c, k.c3, mad(v_e - v_c, k.c1, -(u_n * k.c2)));
        vort = (mb & LUU) ? vo : 0.0;
    }
    const double z = (vort + k.cor) * hh;
    o.zx = z * (v_e + v_c);
    o.zy = z * (u_n + u_c);
    o.t = 0.0;
    o.ss = 0.0;
    if (LAT) {
        // mixing.f90:43-44 and :50-51 on the lagged velocities
        const double st = mad(up_c - up_w, k.s1, -(k.rxy * mad(vp_c, k.rdxh, -(vp_s * k.rdxh_s))));
        const double ssv = mad(k.rxyb, mad(up_n, k.rdxt_n, -(up_c * k.rdxt)), (vp_e - vp_c) * k.s2);
        const double str_t = (mb & LU) ? st : 0.0;
        const double str_s = (mb & LUU) ? ssv : 0.0;
        o.t = mu_c * q_c * str_t;
        o.ss = (0.25 * musum) * hh * str_s;
    }
    return o;
}

SWF_HD ACoef load_acoef(const double *row)
{
    ACoef k;
    k.ku = row[FC_KU]; k.area = row[FC_AREA]; k.area_n = row[FC_AREA_N]; k.kv = row[FC_KV]; k.kh = row[FC_KH];
    k.dyh = row[FC_DYH]; k.dxh = row[FC_DXH]; k.c1 = row[FC_C1]; k.c2 = row[FC_C2]; k.c3 = row[FC_C3];
    k.cor = row[FC_COR]; k.s1 = row[FC_S1]; k.rxy = row[FC_RXY]; k.rdxh = row[FC_RDXH]; k.rdxh_s = row[FC_RDXH_S];
    k.rxyb = row[FC_RXYB]; k.rdxt = row[FC_RDXT]; k.rdxt_n = row[FC_RDXT_N]; k.s2 = row[FC_S2];
    return k;
}

// ---- advective face fluxes (vel_ssh.f90:326-340, 351-365), times 4 ------------------------------
struct Flux {
    double fxp, fyp;    // x momentum through the east face / the north face (luu-masked)
    double fxpy, fypy;  // y momentum through the east face / the north face
};

SWF_HD Flux stage_flux(bool luu_c, double uh_c, double uh_e, double uh_n, double vh_c, double vh_e, double vh_n,
                       double u_c, double u_e, double u_n, double v_c, double v_e, double v_n)
{
    Flux f;
    f.fxp = (uh_c + uh_e) * (u_c + u_e);
    const double fy = (vh_c + vh_e) * (u_n + u_c);
    f.fyp = luu_c ? fy : 0.0;
    f.fxpy = (uh_c + uh_n) * (v_e + v_c);
    f.fypy = (vh_c + vh_n) * (v_c + v_n);
    return f;
}

// ---- stage B: K1, K4, K6, K7, K8, K11 for one cell -----------------------------------------------
struct BCoef {
    double ku, area, area_n, kv, cssh, cu, gx, dcx, dxb2, dxb2_s, rdxt, cv, gy, dx2, dx2_n, rdxh, dcy, tau;
};

SWF_HD BCoef load_bcoef(const double *row, double tau)
{
    BCoef k;
    k.ku = row[FC_KU]; k.area = row[FC_AREA]; k.area_n = row[FC_AREA_N]; k.kv = row[FC_KV]; k.cssh = row[FC_CSSH];
    k.cu = row[FC_CU]; k.gx = row[FC_GX]; k.dcx = row[FC_DCX]; k.dxb2 = row[FC_DXB2]; k.dxb2_s = row[FC_DXB2_S];
    k.rdxt = row[FC_RDXT_B]; k.cv = row[FC_CV]; k.gy = row[FC_GY]; k.dx2 = row[FC_DX2]; k.dx2_n = row[FC_DX2_N];
    k.rdxh = row[FC_RDXH_B]; k.dcy = row[FC_DCY]; k.tau = tau;
    return k;
}

// new values of one cell before the masks choose between them and the old ones
struct BRaw { double sshn, sshpf, un, upf, vn, vpf; };
struct BOut { double ssh, sshp, u, up, v, vp; int bad; };

SWF_HD double filt(double x, double xn, double xp, double ts_half)  // vel_ssh.f90:230
{
    return mad(ts_half, mad(-2.0, x, xn) + xp, x);
}

// sp_a = qpm_c + qpm_e, pp = qpm_c*area + qpm_n*area_n with qpm = lu ? hhq_rest + sshp*ffs : 0.
// f = fluxes owned by this cell, fxp_w / fxpy_w those of the west neighbour, fyp_s / fypy_s of the south one.
// The three combinations with the row below arrive ready-made (the marching kernel forms them before it
// overwrites that row's registers): dvh = vh_c - vh_s, dss = dxb2*ss_c - dxb2_s*ss_s (south_ss), zxs = zx_c + zx_s.
SWF_HD double south_ss(const BCoef &k, double ss_c, double ss_s) { return mad(k.dxb2, ss_c, -(k.dxb2_s * ss_s)); }
template <bool TRANS, bool LAT>
SWF_HD BRaw stage_b_raw(const BCoef &k, int nsea_u, int nsea_v, double ts_half,
                        double ssh_c, double ssh_e, double ssh_n, double sshp_c, double sp_a, double pp,
                        double u_c, double up_c, double v_c, double vp_c,
                        double rhu, double rhv, double uh_c, double uh_w, double dvh,
                        double t_c, double t_e, double t_n, double ss_c, double dss, double ss_w,
                        double zxs, double zy_c, double zy_w,
                        const Flux &f, double fxp_w, double fyp_s, double fxpy_w, double fypy_s,
                        double rhsx, double rhsy, double rdx, double rdy)
{
    BRaw o;
    // K1 (vel_ssh.f90:98-100) + K8
    o.sshn = mad(-k.cssh, (uh_c - uh_w) + dvh, sshp_c);
    o.sshpf = filt(ssh_c, o.sshn, sshp_c, ts_half);

    // lagged depths on U / V points (depth.f90:62-63, 73-74)
    const double hup = sp_a * (k.ku * inv2(nsea_u));
    const double hvp = pp * (k.kv * inv2(nsea_v));

    // zonal velocity: un = (up*bp0 + grx) / bp, bp = hhu dxt dyh / (2 tau)   (vel_ssh.f90:167-176)
    double rx = rhsx;
    {
        double adv4 = zxs;                                           // Coriolis (+ vorticity part of K4), times 4
        if (TRANS) adv4 -= (f.fxp - fxp_w) + (f.fyp - fyp_s);
        rx = mad(0.25, adv4, rx);
        if (LAT) rx += mad(t_e - t_c, k.dcx, dss * k.rdxt);
    }
    o.un = mad(mad(k.cu, rx, up_c * hup), rhu, -mad(k.gx, ssh_e - ssh_c, k.tau * rdx * up_c));
    o.upf = filt(u_c, o.un, up_c, ts_half);

    // meridional velocity (vel_ssh.f90:181-190)
    double ry = rhsy;
    {
        double adv4 = -(zy_c + zy_w);
        if (TRANS) adv4 -= (f.fxpy - fxpy_w) + (f.fypy - fypy_s);
        ry = mad(0.25, adv4, ry);
        if (LAT) ry += mad(ss_c - ss_w, k.dcy, -(mad(k.dx2_n, t_n, -(k.dx2 * t_c)) * k.rdxh));
    }
    o.vn = mad(mad(k.cv, ry, vp_c * hvp), rhv, -mad(k.gy, ssh_n - ssh_c, k.tau * rdy * vp_c));
    o.vpf = filt(v_c, o.vn, vp_c, ts_half);
    return o;
}

SWF_HD bool ssh_bad(double sshn) { return !(sshn < 10000.0 && sshn > -10000.0); }  // vel_ssh.f90:55

// the masks choose (K8: vel_ssh.f90:225-243; K11)
template <bool TRANS, bool LAT>
SWF_HD BOut stage_b(const BCoef &k, unsigned mb, int nsea_u, int nsea_v, double ts_half,
                    double ssh_c, double ssh_e, double ssh_n, double sshp_c, double sp_a, double pp,
                    double u_c, double up_c, double v_c, double vp_c,
                    double rhu, double rhv, double uh_c, double uh_w, double vh_c, double vh_s,
                    double t_c, double t_e, double t_n, double ss_c, double ss_s, double ss_w,
                    double zx_c, double zx_s, double zy_c, double zy_w,
                    const Flux &f, double fxp_w, double fyp_s, double fxpy_w, double fypy_s,
                    double rhsx, double rhsy, double rdx, double rdy)
{
    const BRaw r = stage_b_raw<TRANS, LAT>(k, nsea_u, nsea_v, ts_half, ssh_c, ssh_e, ssh_n, sshp_c, sp_a, pp, u_c, up_c, v_c,
                                           vp_c, rhu, rhv, uh_c, uh_w, vh_c - vh_s, t_c, t_e, t_n, ss_c,
                                           south_ss(k, ss_c, ss_s), ss_w, zx_c + zx_s, zy_c, zy_w, f, fxp_w, fyp_s, fxpy_w,
                                           fypy_s, rhsx, rhsy, rdx, rdy);
    BOut o;
    const bool sea = (mb & LU) != 0, wu = (mb & LCU) != 0, wv = (mb & LCV) != 0;
    o.ssh = sea ? r.sshn : ssh_c;
    o.sshp = sea ? r.sshpf : sshp_c;
    o.bad = sea && ssh_bad(r.sshn);
    o.u = wu ? r.un : u_c;
    o.up = wu ? r.upf : up_c;
    o.v = wv ? r.vn : v_c;
    o.vp = wv ? r.vpf : vp_c;
    return o;
}


// ---- tracer step (control/tracer.f90:44-61: fluxes + update + filter) in tolerance arithmetic ---------
struct TCoef { double ku, area, area_n, kv, dyh, dxh, dyh_rdxt, dxh_rdyt, ctr; };

SWF_HD TCoef load_tcoef(const double *row)
{
    TCoef k;
    k.ku = row[FT_KU]; k.area = row[FT_AREA]; k.area_n = row[FT_AREA_N]; k.kv = row[FT_KV]; k.dyh = row[FT_DYH];
    k.dxh = row[FT_DXH]; k.dyh_rdxt = row[FT_DYH_RDXT]; k.dxh_rdyt = row[FT_DXH_RDYT]; k.ctr = row[FT_CTR];
    return k;
}

struct TFlux { double fx, fy; };

// total fluxes through the east / north face of cell c (leapfrog_tracer.f90:59-90), factor_mu = 1.
// qm_* = lu ? hhq_rest + ssh_new*ffs : 0: the depths on U / V points are those of the NEW level (what K10 left).
SWF_HD TFlux tracer_flux(const TCoef &k, unsigned mb, int nsea_u, int nsea_v, double qm_c, double qm_e, double qm_n,
                         double u_c, double v_c, double mu_c, double mu_e, double mu_n, double ff_c, double ff_e, double ff_n)
{
    TFlux f;
    const double hhu = (qm_c + qm_e) * (k.ku * inv2(nsea_u));
    const double hhv = mad(qm_c, k.area, qm_n * k.area_n) * (k.kv * inv2(nsea_v));
    const double gx = mad(0.5 * (mu_c + mu_e) * k.dyh_rdxt, ff_e - ff_c, -(u_c * k.dyh * (0.5 * (ff_c + ff_e))));
    const double gy = mad(0.5 * (mu_c + mu_n) * k.dxh_rdyt, ff_n - ff_c, -(v_c * k.dxh * (0.5 * (ff_c + ff_n))));
    f.fx = (mb & LCU) ? hhu * gx : 0.0;
    f.fy = (mb & LCV) ? hhv * gy : 0.0;
    return f;
}

struct TOut { double ffn, ffpf; };

// leapfrog_tracer.f90:128-134 and :163: ffn = (bp0*ffp + rhs) / bp with bp = hhq_rest dx dy / (2 tau),
// bp0 = (hhq_rest + sshp_new*ffs) dx dy / (2 tau); then the time filter.  qp_c = hhq_rest + sshp_new*ffs.
SWF_HD TOut tracer_update(const TCoef &k, double ts_half, double h_c, double qp_c, double ff_c, double ffp_c,
                          double fx_c, double fx_w, double fy_c, double fy_s)
{
    TOut o;
    const double rhs = (fx_c - fx_w) + (fy_c - fy_s);
    o.ffn = mad(rhs, k.ctr, qp_c * ffp_c) * frcp(h_c);
    o.ffpf = filt(ff_c, o.ffn, ffp_c, ts_half);
    return o;
}

}  // namespace swf
