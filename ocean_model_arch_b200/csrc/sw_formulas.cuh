// sw_formulas.cuh -- per-cell arithmetic of the shallow-water step, shared by the 1:1 kernels
// (Level A) and the fused kernels (Level B) so that both evaluate bit-identical expressions.
//
// Every function follows one statement of the reference's Fortran (file:line cited) with
// Fortran's evaluation rules: equal-precedence operators left to right, a binary operation in
// the wider kind of its operands (real(4)*real(4) stays real(4)), successive divisions are true
// divisions.  This translation unit MUST be compiled with -fmad=false (and the default
// -prec-div=true -ftz=false) so no multiply-add is contracted: the results are then bitwise
// equal to a strict IEEE CPU evaluation.
//
// The reference's real(4) metric / Coriolis arrays enter through a "metric accessor" M that
// returns each quantity ALREADY PROMOTED to double (promotion is exact, so where it happens does
// not change a bit).  real(4) sub-expressions of the reference (dx*dy, dy**2, dy/dx, ...) are
// separate accessor methods that evaluate in real(4) first.  Two accessors exist:
//   MetGen  reads the real(4) 2-D arrays and converts per use (any grid; Level A and fallback);
//   MetRow  reads per-row tables of doubles built once on the device (grids whose metrics depend
//           on the row only: carthesian and unrotated spherical, i.e. every shipped setup).
// MetRow exists because on B200 the F2F.F64.F32 conversion issues on the quarter-rate XU pipe:
// ncu showed the converting kernel XU-bound (62 % XU, 26 % fp64, 25 % DRAM).
#pragma once
#include <cuda_runtime.h>

#include "sw_tables.h"

namespace swcu {

// Geometry of one block array on the device: element (m,n) at base[(n-by1)*pitch + (m-bx1)].
struct Geo {
    int nx_start, nx_end, ny_start, ny_end;
    int bx1, bx2, by1, by2;
    int pitch;  // elements per row (>= bx2-bx1+1)
};

__device__ __forceinline__ long ix(const Geo &g, int m, int n)
{
    return (long)(n - g.by1) * g.pitch + (m - g.bx1);
}

__device__ __forceinline__ bool on(float mask) { return mask > 0.5f; }

// ---- metric accessors -------------------------------------------------------------------------
// c = flat index of the cell, r = its row (n - by1).  Neighbours: (m+-1) -> (c+-1, r),
// (n+-1) -> (c+-p, r+-1).
struct MetGen {
    static constexpr bool kRecip = false;  // no reciprocal tables: divisions stay hardware divisions
    const float *dx_, *dy_, *dxt_, *dyt_, *dxh_, *dyh_, *dxb_, *dyb_, *rlh_;
    __device__ __forceinline__ double r_dxt(long, int) const { return 0.0; }
    __device__ __forceinline__ double r_dyt(long, int) const { return 0.0; }
    __device__ __forceinline__ double r_dxh(long, int) const { return 0.0; }
    __device__ __forceinline__ double r_dyh(long, int) const { return 0.0; }
    __device__ __forceinline__ double r_dxb(long, int) const { return 0.0; }
    __device__ __forceinline__ double r_dyb(long, int) const { return 0.0; }
    __device__ __forceinline__ double r_area(long, int) const { return 0.0; }
    __device__ __forceinline__ double dx(long c, int) const { return (double)dx_[c]; }
    __device__ __forceinline__ double dy(long c, int) const { return (double)dy_[c]; }
    __device__ __forceinline__ double dxt(long c, int) const { return (double)dxt_[c]; }
    __device__ __forceinline__ double dyt(long c, int) const { return (double)dyt_[c]; }
    __device__ __forceinline__ double dxh(long c, int) const { return (double)dxh_[c]; }
    __device__ __forceinline__ double dyh(long c, int) const { return (double)dyh_[c]; }
    __device__ __forceinline__ double dxb(long c, int) const { return (double)dxb_[c]; }
    __device__ __forceinline__ double dyb(long c, int) const { return (double)dyb_[c]; }
    __device__ __forceinline__ double rlh(long c, int) const { return (double)rlh_[c]; }
    // real(4) sub-expressions of the reference
    __device__ __forceinline__ double area(long c, int) const { return (double)(dx_[c] * dy_[c]); }    // dx*dy
    __device__ __forceinline__ double dy2(long c, int) const { return (double)(dy_[c] * dy_[c]); }     // dy**2
    __device__ __forceinline__ double dx2(long c, int) const { return (double)(dx_[c] * dx_[c]); }
    __device__ __forceinline__ double dxb2(long c, int) const { return (double)(dxb_[c] * dxb_[c]); }
    __device__ __forceinline__ double dyb2(long c, int) const { return (double)(dyb_[c] * dyb_[c]); }
    __device__ __forceinline__ double ryx(long c, int) const { return (double)(dy_[c] / dx_[c]); }     // dy/dx
    __device__ __forceinline__ double rxy(long c, int) const { return (double)(dx_[c] / dy_[c]); }     // dx/dy
    __device__ __forceinline__ double rxyb(long c, int) const { return (double)(dxb_[c] / dyb_[c]); }  // dxb/dyb
    __device__ __forceinline__ double ryxb(long c, int) const { return (double)(dyb_[c] / dxb_[c]); }  // dyb/dxb
};

// (enum MetTab: sw_tables.h, shared with the host-compilable sw_fast.cuh)

struct MetRow {
    static constexpr bool kRecip = true;
    const double *tab;  // [T_COUNT][h]: global tables, or a shared-memory copy of rows r0 .. r0+h-1
    int h;
    int r0;
#define SWCU_ROW(name, T) \
    __device__ __forceinline__ double name(long, int r) const { return tab[(T) * h + (r - r0)]; }
    SWCU_ROW(dx, T_DX) SWCU_ROW(dy, T_DY) SWCU_ROW(dxt, T_DXT) SWCU_ROW(dyt, T_DYT)
    SWCU_ROW(dxh, T_DXH) SWCU_ROW(dyh, T_DYH) SWCU_ROW(dxb, T_DXB) SWCU_ROW(dyb, T_DYB)
    SWCU_ROW(rlh, T_RLH) SWCU_ROW(area, T_AREA) SWCU_ROW(dy2, T_DY2) SWCU_ROW(dx2, T_DX2)
    SWCU_ROW(dxb2, T_DXB2) SWCU_ROW(dyb2, T_DYB2) SWCU_ROW(ryx, T_RYX) SWCU_ROW(rxy, T_RXY)
    SWCU_ROW(rxyb, T_RXYB) SWCU_ROW(ryxb, T_RYXB)
    SWCU_ROW(r_dxt, T_RDXT) SWCU_ROW(r_dyt, T_RDYT) SWCU_ROW(r_dxh, T_RDXH) SWCU_ROW(r_dyh, T_RDYH)
    SWCU_ROW(r_dxb, T_RDXB) SWCU_ROW(r_dyb, T_RDYB) SWCU_ROW(r_area, T_RAREA)
#undef SWCU_ROW
};

// ---- exact division by a tabulated divisor ------------------------------------------------------
// Correctly rounded a/b from y = RN(1/b) without the hardware division sequence (MUFU.RCP64H on the
// slow XU pipe + ~8 dependent DFMA + range checks).  q0 = RN(a*y) is within 2 ulp of a/b; one
// residual correction makes q1 faithful (error <= 1/2 ulp + 2^-104 relative); Markstein's theorem
// (y the correctly rounded reciprocal, q1 faithful, r1 = a - b*q1 exact in one FMA) then gives
// RN(q1 + r1*y) = RN(a/b), i.e. the same bits as the IEEE division the reference performs.
// Valid for normal-range operands (every quantity of this model).  The only case the FMA chain gets
// wrong is the SIGN of a zero quotient (a = -0 gives +0); since sign(a/b) = sign(a*y) always, the sign
// bit of q0 is copied onto the result (one LOP3).  Explicit fma() is used on purpose: -fmad=false only
// forbids implicit contraction.
__device__ __forceinline__ double mdiv(double a, double b, double y)
{
    const double q0 = a * y;
    const double r0 = fma(-b, q0, a);
    const double q1 = fma(r0, y, q0);
    const double r1 = fma(-b, q1, a);
    const double q = fma(r1, y, q1);
    return __hiloint2double((__double2hiint(q) & 0x7fffffff) | (__double2hiint(q0) & 0x80000000), __double2loint(q));
}
template <class M>
__device__ __forceinline__ double dv(double a, double b, double y)
{
    if constexpr (M::kRecip) return mdiv(a, b, y);
    else return a / b;
}

// x / slu for slu = dble(lu+lu[+lu+lu]) with 0/1 masks, nsea = number of sea cells among them.
// Division by 1, 2, 4 is an exact scaling, so multiplying by the exact reciprocal gives the same
// bits as the reference's division; 3 goes through mdiv with RN(1/3).  (nsea = 0 only on masked-out
// cells.)
__device__ __forceinline__ double div_slu(double x, int nsea)
{
    if (nsea == 3) return mdiv(x, 3.0, 1.0 / 3.0);
    return x * (nsea == 2 ? 0.5 : (nsea == 4 ? 0.25 : 1.0));
}

// ---- formulas -----------------------------------------------------------------------------------

// kernel/shallow_water/vel_ssh.f90:98-100
template <class M>
__device__ __forceinline__ double f_sshn(long c, int r, int p, double tau, const M &m,
        const double *__restrict__ hhu, const double *__restrict__ hhv,
        const double *__restrict__ sshp, const double *__restrict__ u, const double *__restrict__ v)
{
    const long w = c - 1, s = c - p;
    const double div = u[c] * hhu[c] * m.dyh(c, r) - u[w] * hhu[w] * m.dyh(w, r)
                     + v[c] * hhv[c] * m.dxh(c, r) - v[s] * hhv[s] * m.dxh(s, r - 1);
    return sshp[c] + 2.0 * tau * (-dv<M>(div, m.area(c, r), m.r_area(c, r)));
}

// kernel/shallow_water/depth.f90:59-61 (and :70-72 with the +y neighbour).  lu_* are the real(4)
// masks promoted to double (their real(4) sum is exact, so summing the doubles gives the same slu).
__device__ __forceinline__ double f_interp2(double hq_c, double hq_e,
        double dx_c, double dy_c, double lu_c, double dx_e, double dy_e, double lu_e, double d1, double d2)
{
    const double slu = lu_c + lu_e;
    return (hq_c * dx_c * dy_c * lu_c + hq_e * dx_e * dy_e * lu_e) / slu / d1 / d2;
}

// f_interp2 / f_interp4 for masks known to be exactly 0 or 1 (fused path: masks come from bits)
template <class M>
__device__ __forceinline__ double f_interp2_b(double hq_c, double hq_e,
        double dx_c, double dy_c, double lu_c, double dx_e, double dy_e, double lu_e, int nsea,
        double d1, double rd1, double d2, double rd2)
{
    return dv<M>(dv<M>(div_slu(hq_c * dx_c * dy_c * lu_c + hq_e * dx_e * dy_e * lu_e, nsea), d1, rd1), d2, rd2);
}
template <class M>
__device__ __forceinline__ double f_interp4_b(double hq_c, double hq_e, double hq_n, double hq_en,
        double dx_c, double dy_c, double lu_c, double dx_e, double dy_e, double lu_e,
        double dx_n, double dy_n, double lu_n, double dx_en, double dy_en, double lu_en, int nsea,
        double dxb, double rdxb, double dyb, double rdyb)
{
    return dv<M>(dv<M>(div_slu(hq_c * dx_c * dy_c * lu_c + hq_e * dx_e * dy_e * lu_e
                             + hq_n * dx_n * dy_n * lu_n + hq_en * dx_en * dy_en * lu_en, nsea), dxb, rdxb), dyb, rdyb);
}

// x / slu for an arbitrary real(4)-sum slu: 1, 2 and 4 are exact scalings (same bits as the division),
// everything else divides
__device__ __forceinline__ double div_slu_any(double x, double slu)
{
    if (slu == 1.0) return x;
    if (slu == 2.0) return x * 0.5;
    if (slu == 4.0) return x * 0.25;
    return x / slu;
}
// numerator of depth.f90:60-61
__device__ __forceinline__ double hq_sum2(double hq_c, double hq_e, double dx_c, double dy_c, double lu_c,
                                          double dx_e, double dy_e, double lu_e)
{
    return hq_c * dx_c * dy_c * lu_c + hq_e * dx_e * dy_e * lu_e;
}

// kernel/shallow_water/depth.f90:81-85
__device__ __forceinline__ double f_interp4(double hq_c, double hq_e, double hq_n, double hq_en,
        double dx_c, double dy_c, double lu_c, double dx_e, double dy_e, double lu_e,
        double dx_n, double dy_n, double lu_n, double dx_en, double dy_en, double lu_en, double dxb, double dyb)
{
    const double slu = lu_c + lu_e + lu_n + lu_en;
    return (hq_c * dx_c * dy_c * lu_c + hq_e * dx_e * dy_e * lu_e
          + hq_n * dx_n * dy_n * lu_n + hq_en * dx_en * dy_en * lu_en) / slu / dxb / dyb;
}

// kernel/shallow_water/vel_ssh.f90:273-275
template <class M>
__device__ __forceinline__ double f_vort(long c, int r, int p, const M &m,
        const double *__restrict__ u, const double *__restrict__ v)
{
    const long e = c + 1, no = c + p;
    return (v[e] * m.dyt(e, r) - v[c] * m.dyt(c, r))
         - (u[no] * m.dxt(no, r + 1) - u[c] * m.dxt(c, r))
         - ((v[e] - v[c]) * m.dyb(c, r) - (u[no] - u[c]) * m.dxb(c, r));
}

// kernel/shallow_water/mixing.f90:43-44
template <class M>
__device__ __forceinline__ double f_str_t(long c, int r, int p, const M &m,
        const double *__restrict__ u, const double *__restrict__ v)
{
    const long w = c - 1, s = c - p;
    return m.ryx(c, r) * (dv<M>(u[c], m.dyh(c, r), m.r_dyh(c, r)) - dv<M>(u[w], m.dyh(w, r), m.r_dyh(w, r)))
         - m.rxy(c, r) * (dv<M>(v[c], m.dxh(c, r), m.r_dxh(c, r)) - dv<M>(v[s], m.dxh(s, r - 1), m.r_dxh(s, r - 1)));
}

// kernel/shallow_water/mixing.f90:50-51
template <class M>
__device__ __forceinline__ double f_str_s(long c, int r, int p, const M &m,
        const double *__restrict__ u, const double *__restrict__ v)
{
    const long e = c + 1, no = c + p;
    return m.rxyb(c, r) * (dv<M>(u[no], m.dxt(no, r + 1), m.r_dxt(no, r + 1)) - dv<M>(u[c], m.dxt(c, r), m.r_dxt(c, r)))
         + m.ryxb(c, r) * (dv<M>(v[e], m.dyt(e, r), m.r_dyt(e, r)) - dv<M>(v[c], m.dyt(c, r), m.r_dyt(c, r)));
}

// kernel/shallow_water/vel_ssh.f90:326-340 ; luu_c / luu_s = dble(luu(m,n)) / dble(luu(m,n-1))
template <class M>
__device__ __forceinline__ double f_rhsx_adv(long c, int r, int p, const M &m, double luu_c, double luu_s,
        const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ vort,
        const double *__restrict__ hu, const double *__restrict__ hv, const double *__restrict__ hh)
{
    const long e = c + 1, w = c - 1, no = c + p, s = c - p, es = c + 1 - p;
    const double fx_p = (u[c] * m.dyh(c, r) * hu[c] + u[e] * m.dyh(e, r) * hu[e]) / 2.0 * (u[c] + u[e]) / 2.0;
    const double fx_m = (u[c] * m.dyh(c, r) * hu[c] + u[w] * m.dyh(w, r) * hu[w]) / 2.0 * (u[c] + u[w]) / 2.0;
    const double fy_p = (v[c] * m.dxh(c, r) * hv[c] + v[e] * m.dxh(e, r) * hv[e]) / 2.0 * (u[no] + u[c]) / 2.0 * luu_c;
    const double fy_m = (v[s] * m.dxh(s, r - 1) * hv[s] + v[es] * m.dxh(es, r - 1) * hv[es]) / 2.0 * (u[s] + u[c]) / 2.0 * luu_s;
    return -(fx_p - fx_m + fy_p - fy_m)
         + (vort[c] * hh[c] * (v[e] + v[c]) + vort[s] * hh[s] * (v[es] + v[s])) / 4.0;
}

// kernel/shallow_water/vel_ssh.f90:351-365
template <class M>
__device__ __forceinline__ double f_rhsy_adv(long c, int r, int p, const M &m,
        const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ vort,
        const double *__restrict__ hu, const double *__restrict__ hv, const double *__restrict__ hh)
{
    const long e = c + 1, w = c - 1, no = c + p, s = c - p, wn = c - 1 + p;
    const double fy_p = (v[c] * m.dxh(c, r) * hv[c] + v[no] * m.dxh(no, r + 1) * hv[no]) / 2.0 * (v[c] + v[no]) / 2.0;
    const double fy_m = (v[c] * m.dxh(c, r) * hv[c] + v[s] * m.dxh(s, r - 1) * hv[s]) / 2.0 * (v[c] + v[s]) / 2.0;
    const double fx_p = (u[c] * m.dyh(c, r) * hu[c] + u[no] * m.dyh(no, r + 1) * hu[no]) / 2.0 * (v[e] + v[c]) / 2.0;
    const double fx_m = (u[w] * m.dyh(w, r) * hu[w] + u[wn] * m.dyh(wn, r + 1) * hu[wn]) / 2.0 * (v[w] + v[c]) / 2.0;
    return -(fx_p - fx_m + fy_p - fy_m)
         - (vort[c] * hh[c] * (u[no] + u[c]) + vort[w] * hh[w] * (u[wn] + u[w])) / 4.0;
}

// kernel/shallow_water/vel_ssh.f90:422-428 ; hq_c / hq_e are hq(m,n) / hq(m+1,n)
template <class M>
__device__ __forceinline__ double f_rhsx_dif(long c, int r, int p, const M &m, double hq_c, double hq_e,
        const double *__restrict__ mu, const double *__restrict__ str_t, const double *__restrict__ str_s,
        const double *__restrict__ hh)
{
    const long e = c + 1, no = c + p, s = c - p, en = c + 1 + p, es = c + 1 - p;
    const double muh_p = (mu[c] + mu[e] + mu[no] + mu[en]) / 4.0;
    const double muh_m = (mu[c] + mu[e] + mu[s] + mu[es]) / 4.0;
    return dv<M>(m.dy2(e, r) * mu[e] * hq_e * str_t[e] - m.dy2(c, r) * mu[c] * hq_c * str_t[c], m.dyh(c, r), m.r_dyh(c, r))
         + dv<M>(m.dxb2(c, r) * muh_p * hh[c] * str_s[c] - m.dxb2(s, r - 1) * muh_m * hh[s] * str_s[s],
                 m.dxt(c, r), m.r_dxt(c, r));
}

// kernel/shallow_water/vel_ssh.f90:438-444 ; hq_c / hq_n are hq(m,n) / hq(m,n+1)
template <class M>
__device__ __forceinline__ double f_rhsy_dif(long c, int r, int p, const M &m, double hq_c, double hq_n,
        const double *__restrict__ mu, const double *__restrict__ str_t, const double *__restrict__ str_s,
        const double *__restrict__ hh)
{
    const long e = c + 1, w = c - 1, no = c + p, en = c + 1 + p, wn = c - 1 + p;
    const double muh_p = (mu[c] + mu[e] + mu[no] + mu[en]) / 4.0;
    const double muh_m = (mu[c] + mu[w] + mu[no] + mu[wn]) / 4.0;
    return -dv<M>(m.dx2(no, r + 1) * mu[no] * hq_n * str_t[no] - m.dx2(c, r) * mu[c] * hq_c * str_t[c],
                  m.dxh(c, r), m.r_dxh(c, r))
         + dv<M>(m.dyb2(c, r) * muh_p * hh[c] * str_s[c] - m.dyb2(w, r) * muh_m * hh[w] * str_s[w],
                 m.dyt(c, r), m.r_dyt(c, r));
}

// x / tau.  When tau is a power of two the division is an exact scaling and x * (1/tau) gives the
// same bits (the host checks frexp(tau) == 0.5 * 2^e); otherwise mdiv with rtau = RN(1/tau) from the
// host's IEEE division.  exact = 0 (Level A, which has no host-side setup) keeps the hardware division.
struct Tau {
    double tau, rtau;
    int pow2, exact;
    __device__ __forceinline__ double div(double x) const
    {
        if (pow2) return x * rtau;
        return exact ? mdiv(x, tau, rtau) : x / tau;
    }
};

// kernel/shallow_water/vel_ssh.f90:167-176 ; FreeFallAcc is real(4) 9.8 promoted;
// rd = dble(rdis(m,n)+rdis(m+1,n)), the real(4) sum promoted
template <class M>
__device__ __forceinline__ double f_un(long c, int r, int p, const Tau &tau, const M &m,
        double hu_c, double hun_c, double hup_c, double rhsx, double rhsx_dif, double rhsx_adv, double rd,
        const double *__restrict__ hhh, const double *__restrict__ ssh,
        const double *__restrict__ v, const double *__restrict__ up)
{
    const long e = c + 1, s = c - p, es = c + 1 - p;
    const double g = (double)9.8f;
    const double dxt = m.dxt(c, r), dyh = m.dyh(c, r);
    const double bp = tau.div(hun_c * dxt * dyh / 2.0);
    const double bp0 = tau.div(hup_c * dxt * dyh / 2.0);
    const double slx = -g * (ssh[e] - ssh[c]) * dyh * hu_c;
    const double grx = rhsx + slx + rhsx_dif + rhsx_adv
                     - rd / 2.0 * up[c] * dxt * dyh * hu_c
                     + (m.rlh(c, r) * hhh[c] * m.dxb(c, r) * m.dyb(c, r) * (v[e] + v[c])
                      + m.rlh(s, r - 1) * hhh[s] * m.dxb(s, r - 1) * m.dyb(s, r - 1) * (v[es] + v[s])) / 4.0;
    return (up[c] * bp0 + grx) / bp;
}

// kernel/shallow_water/vel_ssh.f90:181-190 ; rd = dble(rdis(m,n)+rdis(m,n+1))
template <class M>
__device__ __forceinline__ double f_vn(long c, int r, int p, const Tau &tau, const M &m,
        double hv_c, double hvn_c, double hvp_c, double rhsy, double rhsy_dif, double rhsy_adv, double rd,
        const double *__restrict__ hhh, const double *__restrict__ ssh,
        const double *__restrict__ u, const double *__restrict__ vp)
{
    const long w = c - 1, no = c + p, wn = c - 1 + p;
    const double g = (double)9.8f;
    const double dyt = m.dyt(c, r), dxh = m.dxh(c, r);
    const double bp = tau.div(hvn_c * dyt * dxh / 2.0);
    const double bp0 = tau.div(hvp_c * dyt * dxh / 2.0);
    const double sly = -g * (ssh[no] - ssh[c]) * dxh * hv_c;
    const double gry = rhsy + sly + rhsy_dif + rhsy_adv
                     - rd / 2.0 * vp[c] * dxh * dyt * hv_c
                     - (m.rlh(c, r) * hhh[c] * m.dxb(c, r) * m.dyb(c, r) * (u[no] + u[c])
                      + m.rlh(w, r) * hhh[w] * m.dxb(w, r) * m.dyb(w, r) * (u[wn] + u[w])) / 4.0;
    return (vp[c] * bp0 + gry) / bp;
}

// kernel/shallow_water/vel_ssh.f90:230 (and depth.f90:189, leapfrog_tracer.f90:163): Asselin filter
__device__ __forceinline__ double f_filter(double x, double xn, double xp, double ts)
{
    return x + ts * (xn - 2.0 * x + xp) / 2.0;
}

}  // namespace swcu
