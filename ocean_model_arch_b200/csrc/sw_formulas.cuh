// sw_formulas.cuh -- per-cell arithmetic of the shallow-water step, shared by the 1:1 kernels
// (Level A) and the fused kernels (Level B) so that both evaluate bit-identical expressions.
//
// Every function follows one statement of the reference's Fortran (file:line cited) with
// Fortran's evaluation rules: equal-precedence operators left to right, a binary operation in
// the wider kind of its operands (float*float stays float), successive divisions are true
// divisions.  This translation unit MUST be compiled with -fmad=false (and the default
// -prec-div=true -ftz=false) so no multiply-add is contracted: the results are then bitwise
// equal to a strict IEEE CPU evaluation.
#pragma once
#include <cuda_runtime.h>

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

// kernel/shallow_water/vel_ssh.f90:98-100
__device__ __forceinline__ double f_sshn(long c, int p, double tau,
        const float *__restrict__ dx, const float *__restrict__ dy,
        const float *__restrict__ dxh, const float *__restrict__ dyh,
        const double *__restrict__ hhu, const double *__restrict__ hhv,
        const double *__restrict__ sshp, const double *__restrict__ u, const double *__restrict__ v)
{
    const long w = c - 1, s = c - p;
    const float area = dx[c] * dy[c];
    const double div = u[c] * hhu[c] * dyh[c] - u[w] * hhu[w] * dyh[w]
                     + v[c] * hhv[c] * dxh[c] - v[s] * hhv[s] * dxh[s];
    return sshp[c] + 2.0 * tau * (-(div / area));
}

// kernel/shallow_water/depth.f90:59-61 (and :70-72 with the +y neighbour)
__device__ __forceinline__ double f_interp2(double hq_c, double hq_e,
        float dx_c, float dy_c, float lu_c, float dx_e, float dy_e, float lu_e, float d1, float d2)
{
    const double slu = (double)(lu_c + lu_e);
    return (hq_c * dx_c * dy_c * (double)lu_c + hq_e * dx_e * dy_e * (double)lu_e) / slu / d1 / d2;
}

// kernel/shallow_water/depth.f90:81-85
__device__ __forceinline__ double f_interp4(double hq_c, double hq_e, double hq_n, double hq_en,
        float dx_c, float dy_c, float lu_c, float dx_e, float dy_e, float lu_e,
        float dx_n, float dy_n, float lu_n, float dx_en, float dy_en, float lu_en, float dxb, float dyb)
{
    const double slu = (double)(lu_c + lu_e + lu_n + lu_en);
    return (hq_c * dx_c * dy_c * (double)lu_c + hq_e * dx_e * dy_e * (double)lu_e
          + hq_n * dx_n * dy_n * (double)lu_n + hq_en * dx_en * dy_en * (double)lu_en) / slu / dxb / dyb;
}

// kernel/shallow_water/vel_ssh.f90:273-275
__device__ __forceinline__ double f_vort(long c, int p,
        const float *__restrict__ dxt, const float *__restrict__ dyt,
        const float *__restrict__ dxb, const float *__restrict__ dyb,
        const double *__restrict__ u, const double *__restrict__ v)
{
    const long e = c + 1, no = c + p;
    return (v[e] * dyt[e] - v[c] * dyt[c])
         - (u[no] * dxt[no] - u[c] * dxt[c])
         - ((v[e] - v[c]) * dyb[c] - (u[no] - u[c]) * dxb[c]);
}

// kernel/shallow_water/mixing.f90:43-44
__device__ __forceinline__ double f_str_t(long c, int p,
        const float *__restrict__ dx, const float *__restrict__ dy,
        const float *__restrict__ dxh, const float *__restrict__ dyh,
        const double *__restrict__ u, const double *__restrict__ v)
{
    const long w = c - 1, s = c - p;
    const float ryx = dy[c] / dx[c], rxy = dx[c] / dy[c];
    return ryx * (u[c] / dyh[c] - u[w] / dyh[w]) - rxy * (v[c] / dxh[c] - v[s] / dxh[s]);
}

// kernel/shallow_water/mixing.f90:50-51
__device__ __forceinline__ double f_str_s(long c, int p,
        const float *__restrict__ dxt, const float *__restrict__ dyt,
        const float *__restrict__ dxb, const float *__restrict__ dyb,
        const double *__restrict__ u, const double *__restrict__ v)
{
    const long e = c + 1, no = c + p;
    const float rxy = dxb[c] / dyb[c], ryx = dyb[c] / dxb[c];
    return rxy * (u[no] / dxt[no] - u[c] / dxt[c]) + ryx * (v[e] / dyt[e] - v[c] / dyt[c]);
}

// kernel/shallow_water/vel_ssh.f90:326-340
__device__ __forceinline__ double f_rhsx_adv(long c, int p, float luu_c, float luu_s,
        const float *__restrict__ dxh, const float *__restrict__ dyh,
        const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ vort,
        const double *__restrict__ hu, const double *__restrict__ hv, const double *__restrict__ hh)
{
    const long e = c + 1, w = c - 1, no = c + p, s = c - p, es = c + 1 - p;
    const double fx_p = (u[c] * dyh[c] * hu[c] + u[e] * dyh[e] * hu[e]) / 2.0 * (u[c] + u[e]) / 2.0;
    const double fx_m = (u[c] * dyh[c] * hu[c] + u[w] * dyh[w] * hu[w]) / 2.0 * (u[c] + u[w]) / 2.0;
    const double fy_p = (v[c] * dxh[c] * hv[c] + v[e] * dxh[e] * hv[e]) / 2.0 * (u[no] + u[c]) / 2.0 * (double)luu_c;
    const double fy_m = (v[s] * dxh[s] * hv[s] + v[es] * dxh[es] * hv[es]) / 2.0 * (u[s] + u[c]) / 2.0 * (double)luu_s;
    return -(fx_p - fx_m + fy_p - fy_m)
         + (vort[c] * hh[c] * (v[e] + v[c]) + vort[s] * hh[s] * (v[es] + v[s])) / 4.0;
}

// kernel/shallow_water/vel_ssh.f90:351-365
__device__ __forceinline__ double f_rhsy_adv(long c, int p,
        const float *__restrict__ dxh, const float *__restrict__ dyh,
        const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ vort,
        const double *__restrict__ hu, const double *__restrict__ hv, const double *__restrict__ hh)
{
    const long e = c + 1, w = c - 1, no = c + p, s = c - p, wn = c - 1 + p;
    const double fy_p = (v[c] * dxh[c] * hv[c] + v[no] * dxh[no] * hv[no]) / 2.0 * (v[c] + v[no]) / 2.0;
    const double fy_m = (v[c] * dxh[c] * hv[c] + v[s] * dxh[s] * hv[s]) / 2.0 * (v[c] + v[s]) / 2.0;
    const double fx_p = (u[c] * dyh[c] * hu[c] + u[no] * dyh[no] * hu[no]) / 2.0 * (v[e] + v[c]) / 2.0;
    const double fx_m = (u[w] * dyh[w] * hu[w] + u[wn] * dyh[wn] * hu[wn]) / 2.0 * (v[w] + v[c]) / 2.0;
    return -(fx_p - fx_m + fy_p - fy_m)
         - (vort[c] * hh[c] * (u[no] + u[c]) + vort[w] * hh[w] * (u[wn] + u[w])) / 4.0;
}

// kernel/shallow_water/vel_ssh.f90:422-428 ; hq_c / hq_e are hq(m,n) / hq(m+1,n)
__device__ __forceinline__ double f_rhsx_dif(long c, int p, double hq_c, double hq_e,
        const float *__restrict__ dy, const float *__restrict__ dxt, const float *__restrict__ dyh,
        const float *__restrict__ dxb,
        const double *__restrict__ mu, const double *__restrict__ str_t, const double *__restrict__ str_s,
        const double *__restrict__ hh)
{
    const long e = c + 1, no = c + p, s = c - p, en = c + 1 + p, es = c + 1 - p;
    const double muh_p = (mu[c] + mu[e] + mu[no] + mu[en]) / 4.0;
    const double muh_m = (mu[c] + mu[e] + mu[s] + mu[es]) / 4.0;
    const float dy2e = dy[e] * dy[e], dy2c = dy[c] * dy[c];
    const float dxb2c = dxb[c] * dxb[c], dxb2s = dxb[s] * dxb[s];
    return (dy2e * mu[e] * hq_e * str_t[e] - dy2c * mu[c] * hq_c * str_t[c]) / dyh[c]
         + (dxb2c * muh_p * hh[c] * str_s[c] - dxb2s * muh_m * hh[s] * str_s[s]) / dxt[c];
}

// kernel/shallow_water/vel_ssh.f90:438-444 ; hq_c / hq_n are hq(m,n) / hq(m,n+1)
__device__ __forceinline__ double f_rhsy_dif(long c, int p, double hq_c, double hq_n,
        const float *__restrict__ dx, const float *__restrict__ dyt, const float *__restrict__ dxh,
        const float *__restrict__ dyb,
        const double *__restrict__ mu, const double *__restrict__ str_t, const double *__restrict__ str_s,
        const double *__restrict__ hh)
{
    const long e = c + 1, w = c - 1, no = c + p, en = c + 1 + p, wn = c - 1 + p;
    const double muh_p = (mu[c] + mu[e] + mu[no] + mu[en]) / 4.0;
    const double muh_m = (mu[c] + mu[w] + mu[no] + mu[wn]) / 4.0;
    const float dx2n = dx[no] * dx[no], dx2c = dx[c] * dx[c];
    const float dyb2c = dyb[c] * dyb[c], dyb2w = dyb[w] * dyb[w];
    return -(dx2n * mu[no] * hq_n * str_t[no] - dx2c * mu[c] * hq_c * str_t[c]) / dxh[c]
         + (dyb2c * muh_p * hh[c] * str_s[c] - dyb2w * muh_m * hh[w] * str_s[w]) / dyt[c];
}

// kernel/shallow_water/vel_ssh.f90:167-176 ; FreeFallAcc is real(4) 9.8 promoted
__device__ __forceinline__ double f_un(long c, int p, double tau,
        double hu_c, double hun_c, double hup_c, double rhsx, double rhsx_dif, double rhsx_adv,
        float rd /* rdis(m,n)+rdis(m+1,n), a real(4) sum */,
        const float *__restrict__ dxt, const float *__restrict__ dyh,
        const float *__restrict__ dxb, const float *__restrict__ dyb,
        const float *__restrict__ rlh_s,
        const double *__restrict__ hhh, const double *__restrict__ ssh,
        const double *__restrict__ v, const double *__restrict__ up)
{
    const long e = c + 1, s = c - p, es = c + 1 - p;
    const double g = (double)9.8f;
    const double bp = hun_c * dxt[c] * dyh[c] / 2.0 / tau;
    const double bp0 = hup_c * dxt[c] * dyh[c] / 2.0 / tau;
    const double slx = -g * (ssh[e] - ssh[c]) * dyh[c] * hu_c;
    const double grx = rhsx + slx + rhsx_dif + rhsx_adv
                     - rd / 2.0 * up[c] * dxt[c] * dyh[c] * hu_c
                     + (rlh_s[c] * hhh[c] * dxb[c] * dyb[c] * (v[e] + v[c])
                      + rlh_s[s] * hhh[s] * dxb[s] * dyb[s] * (v[es] + v[s])) / 4.0;
    return (up[c] * bp0 + grx) / bp;
}

// kernel/shallow_water/vel_ssh.f90:181-190
__device__ __forceinline__ double f_vn(long c, int p, double tau,
        double hv_c, double hvn_c, double hvp_c, double rhsy, double rhsy_dif, double rhsy_adv,
        float rd /* rdis(m,n)+rdis(m,n+1), a real(4) sum */,
        const float *__restrict__ dyt, const float *__restrict__ dxh,
        const float *__restrict__ dxb, const float *__restrict__ dyb,
        const float *__restrict__ rlh_s,
        const double *__restrict__ hhh, const double *__restrict__ ssh,
        const double *__restrict__ u, const double *__restrict__ vp)
{
    const long w = c - 1, no = c + p, wn = c - 1 + p;
    const double g = (double)9.8f;
    const double bp = hvn_c * dyt[c] * dxh[c] / 2.0 / tau;
    const double bp0 = hvp_c * dyt[c] * dxh[c] / 2.0 / tau;
    const double sly = -g * (ssh[no] - ssh[c]) * dxh[c] * hv_c;
    const double gry = rhsy + sly + rhsy_dif + rhsy_adv
                     - rd / 2.0 * vp[c] * dxh[c] * dyt[c] * hv_c
                     - (rlh_s[c] * hhh[c] * dxb[c] * dyb[c] * (u[no] + u[c])
                      + rlh_s[w] * hhh[w] * dxb[w] * dyb[w] * (u[wn] + u[w])) / 4.0;
    return (vp[c] * bp0 + gry) / bp;
}

// kernel/shallow_water/vel_ssh.f90:230 (and depth.f90:189, leapfrog_tracer.f90:163): Asselin filter
__device__ __forceinline__ double f_filter(double x, double xn, double xp, double ts)
{
    return x + ts * (xn - 2.0 * x + xp) / 2.0;
}

}  // namespace swcu
