// sw_kernels_ref.cu -- Level A: one CUDA kernel per reference kernel (K1..K11 + 3 tracer kernels),
// written for sm_100a.  One thread per cell, warps run along the contiguous m (x) dimension so
// every global access is coalesced; masked cells are handled with predicated stores (the value is
// always computed, `if (mask>0.5)` guards only the store, exactly like the reference keeps the old
// value).  The arithmetic lives in sw_formulas.cuh.  Compile with -fmad=false.
#include "sw_common.h"

namespace swcu {

namespace {

constexpr int BX = 64;  // threads along m (2 warps = 512 contiguous bytes of fp64 per row)
constexpr int BY = 4;   // rows per CTA

inline dim3 grid_for(int m0, int m1, int n0, int n1)
{
    return dim3((unsigned)((m1 - m0 + BX) / BX), (unsigned)((n1 - n0 + BY) / BY), 1);
}
inline int launched(const char *what)
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SWCU_OK : cuda_fail(e, what);
}

// Level B (REFERENCE mode) may hand the kernels the context's per-row metric tables: ROW = true reads
// them (no real(4)->real(8) conversions, exact mdiv divisions), ROW = false reads the real(4) arrays
// the reference passes (Level A, any grid).
template <bool ROW> struct Pick {
    static __device__ __forceinline__ const MetGen &get(const MetGen &a, const MetRow &) { return a; }
};
template <> struct Pick<true> {
    static __device__ __forceinline__ const MetRow &get(const MetGen &, const MetRow &b) { return b; }
};
#define SWCU_LAUNCH_ROW(kern, grid, ...)                                          \
    do {                                                                          \
        if (mr) kern<true><<<grid, kBlock, 0, st>>>(g, *mr, __VA_ARGS__);         \
        else kern<false><<<grid, kBlock, 0, st>>>(g, MetRow{nullptr, 0, 0}, __VA_ARGS__); \
    } while (0)

#define SWCU_CELL(m0, n0, m1, n1)                          \
    const int m = (m0) + blockIdx.x * BX + threadIdx.x;    \
    const int n = (n0) + blockIdx.y * BY + threadIdx.y;    \
    if (m > (m1) || n > (n1)) return;                      \
    const long c = ix(g, m, n);                            \
    const int p = g.pitch;                                 \
    const int r = n - g.by1;                               \
    (void)p; (void)r

// K1 -- kernel/shallow_water/vel_ssh.f90:94-104
template <bool ROW>
__global__ void __launch_bounds__(BX *BY) k_sw_update_ssh(Geo g, MetRow mr, double tau,
        const float *__restrict__ lu, const float *__restrict__ dx, const float *__restrict__ dy,
        const float *__restrict__ dxh, const float *__restrict__ dyh,
        const double *__restrict__ hhu, const double *__restrict__ hhv, double *__restrict__ sshn,
        const double *__restrict__ sshp, const double *__restrict__ u, const double *__restrict__ v)
{
    SWCU_CELL(g.nx_start, g.ny_start, g.nx_end, g.ny_end);
    const MetGen mg{dx, dy, nullptr, nullptr, dxh, dyh, nullptr, nullptr, nullptr};
    const auto &mt = Pick<ROW>::get(mg, mr);
    // evaluate first, store under the mask: the operand loads then do not wait for the mask load
    // (masked-out lanes may compute Inf/NaN from zero metrics; they are never stored)
    const double val = f_sshn(c, r, p, tau, mt, hhu, hhv, sshp, u, v);
    if (on(lu[c])) sshn[c] = val;
}

// K7 -- kernel/shallow_water/vel_ssh.f90:163-193
template <bool ROW>
__global__ void __launch_bounds__(BX *BY) k_sw_update_uv(Geo g, MetRow mr, Tau tt,
        const float *__restrict__ lcu, const float *__restrict__ lcv,
        const float *__restrict__ dxt, const float *__restrict__ dyt,
        const float *__restrict__ dxh, const float *__restrict__ dyh,
        const float *__restrict__ dxb, const float *__restrict__ dyb,
        const double *__restrict__ hhu, const double *__restrict__ hhun, const double *__restrict__ hhup,
        const double *__restrict__ hhv, const double *__restrict__ hhvn, const double *__restrict__ hhvp,
        const double *__restrict__ hhh, const double *__restrict__ ssh,
        const double *__restrict__ u, double *__restrict__ un, const double *__restrict__ up,
        const double *__restrict__ v, double *__restrict__ vn, const double *__restrict__ vp,
        const float *__restrict__ rdis, const float *__restrict__ rlh_s,
        const double *__restrict__ RHSx, const double *__restrict__ RHSy,
        const double *__restrict__ RHSx_adv, const double *__restrict__ RHSy_adv,
        const double *__restrict__ RHSx_dif, const double *__restrict__ RHSy_dif)
{
    SWCU_CELL(g.nx_start, g.ny_start, g.nx_end, g.ny_end);
    const MetGen mg{nullptr, nullptr, dxt, dyt, dxh, dyh, dxb, dyb, rlh_s};
    const auto &mt = Pick<ROW>::get(mg, mr);
    const double nu = f_un(c, r, p, tt, mt, hhu[c], hhun[c], hhup[c], RHSx[c], RHSx_dif[c], RHSx_adv[c],
                           (double)(rdis[c] + rdis[c + 1]), hhh, ssh, v, up);
    const double nv = f_vn(c, r, p, tt, mt, hhv[c], hhvn[c], hhvp[c], RHSy[c], RHSy_dif[c], RHSy_adv[c],
                           (double)(rdis[c] + rdis[c + p]), hhh, ssh, u, vp);
    if (on(lcu[c])) un[c] = nu;
    if (on(lcv[c])) vn[c] = nv;
}

// K8 -- kernel/shallow_water/vel_ssh.f90:226-243 (range grown by one cell)
__global__ void __launch_bounds__(BX *BY) k_sw_next_step(Geo g, double ts,
        const float *__restrict__ lu, const float *__restrict__ lcu, const float *__restrict__ lcv,
        double *__restrict__ ssh, const double *__restrict__ sshn, double *__restrict__ sshp,
        double *__restrict__ u, const double *__restrict__ un, double *__restrict__ up,
        double *__restrict__ v, const double *__restrict__ vn, double *__restrict__ vp)
{
    SWCU_CELL(g.nx_start - 1, g.ny_start - 1, g.nx_end + 1, g.ny_end + 1);
    if (on(lu[c])) {
        sshp[c] = f_filter(ssh[c], sshn[c], sshp[c], ts);
        ssh[c] = sshn[c];
    }
    if (on(lcu[c])) {
        up[c] = f_filter(u[c], un[c], up[c], ts);
        u[c] = un[c];
    }
    if (on(lcv[c])) {
        vp[c] = f_filter(v[c], vn[c], vp[c], ts);
        v[c] = vn[c];
    }
}

// K3 -- kernel/shallow_water/vel_ssh.f90:269-279
template <bool ROW>
__global__ void __launch_bounds__(BX *BY) k_uv_trans_vort(Geo g, MetRow mr, const float *__restrict__ luu,
        const float *__restrict__ dxt, const float *__restrict__ dyt,
        const float *__restrict__ dxb, const float *__restrict__ dyb,
        const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ vort)
{
    SWCU_CELL(g.nx_start, g.ny_start, g.nx_end, g.ny_end);
    const MetGen mg{nullptr, nullptr, dxt, dyt, nullptr, nullptr, dxb, dyb, nullptr};
    const auto &mt = Pick<ROW>::get(mg, mr);
    const double val = f_vort(c, r, p, mt, u, v);
    if (on(luu[c])) vort[c] = val;
}

// K4 -- kernel/shallow_water/vel_ssh.f90:318-371
template <bool ROW>
__global__ void __launch_bounds__(BX *BY) k_uv_trans(Geo g, MetRow mr,
        const float *__restrict__ lcu, const float *__restrict__ lcv, const float *__restrict__ luu,
        const float *__restrict__ dxh, const float *__restrict__ dyh,
        const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ vort,
        const double *__restrict__ hu, const double *__restrict__ hv, const double *__restrict__ hh,
        double *__restrict__ RHSx, double *__restrict__ RHSy)
{
    SWCU_CELL(g.nx_start, g.ny_start, g.nx_end, g.ny_end);
    const MetGen mg{nullptr, nullptr, nullptr, nullptr, dxh, dyh, nullptr, nullptr, nullptr};
    const auto &mt = Pick<ROW>::get(mg, mr);
    const double rx = f_rhsx_adv(c, r, p, mt, (double)luu[c], (double)luu[c - p], u, v, vort, hu, hv, hh);
    const double ry = f_rhsy_adv(c, r, p, mt, u, v, vort, hu, hv, hh);
    if (on(lcu[c])) RHSx[c] = rx;
    if (on(lcv[c])) RHSy[c] = ry;
}

// K6 -- kernel/shallow_water/vel_ssh.f90:414-450
template <bool ROW>
__global__ void __launch_bounds__(BX *BY) k_uv_diff2(Geo g, MetRow mr,
        const float *__restrict__ lcu, const float *__restrict__ lcv,
        const float *__restrict__ dx, const float *__restrict__ dy,
        const float *__restrict__ dxt, const float *__restrict__ dyt,
        const float *__restrict__ dxh, const float *__restrict__ dyh,
        const float *__restrict__ dxb, const float *__restrict__ dyb,
        const double *__restrict__ mu, const double *__restrict__ str_t, const double *__restrict__ str_s,
        const double *__restrict__ hq, const double *__restrict__ hh,
        double *__restrict__ RHSx, double *__restrict__ RHSy)
{
    SWCU_CELL(g.nx_start, g.ny_start, g.nx_end, g.ny_end);
    const MetGen mg{dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, nullptr};
    const auto &mt = Pick<ROW>::get(mg, mr);
    const double rx = f_rhsx_dif(c, r, p, mt, hq[c], hq[c + 1], mu, str_t, str_s, hh);
    const double ry = f_rhsy_dif(c, r, p, mt, hq[c], hq[c + p], mu, str_t, str_s, hh);
    if (on(lcu[c])) RHSx[c] = rx;
    if (on(lcv[c])) RHSy[c] = ry;
}

// K5 -- kernel/shallow_water/mixing.f90:38-56
template <bool ROW>
__global__ void __launch_bounds__(BX *BY) k_stress_components(Geo g, MetRow mr,
        const float *__restrict__ lu, const float *__restrict__ luu,
        const float *__restrict__ dx, const float *__restrict__ dy,
        const float *__restrict__ dxt, const float *__restrict__ dyt,
        const float *__restrict__ dxh, const float *__restrict__ dyh,
        const float *__restrict__ dxb, const float *__restrict__ dyb,
        const double *__restrict__ u, const double *__restrict__ v,
        double *__restrict__ str_t, double *__restrict__ str_s)
{
    SWCU_CELL(g.nx_start, g.ny_start, g.nx_end, g.ny_end);
    const MetGen mg{dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, nullptr};
    const auto &mt = Pick<ROW>::get(mg, mr);
    const double st = f_str_t(c, r, p, mt, u, v), ss = f_str_s(c, r, p, mt, u, v);
    if (on(lu[c])) str_t[c] = st;
    if (on(luu[c])) str_s[c] = ss;
}

// the three interpolations of hh_init / hh_update for one source depth (depth.f90:57-94);
// the T-point depth of the neighbours is re-evaluated in registers (same expression as the
// whole-array statement) so the kernel needs no grid-wide ordering.
struct Interp3 { double hu, hv, hh; };
template <class M>
__device__ __forceinline__ Interp3 interp3(double q_c, double q_e, double q_n, double q_en, long c, int r, int p,
        const float *__restrict__ lu, const M &mt)
{
    const long e = c + 1, no = c + p, en = c + 1 + p;
    const double dx_c = mt.dx(c, r), dy_c = mt.dy(c, r), dx_e = mt.dx(e, r), dy_e = mt.dy(e, r);
    const double dx_n = mt.dx(no, r + 1), dy_n = mt.dy(no, r + 1), dx_en = mt.dx(en, r + 1), dy_en = mt.dy(en, r + 1);
    const double lu_c = (double)lu[c], lu_e = (double)lu[e], lu_n = (double)lu[no], lu_en = (double)lu[en];
    Interp3 o;
    // dble(lu+lu): the real(4) sum is exact for 0/1 masks and promotes exactly
    const double su = hq_sum2(q_c, q_e, dx_c, dy_c, lu_c, dx_e, dy_e, lu_e);
    const double sv = hq_sum2(q_c, q_n, dx_c, dy_c, lu_c, dx_n, dy_n, lu_n);
    const double sh = q_c * dx_c * dy_c * lu_c + q_e * dx_e * dy_e * lu_e + q_n * dx_n * dy_n * lu_n
                    + q_en * dx_en * dy_en * lu_en;
    o.hu = dv<M>(dv<M>(div_slu_any(su, (double)(lu[c] + lu[e])), mt.dxt(c, r), mt.r_dxt(c, r)), mt.dyh(c, r), mt.r_dyh(c, r));
    o.hv = dv<M>(dv<M>(div_slu_any(sv, (double)(lu[c] + lu[no])), mt.dxh(c, r), mt.r_dxh(c, r)), mt.dyt(c, r), mt.r_dyt(c, r));
    o.hh = dv<M>(dv<M>(div_slu_any(sh, (double)(lu[c] + lu[e] + lu[no] + lu[en])), mt.dxb(c, r), mt.r_dxb(c, r)),
                 mt.dyb(c, r), mt.r_dyb(c, r));
    return o;
}

// K10 -- kernel/shallow_water/depth.f90:48-97
template <bool ROW>
__global__ void __launch_bounds__(BX *BY) k_hh_init(Geo g, MetRow mr, double ffs,
        const float *__restrict__ lu, const float *__restrict__ llu, const float *__restrict__ llv,
        const float *__restrict__ luh,
        const float *__restrict__ dx, const float *__restrict__ dy,
        const float *__restrict__ dxt, const float *__restrict__ dyt,
        const float *__restrict__ dxh, const float *__restrict__ dyh,
        const float *__restrict__ dxb, const float *__restrict__ dyb,
        double *__restrict__ hq, double *__restrict__ hqp, double *__restrict__ hqn,
        double *__restrict__ hu, double *__restrict__ hup, double *__restrict__ hun,
        double *__restrict__ hv, double *__restrict__ hvp, double *__restrict__ hvn,
        double *__restrict__ hh, double *__restrict__ hhp, double *__restrict__ hhn,
        const double *__restrict__ sh, const double *__restrict__ shp, const double *__restrict__ h_r)
{
    SWCU_CELL(g.bx1, g.by1, g.bx2, g.by2);
    const double q = h_r[c] + sh[c] * ffs, qp = h_r[c] + shp[c] * ffs, qn = h_r[c];
    hq[c] = q; hqp[c] = qp; hqn[c] = qn;  // whole-array statements, depth.f90:48-50
    if (m < g.nx_start - 1 || m > g.nx_end || n < g.ny_start - 1 || n > g.ny_end) return;
    const long e = c + 1, no = c + p, en = c + 1 + p;
    const bool wu = on(llu[c]), wv = on(llv[c]), wh = on(luh[c]);   // used for the stores only
    const MetGen mg{dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, nullptr};
    const auto &mt = Pick<ROW>::get(mg, mr);
    const Interp3 i0 = interp3(q, h_r[e] + sh[e] * ffs, h_r[no] + sh[no] * ffs, h_r[en] + sh[en] * ffs, c, r, p, lu, mt);
    const Interp3 ip = interp3(qp, h_r[e] + shp[e] * ffs, h_r[no] + shp[no] * ffs, h_r[en] + shp[en] * ffs, c, r, p, lu, mt);
    const Interp3 in = interp3(qn, h_r[e], h_r[no], h_r[en], c, r, p, lu, mt);
    if (wu) { hu[c] = i0.hu; hup[c] = ip.hu; hun[c] = in.hu; }
    if (wv) { hv[c] = i0.hv; hvp[c] = ip.hv; hvn[c] = in.hv; }
    if (wh) { hh[c] = i0.hh; hhp[c] = ip.hh; hhn[c] = in.hh; }
}

// K2 -- kernel/shallow_water/depth.f90:129-160
template <bool ROW>
__global__ void __launch_bounds__(BX *BY) k_hh_update(Geo g, MetRow mr,
        const float *__restrict__ lu, const float *__restrict__ llu, const float *__restrict__ llv,
        const float *__restrict__ luh,
        const float *__restrict__ dx, const float *__restrict__ dy,
        const float *__restrict__ dxt, const float *__restrict__ dyt,
        const float *__restrict__ dxh, const float *__restrict__ dyh,
        const float *__restrict__ dxb, const float *__restrict__ dyb,
        double *__restrict__ hqn, double *__restrict__ hun, double *__restrict__ hvn, double *__restrict__ hhn,
        const double *__restrict__ sh, const double *__restrict__ h_r)
{
    SWCU_CELL(g.bx1, g.by1, g.bx2, g.by2);
    const double qn = h_r[c] + sh[c];
    hqn[c] = qn;  // depth.f90:129
    if (m < g.nx_start - 1 || m > g.nx_end || n < g.ny_start - 1 || n > g.ny_end) return;
    const long e = c + 1, no = c + p, en = c + 1 + p;
    const bool wu = on(llu[c]), wv = on(llv[c]), wh = on(luh[c]);   // used for the stores only
    const MetGen mg{dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, nullptr};
    const auto &mt = Pick<ROW>::get(mg, mr);
    const Interp3 in = interp3(qn, h_r[e] + sh[e], h_r[no] + sh[no], h_r[en] + sh[en], c, r, p, lu, mt);
    if (wu) hun[c] = in.hu;
    if (wv) hvn[c] = in.hv;
    if (wh) hhn[c] = in.hh;
}

// K9 -- kernel/shallow_water/depth.f90:185-209
__global__ void __launch_bounds__(BX *BY) k_hh_shift(Geo g, double ts,
        const float *__restrict__ lu, const float *__restrict__ llu, const float *__restrict__ llv,
        const float *__restrict__ luh,
        double *__restrict__ hq, double *__restrict__ hqp, const double *__restrict__ hqn,
        double *__restrict__ hu, double *__restrict__ hup, const double *__restrict__ hun,
        double *__restrict__ hv, double *__restrict__ hvp, const double *__restrict__ hvn,
        double *__restrict__ hh, double *__restrict__ hhp, const double *__restrict__ hhn)
{
    SWCU_CELL(g.nx_start - 1, g.ny_start - 1, g.nx_end + 1, g.ny_end + 1);
    if (on(llu[c])) { hup[c] = f_filter(hu[c], hun[c], hup[c], ts); hu[c] = hun[c]; }
    if (on(llv[c])) { hvp[c] = f_filter(hv[c], hvn[c], hvp[c], ts); hv[c] = hvn[c]; }
    if (on(lu[c]))  { hqp[c] = f_filter(hq[c], hqn[c], hqp[c], ts); hq[c] = hqn[c]; }
    if (on(luh[c])) { hhp[c] = f_filter(hh[c], hhn[c], hhp[c], ts); hh[c] = hhn[c]; }
}

// K11 -- kernel/shallow_water/vel_ssh.f90:52-66
__global__ void __launch_bounds__(BX *BY) k_check_ssh_err(Geo g, const float *__restrict__ lu,
        const double *__restrict__ ssh, int *__restrict__ bad)
{
    SWCU_CELL(g.nx_start, g.ny_start, g.nx_end, g.ny_end);
    if (on(lu[c])) {
        const double s = ssh[c];
        if (!(s < 10000.0 && s > -10000.0)) atomicAdd(bad, 1);
    }
}

// kernel/tracer/leapfrog_tracer.f90:55-96
__global__ void __launch_bounds__(BX *BY) k_tran_diff_fluxes(Geo g,
        const float *__restrict__ lcu, const float *__restrict__ lcv,
        const float *__restrict__ dxt, const float *__restrict__ dyt,
        const float *__restrict__ dxh, const float *__restrict__ dyh,
        const double *__restrict__ hhu, const double *__restrict__ hhv, const double *__restrict__ ff,
        const double *__restrict__ uu, const double *__restrict__ vv, const double *__restrict__ mu,
        double factor_mu, double *__restrict__ flux_x, double *__restrict__ flux_y)
{
    SWCU_CELL(g.nx_start, g.ny_start, g.nx_end, g.ny_end);
    const long e = c + 1, no = c + p;
    if (on(lcu[c])) {
        const double dfdx = ff[e] - ff[c];
        const double mu_1d = (mu[c] + mu[e]) / 2.0 * factor_mu * dyh[c] / dxt[c];
        const double flux_diff = mu_1d * hhu[c] * dfdx;
        const double flux_adv = -uu[c] * hhu[c] * dyh[c] * (ff[c] + ff[e]) / 2.0;
        flux_x[c] = flux_adv + flux_diff + 0.0;
    }
    if (on(lcv[c])) {
        const double dfdy = ff[no] - ff[c];
        const double mu_1d = (mu[c] + mu[no]) / 2.0 * factor_mu * dxh[c] / dyt[c];
        const double flux_diff = mu_1d * hhv[c] * dfdy;
        const double flux_adv = -vv[c] * hhv[c] * dxh[c] * (ff[c] + ff[no]) / 2.0;
        flux_y[c] = flux_adv + flux_diff + 0.0;
    }
}

// kernel/tracer/leapfrog_tracer.f90:125-139
__global__ void __launch_bounds__(BX *BY) k_tran_diff_tracer(Geo g, const float *__restrict__ lu,
        const float *__restrict__ dx, const float *__restrict__ dy, double tau,
        const double *__restrict__ hhqn, const double *__restrict__ hhqp,
        const double *__restrict__ flux_x, const double *__restrict__ flux_y,
        const double *__restrict__ ffp, double *__restrict__ ffn)
{
    SWCU_CELL(g.nx_start, g.ny_start, g.nx_end, g.ny_end);
    if (on(lu[c])) {
        const double bp = hhqn[c] * dx[c] * dy[c] / tau / 2.0;
        const double bp0 = hhqp[c] * dx[c] * dy[c] / tau / 2.0;
        const double rhs = flux_x[c] - flux_x[c - 1] + flux_y[c] - flux_y[c - p];
        const double eta = bp0 * ffp[c] + rhs;
        ffn[c] = eta / bp;
    }
}

// kernel/tracer/leapfrog_tracer.f90:159-168
__global__ void __launch_bounds__(BX *BY) k_tracer_next_step(Geo g, double ts, const float *__restrict__ lu,
        const double *__restrict__ ffn, double *__restrict__ ffp, double *__restrict__ ff)
{
    SWCU_CELL(g.nx_start - 1, g.ny_start - 1, g.nx_end + 1, g.ny_end + 1);
    if (on(lu[c])) {
        ffp[c] = f_filter(ff[c], ffn[c], ffp[c], ts);
        ff[c] = ffn[c];
    }
}

const dim3 kBlock(BX, BY, 1);

}  // namespace

#define GRID_S grid_for(g.nx_start, g.nx_end, g.ny_start, g.ny_end)
#define GRID_SP grid_for(g.nx_start - 1, g.nx_end + 1, g.ny_start - 1, g.ny_end + 1)
#define GRID_ALL grid_for(g.bx1, g.bx2, g.by1, g.by2)

int launch_sw_update_ssh(const Geo &g, double tau, const float *lu, const float *dx, const float *dy,
        const float *dxh, const float *dyh, const double *hhu, const double *hhv, double *sshn,
        const double *sshp, const double *u, const double *v, cudaStream_t st, const MetRow *mr)
{
    SWCU_LAUNCH_ROW(k_sw_update_ssh, GRID_S, tau, lu, dx, dy, dxh, dyh, hhu, hhv, sshn, sshp, u, v);
    return launched("sw_update_ssh");
}

int launch_sw_update_uv(const Geo &g, double tau, const float *lcu, const float *lcv,
        const float *dxt, const float *dyt, const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        const double *hhu, const double *hhun, const double *hhup,
        const double *hhv, const double *hhvn, const double *hhvp, const double *hhh, const double *ssh,
        const double *u, double *un, const double *up, const double *v, double *vn, const double *vp,
        const float *rdis, const float *rlh_s, const double *RHSx, const double *RHSy,
        const double *RHSx_adv, const double *RHSy_adv, const double *RHSx_dif, const double *RHSy_dif,
        cudaStream_t st, const MetRow *mr, const Tau *tau_exact)
{
    const Tau tt = tau_exact ? *tau_exact : Tau{tau, 0.0, 0, 0};  // plain Level A keeps the hardware division
    SWCU_LAUNCH_ROW(k_sw_update_uv, GRID_S, tt, lcu, lcv, dxt, dyt, dxh, dyh, dxb, dyb,
            hhu, hhun, hhup, hhv, hhvn, hhvp, hhh, ssh, u, un, up, v, vn, vp, rdis, rlh_s,
            RHSx, RHSy, RHSx_adv, RHSy_adv, RHSx_dif, RHSy_dif);
    return launched("sw_update_uv");
}

int launch_sw_next_step(const Geo &g, double ts, const float *lu, const float *lcu, const float *lcv,
        double *ssh, double *sshn, double *sshp, double *u, double *un, double *up,
        double *v, double *vn, double *vp, cudaStream_t st)
{
    k_sw_next_step<<<GRID_SP, kBlock, 0, st>>>(g, ts, lu, lcu, lcv, ssh, sshn, sshp, u, un, up, v, vn, vp);
    return launched("sw_next_step");
}

int launch_uv_trans_vort(const Geo &g, const float *luu, const float *dxt, const float *dyt,
        const float *dxb, const float *dyb, const double *u, const double *v, double *vort, cudaStream_t st, const MetRow *mr)
{
    SWCU_LAUNCH_ROW(k_uv_trans_vort, GRID_S, luu, dxt, dyt, dxb, dyb, u, v, vort);
    return launched("uv_trans_vort");
}

int launch_uv_trans(const Geo &g, const float *lcu, const float *lcv, const float *luu,
        const float *dxh, const float *dyh, const double *u, const double *v, const double *vort,
        const double *hu, const double *hv, const double *hh, double *RHSx, double *RHSy, cudaStream_t st, const MetRow *mr)
{
    SWCU_LAUNCH_ROW(k_uv_trans, GRID_S, lcu, lcv, luu, dxh, dyh, u, v, vort, hu, hv, hh, RHSx, RHSy);
    return launched("uv_trans");
}

int launch_uv_diff2(const Geo &g, const float *lcu, const float *lcv,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        const double *mu, const double *str_t, const double *str_s, const double *hq, const double *hh,
        double *RHSx, double *RHSy, cudaStream_t st, const MetRow *mr)
{
    SWCU_LAUNCH_ROW(k_uv_diff2, GRID_S, lcu, lcv, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb,
                                          mu, str_t, str_s, hq, hh, RHSx, RHSy);
    return launched("uv_diff2");
}

int launch_stress_components(const Geo &g, const float *lu, const float *luu,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        const double *u, const double *v, double *str_t, double *str_s, cudaStream_t st, const MetRow *mr)
{
    SWCU_LAUNCH_ROW(k_stress_components, GRID_S, lu, luu, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb,
                                                   u, v, str_t, str_s);
    return launched("stress_components");
}

int launch_hh_init(const Geo &g, int ffs, const float *lu, const float *llu, const float *llv, const float *luh,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        double *hq, double *hqp, double *hqn, double *hu, double *hup, double *hun,
        double *hv, double *hvp, double *hvn, double *hh, double *hhp, double *hhn,
        const double *sh, const double *shp, const double *h_r, cudaStream_t st, const MetRow *mr)
{
    SWCU_LAUNCH_ROW(k_hh_init, GRID_ALL, (double)ffs, lu, llu, llv, luh, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb,
            hq, hqp, hqn, hu, hup, hun, hv, hvp, hvn, hh, hhp, hhn, sh, shp, h_r);
    return launched("hh_init");
}

int launch_hh_update(const Geo &g, const float *lu, const float *llu, const float *llv, const float *luh,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        double *hqn, double *hun, double *hvn, double *hhn, const double *sh, const double *h_r, cudaStream_t st, const MetRow *mr)
{
    SWCU_LAUNCH_ROW(k_hh_update, GRID_ALL, lu, llu, llv, luh, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb,
                                             hqn, hun, hvn, hhn, sh, h_r);
    return launched("hh_update");
}

int launch_hh_shift(const Geo &g, double ts, const float *lu, const float *llu, const float *llv, const float *luh,
        double *hq, double *hqp, double *hqn, double *hu, double *hup, double *hun,
        double *hv, double *hvp, double *hvn, double *hh, double *hhp, double *hhn, cudaStream_t st)
{
    k_hh_shift<<<GRID_SP, kBlock, 0, st>>>(g, ts, lu, llu, llv, luh, hq, hqp, hqn, hu, hup, hun,
                                           hv, hvp, hvn, hh, hhp, hhn);
    return launched("hh_shift");
}

int launch_check_ssh_err(const Geo &g, const float *lu, const double *ssh, int *bad, cudaStream_t st)
{
    k_check_ssh_err<<<GRID_S, kBlock, 0, st>>>(g, lu, ssh, bad);
    return launched("check_ssh_err");
}

int launch_tran_diff_fluxes(const Geo &g, const float *lcu, const float *lcv,
        const float *dxt, const float *dyt, const float *dxh, const float *dyh,
        const double *hhu, const double *hhv, const double *ff, const double *uu, const double *vv,
        const double *mu, double factor_mu, double *flux_x, double *flux_y, cudaStream_t st)
{
    k_tran_diff_fluxes<<<GRID_S, kBlock, 0, st>>>(g, lcu, lcv, dxt, dyt, dxh, dyh, hhu, hhv, ff, uu, vv, mu,
                                                  factor_mu, flux_x, flux_y);
    return launched("tran_diff_fluxes");
}

int launch_tran_diff_tracer(const Geo &g, const float *lu, const float *dx, const float *dy, double tau,
        const double *hhqn, const double *hhqp, const double *flux_x, const double *flux_y,
        const double *ffp, double *ffn, cudaStream_t st)
{
    k_tran_diff_tracer<<<GRID_S, kBlock, 0, st>>>(g, lu, dx, dy, tau, hhqn, hhqp, flux_x, flux_y, ffp, ffn);
    return launched("tran_diff_tracer");
}

int launch_tracer_next_step(const Geo &g, double ts, const float *lu, const double *ffn, double *ffp, double *ff,
        cudaStream_t st)
{
    k_tracer_next_step<<<GRID_SP, kBlock, 0, st>>>(g, ts, lu, ffn, ffp, ff);
    return launched("tracer_next_step");
}

}  // namespace swcu
