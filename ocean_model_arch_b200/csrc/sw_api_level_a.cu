// sw_api_level_a.cu -- extern "C" Level-A entry points (include/swcuda.h): the reference's
// per-block kernel calls on device pointers, reference argument order.
#include <cstdarg>
#include <cstdio>

#include "sw_common.h"

namespace swcu {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return SWCU_ERR_CUDA;
}

int check_dims(const swcu_dims *d)
{
    if (!d) { set_error("dims is NULL"); return SWCU_ERR_ARG; }
    if (d->nx_end < d->nx_start || d->ny_end < d->ny_start ||
        d->bnd_x1 > d->nx_start - 2 || d->bnd_x2 < d->nx_end + 2 ||
        d->bnd_y1 > d->ny_start - 2 || d->bnd_y2 < d->ny_end + 2) {
        set_error("bad dims: interior [%d..%d]x[%d..%d] needs a 2-cell border inside [%d..%d]x[%d..%d]",
                  d->nx_start, d->nx_end, d->ny_start, d->ny_end, d->bnd_x1, d->bnd_x2, d->bnd_y1, d->bnd_y2);
        return SWCU_ERR_ARG;
    }
    return SWCU_OK;
}

}  // namespace swcu

using namespace swcu;

#define GEO_OR_RETURN                                  \
    if (int rc__ = check_dims(d)) return rc__;         \
    const Geo g = make_geo(*d, width(*d));             \
    cudaStream_t st = (cudaStream_t)stream

extern "C" {

const char *swcu_last_error(void) { return swcu::g_err; }
int swcu_version(void) { return 100; }
int swcu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int swcu_sw_update_ssh_kernel(const swcu_dims *d, double tau,
        const float *lu, const float *dx, const float *dy, const float *dxh, const float *dyh,
        const double *hhu, const double *hhv, double *sshn, const double *sshp,
        const double *ubrtr, const double *vbrtr, void *stream)
{
    GEO_OR_RETURN;
    return launch_sw_update_ssh(g, tau, lu, dx, dy, dxh, dyh, hhu, hhv, sshn, sshp, ubrtr, vbrtr, st);
}

int swcu_sw_update_uv(const swcu_dims *d, double tau, const float *lcu, const float *lcv,
        const float *dxt, const float *dyt, const float *dxh, const float *dyh,
        const float *dxb, const float *dyb,
        const double *hhu, const double *hhun, const double *hhup,
        const double *hhv, const double *hhvn, const double *hhvp,
        const double *hhh, const double *ssh,
        const double *ubrtr, double *ubrtrn, const double *ubrtrp,
        const double *vbrtr, double *vbrtrn, const double *vbrtrp,
        const float *rdis, const float *rlh_s,
        const double *RHSx, const double *RHSy, const double *RHSx_adv, const double *RHSy_adv,
        const double *RHSx_dif, const double *RHSy_dif, void *stream)
{
    GEO_OR_RETURN;
    return launch_sw_update_uv(g, tau, lcu, lcv, dxt, dyt, dxh, dyh, dxb, dyb, hhu, hhun, hhup, hhv, hhvn, hhvp,
                               hhh, ssh, ubrtr, ubrtrn, ubrtrp, vbrtr, vbrtrn, vbrtrp, rdis, rlh_s,
                               RHSx, RHSy, RHSx_adv, RHSy_adv, RHSx_dif, RHSy_dif, st);
}

int swcu_sw_next_step(const swcu_dims *d, double time_smooth,
        const float *lu, const float *lcu, const float *lcv,
        double *ssh, double *sshn, double *sshp,
        double *ubrtr, double *ubrtrn, double *ubrtrp,
        double *vbrtr, double *vbrtrn, double *vbrtrp, void *stream)
{
    GEO_OR_RETURN;
    return launch_sw_next_step(g, time_smooth, lu, lcu, lcv, ssh, sshn, sshp, ubrtr, ubrtrn, ubrtrp,
                               vbrtr, vbrtrn, vbrtrp, st);
}

int swcu_uv_trans_vort_kernel(const swcu_dims *d, const float *luu,
        const float *dxt, const float *dyt, const float *dxb, const float *dyb,
        const double *u, const double *v, double *vort, void *stream)
{
    GEO_OR_RETURN;
    return launch_uv_trans_vort(g, luu, dxt, dyt, dxb, dyb, u, v, vort, st);
}

int swcu_uv_trans_kernel(const swcu_dims *d, const float *lcu, const float *lcv, const float *luu,
        const float *dxh, const float *dyh, const double *u, const double *v, const double *vort,
        const double *hq, const double *hu, const double *hv, const double *hh,
        double *RHSx, double *RHSy, void *stream)
{
    (void)hq;
    GEO_OR_RETURN;
    return launch_uv_trans(g, lcu, lcv, luu, dxh, dyh, u, v, vort, hu, hv, hh, RHSx, RHSy, st);
}

int swcu_uv_diff2_kernel(const swcu_dims *d, const float *lcu, const float *lcv,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        const double *mu, const double *str_t, const double *str_s,
        const double *hq, const double *hu, const double *hv, const double *hh,
        double *RHSx, double *RHSy, void *stream)
{
    (void)hu; (void)hv;
    GEO_OR_RETURN;
    return launch_uv_diff2(g, lcu, lcv, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, mu, str_t, str_s, hq, hh,
                           RHSx, RHSy, st);
}

int swcu_stress_components_kernel(const swcu_dims *d, const float *lu, const float *luu,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        const double *u, const double *v, double *str_t, double *str_s, void *stream)
{
    GEO_OR_RETURN;
    return launch_stress_components(g, lu, luu, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, u, v, str_t, str_s, st);
}

int swcu_hh_init_kernel(const swcu_dims *d, int full_free_surface,
        const float *lu, const float *llu, const float *llv, const float *luh,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        double *hq, double *hqp, double *hqn, double *hu, double *hup, double *hun,
        double *hv, double *hvp, double *hvn, double *hh, double *hhp, double *hhn,
        const double *sh, const double *shp, const double *h_r, void *stream)
{
    GEO_OR_RETURN;
    return launch_hh_init(g, full_free_surface, lu, llu, llv, luh, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb,
                          hq, hqp, hqn, hu, hup, hun, hv, hvp, hvn, hh, hhp, hhn, sh, shp, h_r, st);
}

int swcu_hh_update_kernel(const swcu_dims *d,
        const float *lu, const float *llu, const float *llv, const float *luh,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        double *hqn, double *hun, double *hvn, double *hhn,
        const double *sh, const double *h_r, void *stream)
{
    GEO_OR_RETURN;
    return launch_hh_update(g, lu, llu, llv, luh, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, hqn, hun, hvn, hhn,
                            sh, h_r, st);
}

int swcu_hh_shift_kernel(const swcu_dims *d, double time_smooth,
        const float *lu, const float *llu, const float *llv, const float *luh,
        double *hq, double *hqp, double *hqn, double *hu, double *hup, double *hun,
        double *hv, double *hvp, double *hvn, double *hh, double *hhp, double *hhn, void *stream)
{
    GEO_OR_RETURN;
    return launch_hh_shift(g, time_smooth, lu, llu, llv, luh, hq, hqp, hqn, hu, hup, hun, hv, hvp, hvn,
                           hh, hhp, hhn, st);
}

int swcu_check_ssh_err_kernel(const swcu_dims *d, const float *lu, const double *ssh,
        int *bad_count, void *stream)
{
    GEO_OR_RETURN;
    return launch_check_ssh_err(g, lu, ssh, bad_count, st);
}

int swcu_tran_diff_fluxes_kernel(const swcu_dims *d, const float *lcu, const float *lcv,
        const float *dxt, const float *dyt, const float *dxh, const float *dyh,
        const double *hhu, const double *hhv, const double *ff, const double *ffp,
        const double *uu, const double *vv, const double *mu, double factor_mu,
        double *flux_x, double *flux_y, void *stream)
{
    (void)ffp;
    GEO_OR_RETURN;
    return launch_tran_diff_fluxes(g, lcu, lcv, dxt, dyt, dxh, dyh, hhu, hhv, ff, uu, vv, mu, factor_mu,
                                   flux_x, flux_y, st);
}

int swcu_tran_diff_tracer_kernel(const swcu_dims *d, const float *lu, const float *dx, const float *dy,
        double tau, const double *hhqn, const double *hhqp,
        const double *flux_x, const double *flux_y, const double *ffp, double *ffn, void *stream)
{
    GEO_OR_RETURN;
    return launch_tran_diff_tracer(g, lu, dx, dy, tau, hhqn, hhqp, flux_x, flux_y, ffp, ffn, st);
}

int swcu_tracer_next_step_kernel(const swcu_dims *d, double time_smooth, const float *lu,
        const double *ffn, double *ffp, double *ff, void *stream)
{
    GEO_OR_RETURN;
    return launch_tracer_next_step(g, time_smooth, lu, ffn, ffp, ff, st);
}

}  // extern "C"
