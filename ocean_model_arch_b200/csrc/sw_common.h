// sw_common.h -- host-side helpers shared by the translation units of libswcuda.so
#pragma once
#include <cuda_runtime.h>

#include "../../include/swcuda.h"
#include "sw_formulas.cuh"

namespace swcu {

// Records the message returned by swcu_last_error() (thread-local).
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define SWCU_CUDA(call)                                              \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return ::swcu::cuda_fail(e__, #call); \
    } while (0)

inline Geo make_geo(const swcu_dims &d, int pitch)
{
    Geo g;
    g.nx_start = d.nx_start; g.nx_end = d.nx_end; g.ny_start = d.ny_start; g.ny_end = d.ny_end;
    g.bx1 = d.bnd_x1; g.bx2 = d.bnd_x2; g.by1 = d.bnd_y1; g.by2 = d.bnd_y2;
    g.pitch = pitch;
    return g;
}
inline int width(const swcu_dims &d) { return d.bnd_x2 - d.bnd_x1 + 1; }
inline int height(const swcu_dims &d) { return d.bnd_y2 - d.bnd_y1 + 1; }
int check_dims(const swcu_dims *d);

// ---- launchers of the 1:1 kernels (sw_kernels_ref.cu).  All return SWCU_OK / SWCU_ERR_CUDA. ----
int launch_sw_update_ssh(const Geo &g, double tau, const float *lu, const float *dx, const float *dy,
        const float *dxh, const float *dyh, const double *hhu, const double *hhv, double *sshn,
        const double *sshp, const double *u, const double *v, cudaStream_t st, const MetRow *mr = nullptr);
int launch_sw_update_uv(const Geo &g, double tau, const float *lcu, const float *lcv,
        const float *dxt, const float *dyt, const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        const double *hhu, const double *hhun, const double *hhup,
        const double *hhv, const double *hhvn, const double *hhvp, const double *hhh, const double *ssh,
        const double *u, double *un, const double *up, const double *v, double *vn, const double *vp,
        const float *rdis, const float *rlh_s, const double *RHSx, const double *RHSy,
        const double *RHSx_adv, const double *RHSy_adv, const double *RHSx_dif, const double *RHSy_dif,
        cudaStream_t st, const MetRow *mr = nullptr, const Tau *tau_exact = nullptr);
int launch_sw_next_step(const Geo &g, double ts, const float *lu, const float *lcu, const float *lcv,
        double *ssh, double *sshn, double *sshp, double *u, double *un, double *up,
        double *v, double *vn, double *vp, cudaStream_t st);
int launch_uv_trans_vort(const Geo &g, const float *luu, const float *dxt, const float *dyt,
        const float *dxb, const float *dyb, const double *u, const double *v, double *vort, cudaStream_t st, const MetRow *mr = nullptr);
int launch_uv_trans(const Geo &g, const float *lcu, const float *lcv, const float *luu,
        const float *dxh, const float *dyh, const double *u, const double *v, const double *vort,
        const double *hu, const double *hv, const double *hh, double *RHSx, double *RHSy, cudaStream_t st, const MetRow *mr = nullptr);
int launch_uv_diff2(const Geo &g, const float *lcu, const float *lcv,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        const double *mu, const double *str_t, const double *str_s, const double *hq, const double *hh,
        double *RHSx, double *RHSy, cudaStream_t st, const MetRow *mr = nullptr);
int launch_stress_components(const Geo &g, const float *lu, const float *luu,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        const double *u, const double *v, double *str_t, double *str_s, cudaStream_t st, const MetRow *mr = nullptr);
int launch_hh_init(const Geo &g, int ffs, const float *lu, const float *llu, const float *llv, const float *luh,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        double *hq, double *hqp, double *hqn, double *hu, double *hup, double *hun,
        double *hv, double *hvp, double *hvn, double *hh, double *hhp, double *hhn,
        const double *sh, const double *shp, const double *h_r, cudaStream_t st, const MetRow *mr = nullptr);
int launch_hh_update(const Geo &g, const float *lu, const float *llu, const float *llv, const float *luh,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        double *hqn, double *hun, double *hvn, double *hhn, const double *sh, const double *h_r, cudaStream_t st, const MetRow *mr = nullptr);
int launch_hh_shift(const Geo &g, double ts, const float *lu, const float *llu, const float *llv, const float *luh,
        double *hq, double *hqp, double *hqn, double *hu, double *hup, double *hun,
        double *hv, double *hvp, double *hvn, double *hh, double *hhp, double *hhn, cudaStream_t st);
int launch_check_ssh_err(const Geo &g, const float *lu, const double *ssh, int *bad, cudaStream_t st);
int launch_tran_diff_fluxes(const Geo &g, const float *lcu, const float *lcv,
        const float *dxt, const float *dyt, const float *dxh, const float *dyh,
        const double *hhu, const double *hhv, const double *ff, const double *uu, const double *vv,
        const double *mu, double factor_mu, double *flux_x, double *flux_y, cudaStream_t st);
int launch_tran_diff_tracer(const Geo &g, const float *lu, const float *dx, const float *dy, double tau,
        const double *hhqn, const double *hhqp, const double *flux_x, const double *flux_y,
        const double *ffp, double *ffn, cudaStream_t st);
int launch_tracer_next_step(const Geo &g, double ts, const float *lu, const double *ffn, double *ffp, double *ff,
        cudaStream_t st);

}  // namespace swcu
