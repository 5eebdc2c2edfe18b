// sw_fused.h -- argument block and launchers of the fused Level-B kernels (sw_kernels_fused.cu)
#pragma once
#include "sw_common.h"

namespace swcu {

// bit <=> the reference's real(4) mask of the same name is > 0.5 (core/grid.f90:24-31)
enum : int { MB_LU = 1, MB_LCU = 2, MB_LCV = 4, MB_LUU = 8, MB_LUH = 16, MB_LLU = 32, MB_LLV = 64 };

struct FusedArgs {
    // time-level-n state (read) and n+1 state (written): ping-pong buffers
    const double *ssh, *sshp, *u, *up, *v, *vp;
    double *ssh_o, *sshp_o, *u_o, *up_o, *v_o, *vp_o;
    // static real(8)
    const double *h_r, *mu;
    const double *RHSx, *RHSy;  // nullptr <=> identically zero (the reference never assigns them)
    // scratch written by prep, read by update
    double *hu, *hv, *hh, *vort, *str_t, *str_s;
    // static real(4)
    const float *dx, *dy, *dxt, *dyt, *dxh, *dyh, *dxb, *dyb, *rlh_s;
    const float *rdis;  // nullptr <=> identically zero
    const unsigned char *mask;
    int *bad;  // K11 counter
    double tau, ts, ffs;
    int trans, lat;
};

// prep on rows [n0..n1] (columns nx_start-1 .. nx_end+1); update on rows [n0..n1] (columns of S)
int launch_prep(const Geo &g, const FusedArgs &a, int n0, int n1, cudaStream_t st);
int launch_update(const Geo &g, const FusedArgs &a, int n0, int n1, cudaStream_t st);
int launch_mask_set(long total, const float *src, unsigned char *bits, int bit, cudaStream_t st);
int launch_mask_get(long total, float *dst, const unsigned char *bits, int bit, cudaStream_t st);

}  // namespace swcu
