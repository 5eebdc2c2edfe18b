// sw_kernels_fused.cu -- Level B fused step: the reference's K1..K11 collapsed into two launches.
//
//   prep   (A): from the time-level-n state (ssh, u, v, up, vp) and the static fields, the
//               neighbour-shared intermediates: depths on U/V/H points (K10/K2: depth.f90:57-94),
//               vorticity (K3: vel_ssh.f90:273-275) and stress components (K5: mixing.f90:43-51).
//               Range: interior grown by one cell (their consumers read them at +-1).
//   update (B): K1 (ssh), K4 (transport), K6 (diffusion), K7 (u, v), K8 (Asselin filter + level
//               rotation), K11 (blow-up guard) for one cell, writing the n+1 state into the
//               ping-pong buffers.  hhu_p / hhv_p (K10 with sshp) are evaluated in registers.
//
// Facts from the reference that make this legal:
//   - K1 and K7 read only level-n fields (sw_interface.f90:326-327, :359);
//   - K2's hhq_n = hhq_rest + ssh is bitwise hhq (K10 uses ssh*dfloat(1)), so hhu_n == hhu etc.;
//   - K9's outputs are overwritten by K10 wherever they are read;
//   - every depth field is a pure function of (hhq_rest, ssh, sshp, masks, metrics).
// Masks arrive packed one byte per cell (bit set <=> reference mask > 0.5); stores are selects,
// never multiplications by the mask (masked lanes may hold Inf/NaN).
// Arithmetic is the shared sw_formulas.cuh, so results are bitwise those of the 1:1 kernels.
#include "sw_fused.h"

namespace swcu {

namespace {

constexpr int BX = 64;
constexpr int BY = 4;

__device__ __forceinline__ float mf(unsigned char bits, int bit) { return (bits & bit) ? 1.0f : 0.0f; }

template <bool TRANS, bool LAT>
__global__ void __launch_bounds__(BX *BY) k_prep(Geo g, FusedArgs a, int n0, int n1)
{
    const int m = g.nx_start - 1 + blockIdx.x * BX + threadIdx.x;
    const int n = n0 + blockIdx.y * BY + threadIdx.y;
    if (m > g.nx_end + 1 || n > n1) return;
    const long c = ix(g, m, n);
    const int p = g.pitch;
    const long e = c + 1, no = c + p, en = c + 1 + p;
    const unsigned char mb = a.mask[c];
    const double ffs = a.ffs;

    // K10: depth.f90:48, 57-94 (hq = h_r + sh*ffs re-evaluated per neighbour)
    const double q_c = a.h_r[c] + a.ssh[c] * ffs, q_e = a.h_r[e] + a.ssh[e] * ffs;
    const double q_n = a.h_r[no] + a.ssh[no] * ffs, q_en = a.h_r[en] + a.ssh[en] * ffs;
    const float lu_c = mf(mb, MB_LU), lu_e = mf(a.mask[e], MB_LU);
    const float lu_n = mf(a.mask[no], MB_LU), lu_en = mf(a.mask[en], MB_LU);
    const float dx_c = a.dx[c], dy_c = a.dy[c], dx_e = a.dx[e], dy_e = a.dy[e];
    const float dx_n = a.dx[no], dy_n = a.dy[no], dx_en = a.dx[en], dy_en = a.dy[en];
    double hu = 0.0, hv = 0.0, hh = 0.0;
    if (mb & MB_LLU) hu = f_interp2(q_c, q_e, dx_c, dy_c, lu_c, dx_e, dy_e, lu_e, a.dxt[c], a.dyh[c]);
    if (mb & MB_LLV) hv = f_interp2(q_c, q_n, dx_c, dy_c, lu_c, dx_n, dy_n, lu_n, a.dxh[c], a.dyt[c]);
    if (mb & MB_LUH)
        hh = f_interp4(q_c, q_e, q_n, q_en, dx_c, dy_c, lu_c, dx_e, dy_e, lu_e, dx_n, dy_n, lu_n,
                       dx_en, dy_en, lu_en, a.dxb[c], a.dyb[c]);
    a.hu[c] = hu; a.hv[c] = hv; a.hh[c] = hh;

    if (TRANS) a.vort[c] = (mb & MB_LUU) ? f_vort(c, p, a.dxt, a.dyt, a.dxb, a.dyb, a.u, a.v) : 0.0;
    if (LAT) {
        a.str_t[c] = (mb & MB_LU) ? f_str_t(c, p, a.dx, a.dy, a.dxh, a.dyh, a.up, a.vp) : 0.0;
        a.str_s[c] = (mb & MB_LUU) ? f_str_s(c, p, a.dxt, a.dyt, a.dxb, a.dyb, a.up, a.vp) : 0.0;
    }
}

template <bool TRANS, bool LAT, bool HAS_RHS, bool HAS_RDISS>
__global__ void __launch_bounds__(BX *BY) k_update(Geo g, FusedArgs a, int n0, int n1)
{
    const int m = g.nx_start + blockIdx.x * BX + threadIdx.x;
    const int n = n0 + blockIdx.y * BY + threadIdx.y;
    if (m > g.nx_end || n > n1) return;
    const long c = ix(g, m, n);
    const int p = g.pitch;
    const long e = c + 1, no = c + p;
    const unsigned char mb = a.mask[c];
    const double tau = a.tau, ts = a.ts, ffs = a.ffs;

    const double ssh_c = a.ssh[c], sshp_c = a.sshp[c];
    const double u_c = a.u[c], up_c = a.up[c], v_c = a.v[c], vp_c = a.vp[c];

    // K1
    double ssh_new = ssh_c, sshp_new = sshp_c;
    if (mb & MB_LU) {
        const double sshn = f_sshn(c, p, tau, a.dx, a.dy, a.dxh, a.dyh, a.hu, a.hv, a.sshp, a.u, a.v);
        sshp_new = f_filter(ssh_c, sshn, sshp_c, ts);  // K8, vel_ssh.f90:230-231
        ssh_new = sshn;
        if (!(sshn < 10000.0 && sshn > -10000.0)) atomicAdd(a.bad, 1);  // K11, vel_ssh.f90:55
    }
    a.ssh_o[c] = ssh_new; a.sshp_o[c] = sshp_new;

    double u_new = u_c, up_new = up_c, v_new = v_c, vp_new = vp_c;
    if (mb & (MB_LCU | MB_LCV)) {
        const double h_c = a.h_r[c];
        const double q_c = h_c + ssh_c * ffs, qp_c = h_c + sshp_c * ffs;
        const float dx_c = a.dx[c], dy_c = a.dy[c];
        const float lu_c = mf(mb, MB_LU);
        if (mb & MB_LCU) {
            const double h_e = a.h_r[e];
            const double q_e = h_e + a.ssh[e] * ffs, qp_e = h_e + a.sshp[e] * ffs;
            const double hu_c = a.hu[c];
            const double hup_c = f_interp2(qp_c, qp_e, dx_c, dy_c, lu_c, a.dx[e], a.dy[e], mf(a.mask[e], MB_LU),
                                           a.dxt[c], a.dyh[c]);  // K10 with shp, depth.f90:62-63
            const double adv = TRANS ? f_rhsx_adv(c, p, mf(mb, MB_LUU), mf(a.mask[c - p], MB_LUU), a.dxh, a.dyh, a.u, a.v, a.vort, a.hu, a.hv, a.hh) : 0.0;
            const double dif = LAT ? f_rhsx_dif(c, p, q_c, q_e, a.dy, a.dxt, a.dyh, a.dxb, a.mu, a.str_t, a.str_s, a.hh) : 0.0;
            const double rhs = HAS_RHS ? a.RHSx[c] : 0.0;
            const float rd = HAS_RDISS ? a.rdis[c] + a.rdis[e] : 0.0f + 0.0f;
            const double un = f_un(c, p, tau, hu_c, hu_c, hup_c, rhs, dif, adv, rd, a.dxt, a.dyh, a.dxb, a.dyb,
                                   a.rlh_s, a.hh, a.ssh, a.v, a.up);
            up_new = f_filter(u_c, un, up_c, ts);
            u_new = un;
        }
        if (mb & MB_LCV) {
            const double h_n = a.h_r[no];
            const double q_n = h_n + a.ssh[no] * ffs, qp_n = h_n + a.sshp[no] * ffs;
            const double hv_c = a.hv[c];
            const double hvp_c = f_interp2(qp_c, qp_n, dx_c, dy_c, lu_c, a.dx[no], a.dy[no], mf(a.mask[no], MB_LU),
                                           a.dxh[c], a.dyt[c]);  // depth.f90:73-74
            const double adv = TRANS ? f_rhsy_adv(c, p, a.dxh, a.dyh, a.u, a.v, a.vort, a.hu, a.hv, a.hh) : 0.0;
            const double dif = LAT ? f_rhsy_dif(c, p, q_c, q_n, a.dx, a.dyt, a.dxh, a.dyb, a.mu, a.str_t, a.str_s, a.hh) : 0.0;
            const double rhs = HAS_RHS ? a.RHSy[c] : 0.0;
            const float rd = HAS_RDISS ? a.rdis[c] + a.rdis[no] : 0.0f + 0.0f;
            const double vn = f_vn(c, p, tau, hv_c, hv_c, hvp_c, rhs, dif, adv, rd, a.dyt, a.dxh, a.dxb, a.dyb,
                                   a.rlh_s, a.hh, a.ssh, a.u, a.vp);
            vp_new = f_filter(v_c, vn, vp_c, ts);
            v_new = vn;
        }
    }
    a.u_o[c] = u_new; a.up_o[c] = up_new; a.v_o[c] = v_new; a.vp_o[c] = vp_new;
}

}  // namespace

int launch_prep(const Geo &g, const FusedArgs &a, int n0, int n1, cudaStream_t st)
{
    if (n1 < n0) return SWCU_OK;
    const dim3 block(BX, BY, 1);
    const dim3 grid((unsigned)((g.nx_end - g.nx_start + 2 + BX) / BX), (unsigned)((n1 - n0 + BY) / BY), 1);
    if (a.trans && a.lat) k_prep<true, true><<<grid, block, 0, st>>>(g, a, n0, n1);
    else if (a.trans) k_prep<true, false><<<grid, block, 0, st>>>(g, a, n0, n1);
    else if (a.lat) k_prep<false, true><<<grid, block, 0, st>>>(g, a, n0, n1);
    else k_prep<false, false><<<grid, block, 0, st>>>(g, a, n0, n1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SWCU_OK : cuda_fail(e, "prep");
}

template <bool T, bool L>
static void update_dispatch(const Geo &g, const FusedArgs &a, int n0, int n1, dim3 grid, dim3 block, cudaStream_t st)
{
    const bool r = a.RHSx != nullptr, d = a.rdis != nullptr;
    if (r && d) k_update<T, L, true, true><<<grid, block, 0, st>>>(g, a, n0, n1);
    else if (r) k_update<T, L, true, false><<<grid, block, 0, st>>>(g, a, n0, n1);
    else if (d) k_update<T, L, false, true><<<grid, block, 0, st>>>(g, a, n0, n1);
    else k_update<T, L, false, false><<<grid, block, 0, st>>>(g, a, n0, n1);
}

int launch_update(const Geo &g, const FusedArgs &a, int n0, int n1, cudaStream_t st)
{
    if (n1 < n0) return SWCU_OK;
    const dim3 block(BX, BY, 1);
    const dim3 grid((unsigned)((g.nx_end - g.nx_start + BX) / BX), (unsigned)((n1 - n0 + BY) / BY), 1);
    if (a.trans && a.lat) update_dispatch<true, true>(g, a, n0, n1, grid, block, st);
    else if (a.trans) update_dispatch<true, false>(g, a, n0, n1, grid, block, st);
    else if (a.lat) update_dispatch<false, true>(g, a, n0, n1, grid, block, st);
    else update_dispatch<false, false>(g, a, n0, n1, grid, block, st);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SWCU_OK : cuda_fail(e, "update");
}

// ---- mask packing -----------------------------------------------------------------------------
namespace {
__global__ void k_mask_set(long total, const float *__restrict__ src, unsigned char *__restrict__ bits, int bit)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const unsigned char b = bits[i];
    bits[i] = src[i] > 0.5f ? (unsigned char)(b | bit) : (unsigned char)(b & ~bit);
}
__global__ void k_mask_get(long total, float *__restrict__ dst, const unsigned char *__restrict__ bits, int bit)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    dst[i] = (bits[i] & bit) ? 1.0f : 0.0f;
}
}  // namespace

int launch_mask_set(long total, const float *src, unsigned char *bits, int bit, cudaStream_t st)
{
    k_mask_set<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, src, bits, bit);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SWCU_OK : cuda_fail(e, "mask_set");
}
int launch_mask_get(long total, float *dst, const unsigned char *bits, int bit, cudaStream_t st)
{
    k_mask_get<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, dst, bits, bit);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SWCU_OK : cuda_fail(e, "mask_get");
}

}  // namespace swcu
