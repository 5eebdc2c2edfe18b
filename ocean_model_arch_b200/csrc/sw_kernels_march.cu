// sw_kernels_march.cu -- k_march: the whole shallow-water step (K1..K11) in ONE launch in tolerance
// mode (sw_fast.cuh), built for the HBM roofline of B200 instead of its fp64 pipe.
//
// Work decomposition.  A WARP owns 28 output columns (its 32 lanes carry a 2-column halo on each
// side) and marches up a band of rows, one row per iteration.  There is no CTA-wide barrier and no
// shared-memory tile of intermediates:
//   - the eight input arrays (ssh, sshp, u, up, v, vp, hhq_rest, mu), the mask bytes and the row's
//     coefficients arrive through a per-warp ring of RING (6) rows in shared memory, filled with cp.async
//     (LDGSTS, zero-filled outside the array) RING-3 rows ahead of their use: all of a warp's HBM reads
//     are in flight while it computes, and nothing in the loop waits for a global load;
//   - raw inputs are read from the ring where they are needed (rows b .. b+2 stay there for the whole
//     iteration); only computed values live across rows;
//   - stage A (depths, volume fluxes, stresses, vorticity: the quantities a U/V/H/T point owns) is
//     evaluated one row ahead of stage B; its results rotate through registers the same way and reach
//     the east / west neighbours by warp shuffles;
//   - the row's coefficients (a 256-byte table row, sw_fast.cuh) travel through the same ring and are read
//     with broadcast 16-byte loads (as plain global loads their L2 latency was 59 % of all stall samples:
//     profiles/r02_march_v1_coef_ldg.txt).
// Per cell-row this costs 118 fp64 instructions, ~28 shared-memory loads of inputs, 19 of coefficients and 23
// shuffles, against ~490 fp64 instructions and ~160 shared-memory accesses of the bitwise kernel k_step.
//
// Launch geometry: ncol = ceil(columns / 28) warp columns x nbands bands of rows, nbands chosen so
// that all warps are resident at once (one wave, equal work); with enough land, 128-row bands whose all-land
// members exit at once.  Boundary strips of a multi-GPU step are the same kernel on 2 rows (NCCL path) or, over
// peer memory, the PEER instantiation: a small concurrent launch that also stores the strips into the
// neighbours' halo rows and publishes the step counter (MarchPeer, sw_fused.h).  profiles/r02_notes.md has the
// measurements behind every choice in this file.
#include <cstring>

#include "sw_fast.cuh"
#include "sw_fused.h"

namespace swcu {

namespace {

#ifndef SWCU_MW
#define SWCU_MW 4
#endif
#ifndef SWCU_RING
#define SWCU_RING 6   // measured: 4 rows 0.1259, 5 rows 0.1235, 6 rows 0.1200, 8 rows 0.1231 ms (2048^2, same box)
#endif
constexpr int MW = SWCU_MW;  // warps per CTA (the CTA is only a container: warps never synchronise with each other)
constexpr int WOUT = 28;   // output columns per warp
constexpr int NARR = 8;    // staged input arrays
constexpr int NROW = NARR + 1;  // 256-byte rows per ring slot: the eight arrays + the row's coefficients
constexpr int RING = SWCU_RING;  // rows in the per-warp ring
constexpr int PADW = 2;    // doubles of padding at both ends of a warp's ring (lane -1 / lane 32 reads)
constexpr int RING_DOUBLES = RING * NROW * 32 + 2 * PADW;
// Two-row unroll of the marching loop (computed values then swap between two register banks without copies).
// Measured SLOWER than the plain loop that copies ~18 doubles per row (0.1294 vs 0.1175 ms at 2048^2, same box):
// the unrolled body is 13 KB of SASS, the plain one 7 KB.  Kept as an option.
#ifndef SWCU_MARCH_UNROLL2
#define SWCU_MARCH_UNROLL2 0
#endif
#ifndef SWCU_RING_TMA
#define SWCU_RING_TMA 0
#endif
// Ring fill: 0 (default) = per-lane 16-byte cp.async (LDGSTS); 1 = TMA bulk copies (cp.async.bulk, UBLKCP in SASS)
// issued by one lane and completing on one mbarrier per ring row.  Measured at 2048^2 on one B200 (same box, same
// call): 0.1315 ms with cp.async, 0.1346 ms with bulk copies -- one lane issuing ten copies per row serialises
// what 32 lanes issue in five instructions, and every row pays an mbarrier try_wait.  The bulk path cannot zero-fill, so it reads
// whole 256-byte row segments that may run up to 31 columns past the pitch (into the next row, or into the
// slack every plane is allocated with): those columns only ever feed discarded lanes.
constexpr bool RING_TMA = SWCU_RING_TMA != 0;
constexpr int MASK_ROW = RING_TMA ? 48 : 32;   // bytes per mask ring row (bulk copies start 16-byte aligned)
constexpr int MASK_RING_BYTES = RING * MASK_ROW + 16;  // (+ pad: lane 31 reads its east neighbour)
constexpr int MBAR_BYTES = RING * 8;
constexpr size_t MARCH_SMEM = (size_t)MW * (RING_DOUBLES * sizeof(double) + MASK_RING_BYTES + MBAR_BYTES);
constexpr unsigned ROW_TX_BYTES = NROW * 256 + MASK_ROW;

enum { A_SSH, A_SSHP, A_U, A_UP, A_V, A_VP, A_H, A_MU };

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp16(unsigned dst, const void *src, bool valid)
{
    const unsigned sz = valid ? 16u : 0u;  // src-size 0: nothing is read, the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp4(unsigned dst, const void *src, bool valid)
{
    const unsigned sz = valid ? 4u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
// ---- TMA bulk copies + mbarrier (one per ring row and warp)
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 64-bit shuffles as two 32-bit ones on the unpacked halves (plain intrinsics: the compiler may schedule
// them freely between the two unrolled rows)
__device__ __forceinline__ double shfl_dn(double x)
{
    const int lo = __shfl_down_sync(0xffffffffu, __double2loint(x), 1);
    const int hi = __shfl_down_sync(0xffffffffu, __double2hiint(x), 1);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_up(double x)
{
    const int lo = __shfl_up_sync(0xffffffffu, __double2loint(x), 1);
    const int hi = __shfl_up_sync(0xffffffffu, __double2hiint(x), 1);
    return __hiloint2double(hi, lo);
}

// predicated (never branching) store.  No "memory" clobber: the output planes are never read by this
// kernel, so the store may move freely among the shared-memory loads of the next row.
__device__ __forceinline__ void st_if(bool pred, double *ptr, double v)
{
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %0, 0; @p st.global.f64 [%1], %2; }" ::"r"((int)pred), "l"(ptr), "d"(v));
}

// Waits until the flag word (written by a neighbouring GPU over NVLink) reaches `want`.  A neighbour that never
// gets there (a dead process) must not hang this GPU for ever: after ~2^26 polls (several seconds) the wait gives
// up and raises *timeout, which swcu_synchronize reports as an error; the step's results are then invalid.
__device__ __forceinline__ void spin_until(const unsigned long long *flag, unsigned long long want, int *timeout)
{
    const volatile unsigned long long *f = flag;
    for (unsigned n = 0; *f < want; ++n) {
        if (n >> 26) { atomicExch(timeout, 1); break; }
        __nanosleep(64);
    }
}

struct MarchIn {
    const double *in[NARR];  // ssh sshp u up v vp hhq_rest mu
};

// One warp's march over rows [bs..be] of warp column `col`.  side >= 0: the rows are a boundary strip whose
// results are stored also into that neighbour's halo rows (side 0 = below, 1 = above) -- a rare, warp-uniform
// branch that fetches the neighbour's plane pointers from the parameter bank only when it is taken.
template <bool TRANS, bool LAT, bool FFS, bool HAS_RHS, bool HAS_RDISS, bool PUSH>
__device__ __forceinline__ void march_warp(const Geo &g, const FusedArgs &a, const MarchIn &src, const MarchPeer &peer,
                                           unsigned char *smem_raw, const int lane, const int wib, const int col,
                                           const int bs, const int be, const int side)
{
    using namespace swf;
    double *ring = reinterpret_cast<double *>(smem_raw) + (size_t)wib * RING_DOUBLES + PADW;
    const unsigned ring_s = smem_addr(ring);
    const unsigned char *mring = smem_raw + (size_t)MW * RING_DOUBLES * sizeof(double) + (size_t)wib * MASK_RING_BYTES;
    const unsigned mring_s = smem_addr(mring);
    const unsigned bars = smem_addr(smem_raw + (size_t)MW * (RING_DOUBLES * sizeof(double) + MASK_RING_BYTES) + (size_t)wib * MBAR_BYTES);
    const int p = g.pitch, h = g.by2 - g.by1 + 1;
    const int ax = g.nx_start - 2 - g.bx1 + WOUT * col;  // array column of lane 0
    const int ac = ax + lane;                            // my array column
    const int r_first = bs - 2 - g.by1;                  // array row of the first staged row (>= 0)
    const int r_last = be + 2 - g.by1;                   // last row anybody reads (<= h-1)
    const int mo = RING_TMA ? (ax & 15) : 0;             // my mask byte within a mask ring row: mo + lane

    // ---- ring fill.  TMA: lane 0 arms the row's mbarrier and issues ten bulk copies (eight array rows, the
    // coefficient row, the mask bytes).  cp.async: one array row = 16 chunks of 16 B, a warp instruction
    // moves two arrays' rows.
    const int chunk = lane & 15, half = lane >> 4;
    const int ccol = ax + 2 * chunk;
    const bool col_ok = ccol + 1 < p;
    auto issue_row = [&](int r, int slot) {  // r: array row
        if (RING_TMA) {
            if (lane == 0 && r <= r_last) {
                const unsigned bar = bars + 8u * slot;
                mbar_expect(bar, ROW_TX_BYTES);
                const long off = (long)r * p + ax;
                const unsigned dst = ring_s + (unsigned)(slot * NROW * 256);
#pragma unroll
                for (int k = 0; k < NARR; ++k) bulk_g2s(dst + 256u * k, src.in[k] + off, 256u, bar);
                bulk_g2s(dst + 256u * NARR, a.fc + (long)r * swf::FC_STRIDE, 256u, bar);
                bulk_g2s(mring_s + (unsigned)(slot * MASK_ROW), a.mask + (long)r * p + (ax & ~15), (unsigned)MASK_ROW, bar);
            }
            return;
        }
        const bool ok = col_ok && r >= 0 && r <= r_last && r < h;
        const long off = ok ? (long)r * p + ccol : 0;
#pragma unroll
        for (int i = 0; i < NARR / 2; ++i) {
            const double *base = half ? src.in[2 * i + 1] : src.in[2 * i];
            cp16(ring_s + (unsigned)(((slot * NROW + 2 * i + half) * 32 + 2 * chunk) * sizeof(double)), base + off, ok);
        }
        // the row's coefficients (sw_fast.cuh: 32 doubles per row) ride in the same group
        const bool rok = r >= 0 && r <= r_last && r < h;
        if (half == 0) {
            cp16(ring_s + (unsigned)(((slot * NROW + NARR) * 32 + 2 * chunk) * sizeof(double)),
                 a.fc + (rok ? (long)r * swf::FC_STRIDE + 2 * chunk : 0), rok);
        } else if (chunk < 8) {  // ... and its 32 mask bytes (4-byte chunks; 0 outside the plane)
            const bool mok = rok && ax + 4 * chunk + 3 < p;
            cp4(mring_s + (unsigned)(slot * 32 + 4 * chunk), a.mask + (mok ? (long)r * p + ax + 4 * chunk : 0), mok);
        }
        cp_commit();
    };
    // row `j` (counted from r_first) of the ring has landed: j-th use of slot j % RING
    auto wait_row = [&](int slot, int use) {
        if (RING_TMA) mbar_wait(bars + 8u * slot, (unsigned)use & 1u);
    };
    if (RING_TMA) {
        if (lane == 0)
            for (int j = 0; j < RING; ++j) mbar_init(bars + 8u * j, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
#define RNG(slot, arr, dl) ring[((slot) * NROW + (arr)) * 32 + lane + (dl)]

#pragma unroll
    for (int j = 0; j < RING; ++j) issue_row(r_first + j, j);

    // ---- state.  Raw inputs are re-read from the ring where they are needed (rows b .. b+2 stay there for the
    // whole iteration); only COMPUTED values live across iterations, in two banks: in the iteration that outputs
    // row b, bank X holds row b's values and bank Y receives row b+1's.  (With SWCU_MARCH_UNROLL2 the banks swap
    // roles every row and nothing is copied; the default plain loop copies Y to X after every row and is faster.)
    struct Bank {
        AOut A;            // stage A of the bank's row
        double q, qm, s;   // thickness at T points of the row ABOVE the bank's row: unmasked, masked, qm_c + qm_e
        double qpm, sp;    // lagged masked thickness of the bank's row and qpm_c + qpm_e
        double fyp, fypy;  // northward face fluxes of the row below the bank's row
    } S[2];
    S[0].A = AOut{0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    S[1].A = S[0].A;
    S[0].qpm = S[0].sp = S[0].fyp = S[0].fypy = 0.0;
    S[1].qpm = S[1].sp = S[1].fyp = S[1].fypy = 0.0;
    unsigned mb0 = 0u, mb1, lue0 = 0u, lue1;
    float rd_n = 0.0f;

    // prologue: what stage A of row bs-1 needs from row bs-1 (slot 1)
    if (RING_TMA) { wait_row(0, 0); wait_row(1, 0); }
    else cp_wait<RING - 2>();
    __syncwarp();
    mb1 = mring[1 * MASK_ROW + mo + lane];
    {
        const double h1 = RNG(1, A_H, 0);
        S[0].q = FFS ? h1 + RNG(1, A_SSH, 0) : h1;
    }
    S[0].qm = (mb1 & LU) ? S[0].q : 0.0;
    S[0].s = S[0].qm + shfl_dn(S[0].qm);
    S[1].q = S[1].qm = S[1].s = 0.0;
    lue1 = __shfl_down_sync(0xffffffffu, mb1 & LU, 1);

    const double ts_half = 0.5 * a.ts;
    const bool lane_out = lane >= 2 && lane < 2 + WOUT && (g.bx1 + ac) <= g.nx_end;
    int s0 = 0, s1i = 1, s2i = 2;  // ring slots of rows b, b+1, b+2

    auto row = [&](const int b, Bank &X, Bank &Y) {
        const int rb = b - g.by1;  // array row of b
        const double *r0 = ring + s0 * (NROW * 32) + lane, *r1 = ring + s1i * (NROW * 32) + lane,
                     *r2 = ring + s2i * (NROW * 32) + lane;
#define AT(rp, arr, dl) (rp)[(arr) * 32 + (dl)]

        if (RING_TMA) wait_row(s2i, (rb - r_first + 2) / RING);   // row b+2 has landed
        else cp_wait<RING - 3>();                                  // ... (my copies) ...
        __syncwarp();                                              // ... and everybody else's

        const unsigned mb2 = mring[s2i * MASK_ROW + mo + lane];
        // coefficient rows from the ring (uniform addresses: broadcast 16-byte loads; pair k = columns 2k, 2k+1)
        ACoef ka;
        BCoef kb;
        {
            const double2 *ra = reinterpret_cast<const double2 *>(ring + (s1i * NROW + NARR) * 32);
            const double2 *rw = reinterpret_cast<const double2 *>(ring + (s0 * NROW + NARR) * 32);
            double2 t;
            t = ra[0]; ka.ku = t.x; ka.area = t.y;
            t = ra[1]; ka.area_n = t.x; ka.kv = t.y;
            t = ra[2]; ka.kh = t.x; ka.dyh = t.y;
            t = ra[3]; ka.dxh = t.x; ka.c1 = t.y;
            t = ra[4]; ka.c2 = t.x; ka.c3 = t.y;
            t = ra[5]; ka.cor = t.x; ka.s1 = t.y;
            t = ra[6]; ka.rxy = t.x; ka.rdxh = t.y;
            t = ra[7]; ka.rdxh_s = t.x; ka.rxyb = t.y;
            t = ra[8]; ka.rdxt = t.x; ka.rdxt_n = t.y;
            t = ra[9]; ka.s2 = t.x;
            t = rw[0]; kb.ku = t.x; kb.area = t.y;
            t = rw[1]; kb.area_n = t.x; kb.kv = t.y;
            t = rw[9]; kb.cssh = t.y;
            t = rw[10]; kb.cu = t.x; kb.gx = t.y;
            t = rw[11]; kb.dcx = t.x; kb.dxb2 = t.y;
            t = rw[12]; kb.dxb2_s = t.x; kb.cv = t.y;
            t = rw[13]; kb.gy = t.x; kb.dx2 = t.y;
            t = rw[14]; kb.dx2_n = t.x; kb.dcy = t.y;
            t = rw[15]; kb.rdxt = t.x; kb.rdxh = t.y;
            kb.tau = a.tau.tau;
        }

        // what stage B needs from the row below (bank Y still holds row b-1), before stage A overwrites it
        const double dvh = X.A.vh - Y.A.vh;
        const double dss = south_ss(kb, X.A.ss, Y.A.ss);
        const double zxs = X.A.zx + Y.A.zx;
        const double fyp_s = X.fyp, fypy_s = X.fypy;

        // ---- stage A, row b+1 (into bank Y)
        const double h2 = AT(r2, A_H, 0);
        const double q2 = FFS ? h2 + AT(r2, A_SSH, 0) : h2;
        const double qm2 = (mb2 & LU) ? q2 : 0.0;
        const double s2 = qm2 + shfl_dn(qm2);
        const unsigned lue2 = __shfl_down_sync(0xffffffffu, mb2 & LU, 1);
        const double mu1 = AT(r1, A_MU, 0);
        const double mus = mu1 + AT(r2, A_MU, 0);
        const double musum = mus + shfl_dn(mus);
        const int bc1 = (int)(mb1 & LU), bc2 = (int)(mb2 & LU);
        const double u1 = AT(r1, A_U, 0), v1 = AT(r1, A_V, 0), v_e1 = AT(r1, A_V, 1);
        const double vp0 = AT(r0, A_VP, 0);
        Y.A = stage_a<TRANS, LAT>(ka, mb1, bc1 + (int)lue1, bc1 + bc2, bc1 + (int)lue1 + bc2 + (int)lue2, X.q, X.qm, qm2, X.s, s2,
                                  u1, AT(r2, A_U, 0), v1, v_e1, AT(r1, A_UP, 0), AT(r1, A_UP, -1), AT(r2, A_UP, 0),
                                  AT(r1, A_VP, 0), vp0, AT(r1, A_VP, 1), mu1, musum);
        Y.q = q2; Y.qm = qm2; Y.s = s2;

        // ---- stage B, row b
        const double h1 = AT(r1, A_H, 0);
        const double qpm1 = (mb1 & LU) ? (FFS ? h1 + AT(r1, A_SSHP, 0) : h1) : 0.0;
        Y.qpm = qpm1;
        Y.sp = qpm1 + shfl_dn(qpm1);
        const double pp = mad(X.qpm, kb.area, qpm1 * kb.area_n);
        const double uh_w0 = shfl_up(X.A.uh);
        const double u0 = AT(r0, A_U, 0), v0 = AT(r0, A_V, 0);
        Flux f = {0.0, 0.0, 0.0, 0.0};
        double fxp_w = 0.0, fxpy_w = 0.0;
        if (TRANS) {
            f = stage_flux((mb0 & LUU) != 0, X.A.uh, shfl_dn(X.A.uh), Y.A.uh, X.A.vh, shfl_dn(X.A.vh), Y.A.vh, u0, AT(r0, A_U, 1),
                           u1, v0, AT(r0, A_V, 1), v1);
            fxp_w = shfl_up(f.fxp);
            fxpy_w = shfl_up(f.fxpy);
        }
        Y.fyp = f.fyp; Y.fypy = f.fypy;
        double t_e0 = 0.0, ss_w0 = 0.0;
        if (LAT) { t_e0 = shfl_dn(X.A.t); ss_w0 = shfl_up(X.A.ss); }
        const double zy_w0 = shfl_up(X.A.zy);

        const bool out = lane_out && b >= bs;
        const long gc = (long)rb * p + ac;
        double rhsx = 0.0, rhsy = 0.0, rdx = 0.0, rdy = 0.0;
        if (HAS_RHS && out) { rhsx = a.RHSx[gc]; rhsy = a.RHSy[gc]; }
        if (HAS_RDISS) {  // vel_ssh.f90:171,185: dble(rdis(m,n) + rdis(m+1,n)), dble(rdis(m,n) + rdis(m,n+1))
            const float rd_c = rd_n;  // loaded as the row above one iteration ago
            rd_n = (ac < p && rb + 1 < h) ? __ldg(a.rdis + (long)(rb + 1) * p + ac) : 0.0f;
            const float rd_e = __shfl_down_sync(0xffffffffu, rd_c, 1);
            rdx = (double)(rd_c + rd_e);
            rdy = (double)(rd_c + rd_n);
        }
        const int bc0 = (int)(mb0 & LU);
        const BRaw o = stage_b_raw<TRANS, LAT>(kb, bc0 + (int)lue0, bc0 + bc1, ts_half, AT(r0, A_SSH, 0), AT(r0, A_SSH, 1),
                                               AT(r1, A_SSH, 0), AT(r0, A_SSHP, 0), X.sp, pp, u0, AT(r0, A_UP, 0), v0, vp0,
                                               X.A.rhu, X.A.rhv, X.A.uh, uh_w0, dvh, X.A.t, t_e0, Y.A.t, X.A.ss, dss, ss_w0,
                                               zxs, X.A.zy, zy_w0, f, fxp_w, fyp_s, fxpy_w, fypy_s, rhsx, rhsy, rdx, rdy);
        // Masked-out cells keep their values, and BOTH ping-pong buffers already hold them (the context copies
        // the planes once): only the cells the reference's kernels assign are stored (vel_ssh.f90:225-243).
        const bool sea = out && (mb0 & LU), wu = out && (mb0 & LCU), wv = out && (mb0 & LCV);
        st_if(sea, a.ssh_o + gc, o.sshn); st_if(sea, a.sshp_o + gc, o.sshpf);
        st_if(wu, a.u_o + gc, o.un); st_if(wu, a.up_o + gc, o.upf);
        st_if(wv, a.v_o + gc, o.vn); st_if(wv, a.vp_o + gc, o.vpf);
        if (sea && ssh_bad(o.sshn)) atomicAdd(a.bad, 1);  // K11
        if (PUSH && side >= 0 && b >= bs) {
            // boundary strip: the same cells go straight into the neighbour's halo rows (NVLink stores).  The
            // neighbour must have declared those rows of its write buffers free for this step.
            if (b == bs && !(peer.dbg & 1)) {
                if (lane == 0) spin_until(side ? peer.free_[1] : peer.free_[0], peer.tick, peer.timeout);
                __syncwarp();
                __threadfence_system();
            }
            const bool ps = !(peer.dbg & 2);
            double *const *po = side ? peer.out[1] : peer.out[0];
            st_if(sea && ps, po[0] + gc, o.sshn); st_if(sea && ps, po[1] + gc, o.sshpf);
            st_if(wu && ps, po[2] + gc, o.un); st_if(wu && ps, po[3] + gc, o.upf);
            st_if(wv && ps, po[4] + gc, o.vn); st_if(wv && ps, po[5] + gc, o.vpf);
        }
#undef AT

        // ---- row b leaves the window: refill its ring slot
        __syncwarp();
        issue_row(rb + RING, s0);
        s0 = s1i; s1i = s2i; s2i = s2i + 1 == RING ? 0 : s2i + 1;
        mb0 = mb1; mb1 = mb2; lue0 = lue1; lue1 = lue2;
    };
#if SWCU_MARCH_UNROLL2
    for (int b = bs - 2; b <= be; b += 2) {
        row(b, S[0], S[1]);
        if (b + 1 <= be) row(b + 1, S[1], S[0]);
    }
#else
#pragma unroll 1
    for (int b = bs - 2; b <= be; ++b) {
        row(b, S[0], S[1]);
        const AOut keep = S[0].A;   // row b's stage A is still the "row below" of the next iteration
        S[0] = S[1];
        S[1].A = keep;
    }
#endif
    cp_wait<0>();
#undef RNG
}

// MINB = CTAs per SM the register budget is sized for (2: 255 registers, 8 warps per SM; 3: 168 registers,
// 12 warps per SM).
// PEER = false: the lean kernel -- every warp marches its band, nothing else is compiled in (a segment loop
// around the march body costs every warp 4.5 %, a rarely taken push branch inside it another 4 %: measured).
// PEER = true: the boundary-strip kernel of the fused halo push -- the warps of band 0 work on the lower strip,
// those of band nbands-1 on the upper one (both if there is one band): wait for the neighbour's rows of the
// previous step, compute the strip, store it also into the neighbour's halo rows, publish the step counter.
// It runs concurrently with the lean kernel (other stream, disjoint rows).
template <bool TRANS, bool LAT, bool FFS, bool HAS_RHS, bool HAS_RDISS, int MINB, bool PEER>
__global__ void __launch_bounds__(MW * 32, MINB)
k_march(Geo g, FusedArgs a, MarchIn src, MarchPlan pl, MarchPeer peer)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int wid = blockIdx.x * MW + wib;
    if (wid >= pl.nwarps) return;
    const int band = wid / pl.ncol, col = wid - band * pl.ncol;
    if (!PEER) {
        int bs, be;
        march_band_rows(pl, band, &bs, &be);
        if (be < bs) return;
        if (pl.band_land && pl.band_land[wid]) return;  // every output cell of this warp's band is land
        march_warp<TRANS, LAT, FFS, HAS_RHS, HAS_RDISS, false>(g, a, src, peer, smem_raw, lane, wib, col, bs, be, -1);
        return;
    }
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        if (peer.out[side][0] == nullptr || band != (side ? pl.nbands - 1 : 0)) continue;
        const int bs = side == 0 ? peer.lo0 : peer.hi0, be = side == 0 ? peer.lo1 : peer.hi1;
        if (be >= bs && !(peer.dbg & 1)) {
            // the strip reads my halo rows of the current state: the neighbour's push of the previous step
            if (lane == 0) spin_until(side ? peer.ready_in[1] : peer.ready_in[0], peer.tick - 1, peer.timeout);
            __syncwarp();
        }
        if (be >= bs) march_warp<TRANS, LAT, FFS, HAS_RHS, HAS_RDISS, true>(g, a, src, peer, smem_raw, lane, wib, col, bs, be, side);
        __threadfence_system();  // every lane: its stores into the neighbour's memory are visible system-wide ...
        __syncwarp();
        if (lane == 0) {          // ... before the last strip warp of this side publishes the step counter there
            unsigned *cnt = side ? peer.count[1] : peer.count[0];
            if (atomicAdd(cnt, 1u) == (unsigned)pl.ncol - 1u) {
                *cnt = 0;
                __threadfence_system();
                *reinterpret_cast<volatile unsigned long long *>(side ? peer.ready[1] : peer.ready[0]) = peer.tick;
            }
        }
        __syncwarp();
    }
}


// ---- expl_tracer (control/tracer.f90:44-61) for ONE tracer field in tolerance arithmetic, same marching scheme:
// a warp owns 28 output columns and walks up a band of rows; the new level's ssh, sshp, u, v, the tracer's two
// levels, hhq_rest and mu arrive through the per-warp ring; the only values that live across rows are the
// northward flux of the row below and (by shuffle) the eastward flux of the west neighbour.
constexpr int TR_RING = 6;   // rows t, t+1 live + four in flight: small enough for three CTAs per SM
constexpr int TR_RING_DOUBLES = TR_RING * NROW * 32 + 2 * PADW;
constexpr int TR_MASK_RING_BYTES = TR_RING * 32 + 16;
constexpr size_t TRACER_SMEM = (size_t)MW * (TR_RING_DOUBLES * sizeof(double) + TR_MASK_RING_BYTES);
enum { T_SSH = 0, T_SSHP = 1, T_U = 2, T_FF = 3, T_V = 4, T_FFP = 5, T_H = 6, T_MU = 7 };

template <bool FFS>
__global__ void __launch_bounds__(MW * 32, 3)
k_tracer_march(Geo g, FusedArgs a, MarchIn src, MarchPlan pl)
{
    using namespace swf;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int wid = blockIdx.x * MW + wib;
    if (wid >= pl.nwarps) return;
    const int band = wid / pl.ncol, col = wid - band * pl.ncol;
    int bs, be;
    march_band_rows(pl, band, &bs, &be);
    if (be < bs) return;
    if (pl.band_land && pl.band_land[wid]) return;

    double *ring = reinterpret_cast<double *>(smem_raw) + (size_t)wib * TR_RING_DOUBLES + PADW;
    const unsigned ring_s = smem_addr(ring);
    const unsigned char *mring = smem_raw + (size_t)MW * TR_RING_DOUBLES * sizeof(double) + (size_t)wib * TR_MASK_RING_BYTES;
    const unsigned mring_s = smem_addr(mring);
    const int p = g.pitch, h = g.by2 - g.by1 + 1;
    const int ax = g.nx_start - 2 - g.bx1 + WOUT * col;
    const int ac = ax + lane;
    const int r_first = bs - 1 - g.by1, r_last = be + 1 - g.by1;
    const int chunk = lane & 15, half = lane >> 4;
    const int ccol = ax + 2 * chunk;
    const bool col_ok = ccol + 1 < p;
    auto issue_row = [&](int r, int slot) {
        const bool rok = r >= 0 && r <= r_last && r < h;
        const bool ok = col_ok && rok;
        const long off = ok ? (long)r * p + ccol : 0;
#pragma unroll
        for (int i = 0; i < NARR / 2; ++i) {
            const double *base = half ? src.in[2 * i + 1] : src.in[2 * i];
            cp16(ring_s + (unsigned)(((slot * NROW + 2 * i + half) * 32 + 2 * chunk) * sizeof(double)), base + off, ok);
        }
        if (half == 0) {
            if (chunk < FT_STRIDE / 2)
                cp16(ring_s + (unsigned)(((slot * NROW + NARR) * 32 + 2 * chunk) * sizeof(double)),
                     a.ft + (rok ? (long)r * FT_STRIDE + 2 * chunk : 0), rok);
        } else if (chunk < 8) {
            const bool mok = rok && ax + 4 * chunk + 3 < p;
            cp4(mring_s + (unsigned)(slot * 32 + 4 * chunk), a.mask + (mok ? (long)r * p + ax + 4 * chunk : 0), mok);
        }
        cp_commit();
    };
#pragma unroll
    for (int j = 0; j < TR_RING; ++j) issue_row(r_first + j, j);

    const double ts_half = 0.5 * a.ts;
    const bool lane_out = lane >= 2 && lane < 2 + WOUT && (g.bx1 + ac) <= g.nx_end;
    double fy_s = 0.0;
    int s0 = 0, s1 = 1;
#pragma unroll 2
    for (int t = bs - 1; t <= be; ++t) {
        const int rt = t - g.by1;
        cp_wait<TR_RING - 2>();
        __syncwarp();
        const double *r0 = ring + s0 * (NROW * 32) + lane, *r1 = ring + s1 * (NROW * 32) + lane;
#define AT(rp, arr, dl) (rp)[(arr) * 32 + (dl)]
        TCoef k;
        {
            const double2 *rw = reinterpret_cast<const double2 *>(ring + (s0 * NROW + NARR) * 32);
            double2 q;
            q = rw[0]; k.ku = q.x; k.area = q.y;
            q = rw[1]; k.area_n = q.x; k.kv = q.y;
            q = rw[2]; k.dyh = q.x; k.dxh = q.y;
            q = rw[3]; k.dyh_rdxt = q.x; k.dxh_rdyt = q.y;
            q = rw[4]; k.ctr = q.x;
        }
        const unsigned mb0 = mring[s0 * 32 + lane], mbe = mring[s0 * 32 + lane + 1], mb1 = mring[s1 * 32 + lane];
        const double h0 = AT(r0, T_H, 0);
        const double qm_c = (mb0 & LU) ? (FFS ? h0 + AT(r0, T_SSH, 0) : h0) : 0.0;
        const double he = AT(r0, T_H, 1), hn = AT(r1, T_H, 0);
        const double qm_e = (mbe & LU) ? (FFS ? he + AT(r0, T_SSH, 1) : he) : 0.0;
        const double qm_n = (mb1 & LU) ? (FFS ? hn + AT(r1, T_SSH, 0) : hn) : 0.0;
        const int bc = (int)(mb0 & LU);
        const double ff0 = AT(r0, T_FF, 0);
        const TFlux f = tracer_flux(k, mb0, bc + (int)(mbe & LU), bc + (int)(mb1 & LU), qm_c, qm_e, qm_n, AT(r0, T_U, 0),
                                    AT(r0, T_V, 0), AT(r0, T_MU, 0), AT(r0, T_MU, 1), AT(r1, T_MU, 0), ff0, AT(r0, T_FF, 1),
                                    AT(r1, T_FF, 0));
        const double fx_w = shfl_up(f.fx);
        const double ffp0 = AT(r0, T_FFP, 0);
        const TOut o = tracer_update(k, ts_half, h0, FFS ? h0 + AT(r0, T_SSHP, 0) : h0, ff0, ffp0, f.fx, fx_w, f.fy, fy_s);
        const bool out = lane_out && t >= bs && (mb0 & LU);
        const long gc = (long)rt * p + ac;
        st_if(out, a.ff_o + gc, o.ffn);
        st_if(out, a.ffp_o + gc, o.ffpf);
        fy_s = f.fy;
#undef AT
        __syncwarp();
        issue_row(rt + TR_RING, s0);
        s0 = s1; s1 = s1 + 1 == TR_RING ? 0 : s1 + 1;
    }
    cp_wait<0>();
}

inline int launched(const char *what)
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SWCU_OK : cuda_fail(e, what);
}

template <bool T, bool L, bool F, bool R, bool D, int MINB, bool PEER>
int march_launch_b(const Geo &g, const FusedArgs &a, const MarchIn &src, const MarchPlan &pl, const MarchPeer &peer,
                   cudaStream_t st)
{
    static unsigned long long attr_set = 0;  // per device opt-in for > 48 KB of dynamic shared memory
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(__atomic_load_n(&attr_set, __ATOMIC_ACQUIRE) >> (dev & 63) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k_march<T, L, F, R, D, MINB, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)MARCH_SMEM);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_march)");
        __atomic_fetch_or(&attr_set, 1ull << (dev & 63), __ATOMIC_RELEASE);
    }
    const unsigned grid = (unsigned)((pl.nwarps + MW - 1) / MW);
    k_march<T, L, F, R, D, MINB, PEER><<<grid, MW * 32, MARCH_SMEM, st>>>(g, a, src, pl, peer);
    return launched("k_march");
}
template <bool T, bool L, bool F, bool R, bool D>
int march_launch(const Geo &g, const FusedArgs &a, const MarchIn &src, const MarchPlan &pl, const MarchPeer &peer,
                 cudaStream_t st)
{
    if (peer.out[0][0] || peer.out[1][0]) return march_launch_b<T, L, F, R, D, 2, true>(g, a, src, pl, peer, st);
    return pl.minb == 3 ? march_launch_b<T, L, F, R, D, 3, false>(g, a, src, pl, peer, st)
                        : march_launch_b<T, L, F, R, D, 2, false>(g, a, src, pl, peer, st);
}

template <bool T, bool L, bool F>
int march_dispatch2(const Geo &g, const FusedArgs &a, const MarchIn &src, const MarchPlan &pl, const MarchPeer &peer,
                    cudaStream_t st)
{
    const bool r = a.RHSx != nullptr, d = a.rdis != nullptr;
    if (r && d) return march_launch<T, L, F, true, true>(g, a, src, pl, peer, st);
    if (r) return march_launch<T, L, F, true, false>(g, a, src, pl, peer, st);
    if (d) return march_launch<T, L, F, false, true>(g, a, src, pl, peer, st);
    return march_launch<T, L, F, false, false>(g, a, src, pl, peer, st);
}

template <bool T, bool L>
int march_dispatch1(const Geo &g, const FusedArgs &a, const MarchIn &src, const MarchPlan &pl, const MarchPeer &peer,
                    cudaStream_t st)
{
    return a.ffs != 0.0 ? march_dispatch2<T, L, true>(g, a, src, pl, peer, st) : march_dispatch2<T, L, false>(g, a, src, pl, peer, st);
}

// one thread per table row
__global__ void k_build_fast(const double *__restrict__ tab, int h, double tau, double *__restrict__ fc, double *__restrict__ ft)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= h) return;
    double row[swf::FC_STRIDE];
    swf::build_fast_row(tab, h, r, tau, row);
#pragma unroll
    for (int k = 0; k < swf::FC_STRIDE; ++k) fc[(long)r * swf::FC_STRIDE + k] = row[k];
    swf::build_tracer_row(tab, h, r, tau, row);
#pragma unroll
    for (int k = 0; k < swf::FT_STRIDE; ++k) ft[(long)r * swf::FT_STRIDE + k] = row[k];
}

// band_land[w] = 1 <=> no sea cell among the output cells of warp w's band
__global__ void k_band_land(Geo g, const unsigned char *__restrict__ mask, MarchPlan pl, unsigned char *__restrict__ out,
                            int *__restrict__ nland)
{
    const int w = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (w >= pl.ncol * pl.nbands) return;
    const int band = w / pl.ncol, col = w - band * pl.ncol;
    int bs, be;
    march_band_rows(pl, band, &bs, &be);
    const int m = g.nx_start + WOUT * col + lane;
    int any = 0;
    if (lane < WOUT && m <= g.nx_end)
        for (int n = bs; n <= be; ++n) any |= mask[ix(g, m, n)] & MB_LU;
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) {
        out[w] = any ? 0 : 1;
        if (!any && nland) atomicAdd(nland, 1);
    }
}

}  // namespace

bool march_supported(const Geo &g, const FusedArgs &a)
{
    // 16-byte cp.async chunks: the first staged column of every warp must be even in array coordinates
    // 16-byte chunks of the arrays and 4-byte chunks of the mask plane: the first staged column of a warp
    // (28 columns apart) must be a multiple of 4 in array coordinates
    return a.tab != nullptr && a.fc != nullptr && (g.nx_start - 2 - g.bx1) >= 0 && (g.nx_start - 2 - g.bx1) % 4 == 0 &&
           g.pitch % 4 == 0 && g.ny_start - 2 >= g.by1 && g.ny_end + 2 <= g.by2;
}

// Warp columns x bands for rows [n0..n1].  band_rows = 0: as many bands as keep every warp resident at once
// (max_warps = SMs x resident warps; one wave, equal work), but bands of at least 16 rows (each band pays 2
// warm-up rows).  band_rows > 0: bands of about that many rows -- more warps than fit at once, the hardware
// block scheduler balances them; for basins with land, where all-land bands cost nothing.
void march_plan(const Geo &g, int n0, int n1, int max_warps, MarchPlan *pl, int band_rows)
{
    pl->n0 = n0; pl->n1 = n1;
    pl->ncol = (g.nx_end - g.nx_start + WOUT) / WOUT;
    const int nrows = n1 - n0 + 1;
    const int min_rows = 16;
    int nb;
    if (band_rows > 0) {
        nb = (nrows + band_rows - 1) / band_rows;
    } else {
        nb = max_warps / (pl->ncol > 0 ? pl->ncol : 1);
        if (nb > (nrows + min_rows - 1) / min_rows) nb = (nrows + min_rows - 1) / min_rows;
    }
    if (nb < 1) nb = 1;
    pl->nbands = nb;
    pl->nwarps = pl->ncol * nb;
    pl->band_land = nullptr;
    pl->minb = 2;
    pl->late_lo = pl->late_hi = pl->late_cut = 0;
}

int march_resident_warps(int device, int minb)
{
    int sms = 0, per_sm = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 148 * 4 * minb;
    cudaError_t e;
    if (minb == 3) {
        e = cudaFuncSetAttribute(k_march<true, true, true, false, false, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MARCH_SMEM);
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_march<true, true, true, false, false, 3, false>, MW * 32, MARCH_SMEM);
    } else {
        e = cudaFuncSetAttribute(k_march<true, true, true, false, false, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MARCH_SMEM);
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_march<true, true, true, false, false, 2, false>, MW * 32, MARCH_SMEM);
    }
    if (e != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = minb; }
    return sms * per_sm * MW;
}

// SMs x resident warps of k_tracer_march (fewer registers than k_march: three CTAs per SM where shared memory allows)
int march_tracer_resident_warps(int device)
{
    int sms = 0, per_sm = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 148 * 8;
    cudaError_t e = cudaFuncSetAttribute(k_tracer_march<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRACER_SMEM);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tracer_march<true>, MW * 32, TRACER_SMEM);
    if (e != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 2; }
    return sms * per_sm * MW;
}

int march_strip_warps(const Geo &g) { return (g.nx_end - g.nx_start + WOUT) / WOUT; }

int launch_march(const Geo &g, const FusedArgs &a, const MarchPlan &pl, cudaStream_t st, const MarchPeer *peer_in)
{
    if (pl.nwarps < 1 || (pl.n1 < pl.n0 && !peer_in)) return SWCU_OK;
    MarchPeer peer;
    if (peer_in) peer = *peer_in;
    else memset(&peer, 0, sizeof(peer));
    MarchIn src;
    src.in[A_SSH] = a.ssh; src.in[A_SSHP] = a.sshp; src.in[A_U] = a.u; src.in[A_UP] = a.up; src.in[A_V] = a.v;
    src.in[A_VP] = a.vp; src.in[A_H] = a.h_r; src.in[A_MU] = a.mu;
    if (a.trans && a.lat) return march_dispatch1<true, true>(g, a, src, pl, peer, st);
    if (a.trans) return march_dispatch1<true, false>(g, a, src, pl, peer, st);
    if (a.lat) return march_dispatch1<false, true>(g, a, src, pl, peer, st);
    return march_dispatch1<false, false>(g, a, src, pl, peer, st);
}

int launch_tracer_march(const Geo &g, const FusedArgs &a, const MarchPlan &pl, cudaStream_t st)
{
    if (pl.nwarps < 1 || pl.n1 < pl.n0) return SWCU_OK;
    static unsigned long long attr_set = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(__atomic_load_n(&attr_set, __ATOMIC_ACQUIRE) >> (dev & 63) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k_tracer_march<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRACER_SMEM);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_tracer_march<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRACER_SMEM);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_tracer_march)");
        __atomic_fetch_or(&attr_set, 1ull << (dev & 63), __ATOMIC_RELEASE);
    }
    MarchIn src;
    src.in[T_SSH] = a.ssh_o; src.in[T_SSHP] = a.sshp_o; src.in[T_U] = a.u_o; src.in[T_FF] = a.ff; src.in[T_V] = a.v_o;
    src.in[T_FFP] = a.ffp; src.in[T_H] = a.h_r; src.in[T_MU] = a.mu;
    const unsigned grid = (unsigned)((pl.nwarps + MW - 1) / MW);
    if (a.ffs != 0.0) k_tracer_march<true><<<grid, MW * 32, TRACER_SMEM, st>>>(g, a, src, pl);
    else k_tracer_march<false><<<grid, MW * 32, TRACER_SMEM, st>>>(g, a, src, pl);
    return launched("k_tracer_march");
}

int launch_build_fast(const double *tab, int h, double tau, double *fc, double *ft, cudaStream_t st)
{
    k_build_fast<<<(unsigned)((h + 63) / 64), 64, 0, st>>>(tab, h, tau, fc, ft);
    return launched("build_fast");
}

int launch_band_land(const Geo &g, const unsigned char *mask, const MarchPlan &pl, unsigned char *out, int *nland_dev,
                     cudaStream_t st)
{
    const int n = pl.ncol * pl.nbands;
    if (n < 1 || pl.n1 < pl.n0) return SWCU_OK;
    k_band_land<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(g, mask, pl, out, nland_dev);
    return launched("band_land");
}

}  // namespace swcu
