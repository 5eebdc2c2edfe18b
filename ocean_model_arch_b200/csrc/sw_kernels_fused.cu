// sw_kernels_fused.cu -- Level B fused step: the reference's K1..K11 collapsed into ONE launch
// (k_step: both stages below on shared-memory tiles staged by TMA; needs the per-row metric tables)
// or into two launches from global memory (k_prep + k_update; any grid), plus k_tracer.
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
// Metrics come through an accessor (sw_formulas.cuh): MetRow (per-row double tables, no
// conversions, no HBM traffic) when every metric array is constant along m, else MetGen.
// Arithmetic is the shared sw_formulas.cuh, so results are bitwise those of the 1:1 kernels.
#include "sw_fused.h"

namespace swcu {

namespace {

constexpr int BX = 64;
constexpr int BY = 4;

template <bool ROW> struct MetOf;
template <> struct MetOf<true> {
    typedef MetRow type;
    static __device__ __forceinline__ MetRow make(const FusedArgs &a) { return MetRow{a.tab, a.tab_h, 0}; }
};
template <> struct MetOf<false> {
    typedef MetGen type;
    static __device__ __forceinline__ MetGen make(const FusedArgs &a)
    {
        return MetGen{a.dx, a.dy, a.dxt, a.dyt, a.dxh, a.dyh, a.dxb, a.dyb, a.rlh_s};
    }
};

// ---- two-launch path: stages A and B straight from global memory (any metric arrays) ----------
template <bool ROW, bool TRANS, bool LAT>
__global__ void __launch_bounds__(BX *BY) k_prep(Geo g, FusedArgs a, int n0, int n1)
{
    const int m = g.nx_start - 1 + blockIdx.x * BX + threadIdx.x;
    const int n = n0 + blockIdx.y * BY + threadIdx.y;
    if (m > g.nx_end + 1 || n > n1) return;
    const long c = ix(g, m, n);
    const typename MetOf<ROW>::type mt = MetOf<ROW>::make(a);
    const PrepOut o = prep_cell<TRANS, LAT>(c, n - g.by1, g.pitch, mt, a.ffs, a.mask, a.ssh, a.h_r, a.u, a.v, a.up, a.vp);
    a.hu[c] = o.hu; a.hv[c] = o.hv; a.hh[c] = o.hh;
    if (TRANS) a.vort[c] = o.vort;
    if (LAT) { a.str_t[c] = o.str_t; a.str_s[c] = o.str_s; }
}

template <bool ROW, bool TRANS, bool LAT, bool HAS_RHS, bool HAS_RDISS>
__global__ void __launch_bounds__(BX *BY) k_update(Geo g, FusedArgs a, int n0, int n1)
{
    const int m = g.nx_start + blockIdx.x * BX + threadIdx.x;
    const int n = n0 + blockIdx.y * BY + threadIdx.y;
    if (m > g.nx_end || n > n1) return;
    const long c = ix(g, m, n);
    const int p = g.pitch;
    const typename MetOf<ROW>::type mt = MetOf<ROW>::make(a);
    const double rhsx = HAS_RHS ? a.RHSx[c] : 0.0, rhsy = HAS_RHS ? a.RHSy[c] : 0.0;
    const double rdx = HAS_RDISS ? (double)(a.rdis[c] + a.rdis[c + 1]) : 0.0;
    const double rdy = HAS_RDISS ? (double)(a.rdis[c] + a.rdis[c + p]) : 0.0;
    const UpdOut o = update_cell<TRANS, LAT>(c, n - g.by1, p, mt, a.tau, a.ts, a.ffs, a.mask, a.ssh, a.sshp, a.u, a.up,
                                             a.v, a.vp, a.h_r, a.mu, a.hu, a.hv, a.hh, a.vort, a.str_t, a.str_s,
                                             rhsx, rhsy, rdx, rdy);
    a.ssh_o[c] = o.ssh; a.sshp_o[c] = o.sshp; a.u_o[c] = o.u; a.up_o[c] = o.up; a.v_o[c] = o.v; a.vp_o[c] = o.vp;
    if (o.bad) atomicAdd(a.bad, 1);  // K11
}

// ---- tracer: fluxes + update + filter in one launch (after the step, on the n+1 state) ----------
template <bool ROW>
__global__ void __launch_bounds__(BX *BY) k_tracer(Geo g, FusedArgs a, int n0, int n1)
{
    const int m = g.nx_start + blockIdx.x * BX + threadIdx.x;
    const int n = n0 + blockIdx.y * BY + threadIdx.y;
    if (m > g.nx_end || n > n1) return;
    const long c = ix(g, m, n);
    const typename MetOf<ROW>::type mt = MetOf<ROW>::make(a);
    const TracerOut o = tracer_cell(c, n - g.by1, g.pitch, mt, a.tau, a.ts, a.ffs, a.mask, a.ssh_o, a.sshp_o, a.h_r,
                                    a.u_o, a.v_o, a.mu, a.ff, a.ffp);
    a.ff_o[c] = o.ff; a.ffp_o[c] = o.ffp;
}

// ---- one-launch path: TMA-staged shared-memory tiles, stages A and B in one kernel --------------
// A CTA owns TX x TY output cells.  One thread issues eight cp.async.bulk.tensor.2d loads (ssh,
// sshp, u, up, v, vp, hhq_rest, mu; box = tile + 2-cell halo, out-of-array cells zero-filled) that
// complete on one mbarrier; stage A then fills the six intermediate tiles for the tile grown by one
// cell, stage B updates the tile and stores the six n+1 arrays.  Nothing but the prognostic state
// touches HBM: 8 reads + 6 writes + 1 mask byte = 113 B per cell instead of 326 B.
constexpr int TX = 32, HALO = 2, IW = TX + 2 * HALO;   // 36 columns per tile row
constexpr int N_IN = 8, N_S1 = 6;

// tile variants: TY_ rows per tile, R_ vertically adjacent cells per thread in stage B, MINB_ CTAs per SM
template <int TY_, int R_, int MINB_> struct TileCfg {
    static constexpr int TY = TY_, R = R_, MINB = MINB_;
    static constexpr int IH = TY + 2 * HALO, TILE = IW * IH;
    static constexpr int THREADS = TX * TY / R;
    static constexpr int MASK_BYTES = (TILE + 127) / 128 * 128;
    static constexpr int TROWS = IH + 1;  // metric-table rows a tile touches (r .. r+1)
    static constexpr size_t SMEM = (size_t)(N_IN + N_S1) * TILE * sizeof(double) + MASK_BYTES + 16
                                 + (size_t)T_COUNT * TROWS * sizeof(double);
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <class CFG, bool TRANS, bool LAT, bool HAS_RHS, bool HAS_RDISS>
__global__ void __launch_bounds__(CFG::THREADS, CFG::MINB)
k_step(const __grid_constant__ StepMaps maps, Geo g, FusedArgs a, int n0, int n1)
{
    constexpr int TY = CFG::TY, R = CFG::R, IH = CFG::IH, TILE = CFG::TILE, THREADS = CFG::THREADS;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double *in = reinterpret_cast<double *>(smem_raw);
    double *s1 = in + N_IN * TILE;
    unsigned char *mk = reinterpret_cast<unsigned char *>(s1 + N_S1 * TILE);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(mk + CFG::MASK_BYTES);
    double *stab = reinterpret_cast<double *>(mk + CFG::MASK_BYTES + 16);  // [T_COUNT][TROWS]

    if (a.tile_land && a.tile_land[blockIdx.y * gridDim.x + blockIdx.x]) return;  // all-land tile (CTA-uniform)

    const int tid = threadIdx.x;
    const int i0 = g.nx_start + blockIdx.x * TX, j0 = n0 + blockIdx.y * TY;
    const int ax = i0 - HALO - g.bx1, ay = j0 - HALO - g.by1;  // tile origin in array coordinates (>= 0)
    const int h = g.by2 - g.by1 + 1;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                     "r"((unsigned)(N_IN * TILE * sizeof(double))) : "memory");
#pragma unroll
        for (int k = 0; k < N_IN; ++k)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(in + k * TILE)), "l"(reinterpret_cast<unsigned long long>(&maps.m[k])),
                         "r"(ax), "r"(ay), "r"(smem_u32(bar)) : "memory");
    }
    // Mask bytes and this tile's rows of the metric tables: plain loads, ALL issued before any is
    // consumed, so they cost one memory round trip that overlaps the TMA flight (a load-then-store
    // loop cost three serialized round trips per CTA: 25 % of all stall samples in ncu).
    {
        constexpr int WORDS_PER_ROW = IW / 4;                  // 9 four-byte words per 36-byte mask row
        constexpr int NW = WORDS_PER_ROW * IH;                 // rows start 4-byte aligned: ax % 32 == 0
        constexpr int MW = (NW + THREADS - 1) / THREADS;
        constexpr int NT = T_COUNT * CFG::TROWS;
        constexpr int MT = (NT + THREADS - 1) / THREADS;
        unsigned mw[MW];
        double tv[MT];
#pragma unroll
        for (int k = 0; k < MW; ++k) {
            const int idx = tid + k * THREADS, ly = idx / WORDS_PER_ROW, wx = idx % WORDS_PER_ROW;
            const int gy = ay + ly, gx = ax + 4 * wx;
            // columns past w but inside the pitch are zero padding; rows past h are outside the plane
            mw[k] = (idx < NW && gy < h && gx + 3 < g.pitch)
                        ? __ldg(reinterpret_cast<const unsigned *>(a.mask + (long)gy * g.pitch + gx)) : 0u;
        }
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            const int idx = tid + k * THREADS, t = idx / CFG::TROWS, rr = idx % CFG::TROWS;
            tv[k] = idx < NT ? __ldg(a.tab + (long)t * a.tab_h + ay + rr) : 0.0;  // the table carries slack rows past h
        }
#pragma unroll
        for (int k = 0; k < MW; ++k) {
            const int idx = tid + k * THREADS;
            if (idx < NW) reinterpret_cast<unsigned *>(mk)[idx] = mw[k];
        }
#pragma unroll
        for (int k = 0; k < MT; ++k) {
            const int idx = tid + k * THREADS;
            if (idx < NT) stab[idx] = tv[k];
        }
    }
    {
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
    }
    __syncthreads();

    const double *ssh = in, *sshp = in + TILE, *u = in + 2 * TILE, *up = in + 3 * TILE, *v = in + 4 * TILE,
                 *vp = in + 5 * TILE, *h_r = in + 6 * TILE, *mu = in + 7 * TILE;
    double *hu = s1, *hv = s1 + TILE, *hh = s1 + 2 * TILE, *vort = s1 + 3 * TILE, *str_t = s1 + 4 * TILE,
           *str_s = s1 + 5 * TILE;
    const MetRow mt{stab, CFG::TROWS, ay};

    // stage A on the tile grown by one cell: local columns 1..TX+2, rows 1..TY+2
    constexpr int AW = TX + 2, AH = TY + 2;
    for (int idx = tid; idx < AW * AH; idx += THREADS) {
        const int lx = 1 + idx % AW, ly = 1 + idx / AW;
        const int c = ly * IW + lx;
        const PrepOut o = prep_cell<TRANS, LAT>(c, ay + ly, IW, mt, a.ffs, mk, ssh, h_r, u, v, up, vp);
        hu[c] = o.hu; hv[c] = o.hv; hh[c] = o.hh; vort[c] = o.vort; str_t[c] = o.str_t; str_s[c] = o.str_s;
    }
    __syncthreads();

    // stage B: thread (tx, ty) updates the R vertically adjacent cells (tx, R*ty .. R*ty+R-1), so the
    // compiler shares their common neighbour loads and sub-expressions
    const int tx = tid & (TX - 1), ty = tid / TX;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int lx = HALO + tx, ly = HALO + ty * R + k;
        const int m = i0 + tx, n = j0 + ty * R + k;
        if (m > g.nx_end || n > n1) continue;
        const int c = ly * IW + lx;
        const long gc = ix(g, m, n);
        const double rhsx = HAS_RHS ? a.RHSx[gc] : 0.0, rhsy = HAS_RHS ? a.RHSy[gc] : 0.0;
        const double rdx = HAS_RDISS ? (double)(a.rdis[gc] + a.rdis[gc + 1]) : 0.0;
        const double rdy = HAS_RDISS ? (double)(a.rdis[gc] + a.rdis[gc + g.pitch]) : 0.0;
        const UpdOut o = update_cell<TRANS, LAT>(c, ay + ly, IW, mt, a.tau, a.ts, a.ffs, mk, ssh, sshp, u, up, v, vp, h_r, mu,
                                                 hu, hv, hh, vort, str_t, str_s, rhsx, rhsy, rdx, rdy);
        a.ssh_o[gc] = o.ssh; a.sshp_o[gc] = o.sshp; a.u_o[gc] = o.u; a.up_o[gc] = o.up; a.v_o[gc] = o.v; a.vp_o[gc] = o.vp;
        if (o.bad) atomicAdd(a.bad, 1);  // K11
    }
}

inline int launched(const char *what)
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SWCU_OK : cuda_fail(e, what);
}

template <bool ROW>
void prep_dispatch(const Geo &g, const FusedArgs &a, int n0, int n1, dim3 grid, dim3 block, cudaStream_t st)
{
    if (a.trans && a.lat) k_prep<ROW, true, true><<<grid, block, 0, st>>>(g, a, n0, n1);
    else if (a.trans) k_prep<ROW, true, false><<<grid, block, 0, st>>>(g, a, n0, n1);
    else if (a.lat) k_prep<ROW, false, true><<<grid, block, 0, st>>>(g, a, n0, n1);
    else k_prep<ROW, false, false><<<grid, block, 0, st>>>(g, a, n0, n1);
}

template <bool ROW, bool T, bool L>
void update_dispatch2(const Geo &g, const FusedArgs &a, int n0, int n1, dim3 grid, dim3 block, cudaStream_t st)
{
    const bool r = a.RHSx != nullptr, d = a.rdis != nullptr;
    if (r && d) k_update<ROW, T, L, true, true><<<grid, block, 0, st>>>(g, a, n0, n1);
    else if (r) k_update<ROW, T, L, true, false><<<grid, block, 0, st>>>(g, a, n0, n1);
    else if (d) k_update<ROW, T, L, false, true><<<grid, block, 0, st>>>(g, a, n0, n1);
    else k_update<ROW, T, L, false, false><<<grid, block, 0, st>>>(g, a, n0, n1);
}

template <bool ROW>
void update_dispatch(const Geo &g, const FusedArgs &a, int n0, int n1, dim3 grid, dim3 block, cudaStream_t st)
{
    if (a.trans && a.lat) update_dispatch2<ROW, true, true>(g, a, n0, n1, grid, block, st);
    else if (a.trans) update_dispatch2<ROW, true, false>(g, a, n0, n1, grid, block, st);
    else if (a.lat) update_dispatch2<ROW, false, true>(g, a, n0, n1, grid, block, st);
    else update_dispatch2<ROW, false, false>(g, a, n0, n1, grid, block, st);
}

}  // namespace

int launch_prep(const Geo &g, const FusedArgs &a, int n0, int n1, cudaStream_t st)
{
    if (n1 < n0) return SWCU_OK;
    const dim3 block(BX, BY, 1);
    const dim3 grid((unsigned)((g.nx_end - g.nx_start + 2 + BX) / BX), (unsigned)((n1 - n0 + BY) / BY), 1);
    if (a.tab) prep_dispatch<true>(g, a, n0, n1, grid, block, st);
    else prep_dispatch<false>(g, a, n0, n1, grid, block, st);
    return launched("prep");
}

int launch_update(const Geo &g, const FusedArgs &a, int n0, int n1, cudaStream_t st)
{
    if (n1 < n0) return SWCU_OK;
    const dim3 block(BX, BY, 1);
    const dim3 grid((unsigned)((g.nx_end - g.nx_start + BX) / BX), (unsigned)((n1 - n0 + BY) / BY), 1);
    if (a.tab) update_dispatch<true>(g, a, n0, n1, grid, block, st);
    else update_dispatch<false>(g, a, n0, n1, grid, block, st);
    return launched("update");
}

int launch_tracer(const Geo &g, const FusedArgs &a, int n0, int n1, cudaStream_t st)
{
    if (n1 < n0) return SWCU_OK;
    const dim3 block(BX, BY, 1);
    const dim3 grid((unsigned)((g.nx_end - g.nx_start + BX) / BX), (unsigned)((n1 - n0 + BY) / BY), 1);
    if (a.tab) k_tracer<true><<<grid, block, 0, st>>>(g, a, n0, n1);
    else k_tracer<false><<<grid, block, 0, st>>>(g, a, n0, n1);
    return launched("tracer");
}

namespace {
template <class CFG, bool T, bool L, bool R, bool D>
int step_launch(const StepMaps &maps, const Geo &g, const FusedArgs &a, int n0, int n1, cudaStream_t st)
{
    // the opt-in for > 48 KB of dynamic shared memory is per device: remember which devices have it
    static unsigned long long attr_set = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(__atomic_load_n(&attr_set, __ATOMIC_ACQUIRE) >> (dev & 63) & 1ull)) {   // contexts may step from several host threads
        cudaError_t e = cudaFuncSetAttribute(k_step<CFG, T, L, R, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)CFG::SMEM);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_step)");
        __atomic_fetch_or(&attr_set, 1ull << (dev & 63), __ATOMIC_RELEASE);
    }
    const dim3 grid((unsigned)((g.nx_end - g.nx_start + TX) / TX), (unsigned)((n1 - n0 + CFG::TY) / CFG::TY), 1);
    k_step<CFG, T, L, R, D><<<grid, CFG::THREADS, CFG::SMEM, st>>>(maps, g, a, n0, n1);
    return launched("step_tiled");
}

template <class CFG>
int step_dispatch(const StepMaps &maps, const Geo &g, const FusedArgs &a, int n0, int n1, cudaStream_t st)
{
    const bool r = a.RHSx != nullptr, d = a.rdis != nullptr;
    if (a.trans && a.lat && !r && !d) return step_launch<CFG, true, true, false, false>(maps, g, a, n0, n1, st);
    if (a.trans && a.lat && !r && d) return step_launch<CFG, true, true, false, true>(maps, g, a, n0, n1, st);
    if (a.trans && a.lat && r && d) return step_launch<CFG, true, true, true, true>(maps, g, a, n0, n1, st);
    if (a.trans && a.lat && r && !d) return step_launch<CFG, true, true, true, false>(maps, g, a, n0, n1, st);
    return -1;  // other flag combinations use the two-launch path
}
}  // namespace

// One launch for rows [n0..n1]: needs the per-row metric tables (a.tab) and the eight tensor maps.
// Returns -1 if this flag combination has no tiled instantiation (caller falls back to two launches).
int launch_step_tiled(const StepMaps &maps, const Geo &g, const FusedArgs &a, int n0, int n1, int variant, cudaStream_t st)
{
    if (n1 < n0) return SWCU_OK;
    switch (variant) {
        case 1: return step_dispatch<TileCfg<16, 2, 2>>(maps, g, a, n0, n1, st);
        case 2: return step_dispatch<TileCfg<16, 1, 2>>(maps, g, a, n0, n1, st);
        case 3: return step_dispatch<TileCfg<8, 1, 3>>(maps, g, a, n0, n1, st);
        case 4: return step_dispatch<TileCfg<8, 2, 4>>(maps, g, a, n0, n1, st);
        case 5: return step_dispatch<TileCfg<16, 4, 2>>(maps, g, a, n0, n1, st);
        default: return step_dispatch<TileCfg<8, 2, 4>>(maps, g, a, n0, n1, st);
    }
}

// (the mask tile is staged with 4-byte loads: tile rows must start 4-byte aligned in the mask plane)
bool step_tiled_supported(const Geo &g, const FusedArgs &a)
{
    return a.trans && a.lat && a.tab != nullptr && (g.nx_start - HALO - g.bx1) % 4 == 0;
}

namespace {
int variant_ty(int variant) { return (variant == 3 || variant == 4 || variant < 1 || variant > 5) ? 8 : 16; }

// one warp per tile: OR of the lu bits of the tile's output cells
__global__ void k_tile_land(Geo g, const unsigned char *__restrict__ mask, int ty, int ntx, int nty, int n0, int n1,
                            unsigned char *__restrict__ tile_land)
{
    const int t = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (t >= ntx * nty) return;
    const int bx = t % ntx, by = t / ntx;
    const int m = g.nx_start + bx * TX + lane;
    int any = 0;
    for (int k = 0; k < ty; ++k) {
        const int n = n0 + by * ty + k;
        if (m <= g.nx_end && n <= n1) any |= mask[ix(g, m, n)] & MB_LU;
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) tile_land[t] = any ? 0 : 1;
}
}  // namespace

void step_tile_grid(const Geo &g, int variant, int n0, int n1, int *ntx, int *nty)
{
    const int ty = variant_ty(variant);
    *ntx = (g.nx_end - g.nx_start + TX) / TX;
    *nty = n1 >= n0 ? (n1 - n0 + ty) / ty : 0;
}

int launch_tile_land(const Geo &g, const unsigned char *mask, int variant, int n0, int n1, unsigned char *tile_land,
                     cudaStream_t st)
{
    int ntx = 0, nty = 0;
    step_tile_grid(g, variant, n0, n1, &ntx, &nty);
    const int tiles = ntx * nty;
    if (tiles == 0) return SWCU_OK;
    k_tile_land<<<(unsigned)((tiles + 7) / 8), 256, 0, st>>>(g, mask, variant_ty(variant), ntx, nty, n0, n1, tile_land);
    return launched("tile_land");
}

// TMA box (columns, rows) of a tile variant
void step_tile_box(int variant, int *box_w, int *box_h)
{
    *box_w = IW;
    *box_h = (variant == 3 || variant == 4 ? 8 : 16) + 2 * HALO;
}

// ---- per-row metric tables ----------------------------------------------------------------------
namespace {

// One thread per row: the table entries, evaluated exactly like MetGen does for column nx_start.
__global__ void k_build_tables(Geo g, MetGen mg, double *__restrict__ tab, int h)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= h) return;
    const long c = (long)r * g.pitch + (g.nx_start - g.bx1);
    tab[T_DX * h + r] = mg.dx(c, r);     tab[T_DY * h + r] = mg.dy(c, r);
    tab[T_DXT * h + r] = mg.dxt(c, r);   tab[T_DYT * h + r] = mg.dyt(c, r);
    tab[T_DXH * h + r] = mg.dxh(c, r);   tab[T_DYH * h + r] = mg.dyh(c, r);
    tab[T_DXB * h + r] = mg.dxb(c, r);   tab[T_DYB * h + r] = mg.dyb(c, r);
    tab[T_RLH * h + r] = mg.rlh(c, r);
    tab[T_AREA * h + r] = mg.area(c, r);
    tab[T_DY2 * h + r] = mg.dy2(c, r);   tab[T_DX2 * h + r] = mg.dx2(c, r);
    tab[T_DXB2 * h + r] = mg.dxb2(c, r); tab[T_DYB2 * h + r] = mg.dyb2(c, r);
    tab[T_RYX * h + r] = mg.ryx(c, r);   tab[T_RXY * h + r] = mg.rxy(c, r);
    tab[T_RXYB * h + r] = mg.rxyb(c, r); tab[T_RYXB * h + r] = mg.ryxb(c, r);
    // correctly rounded reciprocals for mdiv() (IEEE division, done once per row)
    tab[T_RDXT * h + r] = 1.0 / mg.dxt(c, r); tab[T_RDYT * h + r] = 1.0 / mg.dyt(c, r);
    tab[T_RDXH * h + r] = 1.0 / mg.dxh(c, r); tab[T_RDYH * h + r] = 1.0 / mg.dyh(c, r);
    tab[T_RDXB * h + r] = 1.0 / mg.dxb(c, r); tab[T_RDYB * h + r] = 1.0 / mg.dyb(c, r);
    tab[T_RAREA * h + r] = 1.0 / mg.area(c, r);
}

// *nonrow += number of cells in columns [bx1+1 .. bx2-1] whose nine real(4) metric values differ
// (bitwise) from the value in column nx_start of the same row.  The outermost columns are the
// reference's never-read zero frame and are excluded.
__global__ void k_check_rowconst(Geo g, const float *const *arrs, int narr, int *nonrow)
{
    const int m = g.bx1 + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (m > g.bx2 - 1) return;
    const long c = (long)r * g.pitch + (m - g.bx1);
    const long c0 = (long)r * g.pitch + (g.nx_start - g.bx1);
    int bad = 0;
    for (int k = 0; k < narr; ++k)
        bad |= __float_as_int(arrs[k][c]) != __float_as_int(arrs[k][c0]);
    if (bad) atomicAdd(nonrow, 1);
}

}  // namespace

int launch_build_tables(const Geo &g, const FusedArgs &a, double *tab, int h, int *nonrow_dev,
                        const float *const *arr_list_dev, cudaStream_t st)
{
    const MetGen mg{a.dx, a.dy, a.dxt, a.dyt, a.dxh, a.dyh, a.dxb, a.dyb, a.rlh_s};
    k_build_tables<<<(unsigned)((h + 127) / 128), 128, 0, st>>>(g, mg, tab, h);
    const int wcols = g.bx2 - g.bx1 - 1;
    k_check_rowconst<<<dim3((unsigned)((wcols + 255) / 256), (unsigned)h, 1), 256, 0, st>>>(g, arr_list_dev, 9, nonrow_dev);
    return launched("build_tables");
}

// ---- output record (control/output.f90:101-135, tools/io.f90:343-348) --------------------------
namespace {
__global__ void k_output_record(Geo g, const double *__restrict__ f, const unsigned char *__restrict__ bits,
                                const float *__restrict__ lu, float *__restrict__ out)
{
    const int m = g.nx_start + blockIdx.x * blockDim.x + threadIdx.x;
    const int n = g.ny_start + blockIdx.y;
    if (m > g.nx_end) return;
    const long c = ix(g, m, n);
    const bool sea = bits ? (bits[c] & MB_LU) != 0 : !(fabsf(lu[c]) < 0.5f);
    const long o = (long)(n - g.ny_start) * (g.nx_end - g.nx_start + 1) + (m - g.nx_start);
    out[o] = sea ? (float)f[c] : -1.0e32f;   // real(x, wp4) rounds to nearest, like cvt.rn.f32.f64
}
}  // namespace

int launch_output_record(const Geo &g, const double *field, const unsigned char *mask_bits, const float *lu,
                         float *out, cudaStream_t st)
{
    const dim3 grid((unsigned)((g.nx_end - g.nx_start + 256) / 256), (unsigned)(g.ny_end - g.ny_start + 1), 1);
    k_output_record<<<grid, 256, 0, st>>>(g, field, mask_bits, lu, out);
    return launched("output_record");
}

// ---- self-test of mdiv ---------------------------------------------------------------------------
namespace {
__device__ __forceinline__ unsigned long long splitmix(unsigned long long &x)
{
    unsigned long long z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// a double with random sign and mantissa and an exponent in [-span, span]
__device__ __forceinline__ double rnd_double(unsigned long long &st, int span)
{
    const unsigned long long r = splitmix(st);
    const long long e = 1023 + (long long)(splitmix(st) % (unsigned long long)(2 * span + 1)) - span;
    return __longlong_as_double((long long)((r & 0x800fffffffffffffull) | ((unsigned long long)e << 52)));
}
__global__ void k_selftest_mdiv(long n, unsigned long long seed, unsigned long long *bad)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long st = seed + 0x632be59bd9b4e019ull * (unsigned long long)(i + 1);
    double a = rnd_double(st, 200);
    const unsigned sel = (unsigned)(splitmix(st) & 7u);
    if (sel == 0) a = 0.0;
    if (sel == 1) a = -0.0;
    double b;
    if (splitmix(st) & 1ull) b = (double)(float)rnd_double(st, 60);   // like a promoted real(4) metric
    else b = rnd_double(st, 200);
    if (splitmix(st) & 1ull) b = fabs(b);
    const double y = 1.0 / b;
    const double q = mdiv(a, b, y), want = a / b;
    if (__double_as_longlong(q) != __double_as_longlong(want)) atomicAdd(bad, 1ull);
}
}  // namespace

int launch_selftest_mdiv(long n, unsigned long long seed, unsigned long long *bad_dev, cudaStream_t st)
{
    k_selftest_mdiv<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, seed, bad_dev);
    return launched("selftest_mdiv");
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
    return launched("mask_set");
}
int launch_mask_get(long total, float *dst, const unsigned char *bits, int bit, cudaStream_t st)
{
    k_mask_get<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, dst, bits, bit);
    return launched("mask_get");
}

}  // namespace swcu
