// sw_fused.h -- argument block and launchers of the fused Level-B kernels (sw_kernels_fused.cu)
#pragma once
#include <cuda.h>

#include "sw_cells.cuh"
#include "sw_common.h"

namespace swcu {

// TMA descriptors of the eight arrays the tiled kernel loads: ssh, sshp, u, up, v, vp, hhq_rest, mu
struct StepMaps { CUtensorMap m[8]; };

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
    const double *tab;  // per-row metric tables [T_COUNT][tab_h], nullptr <=> use the 2-D real(4) arrays
    int tab_h;
    const double *fc;   // tolerance mode: per-row coefficient table [tab_h][FC_STRIDE] (sw_fast.cuh), else nullptr
    const double *ft;   //   ... and the tracer step's [tab_h][FT_STRIDE]
    const unsigned char *mask;
    // one byte per tile of the main (interior / full-range) tiled launch: 1 <=> every output cell of the tile is land, so
    // the tile is skipped (land cells never change; both ping-pong buffers already hold them).
    // nullptr for strip launches / other paths.  The B200 analogue of the reference's land-block
    // skipping (core/decomposition.f90:508-512 counts all-land blocks).
    const unsigned char *tile_land;
    int *bad;  // K11 counter
    Tau tau;
    double ts, ffs;
    int trans, lat;
    // tracer (ff1): level-n arrays read, n+1 written (ping-pong); nullptr without tracers
    const double *ff, *ffp;
    double *ff_o, *ffp_o;
};

// Geometry of one k_march launch (sw_kernels_march.cu): warp w works on warp column w % ncol (28 output
// columns) and on band w / ncol of the rows [n0..n1] (bands of equal height, +-1 row).
struct MarchPlan {
    int n0, n1;
    int ncol, nbands, nwarps;        // nwarps = ncol * nbands
    // The last `late_hi` bands are `late_cut` rows shorter than the others: with a fused halo push the strip
    // kernel runs concurrently and holds a few CTA slots for the first microseconds, so the CTAs of the main
    // launch that are scheduled last start that much later.  (late_lo: the same for band 0; unused.)
    int late_lo, late_hi, late_cut;
    const unsigned char *band_land;  // [nwarps]: 1 <=> the warp's output cells are all land (skipped), or nullptr
    int minb;                        // register budget variant: sized for 2 or 3 CTAs per SM
};
// Halo push fused into k_march (neighbouring blocks in other processes of the node, peer-mapped memory):
// a small k_march<PEER> launch on the high-priority stream, concurrent with the lean main launch (disjoint rows):
// its warps wait for the neighbour's rows of the previous step, compute the boundary strip of their side, store
// it ALSO into that neighbour's halo rows over NVLink, and the last of them publishes `tick` in the neighbour's
// "halo ready" word.
// Side 0 = the block below (rank-1), 1 = the block above.  A side is active iff out[side][0] != nullptr.
struct MarchPeer {
    int lo0, lo1, hi0, hi1;              // rows of the lower / upper strip (inclusive; empty if x1 < x0)
    double *out[2][6];                   // the neighbours' write planes, shifted so that MY element index applies
    unsigned long long *ready[2];        // the neighbours' "halo ready" words (peer-mapped)
    const unsigned long long *ready_in[2];  // MY "halo ready" words: the neighbour's push of step tick-1 has landed
    const unsigned long long *free_[2];  // my "your rows in my write buffers may be overwritten" words
    unsigned long long tick;
    unsigned *count[2];                  // strip-warp counters (zero between launches)
    int *timeout;                        // raised when a neighbour's flag never arrived (device int)
    int dbg;                             // timing experiments only (SWCU_PEER_DBG): 1 = no wait for "free", 2 = no peer stores
};
// rows [*bs .. *be] of band `band`
__host__ __device__ inline int march_band_start(const MarchPlan &pl, int band)
{
    const long R = (long)(pl.n1 - pl.n0 + 1) + (long)pl.late_cut * (pl.late_lo + pl.late_hi);
    const int first_late = pl.nbands - pl.late_hi;
    return pl.n0 + (int)((long)band * R / pl.nbands) - (band >= 1 ? pl.late_lo * pl.late_cut : 0) -
           (band > first_late ? (band - first_late) * pl.late_cut : 0);
}
__host__ __device__ inline void march_band_rows(const MarchPlan &pl, int band, int *bs, int *be)
{
    *bs = march_band_start(pl, band);
    *be = march_band_start(pl, band + 1) - 1;
}
bool march_supported(const Geo &g, const FusedArgs &a);
void march_plan(const Geo &g, int n0, int n1, int max_warps, MarchPlan *pl, int band_rows = 0);
int march_resident_warps(int device, int minb);  // SMs x resident warps of k_march: the size of one full wave
int march_tracer_resident_warps(int device);
int launch_march(const Geo &g, const FusedArgs &a, const MarchPlan &pl, cudaStream_t st, const MarchPeer *peer = nullptr);
int march_strip_warps(const Geo &g);  // warps one boundary strip takes (= warp columns)
int launch_build_fast(const double *tab, int h, double tau, double *fc, double *ft, cudaStream_t st);
int launch_tracer_march(const Geo &g, const FusedArgs &a, const MarchPlan &pl, cudaStream_t st);
int launch_band_land(const Geo &g, const unsigned char *mask, const MarchPlan &pl, unsigned char *out, int *nland_dev,
                     cudaStream_t st);

// prep on rows [n0..n1] (columns nx_start-1 .. nx_end+1); update on rows [n0..n1] (columns of S)
int launch_prep(const Geo &g, const FusedArgs &a, int n0, int n1, cudaStream_t st);
int launch_update(const Geo &g, const FusedArgs &a, int n0, int n1, cudaStream_t st);
// the whole step for rows [n0..n1] in ONE launch (TMA-staged shared-memory tiles); needs a.tab
int launch_step_tiled(const StepMaps &maps, const Geo &g, const FusedArgs &a, int n0, int n1, int variant,
                      cudaStream_t st);
bool step_tiled_supported(const Geo &g, const FusedArgs &a);
void step_tile_box(int variant, int *box_w, int *box_h);
// tile grid (columns, rows of tiles) of the full-range launch, and the all-land flags for it
void step_tile_grid(const Geo &g, int variant, int n0, int n1, int *ntx, int *nty);
int launch_tile_land(const Geo &g, const unsigned char *mask, int variant, int n0, int n1, unsigned char *tile_land,
                     cudaStream_t st);
// Builds the per-row tables from column nx_start of the nine real(4) arrays and counts (into
// *nonrow_dev) the cells whose values differ from their row's entry.
int launch_build_tables(const Geo &g, const FusedArgs &a, double *tab, int h, int *nonrow_dev,
                        const float *const *arr_list_dev, cudaStream_t st);
// expl_tracer for rows [n0..n1] in one launch; reads a.ssh_o .. a.v_o (the state the step just wrote)
int launch_tracer(const Geo &g, const FusedArgs &a, int n0, int n1, cudaStream_t st);
// fp32 output record of the interior with undef on land (either a mask byte plane or a real(4) lu)
int launch_output_record(const Geo &g, const double *field, const unsigned char *mask_bits, const float *lu,
                         float *out, cudaStream_t st);
int launch_selftest_mdiv(long n, unsigned long long seed, unsigned long long *bad_dev, cudaStream_t st);
int launch_mask_set(long total, const float *src, unsigned char *bits, int bit, cudaStream_t st);
int launch_mask_get(long total, float *dst, const unsigned char *bits, int bit, cudaStream_t st);

// halo rows pushed into the neighbours' buffers over peer memory (sw_init.cu)
struct PushArgs {
    const double *src[2][8];       // [side][array]: first row to send
    double *dst[2][8];             // [side][array]: where it lands in the neighbour's plane (peer-mapped)
    unsigned long long *flag[2];   // the neighbours' "halo ready" words (peer-mapped), or nullptr
    unsigned long long value;
    long count;                    // doubles per array and side
    unsigned *counter;             // zero-initialised CTA counter
};
int launch_push_halo(const PushArgs &p, int narrays, cudaStream_t st);
int launch_signal(unsigned long long *a, unsigned long long *b, unsigned long long value, cudaStream_t st);

// device-side input construction (sw_init.cu)
int launch_init_lu(const Geo &g, int w, int h, int nx, int ny, const int *land_dev, unsigned char *lu_dev, cudaStream_t st);
int launch_init_masks(const Geo &g, int w, int h, const unsigned char *lu_dev, unsigned char *bits, float *const f[7],
                      cudaStream_t st);
int launch_expand_rows(const Geo &g, int w, int h, int nx, float *const dst[9], const float *prof_dev, cudaStream_t st);
int launch_fill8(int w, int h, int pitch, double *dst, double value, cudaStream_t st);
int launch_fill4(int w, int h, int pitch, float *dst, float value, cudaStream_t st);

}  // namespace swcu
