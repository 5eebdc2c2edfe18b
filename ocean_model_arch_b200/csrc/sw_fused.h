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
    int ncol, nbands, nwarps;
    const unsigned char *band_land;  // [nwarps]: 1 <=> the warp's output cells are all land (skipped), or nullptr
    int minb;                        // register budget variant: sized for 2 or 3 CTAs per SM
};
bool march_supported(const Geo &g, const FusedArgs &a);
void march_plan(const Geo &g, int n0, int n1, int max_warps, MarchPlan *pl);
int march_resident_warps(int device, int minb);  // SMs x resident warps of k_march: the size of one full wave
int launch_march(const Geo &g, const FusedArgs &a, const MarchPlan &pl, cudaStream_t st);
int launch_build_fast(const double *tab, int h, double tau, double *fc, cudaStream_t st);
int launch_band_land(const Geo &g, const unsigned char *mask, const MarchPlan &pl, unsigned char *out, cudaStream_t st);

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
