/*
 * swcuda.h -- C ABI of libswcuda.so: the B200-native (sm_100a) kernel layer for the fp64
 * shallow-water time step of Andrcraft9/ocean_model_arch.
 *
 * This is the drop-in boundary.  A reference maintainer binds these entry points from Fortran
 * with iso_c_binding (see INTEGRATION.md and fortran/sw_interface_cuda.f90) in place of the
 * per-block kernel envokes of interface/shallow_water/sw_interface.f90 and the CUDA-Fortran
 * launch wrappers of gpu/interface/sw_interface_gpu.f90.  All `file:line` citations are relative
 * to the reference repository root.
 *
 * Conventions
 *   - plain C types only; `int` is 32-bit like Fortran's default integer;
 *   - arrays are the reference's explicit-shape block arrays A(bnd_x1:bnd_x2, bnd_y1:bnd_y2),
 *     column-major, m (x) contiguous: element (m,n) at
 *     base[(n-bnd_y1)*(bnd_x2-bnd_x1+1) + (m-bnd_x1)]  (core/decomposition.f90:493-503);
 *   - real(8) -> double, real(4) -> float (masks, metrics, Coriolis, friction are real(4));
 *   - every function returns SWCU_OK (0) or an error code; swcu_last_error() gives the text.
 *     The reference ignores CUDA istat and aborts through abort_model (shared/errors.f90:30-37);
 *     the Fortran shim calls abort_model on any non-zero return;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     SWCU_ERR_CUDA.
 */
#ifndef SWCUDA_H
#define SWCUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define SWCU_OK 0
#define SWCU_ERR_CUDA 1   /* CUDA runtime error (text in swcu_last_error) */
#define SWCU_ERR_ARG 2    /* bad argument / unknown field id */
#define SWCU_ERR_NCCL 3   /* NCCL error or NCCL library not loadable */
#define SWCU_ERR_STATE 4  /* call not valid in the context's current mode/state */
#define SWCU_ERR_BLOWUP 5 /* check_ssh_err found |ssh| >= 1e4 or NaN on a sea cell */

/* The eight leading integer arguments of every reference kernel
 * (e.g. kernel/shallow_water/vel_ssh.f90:69-72), i.e. domain%bnx_start(k) ... domain%bbnd_y2(k)
 * as passed by the binders (interface/shallow_water/sw_interface.f90:46-47). */
typedef struct swcu_dims {
    int nx_start, nx_end, ny_start, ny_end;
    int bnd_x1, bnd_x2, bnd_y1, bnd_y2;
} swcu_dims;

const char *swcu_last_error(void);
int swcu_version(void);
/* Number of visible CUDA devices (0 when there is none); never fails. */
int swcu_device_count(void);

/* ------------------------------------------------------------------------------------------
 * Level A -- stateless 1:1 kernels on DEVICE pointers, reference argument order.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).  Launches are
 * asynchronous.  Each replaces one CPU kernel / one CUDA-Fortran kernel of the reference:
 * ------------------------------------------------------------------------------------------ */

/* K1  kernel/shallow_water/vel_ssh.f90:69-106  (gpu/kernel/vel_ssh_gpu.f90:24-61) */
int swcu_sw_update_ssh_kernel(const swcu_dims *d, double tau,
        const float *lu, const float *dx, const float *dy, const float *dxh, const float *dyh,
        const double *hhu, const double *hhv, double *sshn, const double *sshp,
        const double *ubrtr, const double *vbrtr, void *stream);

/* K7  kernel/shallow_water/vel_ssh.f90:108-195  (gpu/kernel/vel_ssh_gpu.f90:63-153) */
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
        const double *RHSx_dif, const double *RHSy_dif, void *stream);

/* K8  kernel/shallow_water/vel_ssh.f90:197-245  (gpu/kernel/vel_ssh_gpu.f90:155-203) */
int swcu_sw_next_step(const swcu_dims *d, double time_smooth,
        const float *lu, const float *lcu, const float *lcv,
        double *ssh, double *sshn, double *sshp,
        double *ubrtr, double *ubrtrn, double *ubrtrp,
        double *vbrtr, double *vbrtrn, double *vbrtrp, void *stream);

/* K3  kernel/shallow_water/vel_ssh.f90:247-281  (nlev = 1) */
int swcu_uv_trans_vort_kernel(const swcu_dims *d, const float *luu,
        const float *dxt, const float *dyt, const float *dxb, const float *dyb,
        const double *u, const double *v, double *vort, void *stream);

/* K4  kernel/shallow_water/vel_ssh.f90:283-373  (nlev = 1; hq is accepted and unused) */
int swcu_uv_trans_kernel(const swcu_dims *d, const float *lcu, const float *lcv, const float *luu,
        const float *dxh, const float *dyh, const double *u, const double *v, const double *vort,
        const double *hq, const double *hu, const double *hv, const double *hh,
        double *RHSx, double *RHSy, void *stream);

/* K6  kernel/shallow_water/vel_ssh.f90:375-452  (nlev = 1; hu, hv accepted and unused) */
int swcu_uv_diff2_kernel(const swcu_dims *d, const float *lcu, const float *lcv,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        const double *mu, const double *str_t, const double *str_s,
        const double *hq, const double *hu, const double *hv, const double *hh,
        double *RHSx, double *RHSy, void *stream);

/* K5  kernel/shallow_water/mixing.f90:14-58  (nlev = 1) */
int swcu_stress_components_kernel(const swcu_dims *d, const float *lu, const float *luu,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        const double *u, const double *v, double *str_t, double *str_s, void *stream);

/* K10 kernel/shallow_water/depth.f90:14-99.  full_free_surface is a module variable in the
 * reference (config_sw_module); here it is an argument.  sh, shp, h_r are read-only. */
int swcu_hh_init_kernel(const swcu_dims *d, int full_free_surface,
        const float *lu, const float *llu, const float *llv, const float *luh,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        double *hq, double *hqp, double *hqn, double *hu, double *hup, double *hun,
        double *hv, double *hvp, double *hvn, double *hh, double *hhp, double *hhn,
        const double *sh, const double *shp, const double *h_r, void *stream);

/* K2  kernel/shallow_water/depth.f90:101-162 */
int swcu_hh_update_kernel(const swcu_dims *d,
        const float *lu, const float *llu, const float *llv, const float *luh,
        const float *dx, const float *dy, const float *dxt, const float *dyt,
        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
        double *hqn, double *hun, double *hvn, double *hhn,
        const double *sh, const double *h_r, void *stream);

/* K9  kernel/shallow_water/depth.f90:164-211.  time_smooth is a module variable in the
 * reference; here it is an argument. */
int swcu_hh_shift_kernel(const swcu_dims *d, double time_smooth,
        const float *lu, const float *llu, const float *llv, const float *luh,
        double *hq, double *hqp, double *hqn, double *hu, double *hup, double *hun,
        double *hv, double *hvp, double *hvn, double *hh, double *hhp, double *hhn, void *stream);

/* K11 kernel/shallow_water/vel_ssh.f90:40-67.  Instead of abort_model the number of offending
 * sea cells is ADDED to *bad_count (a device int the caller zeroes). */
int swcu_check_ssh_err_kernel(const swcu_dims *d, const float *lu, const double *ssh,
        int *bad_count, void *stream);

/* tracer kernels, kernel/tracer/leapfrog_tracer.f90:13-98, 100-141, 143-170 */
int swcu_tran_diff_fluxes_kernel(const swcu_dims *d, const float *lcu, const float *lcv,
        const float *dxt, const float *dyt, const float *dxh, const float *dyh,
        const double *hhu, const double *hhv, const double *ff, const double *ffp,
        const double *uu, const double *vv, const double *mu, double factor_mu,
        double *flux_x, double *flux_y, void *stream);
int swcu_tran_diff_tracer_kernel(const swcu_dims *d, const float *lu, const float *dx, const float *dy,
        double tau, const double *hhqn, const double *hhqp,
        const double *flux_x, const double *flux_y, const double *ffp, double *ffn, void *stream);
int swcu_tracer_next_step_kernel(const swcu_dims *d, double time_smooth, const float *lu,
        const double *ffn, double *ffp, double *ff, void *stream);

/* ------------------------------------------------------------------------------------------
 * Level B -- resident context: one block of the reference's decomposition lives on one GPU.
 * Replaces field_device storage (core/data_types.f90:56-59), init_device_data
 * (control/init_data.f90:127-145), expl_shallow_water_gpu (control/shallow_water/
 * shallow_water.f90:184-251), expl_tracer and the mode-5 host-staged halo syncs
 * (shared/mpp/sync.f90:378-536).  The library owns all device memory.
 * ------------------------------------------------------------------------------------------ */

/* field ids: every array of ocean_type (core/ocean.f90:14-48) and grid_type (core/grid.f90:23-90)
 * that the shallow-water / tracer step touches */
enum swcu_field {
    /* real(8) */
    SWCU_F_SSH = 0, SWCU_F_SSHN, SWCU_F_SSHP,
    SWCU_F_UBRTR, SWCU_F_UBRTRN, SWCU_F_UBRTRP,
    SWCU_F_VBRTR, SWCU_F_VBRTRN, SWCU_F_VBRTRP,
    SWCU_F_RHSX, SWCU_F_RHSY, SWCU_F_RHSX_ADV, SWCU_F_RHSY_ADV, SWCU_F_RHSX_DIF, SWCU_F_RHSY_DIF,
    SWCU_F_MU, SWCU_F_STR_T, SWCU_F_STR_S, SWCU_F_VORT,
    SWCU_F_HHQ_REST, SWCU_F_HHQ, SWCU_F_HHQ_P, SWCU_F_HHQ_N,
    SWCU_F_HHU, SWCU_F_HHU_P, SWCU_F_HHU_N, SWCU_F_HHV, SWCU_F_HHV_P, SWCU_F_HHV_N,
    SWCU_F_HHH, SWCU_F_HHH_P, SWCU_F_HHH_N,
    SWCU_F_FLUX_X, SWCU_F_FLUX_Y, SWCU_F_FF1, SWCU_F_FF1N, SWCU_F_FF1P,
    SWCU_NF8,
    /* real(4) */
    SWCU_F_LU = 100, SWCU_F_LUU, SWCU_F_LUH, SWCU_F_LCU, SWCU_F_LCV, SWCU_F_LLU, SWCU_F_LLV,
    SWCU_F_DX, SWCU_F_DY, SWCU_F_DXT, SWCU_F_DYT, SWCU_F_DXH, SWCU_F_DYH, SWCU_F_DXB, SWCU_F_DYB,
    SWCU_F_RLH_S, SWCU_F_R_DISS,
    SWCU_F4_END
};
#define SWCU_NF4 (SWCU_F4_END - 100)

/* step structure */
#define SWCU_MODE_REFERENCE 0 /* the reference's 11-kernel sequence, one launch per kernel (K1..K11) */
#define SWCU_MODE_FUSED 1     /* the whole step in ONE launch (K1..K11 fused: k_march, or k_step with "exact" = 1), or
                               * prep + update launches when the grid's metrics vary along x; + one launch per tracer */

typedef struct swcu_params {
    int full_free_surface, trans_terms, ksw_lat; /* sw.par 1-3 (configs/sw.f90:34-36) */
    double time_smooth;                          /* sw.par 4 */
    int use_tracers;                             /* sw.par 6 (0/1; one tracer field, ff1(1)) */
    int mode;                                    /* SWCU_MODE_* */
} swcu_params;

typedef struct swcu_ctx swcu_ctx;

/* Creates the context on CUDA device `device` for one block.  All fields start zero-filled, like
 * the reference's allocations (core/data_types.f90:529). */
int swcu_create(swcu_ctx **out, const swcu_dims *dims, const swcu_params *params, int device);
int swcu_destroy(swcu_ctx *ctx);

/* Host <-> device copies of one whole block array in the reference layout (host pointer =
 * block(k)%field).  Replaces sync_host_device (core/data_types.f90:934-950).  In FUSED mode the
 * derived fields (hh*, vort, str_*, RHS*_adv, RHS*_dif, sshn/ubrtrn/vbrtrn) are recomputed on demand by
 * swcu_download from the resident state, giving the values the reference holds after the same
 * number of steps. */
int swcu_upload(swcu_ctx *ctx, int field, const void *host);
int swcu_download(swcu_ctx *ctx, int field, void *host);
/* Upload of `nrows` consecutive rows starting at 0-based array row `first_row` (reference row
 * n = bnd_y1 + first_row); host points at nrows x (bnd_x2-bnd_x1+1) contiguous values.  Lets a host
 * with less memory than the device stream a block in stripes. */
int swcu_upload_rows(swcu_ctx *ctx, int field, const void *host, int first_row, int nrows);
/* The record local_output writes for a field (control/output.f90:101-135, tools/io.f90:325-352):
 * the block interior nx_start..nx_end x ny_start..ny_end, converted to real(4)
 * (copy_from_real8, core/data_types.f90:438-452), with undef = -1.0e32 (shared/system.f90:10) where
 * |lu| < 0.5, m fastest.  Built on the device; `host` receives (nx_end-nx_start+1)*(ny_end-ny_start+1)
 * floats -- a quarter of the bytes of the full real(8) block the reference copies back first
 * (model.f90:181).  real(8) fields only. */
int swcu_output_record(swcu_ctx *ctx, int field, float *host);
/* Same, but the other side is a DEVICE pointer in the reference layout (for callers that keep
 * their own device arrays, e.g. CUDA-Fortran field_device). */
int swcu_upload_from_device(swcu_ctx *ctx, int field, const void *dev);
int swcu_download_to_device(swcu_ctx *ctx, int field, void *dev);

/* Tuning / testing knobs.  "metric_tables": 1 (default) = use per-row tables of the metric arrays
 * when every one of them is constant along m (carthesian and unrotated spherical grids), 0 = always
 * read the 2-D real(4) arrays.  "tiled": 1 (default) = with metric tables in use, run the whole
 * step as ONE launch of the TMA-staged shared-memory kernel; 0 = two launches (prep + update) from
 * global memory.  "land_skip": 1 (default) = CTAs of the tiled kernel whose 32x8 output cells are all
 * land exit before loading anything (flags rebuilt after a mask upload), 0 = compute every tile.
 * "tile_variant": 1..5 selects the tile shape / cells per thread of the tiled kernel (default 4:
 * 32x8 cells, 128 threads, 2 cells per thread, 4 CTAs per SM).  Results are bitwise identical in every
 * combination.
 * "exact": 0 (default; the environment variable SWCU_EXACT=1 changes the default of new contexts) = the fused
 * step evaluates the reference's scheme with re-associated arithmetic (k_march: products of reciprocals as
 * per-row coefficients, shared face fluxes, explicit fma) -- ssh / u / v within a relative L2 of 1e-12 of the
 * reference's CPU path after 1000 steps (measured ~1e-15), masked-out and land cells bit-exact, results
 * independent of the block decomposition; 1 = every expression in the reference's operation order, bitwise equal
 * to the CPU path (k_step).  "march_minb": 2 (default) or 3, the register budget variant of k_march.
 * "tracer_num": number of tracer fields (sw.par line 7, control/tracer.f90:42; default 1; set once after
 * swcu_create with use_tracers = 1, before uploads of tracers 2..n); "tracer_select": the tracer (0-based) that
 * the field ids SWCU_F_FF1 / FF1N / FF1P address in uploads, downloads, fills and output records.  A step
 * advances every tracer.  The peer-memory halo path carries one tracer; several need a communicator. */
int swcu_set_option(swcu_ctx *ctx, const char *name, int value);
/* 1 if the last step used the per-row metric tables, 0 if it read the 2-D arrays. */
int swcu_uses_metric_tables(const swcu_ctx *ctx);

/* The envoke(hh_init) of init_ocean_data (control/init_data.f90:60-63): K10 + its halo sync on the
 * resident arrays, to be called once after the initial uploads.  A no-op in FUSED mode, where the
 * depth fields are functions of the resident state. */
int swcu_envoke_hh_init(swcu_ctx *ctx);

/* Kernel ids for the per-kernel envokes below: the (kernel, sync) pairs expl_shallow_water and
 * expl_tracer hand to envoke (control/shallow_water/shallow_water.f90:36-92, control/tracer.f90:50-60). */
enum swcu_kernel {
    SWCU_K_SW_UPDATE_SSH = 1, SWCU_K_HH_UPDATE, SWCU_K_UV_TRANS_VORT, SWCU_K_UV_TRANS,
    SWCU_K_STRESS_COMPONENTS, SWCU_K_UV_DIFF2, SWCU_K_SW_UPDATE_UV, SWCU_K_SW_NEXT_STEP,
    SWCU_K_HH_SHIFT, SWCU_K_HH_INIT, SWCU_K_CHECK_SSH_ERR,
    SWCU_K_TRAN_DIFF_FLUXES, SWCU_K_TRAN_DIFF_TRACER, SWCU_K_TRACER_NEXT_STEP
};
/* envoke_<name>_kernel(k, param) / envoke_<name>_sync(k, sync_parameters) of
 * interface/shallow_water/sw_interface.f90:42-408 and interface/tracer/tracer_interface.f90:28-102 on
 * the context's RESIDENT arrays (SWCU_MODE_REFERENCE): the binder's choice of arrays, then the 1:1
 * kernel; the sync exchanges the fields that binder's sync lists (width 1) with the neighbouring
 * blocks when a communicator is attached, else it is a no-op like a one-block hybrid_sync.  With these
 * two the reference's algorithm layer can keep its own envoke(sub_kernel, sub_sync, parameters)
 * sequence unchanged and still run on the device. */
int swcu_envoke_kernel(swcu_ctx *ctx, int kernel_id, double tau);
int swcu_envoke_sync(swcu_ctx *ctx, int kernel_id);

/* nsteps x expl_shallow_water(tau) [+ expl_tracer(tau) when use_tracers], asynchronous on the
 * context's stream; no host round trip.  With a communicator attached (below) every step
 * exchanges halos with the neighbouring blocks. */
int swcu_step(swcu_ctx *ctx, double tau, int nsteps);
/* Waits for the context's streams; returns SWCU_ERR_BLOWUP if K11 flagged cells since the last
 * call (their count in *bad_cells when non-NULL). */
int swcu_synchronize(swcu_ctx *ctx, long *bad_cells);

/* Timing helpers on the context's stream (CUDA events), so host languages without a CUDA binding
 * can time the step the way mpp_device_time_model_step does (shared/mpp/mpp.f90:406-423). */
int swcu_timer_start(swcu_ctx *ctx);
int swcu_timer_stop(swcu_ctx *ctx, float *elapsed_ms);
/* Runs nsteps like swcu_step but brackets every kernel launch with CUDA events on the context's
 * stream and returns the summed device time (ms) and launch count of the two fused kernels
 * (prep, update).  FUSED mode only; used for per-kernel roofline numbers. */
int swcu_profile_steps(swcu_ctx *ctx, double tau, int nsteps,
                       float *prep_ms, long *prep_launches, float *update_ms, long *update_launches);
/* Self-test of the exact division used by the fused kernels: on `n` pseudo-random operand pairs
 * (dividends of both signs over 2^-200..2^200 incl. zeros of both signs, divisors = real(4) values
 * promoted to double like the metric arrays, and arbitrary doubles) compares mdiv(a, b, RN(1/b)) with
 * the IEEE division a/b BITWISE on the device and returns the number of mismatches. */
int swcu_selftest_mdiv(long n, unsigned long long seed, long *mismatches);
/* Launch count of this library's kernels on this context since creation. */
long swcu_launch_count(const swcu_ctx *ctx);
/* Bytes of device memory held by the context. */
long swcu_device_bytes(const swcu_ctx *ctx);
/* The context's compute stream as a cudaStream_t (void*). */
void *swcu_stream(swcu_ctx *ctx);

/* Multi-GPU: one block per GPU, y-slab decomposition (parallel.par: bppnx = 1, bppny = nranks;
 * core/decomposition.f90:468-486).  Replaces hybrid_sync / syncborder_data2D_real8
 * (shared/mpp/sync.f90:294-374, syncborder_block2D_gen_all.fi) with ncclSend/ncclRecv of halo
 * rows over NVLink on a side stream, overlapped with interior compute.
 * swcu_comm_unique_id fills a 128-byte ncclUniqueId on one rank; the caller broadcasts it
 * (MPI_Bcast in the Fortran host, torch.distributed in the Python host). */
int swcu_comm_unique_id(void *id128);
int swcu_comm_init(swcu_ctx *ctx, int nranks, int rank, const void *id128);
int swcu_comm_destroy(swcu_ctx *ctx);
/* The row bookkeeping of the halo exchange, exposed so hosts can test it without a GPU: for the
 * neighbour below (side = 0, rank-1) or above (side = 1, rank+1) and a halo of `nrows` rows, the
 * 0-based array row where the rows to SEND start and where the RECEIVED rows land
 * (array row of reference row n is n - bnd_y1; get_boundary_points_of_block /
 * get_halo_points_of_block, core/decomposition.f90:94-154, 230-290, widened to nrows). */
int swcu_halo_plan(const swcu_dims *dims, int nrows, int side, int *send_row, int *recv_row);
/* The row bookkeeping of k_march, exposed so it can be tested without a GPU: rows [*first .. *last] (inclusive;
 * empty if *last < *first) of band `band` when rows [n0 .. n1] are cut into `nbands` bands of which the last
 * `late_bands` are `late_cut` rows shorter than the others (the bands whose CTAs start late beside a concurrent
 * strip launch).  The bands tile [n0 .. n1] exactly once, in order. */
int swcu_march_band_rows(int n0, int n1, int nbands, int late_bands, int late_cut, int band, int *first, int *last);
/* One explicit halo exchange of a field (all ranks call it): the analogue of
 * `call sync(domain, data2d)` (shared/mpp/sync.f90:541-556) for init-time use.  Over a communicator or
 * in-process links; SWCU_ERR_STATE on a block that uses peer memory (which carries only the step's arrays). */
int swcu_halo_exchange(swcu_ctx *ctx, int field);
/* PRECONDITION of SWCU_MODE_FUSED, and the call that establishes it.  The fused step evaluates the depth,
 * vorticity and stress fields of the cells one layer OUTSIDE a block's interior itself (instead of
 * exchanging them after every kernel like the reference does), so it reads the inputs -- masks, the nine
 * metric / Coriolis arrays, hhq_rest, mu, r_diss, RHSx/RHSy and the prognostic arrays -- TWO layers outside
 * the interior.  The reference fills its block arrays only one layer out (grid_kernels.f90 loops over
 * nx_start-1 .. nx_end+1; its syncs have width 1, core/decomposition.f90:230-270).  After the initial uploads
 * and after the neighbours are attached (swcu_comm_init or swcu_link; NOT peer memory -- attach that
 * afterwards), every block calls swcu_widen_halos once: it exchanges two halo layers of every resident input
 * (corners included for linked blocks).  Inputs built by swcu_init_grid / uploaded from globally generated
 * arrays already satisfy the precondition, and the call is then a harmless repetition.  At the GLOBAL
 * boundary the two outer layers must be land (the reference's masks guarantee a land frame of width 2,
 * tools/io.f90:49-59), where the values never matter.  No-op in SWCU_MODE_REFERENCE and without neighbours. */
int swcu_widen_halos(swcu_ctx *ctx);

/* The same y-slab exchange WITHOUT NCCL, for ranks on one node: every rank exports IPC handles of its
 * prognostic buffers (swcu_peer_export fills SWCU_PEER_BLOB_BYTES bytes), the host layer hands each blob
 * to the two neighbouring ranks (MPI_Sendrecv / torch.distributed), and swcu_peer_attach maps them
 * (side 0 = the block below, 1 = the block above).  Each step the boundary strips are computed
 * first, one small kernel stores them straight into the neighbours' halo rows over NVLink and
 * publishes a step counter there; the neighbour's stream waits for that counter with a stream memory
 * operation (cuStreamWaitValue64), so there is no collective, no copy kernel on the receiving GPU and
 * no host synchronisation in the step.  FUSED mode; all ranks must step in lockstep (same nsteps).
 * Mutually exclusive with swcu_comm_init and swcu_link. */
#define SWCU_PEER_BLOB_BYTES 2048
int swcu_peer_export(swcu_ctx *ctx, void *blob);
int swcu_peer_attach(swcu_ctx *ctx, int side, const void *blob);
int swcu_peer_detach(swcu_ctx *ctx); /* unmaps both sides (all ranks call it before re-attaching) */

/* Several blocks per process (parallel.par bppnx x bppny > 1 x 1, and the reference's _GPU_MULTI_
 * one-process-many-GPUs mode): swcu_link ties two contexts of the same process that are neighbours
 * in a tensor-product block grid -- side, or corner; the direction follows from their dims, the
 * contexts may live on the same device or on two devices (peer access is enabled when the topology
 * offers it).  Linked blocks step together through swcu_step_group, which replaces the same-rank
 * block-to-block halo copy of shared/mpp/syncborder_block2D_gen_all.fi:218-249: block arrays use
 * global indices, so every block pulls the cells it lacks from the same (m, n) of its neighbour's
 * array with one strided device-to-device copy per direction, ordered by events on the compute
 * streams (no host synchronisation inside a step).  FUSED mode pulls 2 halo layers of the six
 * prognostic arrays once per step; REFERENCE mode pulls 1 layer after every kernel whose
 * envoke_*_sync is non-empty.  A group lists every block its members are linked to; unlinked
 * blocks may ride along.  swcu_step on a linked context is an error (SWCU_ERR_STATE), as is mixing
 * links with a communicator.  swcu_envoke_sync / swcu_halo_exchange on a linked context pull from
 * the neighbours after synchronising their streams (host-blocking; the caller has issued the
 * producing kernel on all blocks first, as the reference's loop over blocks does). */
int swcu_link(swcu_ctx *a, swcu_ctx *b);
int swcu_unlink(swcu_ctx *ctx); /* drops all links of ctx (swcu_destroy does this too) */
int swcu_step_group(swcu_ctx *const *ctxs, int n, double tau, int nsteps);

/* ------------------------------------------------------------------------------------------
 * Host-side input construction (C++, no GPU needed): what init_grid_data / init_ocean_data
 * (control/init_data.f90:29-125) produce for one block, in the reference layout.  These mirror
 * lu_init_kernel / lu_lv_init_kernel / grid_base_init_kernel / grid_geo_init_kernel
 * (kernel/service/grid_kernels.f90:18-204, 206-538) and gaussian_elimination_kernel
 * (kernel/shallow_water/vel_ssh.f90:15-38).
 * ------------------------------------------------------------------------------------------ */
typedef struct swh_basin {
    int nx, ny;                              /* basin.par 1-2 */
    double dxst, dyst, rlon, rlat;           /* basin.par 6-9 */
    int curve_grid;                          /* basin.par 12: 0 carthesian, 1 spherical */
    double rotation_on_lon, rotation_on_lat; /* basin.par 13-14 */
} swh_basin;

/* mask: global nx*ny ints ((m,n) at mask[(n-1)*nx+(m-1)], 0 = sea) or NULL for "none". */
int swh_masks(const swh_basin *b, const swcu_dims *d, const int *mask,
              float *lu, float *luu, float *luh, float *lcu, float *lcv, float *llu, float *llv);
int swh_metrics(const swh_basin *b, const swcu_dims *d,
                float *dx, float *dy, float *dxt, float *dyt, float *dxh, float *dyh,
                float *dxb, float *dyb, float *rlh_s);
int swh_gaussian(const swcu_dims *d, const float *lu, double *ssh, double sigma, int nx0, int ny0);
/* block_uniform_decomposition (core/decomposition.f90:427-503): interior start/size of block
 * `i` of `nb` along an axis of `ncells` computational cells. */
int swh_uniform_split(int ncells, int nb, int i, int *start, int *size);
/* Which rank (or GPU of one process) owns which block of a bnx x bny block grid; arrays are indexed
 * [bn*bnx + bm] with 0-based block coordinates.
 *   swh_block_weights      bglob_weight = sea cells per block (core/decomposition.f90:505-520);
 *                          mask as for swh_masks (NULL = "none")
 *   swh_hilbert_d2xy       shared/mpp/hilbert_curve.f90:13-61
 *   swh_hilbert_partition  create_hilbert_curve_decomposition (core/decomposition.f90:532-612):
 *                          needs bnx = bny = 2^M; consecutive pieces of the Hilbert walk with about
 *                          equal weight; `powers` = compute_powers per rank or NULL
 *   swh_uniform_partition  create_uniform_decomposition (core/decomposition.f90:614-670)
 * owner = -1 marks a land-only block (no rank holds it). */
/*   swh_balanced_slabs     y-slabs (one block per GPU) of about equal work instead of equal height: the
 *                          same sea-weight idea applied to the slab cut.  Work = tiles of
 *                          tile_cols x band_rows cells; an all-land tile (which the fused step skips)
 *                          counts `land_cost` of a tile with sea.  Cuts fall on band boundaries.
 *                          start[r], size[r] (r < nranks) as in swh_uniform_split. */
int swh_balanced_slabs(int nx, int ny, const int *mask, int nranks, int band_rows, int tile_cols, double land_cost,
                       int *start, int *size);
int swh_block_weights(int nx, int ny, int bnx, int bny, const int *mask, double *weights);
int swh_hilbert_d2xy(int order, int d, int *x, int *y);
int swh_hilbert_partition(int nb, const double *weights, int nranks, const double *powers, int *owner);
int swh_uniform_partition(int bnx, int bny, int px, int py, const double *weights, int *owner);

/* ------------------------------------------------------------------------------------------
 * Device-side construction of a resident context's static inputs: init_grid_data
 * (control/init_data.f90:96-125) without building and uploading sixteen 2-D arrays.
 * swcu_init_grid fills the seven masks and the nine metric / Coriolis arrays of the context:
 *   - lu from `mask` (the GLOBAL nx*ny integer mask in the layout of swh_masks, only the block's
 *     window is copied to the device; NULL = "none", the rectangular basin of tools/io.f90:49-59) and
 *     the derived masks with the reference's tests (kernel/service/grid_kernels.f90:18-92), on the device;
 *   - dx .. dyb, rlh_s: where they are constant along x (carthesian, or spherical with
 *     rotation_on_lat = 0) one column is evaluated on the host by swh_metrics -- same libm calls,
 *     same real(4) roundings -- and spread over the rows on the device; with a rotated pole the arrays
 *     are built on the host and uploaded, as swh_metrics + swcu_upload would.
 * Results equal swh_masks / swh_metrics + swcu_upload bit for bit.
 * swcu_fill / swcu_copy_field are data2D%fill and %copy_from on the resident arrays (whole block
 * array incl. frame, core/data_types.f90:665-716): hhq_rest = 100, mu = 0, sshp = ssh ...
 * (control/init_data.f90:57-58,76-77,112-114).  Fields FUSED mode does not keep are accepted and
 * ignored, like swcu_upload does. */
int swcu_init_grid(swcu_ctx *ctx, const swh_basin *basin, const int *mask);
int swcu_fill(swcu_ctx *ctx, int field, double value);
int swcu_copy_field(swcu_ctx *ctx, int dst_field, int src_field);

#ifdef __cplusplus
}
#endif
#endif
