/*
 * sw_oracle.h -- CPU ORACLE for the shallow-water step of Andrcraft9/ocean_model_arch.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (ocean_model_arch_b200/) may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / reported CPU baseline.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or fixtures for
 * this path, and it cannot be compiled here (no Fortran compiler, no MPI).
 * This file is a plain-C restatement of the reference's Fortran, honouring Fortran's mixed
 * real(4)/real(8) promotion rules, evaluated strictly (no FMA contraction, no reassociation:
 * build with -ffp-contract=off -fno-fast-math).  Every function cites the reference file:line
 * it follows (paths relative to /root/reference).
 * What stands in for the missing pins (tests/test_oracle.py): committed digests, an independent
 * NumPy restatement written from the Fortran (tests/np_restatement.py) that must agree bitwise on
 * every array, analytic known answers for every term of the scheme, and the invariants of
 * SURVEY.md 8c (decomposition invariance, conservation, untouched land).
 */
#ifndef SW_ORACLE_H
#define SW_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* The eight leading integer arguments of every reference kernel
 * (kernel/shallow_water/vel_ssh.f90:69-70 et al.). */
#define SWO_DIMS int nx_start, int nx_end, int ny_start, int ny_end, \
                 int bnd_x1, int bnd_x2, int bnd_y1, int bnd_y2

/* ---- per-kernel restatements (stateless, reference argument order) ---- */
void swo_gaussian_elimination_kernel(SWO_DIMS, const float *lu, double *ssh, double sigma, int nx0, int ny0);
int  swo_check_ssh_err_kernel(SWO_DIMS, const float *lu, const double *ssh);
void swo_sw_update_ssh_kernel(SWO_DIMS, double tau, const float *lu, const float *dx, const float *dy,
                              const float *dxh, const float *dyh, const double *hhu, const double *hhv,
                              double *sshn, const double *sshp, const double *ubrtr, const double *vbrtr);
void swo_sw_update_uv(SWO_DIMS, double tau, const float *lcu, const float *lcv,
                      const float *dxt, const float *dyt, const float *dxh, const float *dyh,
                      const float *dxb, const float *dyb,
                      const double *hhu, const double *hhun, const double *hhup,
                      const double *hhv, const double *hhvn, const double *hhvp,
                      const double *hhh, const double *ssh,
                      const double *ubrtr, double *ubrtrn, const double *ubrtrp,
                      const double *vbrtr, double *vbrtrn, const double *vbrtrp,
                      const float *rdis, const float *rlh_s,
                      const double *RHSx, const double *RHSy, const double *RHSx_adv, const double *RHSy_adv,
                      const double *RHSx_dif, const double *RHSy_dif);
void swo_sw_next_step(SWO_DIMS, double time_smooth, const float *lu, const float *lcu, const float *lcv,
                      double *ssh, double *sshn, double *sshp,
                      double *ubrtr, double *ubrtrn, double *ubrtrp,
                      double *vbrtr, double *vbrtrn, double *vbrtrp);
void swo_uv_trans_vort_kernel(SWO_DIMS, const float *luu, const float *dxt, const float *dyt,
                              const float *dxb, const float *dyb, const double *u, const double *v, double *vort);
void swo_uv_trans_kernel(SWO_DIMS, const float *lcu, const float *lcv, const float *luu,
                         const float *dxh, const float *dyh, const double *u, const double *v, const double *vort,
                         const double *hq, const double *hu, const double *hv, const double *hh,
                         double *RHSx, double *RHSy);
void swo_uv_diff2_kernel(SWO_DIMS, const float *lcu, const float *lcv,
                         const float *dx, const float *dy, const float *dxt, const float *dyt,
                         const float *dxh, const float *dyh, const float *dxb, const float *dyb,
                         const double *mu, const double *str_t, const double *str_s,
                         const double *hq, const double *hu, const double *hv, const double *hh,
                         double *RHSx, double *RHSy);
void swo_stress_components_kernel(SWO_DIMS, const float *lu, const float *luu,
                                  const float *dx, const float *dy, const float *dxt, const float *dyt,
                                  const float *dxh, const float *dyh, const float *dxb, const float *dyb,
                                  const double *u, const double *v, double *str_t, double *str_s);
void swo_hh_init_kernel(SWO_DIMS, int full_free_surface,
                        const float *lu, const float *llu, const float *llv, const float *luh,
                        const float *dx, const float *dy, const float *dxt, const float *dyt,
                        const float *dxh, const float *dyh, const float *dxb, const float *dyb,
                        double *hq, double *hqp, double *hqn, double *hu, double *hup, double *hun,
                        double *hv, double *hvp, double *hvn, double *hh, double *hhp, double *hhn,
                        const double *sh, const double *shp, const double *h_r);
void swo_hh_update_kernel(SWO_DIMS, const float *lu, const float *llu, const float *llv, const float *luh,
                          const float *dx, const float *dy, const float *dxt, const float *dyt,
                          const float *dxh, const float *dyh, const float *dxb, const float *dyb,
                          double *hqn, double *hun, double *hvn, double *hhn, const double *sh, const double *h_r);
void swo_hh_shift_kernel(SWO_DIMS, double time_smooth,
                         const float *lu, const float *llu, const float *llv, const float *luh,
                         double *hq, double *hqp, double *hqn, double *hu, double *hup, double *hun,
                         double *hv, double *hvp, double *hvn, double *hh, double *hhp, double *hhn);
void swo_tran_diff_fluxes_kernel(SWO_DIMS, const float *lcu, const float *lcv,
                                 const float *dxt, const float *dyt, const float *dxh, const float *dyh,
                                 const double *hhu, const double *hhv, const double *ff, const double *ffp,
                                 const double *uu, const double *vv, const double *mu, double factor_mu,
                                 double *flux_x, double *flux_y);
void swo_tran_diff_tracer_kernel(SWO_DIMS, const float *lu, const float *dx, const float *dy, double tau,
                                 const double *hhqn, const double *hhqp, const double *flux_x, const double *flux_y,
                                 const double *ffp, double *ffn);
void swo_tracer_next_step_kernel(SWO_DIMS, double time_smooth, const float *lu,
                                 const double *ffn, double *ffp, double *ff);

/* grid construction kernels (kernel/service/grid_kernels.f90) */
void swo_lu_init_kernel(int bnd_x1, int bnd_x2, int bnd_y1, int bnd_y2, int nx, int ny,
                        const int *mask, float *lu, float *lu1);
void swo_lu_lv_init_kernel(int bnd_x1, int bnd_x2, int bnd_y1, int bnd_y2, const float *lu,
                           float *luh, float *luu, float *llu, float *llv, float *lcu, float *lcv);

/* ---- whole-model restatement (model.f90 init + time loop, block mode) ---- */
typedef struct swo_config {
    int nx, ny;                 /* basin.par 1-2 (configs/basinpar.f90:64-65) */
    double dxst, dyst, rlon, rlat;          /* basin.par 6-9 */
    int curve_grid;             /* 0 carthesian, 1 spherical (basin.par 12) */
    double rotation_on_lon, rotation_on_lat;/* basin.par 13-14 */
    int full_free_surface, trans_terms, ksw_lat;  /* sw.par 1-3 */
    double time_smooth, lvisc_2;            /* sw.par 4-5 */
    int use_tracers, tracer_num;            /* sw.par 6-7 */
    float time_step;            /* ocean_run.par 2, real(4) (tools/time_manager.f90:52) */
    int bnx, bny;               /* total block grid (parallel.par 3-4 with _DD_MANUAL_BLOCK_GRID_) */
    /* Deviations from the shipped setup, off by default (BASELINE config 4): */
    int keep_mu;                /* 1: do NOT zero mu after filling it with lvisc_2 (quirk: init_data.f90:76-77) */
    float r_diss;               /* constant Rayleigh friction written to r_diss (reference leaves it 0) */
    double hhq_rest;            /* flat bottom depth, reference uses 100.0 (init_data.f90:112-114) */
    int nthreads;               /* OpenMP threads over blocks; 0 = leave runtime default */
} swo_config;

typedef struct swo_model swo_model;

/* mask: nx*ny ints, element (m,n) (1-based) at mask[(n-1)*nx+(m-1)], 0 = sea; NULL = "none"
 * (closed rectangle with a 2-cell land frame, tools/io.f90:49-59). */
swo_model *swo_create(const swo_config *cfg, const int *mask);
void swo_destroy(swo_model *m);
/* nsteps x (expl_shallow_water ; expl_tracer).  Returns number of check_ssh_err failures. */
long swo_step(swo_model *m, int nsteps);
/* Gather field `name` of every block's interior+frame into a global nx*ny array
 * (out8 for real(8) fields, out4 for real(4) ones; pass the other as NULL).  Returns 0 on
 * success, 1 if the name is unknown, 2 if the wrong pointer kind was given.
 * Cells are taken from the block that owns them (interior), the outer 2-cell frame from the
 * adjacent border blocks. */
int swo_get_field(const swo_model *m, const char *name, double *out8, float *out4);
/* Overwrite a field of every block (including its halo frame) from a global array. */
int swo_set_field(swo_model *m, const char *name, const double *in8, const float *in4);
/* Raw access to block k's copy (bnd_x1:bnd_x2, bnd_y1:bnd_y2) for decomposition tests. */
int swo_block_count(const swo_model *m);
int swo_block_dims(const swo_model *m, int k, int dims8[8]);
void *swo_block_field(swo_model *m, int k, const char *name, int *is_real8);

#ifdef __cplusplus
}
#endif
#endif
