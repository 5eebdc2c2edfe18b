/*
 * sw_driver.c -- `program model` (model.f90:47-200) restricted to the shallow-water path, in plain C
 * on top of include/swcuda.h: what a host that is not Python and not Fortran links against.
 *
 *   cc -std=c99 -Iinclude examples/sw_driver.c -Locean_model_arch_b200 -lswcuda \
 *      -Wl,-rpath,$PWD/ocean_model_arch_b200 -o sw_driver
 *   ./sw_driver NX NY NSTEPS [mode: fused|reference] [bnx bny]
 *
 * Builds the inputs of init_grid_data / init_ocean_data with the library's host functions (swh_*),
 * keeps one resident context per block (bnx x bny blocks, all on device 0, linked as neighbours),
 * runs NSTEPS of tau = 1 s, and prints for ssh, ubrtr, vbrtr the FNV-1a hash of the interior cells
 * in global row-major order plus max|ssh| -- numbers tests/test_gpu_driver.py recomputes from the
 * oracle.  Exit status is non-zero on any library error (text from swcu_last_error()).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "swcuda.h"

#define CHECK(call)                                                                       \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != SWCU_OK) {                                                             \
            fprintf(stderr, "%s failed with %d: %s\n", #call, rc_, swcu_last_error());   \
            exit(1);                                                                      \
        }                                                                                 \
    } while (0)

typedef struct block {
    swcu_dims d;
    swcu_ctx *ctx;
    int w, h;
} block;

static uint64_t fnv1a(uint64_t hsh, const void *p, size_t n)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; ++i) { hsh ^= b[i]; hsh *= 1099511628211ULL; }
    return hsh;
}

static void *zalloc(size_t n) { void *p = calloc(n, 1); if (!p) { fprintf(stderr, "out of memory\n"); exit(1); } return p; }

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s NX NY NSTEPS [fused|reference] [bnx bny]\n", argv[0]); return 2; }
    const int nx = atoi(argv[1]), ny = atoi(argv[2]), nsteps = atoi(argv[3]);
    const int mode = (argc > 4 && !strcmp(argv[4], "reference")) ? SWCU_MODE_REFERENCE : SWCU_MODE_FUSED;
    const int bnx = argc > 6 ? atoi(argv[5]) : 1, bny = argc > 6 ? atoi(argv[6]) : 1;
    const int nb = bnx * bny;
    if (swcu_device_count() < 1) { fprintf(stderr, "no CUDA device (the library has no CPU path)\n"); return 3; }

    /* the shipped basin.par / sw.par values (configs/basinpar.f90:64-83, configs/sw.f90:34-41) */
    const swh_basin basin = {nx, ny, 0.00312, 0.00225, 34.75156, 44.801125, 1, 0.0, 0.0};
    const swcu_params par = {1, 1, 1, 0.5, 0, mode};
    const double tau = 1.0;

    block *blk = (block *)zalloc(sizeof(block) * (size_t)nb);
    swcu_ctx **group = (swcu_ctx **)zalloc(sizeof(swcu_ctx *) * (size_t)nb);
    for (int bn = 0; bn < bny; ++bn)
        for (int bm = 0; bm < bnx; ++bm) {
            block *b = &blk[bn * bnx + bm];
            int xs, xn, ys, yn;   /* block_uniform_decomposition, core/decomposition.f90:427-503 */
            CHECK(swh_uniform_split(nx - 4, bnx, bm, &xs, &xn));
            CHECK(swh_uniform_split(ny - 4, bny, bn, &ys, &yn));
            const swcu_dims d = {3 + xs, 3 + xs + xn - 1, 3 + ys, 3 + ys + yn - 1,
                                 3 + xs - 2, 3 + xs + xn + 1, 3 + ys - 2, 3 + ys + yn + 1};
            b->d = d; b->w = d.bnd_x2 - d.bnd_x1 + 1; b->h = d.bnd_y2 - d.bnd_y1 + 1;
            const size_t n = (size_t)b->w * b->h;
            float *f4[16];
            for (int i = 0; i < 16; ++i) f4[i] = (float *)zalloc(n * sizeof(float));
            CHECK(swh_masks(&basin, &d, NULL, f4[0], f4[1], f4[2], f4[3], f4[4], f4[5], f4[6]));
            CHECK(swh_metrics(&basin, &d, f4[7], f4[8], f4[9], f4[10], f4[11], f4[12], f4[13], f4[14], f4[15]));
            double *ssh = (double *)zalloc(n * sizeof(double)), *hrest = (double *)zalloc(n * sizeof(double));
            for (size_t i = 0; i < n; ++i) hrest[i] = 100.0;   /* control/init_data.f90:112-114 */
            /* Gaussian bump on every sea cell of the global interior this block's array covers */
            swcu_dims wide = d;
            wide.nx_start = d.bnd_x1 > 3 ? d.bnd_x1 : 3; wide.nx_end = d.bnd_x2 < nx - 2 ? d.bnd_x2 : nx - 2;
            wide.ny_start = d.bnd_y1 > 3 ? d.bnd_y1 : 3; wide.ny_end = d.bnd_y2 < ny - 2 ? d.bnd_y2 : ny - 2;
            CHECK(swh_gaussian(&wide, f4[0], ssh, 1.0, nx / 2, ny / 2));

            CHECK(swcu_create(&b->ctx, &d, &par, 0));
            const int ids4[16] = {SWCU_F_LU, SWCU_F_LUU, SWCU_F_LUH, SWCU_F_LCU, SWCU_F_LCV, SWCU_F_LLU, SWCU_F_LLV,
                                  SWCU_F_DX, SWCU_F_DY, SWCU_F_DXT, SWCU_F_DYT, SWCU_F_DXH, SWCU_F_DYH, SWCU_F_DXB,
                                  SWCU_F_DYB, SWCU_F_RLH_S};
            for (int i = 0; i < 16; ++i) CHECK(swcu_upload(b->ctx, ids4[i], f4[i]));
            CHECK(swcu_upload(b->ctx, SWCU_F_HHQ_REST, hrest));
            CHECK(swcu_upload(b->ctx, SWCU_F_SSH, ssh));
            CHECK(swcu_upload(b->ctx, SWCU_F_SSHP, ssh));
            CHECK(swcu_upload(b->ctx, SWCU_F_SSHN, ssh));
            CHECK(swcu_envoke_hh_init(b->ctx));          /* control/init_data.f90:60-63 */
            CHECK(swcu_synchronize(b->ctx, NULL));
            for (int i = 0; i < 16; ++i) free(f4[i]);
            free(ssh); free(hrest);
            group[bn * bnx + bm] = b->ctx;
        }
    /* neighbours: E, N, NE, NW of every block, each pair once */
    const int dm[4] = {1, 0, 1, -1}, dn[4] = {0, 1, 1, 1};
    for (int bn = 0; bn < bny; ++bn)
        for (int bm = 0; bm < bnx; ++bm)
            for (int k = 0; k < 4; ++k) {
                const int om = bm + dm[k], on = bn + dn[k];
                if (om < 0 || om >= bnx || on >= bny) continue;
                CHECK(swcu_link(blk[bn * bnx + bm].ctx, blk[on * bnx + om].ctx));
            }

    if (nb == 1) CHECK(swcu_step(group[0], tau, nsteps));
    else CHECK(swcu_step_group(group, nb, tau, nsteps));
    for (int k = 0; k < nb; ++k) {
        long bad = 0;
        CHECK(swcu_synchronize(group[k], &bad));   /* SWCU_ERR_BLOWUP if K11 fired */
    }

    /* assemble the global interior and hash it */
    const int fields[3] = {SWCU_F_SSH, SWCU_F_UBRTR, SWCU_F_VBRTR};
    const char *names[3] = {"ssh", "ubrtr", "vbrtr"};
    const size_t gw = (size_t)(nx - 4), gh = (size_t)(ny - 4);
    double *global = (double *)zalloc(gw * gh * sizeof(double));
    for (int f = 0; f < 3; ++f) {
        for (int k = 0; k < nb; ++k) {
            const block *b = &blk[k];
            double *a = (double *)zalloc((size_t)b->w * b->h * sizeof(double));
            CHECK(swcu_download(b->ctx, fields[f], a));
            for (int n = b->d.ny_start; n <= b->d.ny_end; ++n)
                for (int m = b->d.nx_start; m <= b->d.nx_end; ++m)
                    global[(size_t)(n - 3) * gw + (size_t)(m - 3)] = a[(size_t)(n - b->d.bnd_y1) * b->w + (m - b->d.bnd_x1)];
            free(a);
        }
        double amax = 0.0;
        for (size_t i = 0; i < gw * gh; ++i) if (fabs(global[i]) > amax) amax = fabs(global[i]);
        printf("%s fnv1a=%016llx max_abs=%.17g\n", names[f],
               (unsigned long long)fnv1a(14695981039346656037ULL, global, gw * gh * sizeof(double)), amax);
    }
    long launches = 0;
    for (int k = 0; k < nb; ++k) launches += swcu_launch_count(group[k]);
    printf("blocks=%d steps=%d launches=%ld mode=%s\n", nb, nsteps, launches, mode == SWCU_MODE_FUSED ? "fused" : "reference");
    for (int k = 0; k < nb; ++k) CHECK(swcu_destroy(group[k]));
    free(global); free(group); free(blk);
    return 0;
}
