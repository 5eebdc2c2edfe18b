// sw_host.cpp -- host-side (CPU, one-time) construction of the hot path's inputs for one block:
// sea/land masks, metric arrays, Coriolis parameter and the Gaussian initial state, in the
// reference's block-array layout.  This is the C++ stand-in for the part of the reference's
// Fortran host (init_grid_data / init_ocean_data, control/init_data.f90:29-125) that feeds the
// kernel layer; it is off the timed path.  libm's cos/sin/exp are used exactly where the
// reference uses dcos/dsin/dexp so a gfortran build on the same libm produces the same bits.
#include <cmath>
#include <cstddef>
#include <vector>

#include "../../include/swcuda.h"

namespace {

struct Lay {
    int bx1, by1, ld;
    size_t operator()(int m, int n) const { return (size_t)(n - by1) * ld + (size_t)(m - bx1); }
};
inline Lay lay(const swcu_dims &d) { return Lay{d.bnd_x1, d.bnd_y1, d.bnd_x2 - d.bnd_x1 + 1}; }

// shared/constants.f90:11-23 (real(4) parameters; dPi is the truncated literal)
const float kPi4 = 3.1415926f;
const float kRadEarth = 6371000.0f;
const float kEarthAngVel = 7.2921159e-5f;
const double kDPi = 3.14159265358979;
const double kLatExtr = 89.99999;

// core/math_tools.f90:28-40
double dcosd(double x) { return std::cos((x / 180.0) * kDPi); }
double dsind(double x) { return std::sin((x / 180.0) * kDPi); }

inline int imax(int a, int b) { return a > b ? a : b; }
inline int imin(int a, int b) { return a < b ? a : b; }

}  // namespace

extern "C" {

// lu_init_kernel + lu_lv_init_kernel (kernel/service/grid_kernels.f90:18-92).  lu is taken from
// the global integer mask on the whole block array; the six derived masks are evaluated on
// [bnd1 .. bnd2-1] like the reference does, the last row/column stays 0.
int swh_masks(const swh_basin *b, const swcu_dims *d, const int *mask,
              float *lu, float *luu, float *luh, float *lcu, float *lcv, float *llu, float *llv)
{
    if (!b || !d || !lu) return SWCU_ERR_ARG;
    const Lay L = lay(*d);
    const int nx = b->nx, ny = b->ny;
    const size_t total = (size_t)L.ld * (size_t)(d->bnd_y2 - d->bnd_y1 + 1);
    float *out[7] = {lu, luu, luh, lcu, lcv, llu, llv};
    for (float *p : out)
        if (p) for (size_t i = 0; i < total; ++i) p[i] = 0.0f;
    for (int n = d->bnd_y1; n <= d->bnd_y2; ++n)
        for (int m = d->bnd_x1; m <= d->bnd_x2; ++m) {
            if (m < 1 || m > nx || n < 1 || n > ny) return SWCU_ERR_ARG;
            int land;
            if (mask) land = mask[(size_t)(n - 1) * nx + (m - 1)];
            else land = (m < 3 || m > nx - 2 || n < 3 || n > ny - 2) ? 1 : 0;  // tools/io.f90:49-59
            if (land == 0) lu[L(m, n)] = 1.0f;
        }
    for (int n = d->bnd_y1; n <= d->bnd_y2 - 1; ++n)
        for (int m = d->bnd_x1; m <= d->bnd_x2 - 1; ++m) {
            const float a = lu[L(m, n)], e = lu[L(m + 1, n)], no = lu[L(m, n + 1)], en = lu[L(m + 1, n + 1)];
            const size_t c = L(m, n);
            if (luh && a + e + no + en > 0.5f) luh[c] = 1.0f;
            if (luu && a * e * no * en > 0.5f) luu[c] = 1.0f;
            if (llu && a + e > 0.5f) llu[c] = 1.0f;
            if (llv && a + no > 0.5f) llv[c] = 1.0f;
            if (lcu && a * e > 0.5f) lcu[c] = 1.0f;
            if (lcv && a * no > 0.5f) lcv[c] = 1.0f;
        }
    return SWCU_OK;
}

// grid_base_init_kernel + grid_geo_init_kernel for a uniform grid (xgr_type = ygr_type = 0):
// kernel/service/grid_kernels.f90:164-201, 246-421 with grid_parameters_carthesian / _spherical
// (kernel/service/grid_parameters.f90:16-181).  Filled where the reference's global significant
// area [2..nx-1]x[2..ny-1] meets the block array (for a single block that is exactly the
// reference's nx_start-1 .. nx_end+1 range); everything else stays 0 like the reference's
// zero-initialised allocation.  rlh_s covers the whole array (grid_kernels.f90:201).
int swh_metrics(const swh_basin *b, const swcu_dims *d,
                float *dx, float *dy, float *dxt, float *dyt, float *dxh, float *dyh,
                float *dxb, float *dyb, float *rlh_s)
{
    if (!b || !d) return SWCU_ERR_ARG;
    if (b->curve_grid != 0 && b->curve_grid != 1) return SWCU_ERR_ARG;
    const Lay L = lay(*d);
    const size_t total = (size_t)L.ld * (size_t)(d->bnd_y2 - d->bnd_y1 + 1);
    float *all[9] = {dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, rlh_s};
    for (float *p : all) {
        if (!p) return SWCU_ERR_ARG;
        for (size_t i = 0; i < total; ++i) p[i] = 0.0f;
    }
    const float pip180 = kPi4 / 180.0f;
    const float sx = (float)b->dxst * pip180 * kRadEarth;  // sngl(dxst)*pip180*RadEarth
    const float sy = (float)b->dyst * pip180 * kRadEarth;
    for (size_t i = 0; i < total; ++i) rlh_s[i] = 2.0f * kEarthAngVel;

    const int m1 = imax(d->bnd_x1, 2), m2 = imin(d->bnd_x2, b->nx - 1);
    const int n1 = imax(d->bnd_y1, 2), n2 = imin(d->bnd_y2, b->ny - 1);
    // model coordinates, grid_kernels.f90:118-148 (mmm = nnn = 3)
    auto xt = [&](int m) { return b->rlon + (double)(m - 3) * b->dxst; };
    auto yt = [&](int n) { return b->rlat + (double)(n - 3) * b->dyst; };
    auto xu = [&](int m) { return (xt(m) + xt(m + 1)) / 2.0; };
    auto yv = [&](int n) { return (yt(n) + yt(n + 1)) / 2.0; };

    const double sinlat_extr = dsind(kLatExtr);
    const float sq2 = std::sqrt(2.0f);
    for (int n = n1; n <= n2; ++n) {
        float cos_t = 1.0f, cos_v = 1.0f;
        if (b->curve_grid == 1) {
            const double lat_t = std::fmax(std::fmin(yt(n), kLatExtr), -kLatExtr);
            const double lat_v = std::fmax(std::fmin(yv(n), kLatExtr), -kLatExtr);
            cos_t = (float)dcosd(lat_t);
            cos_v = (float)dcosd(lat_v);
        }
        for (int m = m1; m <= m2; ++m) {
            const size_t c = L(m, n);
            dy[c] = sy; dyt[c] = sy; dyh[c] = sy; dyb[c] = sy;  // metr_y * 1.0
            if (b->curve_grid == 1) {
                dx[c] = sx * cos_t;   // T grid: (xt, yt)
                dxt[c] = sx * cos_t;  // U grid: (xu, yt)
                dxh[c] = sx * cos_v;  // V grid: (xt, yv)
                dxb[c] = sx * cos_v;  // H grid: (xu, yv)
                // H-grid call carries key_cor = 1: cor_sin = cor_sin * sngl(sin_lat)
                double sin_lat = dsind(yv(n)) * dcosd(b->rotation_on_lat)
                               + dcosd(xu(m)) * dcosd(yv(n)) * dsind(b->rotation_on_lat);
                sin_lat = std::fmin(std::fmax(sin_lat, -sinlat_extr), sinlat_extr);
                rlh_s[c] = rlh_s[c] * (float)sin_lat;
            } else {
                dx[c] = sx; dxt[c] = sx; dxh[c] = sx; dxb[c] = sx;
                rlh_s[c] = rlh_s[c] / sq2;  // grid_parameters.f90:71
            }
        }
    }
    return SWCU_OK;
}

// gaussian_elimination_kernel (kernel/shallow_water/vel_ssh.f90:15-38) over the block's
// nx_start..nx_end x ny_start..ny_end of `d` (callers widen d's interior to cover halo rows that
// a neighbouring block owns, which is what the reference's following sync delivers).
int swh_gaussian(const swcu_dims *d, const float *lu, double *ssh, double sigma, int nx0, int ny0)
{
    if (!d || !lu || !ssh) return SWCU_ERR_ARG;
    const Lay L = lay(*d);
    for (int n = d->ny_start; n <= d->ny_end; ++n)
        for (int m = d->nx_start; m <= d->nx_end; ++m) {
            const size_t c = L(m, n);
            if (lu[c] > 0.5f) {
                const double ddx = (double)(m - nx0) / (nx0 * 0.25);
                const double ddy = (double)(n - ny0) / (ny0 * 0.25);
                ssh[c] = (1.0 / (std::sqrt(2 * kDPi) * sigma)) * std::exp(-((ddx * ddx + ddy * ddy) / (2 * sigma * sigma)));
            }
        }
    return SWCU_OK;
}

// core/decomposition.f90:441-458: block sizes floor((N - done)/(blocks left)), last takes the rest
int swh_uniform_split(int ncells, int nb, int i, int *start, int *size)
{
    if (nb < 1 || i < 0 || i >= nb || !start || !size) return SWCU_ERR_ARG;
    int done = 0;
    for (int k = 0; k <= i; ++k) {
        const int s = (k == nb - 1) ? ncells - done : (int)std::floor((float)(ncells - done) / (float)(nb - k));
        if (s <= 0) return SWCU_ERR_ARG;
        if (k == i) { *start = done; *size = s; }
        done += s;
    }
    return SWCU_OK;
}

// bglob_weight of block_uniform_decomposition (core/decomposition.f90:505-520, without
// _DD_BINARY_BLOCK_WEIGHTS_): the number of sea cells in each block's interior.
int swh_block_weights(int nx, int ny, int bnx, int bny, const int *mask, double *weights)
{
    if (!weights || nx < 5 || ny < 5 || bnx < 1 || bny < 1) return SWCU_ERR_ARG;
    for (int bn = 0; bn < bny; ++bn) {
        int ys = 0, yn = 0;
        if (int rc = swh_uniform_split(ny - 4, bny, bn, &ys, &yn)) return rc;
        for (int bm = 0; bm < bnx; ++bm) {
            int xs = 0, xn = 0;
            if (int rc = swh_uniform_split(nx - 4, bnx, bm, &xs, &xn)) return rc;
            double wgt = 0.0;
            if (!mask) wgt = (double)xn * (double)yn;  // "none": every interior cell is sea
            else
                for (int n = 3 + ys; n < 3 + ys + yn; ++n)
                    for (int m = 3 + xs; m < 3 + xs + xn; ++m)
                        wgt += mask[(size_t)(n - 1) * nx + (m - 1)] == 0 ? 1.0 : 0.0;
            weights[(size_t)bn * bnx + bm] = wgt;
        }
    }
    return SWCU_OK;
}

// Cell number d of the Hilbert curve of order `order` (shared/mpp/hilbert_curve.f90:13-61, the
// classic quadrant walk: at each scale pick the quadrant from two bits of d, undo its rotation).
int swh_hilbert_d2xy(int order, int d, int *x, int *y)
{
    if (order < 0 || order > 15 || d < 0 || !x || !y) return SWCU_ERR_ARG;
    int px = 0, py = 0;
    for (int side = 1, rest = d; side < (1 << order); side *= 2, rest /= 4) {
        const int qx = (rest >> 1) & 1, qy = (rest ^ qx) & 1;
        if (qy == 0) {                 // quadrants 0 and 3 are transposed, 3 is also mirrored
            if (qx == 1) { px = side - 1 - px; py = side - 1 - py; }
            const int t = px; px = py; py = t;
        }
        px += side * qx; py += side * qy;
    }
    *x = px; *y = py;
    return SWCU_OK;
}

// create_hilbert_curve_decomposition (core/decomposition.f90:532-612): walk the nb x nb blocks along
// the Hilbert curve and cut the walk into `nranks` consecutive pieces of about equal weight
// (piece i aims at the remaining weight times power_i / sum of the remaining powers; a block goes to
// the next piece when the piece with and without it straddles that mean).  owner[bn*nb + bm] = rank,
// -1 for blocks of zero weight.  `powers` may be NULL (all ranks equal).
int swh_hilbert_partition(int nb, const double *weights, int nranks, const double *powers, int *owner)
{
    if (!weights || !owner || nb < 1 || nranks < 1) return SWCU_ERR_ARG;
    int order = 0;
    while ((1 << order) < nb) ++order;
    if ((1 << order) != nb) return SWCU_ERR_ARG;   // "Can`t build Hilbert curve for this geometry"
    auto power_from = [&](int i) { double s = 0.0; for (int k = i; k < nranks; ++k) s += powers ? powers[k] : 1.0; return s; };
    auto power = [&](int i) { return powers ? powers[i] : 1.0; };
    double total = 0.0;
    for (int k = 0; k < nb * nb; ++k) total += weights[k];
    double mean = total * power(0) / power_from(0), piece = 0.0, assigned = 0.0;
    int rank = 0;
    for (int d = 0; d < nb * nb; ++d) {
        int bm = 0, bn = 0;
        swh_hilbert_d2xy(order, d, &bm, &bn);
        const size_t k = (size_t)bn * nb + bm;
        const double wgt = weights[k];
        if (wgt == 0.0) { owner[k] = -1; continue; }
        piece += wgt;
        if (piece + (piece - wgt) > 2.0 * mean) {
            if (rank + 1 < nranks) {
                ++rank;
                mean = (total - assigned) * power(rank) / power_from(rank);
            }   // else: the reference keeps everything that is left on the last rank
            piece = wgt;
        }
        owner[k] = rank;
        assigned += wgt;
    }
    return SWCU_OK;
}

// create_uniform_decomposition (core/decomposition.f90:614-670): a px x py process grid, each rank
// owning a (bnx/px) x (bny/py) rectangle of blocks; rank = cart rank with y fastest (MPI_Cart_create
// row-major over (x, y)); -1 for blocks of zero weight.
int swh_uniform_partition(int bnx, int bny, int px, int py, const double *weights, int *owner)
{
    if (!owner || bnx < 1 || bny < 1 || px < 1 || py < 1 || bnx % px || bny % py) return SWCU_ERR_ARG;
    const int lx = bnx / px, ly = bny / py;
    for (int bn = 0; bn < bny; ++bn)
        for (int bm = 0; bm < bnx; ++bm) {
            const size_t k = (size_t)bn * bnx + bm;
            owner[k] = (weights && weights[k] == 0.0) ? -1 : (bm / lx) * py + (bn / ly);
        }
    return SWCU_OK;
}

// y-slabs of about equal WORK instead of equal height: the sea-weight balancing of the reference's
// load-balanced decomposition (core/decomposition.f90:505-520,560-600) applied to the one-block-per-GPU
// slab cut.  The unit of work is what the fused step really skips: a tile_cols x band_rows tile
// whose cells are all land costs `land_cost` of a tile with sea in it.  Rows are cut at band
// boundaries counted from the first computational row, so every slab's tile grid coincides with the
// global one.  start/size as in swh_uniform_split (offsets in computational rows), for all ranks.
int swh_balanced_slabs(int nx, int ny, const int *mask, int nranks, int band_rows, int tile_cols, double land_cost,
                       int *start, int *size)
{
    if (!start || !size || nx < 5 || ny < 5 || nranks < 1 || band_rows < 1 || tile_cols < 1) return SWCU_ERR_ARG;
    const int rows = ny - 4, cols = nx - 4;
    const int nbands = (rows + band_rows - 1) / band_rows, ntx = (cols + tile_cols - 1) / tile_cols;
    if (nbands < nranks) return SWCU_ERR_ARG;
    std::vector<double> w((size_t)nbands, 0.0);
    double total = 0.0;
    for (int b = 0; b < nbands; ++b) {
        const int n0 = 3 + b * band_rows, n1 = imin(n0 + band_rows - 1, ny - 2);
        for (int t = 0; t < ntx; ++t) {
            const int m0 = 3 + t * tile_cols, m1 = imin(m0 + tile_cols - 1, nx - 2);
            bool sea = mask == nullptr;
            for (int n = n0; n <= n1 && !sea; ++n)
                for (int m = m0; m <= m1 && !sea; ++m) sea = mask[(size_t)(n - 1) * nx + (m - 1)] == 0;
            w[b] += sea ? 1.0 : land_cost;
        }
        total += w[b];
    }
    int b = 0;
    double cum = 0.0;
    for (int r = 0; r < nranks; ++r) {
        const int first = b;
        const double target = total * (double)(r + 1) / (double)nranks;
        // at least one band each, and leave one band for every rank still to come
        do { cum += w[b]; ++b; } while (b < nbands - (nranks - 1 - r) && cum + 0.5 * w[b] <= target);
        if (r == nranks - 1) b = nbands;
        start[r] = first * band_rows;
        size[r] = imin(b * band_rows, rows) - start[r];
    }
    return SWCU_OK;
}

}  // extern "C"
