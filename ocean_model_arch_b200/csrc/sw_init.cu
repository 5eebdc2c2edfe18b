// sw_init.cu -- device-side construction of a block's static inputs: what
// init_grid_data leaves in grid_data (control/init_data.f90:96-125) without building the 2-D arrays
// on the host and uploading them.
//
//   masks    lu from the global integer mask (sea where it is 0, kernel/service/grid_kernels.f90:18-54;
//            "none" = rectangular basin with a 2-cell land frame, tools/io.f90:49-59), then the six
//            derived masks with the reference's `> 0.5` tests on sums / products of lu over
//            [bnd1 .. bnd2-1] (grid_kernels.f90:56-92).  Integer logic: bit-exact by construction.
//   metrics  for grids whose metric arrays are constant along x (carthesian, or spherical without a
//            rotated pole) the host evaluates ONE column with the very code of swh_metrics -- so the
//            libm cos / sin calls and every real(4) rounding are the same -- and the device copies
//            that column across the row, honouring the "filled on [2..nx-1] x [2..ny-1] only" rule
//            (grid_kernels.f90:152-198; rlh_s is set on the whole array first, :201).
//   fill / copy  data2D%fill and %copy_from act on the whole block array incl. frame
//            (core/data_types.f90:665-716).
#include "sw_fused.h"

namespace swcu {

namespace {

inline int launched_init(const char *what)
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SWCU_OK : cuda_fail(e, what);
}

// lu as one byte per array cell.  `land`: the block's window of the global mask (w x h ints, row
// pitch w) or nullptr for the frame rule.
__global__ void k_init_lu(Geo g, int w, int h, int nx, int ny, const int *__restrict__ land, unsigned char *__restrict__ lu)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= w || j >= h) return;
    const int m = g.bx1 + i, n = g.by1 + j;
    int is_land;
    if (land) is_land = land[(long)j * w + i];
    else is_land = (m < 3 || m > nx - 2 || n < 3 || n > ny - 2) ? 1 : 0;
    lu[(long)j * w + i] = is_land == 0 ? 1 : 0;
}

// The seven masks of one cell from the lu bytes.  bits != nullptr: FUSED byte plane (pitch g.pitch);
// otherwise the seven real(4) planes of REFERENCE mode.
__global__ void k_init_masks(Geo g, int w, int h, const unsigned char *__restrict__ lu, unsigned char *__restrict__ bits,
                             float *__restrict__ f_lu, float *__restrict__ f_luu, float *__restrict__ f_luh,
                             float *__restrict__ f_lcu, float *__restrict__ f_lcv, float *__restrict__ f_llu,
                             float *__restrict__ f_llv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= w || j >= h) return;
    const int a = lu[(long)j * w + i];
    int out = a ? MB_LU : 0;
    if (i < w - 1 && j < h - 1) {   // derived masks live on [bnd1 .. bnd2-1]; the last row / column stays 0
        const int e = lu[(long)j * w + i + 1], no = lu[(long)(j + 1) * w + i], en = lu[(long)(j + 1) * w + i + 1];
        if (a + e + no + en > 0) out |= MB_LUH;   // lu + lu + lu + lu > 0.5
        if (a & e & no & en) out |= MB_LUU;       // product > 0.5
        if (a + e > 0) out |= MB_LLU;
        if (a + no > 0) out |= MB_LLV;
        if (a & e) out |= MB_LCU;
        if (a & no) out |= MB_LCV;
    }
    const long c = (long)j * g.pitch + i;
    if (bits) { bits[c] = (unsigned char)out; return; }
    f_lu[c] = (out & MB_LU) ? 1.0f : 0.0f;   f_luu[c] = (out & MB_LUU) ? 1.0f : 0.0f;
    f_luh[c] = (out & MB_LUH) ? 1.0f : 0.0f; f_lcu[c] = (out & MB_LCU) ? 1.0f : 0.0f;
    f_lcv[c] = (out & MB_LCV) ? 1.0f : 0.0f; f_llu[c] = (out & MB_LLU) ? 1.0f : 0.0f;
    f_llv[c] = (out & MB_LLV) ? 1.0f : 0.0f;
}

struct RowFill { float *dst[9]; };

// prof: 9 arrays x 2 columns x h rows: [k][0][j] = value on columns outside [2..nx-1], [k][1][j] inside
__global__ void k_expand_rows(Geo g, int w, int h, int nx, RowFill out, const float *__restrict__ prof)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= w || j >= h) return;
    const int m = g.bx1 + i;
    const int inside = (m >= 2 && m <= nx - 1) ? 1 : 0;
    const long c = (long)j * g.pitch + i;
#pragma unroll
    for (int k = 0; k < 9; ++k) out.dst[k][c] = prof[((long)k * 2 + inside) * h + j];
}

// Halo push over peer memory (NVLink): CTA (k, side) copies `count` doubles of array k -- the rows the
// neighbour on `side` lacks -- from this GPU's buffer into the neighbour's halo rows through a
// peer-mapped pointer.  The last CTA to finish publishes `value` in both neighbours' flag words
// (system-scope fence first), which their streams wait on with a stream memory operation.
__global__ void k_push_halo(PushArgs p)
{
    const int k = blockIdx.x, side = blockIdx.y;
    const double *src = p.src[side][k];
    double *dst = p.dst[side][k];
    if (src && dst)
        for (long i = threadIdx.x; i < p.count; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(p.counter, 1u);
        if (done == gridDim.x * gridDim.y - 1) {
            *p.counter = 0;
            __threadfence_system();
            for (int s = 0; s < 2; ++s)
                if (p.flag[s]) *reinterpret_cast<volatile unsigned long long *>(p.flag[s]) = p.value;
        }
    }
}

__global__ void k_signal(unsigned long long *a, unsigned long long *b, unsigned long long value)
{
    __threadfence_system();
    if (a) *reinterpret_cast<volatile unsigned long long *>(a) = value;
    if (b) *reinterpret_cast<volatile unsigned long long *>(b) = value;
}

template <typename T>
__global__ void k_fill(int w, int h, int pitch, T *__restrict__ dst, T value)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= w || j >= h) return;
    dst[(long)j * pitch + i] = value;
}

inline dim3 grid2d(int w, int h) { return dim3((unsigned)((w + 255) / 256), (unsigned)h, 1); }

}  // namespace

int launch_init_lu(const Geo &g, int w, int h, int nx, int ny, const int *land_dev, unsigned char *lu_dev, cudaStream_t st)
{
    k_init_lu<<<grid2d(w, h), 256, 0, st>>>(g, w, h, nx, ny, land_dev, lu_dev);
    return launched_init("init_lu");
}

int launch_init_masks(const Geo &g, int w, int h, const unsigned char *lu_dev, unsigned char *bits, float *const f[7],
                      cudaStream_t st)
{
    k_init_masks<<<grid2d(w, h), 256, 0, st>>>(g, w, h, lu_dev, bits, f ? f[0] : nullptr, f ? f[1] : nullptr,
                                               f ? f[2] : nullptr, f ? f[3] : nullptr, f ? f[4] : nullptr,
                                               f ? f[5] : nullptr, f ? f[6] : nullptr);
    return launched_init("init_masks");
}

int launch_expand_rows(const Geo &g, int w, int h, int nx, float *const dst[9], const float *prof_dev, cudaStream_t st)
{
    RowFill out;
    for (int k = 0; k < 9; ++k) out.dst[k] = dst[k];
    k_expand_rows<<<grid2d(w, h), 256, 0, st>>>(g, w, h, nx, out, prof_dev);
    return launched_init("expand_rows");
}

int launch_push_halo(const PushArgs &p, int narrays, cudaStream_t st)
{
    k_push_halo<<<dim3((unsigned)narrays, 2, 1), 256, 0, st>>>(p);
    return launched_init("push_halo");
}

int launch_signal(unsigned long long *a, unsigned long long *b, unsigned long long value, cudaStream_t st)
{
    k_signal<<<1, 1, 0, st>>>(a, b, value);
    return launched_init("signal");
}

int launch_fill8(int w, int h, int pitch, double *dst, double value, cudaStream_t st)
{
    k_fill<double><<<grid2d(w, h), 256, 0, st>>>(w, h, pitch, dst, value);
    return launched_init("fill");
}
int launch_fill4(int w, int h, int pitch, float *dst, float value, cudaStream_t st)
{
    k_fill<float><<<grid2d(w, h), 256, 0, st>>>(w, h, pitch, dst, value);
    return launched_init("fill");
}

}  // namespace swcu
