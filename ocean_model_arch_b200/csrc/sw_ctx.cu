// sw_ctx.cu -- Level B: the resident context (include/swcuda.h).  One block of the reference's
// decomposition lives on one GPU; fields stay in HBM between steps (no per-step host round trip).
//
// Device layout: every field is a pitched 2-D array, pitch = width rounded up to 16 elements, so
// each row starts on a 128-byte (fp64) / 64-byte (fp32) boundary; element (m,n) of the
// reference's A(bnd_x1:bnd_x2, bnd_y1:bnd_y2) is at base[(n-bnd_y1)*pitch + (m-bnd_x1)].
//
// REFERENCE mode keeps all 32 real(8) + 17 real(4) arrays of ocean_type / grid_type and launches
// the reference's kernel sequence (control/shallow_water/shallow_water.f90:22-94).
// FUSED mode keeps the six prognostic arrays twice (ping-pong), hhq_rest, mu, six scratch arrays,
// ten real(4) arrays, per-row metric tables and one mask byte per cell; a step is ONE launch of the
// TMA-tiled kernel (or prep + update when the metrics vary along x) plus one launch per tracer step.
//
// Neighbouring blocks, three ways (one per context):
//   communicator  one block per process and GPU, y-slabs: ncclSend/ncclRecv of halo rows on a side
//                 stream (replaces shared/mpp/sync.f90:294-374 + syncborder_block2D_gen_all.fi)
//   peer memory   the same cut for processes on one node: boundary rows are stored straight into the
//                 neighbours' buffers (CUDA IPC), streams wait on step counters
//   links         several blocks of ONE process, any bnx x bny cut, one GPU or several: strided
//                 device-to-device pulls ordered by events (syncborder_block2D_gen_all.fi:218-249)
// In FUSED mode the two boundary strips are computed first and the exchange overlaps the interior.
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "sw_fast.cuh"
#include "sw_fused.h"

using namespace swcu;

namespace {

// ---- NCCL through dlopen: libswcuda.so has no link-time NCCL dependency; single-GPU users never
// load it, multi-GPU hosts get whatever libnccl.so.2 the process already holds (torch's) --------
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load()
{
    if (g_nccl.h) return SWCU_OK;
    const char *names[] = {getenv("SWCU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { set_error("cannot dlopen libnccl.so.2 (set SWCU_NCCL_LIB): %s", dlerror()); return SWCU_ERR_NCCL; }
#define L(sym)                                                                        \
    *(void **)(&g_nccl.sym) = dlsym(h, "nccl" #sym);                                  \
    if (!g_nccl.sym) { set_error("libnccl lacks nccl" #sym); return SWCU_ERR_NCCL; }
    L(GetUniqueId) L(CommInitRank) L(CommDestroy) L(GroupStart) L(GroupEnd) L(Send) L(Recv) L(GetErrorString)
#undef L
    g_nccl.h = h;
    return SWCU_OK;
}
int nccl_fail(ncclResult_t r, const char *what)
{
    set_error("NCCL error %d (%s) in %s", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
    return SWCU_ERR_NCCL;
}
#define SWCU_NCCL(call)                                        \
    do {                                                       \
        ncclResult_t r__ = (call);                             \
        if (r__ != ncclSuccess) return nccl_fail(r__, #call);  \
    } while (0)

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int encode_load()
{
    if (g_encode) return SWCU_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        cudaGetLastError();
        return SWCU_ERR_CUDA;
    }
    g_encode = (EncodeTiledFn)fn;
    return SWCU_OK;
}

// cuStreamWaitValue64 (stream memory operation) through the runtime's driver entry point
typedef CUresult (*WaitValueFn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
WaitValueFn g_wait_value = nullptr;

int wait_value_load()
{
    if (g_wait_value) return SWCU_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuStreamWaitValue64", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        set_error("cuStreamWaitValue64 not available from the driver");
        cudaGetLastError();
        return SWCU_ERR_CUDA;
    }
    g_wait_value = (WaitValueFn)fn;
    return SWCU_OK;
}

// cuStreamWriteValue64: a stream-ordered 8-byte store (here: into a neighbour's peer-mapped flag word)
WaitValueFn g_write_value = nullptr;
int write_value_load()
{
    if (g_write_value) return SWCU_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuStreamWriteValue64", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        set_error("cuStreamWriteValue64 not available from the driver");
        cudaGetLastError();
        return SWCU_ERR_CUDA;
    }
    g_write_value = (WaitValueFn)fn;
    return SWCU_OK;
}

#define RC(call) do { if (int rc__ = (call)) return rc__; } while (0)

// field-by-field (the struct has padding bytes a Fortran or C caller need not zero)
bool same_params(const swcu_params &a, const swcu_params &b)
{
    return a.full_free_surface == b.full_free_surface && a.trans_terms == b.trans_terms && a.ksw_lat == b.ksw_lat &&
           a.time_smooth == b.time_smooth && a.use_tracers == b.use_tracers && a.mode == b.mode;
}

const int kState[6] = {SWCU_F_SSH, SWCU_F_SSHP, SWCU_F_UBRTR, SWCU_F_UBRTRP, SWCU_F_VBRTR, SWCU_F_VBRTRP};

int mask_bit(int field)
{
    switch (field) {
        case SWCU_F_LU: return MB_LU;
        case SWCU_F_LCU: return MB_LCU;
        case SWCU_F_LCV: return MB_LCV;
        case SWCU_F_LUU: return MB_LUU;
        case SWCU_F_LUH: return MB_LUH;
        case SWCU_F_LLU: return MB_LLU;
        case SWCU_F_LLV: return MB_LLV;
        default: return 0;
    }
}

}  // namespace

struct swcu_ctx {
    swcu_dims d;
    swcu_params p;
    int device = 0;
    Geo g;
    int pitch = 0, w = 0, h = 0;
    size_t plane = 0;  // pitch * h elements
    cudaStream_t st = nullptr, comm_st = nullptr, bnd_st = nullptr;
    cudaEvent_t ev_bnd = nullptr, ev_comm = nullptr, ev_start = nullptr, t0 = nullptr, t1 = nullptr;
    double *f8[SWCU_NF8] = {};
    float *f4[SWCU_NF4] = {};
    double *alt[6] = {};  // FUSED: second copy of the prognostic arrays (ping-pong)
    double *alt_ff[2] = {};  // FUSED + tracers: second copy of ff1, ff1p
    // tracer_num > 1 (control/tracer.f90:42, core/ocean.f90:91-94): every tracer has its own planes; the ones of
    // tracer `tr_bound` are those f8[SWCU_F_FF1 / FF1N / FF1P] and alt_ff[] point at; `tr_sel` is the tracer the
    // caller's uploads / downloads address (option "tracer_select")
    struct TracerPlanes { double *ff = nullptr, *ffn = nullptr, *ffp = nullptr, *alt[2] = {nullptr, nullptr}; };
    std::vector<TracerPlanes> tr;
    int tr_bound = 0, tr_sel = 0;
    unsigned char *mask = nullptr;
    bool alt_dirty = true;
    bool has_rhs = false, has_rdiss = false;
    int *bad_dev = nullptr;
    int *bad_host = nullptr;  // pinned
    long launches = 0;
    long bytes = 0;
    long steps_done = 0;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    // in-process neighbours (swcu_link), by direction: S N W E SW SE NW NE
    swcu_ctx *nbr[8] = {};
    int nlinks = 0;
    FusedArgs fa;  // arguments of the step in flight (FUSED)
    // halo exchange over peer memory between PROCESSES (swcu_peer_export / swcu_peer_attach)
    struct PeerLink {
        bool on = false;
        double *set[2][8] = {};               // the neighbour's planes, peer-mapped: [set][ssh sshp u up v vp ff1 ff1p]
        unsigned long long *flags = nullptr;  // the neighbour's flag words, peer-mapped
        int by1 = 0;
        std::vector<void *> opened;
    } peer[2];                                // 0 = below (rank-1), 1 = above (rank+1)
    unsigned long long *flags = nullptr;      // mine: READY lo/hi, FREE lo/hi, READY_FF lo/hi
    unsigned *push_count = nullptr;           // counters: [0,1] k_push_halo, [2,3] strip warps of k_march, [5] timeout flag
    double *set_ptr[2][8] = {};               // my planes in export order
    int cur_set = 0;                          // set_ptr[cur_set] holds the current state
    // per-row metric tables (FUSED): rebuilt after a metric upload, used when all arrays are row-constant
    bool metrics_dirty = true, want_tables = true, use_tables = false;
    double *tab = nullptr;
    const float **arr_list_dev = nullptr;
    int *nonrow_dev = nullptr;
    bool masks_dirty = true, want_land_skip = true;  // all-land tile flags, rebuilt after a mask upload
    unsigned char *tile_land = nullptr;
    size_t tile_land_cap = 0;
    int tile_land_n0 = 0, tile_land_n1 = -1;  // row range of the launch the flags describe
    bool want_tiled = true;                   // one-launch TMA-tiled step when the tables are usable
    int tile_variant = 4;
    // tolerance mode (sw_fast.cuh / k_march): "exact" = 0.  The coefficient table depends on tau.
    bool exact = false;
    double *fc = nullptr, *ft = nullptr;
    double fc_tau = 0.0;
    bool fc_valid = false;
    int march_warps = 0;                      // SMs x resident warps of k_march on this device
    int tracer_warps = 0;                     //   ... of k_tracer_march
    int march_minb = 2;
    MarchPlan plan_main = {0, -1, 0, 0, 0, 0, 0, 0, nullptr, 2};
    int plan_sides = -1;
    unsigned char *band_land = nullptr;
    size_t band_land_cap = 0;
    std::map<const void *, CUtensorMap> tmaps;  // TMA descriptors by array base pointer
    // per-launch event pairs, filled only inside swcu_profile_steps
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;  // begin, end, begin, end, ...
    std::vector<int> prof_kind;        // 0 = prep, 1 = update
};

namespace {

int dev_alloc(swcu_ctx *c, void **ptr, size_t bytes)
{
    // + slack: k_march's bulk copies read whole 256-byte row segments that may run past the last row's pitch
    SWCU_CUDA(cudaMalloc(ptr, bytes + 512));
    SWCU_CUDA(cudaMemsetAsync(*ptr, 0, bytes + 512, c->st));
    c->bytes += (long)bytes + 512;
    return SWCU_OK;
}
int alloc8(swcu_ctx *c, int f)
{
    if (c->f8[f]) return SWCU_OK;
    return dev_alloc(c, (void **)&c->f8[f], c->plane * sizeof(double));
}
int alloc4(swcu_ctx *c, int f)
{
    if (c->f4[f - 100]) return SWCU_OK;
    return dev_alloc(c, (void **)&c->f4[f - 100], c->plane * sizeof(float));
}
inline float *F4(swcu_ctx *c, int f) { return c->f4[f - 100]; }
bool is_f8(int f) { return f >= 0 && f < SWCU_NF8; }
bool is_f4(int f) { return f >= 100 && f < SWCU_F4_END; }
bool is_tracer_field(int f) { return f >= SWCU_F_FLUX_X && f <= SWCU_F_FF1P; }

struct Use { int dev; explicit Use(int d) { cudaGetDevice(&dev); if (dev != d) cudaSetDevice(d); else dev = -1; }
             ~Use() { if (dev >= 0) cudaSetDevice(dev); } };

// which real(8) fields the FUSED mode keeps resident
bool fused_keeps8(const swcu_ctx *c, int f)
{
    switch (f) {
        case SWCU_F_SSH: case SWCU_F_SSHP: case SWCU_F_UBRTR: case SWCU_F_UBRTRP: case SWCU_F_VBRTR: case SWCU_F_VBRTRP:
        case SWCU_F_HHQ_REST: case SWCU_F_MU:
        case SWCU_F_HHU: case SWCU_F_HHV: case SWCU_F_HHH: case SWCU_F_VORT: case SWCU_F_STR_T: case SWCU_F_STR_S:
            return true;
        case SWCU_F_RHSX: case SWCU_F_RHSY: return c->has_rhs;
        default: return false;
    }
}
bool fused_keeps4(const swcu_ctx *c, int f)
{
    if (f >= SWCU_F_DX && f <= SWCU_F_RLH_S) return true;
    if (f == SWCU_F_R_DISS) return c->has_rdiss;
    return false;
}

int ntracers(const swcu_ctx *c) { return c->p.use_tracers ? (c->tr.empty() ? 1 : (int)c->tr.size()) : 0; }

// makes tracer k the one the field ids FF1 / FF1N / FF1P and alt_ff address
void tracer_bind(swcu_ctx *c, int k)
{
    if (c->tr.empty() || k == c->tr_bound) return;
    swcu_ctx::TracerPlanes &cur = c->tr[c->tr_bound];
    cur.ff = c->f8[SWCU_F_FF1]; cur.ffn = c->f8[SWCU_F_FF1N]; cur.ffp = c->f8[SWCU_F_FF1P];
    cur.alt[0] = c->alt_ff[0]; cur.alt[1] = c->alt_ff[1];
    const swcu_ctx::TracerPlanes &nx = c->tr[k];
    c->f8[SWCU_F_FF1] = nx.ff; c->f8[SWCU_F_FF1N] = nx.ffn; c->f8[SWCU_F_FF1P] = nx.ffp;
    c->alt_ff[0] = nx.alt[0]; c->alt_ff[1] = nx.alt[1];
    c->tr_bound = k;
}
// the bound tracer's write buffers become its current ones (FUSED ping-pong)
void tracer_swap(swcu_ctx *c)
{
    double *t = c->f8[SWCU_F_FF1]; c->f8[SWCU_F_FF1] = c->alt_ff[0]; c->alt_ff[0] = t;
    t = c->f8[SWCU_F_FF1P]; c->f8[SWCU_F_FF1P] = c->alt_ff[1]; c->alt_ff[1] = t;
}

int state_slot(int f)
{
    for (int i = 0; i < 6; ++i) if (kState[i] == f) return i;
    return -1;
}

// ---- halo rows over NCCL (y-slabs) ------------------------------------------------------------
// Sends my first / last `nrows` interior rows to rank-1 / rank+1 and receives their rows into my
// lower / upper halo rows.  Must be called between ncclGroupStart/End.
template <typename T>
int exchange_rows(swcu_ctx *c, T *base, int nrows, cudaStream_t st)
{
    const ncclDataType_t dt = sizeof(T) == 8 ? ncclFloat64 : (sizeof(T) == 4 ? ncclFloat32 : ncclUint8);
    const size_t cnt = (size_t)nrows * c->pitch;
    const int peer[2] = {c->rank - 1, c->rank + 1};
    for (int side = 0; side < 2; ++side) {
        if (peer[side] < 0 || peer[side] >= c->nranks) continue;
        int srow = 0, rrow = 0;
        RC(swcu_halo_plan(&c->d, nrows, side, &srow, &rrow));
        SWCU_NCCL(g_nccl.Send(base + (size_t)srow * c->pitch, cnt, dt, peer[side], c->comm, st));
        SWCU_NCCL(g_nccl.Recv(base + (size_t)rrow * c->pitch, cnt, dt, peer[side], c->comm, st));
    }
    return SWCU_OK;
}

int prof_mark(swcu_ctx *c, int kind, bool begin, cudaStream_t st = nullptr)
{
    if (!c->prof) return SWCU_OK;
    cudaEvent_t e;
    SWCU_CUDA(cudaEventCreate(&e));
    SWCU_CUDA(cudaEventRecord(e, st ? st : c->st));
    c->prof_ev.push_back(e);
    if (begin) c->prof_kind.push_back(kind);
    return SWCU_OK;
}
#define PROF(kind, call) do { RC(prof_mark(c, kind, true)); RC(call); RC(prof_mark(c, kind, false)); } while (0)

int prepare_metrics(swcu_ctx *c);

// One envoke_<name>_kernel(k, param) of interface/shallow_water/sw_interface.f90:42-403 (and
// interface/tracer/tracer_interface.f90:28-96) on the context's resident arrays: the binder's choice of
// which ocean_data / grid_data array goes to which dummy argument, then the 1:1 kernel.
int envoke_kernel(swcu_ctx *c, int kid, double tau)
{
    const Geo &g = c->g;
    cudaStream_t st = c->st;
    double **F = c->f8;
    const swcu_params &p = c->p;
    float *lu = F4(c, SWCU_F_LU), *luu = F4(c, SWCU_F_LUU), *luh = F4(c, SWCU_F_LUH), *lcu = F4(c, SWCU_F_LCU),
          *lcv = F4(c, SWCU_F_LCV), *llu = F4(c, SWCU_F_LLU), *llv = F4(c, SWCU_F_LLV);
    float *dx = F4(c, SWCU_F_DX), *dy = F4(c, SWCU_F_DY), *dxt = F4(c, SWCU_F_DXT), *dyt = F4(c, SWCU_F_DYT),
          *dxh = F4(c, SWCU_F_DXH), *dyh = F4(c, SWCU_F_DYH), *dxb = F4(c, SWCU_F_DXB), *dyb = F4(c, SWCU_F_DYB);
    // With row-constant metrics the 1:1 kernels read the context's per-row tables (no conversions, exact
    // mdiv divisions); otherwise the real(4) arrays, exactly as Level A does.  Same bits either way.
    if (c->metrics_dirty) RC(prepare_metrics(c));
    const MetRow mrow{c->tab, c->h, 0};
    const MetRow *mr = c->use_tables ? &mrow : nullptr;
    c->launches++;
    switch (kid) {
        case SWCU_K_SW_UPDATE_SSH:  // sw_interface.f90:310-328
            return launch_sw_update_ssh(g, tau, lu, dx, dy, dxh, dyh, F[SWCU_F_HHU], F[SWCU_F_HHV], F[SWCU_F_SSHN],
                                        F[SWCU_F_SSHP], F[SWCU_F_UBRTR], F[SWCU_F_VBRTR], st, mr);
        case SWCU_K_HH_UPDATE:  // :145-169 (takes ssh)
            return launch_hh_update(g, lu, llu, llv, luh, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, F[SWCU_F_HHQ_N],
                                    F[SWCU_F_HHU_N], F[SWCU_F_HHV_N], F[SWCU_F_HHH_N], F[SWCU_F_SSH], F[SWCU_F_HHQ_REST],
                                    st, mr);
        case SWCU_K_UV_TRANS_VORT:  // :211-229
            return launch_uv_trans_vort(g, luu, dxt, dyt, dxb, dyb, F[SWCU_F_UBRTR], F[SWCU_F_VBRTR], F[SWCU_F_VORT], st, mr);
        case SWCU_K_UV_TRANS:  // :238-262
            return launch_uv_trans(g, lcu, lcv, luu, dxh, dyh, F[SWCU_F_UBRTR], F[SWCU_F_VBRTR], F[SWCU_F_VORT],
                                   F[SWCU_F_HHU], F[SWCU_F_HHV], F[SWCU_F_HHH], F[SWCU_F_RHSX_ADV], F[SWCU_F_RHSY_ADV], st, mr);
        case SWCU_K_STRESS_COMPONENTS:  // :110-134 (takes ubrtrp, vbrtrp)
            return launch_stress_components(g, lu, luu, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, F[SWCU_F_UBRTRP],
                                            F[SWCU_F_VBRTRP], F[SWCU_F_STR_T], F[SWCU_F_STR_S], st, mr);
        case SWCU_K_UV_DIFF2:  // :273-302
            return launch_uv_diff2(g, lcu, lcv, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, F[SWCU_F_MU], F[SWCU_F_STR_T],
                                   F[SWCU_F_STR_S], F[SWCU_F_HHQ], F[SWCU_F_HHH], F[SWCU_F_RHSX_DIF], F[SWCU_F_RHSY_DIF], st, mr);
        case SWCU_K_SW_UPDATE_UV: {  // :337-374
            Tau tt;
            int ex = 0;
            const double mant = frexp(tau, &ex);
            tt.tau = tau; tt.rtau = 1.0 / tau; tt.pow2 = (mant == 0.5 && tau > 1e-300 && tau < 1e300) ? 1 : 0; tt.exact = 1;
            return launch_sw_update_uv(g, tau, lcu, lcv, dxt, dyt, dxh, dyh, dxb, dyb, F[SWCU_F_HHU], F[SWCU_F_HHU_N],
                                       F[SWCU_F_HHU_P], F[SWCU_F_HHV], F[SWCU_F_HHV_N], F[SWCU_F_HHV_P], F[SWCU_F_HHH],
                                       F[SWCU_F_SSH], F[SWCU_F_UBRTR], F[SWCU_F_UBRTRN], F[SWCU_F_UBRTRP], F[SWCU_F_VBRTR],
                                       F[SWCU_F_VBRTRN], F[SWCU_F_VBRTRP], F4(c, SWCU_F_R_DISS), F4(c, SWCU_F_RLH_S),
                                       F[SWCU_F_RHSX], F[SWCU_F_RHSY], F[SWCU_F_RHSX_ADV], F[SWCU_F_RHSY_ADV],
                                       F[SWCU_F_RHSX_DIF], F[SWCU_F_RHSY_DIF], st, mr, &tt);
        }
        case SWCU_K_SW_NEXT_STEP:  // :384-403
            return launch_sw_next_step(g, p.time_smooth, lu, lcu, lcv, F[SWCU_F_SSH], F[SWCU_F_SSHN], F[SWCU_F_SSHP],
                                       F[SWCU_F_UBRTR], F[SWCU_F_UBRTRN], F[SWCU_F_UBRTRP], F[SWCU_F_VBRTR],
                                       F[SWCU_F_VBRTRN], F[SWCU_F_VBRTRP], st);
        case SWCU_K_HH_SHIFT:  // :181-203
            return launch_hh_shift(g, p.time_smooth, lu, llu, llv, luh, F[SWCU_F_HHQ], F[SWCU_F_HHQ_P], F[SWCU_F_HHQ_N],
                                   F[SWCU_F_HHU], F[SWCU_F_HHU_P], F[SWCU_F_HHU_N], F[SWCU_F_HHV], F[SWCU_F_HHV_P],
                                   F[SWCU_F_HHV_N], F[SWCU_F_HHH], F[SWCU_F_HHH_P], F[SWCU_F_HHH_N], st);
        case SWCU_K_HH_INIT:  // :42-75
            return launch_hh_init(g, p.full_free_surface, lu, llu, llv, luh, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb,
                                  F[SWCU_F_HHQ], F[SWCU_F_HHQ_P], F[SWCU_F_HHQ_N], F[SWCU_F_HHU], F[SWCU_F_HHU_P],
                                  F[SWCU_F_HHU_N], F[SWCU_F_HHV], F[SWCU_F_HHV_P], F[SWCU_F_HHV_N], F[SWCU_F_HHH],
                                  F[SWCU_F_HHH_P], F[SWCU_F_HHH_N], F[SWCU_F_SSH], F[SWCU_F_SSHP], F[SWCU_F_HHQ_REST], st, mr);
        case SWCU_K_CHECK_SSH_ERR:  // :93-102
            return launch_check_ssh_err(g, lu, F[SWCU_F_SSH], c->bad_dev, st);
        case SWCU_K_TRAN_DIFF_FLUXES:  // tracer_interface.f90:28-49
            if (!p.use_tracers) break;
            return launch_tran_diff_fluxes(g, lcu, lcv, dxt, dyt, dxh, dyh, F[SWCU_F_HHU], F[SWCU_F_HHV], F[SWCU_F_FF1],
                                           F[SWCU_F_UBRTR], F[SWCU_F_VBRTR], F[SWCU_F_MU], 1.0, F[SWCU_F_FLUX_X],
                                           F[SWCU_F_FLUX_Y], st);
        case SWCU_K_TRAN_DIFF_TRACER:  // :59-74
            if (!p.use_tracers) break;
            return launch_tran_diff_tracer(g, lu, dx, dy, tau, F[SWCU_F_HHQ_N], F[SWCU_F_HHQ_P], F[SWCU_F_FLUX_X],
                                           F[SWCU_F_FLUX_Y], F[SWCU_F_FF1P], F[SWCU_F_FF1N], st);
        case SWCU_K_TRACER_NEXT_STEP:  // :83-96
            if (!p.use_tracers) break;
            return launch_tracer_next_step(g, p.time_smooth, lu, F[SWCU_F_FF1N], F[SWCU_F_FF1P], F[SWCU_F_FF1], st);
        default: break;
    }
    c->launches--;
    set_error("unknown kernel id %d (or tracer kernel without use_tracers)", kid);
    return SWCU_ERR_ARG;
}

// The matching envoke_<name>_sync (sw_interface.f90:77-408, tracer_interface.f90:51-102): which fields
// are halo-synced after the kernel (width 1, like hybrid_sync).  Returns the count, -1 for a bad id.
int sync_list(int kid, int out[3])
{
    auto set = [&](std::initializer_list<int> l) { int n = 0; for (int f : l) out[n++] = f; return n; };
    switch (kid) {
        case SWCU_K_SW_UPDATE_SSH: return set({SWCU_F_SSHN});
        case SWCU_K_HH_UPDATE: return set({SWCU_F_HHU_N, SWCU_F_HHV_N, SWCU_F_HHH_N});
        case SWCU_K_UV_TRANS_VORT: return set({SWCU_F_VORT});
        case SWCU_K_UV_TRANS: return set({SWCU_F_HHU_P, SWCU_F_HHV_P, SWCU_F_HHH_P});  // "lazy"
        case SWCU_K_STRESS_COMPONENTS: return set({SWCU_F_STR_T, SWCU_F_STR_S});
        case SWCU_K_SW_UPDATE_UV: return set({SWCU_F_VBRTRN, SWCU_F_UBRTRN});
        case SWCU_K_HH_INIT: return set({SWCU_F_HHU, SWCU_F_HHV, SWCU_F_HHH});
        case SWCU_K_TRAN_DIFF_FLUXES: return set({SWCU_F_FLUX_X, SWCU_F_FLUX_Y});
        case SWCU_K_TRAN_DIFF_TRACER: return set({SWCU_F_FF1N});
        case SWCU_K_UV_DIFF2: case SWCU_K_SW_NEXT_STEP: case SWCU_K_HH_SHIFT: case SWCU_K_CHECK_SSH_ERR:
        case SWCU_K_TRACER_NEXT_STEP: return 0;  // empty syncs in the reference
        default: set_error("unknown kernel id %d", kid); return -1;
    }
}

int envoke_sync_linked(swcu_ctx *c, const int *f, int n);

int envoke_sync(swcu_ctx *c, int kid)
{
    int f[3];
    const int n = sync_list(kid, f);
    if (n < 0) return SWCU_ERR_ARG;
    if (c->nlinks && n) return envoke_sync_linked(c, f, n);
    if (!c->comm || n == 0) return SWCU_OK;
    SWCU_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i)
        if (int rc = exchange_rows(c, c->f8[f[i]], 1, c->st)) { g_nccl.GroupEnd(); return rc; }
    SWCU_NCCL(g_nccl.GroupEnd());
    return SWCU_OK;
}

#define ENVOKE(kid) do { RC(envoke_kernel(c, kid, tau)); RC(envoke_sync(c, kid)); } while (0)

// control/shallow_water/shallow_water.f90:22-94, then control/tracer.f90:44-61
int step_reference(swcu_ctx *c, double tau)
{
    const swcu_params &p = c->p;
    ENVOKE(SWCU_K_SW_UPDATE_SSH);
    if (p.full_free_surface > 0) ENVOKE(SWCU_K_HH_UPDATE);
    if (p.trans_terms > 0) { ENVOKE(SWCU_K_UV_TRANS_VORT); ENVOKE(SWCU_K_UV_TRANS); }
    if (p.ksw_lat > 0) { ENVOKE(SWCU_K_STRESS_COMPONENTS); ENVOKE(SWCU_K_UV_DIFF2); }
    ENVOKE(SWCU_K_SW_UPDATE_UV);
    ENVOKE(SWCU_K_SW_NEXT_STEP);
    if (p.full_free_surface > 0) { ENVOKE(SWCU_K_HH_SHIFT); ENVOKE(SWCU_K_HH_INIT); }
    ENVOKE(SWCU_K_CHECK_SSH_ERR);
    for (int k = 0; k < ntracers(c); ++k) {   // control/tracer.f90:42: do k = 1, tracer_num
        tracer_bind(c, k);
        ENVOKE(SWCU_K_TRAN_DIFF_FLUXES);
        ENVOKE(SWCU_K_TRAN_DIFF_TRACER);
        ENVOKE(SWCU_K_TRACER_NEXT_STEP);
    }
    tracer_bind(c, c->tr_sel);
    return SWCU_OK;
}

void fill_static_args(swcu_ctx *c, FusedArgs &a)
{
    a.dx = F4(c, SWCU_F_DX); a.dy = F4(c, SWCU_F_DY); a.dxt = F4(c, SWCU_F_DXT); a.dyt = F4(c, SWCU_F_DYT);
    a.dxh = F4(c, SWCU_F_DXH); a.dyh = F4(c, SWCU_F_DYH); a.dxb = F4(c, SWCU_F_DXB); a.dyb = F4(c, SWCU_F_DYB);
    a.rlh_s = F4(c, SWCU_F_RLH_S); a.rdis = c->has_rdiss ? F4(c, SWCU_F_R_DISS) : nullptr;
    a.mask = c->mask; a.bad = c->bad_dev;
    a.tile_land = nullptr;
}

// (re)builds the per-row metric tables and decides whether they may replace the 2-D arrays
int prepare_metrics(swcu_ctx *c)
{
    c->metrics_dirty = false;
    c->use_tables = false;
    c->fc_valid = false;
    if (!c->want_tables) return SWCU_OK;
    FusedArgs a;
    fill_static_args(c, a);
    if (!c->tab) {
        // + slack: tiles that hang over the top of the array index a few rows past h
        RC(dev_alloc(c, (void **)&c->tab, ((size_t)T_COUNT * c->h + 64) * sizeof(double)));
        RC(dev_alloc(c, (void **)&c->arr_list_dev, 9 * sizeof(float *)));
        RC(dev_alloc(c, (void **)&c->nonrow_dev, sizeof(int)));
    }
    const float *list[9] = {a.dx, a.dy, a.dxt, a.dyt, a.dxh, a.dyh, a.dxb, a.dyb, a.rlh_s};
    SWCU_CUDA(cudaMemcpyAsync(c->arr_list_dev, list, sizeof(list), cudaMemcpyHostToDevice, c->st));
    SWCU_CUDA(cudaMemsetAsync(c->nonrow_dev, 0, sizeof(int), c->st));
    RC(launch_build_tables(c->g, a, c->tab, c->h, c->nonrow_dev, c->arr_list_dev, c->st));
    int nonrow = 1;
    SWCU_CUDA(cudaMemcpyAsync(&nonrow, c->nonrow_dev, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    SWCU_CUDA(cudaStreamSynchronize(c->st));
    c->use_tables = nonrow == 0;
    return SWCU_OK;
}

// TMA descriptor of one pitched fp64 plane: dims (w, h), row stride pitch*8 B, box = tile + halo
int tensor_map_for(swcu_ctx *c, const double *base, CUtensorMap *out)
{
    auto it = c->tmaps.find(base);
    if (it != c->tmaps.end()) { *out = it->second; return SWCU_OK; }
    RC(encode_load());
    int bw = 0, bh = 0;
    step_tile_box(c->tile_variant, &bw, &bh);
    const cuuint64_t dims[2] = {(cuuint64_t)c->w, (cuuint64_t)c->h};
    const cuuint64_t strides[1] = {(cuuint64_t)c->pitch * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return SWCU_ERR_CUDA; }
    c->tmaps[base] = m;
    *out = m;
    return SWCU_OK;
}

// ---- halo rows over peer memory (processes on one node; swcu_peer_export / swcu_peer_attach) --------
// Pushes my first / last two interior rows of arrays first .. first+n-1 (export order) of the WRITE
// set into the neighbours' halo rows of THEIR write set and publishes `tick` in their flag word
// flag0 + (0: written by the block below them, 1: by the block above).  The push first waits until
// each neighbour has declared its write buffers free for this step (FREE words 2, 3).
int peer_push(swcu_ctx *c, cudaStream_t st, int first, int n, unsigned long long tick, int flag0)
{
    RC(wait_value_load());
    PushArgs p = {};
    const int wset = c->cur_set ^ 1;
    for (int side = 0; side < 2; ++side) {
        const swcu_ctx::PeerLink &pl = c->peer[side];
        if (!pl.on) continue;
        CUresult r = g_wait_value((CUstream)st, (CUdeviceptr)(c->flags + 2 + side), tick, 0 /* GEQ */);
        if (r != CUDA_SUCCESS) { set_error("cuStreamWaitValue64 failed with %d", (int)r); return SWCU_ERR_CUDA; }
        int srow = 0, rrow = 0;
        RC(swcu_halo_plan(&c->d, 2, side, &srow, &rrow));
        const long drow = (long)(c->d.bnd_y1 + srow) - pl.by1;   // same global rows in the neighbour's array
        for (int k = 0; k < n; ++k) {
            p.src[side][k] = c->set_ptr[wset][first + k] + (size_t)srow * c->pitch;
            p.dst[side][k] = pl.set[wset][first + k] + (size_t)drow * c->pitch;
        }
        p.flag[side] = pl.flags + flag0 + (side == 0 ? 1 : 0);   // I am the block above my lower neighbour
    }
    p.value = tick;
    p.count = 2L * c->pitch;
    p.counter = c->push_count + (flag0 ? 1 : 0);
    RC(launch_push_halo(p, n, st));
    c->launches++;
    return SWCU_OK;
}

// makes `st` wait until both neighbours have pushed their rows for this tick (flag words flag0, flag0+1)
int peer_wait(swcu_ctx *c, cudaStream_t st, int flag0, unsigned long long tick)
{
    RC(wait_value_load());
    for (int side = 0; side < 2; ++side) {
        if (!c->peer[side].on) continue;
        CUresult r = g_wait_value((CUstream)st, (CUdeviceptr)(c->flags + flag0 + side), tick, 0 /* GEQ */);
        if (r != CUDA_SUCCESS) { set_error("cuStreamWaitValue64 failed with %d", (int)r); return SWCU_ERR_CUDA; }
    }
    return SWCU_OK;
}

// The n -> n+1 update of the six prognostic arrays into the write buffers (no tracers, no swap).
int fused_main(swcu_ctx *c, double tau)
{
    const Geo &g = c->g;
    if (c->metrics_dirty) RC(prepare_metrics(c));
    if (c->alt_dirty) {  // the frame / land cells of the write buffers must equal the read buffers
        for (int i = 0; i < 6; ++i)
            SWCU_CUDA(cudaMemcpyAsync(c->alt[i], c->f8[kState[i]], c->plane * sizeof(double),
                                      cudaMemcpyDeviceToDevice, c->st));
        for (int k = 0; k < ntracers(c); ++k) {
            tracer_bind(c, k);
            SWCU_CUDA(cudaMemcpyAsync(c->alt_ff[0], c->f8[SWCU_F_FF1], c->plane * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            SWCU_CUDA(cudaMemcpyAsync(c->alt_ff[1], c->f8[SWCU_F_FF1P], c->plane * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
        }
        tracer_bind(c, c->tr_sel);
        c->alt_dirty = false;
    }
    FusedArgs &a = c->fa;
    a.ssh = c->f8[SWCU_F_SSH]; a.sshp = c->f8[SWCU_F_SSHP]; a.u = c->f8[SWCU_F_UBRTR]; a.up = c->f8[SWCU_F_UBRTRP];
    a.v = c->f8[SWCU_F_VBRTR]; a.vp = c->f8[SWCU_F_VBRTRP];
    a.ssh_o = c->alt[0]; a.sshp_o = c->alt[1]; a.u_o = c->alt[2]; a.up_o = c->alt[3]; a.v_o = c->alt[4]; a.vp_o = c->alt[5];
    a.h_r = c->f8[SWCU_F_HHQ_REST]; a.mu = c->f8[SWCU_F_MU];
    a.RHSx = c->has_rhs ? c->f8[SWCU_F_RHSX] : nullptr; a.RHSy = c->has_rhs ? c->f8[SWCU_F_RHSY] : nullptr;
    a.hu = c->f8[SWCU_F_HHU]; a.hv = c->f8[SWCU_F_HHV]; a.hh = c->f8[SWCU_F_HHH];
    a.vort = c->f8[SWCU_F_VORT]; a.str_t = c->f8[SWCU_F_STR_T]; a.str_s = c->f8[SWCU_F_STR_S];
    fill_static_args(c, a);
    a.tab = c->use_tables ? c->tab : nullptr; a.tab_h = c->h;
    a.fc = a.ft = nullptr;
    if (!c->exact && c->use_tables) {  // tolerance mode: per-row coefficients (they contain tau)
        if (!c->fc) {
            RC(dev_alloc(c, (void **)&c->fc, ((size_t)c->h + 4) * swf::FC_STRIDE * sizeof(double)));
            RC(dev_alloc(c, (void **)&c->ft, ((size_t)c->h + 4) * swf::FT_STRIDE * sizeof(double)));
        }
        if (!c->fc_valid || c->fc_tau != tau) {
            RC(launch_build_fast(c->tab, c->h, tau, c->fc, c->ft, c->st));
            c->fc_valid = true; c->fc_tau = tau; c->launches++;
        }
        a.fc = c->fc;
        a.ft = c->ft;
    }
    {   // x/tau == x*(1/tau) bitwise when tau is a power of two (exact scaling)
        int ex = 0;
        const double mant = frexp(tau, &ex);
        a.tau.tau = tau; a.tau.rtau = 1.0 / tau; a.tau.pow2 = (mant == 0.5 && tau > 1e-300 && tau < 1e300) ? 1 : 0;
        a.tau.exact = 1;
    }
    a.ts = c->p.time_smooth; a.ffs = (double)c->p.full_free_surface;
    a.trans = c->p.trans_terms > 0; a.lat = c->p.ksw_lat > 0;
    a.ff = a.ffp = nullptr; a.ff_o = a.ffp_o = nullptr;
    if (c->p.use_tracers) {
        a.ff = c->f8[SWCU_F_FF1]; a.ffp = c->f8[SWCU_F_FF1P]; a.ff_o = c->alt_ff[0]; a.ffp_o = c->alt_ff[1];
    }

    const int ns = g.ny_start, ne = g.ny_end;
    const bool march = a.fc != nullptr && march_supported(g, a);
    if (march && !c->march_warps) c->march_warps = march_resident_warps(c->device, c->march_minb);
    const bool tiled = !march && c->use_tables && c->want_tiled && step_tiled_supported(g, a);
    // rows of the main launch: everything, or the interior between the two boundary strips
    // neighbours in other processes: over NCCL (communicator) or over peer memory (swcu_peer_attach)
    const bool peers = c->peer[0].on || c->peer[1].on;
    const bool lo = peers ? c->peer[0].on : (c->comm && c->rank > 0);
    const bool hi = peers ? c->peer[1].on : (c->comm && c->rank + 1 < c->nranks);
    int main0 = ns, main1 = ne;
    if (lo) main0 = (ns + 1 < ne ? ns + 1 : ne) + 1;
    if (hi && main0 <= ne) main1 = (ne - 1 > main0 ? ne - 1 : main0) - 1;
    // tolerance mode over peer memory: ONE launch per step, the boundary strips and their push into the
    // neighbours' halo rows are part of k_march (MarchPeer)
    const bool fused_push = march && peers && (ne - ns + 1) >= 4;
    const int sides = fused_push ? (lo ? 1 : 0) + (hi ? 2 : 0) : 0;
    if (march && (c->masks_dirty || c->plan_main.n0 != main0 || c->plan_main.n1 != main1 || c->plan_sides != sides)) {
        // Geometry of the main launch and the all-land flags of its bands.  First with short bands (128 rows): if
        // enough of them are all land (they cost nothing and the block scheduler balances the rest) keep that;
        // otherwise one wave of equal bands, which has the least warm-up and no tail.
        int fine_rows = 128;
        if (const char *e = getenv("SWCU_BAND_ROWS")) fine_rows = atoi(e);
        bool fine = false;
        for (int pass = 0; pass < 2; ++pass) {
            const bool try_fine = pass == 0 && c->want_land_skip && fine_rows > 0 && (main1 - main0 + 1) >= 4 * fine_rows;
            if (pass == 0 && !try_fine) continue;
            march_plan(g, main0, main1, c->march_warps, &c->plan_main, try_fine ? fine_rows : 0);
            const size_t need = (size_t)(c->plan_main.nwarps > 0 ? c->plan_main.nwarps : 1);
            if (need > c->band_land_cap) {
                if (c->band_land) { cudaFree(c->band_land); c->bytes -= (long)c->band_land_cap; }
                c->band_land = nullptr; c->band_land_cap = 0;
                RC(dev_alloc(c, (void **)&c->band_land, need));
                c->band_land_cap = need;
            }
            if (main1 < main0) break;
            SWCU_CUDA(cudaMemsetAsync(c->nonrow_dev, 0, sizeof(int), c->st));
            RC(launch_band_land(g, c->mask, c->plan_main, c->band_land, c->nonrow_dev, c->st));
            if (try_fine) {
                int nland = 0;
                SWCU_CUDA(cudaMemcpyAsync(&nland, c->nonrow_dev, sizeof(int), cudaMemcpyDeviceToHost, c->st));
                SWCU_CUDA(cudaStreamSynchronize(c->st));
                if (nland * 10 >= c->plan_main.nwarps) { fine = true; break; }   // >= 10 % of the bands are all land
            } else break;
        }
        c->plan_sides = sides;
        if (sides && !fine) {
            // the strip kernel runs concurrently and holds ncol / 4 CTA slots per side for the first microseconds:
            // the CTAs of the main launch that are scheduled last (its last band per side) start that much later
            // and get that many rows less to do
            int cut = 24;   // measured at 2048^2 per GPU, 2 GPUs: cut 0 0.1279, 8 0.1288, 16 0.1223, 24 0.1189 ms (1 GPU: 0.1145)
            if (const char *e = getenv("SWCU_LATE_CUT")) cut = atoi(e);
            const int nb = c->plan_main.nbands, late = (sides & 1) + ((sides >> 1) & 1);
            if (cut > 0 && nb > late && (main1 - main0 + 1) / nb > 2 * cut) {
                c->plan_main.late_hi = late;
                c->plan_main.late_cut = cut;
            }
        }
        c->masks_dirty = false;
    }
    if (tiled && (c->masks_dirty || c->tile_land_n0 != main0 || c->tile_land_n1 != main1)) {
        // (re)build the all-land tile flags of the main launch
        int ntx = 0, nty = 0;
        step_tile_grid(g, c->tile_variant, main0, main1, &ntx, &nty);
        const size_t need = (size_t)ntx * nty;
        if (need > c->tile_land_cap) {
            if (c->tile_land) { cudaFree(c->tile_land); c->bytes -= (long)c->tile_land_cap; }
            c->tile_land = nullptr; c->tile_land_cap = 0;
            RC(dev_alloc(c, (void **)&c->tile_land, need ? need : 1));
            c->tile_land_cap = need ? need : 1;
        }
        RC(launch_tile_land(g, c->mask, c->tile_variant, main0, main1, c->tile_land, c->st));
        c->masks_dirty = false; c->tile_land_n0 = main0; c->tile_land_n1 = main1;
    }
    StepMaps maps;
    if (tiled) {
        const double *src[8] = {a.ssh, a.sshp, a.u, a.up, a.v, a.vp, a.h_r, a.mu};
        for (int k = 0; k < 8; ++k) RC(tensor_map_for(c, src[k], &maps.m[k]));
    } else if (!march) {
        PROF(0, launch_prep(g, a, ns - 1, ne + 1, c->st));
        c->launches++;
    }
    // rows [r0..r1] of the n+1 state: one tiled launch, or the update stage of the two-launch path
    auto rows = [&](int r0, int r1, cudaStream_t st) -> int {
        if (r1 < r0) return SWCU_OK;
        // the flags describe the tile grid of the main launch only; the boundary strips do not use them
        a.tile_land = (tiled && c->want_land_skip && r0 == main0 && r1 == main1) ? c->tile_land : nullptr;
        RC(prof_mark(c, 1, true, st));
        if (march) {
            MarchPlan pl;
            if (r0 == main0 && r1 == main1) {
                pl = c->plan_main;
                pl.band_land = c->want_land_skip ? c->band_land : nullptr;
            } else {
                march_plan(g, r0, r1, c->march_warps, &pl);  // boundary strip: a few rows, no flags
            }
            pl.minb = c->march_minb;
            RC(launch_march(g, a, pl, st));
        } else {
            RC(tiled ? launch_step_tiled(maps, g, a, r0, r1, c->tile_variant, st) : launch_update(g, a, r0, r1, st));
        }
        RC(prof_mark(c, 1, false, st));
        c->launches++;
        return SWCU_OK;
    };
    if (!lo && !hi) {
        RC(rows(ns, ne, c->st));
    } else if (fused_push) {
        // ---- tolerance arithmetic over peer memory.  Two launches per step that write disjoint rows and run
        // concurrently: the lean k_march over the interior rows on the compute stream, and the strip kernel
        // (k_march<PEER>: boundary rows + their push into the neighbours' halo rows + the step counter) on the
        // high-priority stream.  No stream ever waits for a neighbour: the strip warps do, inside the kernel.
        const unsigned long long tick = (unsigned long long)c->steps_done + 1;
        RC(write_value_load());
        // Bits for timing experiments only (1: strip warps do not wait for the neighbours' flags, 2: no stores into
        // the neighbours' memory, 4: no "free" signal / tracer wait) -- each of them breaks the exchange, so they
        // are read only from a build-time switch, never from the environment of a production run.
#ifdef SWCU_ENABLE_PEER_DBG
        int dbg = getenv("SWCU_PEER_DBG") ? atoi(getenv("SWCU_PEER_DBG")) : 0;
#else
        const int dbg = 0;
#endif
        // everything that read my write buffers (the previous step, an upload's copy) is on the compute stream
        SWCU_CUDA(cudaEventRecord(c->ev_start, c->st));
        SWCU_CUDA(cudaStreamWaitEvent(c->bnd_st, c->ev_start, 0));
        for (int side = 0; side < 2; ++side) {   // "my write buffers may be written from here on"
            if (!c->peer[side].on || (dbg & 4)) continue;
            CUresult r = g_write_value((CUstream)c->bnd_st, (CUdeviceptr)(c->peer[side].flags + (side == 0 ? 3 : 2)), tick, 0);
            if (r != CUDA_SUCCESS) { set_error("cuStreamWriteValue64 failed with %d", (int)r); return SWCU_ERR_CUDA; }
        }
        MarchPeer mp;
        memset(&mp, 0, sizeof(mp));
        const int e = ns + 1, s2 = ne - 1;
        mp.lo0 = ns; mp.lo1 = lo ? e : ns - 1; mp.hi0 = s2; mp.hi1 = hi ? ne : s2 - 1;
        const int wset = c->cur_set ^ 1;
        for (int side = 0; side < 2; ++side) {
            const swcu_ctx::PeerLink &pl = c->peer[side];
            if (!pl.on) continue;
            const long shift = (long)(c->d.bnd_y1 - pl.by1) * c->pitch;  // same global (m, n) in the neighbour's plane
            for (int k = 0; k < 6; ++k) mp.out[side][k] = pl.set[wset][k] + shift;
            mp.ready[side] = pl.flags + (side == 0 ? 1 : 0);   // I am the block above my lower neighbour
            mp.free_[side] = c->flags + 2 + side;
            mp.ready_in[side] = c->flags + side;
            mp.count[side] = c->push_count + 2 + side;
        }
        mp.tick = tick;
        mp.dbg = dbg;
        mp.timeout = reinterpret_cast<int *>(c->push_count + 5);
        // strips: band 0's warps take the lower strip, the last band's the upper one; no band rows
        MarchPlan sp;
        march_plan(g, ns, ns - 1, c->march_warps, &sp);
        sp.nbands = (lo && hi) ? 2 : 1;
        sp.nwarps = sp.ncol * sp.nbands;
        RC(prof_mark(c, 1, true, c->bnd_st));
        RC(launch_march(g, a, sp, c->bnd_st, &mp));
        RC(prof_mark(c, 1, false, c->bnd_st));
        SWCU_CUDA(cudaEventRecord(c->ev_bnd, c->bnd_st));
        c->launches++;
        // interior
        RC(rows(main0, main1, c->st));
        // whatever comes next on the compute stream (the next step, a tracer kernel, a download) sees the strips
        SWCU_CUDA(cudaStreamWaitEvent(c->st, c->ev_bnd, 0));
        // ... and only a tracer kernel, which reads the new state's halo rows right now, needs the neighbours' rows
        // on the stream; the next step's strip warps wait for them in the kernel
        if (c->p.use_tracers && !(dbg & 4)) RC(peer_wait(c, c->st, 0, tick));
    } else {
        // The two boundary strips (the rows each neighbour needs) run on a high-priority stream
        // concurrently with the interior update; their completion releases the exchange on a second
        // high-priority stream (NCCL) or the push into the neighbours' memory (peer path), and the
        // compute stream joins before the next step.
        const unsigned long long tick = (unsigned long long)c->steps_done + 1;
        int i0 = ns, i1 = ne;
        if (peers && tick > 1) RC(peer_wait(c, c->st, 0, tick - 1));   // (a no-op unless the previous step ran fused)
        if (peers)   // my write buffers may be written by the neighbours from here on (alt sync is done)
            RC(launch_signal(lo ? c->peer[0].flags + 3 : nullptr, hi ? c->peer[1].flags + 2 : nullptr, tick, c->st));
        SWCU_CUDA(cudaEventRecord(c->ev_start, c->st));
        SWCU_CUDA(cudaStreamWaitEvent(c->bnd_st, c->ev_start, 0));
        if (lo) { const int e = ns + 1 < ne ? ns + 1 : ne; RC(rows(ns, e, c->bnd_st)); i0 = e + 1; }
        if (hi && i0 <= ne) { const int s = ne - 1 > i0 ? ne - 1 : i0; RC(rows(s, ne, c->bnd_st)); i1 = s - 1; }
        if (peers) {
            RC(peer_push(c, c->bnd_st, 0, 6, tick, 0));
            SWCU_CUDA(cudaEventRecord(c->ev_bnd, c->bnd_st));
            RC(rows(i0, i1, c->st));
            SWCU_CUDA(cudaStreamWaitEvent(c->st, c->ev_bnd, 0));
            RC(peer_wait(c, c->st, 0, tick));   // my halo rows of the new state have arrived
        } else {
            SWCU_CUDA(cudaEventRecord(c->ev_bnd, c->bnd_st));
            SWCU_CUDA(cudaStreamWaitEvent(c->comm_st, c->ev_bnd, 0));
            SWCU_NCCL(g_nccl.GroupStart());
            for (int i = 0; i < 6; ++i)
                if (int rc = exchange_rows(c, c->alt[i], 2, c->comm_st)) { g_nccl.GroupEnd(); return rc; }
            SWCU_NCCL(g_nccl.GroupEnd());
            SWCU_CUDA(cudaEventRecord(c->ev_comm, c->comm_st));
            RC(rows(i0, i1, c->st));
            SWCU_CUDA(cudaStreamWaitEvent(c->st, c->ev_comm, 0));  // strips -> exchange -> here
        }
    }
    return SWCU_OK;
}

// expl_tracer (control/tracer.f90:44-61) on the state fused_main just wrote (and, with a communicator,
// just exchanged: the compute stream already waits on the exchange event)
int fused_tracer(swcu_ctx *c)
{
    c->fa.ff = c->f8[SWCU_F_FF1]; c->fa.ffp = c->f8[SWCU_F_FF1P]; c->fa.ff_o = c->alt_ff[0]; c->fa.ffp_o = c->alt_ff[1];
    if (c->fa.fc && march_supported(c->g, c->fa)) {   // tolerance mode: the marching tracer kernel
        MarchPlan pl;
        if (!c->tracer_warps) c->tracer_warps = march_tracer_resident_warps(c->device);
        if (c->plan_main.n0 == c->g.ny_start && c->plan_main.n1 == c->g.ny_end && !c->plan_main.late_cut &&
            c->plan_main.nwarps > c->march_warps) {
            pl = c->plan_main;                    // short bands over the same rows: same bands, same all-land flags
            pl.band_land = c->want_land_skip ? c->band_land : nullptr;
        } else {
            march_plan(c->g, c->g.ny_start, c->g.ny_end, c->tracer_warps, &pl);   // one wave of this kernel
        }
        RC(launch_tracer_march(c->g, c->fa, pl, c->st));
    } else {
        RC(launch_tracer(c->g, c->fa, c->g.ny_start, c->g.ny_end, c->st));
    }
    c->launches++;
    if (c->peer[0].on || c->peer[1].on) {
        const unsigned long long tick = (unsigned long long)c->steps_done + 1;
        RC(peer_push(c, c->st, 6, 2, tick, 4));
        RC(peer_wait(c, c->st, 4, tick));
    } else if (c->comm) {
        SWCU_NCCL(g_nccl.GroupStart());
        int rc = exchange_rows(c, c->alt_ff[0], 2, c->st);
        if (!rc) rc = exchange_rows(c, c->alt_ff[1], 2, c->st);
        ncclResult_t r = g_nccl.GroupEnd();
        if (rc) return rc;
        if (r != ncclSuccess) return nccl_fail(r, "ncclGroupEnd");
    }
    return SWCU_OK;
}

// the write buffers become the current state
void fused_swap(swcu_ctx *c)
{
    for (int i = 0; i < 6; ++i) { double *t = c->f8[kState[i]]; c->f8[kState[i]] = c->alt[i]; c->alt[i] = t; }
    c->cur_set ^= 1;
}

int step_fused(swcu_ctx *c, double tau)
{
    RC(fused_main(c, tau));
    for (int k = 0; k < ntracers(c); ++k) {   // control/tracer.f90:42: do k = 1, tracer_num
        tracer_bind(c, k);
        RC(fused_tracer(c));
        tracer_swap(c);
    }
    tracer_bind(c, c->tr_sel);
    fused_swap(c);
    return SWCU_OK;
}

// ---- in-process neighbours (swcu_link / swcu_step_group) ---------------------------------------
// Several blocks of one process -- on one GPU or on several -- exchange halos by pulling: block
// arrays use global indices, so the cells a block lacks are the same (m, n) in the neighbour's
// array and one strided device-to-device copy per direction moves them.  Replaces the same-rank
// block-to-block copy of shared/mpp/syncborder_block2D_gen_all.fi:218-249 (and _GPU_MULTI_).
// Directions: S N W E SW NE SE NW, so that k ^ 1 is the opposite of k.
const int kDirX[8] = {0, 0, -1, 1, -1, 1, 1, -1};
const int kDirY[8] = {-1, 1, 0, 0, -1, 1, -1, 1};

int copy2d(void *dst, size_t dpitch, int ddev, const void *src, size_t spitch, int sdev, size_t wbytes, size_t rows,
           cudaStream_t st)
{
    if (ddev == sdev) {
        SWCU_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, wbytes, rows, cudaMemcpyDeviceToDevice, st));
        return SWCU_OK;
    }
    cudaMemcpy3DPeerParms p = {};
    p.srcDevice = sdev; p.dstDevice = ddev;
    p.srcPtr = make_cudaPitchedPtr(const_cast<void *>(src), spitch, wbytes, rows);
    p.dstPtr = make_cudaPitchedPtr(dst, dpitch, wbytes, rows);
    p.extent = make_cudaExtent(wbytes, rows, 1);
    SWCU_CUDA(cudaMemcpy3DPeerAsync(&p, st));
    return SWCU_OK;
}

// copies the `hw` halo layers that lie towards direction k from neighbour k's interior
template <typename T>
int pull_halo(swcu_ctx *c, int k, T *dst, const T *src, int hw, cudaStream_t st)
{
    const swcu_ctx *p = c->nbr[k];
    const swcu_dims &d = c->d;
    int x0 = d.nx_start, x1 = d.nx_end, y0 = d.ny_start, y1 = d.ny_end;
    if (kDirX[k] > 0) { x0 = d.nx_end + 1; x1 = d.nx_end + hw; } else if (kDirX[k] < 0) { x0 = d.nx_start - hw; x1 = d.nx_start - 1; }
    if (kDirY[k] > 0) { y0 = d.ny_end + 1; y1 = d.ny_end + hw; } else if (kDirY[k] < 0) { y0 = d.ny_start - hw; y1 = d.ny_start - 1; }
    T *dp = dst + (size_t)(y0 - d.bnd_y1) * c->pitch + (x0 - d.bnd_x1);
    const T *sp = src + (size_t)(y0 - p->d.bnd_y1) * p->pitch + (x0 - p->d.bnd_x1);
    return copy2d(dp, (size_t)c->pitch * sizeof(T), c->device, sp, (size_t)p->pitch * sizeof(T), p->device,
                  (size_t)(x1 - x0 + 1) * sizeof(T), (size_t)(y1 - y0 + 1), st);
}

// Lockstep halo pull of `narr` arrays for a group.  Each block's compute stream waits until its
// neighbours have produced their interiors, pulls, and then waits until its neighbours have finished
// reading from it, so the next kernel may overwrite anything.  Nothing blocks the host.
template <typename Get>
int group_pull(swcu_ctx *const *cs, int n, int narr, int hw, Get get)
{
    if (narr == 0) return SWCU_OK;
    for (int i = 0; i < n; ++i) {
        if (!cs[i]->nlinks) continue;
        Use use(cs[i]->device);
        SWCU_CUDA(cudaEventRecord(cs[i]->ev_bnd, cs[i]->st));
    }
    for (int i = 0; i < n; ++i) {
        swcu_ctx *c = cs[i];
        if (!c->nlinks) continue;
        Use use(c->device);
        for (int k = 0; k < 8; ++k)
            if (c->nbr[k]) SWCU_CUDA(cudaStreamWaitEvent(c->st, c->nbr[k]->ev_bnd, 0));
        for (int k = 0; k < 8; ++k) {
            if (!c->nbr[k]) continue;
            for (int a = 0; a < narr; ++a) RC(pull_halo(c, k, get(c, a), (const double *)get(c->nbr[k], a), hw, c->st));
        }
        SWCU_CUDA(cudaEventRecord(c->ev_comm, c->st));
    }
    for (int i = 0; i < n; ++i) {
        swcu_ctx *c = cs[i];
        if (!c->nlinks) continue;
        Use use(c->device);
        for (int k = 0; k < 8; ++k)
            if (c->nbr[k]) SWCU_CUDA(cudaStreamWaitEvent(c->st, c->nbr[k]->ev_comm, 0));
    }
    return SWCU_OK;
}

// Host-blocking pull for one block (per-block envoke_sync / swcu_halo_exchange on linked blocks): the
// caller has already issued the producing kernel on every neighbour, as the reference's loop over
// blocks does before its sync (core/kernel_interface.f90:62-86).
template <typename T, typename Get>
int pull_blocking(swcu_ctx *c, int hw, int narr, Get get)
{
    for (int k = 0; k < 8; ++k)
        if (c->nbr[k]) { Use use(c->nbr[k]->device); SWCU_CUDA(cudaStreamSynchronize(c->nbr[k]->st)); }
    for (int k = 0; k < 8; ++k) {
        if (!c->nbr[k]) continue;
        for (int a = 0; a < narr; ++a) {
            T *dst = get(c, a);
            const T *src = get(c->nbr[k], a);
            if (!dst || !src) { set_error("array not resident on a linked block"); return SWCU_ERR_STATE; }
            RC(pull_halo(c, k, dst, src, hw, c->st));
        }
    }
    SWCU_CUDA(cudaStreamSynchronize(c->st));
    return SWCU_OK;
}

int envoke_sync_linked(swcu_ctx *c, const int *f, int n)
{
    return pull_blocking<double>(c, 1, n, [f](const swcu_ctx *x, int a) { return x->f8[f[a]]; });
}

template <typename T>
int copy_in_rows(swcu_ctx *c, T *dst, const T *src, int row0, int nrows, cudaMemcpyKind kind)
{
    SWCU_CUDA(cudaMemcpy2DAsync(dst + (size_t)row0 * c->pitch, (size_t)c->pitch * sizeof(T), src, (size_t)c->w * sizeof(T),
                                (size_t)c->w * sizeof(T), (size_t)nrows, kind, c->st));
    return SWCU_OK;
}
template <typename T>
int copy_out(swcu_ctx *c, T *dst, const T *src, cudaMemcpyKind kind)
{
    SWCU_CUDA(cudaMemcpy2DAsync(dst, (size_t)c->w * sizeof(T), src, (size_t)c->pitch * sizeof(T),
                                (size_t)c->w * sizeof(T), (size_t)c->h, kind, c->st));
    return SWCU_OK;
}

// FUSED mode keeps no derived arrays.  For swcu_download they are rebuilt with the 1:1 kernels from the
// resident state into scratch planes: the depth fields from the current state (what K10 left), and
// vort / str_* / RHS*_adv / RHS*_dif from the PREVIOUS state, which the ping-pong buffers still hold
// (what K3..K6 of the last step computed).  Before the first step those five are zero, like the
// reference's fresh allocations.
int materialize(swcu_ctx *c, int field, double **out_plane, std::vector<void *> &scratch)
{
    auto plane8 = [&](double **p) -> int {
        SWCU_CUDA(cudaMalloc((void **)p, c->plane * sizeof(double)));
        scratch.push_back(*p);
        SWCU_CUDA(cudaMemsetAsync(*p, 0, c->plane * sizeof(double), c->st));
        return SWCU_OK;
    };
    float *mk[7];
    const int bits[7] = {MB_LU, MB_LUU, MB_LUH, MB_LCU, MB_LCV, MB_LLU, MB_LLV};
    for (int i = 0; i < 7; ++i) {
        SWCU_CUDA(cudaMalloc((void **)&mk[i], c->plane * sizeof(float)));
        scratch.push_back(mk[i]);
        RC(launch_mask_get((long)c->plane, mk[i], c->mask, bits[i], c->st));
    }
    float *lu = mk[0], *luu = mk[1], *luh = mk[2], *lcu = mk[3], *lcv = mk[4], *llu = mk[5], *llv = mk[6];
    float *dx = F4(c, SWCU_F_DX), *dy = F4(c, SWCU_F_DY), *dxt = F4(c, SWCU_F_DXT), *dyt = F4(c, SWCU_F_DYT),
          *dxh = F4(c, SWCU_F_DXH), *dyh = F4(c, SWCU_F_DYH), *dxb = F4(c, SWCU_F_DXB), *dyb = F4(c, SWCU_F_DYB);
    const bool depth = field >= SWCU_F_HHQ && field <= SWCU_F_HHH_N;
    const bool prev = !depth;
    if (prev && c->steps_done == 0) { RC(plane8(out_plane)); return SWCU_OK; }
    // state the requested field derives from
    const double *ssh = prev ? c->alt[0] : c->f8[SWCU_F_SSH], *sshp = prev ? c->alt[1] : c->f8[SWCU_F_SSHP];
    const double *u = prev ? c->alt[2] : c->f8[SWCU_F_UBRTR], *up = prev ? c->alt[3] : c->f8[SWCU_F_UBRTRP];
    const double *v = prev ? c->alt[4] : c->f8[SWCU_F_VBRTR], *vp = prev ? c->alt[5] : c->f8[SWCU_F_VBRTRP];
    double *hh[12];  // hq hqp hqn hu hup hun hv hvp hvn hh hhp hhn
    for (auto &p : hh) RC(plane8(&p));
    RC(launch_hh_init(c->g, c->p.full_free_surface, lu, llu, llv, luh, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb,
                      hh[0], hh[1], hh[2], hh[3], hh[4], hh[5], hh[6], hh[7], hh[8], hh[9], hh[10], hh[11],
                      ssh, sshp, c->f8[SWCU_F_HHQ_REST], c->st));
    if (depth) { *out_plane = hh[field - SWCU_F_HHQ]; return SWCU_OK; }
    double *vort, *str_t, *str_s, *ox, *oy;
    RC(plane8(&vort)); RC(plane8(&str_t)); RC(plane8(&str_s)); RC(plane8(&ox)); RC(plane8(&oy));
    if (c->p.trans_terms > 0) RC(launch_uv_trans_vort(c->g, luu, dxt, dyt, dxb, dyb, u, v, vort, c->st));
    if (c->p.ksw_lat > 0)
        RC(launch_stress_components(c->g, lu, luu, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, up, vp, str_t, str_s, c->st));
    switch (field) {
        case SWCU_F_VORT: *out_plane = vort; return SWCU_OK;
        case SWCU_F_STR_T: *out_plane = str_t; return SWCU_OK;
        case SWCU_F_STR_S: *out_plane = str_s; return SWCU_OK;
        case SWCU_F_RHSX_ADV: case SWCU_F_RHSY_ADV:
            if (c->p.trans_terms > 0)
                RC(launch_uv_trans(c->g, lcu, lcv, luu, dxh, dyh, u, v, vort, hh[3], hh[6], hh[9], ox, oy, c->st));
            *out_plane = field == SWCU_F_RHSX_ADV ? ox : oy;
            return SWCU_OK;
        case SWCU_F_RHSX_DIF: case SWCU_F_RHSY_DIF:
            if (c->p.ksw_lat > 0)
                RC(launch_uv_diff2(c->g, lcu, lcv, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, c->f8[SWCU_F_MU], str_t, str_s,
                                   hh[0], hh[9], ox, oy, c->st));
            *out_plane = field == SWCU_F_RHSX_DIF ? ox : oy;
            return SWCU_OK;
        default: break;
    }
    set_error("field %d cannot be materialised", field);
    return SWCU_ERR_STATE;
}

int upload_impl(swcu_ctx *c, int field, const void *src, bool from_device, int row0 = 0, int nrows = -1)
{
    if (!c || !src) { set_error("null argument"); return SWCU_ERR_ARG; }
    if (nrows < 0) nrows = c->h;
    if (row0 < 0 || nrows < 0 || row0 + nrows > c->h) { set_error("rows [%d, %d) outside the block (%d rows)", row0, row0 + nrows, c->h); return SWCU_ERR_ARG; }
    Use use(c->device);
    const cudaMemcpyKind kind = from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const bool fused = c->p.mode == SWCU_MODE_FUSED;
    if (is_f8(field)) {
        if (is_tracer_field(field) && !c->p.use_tracers) { set_error("tracer field without use_tracers"); return SWCU_ERR_STATE; }
        if (fused && (field == SWCU_F_RHSX || field == SWCU_F_RHSY) && !c->has_rhs) {
            c->has_rhs = true;  // the reference never assigns RHSx/RHSy; first upload makes them resident
            RC(alloc8(c, SWCU_F_RHSX)); RC(alloc8(c, SWCU_F_RHSY));
        }
        if (fused && is_tracer_field(field) && field != SWCU_F_FF1 && field != SWCU_F_FF1P) return SWCU_OK;
        if (fused && !fused_keeps8(c, field) && !is_tracer_field(field)) return SWCU_OK;  // derived: recomputed on device
        RC(copy_in_rows(c, c->f8[field], (const double *)src, row0, nrows, kind));
        if (fused && (state_slot(field) >= 0 || is_tracer_field(field))) c->alt_dirty = true;
        return SWCU_OK;
    }
    if (is_f4(field)) {
        const int bit = mask_bit(field);
        if (fused && bit) {
            // stage through a scratch real(4) plane, then fold into the mask byte
            float *tmp = nullptr;
            const size_t cnt = (size_t)nrows * c->pitch;
            SWCU_CUDA(cudaMalloc((void **)&tmp, cnt * sizeof(float)));
            SWCU_CUDA(cudaMemsetAsync(tmp, 0, cnt * sizeof(float), c->st));
            int rc = SWCU_OK;
            {
                cudaError_t e = cudaMemcpy2DAsync(tmp, (size_t)c->pitch * sizeof(float), src, (size_t)c->w * sizeof(float),
                                                  (size_t)c->w * sizeof(float), (size_t)nrows, kind, c->st);
                if (e != cudaSuccess) rc = cuda_fail(e, "mask upload");
            }
            if (!rc) rc = launch_mask_set((long)cnt, tmp, c->mask + (size_t)row0 * c->pitch, bit, c->st);
            c->masks_dirty = true;
            cudaStreamSynchronize(c->st);
            cudaFree(tmp);
            return rc;
        }
        if (fused && field == SWCU_F_R_DISS && !c->has_rdiss) { c->has_rdiss = true; RC(alloc4(c, SWCU_F_R_DISS)); }
        if (fused && !fused_keeps4(c, field)) return SWCU_OK;
        if (field >= SWCU_F_DX && field <= SWCU_F_RLH_S) c->metrics_dirty = true;
        return copy_in_rows(c, F4(c, field), (const float *)src, row0, nrows, kind);
    }
    set_error("unknown field id %d", field);
    return SWCU_ERR_ARG;
}

int download_impl(swcu_ctx *c, int field, void *dst, bool to_device)
{
    if (!c || !dst) { set_error("null argument"); return SWCU_ERR_ARG; }
    Use use(c->device);
    const cudaMemcpyKind kind = to_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const bool fused = c->p.mode == SWCU_MODE_FUSED;
    int rc = SWCU_OK;
    if (is_f8(field)) {
        const double *src = c->f8[field];
        if (fused) {
            // after K8 the n+1 arrays equal the current ones on every cell they were written
            if (field == SWCU_F_SSHN) src = c->f8[SWCU_F_SSH];
            else if (field == SWCU_F_UBRTRN) src = c->f8[SWCU_F_UBRTR];
            else if (field == SWCU_F_VBRTRN) src = c->f8[SWCU_F_VBRTR];
            else if (field == SWCU_F_FF1N && c->p.use_tracers) src = c->f8[SWCU_F_FF1];
            else if ((field >= SWCU_F_HHQ && field <= SWCU_F_HHH_N) || field == SWCU_F_VORT || field == SWCU_F_STR_T ||
                     field == SWCU_F_STR_S || (field >= SWCU_F_RHSX_ADV && field <= SWCU_F_RHSY_DIF)) {
                std::vector<void *> scratch;
                double *plane = nullptr;
                rc = materialize(c, field, &plane, scratch);
                if (!rc) rc = copy_out(c, (double *)dst, plane, kind);
                cudaStreamSynchronize(c->st);
                for (void *p : scratch) cudaFree(p);
                return rc;
            } else if (!fused_keeps8(c, field) && !((field == SWCU_F_FF1 || field == SWCU_F_FF1P) && c->p.use_tracers)) {
                set_error("field %d is not resident in FUSED mode", field);
                return SWCU_ERR_STATE;
            }
        }
        if (!src) { set_error("field %d not allocated", field); return SWCU_ERR_STATE; }
        rc = copy_out(c, (double *)dst, src, kind);
    } else if (is_f4(field)) {
        const int bit = mask_bit(field);
        if (fused && bit) {
            float *tmp = nullptr;
            SWCU_CUDA(cudaMalloc((void **)&tmp, c->plane * sizeof(float)));
            rc = launch_mask_get((long)c->plane, tmp, c->mask, bit, c->st);
            if (!rc) rc = copy_out(c, (float *)dst, tmp, kind);
            cudaStreamSynchronize(c->st);
            cudaFree(tmp);
            return rc;
        }
        const float *src = F4(c, field);
        if (!src) { set_error("field %d is not resident", field); return SWCU_ERR_STATE; }
        rc = copy_out(c, (float *)dst, src, kind);
    } else {
        set_error("unknown field id %d", field);
        return SWCU_ERR_ARG;
    }
    if (rc) return rc;
    SWCU_CUDA(cudaStreamSynchronize(c->st));
    return SWCU_OK;
}

}  // namespace

extern "C" {

int swcu_create(swcu_ctx **out, const swcu_dims *dims, const swcu_params *params, int device)
{
    if (!out || !params) { set_error("null argument"); return SWCU_ERR_ARG; }
    RC(check_dims(dims));
    if (params->mode != SWCU_MODE_REFERENCE && params->mode != SWCU_MODE_FUSED) { set_error("bad mode"); return SWCU_ERR_ARG; }
    int ndev = 0;
    SWCU_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("device %d of %d", device, ndev); return SWCU_ERR_ARG; }
    Use use(device);
    swcu_ctx *c = new swcu_ctx();
    c->d = *dims; c->p = *params; c->device = device;
    c->w = width(*dims); c->h = height(*dims);
    c->pitch = (c->w + 15) / 16 * 16;
    c->plane = (size_t)c->pitch * c->h;
    c->g = make_geo(*dims, c->pitch);
    if (const char *e = getenv("SWCU_EXACT")) c->exact = atoi(e) != 0;  // default arithmetic of new contexts
    int rc = SWCU_OK;
#define TRY(call) do { if (!rc) rc = (call); } while (0)
#define TRYCUDA(call) do { if (!rc) { cudaError_t e__ = (call); if (e__ != cudaSuccess) rc = cuda_fail(e__, #call); } } while (0)
    TRYCUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    {   // the halo-exchange stream outranks the compute stream so NCCL's copy kernel is scheduled as
        // soon as an SM frees up instead of queueing behind the interior update's CTAs
        int lo = 0, hi = 0;
        TRYCUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        TRYCUDA(cudaStreamCreateWithPriority(&c->comm_st, cudaStreamNonBlocking, hi));
        TRYCUDA(cudaStreamCreateWithPriority(&c->bnd_st, cudaStreamNonBlocking, hi));
    }
    TRYCUDA(cudaEventCreateWithFlags(&c->ev_bnd, cudaEventDisableTiming));
    TRYCUDA(cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming));
    TRYCUDA(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
    TRYCUDA(cudaEventCreate(&c->t0));
    TRYCUDA(cudaEventCreate(&c->t1));
    TRY(dev_alloc(c, (void **)&c->bad_dev, sizeof(int)));
    TRYCUDA(cudaHostAlloc((void **)&c->bad_host, sizeof(int), cudaHostAllocDefault));
    if (params->mode == SWCU_MODE_REFERENCE) {
        for (int f = 0; f < SWCU_NF8; ++f) {
            if (is_tracer_field(f) && !params->use_tracers) continue;
            TRY(alloc8(c, f));
        }
        for (int f = 100; f < SWCU_F4_END; ++f) TRY(alloc4(c, f));
    } else {
        for (int f = 0; f < SWCU_NF8; ++f) if (fused_keeps8(c, f)) TRY(alloc8(c, f));
        for (int i = 0; i < 6; ++i) TRY(dev_alloc(c, (void **)&c->alt[i], c->plane * sizeof(double)));
        if (params->use_tracers) {
            TRY(alloc8(c, SWCU_F_FF1)); TRY(alloc8(c, SWCU_F_FF1P));
            for (int i = 0; i < 2; ++i) TRY(dev_alloc(c, (void **)&c->alt_ff[i], c->plane * sizeof(double)));
        }
        for (int f = 100; f < SWCU_F4_END; ++f) if (fused_keeps4(c, f)) TRY(alloc4(c, f));
        TRY(dev_alloc(c, (void **)&c->mask, c->plane));
        TRY(dev_alloc(c, (void **)&c->flags, 8 * sizeof(unsigned long long)));
        TRY(dev_alloc(c, (void **)&c->push_count, 8 * sizeof(unsigned)));
        for (int i = 0; i < 6; ++i) { c->set_ptr[0][i] = c->f8[kState[i]]; c->set_ptr[1][i] = c->alt[i]; }
        if (params->use_tracers) {
            c->set_ptr[0][6] = c->f8[SWCU_F_FF1]; c->set_ptr[0][7] = c->f8[SWCU_F_FF1P];
            c->set_ptr[1][6] = c->alt_ff[0]; c->set_ptr[1][7] = c->alt_ff[1];
        }
    }
    TRYCUDA(cudaStreamSynchronize(c->st));
    if (rc) { swcu_destroy(c); return rc; }
    *out = c;
    return SWCU_OK;
}

int swcu_destroy(swcu_ctx *c)
{
    if (!c) return SWCU_OK;
    swcu_unlink(c);
    Use use(c->device);
    cudaDeviceSynchronize();
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (auto &pl : c->peer)
        for (void *p : pl.opened) cudaIpcCloseMemHandle(p);
    if (!c->tr.empty()) {
        tracer_bind(c, 0);
        for (size_t k = 1; k < c->tr.size(); ++k) {
            cudaFree(c->tr[k].ff); cudaFree(c->tr[k].ffn); cudaFree(c->tr[k].ffp); cudaFree(c->tr[k].alt[0]); cudaFree(c->tr[k].alt[1]);
        }
    }
    cudaFree(c->flags); cudaFree(c->push_count);
    for (auto &p : c->f8) cudaFree(p);
    for (auto &p : c->f4) cudaFree(p);
    for (auto &p : c->alt) cudaFree(p);
    for (auto &p : c->alt_ff) cudaFree(p);
    cudaFree(c->mask); cudaFree(c->bad_dev);
    cudaFree(c->fc); cudaFree(c->ft); cudaFree(c->band_land);
    cudaFree(c->tab); cudaFree(c->arr_list_dev); cudaFree(c->nonrow_dev); cudaFree(c->tile_land);
    if (c->bad_host) cudaFreeHost(c->bad_host);
    if (c->ev_bnd) cudaEventDestroy(c->ev_bnd);
    if (c->ev_comm) cudaEventDestroy(c->ev_comm);
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    if (c->bnd_st) cudaStreamDestroy(c->bnd_st);
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    if (c->st) cudaStreamDestroy(c->st);
    if (c->comm_st) cudaStreamDestroy(c->comm_st);
    delete c;
    return SWCU_OK;
}

int swcu_upload(swcu_ctx *c, int field, const void *host) { return upload_impl(c, field, host, false); }
int swcu_upload_from_device(swcu_ctx *c, int field, const void *dev) { return upload_impl(c, field, dev, true); }
int swcu_output_record(swcu_ctx *c, int field, float *host)
{
    if (!c || !host) { set_error("null argument"); return SWCU_ERR_ARG; }
    if (!is_f8(field)) { set_error("output records exist for real(8) fields only"); return SWCU_ERR_ARG; }
    Use use(c->device);
    const bool fused = c->p.mode == SWCU_MODE_FUSED;
    const double *src = c->f8[field];
    if (fused) {
        if (field == SWCU_F_SSHN) src = c->f8[SWCU_F_SSH];
        else if (field == SWCU_F_UBRTRN) src = c->f8[SWCU_F_UBRTR];
        else if (field == SWCU_F_VBRTRN) src = c->f8[SWCU_F_VBRTR];
        else if (field == SWCU_F_FF1N && c->p.use_tracers) src = c->f8[SWCU_F_FF1];
        else if (!fused_keeps8(c, field) && !((field == SWCU_F_FF1 || field == SWCU_F_FF1P) && c->p.use_tracers)) src = nullptr;
    }
    if (!src) { set_error("field %d is not resident", field); return SWCU_ERR_STATE; }
    const size_t n = (size_t)(c->d.nx_end - c->d.nx_start + 1) * (size_t)(c->d.ny_end - c->d.ny_start + 1);
    float *tmp = nullptr;
    SWCU_CUDA(cudaMalloc((void **)&tmp, n * sizeof(float)));
    int rc = launch_output_record(c->g, src, fused ? c->mask : nullptr, fused ? nullptr : F4(c, SWCU_F_LU), tmp, c->st);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(host, tmp, n * sizeof(float), cudaMemcpyDeviceToHost, c->st);
        if (e != cudaSuccess) rc = cuda_fail(e, "output record copy");
    }
    cudaStreamSynchronize(c->st);
    cudaFree(tmp);
    c->launches++;
    return rc;
}

int swcu_upload_rows(swcu_ctx *c, int field, const void *host, int first_row, int nrows)
{
    return upload_impl(c, field, host, false, first_row, nrows);
}
int swcu_download(swcu_ctx *c, int field, void *host) { return download_impl(c, field, host, false); }
int swcu_download_to_device(swcu_ctx *c, int field, void *dev) { return download_impl(c, field, dev, true); }

int swcu_set_option(swcu_ctx *c, const char *name, int value)
{
    if (!c || !name) { set_error("null argument"); return SWCU_ERR_ARG; }
    if (!strcmp(name, "metric_tables")) { c->want_tables = value != 0; c->metrics_dirty = true; return SWCU_OK; }
    if (!strcmp(name, "tiled")) { c->want_tiled = value != 0; return SWCU_OK; }
    if (!strcmp(name, "tile_variant")) {
        if (value < 1 || value > 5) { set_error("tile_variant must be 1..5"); return SWCU_ERR_ARG; }
        c->tile_variant = value; c->tmaps.clear(); c->masks_dirty = true;
        return SWCU_OK;
    }
    if (!strcmp(name, "land_skip")) { c->want_land_skip = value != 0; return SWCU_OK; }
    if (!strcmp(name, "tracer_num")) {   // sw.par line 7 (configs/sw.f90:41); once, before the tracers are uploaded
        if (!c->p.use_tracers) { set_error("tracer_num needs use_tracers"); return SWCU_ERR_STATE; }
        if (value < 1 || value > 64) { set_error("tracer_num must be 1..64"); return SWCU_ERR_ARG; }
        if (!c->tr.empty() || c->steps_done) { set_error("tracer_num is set once, before the first step"); return SWCU_ERR_STATE; }
        if (c->peer[0].on || c->peer[1].on) { set_error("the peer-memory halo path carries one tracer: use a communicator"); return SWCU_ERR_STATE; }
        if (value == 1) return SWCU_OK;
        Use use(c->device);
        const bool fused = c->p.mode == SWCU_MODE_FUSED;
        c->tr.resize((size_t)value);
        c->tr[0].ff = c->f8[SWCU_F_FF1]; c->tr[0].ffn = c->f8[SWCU_F_FF1N]; c->tr[0].ffp = c->f8[SWCU_F_FF1P];
        c->tr[0].alt[0] = c->alt_ff[0]; c->tr[0].alt[1] = c->alt_ff[1];
        for (int k = 1; k < value; ++k) {
            RC(dev_alloc(c, (void **)&c->tr[k].ff, c->plane * sizeof(double)));
            RC(dev_alloc(c, (void **)&c->tr[k].ffp, c->plane * sizeof(double)));
            if (fused) {
                RC(dev_alloc(c, (void **)&c->tr[k].alt[0], c->plane * sizeof(double)));
                RC(dev_alloc(c, (void **)&c->tr[k].alt[1], c->plane * sizeof(double)));
            } else {
                RC(dev_alloc(c, (void **)&c->tr[k].ffn, c->plane * sizeof(double)));
            }
        }
        SWCU_CUDA(cudaStreamSynchronize(c->st));
        c->tr_bound = c->tr_sel = 0;
        return SWCU_OK;
    }
    if (!strcmp(name, "tracer_select")) {   // which tracer (0-based) FF1 / FF1N / FF1P address in uploads and downloads
        if (value < 0 || value >= (ntracers(c) > 0 ? ntracers(c) : 1)) { set_error("tracer_select out of range"); return SWCU_ERR_ARG; }
        Use use(c->device);
        SWCU_CUDA(cudaStreamSynchronize(c->st));
        c->tr_sel = value;
        tracer_bind(c, value);
        return SWCU_OK;
    }
    if (!strcmp(name, "march_minb")) {
        if (value != 2 && value != 3) { set_error("march_minb must be 2 or 3"); return SWCU_ERR_ARG; }
        c->march_minb = value; c->march_warps = 0; c->plan_main.n1 = c->plan_main.n0 - 1; c->masks_dirty = true;
        return SWCU_OK;
    }
    if (!strcmp(name, "exact")) { c->exact = value != 0; c->plan_main.n1 = c->plan_main.n0 - 1; c->masks_dirty = true; return SWCU_OK; }
    set_error("unknown option %s", name);
    return SWCU_ERR_ARG;
}
int swcu_uses_metric_tables(const swcu_ctx *c) { return c && c->use_tables ? 1 : 0; }

int swcu_envoke_hh_init(swcu_ctx *c)
{
    if (!c) { set_error("null ctx"); return SWCU_ERR_ARG; }
    if (c->p.mode == SWCU_MODE_FUSED) return SWCU_OK;
    Use use(c->device);
    RC(envoke_kernel(c, SWCU_K_HH_INIT, 0.0));
    RC(envoke_sync(c, SWCU_K_HH_INIT));
    SWCU_CUDA(cudaStreamSynchronize(c->st));
    return SWCU_OK;
}

int swcu_envoke_kernel(swcu_ctx *c, int kernel_id, double tau)
{
    if (!c) { set_error("null ctx"); return SWCU_ERR_ARG; }
    if (c->p.mode != SWCU_MODE_REFERENCE) { set_error("per-kernel envokes need SWCU_MODE_REFERENCE"); return SWCU_ERR_STATE; }
    Use use(c->device);
    return envoke_kernel(c, kernel_id, tau);
}

int swcu_envoke_sync(swcu_ctx *c, int kernel_id)
{
    if (!c) { set_error("null ctx"); return SWCU_ERR_ARG; }
    if (c->p.mode != SWCU_MODE_REFERENCE) { set_error("per-kernel envokes need SWCU_MODE_REFERENCE"); return SWCU_ERR_STATE; }
    Use use(c->device);
    return envoke_sync(c, kernel_id);
}

int swcu_step(swcu_ctx *c, double tau, int nsteps)
{
    if (!c || nsteps < 0) { set_error("bad argument"); return SWCU_ERR_ARG; }
    if (c->nlinks) { set_error("a linked block steps with its neighbours: use swcu_step_group"); return SWCU_ERR_STATE; }
    Use use(c->device);
    for (int i = 0; i < nsteps; ++i) {
        RC(c->p.mode == SWCU_MODE_FUSED ? step_fused(c, tau) : step_reference(c, tau));
        c->steps_done++;
    }
    return SWCU_OK;
}

int swcu_profile_steps(swcu_ctx *c, double tau, int nsteps,
                       float *prep_ms, long *prep_launches, float *update_ms, long *update_launches)
{
    if (!c || nsteps < 0 || !prep_ms || !prep_launches || !update_ms || !update_launches) { set_error("bad argument"); return SWCU_ERR_ARG; }
    if (c->p.mode != SWCU_MODE_FUSED) { set_error("profile_steps needs FUSED mode"); return SWCU_ERR_STATE; }
    if (c->nlinks) { set_error("profile_steps on a linked block"); return SWCU_ERR_STATE; }
    Use use(c->device);
    c->prof = true;
    int rc = SWCU_OK;
    for (int i = 0; i < nsteps && !rc; ++i) { rc = step_fused(c, tau); c->steps_done++; }
    c->prof = false;
    cudaError_t e = cudaStreamSynchronize(c->st);
    float sum[2] = {0.f, 0.f};
    long cnt[2] = {0, 0};
    for (size_t k = 0; k < c->prof_kind.size() && 2 * k + 1 < c->prof_ev.size(); ++k) {
        float ms = 0.f;
        if (e == cudaSuccess && cudaEventElapsedTime(&ms, c->prof_ev[2 * k], c->prof_ev[2 * k + 1]) == cudaSuccess) {
            sum[c->prof_kind[k]] += ms;
            cnt[c->prof_kind[k]]++;
        }
    }
    for (cudaEvent_t ev : c->prof_ev) cudaEventDestroy(ev);
    c->prof_ev.clear(); c->prof_kind.clear();
    *prep_ms = sum[0]; *prep_launches = cnt[0]; *update_ms = sum[1]; *update_launches = cnt[1];
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "profile_steps");
    return SWCU_OK;
}

int swcu_synchronize(swcu_ctx *c, long *bad_cells)
{
    if (!c) { set_error("null ctx"); return SWCU_ERR_ARG; }
    Use use(c->device);
    if ((c->peer[0].on || c->peer[1].on) && c->steps_done > 0)   // the neighbours' rows of the last step have landed
        RC(peer_wait(c, c->st, 0, (unsigned long long)c->steps_done));
    SWCU_CUDA(cudaMemcpyAsync(c->bad_host, c->bad_dev, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    SWCU_CUDA(cudaMemsetAsync(c->bad_dev, 0, sizeof(int), c->st));
    SWCU_CUDA(cudaStreamSynchronize(c->st));
    SWCU_CUDA(cudaStreamSynchronize(c->comm_st));
    SWCU_CUDA(cudaStreamSynchronize(c->bnd_st));
    if (c->push_count && (c->peer[0].on || c->peer[1].on)) {   // did a strip warp give up waiting for a neighbour?
        int to = 0;
        SWCU_CUDA(cudaMemcpy(&to, c->push_count + 5, sizeof(int), cudaMemcpyDeviceToHost));
        if (to) {
            SWCU_CUDA(cudaMemset(c->push_count + 5, 0, sizeof(int)));
            set_error("a neighbouring block never signalled its halo rows / free buffers (peer-memory path timed out)");
            return SWCU_ERR_STATE;
        }
    }
    const long bad = *c->bad_host;
    if (bad_cells) *bad_cells = bad;
    if (bad) { set_error("check_ssh_err: %ld sea cells with |ssh| >= 1e4 or NaN", bad); return SWCU_ERR_BLOWUP; }
    return SWCU_OK;
}

int swcu_timer_start(swcu_ctx *c)
{
    if (!c) return SWCU_ERR_ARG;
    Use use(c->device);
    SWCU_CUDA(cudaEventRecord(c->t0, c->st));
    return SWCU_OK;
}
int swcu_timer_stop(swcu_ctx *c, float *ms)
{
    if (!c || !ms) return SWCU_ERR_ARG;
    Use use(c->device);
    SWCU_CUDA(cudaEventRecord(c->t1, c->st));
    SWCU_CUDA(cudaEventSynchronize(c->t1));
    SWCU_CUDA(cudaEventElapsedTime(ms, c->t0, c->t1));
    return SWCU_OK;
}
int swcu_selftest_mdiv(long n, unsigned long long seed, long *mismatches)
{
    if (n < 1 || !mismatches) { set_error("bad argument"); return SWCU_ERR_ARG; }
    unsigned long long *bad = nullptr, host = 0;
    SWCU_CUDA(cudaMalloc((void **)&bad, sizeof(*bad)));
    SWCU_CUDA(cudaMemset(bad, 0, sizeof(*bad)));
    int rc = launch_selftest_mdiv(n, seed, bad, nullptr);
    if (!rc) {
        cudaError_t e = cudaMemcpy(&host, bad, sizeof(host), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "selftest copy");
    }
    cudaFree(bad);
    *mismatches = (long)host;
    return rc;
}

long swcu_launch_count(const swcu_ctx *c) { return c ? c->launches : 0; }
long swcu_device_bytes(const swcu_ctx *c) { return c ? c->bytes : 0; }
void *swcu_stream(swcu_ctx *c) { return c ? (void *)c->st : nullptr; }

int swcu_comm_unique_id(void *id128)
{
    if (!id128) return SWCU_ERR_ARG;
    RC(nccl_load());
    ncclUniqueId id;
    SWCU_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return SWCU_OK;
}

int swcu_comm_init(swcu_ctx *c, int nranks, int rank, const void *id128)
{
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad argument"); return SWCU_ERR_ARG; }
    if (c->comm) { set_error("communicator already attached"); return SWCU_ERR_STATE; }
    if (nranks == 1) return SWCU_OK;
    if (c->nlinks || c->peer[0].on || c->peer[1].on) { set_error("a block exchanges halos over ONE of: communicator, in-process links, peer memory"); return SWCU_ERR_STATE; }
    RC(nccl_load());
    Use use(c->device);
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    SWCU_NCCL(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
    c->nranks = nranks; c->rank = rank;
    return SWCU_OK;
}

int swcu_comm_destroy(swcu_ctx *c)
{
    if (!c || !c->comm) return SWCU_OK;
    Use use(c->device);
    cudaDeviceSynchronize();
    SWCU_NCCL(g_nccl.CommDestroy(c->comm));
    c->comm = nullptr; c->nranks = 1; c->rank = 0;
    return SWCU_OK;
}

int swcu_link(swcu_ctx *a, swcu_ctx *b)
{
    if (!a || !b || a == b) { set_error("bad argument"); return SWCU_ERR_ARG; }
    if (a->comm || b->comm) { set_error("a block has either in-process links or a communicator"); return SWCU_ERR_STATE; }
    if (!same_params(a->p, b->p)) { set_error("linked blocks need identical parameters"); return SWCU_ERR_ARG; }
    auto rel = [](int a0, int a1, int b0, int b1, int *out) {
        if (b0 == a1 + 1) { *out = 1; return true; }
        if (b1 + 1 == a0) { *out = -1; return true; }
        if (b0 == a0 && b1 == a1) { *out = 0; return true; }
        return false;
    };
    int dx = 0, dy = 0;
    if (!rel(a->d.nx_start, a->d.nx_end, b->d.nx_start, b->d.nx_end, &dx) ||
        !rel(a->d.ny_start, a->d.ny_end, b->d.ny_start, b->d.ny_end, &dy) || (dx == 0 && dy == 0)) {
        set_error("blocks [%d:%d]x[%d:%d] and [%d:%d]x[%d:%d] are not neighbours of one block grid", a->d.nx_start,
                  a->d.nx_end, a->d.ny_start, a->d.ny_end, b->d.nx_start, b->d.nx_end, b->d.ny_start, b->d.ny_end);
        return SWCU_ERR_ARG;
    }
    for (const swcu_ctx *c : {a, b})
        if (c->d.nx_end - c->d.nx_start + 1 < 2 || c->d.ny_end - c->d.ny_start + 1 < 2) {
            set_error("a linked block needs at least 2 x 2 interior cells (halo width 2)");
            return SWCU_ERR_ARG;
        }
    int k = -1;
    for (int i = 0; i < 8; ++i) if (kDirX[i] == dx && kDirY[i] == dy) k = i;
    if (a->nbr[k] || b->nbr[k ^ 1]) { set_error("that side is already linked"); return SWCU_ERR_STATE; }
    if (a->device != b->device) {  // direct peer copies where the topology allows; staged otherwise
        for (int pass = 0; pass < 2; ++pass) {
            const int from = pass ? b->device : a->device, to = pass ? a->device : b->device;
            int can = 0;
            SWCU_CUDA(cudaDeviceCanAccessPeer(&can, from, to));
            if (!can) continue;
            Use use(from);
            cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
    }
    a->nbr[k] = b; b->nbr[k ^ 1] = a;
    a->nlinks++; b->nlinks++;
    return SWCU_OK;
}

int swcu_unlink(swcu_ctx *c)
{
    if (!c) return SWCU_OK;
    for (int k = 0; k < 8; ++k) {
        swcu_ctx *p = c->nbr[k];
        if (!p) continue;
        { Use use(p->device); cudaStreamSynchronize(p->st); }
        p->nbr[k ^ 1] = nullptr; p->nlinks--;
        c->nbr[k] = nullptr; c->nlinks--;
    }
    return SWCU_OK;
}

int swcu_step_group(swcu_ctx *const *cs, int n, double tau, int nsteps)
{
    if (!cs || n < 1 || nsteps < 0) { set_error("bad argument"); return SWCU_ERR_ARG; }
    for (int i = 0; i < n; ++i) {
        if (!cs[i]) { set_error("null block in the group"); return SWCU_ERR_ARG; }
        if (cs[i]->comm) { set_error("a block with a communicator steps with swcu_step"); return SWCU_ERR_STATE; }
        if (!same_params(cs[i]->p, cs[0]->p) || ntracers(cs[i]) != ntracers(cs[0])) { set_error("blocks of a group need identical parameters"); return SWCU_ERR_ARG; }
        for (int j = 0; j < i; ++j) if (cs[j] == cs[i]) { set_error("block listed twice"); return SWCU_ERR_ARG; }
        for (int k = 0; k < 8; ++k) {
            if (!cs[i]->nbr[k]) continue;
            bool found = false;
            for (int j = 0; j < n; ++j) found = found || cs[j] == cs[i]->nbr[k];
            if (!found) { set_error("block %d is linked to a block outside the group", i); return SWCU_ERR_ARG; }
        }
    }
    const swcu_params &p = cs[0]->p;
    // one (kernel, sync) pair of the reference's algorithm layer over all blocks of the group
    auto envoke = [&](int kid) -> int {
        for (int i = 0; i < n; ++i) { Use use(cs[i]->device); RC(envoke_kernel(cs[i], kid, tau)); }
        int f[3];
        const int nf = sync_list(kid, f);
        if (nf < 0) return SWCU_ERR_ARG;
        return group_pull(cs, n, nf, 1, [&f](const swcu_ctx *x, int a) { return x->f8[f[a]]; });
    };
    for (int s = 0; s < nsteps; ++s) {
        if (p.mode == SWCU_MODE_FUSED) {
            for (int i = 0; i < n; ++i) { Use use(cs[i]->device); RC(fused_main(cs[i], tau)); }
            RC(group_pull(cs, n, 6, 2, [](const swcu_ctx *x, int a) { return x->alt[a]; }));
            for (int k = 0; k < ntracers(cs[0]); ++k) {
                for (int i = 0; i < n; ++i) { Use use(cs[i]->device); tracer_bind(cs[i], k); RC(fused_tracer(cs[i])); }
                RC(group_pull(cs, n, 2, 2, [](const swcu_ctx *x, int a) { return x->alt_ff[a]; }));
                for (int i = 0; i < n; ++i) tracer_swap(cs[i]);
            }
            for (int i = 0; i < n; ++i) { tracer_bind(cs[i], cs[i]->tr_sel); fused_swap(cs[i]); }
        } else {  // control/shallow_water/shallow_water.f90:22-94, then control/tracer.f90:44-61
            RC(envoke(SWCU_K_SW_UPDATE_SSH));
            if (p.full_free_surface > 0) RC(envoke(SWCU_K_HH_UPDATE));
            if (p.trans_terms > 0) { RC(envoke(SWCU_K_UV_TRANS_VORT)); RC(envoke(SWCU_K_UV_TRANS)); }
            if (p.ksw_lat > 0) { RC(envoke(SWCU_K_STRESS_COMPONENTS)); RC(envoke(SWCU_K_UV_DIFF2)); }
            RC(envoke(SWCU_K_SW_UPDATE_UV));
            RC(envoke(SWCU_K_SW_NEXT_STEP));
            if (p.full_free_surface > 0) { RC(envoke(SWCU_K_HH_SHIFT)); RC(envoke(SWCU_K_HH_INIT)); }
            RC(envoke(SWCU_K_CHECK_SSH_ERR));
            for (int k = 0; k < ntracers(cs[0]); ++k) {
                for (int i = 0; i < n; ++i) tracer_bind(cs[i], k);
                RC(envoke(SWCU_K_TRAN_DIFF_FLUXES));
                RC(envoke(SWCU_K_TRAN_DIFF_TRACER));
                RC(envoke(SWCU_K_TRACER_NEXT_STEP));
            }
            for (int i = 0; i < n; ++i) tracer_bind(cs[i], cs[i]->tr_sel);
        }
        for (int i = 0; i < n; ++i) cs[i]->steps_done++;
    }
    return SWCU_OK;
}

int swcu_init_grid(swcu_ctx *c, const swh_basin *b, const int *mask)
{
    if (!c || !b) { set_error("null argument"); return SWCU_ERR_ARG; }
    const swcu_dims &d = c->d;
    if (d.bnd_x1 < 1 || d.bnd_x2 > b->nx || d.bnd_y1 < 1 || d.bnd_y2 > b->ny) {
        set_error("block array [%d:%d]x[%d:%d] leaves the %d x %d basin", d.bnd_x1, d.bnd_x2, d.bnd_y1, d.bnd_y2, b->nx, b->ny);
        return SWCU_ERR_ARG;
    }
    if (b->curve_grid != 0 && b->curve_grid != 1) { set_error("curve_grid must be 0 or 1"); return SWCU_ERR_ARG; }
    Use use(c->device);
    const bool fused = c->p.mode == SWCU_MODE_FUSED;
    const int w = c->w, h = c->h;
    int rc = SWCU_OK;
    std::vector<void *> scratch;
    auto tmp = [&](size_t bytes) -> void * {
        void *p = nullptr;
        if (rc) return p;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e != cudaSuccess) { rc = cuda_fail(e, "scratch for swcu_init_grid"); return nullptr; }
        scratch.push_back(p);
        return p;
    };
    // ---- masks
    int *land = nullptr;
    if (mask) {
        land = (int *)tmp((size_t)w * h * sizeof(int));
        if (!rc) {
            cudaError_t e = cudaMemcpy2DAsync(land, (size_t)w * sizeof(int),
                                              mask + (size_t)(d.bnd_y1 - 1) * b->nx + (d.bnd_x1 - 1), (size_t)b->nx * sizeof(int),
                                              (size_t)w * sizeof(int), (size_t)h, cudaMemcpyHostToDevice, c->st);
            if (e != cudaSuccess) rc = cuda_fail(e, "mask window upload");
        }
    }
    unsigned char *lu = (unsigned char *)tmp((size_t)w * h);
    if (!rc) rc = launch_init_lu(c->g, w, h, b->nx, b->ny, land, lu, c->st);
    if (!rc) {
        if (fused) rc = launch_init_masks(c->g, w, h, lu, c->mask, nullptr, c->st);
        else {
            float *const f[7] = {F4(c, SWCU_F_LU), F4(c, SWCU_F_LUU), F4(c, SWCU_F_LUH), F4(c, SWCU_F_LCU),
                                 F4(c, SWCU_F_LCV), F4(c, SWCU_F_LLU), F4(c, SWCU_F_LLV)};
            rc = launch_init_masks(c->g, w, h, lu, nullptr, f, c->st);
        }
        c->masks_dirty = true;
        c->launches += 2;
    }
    // ---- metrics and Coriolis parameter
    const int ids[9] = {SWCU_F_DX, SWCU_F_DY, SWCU_F_DXT, SWCU_F_DYT, SWCU_F_DXH, SWCU_F_DYH, SWCU_F_DXB, SWCU_F_DYB, SWCU_F_RLH_S};
    const bool row_constant = b->curve_grid == 0 || b->rotation_on_lat == 0.0;
    if (!rc && row_constant) {
        // one array column outside [2..nx-1] (m = 1) and one inside (m = 2), evaluated by swh_metrics itself
        const swcu_dims strip = {2, 2, d.ny_start, d.ny_end, 1, 2, d.bnd_y1, d.bnd_y2};
        std::vector<float> col[9];
        float *ptr[9];
        for (int k = 0; k < 9; ++k) { col[k].assign((size_t)2 * h, 0.0f); ptr[k] = col[k].data(); }
        rc = swh_metrics(b, &strip, ptr[0], ptr[1], ptr[2], ptr[3], ptr[4], ptr[5], ptr[6], ptr[7], ptr[8]);
        if (rc) set_error("swh_metrics rejected the basin");
        std::vector<float> prof((size_t)9 * 2 * h);
        for (int k = 0; k < 9; ++k)
            for (int side = 0; side < 2; ++side)
                for (int j = 0; j < h; ++j) prof[((size_t)k * 2 + side) * h + j] = col[k][(size_t)j * 2 + side];
        float *prof_dev = (float *)tmp(prof.size() * sizeof(float));
        if (!rc) {
            cudaError_t e = cudaMemcpyAsync(prof_dev, prof.data(), prof.size() * sizeof(float), cudaMemcpyHostToDevice, c->st);
            if (e != cudaSuccess) rc = cuda_fail(e, "metric profile upload");
        }
        if (!rc) {
            float *dst[9];
            for (int k = 0; k < 9; ++k) dst[k] = F4(c, ids[k]);
            rc = launch_expand_rows(c->g, w, h, b->nx, dst, prof_dev, c->st);
            c->launches++;
        }
        if (!rc) {   // the profile vector must outlive the asynchronous copy
            cudaError_t e = cudaStreamSynchronize(c->st);
            if (e != cudaSuccess) rc = cuda_fail(e, "swcu_init_grid");
        }
    } else if (!rc) {
        // rotated pole: rlh_s varies along x; build the 2-D arrays on the host like the reference and upload them
        std::vector<float> arr[9];
        float *ptr[9];
        for (int k = 0; k < 9; ++k) { arr[k].assign((size_t)w * h, 0.0f); ptr[k] = arr[k].data(); }
        rc = swh_metrics(b, &d, ptr[0], ptr[1], ptr[2], ptr[3], ptr[4], ptr[5], ptr[6], ptr[7], ptr[8]);
        for (int k = 0; k < 9 && !rc; ++k) rc = upload_impl(c, ids[k], ptr[k], false);
        if (!rc) {
            cudaError_t e = cudaStreamSynchronize(c->st);
            if (e != cudaSuccess) rc = cuda_fail(e, "swcu_init_grid");
        }
    }
    c->metrics_dirty = true;
    cudaStreamSynchronize(c->st);
    for (void *p : scratch) cudaFree(p);
    return rc;
}

int swcu_fill(swcu_ctx *c, int field, double value)
{
    if (!c) { set_error("null ctx"); return SWCU_ERR_ARG; }
    Use use(c->device);
    const bool fused = c->p.mode == SWCU_MODE_FUSED;
    if (is_f8(field)) {
        if (is_tracer_field(field) && !c->p.use_tracers) { set_error("tracer field without use_tracers"); return SWCU_ERR_STATE; }
        if (fused && (field == SWCU_F_RHSX || field == SWCU_F_RHSY) && !c->has_rhs) {
            c->has_rhs = true;
            RC(alloc8(c, SWCU_F_RHSX)); RC(alloc8(c, SWCU_F_RHSY));
        }
        if (fused && !fused_keeps8(c, field) && !(c->p.use_tracers && (field == SWCU_F_FF1 || field == SWCU_F_FF1P)))
            return SWCU_OK;  // derived or n+1 array: recomputed on the device
        RC(launch_fill8(c->w, c->h, c->pitch, c->f8[field], value, c->st));
        if (fused && (state_slot(field) >= 0 || is_tracer_field(field))) c->alt_dirty = true;
        c->launches++;
        return SWCU_OK;
    }
    if (is_f4(field)) {
        if (mask_bit(field)) { set_error("masks are set by swcu_init_grid or swcu_upload"); return SWCU_ERR_ARG; }
        if (fused && field == SWCU_F_R_DISS && !c->has_rdiss) { c->has_rdiss = true; RC(alloc4(c, SWCU_F_R_DISS)); }
        if (fused && !fused_keeps4(c, field)) return SWCU_OK;
        if (field >= SWCU_F_DX && field <= SWCU_F_RLH_S) c->metrics_dirty = true;
        RC(launch_fill4(c->w, c->h, c->pitch, F4(c, field), (float)value, c->st));
        c->launches++;
        return SWCU_OK;
    }
    set_error("unknown field id %d", field);
    return SWCU_ERR_ARG;
}

int swcu_copy_field(swcu_ctx *c, int dst_field, int src_field)
{
    if (!c) { set_error("null ctx"); return SWCU_ERR_ARG; }
    if (!is_f8(dst_field) || !is_f8(src_field)) { set_error("copy_field works on real(8) fields"); return SWCU_ERR_ARG; }
    Use use(c->device);
    const bool fused = c->p.mode == SWCU_MODE_FUSED;
    const bool tr_dst = c->p.use_tracers && (dst_field == SWCU_F_FF1 || dst_field == SWCU_F_FF1P);
    if (fused && !fused_keeps8(c, dst_field) && !tr_dst) return SWCU_OK;  // e.g. sshn: not resident, derived
    const double *src = c->f8[src_field];
    double *dst = c->f8[dst_field];
    if (!src || !dst) { set_error("field not resident"); return SWCU_ERR_STATE; }
    SWCU_CUDA(cudaMemcpyAsync(dst, src, c->plane * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
    if (fused && (state_slot(dst_field) >= 0 || tr_dst)) c->alt_dirty = true;
    return SWCU_OK;
}

namespace {
struct PeerBlob {
    unsigned magic;
    int pid, device;
    int by1, ny_start, ny_end, pitch, w, tracers, cur_set;
    long steps_done;
    cudaIpcMemHandle_t mem[2][8];
    cudaIpcMemHandle_t flags;
};
static_assert(sizeof(PeerBlob) <= SWCU_PEER_BLOB_BYTES, "blob size");
const unsigned kPeerMagic = 0x53574355u;
}  // namespace

int swcu_peer_export(swcu_ctx *c, void *blob)
{
    if (!c || !blob) { set_error("null argument"); return SWCU_ERR_ARG; }
    if (c->p.mode != SWCU_MODE_FUSED) { set_error("the peer-memory halo path needs SWCU_MODE_FUSED"); return SWCU_ERR_STATE; }
    if (ntracers(c) > 1) { set_error("the peer-memory halo path carries one tracer: use a communicator"); return SWCU_ERR_STATE; }
    Use use(c->device);
    PeerBlob b;
    memset(&b, 0, sizeof(b));
    b.magic = kPeerMagic; b.pid = (int)getpid(); b.device = c->device;
    b.by1 = c->d.bnd_y1; b.ny_start = c->d.ny_start; b.ny_end = c->d.ny_end; b.pitch = c->pitch; b.w = c->w;
    b.tracers = c->p.use_tracers ? 1 : 0; b.cur_set = c->cur_set; b.steps_done = c->steps_done;
    const int n = c->p.use_tracers ? 8 : 6;
    for (int s = 0; s < 2; ++s)
        for (int k = 0; k < n; ++k) SWCU_CUDA(cudaIpcGetMemHandle(&b.mem[s][k], c->set_ptr[s][k]));
    SWCU_CUDA(cudaIpcGetMemHandle(&b.flags, c->flags));
    memset(blob, 0, SWCU_PEER_BLOB_BYTES);
    memcpy(blob, &b, sizeof(b));
    return SWCU_OK;
}

int swcu_peer_detach(swcu_ctx *c)
{
    if (!c) return SWCU_OK;
    Use use(c->device);
    cudaDeviceSynchronize();
    for (auto &pl : c->peer) {
        for (void *p : pl.opened) cudaIpcCloseMemHandle(p);
        pl = swcu_ctx::PeerLink();
    }
    if (c->flags) SWCU_CUDA(cudaMemset(c->flags, 0, 8 * sizeof(unsigned long long)));
    return SWCU_OK;
}

int swcu_peer_attach(swcu_ctx *c, int side, const void *blob)
{
    if (!c || !blob || (side != 0 && side != 1)) { set_error("bad argument"); return SWCU_ERR_ARG; }
    if (c->p.mode != SWCU_MODE_FUSED) { set_error("the peer-memory halo path needs SWCU_MODE_FUSED"); return SWCU_ERR_STATE; }
    if (c->comm || c->nlinks) { set_error("a block exchanges halos over ONE of: communicator, in-process links, peer memory"); return SWCU_ERR_STATE; }
    if (c->peer[side].on) { set_error("that side is already attached"); return SWCU_ERR_STATE; }
    PeerBlob b;
    memcpy(&b, blob, sizeof(b));
    if (b.magic != kPeerMagic) { set_error("not a swcu_peer_export blob"); return SWCU_ERR_ARG; }
    if (b.pid == (int)getpid()) { set_error("blocks of one process are tied with swcu_link"); return SWCU_ERR_ARG; }
    const bool adjacent = side == 0 ? b.ny_end + 1 == c->d.ny_start : b.ny_start == c->d.ny_end + 1;
    if (!adjacent || b.pitch != c->pitch || b.w != c->w) { set_error("the exported block is not my y-neighbour on that side"); return SWCU_ERR_ARG; }
    if (b.tracers != (c->p.use_tracers ? 1 : 0) || b.cur_set != c->cur_set || b.steps_done != c->steps_done) {
        set_error("neighbours must agree on use_tracers and on the number of steps taken");
        return SWCU_ERR_STATE;
    }
    RC(wait_value_load());
    Use use(c->device);
    swcu_ctx::PeerLink &pl = c->peer[side];
    const int n = c->p.use_tracers ? 8 : 6;
    for (int s = 0; s < 2; ++s)
        for (int k = 0; k < n; ++k) {
            void *p = nullptr;
            SWCU_CUDA(cudaIpcOpenMemHandle(&p, b.mem[s][k], cudaIpcMemLazyEnablePeerAccess));
            pl.opened.push_back(p);
            pl.set[s][k] = (double *)p;
        }
    void *f = nullptr;
    SWCU_CUDA(cudaIpcOpenMemHandle(&f, b.flags, cudaIpcMemLazyEnablePeerAccess));
    pl.opened.push_back(f);
    pl.flags = (unsigned long long *)f;
    pl.by1 = b.by1;
    pl.on = true;
    {   // the halo rows towards this neighbour are valid as of the steps taken so far
        const unsigned long long v = (unsigned long long)c->steps_done;
        SWCU_CUDA(cudaMemcpy(c->flags + side, &v, sizeof(v), cudaMemcpyHostToDevice));
    }
    return SWCU_OK;
}

int swcu_march_band_rows(int n0, int n1, int nbands, int late_bands, int late_cut, int band, int *first, int *last)
{
    if (!first || !last || nbands < 1 || band < 0 || band >= nbands || late_bands < 0 || late_bands > nbands || late_cut < 0) {
        set_error("bad band arguments");
        return SWCU_ERR_ARG;
    }
    MarchPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.n0 = n0; pl.n1 = n1; pl.nbands = nbands; pl.late_hi = late_bands; pl.late_cut = late_cut;
    march_band_rows(pl, band, first, last);
    return SWCU_OK;
}

int swcu_halo_plan(const swcu_dims *d, int nrows, int side, int *send_row, int *recv_row)
{
    if (!d || !send_row || !recv_row || nrows < 1 || nrows > 2 || (side != 0 && side != 1)) {
        set_error("bad halo plan arguments");
        return SWCU_ERR_ARG;
    }
    const int r_first = d->ny_start - d->bnd_y1, r_last = d->ny_end - d->bnd_y1;
    if (r_last - r_first + 1 < nrows) { set_error("block has fewer interior rows than the halo width"); return SWCU_ERR_ARG; }
    if (side == 0) { *send_row = r_first; *recv_row = r_first - nrows; }
    else { *send_row = r_last - nrows + 1; *recv_row = r_last + 1; }
    return SWCU_OK;
}

int swcu_widen_halos(swcu_ctx *c)
{
    if (!c) { set_error("null ctx"); return SWCU_ERR_ARG; }
    if (c->p.mode != SWCU_MODE_FUSED) return SWCU_OK;   // REFERENCE mode keeps the reference's width-1 syncs
    if (c->peer[0].on || c->peer[1].on) {
        set_error("swcu_widen_halos runs over a communicator or in-process links: call it before swcu_peer_attach");
        return SWCU_ERR_STATE;
    }
    if (!c->comm && !c->nlinks) return SWCU_OK;          // no neighbours: nothing to widen
    Use use(c->device);
    std::vector<double *> p8;
    std::vector<float *> p4;
    auto other_tracers = [](swcu_ctx *x, std::vector<double *> &v) {   // tracers 1 .. n-1 (tracer 0 is bound here)
        for (size_t k = 1; k < x->tr.size(); ++k) { v.push_back(x->tr[k].ff); v.push_back(x->tr[k].ffp); }
    };
    tracer_bind(c, 0);
    for (int k = 0; k < 8; ++k) if (c->nbr[k]) tracer_bind(c->nbr[k], 0);
    for (int f = 0; f < SWCU_NF8; ++f) {
        const bool tracer = c->p.use_tracers && (f == SWCU_F_FF1 || f == SWCU_F_FF1P);
        if (c->f8[f] && (state_slot(f) >= 0 || f == SWCU_F_HHQ_REST || f == SWCU_F_MU || tracer ||
                         ((f == SWCU_F_RHSX || f == SWCU_F_RHSY) && c->has_rhs)))
            p8.push_back(c->f8[f]);
    }
    other_tracers(c, p8);
    for (int f = SWCU_F_DX; f <= SWCU_F_RLH_S; ++f) if (F4(c, f)) p4.push_back(F4(c, f));
    if (c->has_rdiss && F4(c, SWCU_F_R_DISS)) p4.push_back(F4(c, SWCU_F_R_DISS));
    if (c->nlinks) {
        // position of each plane in the neighbour's lists is the same (same parameters, same residency)
        int rc = SWCU_OK;
        for (int k = 0; k < 8 && !rc; ++k) {
            swcu_ctx *n = c->nbr[k];
            if (!n) continue;
            { Use un(n->device); SWCU_CUDA(cudaStreamSynchronize(n->st)); }
            std::vector<double *> q8;
            std::vector<float *> q4;
            for (int f = 0; f < SWCU_NF8; ++f) {
                const bool tracer = n->p.use_tracers && (f == SWCU_F_FF1 || f == SWCU_F_FF1P);
                if (n->f8[f] && (state_slot(f) >= 0 || f == SWCU_F_HHQ_REST || f == SWCU_F_MU || tracer ||
                                 ((f == SWCU_F_RHSX || f == SWCU_F_RHSY) && n->has_rhs)))
                    q8.push_back(n->f8[f]);
            }
            other_tracers(n, q8);
            for (int f = SWCU_F_DX; f <= SWCU_F_RLH_S; ++f) if (F4(n, f)) q4.push_back(F4(n, f));
            if (n->has_rdiss && F4(n, SWCU_F_R_DISS)) q4.push_back(F4(n, SWCU_F_R_DISS));
            if (q8.size() != p8.size() || q4.size() != p4.size()) { set_error("linked blocks hold different fields"); return SWCU_ERR_STATE; }
            for (size_t i = 0; i < p8.size() && !rc; ++i) rc = pull_halo<double>(c, k, p8[i], q8[i], 2, c->st);
            for (size_t i = 0; i < p4.size() && !rc; ++i) rc = pull_halo<float>(c, k, p4[i], q4[i], 2, c->st);
            if (!rc) rc = pull_halo<unsigned char>(c, k, c->mask, n->mask, 2, c->st);
        }
        if (rc) return rc;
    } else {
        SWCU_NCCL(g_nccl.GroupStart());
        int rc = SWCU_OK;
        for (size_t i = 0; i < p8.size() && !rc; ++i) rc = exchange_rows(c, p8[i], 2, c->st);
        for (size_t i = 0; i < p4.size() && !rc; ++i) rc = exchange_rows(c, p4[i], 2, c->st);
        if (!rc) rc = exchange_rows(c, c->mask, 2, c->st);
        ncclResult_t r = g_nccl.GroupEnd();
        if (rc) return rc;
        if (r != ncclSuccess) return nccl_fail(r, "ncclGroupEnd");
    }
    SWCU_CUDA(cudaStreamSynchronize(c->st));
    tracer_bind(c, c->tr_sel);
    for (int k = 0; k < 8; ++k) if (c->nbr[k]) tracer_bind(c->nbr[k], c->nbr[k]->tr_sel);
    c->alt_dirty = true; c->metrics_dirty = true; c->masks_dirty = true;
    return SWCU_OK;
}

int swcu_halo_exchange(swcu_ctx *c, int field)
{
    if (!c) return SWCU_ERR_ARG;
    if (c->nlinks) {
        Use use(c->device);
        const int hw = c->p.mode == SWCU_MODE_FUSED ? 2 : 1;
        if (is_f8(field)) {
            if (c->p.mode == SWCU_MODE_FUSED && state_slot(field) >= 0) c->alt_dirty = true;
            return pull_blocking<double>(c, hw, 1, [field](const swcu_ctx *x, int) { return x->f8[field]; });
        }
        if (is_f4(field)) return pull_blocking<float>(c, hw, 1, [field](const swcu_ctx *x, int) { return x->f4[field - 100]; });
        set_error("unknown field id %d", field);
        return SWCU_ERR_ARG;
    }
    if (c->peer[0].on || c->peer[1].on) {
        set_error("swcu_halo_exchange: this block exchanges halos over peer memory, which carries the prognostic arrays "
                  "of the step only; exchange set-up fields over a communicator (or swcu_widen_halos) before swcu_peer_attach");
        return SWCU_ERR_STATE;
    }
    if (!c->comm) return SWCU_OK;
    Use use(c->device);
    const bool fused = c->p.mode == SWCU_MODE_FUSED;
    SWCU_NCCL(g_nccl.GroupStart());
    int rc = SWCU_OK;
    if (is_f8(field) && c->f8[field]) {
        rc = exchange_rows(c, c->f8[field], fused ? 2 : 1, c->st);
        if (fused && state_slot(field) >= 0) c->alt_dirty = true;
    } else if (is_f4(field) && F4(c, field)) rc = exchange_rows(c, F4(c, field), fused ? 2 : 1, c->st);
    else { set_error("field %d is not resident", field); rc = SWCU_ERR_STATE; }
    ncclResult_t r = g_nccl.GroupEnd();
    if (rc) return rc;
    if (r != ncclSuccess) return nccl_fail(r, "ncclGroupEnd");
    SWCU_CUDA(cudaStreamSynchronize(c->st));
    return SWCU_OK;
}

}  // extern "C"
