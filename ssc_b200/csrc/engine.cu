// C ABI + device engine of the B200 patch preconditioner (include/ssc_b200.h).
//
// Stage map (reference loop -> kernel), see DESIGN.md:
//   PCSetUp_PATCH  ssc/libssc.c:1201-1299  -> host patch builder (patch_builder.cpp), K1 extract, K2 factor, PoU weights
//   PCApply_PATCH  ssc/libssc.c:1301-1426  -> halo fill, fused gather/solve/scatter-add kernel per size class,
//                                             overlap sum, BC pass-through, optional weighting
// There is no CPU fallback: every compute entry point fails if no CUDA device is present.
#include "../../include/ssc_b200.h"
#include "common.h"
#include "kernels.cuh"
#include "patch_builder_device.cuh"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>
#include <dlfcn.h>
#include <array>
#include <climits>
#include <atomic>
#include <thread>

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

int ssc_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
extern "C" const char *SSCGetLastError(void) { return g_err; }

#define CUDA_OK(call)                                                                                          \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess)                                                                                 \
            return ssc_set_error(SSC_ERR_GPU, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
    } while (0)
#define CHK(call)                    \
    do {                             \
        int ierr_ = (call);          \
        if (ierr_) return ierr_;     \
    } while (0)

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen: a one-GPU run never needs the library
// ---------------------------------------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_t;
struct Nccl {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_t, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static Nccl g_nccl;
static int nccl_load()
{
    if (g_nccl.h) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !g_nccl.h; ++i) g_nccl.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!g_nccl.h) return ssc_set_error(SSC_ERR_LIB, "cannot load libnccl: %s", dlerror());
#define SYM(f, name)                                                                                 \
    *(void **)(&g_nccl.f) = dlsym(g_nccl.h, name);                                                   \
    if (!g_nccl.f) return ssc_set_error(SSC_ERR_LIB, "libnccl lacks %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return 0;
}
#define NCCL_OK(call)                                                                                     \
    do {                                                                                                  \
        int r_ = (call);                                                                                  \
        if (r_ != 0) return ssc_set_error(SSC_ERR_LIB, "NCCL error %s at %s:%d", g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)
static const int kNcclFloat64 = 8;   // ncclDouble

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct SizeClass {
    int n = 0;
    i64 count = 0, first = 0;     // position in size-sorted order
    i64 goff = 0;                 // offset of the class in d_gtol (entries)
    i64 moff = 0;                 // offset of the class in d_blocks (doubles)
    i64 stride = 0;               // doubles per block
    int ld = 0;                   // row stride inside a block (doubles)
    i64 poff = 0, pstride = 0;    // packed-symmetric storage: class offset and per-patch stride (doubles)
    unsigned pbytes = 0;          // bytes of one packed triangle moved by the bulk copy
    int tp = 0, ppb = 0;          // apply launch shape
    bool refine = false;          // -ssc_sub_solve: this class is applied with one step of iterative refinement against the kept operators
    double kappa = 0.0;           // largest ||B||_1 ||B^-1||_inf of the class (computed when -ssc_sub_solve is refine or auto)
};

struct SSCPatchPC_s {
    int device = 0;
    cudaStream_t stream = nullptr, up_stream = nullptr;
    cudaStream_t side_stream = nullptr, launch_stream = nullptr;   // the small size classes of an apply run beside the big one; launch_apply_class launches on launch_stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap_classes = true;
    int apply_stream_min_waves = 4;  // ... and only when every CTA gets at least this many rounds of patches (tests: 0)
    int apply_stream_ctas = 0;       // CTAs of the persistent kernel (0: fill the GPU; tests use a few to get many rounds on small cases)
    static constexpr int apply_stream_max_n = 96;     // largest patch the persistent apply kernel takes (n = 125: measured no better than the one-patch-per-CTA kernel)
    static constexpr bool apply_exact_groups = true;  // n < 32: a group has exactly n threads (measured against n rounded up to a multiple of 8)
    static constexpr int apply_group_threads = 128;   // threads of a CTA that carries several small patches (n <= 96): 96 / 128 / 192 / 256 measured in round 1
    bool apply_stream_small = true;  // n <= 96: persistent apply kernel with the gathers prefetched across patches
    int sm_count = 148;
    size_t total_mem = (size_t)180 << 30;
    // options (defaults: PCCreate_PATCH ssc/libssc.c:1697-1711)
    bool save_operators = true, partition_of_unity = false, multiplicative = false, print_patches = false;
    bool symmetrise_sweep = false, keep_operators = false, dim_set = false, codim_set = false;
    int sub_pc = SSC_SUB_LU;
    int sub_solve = 0;            // -ssc_sub_solve: 0 inverse (stored-inverse product), 1 refine (+ one refinement step, every class), 2 auto (refine where kappa > refine_kappa)
    int refine_steps = 1;         // -ssc_refine_steps: refinement steps of a refined class (each contracts the error by ||I - S B||)
    double refine_kappa = 1.0e3;  // -ssc_refine_kappa: threshold of auto (kappa * eps = 2e-13 there)
    unsigned long long *d_kappa = nullptr;
    int factor_ctas = 0;          // concurrent CTAs of the blocked factor kernel (0 = fill the SMs)
    int factor_panel = 16;        // panel width of the blocked factor kernel (16 or 32)
    int factor_variant = -1;      // -1 (default) by size: n <= 64 register tiles (1), 65..256 tensor-core Gauss-Jordan on the L2-resident block, several patches per SM (2),
                                  // larger blocked DMMA (3), beyond its shared-memory panels the in-place global-memory Gauss-Jordan (0); 0..3 force one where it applies
    int apply_variant = 1;        // 0: simple loop (also the fallback for patches too large for one thread per row), 1: pipelined LDG.64 (default)
    std::string sub_mat_type, sub_ksp_type = "preonly";
    BuildOptions bopt;
    // host inputs
    HostPlex plex;
    Discretisation disc;
    UserPatches user;
    PatchLists lists;
    bool lists_valid = false, direct_patches = false;
    i64 nlocal = 0, nowned = 0;
    std::vector<i64> globalBc;
    // star forest
    bool sf_set = false;
    std::vector<i64> selfLeaf, selfRoot, sendOff, sendRoot, recvOff, recvLeaf;
    std::vector<int> neigh;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    // NVLink peer exchange (CUDA IPC)
    bool p2p = false;
    PeerScatter h_ps;
    PeerScatter *d_ps = nullptr;
    PeerPush h_push;
    PeerFlags h_flags;
    int *d_ghostPeer = nullptr, *d_ghostRemote = nullptr, *d_sendPeer = nullptr, *d_remoteLeaf = nullptr, *d_iface = nullptr;
    i64 niface = 0;
    double *d_xg = nullptr, *d_yin = nullptr;       // ghost values of x (filled by neighbours), incoming contributions to owned dofs
    unsigned long long *d_flags = nullptr, epoch = 0;
    int *d_timeout = nullptr;
    std::vector<void *> ipc_opened;
    // operator (borrowed)
    i64 csr_nrows = 0, csr_nnz = 0;
    const i64 *csr_rowptr = nullptr;
    const int32_t *csr_colidx = nullptr;
    const double *csr_vals = nullptr;
    int csr_space = SSC_MEM_HOST;
    bool csr_set = false;
    bool op_released = false;        // the caller took the borrowed operator arrays back (SSCPatchReleaseOperator)
    // device state
    bool device_ready = false, setup_done = false;
    std::vector<SizeClass> classes;
    std::vector<i64> sorted2orig, orig2sorted;
    std::vector<i64> colour;                                        // per original patch (multiplicative mode)
    i64 ncolours = 0;
    std::vector<std::vector<std::pair<i64, i64>>> colour_ranges;    // [colour][class] -> sorted positions
    double *d_r = nullptr;
    bool deterministic = false;      // slot-and-sum instead of atomics
    bool device_construct = false;   // -ssc_device_patch_construction: build star patches on the GPU (N2); default: host builder
                                     // (measured no faster end to end: the lists have to come back to the host, profiles/README.md)
    bool lists_from_device = false;
    bool symmetric = false;          // packed lower triangle of the (symmetric) inverse only
    double *d_packed = nullptr;
    i64 packed_elems = 0;
    double *d_ypatch = nullptr;
    i64 *d_slotoff = nullptr;
    int *d_slots = nullptr;
    int *d_gtol = nullptr;
    double *d_blocks = nullptr, *d_blocks0 = nullptr;
    i64 block_elems = 0, sum_n = 0, sum_n2 = 0, max_n = 0;
    int *d_fail = nullptr, *d_anyfail = nullptr;
    i64 failed = 0;
    std::vector<i64> failed_list;
    int *d_bc = nullptr;
    i64 nbc = 0;
    double *d_w = nullptr, *d_mult = nullptr;
    double *d_xl = nullptr, *d_yl = nullptr;      // local (owned+ghost) vectors when the SF is not the identity
    double *d_xs = nullptr, *d_ys = nullptr;      // staging for host-space applies
    int *d_selfLeaf = nullptr, *d_selfRoot = nullptr, *d_sendRoot = nullptr, *d_recvLeaf = nullptr;
    double *d_sendbuf = nullptr, *d_recvbuf = nullptr;
    // device copies of the operator
    i64 *d_rowptr = nullptr;
    int *d_colidx = nullptr;
    double *d_vals = nullptr;
    bool csr_owned = false;
    bool csr_values_only = false;    // SSCPatchSetOperatorValues: the device copy of rowptr / colidx is current, only the values travel
    bool csr_dup = false;            // some row of the CSR has unsorted or repeated columns: the extract kernel accumulates instead of storing
    int *d_flag = nullptr;
    int extract_row_buffers = -1;    // -ssc_extract_row_buffers: warps of an extract CTA that get a shared-memory row buffer (-1: as many as fit; 0: hits go straight to global memory, the path of very large patches)
    i64 csr_dev_rows = 0, csr_dev_nnz = 0;   // shape of the owned device copy
    // element matrices (the reference's own source of the patch operators, ssc/libssc.c:1120-1156): borrowed like the CSR arrays
    bool elem_set = false, elem_owned = false, elem_lists_ready = false, elem_maps_ready = false;
    i64 elem_map_rows = 0, elem_dev_len = 0;
    static constexpr int asm_threads = 256;
    static constexpr int asm_ctas_per_sm = 16;        // blocks being accumulated at once per SM (their in-place working set competes for L2; sweep flat in round 1)
    const double *elem_K = nullptr;
    i64 elem_ncells = 0, elem_t = 0, elem_stride = 0;
    int elem_space = 0;
    double *d_elemK = nullptr;
    i64 *d_pcelloff = nullptr;       // cells of every patch, patches in size-sorted order
    int *d_pcells = nullptr;
    std::vector<int *> d_cnm;        // device copies of the cell-node maps
    ElemSpaces espaces;
    void *d_flush = nullptr;
    // milestones of the values upload: mark k is recorded on up_stream once the first mark_bytes[k] bytes of the CSR values are on their way,
    // so extract + factor of the patches whose rows lie below it start while the rest still travels (SSCPatchSetUp)
    static constexpr int kUploadMarks = 16;
    cudaEvent_t mark_ev[kUploadMarks] = {};
    size_t mark_bytes[kUploadMarks] = {};
    std::atomic<int> marks_recorded{0};
    cudaEvent_t pattern_ev = nullptr;                 // rowptr + colidx are on their way
    std::atomic<int> pattern_recorded{0};
    std::atomic<int> upload_finished{0};              // the uploader thread has returned (with or without an error)
    bool marks_active = false;                        // this set-up uploads host CSR values with milestones
    std::vector<int> patch_maxrow;                    // per sorted patch: the largest row it reads
    std::vector<i64> patch_need;                      // per sorted patch: CSR values it needs = rowptr[maxrow + 1]
    std::atomic<int> upload_pause{0};                 // the staged upload issues no new chunk while set: the dof lists (needed first) get the copy engine to themselves
    std::atomic<int> builder_busy{0};                 // the patch builder is running: only stage_threads staging threads work
    int stage_threads = 0;                            // host threads of the staged upload (chosen per set-up)
    int builder_threads_auto = 0;                     // patch-builder threads when ssc_builder_threads is not set and an upload runs beside it
    static constexpr int kStageRing = 4;
    void *stage[kStageRing] = {nullptr, nullptr, nullptr, nullptr};      // pinned staging ring for large pageable uploads
    cudaEvent_t stage_ev[kStageRing] = {nullptr, nullptr, nullptr, nullptr};
    bool stage_busy[kStageRing] = {false, false, false, false};
    // pipelined host-space apply: H2D of x, patch kernels and D2H of y overlap slab by slab
    int pipe_slabs = 12;
    bool pipe_force = false;
    bool pipe_taper = true;      // slabs grow geometrically from both ends of the patch list
    bool pipe_ready = false, pipe_tried = false;
    std::vector<std::vector<std::pair<i64, i64>>> pipe_ranges;   // [slab][class] -> sorted positions [q0, q1)
    std::vector<i64> pipe_in, pipe_out, pipe_bc;                 // x piece bounds, y piece bounds, bc list offsets per y piece
    cudaStream_t s_in = nullptr, s_out = nullptr;
    // the same with neighbour ranks (peer exchange): pieces of x / y over the owned dofs, slabs that touch ghosts first
    cudaEvent_t trace_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // SSC_PIPE_TRACE timeline of the multi-rank host apply
    struct PeerPipe {
        std::vector<i64> cut, bc_off;                 // piece p = owned dofs [cut[p], cut[p+1]); Dirichlet list offsets per piece
        std::vector<int> copy_order, slab_order;      // H2D order of the pieces (those holding interface values first); launch order of the slabs
        std::vector<std::pair<int, int>> slab_pieces; // pieces [first, last] a slab reads owned entries from
        std::vector<char> slab_ghost, piece_late;     // slab touches ghost dofs; piece holds interface dofs (final only after the fold)
        std::vector<int> piece_ready;                 // position in slab_order after which the piece is complete
        int nsend_pieces = 0, nghost_slabs = 0;
    } pp;
    std::vector<cudaEvent_t> ev_in, ev_k;
    // caller's host vectors page-locked on first use (a PETSc Vec's array is pageable: without this the asynchronous copies of the
    // host-space apply are staged and serialised by the driver and the slab overlap is lost)
    bool register_host = true;       // -ssc_register_host_vectors
    std::vector<std::pair<const void *, size_t>> host_reg;
    i64 host_reg_count = 0;
    // timing
    std::map<std::string, double> event_ms;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ktimer = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kev;
    size_t kev_used = 0;
    double ktimer_ms = 0;
    i64 ktimer_n = 0;
    i64 launches = 0;
};
typedef SSCPatchPC_s PC;

static int set_device(PC *pc)
{
    if (pc->device < 0) return ssc_set_error(SSC_ERR_GPU, "this PC was created without a CUDA device (host-only handle); there is no CPU fallback");
    CUDA_OK(cudaSetDevice(pc->device));
    return 0;
}
template <class T> static int dev_alloc(T **p, i64 count)
{
    if (count <= 0) count = 1;
    cudaError_t e = cudaMalloc((void **)p, (size_t)count * sizeof(T));
    if (e != cudaSuccess) return ssc_set_error(SSC_ERR_MEM, "cudaMalloc of %lld bytes failed: %s", (long long)(count * (i64)sizeof(T)), cudaGetErrorString(e));
    return 0;
}
template <class T> static void dev_free(T *&p) { if (p) cudaFree(p); p = nullptr; }

static int upload_i32(PC *pc, const std::vector<i64> &src, int **dst)
{
    std::vector<int> tmp(src.size());
    for (size_t i = 0; i < src.size(); ++i) {
        if (src[i] < 0 || src[i] > 2147483647LL) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "index %lld does not fit the device's 32-bit dof index", (long long)src[i]);
        tmp[i] = (int)src[i];
    }
    CHK(dev_alloc(dst, (i64)tmp.size()));
    if (!tmp.empty()) CUDA_OK(cudaMemcpyAsync(*dst, tmp.data(), tmp.size() * sizeof(int), cudaMemcpyHostToDevice, pc->stream));
    CUDA_OK(cudaStreamSynchronize(pc->stream));
    return 0;
}

static inline int grid_for(i64 n, int block) { i64 g = (n + block - 1) / block; return (int)std::min<i64>(std::max<i64>(g, 1), 148 * 32); }

// The rest of the engine, by theme.  One translation unit on purpose: every launch site sees the same context
// struct and kernel templates, and the static helpers stay internal.
#include "engine_lifecycle.inc"        // library / lifecycle, options (PCSetFromOptions_PATCH), host inputs
#include "engine_lists.inc"            // first-time set-up: size classes, device lists, star-forest exchange over NCCL
#include "engine_peer.inc"             // NVLink peer exchange (CUDA IPC)
#include "engine_setup.inc"            // operator upload, extract / assemble + factor, weights, SSCPatchSetUp
#include "engine_apply.inc"            // apply: device-resident, multiplicative, pipelined host-space applies
#include "engine_introspect.inc"       // introspection, device plumbing, device patch construction
#include "coarse_amg.inc"              // smoothed-aggregation hierarchy of the coarse solver (host)
#include "engine_loworder.inc"         // low-order (P1) coarse correction
