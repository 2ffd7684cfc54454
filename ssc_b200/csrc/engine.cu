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
};

struct SSCPatchPC_s {
    int device = 0;
    cudaStream_t stream = nullptr, up_stream = nullptr;
    cudaStream_t side_stream = nullptr, launch_stream = nullptr;   // the small size classes of an apply run beside the big one; launch_apply_class launches on launch_stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap_classes = true;
    int apply_stream_min_waves = 4;  // ... and only when every CTA gets at least this many rounds of patches (tests: 0)
    int apply_stream_ctas = 0;       // CTAs of the persistent kernel (0: fill the GPU; tests use a few to get many rounds on small cases)
    int apply_stream_max_n = 96;     // largest patch the persistent apply kernel takes (n = 125: measured no better than the one-patch-per-CTA kernel)
    bool apply_exact_groups = true;  // n < 32: a group has exactly n threads (false: n rounded up to a multiple of 8)
    int apply_group_threads = 128;   // threads of a CTA that carries several small patches (n <= 96): 96 / 128 / 192 / 256 measured, tools/gpu_apply_sizes.py
    bool apply_stream_small = true;  // n <= 96: persistent apply kernel with the gathers prefetched across patches
    int sm_count = 148;
    size_t total_mem = (size_t)180 << 30;
    // options (defaults: PCCreate_PATCH ssc/libssc.c:1697-1711)
    bool save_operators = true, partition_of_unity = false, multiplicative = false, print_patches = false;
    bool symmetrise_sweep = false, keep_operators = false, dim_set = false, codim_set = false;
    int sub_pc = SSC_SUB_LU;
    int factor_tile = 4;          // register tiles of the n <= 128 factor kernel: 4 (8x4, default) or 8 (8x8, measured slower)
    int factor_ctas = 0;          // concurrent CTAs of the blocked factor kernel (0 = fill the SMs)
    int factor_panel = 16;        // panel width of the blocked factor kernel (16 or 32)
    int factor_searchers = 1;     // blocked look-ahead kernel: searcher warps (1 or 2)
    bool factor_remap = true;     // blocked look-ahead kernel: keep the searcher's SM sub-partition nearly free of worker warps
    int factor_variant = -1;      // n <= 128: 0 shared-memory Gauss-Jordan, 1 register-resident, 2 + look-ahead pivot search, 3 blocked look-ahead;
                                  // 4 blocked look-ahead with the searcher two panels ahead;
                                  // -1 (default): for 100 <= n <= 128 (one CTA per SM anyway) 4 with pivoting (lu), 3 without (cholesky); 1 below (many small CTAs per SM hide the latency)
    int apply_variant = 1;        // 0: simple loop, 1: pipelined LDG.64, 2: pipelined LDG.128 (even row stride)
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
    bool sym_two_chunk = false;      // packed apply: two bulk copies split at a column, two threads per row
    bool sym_bulk = true;            // packed apply: TMA bulk copy (true, 3.24 ms at 64^3 Q3) or per-thread cp.async (false, 4.46 ms)
    bool sym_persistent = false;     // persistent ring-buffered form of the packed apply kernel (measured slower: profiles/README.md)
    double *d_packed = nullptr;
    i64 packed_elems = 0;
    double *d_ypatch = nullptr;
    i64 *d_slotoff = nullptr;
    int *d_slots = nullptr;
    int *d_gtol = nullptr;
    double *d_blocks = nullptr, *d_blocks0 = nullptr;
    i64 block_elems = 0, sum_n = 0, sum_n2 = 0, max_n = 0;
    int *d_fail = nullptr;
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
    i64 csr_dev_rows = 0, csr_dev_nnz = 0;   // shape of the owned device copy
    // element matrices (the reference's own source of the patch operators, ssc/libssc.c:1120-1156): borrowed like the CSR arrays
    bool elem_set = false, elem_owned = false, elem_lists_ready = false, elem_maps_ready = false;
    i64 elem_map_rows = 0, elem_dev_len = 0;
    bool asm_shared = false;         // assemble kernel: accumulate in place through L2 (13.7 ms at 64^3 Q3) or in shared memory when the block fits (26 ms)
    int asm_threads = 256;
    int asm_ctas_per_sm = 16;        // blocks being accumulated at once per SM (their in-place working set competes for L2)
    const double *elem_K = nullptr;
    i64 elem_ncells = 0, elem_t = 0, elem_stride = 0;
    int elem_space = 0;
    double *d_elemK = nullptr;
    i64 *d_pcelloff = nullptr;       // cells of every patch, patches in size-sorted order
    int *d_pcells = nullptr;
    std::vector<int *> d_cnm;        // device copies of the cell-node maps
    ElemSpaces espaces;
    void *d_flush = nullptr;
    void *stage[3] = {nullptr, nullptr, nullptr};      // pinned staging ring for large pageable uploads
    cudaEvent_t stage_ev[3] = {nullptr, nullptr, nullptr};
    bool stage_busy[3] = {false, false, false};
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

// ---------------------------------------------------------------------------------------------
// library / lifecycle
// ---------------------------------------------------------------------------------------------
extern "C" int SSCDeviceCount(int *count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return 0;
}
extern "C" int SSCPatchInitializePackage(void) { return 0; }

extern "C" int SSCPatchCreate(int device, SSCPatchPC *out)
{
    if (device == -1) {            // host-only handle: patch construction (no arithmetic) without a GPU
        PC *pc = new PC();
        pc->device = -1;
        *out = pc;
        return 0;
    }
    int n = 0;
    SSCDeviceCount(&n);
    if (n <= 0) return ssc_set_error(SSC_ERR_GPU, "no CUDA device is visible; this library has no CPU fallback");
    if (device < 0 || device >= n) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "device %d out of range [0,%d)", device, n);
    PC *pc = new PC();
    pc->device = device;
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaStreamCreateWithFlags(&pc->stream, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&pc->up_stream, cudaStreamNonBlocking));
    {
        int lo_pri = 0, hi_pri = 0;                       // the small classes' blocks take free SM slots first
        CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        CUDA_OK(cudaStreamCreateWithPriority(&pc->side_stream, cudaStreamNonBlocking, hi_pri));
    }
    CUDA_OK(cudaEventCreateWithFlags(&pc->ev_fork, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&pc->ev_join, cudaEventDisableTiming));
    pc->launch_stream = pc->stream;
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    pc->sm_count = prop.multiProcessorCount;
    pc->total_mem = (size_t)prop.totalGlobalMem;
    CUDA_OK(cudaEventCreate(&pc->ev0));
    CUDA_OK(cudaEventCreate(&pc->ev1));
    *out = pc;
    return 0;
}

static void free_device_state(PC *pc)
{
    if (pc->device < 0) return;
    dev_free(pc->d_packed);
    dev_free(pc->d_ypatch); dev_free(pc->d_slotoff); dev_free(pc->d_slots);
    dev_free(pc->d_r); dev_free(pc->d_gtol); dev_free(pc->d_blocks); dev_free(pc->d_blocks0); dev_free(pc->d_fail); dev_free(pc->d_bc);
    dev_free(pc->d_w); dev_free(pc->d_mult); dev_free(pc->d_xs); dev_free(pc->d_ys);
    dev_free(pc->d_xl); dev_free(pc->d_yl);
    dev_free(pc->d_selfLeaf); dev_free(pc->d_selfRoot); dev_free(pc->d_sendRoot); dev_free(pc->d_recvLeaf);
    dev_free(pc->d_sendbuf); dev_free(pc->d_recvbuf);
    if (pc->csr_owned) { dev_free(pc->d_rowptr); dev_free(pc->d_colidx); dev_free(pc->d_vals); }
    pc->d_rowptr = nullptr; pc->d_colidx = nullptr; pc->d_vals = nullptr; pc->csr_owned = false;
    if (pc->elem_owned) dev_free(pc->d_elemK);
    pc->d_elemK = nullptr; pc->elem_owned = false;
    dev_free(pc->d_pcelloff); dev_free(pc->d_pcells);
    for (int *&q : pc->d_cnm) dev_free(q);
    pc->d_cnm.clear(); pc->elem_lists_ready = false; pc->elem_maps_ready = false; pc->elem_map_rows = 0;
    pc->classes.clear(); pc->device_ready = false; pc->setup_done = false; pc->pipe_ready = false;
}

extern "C" int SSCPatchReset(SSCPatchPC pc)
{
    if (!pc) return ssc_set_error(SSC_ERR_ARG_NULL, "null PC");
    if (pc->device >= 0) {
        CHK(set_device(pc));
        cudaStreamSynchronize(pc->stream);
        free_device_state(pc);
    }
    pc->lists = PatchLists(); pc->lists_valid = false; pc->direct_patches = false;
    pc->plex = HostPlex(); pc->disc = Discretisation(); pc->user = UserPatches();
    pc->csr_set = false; pc->elem_set = false; pc->sf_set = false;
    return 0;
}

extern "C" int SSCPatchDestroy(SSCPatchPC *ppc)
{
    if (!ppc || !*ppc) return 0;
    PC *pc = *ppc;
    if (pc->device < 0) { delete pc; *ppc = nullptr; return 0; }
    cudaSetDevice(pc->device);
    cudaStreamSynchronize(pc->stream);
    free_device_state(pc);
    if (pc->d_flush) cudaFree(pc->d_flush);
    for (int r = 0; r < 3; ++r) if (pc->stage[r]) { cudaFreeHost(pc->stage[r]); cudaEventDestroy(pc->stage_ev[r]); }
    for (void *q : pc->ipc_opened) cudaIpcCloseMemHandle(q);
    pc->p2p = false;
    dev_free(pc->d_xg); dev_free(pc->d_yin); dev_free(pc->d_iface); dev_free(pc->d_flags); dev_free(pc->d_timeout); dev_free(pc->d_ps);
    dev_free(pc->d_ghostPeer); dev_free(pc->d_ghostRemote); dev_free(pc->d_sendPeer); dev_free(pc->d_remoteLeaf);
    if (pc->s_in) { cudaStreamDestroy(pc->s_in); cudaStreamDestroy(pc->s_out); }
    for (auto &ev : pc->trace_ev) if (ev) { cudaEventDestroy(ev); ev = nullptr; }
    for (auto &ev : pc->ev_in) cudaEventDestroy(ev);
    for (auto &ev : pc->ev_k) cudaEventDestroy(ev);
    for (auto &e : pc->kev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (pc->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(pc->comm);
    cudaEventDestroy(pc->ev0); cudaEventDestroy(pc->ev1);
    cudaStreamDestroy(pc->stream);
    cudaStreamDestroy(pc->up_stream);
    if (pc->side_stream) { cudaStreamDestroy(pc->side_stream); cudaEventDestroy(pc->ev_fork); cudaEventDestroy(pc->ev_join); }
    delete pc;
    *ppc = nullptr;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// options
// ---------------------------------------------------------------------------------------------
static int parse_bool(const char *v, bool *out)
{
    std::string s(v ? v : "");
    for (auto &c : s) c = (char)tolower(c);
    if (s == "" || s == "1" || s == "true" || s == "yes" || s == "on") { *out = true; return 0; }
    if (s == "0" || s == "false" || s == "no" || s == "off") { *out = false; return 0; }
    return ssc_set_error(SSC_ERR_USER, "cannot read '%s' as a boolean", v);
}

extern "C" int SSCPatchSetSaveOperators(SSCPatchPC pc, int flg) { pc->save_operators = flg != 0; return 0; }
extern "C" int SSCPatchSetPartitionOfUnity(SSCPatchPC pc, int flg) { pc->partition_of_unity = flg != 0; return 0; }
extern "C" int SSCPatchSetSubMatType(SSCPatchPC pc, const char *t)
{
    std::string s(t ? t : "");
    // every PETSc matrix type the reference could be given stores the same numbers; here they all are dense blocks
    if (s != "dense" && s != "seqdense" && s != "aij" && s != "seqaij")
        return ssc_set_error(SSC_ERR_SUP, "pc_patch_sub_mat_type '%s' is not supported (dense, seqdense, aij, seqaij)", s.c_str());
    pc->sub_mat_type = s;
    return 0;
}

extern "C" int SSCPatchSetOption(SSCPatchPC pc, const char *name_, const char *value)
{
    if (!pc || !name_) return ssc_set_error(SSC_ERR_ARG_NULL, "null argument");
    std::string name(name_);
    if (!name.empty() && name[0] == '-') name = name.substr(1);
    const char *v = value ? value : "";
    if (name == "pc_patch_save_operators") return parse_bool(v, &pc->save_operators);
    if (name == "pc_patch_partition_of_unity") return parse_bool(v, &pc->partition_of_unity);
    if (name == "pc_patch_multiplicative") {
        // The reference removed this mode ("We've removed multiplicative to do BC condensation (for now)",
        // ssc/libssc.c:1568-1572).  north_star asks for it colour by colour; see DESIGN.md section 8 (N4).
        bool flg = false;
        CHK(parse_bool(v, &flg));
        if (flg != pc->multiplicative && pc->device_ready) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "pc_patch_multiplicative must be chosen before the first SetUp");
        pc->multiplicative = flg;
        return 0;
    }
    if (name == "pc_patch_construction_dim") {
        pc->bopt.dim = atoll(v); pc->dim_set = true;
        if (pc->codim_set) return ssc_set_error(SSC_ERR_USER, "Can only set one of dimension or co-dimension");   // :1576-1578
        return 0;
    }
    if (name == "pc_patch_construction_codim") {
        pc->bopt.codim = atoll(v); pc->codim_set = true;
        if (pc->dim_set) return ssc_set_error(SSC_ERR_USER, "Can only set one of dimension or co-dimension");
        return 0;
    }
    if (name == "pc_patch_construction_type") {             // :1580-1603
        std::string s(v);
        if (s == "star") pc->bopt.construct_type = SSC_PATCH_STAR;
        else if (s == "vanka") pc->bopt.construct_type = SSC_PATCH_VANKA;
        else if (s == "user") pc->bopt.construct_type = SSC_PATCH_USER;
        else if (s == "python") pc->bopt.construct_type = SSC_PATCH_PYTHON;
        else return ssc_set_error(SSC_ERR_USER, "Unknown patch construction type");
        return 0;
    }
    if (name == "pc_patch_vanka_dim") { pc->bopt.vankadim = atoll(v); return 0; }
    if (name == "pc_patch_vanka_space")                       // :1589-1592
        return ssc_set_error(SSC_ERR_USER, "-pc_patch_vanka_space has been renamed to -pc_patch_exclude_subspace");
    if (name == "pc_patch_exclude_subspace") { pc->bopt.exclude_subspace = atoll(v); return 0; }
    if (name == "pc_patch_sub_mat_type") return SSCPatchSetSubMatType(pc, v);
    if (name == "pc_patch_print_patches") { CHK(parse_bool(v, &pc->print_patches)); pc->bopt.print_patches = pc->print_patches; pc->lists_valid = pc->lists_valid && !pc->print_patches; return 0; }
    if (name == "pc_patch_symmetrise_sweep") return parse_bool(v, &pc->symmetrise_sweep);
    if (name == "sub_ksp_type") {
        if (std::string(v) != "preonly") return ssc_set_error(SSC_ERR_SUP, "sub_ksp_type '%s' is not supported: patch problems are solved directly (preonly)", v);
        return 0;
    }
    if (name == "sub_pc_type") {
        std::string s(v);
        if (s == "lu") pc->sub_pc = SSC_SUB_LU;
        else if (s == "cholesky") pc->sub_pc = SSC_SUB_CHOLESKY;
        else return ssc_set_error(SSC_ERR_SUP, "sub_pc_type '%s' is not supported (lu, cholesky)", v);
        return 0;
    }
    if (name == "sub_pc_factor_mat_solver_type" || name == "sub_pc_factor_mat_ordering_type") return 0;   // orderings do not change a dense solve
    if (name == "ssc_keep_dofs_table") return parse_bool(v, &pc->bopt.keep_dofs);
    if (name == "ssc_keep_patch_operators") return parse_bool(v, &pc->keep_operators);
    if (name == "ssc_builder_threads") { pc->bopt.nthreads = atoi(v); return 0; }
    if (name == "ssc_factor_variant") { pc->factor_variant = atoi(v); if (pc->factor_variant < -1 || pc->factor_variant > 4) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "ssc_factor_variant must be -1 (automatic) or 0..4"); return 0; }
    if (name == "ssc_symmetric_storage") {
        bool f = false;
        CHK(parse_bool(v, &f));
        if (f != pc->symmetric && pc->device_ready) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "ssc_symmetric_storage must be chosen before the first SetUp");
        pc->symmetric = f;
        return 0;
    }
    if (name == "ssc_symmetric_persistent") return parse_bool(v, &pc->sym_persistent);
    if (name == "ssc_symmetric_bulk_copy") return parse_bool(v, &pc->sym_bulk);
    if (name == "ssc_symmetric_two_chunk") return parse_bool(v, &pc->sym_two_chunk);
    if (name == "ssc_assemble_shared") return parse_bool(v, &pc->asm_shared);
    if (name == "ssc_assemble_ctas_per_sm") { pc->asm_ctas_per_sm = std::max(1, atoi(v)); return 0; }
    if (name == "ssc_assemble_threads") { pc->asm_threads = atoi(v); if (pc->asm_threads != 128 && pc->asm_threads != 256 && pc->asm_threads != 512 && pc->asm_threads != 1024) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "ssc_assemble_threads must be 128, 256, 512 or 1024"); return 0; }
    if (name == "ssc_device_patch_construction") { CHK(parse_bool(v, &pc->device_construct)); pc->lists_valid = false; return 0; }
    if (name == "ssc_deterministic") {
        bool f = false;
        CHK(parse_bool(v, &f));
        if (f != pc->deterministic && pc->device_ready) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "ssc_deterministic must be chosen before the first SetUp");
        pc->deterministic = f;
        return 0;
    }
    if (name == "ssc_factor_remap") return parse_bool(v, &pc->factor_remap);
    if (name == "ssc_factor_searchers") { pc->factor_searchers = atoi(v); return 0; }
    if (name == "ssc_overlap_classes") return parse_bool(v, &pc->overlap_classes);
    if (name == "ssc_apply_stream_small") return parse_bool(v, &pc->apply_stream_small);
    if (name == "ssc_apply_exact_groups") return parse_bool(v, &pc->apply_exact_groups);
    if (name == "ssc_apply_stream_min_waves") { pc->apply_stream_min_waves = atoi(v); return 0; }
    if (name == "ssc_apply_stream_ctas") { pc->apply_stream_ctas = atoi(v); return 0; }
    if (name == "ssc_apply_stream_max_n") { pc->apply_stream_max_n = atoi(v); return 0; }
    if (name == "ssc_apply_group_threads") { pc->apply_group_threads = atoi(v); if (pc->apply_group_threads < 32 || pc->apply_group_threads > 256) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "ssc_apply_group_threads must be in [32, 256]"); return 0; }
    if (name == "ssc_factor_panel") { pc->factor_panel = atoi(v); return 0; }
    if (name == "ssc_factor_ctas") { pc->factor_ctas = atoi(v); return 0; }
    if (name == "ssc_factor_tile") { pc->factor_tile = atoi(v); return 0; }
    if (name == "ssc_host_pipeline_taper") { pc->pipe_ready = false; pc->pipe_tried = false; return parse_bool(v, &pc->pipe_taper); }
    if (name == "ssc_host_pipeline") { const int s = atoi(v); pc->pipe_force = s < 0; pc->pipe_slabs = s < 0 ? -s : s; pc->pipe_ready = false; return 0; }   // negative: use it whatever the problem size
    if (name == "ssc_apply_variant") {
        const int a = atoi(v);
        if (a < 0 || a > 2) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "ssc_apply_variant must be 0, 1 or 2");
        if (pc->device_ready && (a == 2) != (pc->apply_variant == 2)) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "the block layout of variant 2 must be chosen before the first SetUp");
        pc->apply_variant = a;
        return 0;
    }
    return ssc_set_error(SSC_ERR_USER, "unknown option '%s'", name.c_str());
}

// ---------------------------------------------------------------------------------------------
// host inputs
// ---------------------------------------------------------------------------------------------
extern "C" int SSCPatchSetPlex(SSCPatchPC pc, SSCInt dim, SSCInt pEnd, const SSCInt *coneOff, const SSCInt *cones,
                               const SSCInt *depthStart, const SSCInt *depthEnd, const unsigned char *ghost)
{
    if (!pc || !coneOff || !depthStart || !depthEnd) return ssc_set_error(SSC_ERR_ARG_NULL, "null argument");
    HostPlex &P = pc->plex;
    P.dim = dim; P.pEnd = pEnd;
    P.coneOff.assign(coneOff, coneOff + pEnd + 1);
    P.cones.assign(cones, cones + coneOff[pEnd]);
    P.depthStart.assign(depthStart, depthStart + dim + 1);
    P.depthEnd.assign(depthEnd, depthEnd + dim + 1);
    if (ghost) P.ghost.assign(ghost, ghost + pEnd); else P.ghost.clear();
    // supports = transpose of the cone table
    P.suppOff.assign(pEnd + 1, 0);
    for (i64 k = 0; k < (i64)P.cones.size(); ++k) {
        if (P.cones[k] < 0 || P.cones[k] >= pEnd) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "cone entry outside the chart");
        P.suppOff[P.cones[k] + 1]++;
    }
    for (i64 p = 0; p < pEnd; ++p) P.suppOff[p + 1] += P.suppOff[p];
    P.supports.resize(P.cones.size());
    std::vector<i64> fill(P.suppOff.begin(), P.suppOff.end() - 1);
    for (i64 p = 0; p < pEnd; ++p)
        for (i64 k = P.coneOff[p]; k < P.coneOff[p + 1]; ++k) P.supports[fill[P.cones[k]]++] = p;
    P.set = true;
    pc->lists_valid = false;
    return 0;
}

extern "C" int SSCPatchSetCellNumbering(SSCPatchPC pc, const SSCInt *dof, const SSCInt *off)
{
    if (!pc->plex.set) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "set the plex before the cell numbering");
    pc->disc.cnDof.assign(dof, dof + pc->plex.pEnd);
    pc->disc.cnOff.assign(off, off + pc->plex.pEnd);
    pc->disc.cn_set = true;
    pc->lists_valid = false;
    return 0;
}

extern "C" int SSCPatchSetDiscretisationInfo(SSCPatchPC pc, SSCInt nsubspaces, const SSCInt *const *secDof,
                                             const SSCInt *const *secOff, const SSCInt *bs, const SSCInt *nodesPerCell,
                                             const SSCInt *const *cellNodeMap, const SSCInt *subspaceOffsets,
                                             SSCInt numGhostBcs, const SSCInt *ghostBcNodes,
                                             SSCInt numGlobalBcs, const SSCInt *globalBcNodes)
{
    if (!pc->plex.set) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "set the plex before the discretisation info");
    Discretisation &D = pc->disc;
    D.nsub = nsubspaces; D.totalDofsPerCell = 0;
    D.secDof.resize(nsubspaces); D.secOff.resize(nsubspaces);
    D.bs.assign(bs, bs + nsubspaces);
    D.nodesPerCell.assign(nodesPerCell, nodesPerCell + nsubspaces);
    D.subspaceOffsets.assign(subspaceOffsets, subspaceOffsets + nsubspaces + 1);
    D.cellNodeMap.assign(cellNodeMap, cellNodeMap + nsubspaces);
    for (i64 k = 0; k < nsubspaces; ++k) {
        D.secDof[k].assign(secDof[k], secDof[k] + pc->plex.pEnd);
        D.secOff[k].assign(secOff[k], secOff[k] + pc->plex.pEnd);
        D.totalDofsPerCell += nodesPerCell[k] * bs[k];
    }
    D.ghostBcNodes.assign(ghostBcNodes, ghostBcNodes + numGhostBcs);
    D.globalBcNodes.assign(globalBcNodes, globalBcNodes + numGlobalBcs);
    D.set = true;
    pc->globalBc = D.globalBcNodes;
    pc->nlocal = subspaceOffsets[nsubspaces];
    if (!pc->sf_set) pc->nowned = pc->nlocal;
    pc->lists_valid = false; pc->direct_patches = false;
    return 0;
}

extern "C" int SSCPatchSetUserPatches(SSCPatchPC pc, SSCInt nuserIS, const SSCInt *userOff, const SSCInt *userPoints,
                                      SSCInt niter, const SSCInt *iterationSet)
{
    pc->user.off.assign(userOff, userOff + nuserIS + 1);
    pc->user.pts.assign(userPoints, userPoints + userOff[nuserIS]);
    pc->user.iterationSet.assign(iterationSet, iterationSet + niter);
    for (i64 v : pc->user.iterationSet)
        if (v < 0 || v >= nuserIS) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "iteration set entry %lld outside [0,%lld)", (long long)v, (long long)nuserIS);
    pc->user.set = true;
    pc->lists_valid = false;
    return 0;
}

extern "C" int SSCPatchSetPatches(SSCPatchPC pc, SSCInt npatch, const SSCInt *offsets, const SSCInt *gtol,
                                  SSCInt nlocal, SSCInt nowned, SSCInt numGlobalBcs, const SSCInt *globalBcNodes)
{
    if (!pc || !offsets) return ssc_set_error(SSC_ERR_ARG_NULL, "null argument");
    PatchLists &L = pc->lists;
    L = PatchLists();
    L.npatch = npatch; L.vStart = 0;
    L.gtolOff.assign(offsets, offsets + npatch + 1);
    L.gtol.assign(gtol, gtol + offsets[npatch]);
    L.cellOff.assign(npatch + 1, 0);
    if (offsets[0] != 0) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "offsets[0] must be 0");
    for (i64 p = 0; p < npatch; ++p) if (offsets[p + 1] < offsets[p]) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "patch offsets must not decrease (patch %lld)", (long long)p);
    for (i64 g : L.gtol) if (g < 0 || g >= nlocal) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "patch dof %lld outside [0,%lld)", (long long)g, (long long)nlocal);
    {   // a dof may appear once per patch (the reference's hash map guarantees it, ssc/libssc.c:763-767)
        std::vector<i64> stamp((size_t)std::max<i64>(nlocal, 1), -1);
        for (i64 p = 0; p < npatch; ++p)
            for (i64 k = offsets[p]; k < offsets[p + 1]; ++k) {
                if (stamp[(size_t)gtol[k]] == p) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "dof %lld is listed twice in patch %lld", (long long)gtol[k], (long long)p);
                stamp[(size_t)gtol[k]] = p;
            }
    }
    pc->nlocal = nlocal;
    if (!pc->sf_set) pc->nowned = nowned;
    pc->globalBc.assign(globalBcNodes, globalBcNodes + numGlobalBcs);
    pc->lists_valid = true; pc->direct_patches = true;
    free_device_state(pc);
    return 0;
}

extern "C" int SSCPatchSetStarForest(SSCPatchPC pc, SSCInt nowned, SSCInt nself, const SSCInt *selfLeaf, const SSCInt *selfRoot,
                                     SSCInt nneigh, const int *neigh, const SSCInt *sendOff, const SSCInt *sendRoot,
                                     const SSCInt *recvOff, const SSCInt *recvLeaf)
{
    pc->nowned = nowned;
    pc->selfLeaf.assign(selfLeaf, selfLeaf + nself);
    pc->selfRoot.assign(selfRoot, selfRoot + nself);
    pc->neigh.assign(neigh, neigh + nneigh);
    pc->sendOff.assign(sendOff, sendOff + nneigh + 1);
    pc->sendRoot.assign(sendRoot, sendRoot + sendOff[nneigh]);
    pc->recvOff.assign(recvOff, recvOff + nneigh + 1);
    pc->recvLeaf.assign(recvLeaf, recvLeaf + recvOff[nneigh]);
    pc->sf_set = true;
    free_device_state(pc);
    return 0;
}

extern "C" int SSCCommGetUniqueId(char uniqueId[128])
{
    CHK(nccl_load());
    ncclUniqueId_t id;
    NCCL_OK(g_nccl.GetUniqueId(&id));
    memcpy(uniqueId, id.internal, 128);
    return 0;
}

extern "C" int SSCPatchCommInit(SSCPatchPC pc, int nranks, int rank, const char uniqueId[128])
{
    CHK(nccl_load());
    CHK(set_device(pc));
    ncclUniqueId_t id;
    memcpy(id.internal, uniqueId, 128);
    NCCL_OK(g_nccl.CommInitRank(&pc->comm, nranks, id, rank));
    pc->nranks = nranks; pc->rank = rank;
    return 0;
}

extern "C" int SSCPatchSetOperatorCSR(SSCPatchPC pc, SSCInt nrows, const SSCInt *rowptr, const int32_t *colidx,
                                      const double *vals, int memspace)
{
    if (!rowptr || !colidx || !vals) return ssc_set_error(SSC_ERR_ARG_NULL, "null CSR array");
    pc->csr_nrows = nrows; pc->csr_rowptr = rowptr; pc->csr_colidx = colidx; pc->csr_vals = vals;
    pc->csr_space = memspace; pc->csr_set = true;
    pc->elem_set = false; pc->op_released = false;
    return 0;
}

// The reference's route to the patch operators: one totalDofsPerCell x totalDofsPerCell matrix per cell, what its
// compute-operator callback hands to MatSetValues cell by cell (ssc/libssc.c:1150, ssc/patch.py:28-62,210-213).
extern "C" int SSCPatchSetElementMatrices(SSCPatchPC pc, SSCInt ncells, SSCInt dofs_per_cell, const double *Ke, SSCInt cell_stride, int memspace)
{
    if (!pc) return ssc_set_error(SSC_ERR_ARG_NULL, "null PC");
    if (!Ke) return ssc_set_error(SSC_ERR_ARG_NULL, "null element matrices");
    if (ncells < 1 || dofs_per_cell < 1) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "need at least one cell and one dof per cell");
    if (cell_stride != 0 && cell_stride < dofs_per_cell * dofs_per_cell) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "cell stride %lld is smaller than one %lld x %lld matrix (0 = one matrix shared by all cells)", (long long)cell_stride, (long long)dofs_per_cell, (long long)dofs_per_cell);
    pc->elem_K = Ke; pc->elem_ncells = ncells; pc->elem_t = dofs_per_cell; pc->elem_stride = cell_stride; pc->elem_space = memspace;
    pc->elem_set = true;
    pc->csr_set = false; pc->op_released = false;
    return 0;
}

// The operator arrays are only borrowed (ssc/libssc.c never owns the callback's data either): the caller announces that they
// are gone.  Device copies made by the last SSCPatchSetUp stay; a further SSCPatchSetUp needs a new operator first.
extern "C" int SSCPatchReleaseOperator(SSCPatchPC pc)
{
    if (!pc) return ssc_set_error(SSC_ERR_ARG_NULL, "null PC");
    pc->op_released = true;
    pc->csr_rowptr = nullptr; pc->csr_colidx = nullptr; pc->csr_vals = nullptr; pc->elem_K = nullptr;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// setup
// ---------------------------------------------------------------------------------------------
static void choose_launch_shape(SizeClass &c, int variant, int group_threads = 128, bool exact_groups = true)
{
    int n = c.n;
    if (variant == 2) n = (c.n + 1) / 2;            // one thread per pair of rows
    if (n >= 1024) { c.tp = 1024; c.ppb = 1; return; }
    if (n > 96) { c.tp = (n + 31) / 32 * 32; c.ppb = 1; return; }
    c.tp = (n < 32 && exact_groups) ? std::max(n, 1) : (n + 7) / 8 * 8;       // sub-warp groups for small patches: exactly n threads (a thread per patch at degree 1)
    c.ppb = std::max(1, group_threads / c.tp);
}

static int need_lists(PC *pc);
// first-time part of PCSetUp_PATCH (ssc/libssc.c:1210-1250): dof lists -> size classes -> device arrays
static int build_device_lists(PC *pc)
{
    auto t0 = std::chrono::steady_clock::now();
    CHK(need_lists(pc));
    pc->event_ms["PCPATCHCreate"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    const PatchLists &L = pc->lists;
    const i64 np = L.npatch;
    // stable sort of the non-empty patches by size
    std::vector<i64> order;
    order.reserve(np);
    for (i64 p = 0; p < np; ++p) if (L.gtolOff[p + 1] - L.gtolOff[p] > 0) order.push_back(p);   // ssc/libssc.c:1355-1358
    if (pc->multiplicative) {
        // Greedy colouring of the patch conflict graph: two patches conflict when they share a cell, i.e. when the
        // operator can couple their dofs.  Patches of one colour then commute and are applied together.
        if (pc->direct_patches || L.cells.empty()) return ssc_set_error(SSC_ERR_SUP, "multiplicative sweeps need patches built from the mesh topology (cells per patch)");
        if (pc->sf_set) return ssc_set_error(SSC_ERR_SUP, "multiplicative sweeps across ranks are not supported in this build");
        if (pc->partition_of_unity) return ssc_set_error(SSC_ERR_ARG_INCOMP, "partition of unity weighting applies to the additive method only");
        i64 maxcell = -1;
        for (i64 c : L.cells) maxcell = std::max(maxcell, c);
        std::vector<i64> coff((size_t)maxcell + 2, 0);
        for (i64 c : L.cells) coff[(size_t)c + 1]++;
        for (size_t c = 0; c + 1 < coff.size(); ++c) coff[c + 1] += coff[c];
        std::vector<i64> cpatch((size_t)L.cells.size()), fill(coff.begin(), coff.end() - 1);
        for (i64 p = 0; p < np; ++p) for (i64 k = L.cellOff[p]; k < L.cellOff[p + 1]; ++k) cpatch[(size_t)fill[(size_t)L.cells[k]]++] = p;
        pc->colour.assign((size_t)np, -1);
        std::vector<i64> mark;
        pc->ncolours = 0;
        for (i64 p = 0; p < np; ++p) {
            if (L.gtolOff[p + 1] == L.gtolOff[p]) continue;
            mark.assign((size_t)pc->ncolours + 1, 0);
            for (i64 k = L.cellOff[p]; k < L.cellOff[p + 1]; ++k) {
                const i64 c = L.cells[k];
                for (i64 m = coff[(size_t)c]; m < coff[(size_t)c + 1]; ++m) { const i64 q = cpatch[(size_t)m]; if (q != p && pc->colour[(size_t)q] >= 0) mark[(size_t)pc->colour[(size_t)q]] = 1; }
            }
            i64 col = 0;
            while (col < pc->ncolours && mark[(size_t)col]) ++col;
            pc->colour[(size_t)p] = col;
            pc->ncolours = std::max(pc->ncolours, col + 1);
        }
        std::stable_sort(order.begin(), order.end(), [&](i64 a, i64 b) {
            const i64 na = L.gtolOff[a + 1] - L.gtolOff[a], nb = L.gtolOff[b + 1] - L.gtolOff[b];
            return na != nb ? na < nb : pc->colour[(size_t)a] < pc->colour[(size_t)b];
        });
    } else
    std::stable_sort(order.begin(), order.end(), [&](i64 a, i64 b) { return L.gtolOff[a + 1] - L.gtolOff[a] < L.gtolOff[b + 1] - L.gtolOff[b]; });
    pc->sorted2orig = order;
    pc->orig2sorted.assign(np, -1);
    pc->classes.clear();
    pc->sum_n = 0; pc->sum_n2 = 0; pc->max_n = 0;
    i64 goff = 0, moff = 0;
    for (i64 q = 0; q < (i64)order.size(); ++q) {
        const i64 p = order[q], n = L.gtolOff[p + 1] - L.gtolOff[p];
        pc->orig2sorted[p] = q;
        if (pc->classes.empty() || pc->classes.back().n != (int)n) {
            SizeClass c;
            moff = (moff + 15) / 16 * 16;                 // every class starts on a 128-byte line (a class of tiny blocks is unpadded inside)
            c.n = (int)n; c.first = q; c.goff = goff; c.moff = moff;
            c.ld = pc->apply_variant == 2 ? (int)((n + 1) / 2 * 2) : (int)n;
            c.stride = (n * c.ld + 15) / 16 * 16;
            if (n * c.ld <= 4 && pc->apply_variant != 2) c.stride = n * c.ld;      // tiny blocks (degree 1: one entry) are not padded to a 128-byte line
            choose_launch_shape(c, pc->apply_variant, pc->apply_group_threads, pc->apply_exact_groups);
            { const i64 psz = n * (n + 1) / 2; c.pstride = (psz + 15) / 16 * 16; c.pbytes = (unsigned)(((psz + 1) / 2 * 2) * 8); }
            pc->classes.push_back(c);
        }
        SizeClass &c = pc->classes.back();
        c.count++;
        goff += n; moff += c.stride;
        pc->sum_n += n; pc->sum_n2 += n * n; pc->max_n = std::max(pc->max_n, n);
    }
    pc->block_elems = moff;
    {
        i64 po = 0;
        for (SizeClass &c : pc->classes) { c.poff = po; po += c.count * c.pstride; }
        pc->packed_elems = po;
        if (pc->symmetric) for (const SizeClass &c : pc->classes)
            if ((size_t)c.pbytes + (size_t)(2 * c.n + 32) * 8 > 225 * 1024) return ssc_set_error(SSC_ERR_SUP, "ssc_symmetric_storage holds a whole packed block in shared memory: patches up to 238 dofs (found %d)", c.n);
    }
    if (pc->multiplicative) {
        pc->colour_ranges.assign((size_t)pc->ncolours, std::vector<std::pair<i64, i64>>(pc->classes.size(), {0, 0}));
        for (size_t ci = 0; ci < pc->classes.size(); ++ci) {
            const SizeClass &c = pc->classes[ci];
            i64 q = c.first;
            while (q < c.first + c.count) {
                const i64 col = pc->colour[(size_t)order[(size_t)q]];
                i64 q1 = q;
                while (q1 < c.first + c.count && pc->colour[(size_t)order[(size_t)q1]] == col) ++q1;
                pc->colour_ranges[(size_t)col][ci] = {q, q1};
                q = q1;
            }
        }
        CHK(dev_alloc(&pc->d_r, pc->nlocal));
    }
    if (pc->nlocal > 2147483647LL) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "local vector too long for 32-bit device indices");
    pc->event_ms["SSCListsSort"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() - pc->event_ms["PCPATCHCreate"];
    auto t1 = std::chrono::steady_clock::now();
    // dof lists in sorted order, int32
    RawArray<int> g32;                                     // filled by the threads below: no serial zero-fill of 130 MB
    g32.resize((size_t)pc->sum_n);
    {
        std::vector<i64> woff(order.size() + 1, 0);              // write offset of every sorted patch
        for (size_t q = 0; q < order.size(); ++q) woff[q + 1] = woff[q] + (L.gtolOff[(size_t)order[q] + 1] - L.gtolOff[(size_t)order[q]]);
        const int Tn = (int)std::min<size_t>(std::min<size_t>(32, std::max(1u, std::thread::hardware_concurrency())), std::max<size_t>(1, order.size() / 4096));
        auto work = [&](int t) {
            for (size_t q = order.size() * t / Tn; q < order.size() * (t + 1) / Tn; ++q) {
                const i64 p = order[q];
                i64 w = woff[q];
                for (i64 k = L.gtolOff[(size_t)p]; k < L.gtolOff[(size_t)p + 1]; ++k) g32[(size_t)w++] = (int)L.gtol[(size_t)k];
            }
        };
        if (Tn == 1) work(0);
        else { std::vector<std::thread> th; for (int t = 0; t < Tn; ++t) th.emplace_back(work, t); for (auto &x : th) x.join(); }
    }
    pc->event_ms["SSCListsPack"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    t1 = std::chrono::steady_clock::now();
    CHK(dev_alloc(&pc->d_gtol, pc->sum_n));
    if (pc->sum_n) CUDA_OK(cudaMemcpyAsync(pc->d_gtol, g32.data(), (size_t)pc->sum_n * sizeof(int), cudaMemcpyHostToDevice, pc->stream));
    CUDA_OK(cudaStreamSynchronize(pc->stream));
    pc->event_ms["SSCListsUpload"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    CHK(dev_alloc(&pc->d_fail, (i64)order.size()));
    if (pc->deterministic) {
        if (pc->sf_set) return ssc_set_error(SSC_ERR_SUP, "ssc_deterministic is a single-rank mode in this build");
        if (pc->sum_n > 2147483647LL) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "too many patch dofs for 32-bit slot numbers");
        std::vector<i64> soff((size_t)pc->nlocal + 1, 0);
        for (i64 s = 0; s < pc->sum_n; ++s) soff[(size_t)g32[(size_t)s] + 1]++;
        for (i64 g = 0; g < pc->nlocal; ++g) soff[(size_t)g + 1] += soff[(size_t)g];
        std::vector<int> slots((size_t)pc->sum_n);
        std::vector<i64> fill(soff.begin(), soff.end() - 1);
        for (i64 s = 0; s < pc->sum_n; ++s) slots[(size_t)fill[(size_t)g32[(size_t)s]]++] = (int)s;      // ascending slot order per dof
        CHK(dev_alloc(&pc->d_slotoff, pc->nlocal + 1)); CHK(dev_alloc(&pc->d_slots, pc->sum_n)); CHK(dev_alloc(&pc->d_ypatch, pc->sum_n));
        CUDA_OK(cudaMemcpy(pc->d_slotoff, soff.data(), soff.size() * sizeof(i64), cudaMemcpyHostToDevice));
        if (pc->sum_n) CUDA_OK(cudaMemcpy(pc->d_slots, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    // Dirichlet pass-through list, ssc/libssc.c:1408-1413 (only idx < local size)
    std::vector<i64> bc;
    for (i64 d : pc->globalBc) if (d >= 0 && d < pc->nowned) bc.push_back(d);
    pc->nbc = (i64)bc.size();
    CHK(upload_i32(pc, bc, &pc->d_bc));
    // iteration multiplicity for user patches (the iteration set may repeat or omit patches, :1347-1351)
    if (pc->bopt.construct_type >= 2 && pc->user.set && !pc->direct_patches) {
        std::vector<double> mult(order.size(), 0.0);
        for (i64 v : pc->user.iterationSet) if (pc->orig2sorted[v] >= 0) mult[pc->orig2sorted[v]] += 1.0;
        CHK(dev_alloc(&pc->d_mult, (i64)mult.size()));
        if (!mult.empty()) CUDA_OK(cudaMemcpy(pc->d_mult, mult.data(), mult.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    // star forest
    const bool identity = !pc->sf_set;
    if (!identity) {
        CHK(upload_i32(pc, pc->selfLeaf, &pc->d_selfLeaf));
        CHK(upload_i32(pc, pc->selfRoot, &pc->d_selfRoot));
        CHK(upload_i32(pc, pc->sendRoot, &pc->d_sendRoot));
        CHK(upload_i32(pc, pc->recvLeaf, &pc->d_recvLeaf));
        CHK(dev_alloc(&pc->d_sendbuf, (i64)std::max(pc->sendRoot.size(), pc->recvLeaf.size())));
        CHK(dev_alloc(&pc->d_recvbuf, (i64)std::max(pc->sendRoot.size(), pc->recvLeaf.size())));
        if (!pc->p2p && !pc->d_xl) CHK(dev_alloc(&pc->d_xl, pc->nlocal));
        if (!pc->p2p && !pc->d_yl) CHK(dev_alloc(&pc->d_yl, pc->nlocal));
        if (!pc->neigh.empty() && !pc->comm && !pc->p2p) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "the star forest names neighbour ranks but SSCPatchCommInit() has not been called");
    } else if (pc->nowned != pc->nlocal) {
        return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "nowned (%lld) != nlocal (%lld) but no star forest was set", (long long)pc->nowned, (long long)pc->nlocal);
    }
    pc->device_ready = true;
    return 0;
}

// halo fill: SFBcast of ssc/libssc.c:1323-1324 (owner values -> local vector incl. ghosts)
static int sf_bcast(PC *pc, const double *x, double *xl)
{
    const i64 nself = (i64)pc->selfLeaf.size();
    if (nself) { ssc_copy_indexed_kernel<<<grid_for(nself, 256), 256, 0, pc->stream>>>(nself, pc->d_selfLeaf, pc->d_selfRoot, x, xl, 0); pc->launches++; }
    const i64 ns = (i64)pc->sendRoot.size(), nr = (i64)pc->recvLeaf.size();
    if (pc->neigh.empty()) return 0;
    if (ns) { ssc_gather_kernel<<<grid_for(ns, 256), 256, 0, pc->stream>>>(ns, pc->d_sendRoot, x, pc->d_sendbuf); pc->launches++; }
    NCCL_OK(g_nccl.GroupStart());
    for (size_t k = 0; k < pc->neigh.size(); ++k) {
        const i64 s0 = pc->sendOff[k], s1 = pc->sendOff[k + 1], r0 = pc->recvOff[k], r1 = pc->recvOff[k + 1];
        if (s1 > s0) NCCL_OK(g_nccl.Send(pc->d_sendbuf + s0, (size_t)(s1 - s0), kNcclFloat64, pc->neigh[k], pc->comm, pc->stream));
        if (r1 > r0) NCCL_OK(g_nccl.Recv(pc->d_recvbuf + r0, (size_t)(r1 - r0), kNcclFloat64, pc->neigh[k], pc->comm, pc->stream));
    }
    NCCL_OK(g_nccl.GroupEnd());
    if (nr) { ssc_scatter_kernel<<<grid_for(nr, 256), 256, 0, pc->stream>>>(nr, pc->d_recvLeaf, pc->d_recvbuf, xl, 0); pc->launches++; }
    return 0;
}

// overlap sum: SFReduce(MPI_SUM) of ssc/libssc.c:1399-1400 (local contributions, ghosts included -> owners); y zeroed first (:1396)
static int sf_reduce(PC *pc, const double *yl, double *y)
{
    CUDA_OK(cudaMemsetAsync(y, 0, (size_t)pc->nowned * sizeof(double), pc->stream));
    const i64 nself = (i64)pc->selfLeaf.size();
    if (nself) { ssc_copy_indexed_kernel<<<grid_for(nself, 256), 256, 0, pc->stream>>>(nself, pc->d_selfRoot, pc->d_selfLeaf, yl, y, 1); pc->launches++; }
    if (pc->neigh.empty()) return 0;
    const i64 ns = (i64)pc->sendRoot.size(), nr = (i64)pc->recvLeaf.size();
    if (nr) { ssc_gather_kernel<<<grid_for(nr, 256), 256, 0, pc->stream>>>(nr, pc->d_recvLeaf, yl, pc->d_recvbuf); pc->launches++; }
    NCCL_OK(g_nccl.GroupStart());
    for (size_t k = 0; k < pc->neigh.size(); ++k) {
        const i64 s0 = pc->sendOff[k], s1 = pc->sendOff[k + 1], r0 = pc->recvOff[k], r1 = pc->recvOff[k + 1];
        if (r1 > r0) NCCL_OK(g_nccl.Send(pc->d_recvbuf + r0, (size_t)(r1 - r0), kNcclFloat64, pc->neigh[k], pc->comm, pc->stream));
        if (s1 > s0) NCCL_OK(g_nccl.Recv(pc->d_sendbuf + s0, (size_t)(s1 - s0), kNcclFloat64, pc->neigh[k], pc->comm, pc->stream));
    }
    NCCL_OK(g_nccl.GroupEnd());
    if (ns) { ssc_scatter_kernel<<<grid_for(ns, 256), 256, 0, pc->stream>>>(ns, pc->d_sendRoot, pc->d_sendbuf, y, 1); pc->launches++; }
    return 0;
}


// ---------------------------------------------------------------------------------------------
// NVLink peer exchange: the neighbours' local vectors are mapped with CUDA IPC; the halo fill becomes direct stores,
// the overlap sum becomes fp64 reductions issued by the patch kernel itself, and two neighbour barriers per apply
// replace the NCCL send/recv pairs.
// ---------------------------------------------------------------------------------------------
static int peer_barrier(PC *pc)
{
    if (pc->neigh.empty()) return 0;
    pc->epoch++;
    ssc_peer_barrier_kernel<<<1, 32, 0, pc->stream>>>((int)pc->neigh.size(), pc->h_flags, pc->d_flags, pc->epoch, pc->d_timeout);
    pc->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}
static int peer_check(PC *pc)
{
    if (!pc->p2p) return 0;
    int t = 0;
    CUDA_OK(cudaMemcpy(&t, pc->d_timeout, sizeof(int), cudaMemcpyDeviceToHost));
    if (t) return ssc_set_error(SSC_ERR_LIB, "a neighbour rank did not reach the peer barrier within the time limit");
    return 0;
}
__global__ void ssc_peer_pushadd_kernel(ll nghost, int nowned, const double *__restrict__ yl, PeerScatter ps)
{
    for (ll s = (ll)blockIdx.x * blockDim.x + threadIdx.x; s < nghost; s += (ll)gridDim.x * blockDim.x)
        atomicAdd(ps.yl[ps.ghostPeer[s]] + ps.ghostRemote[s], yl[nowned + s]);
}

extern "C" int SSCPatchPeerExport(SSCPatchPC pc, char *handles)
{
    CHK(set_device(pc));
    if (!pc->sf_set) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "set the star forest before exporting peer handles");
    const i64 nghost = pc->nlocal - pc->nowned;
    if (!pc->d_xg) { CHK(dev_alloc(&pc->d_xg, nghost)); CUDA_OK(cudaMemset(pc->d_xg, 0, (size_t)std::max<i64>(nghost, 1) * sizeof(double))); }
    if (!pc->d_yin) { CHK(dev_alloc(&pc->d_yin, pc->nowned)); CUDA_OK(cudaMemset(pc->d_yin, 0, (size_t)std::max<i64>(pc->nowned, 1) * sizeof(double))); }
    if (!pc->d_flags) { CHK(dev_alloc(&pc->d_flags, 16)); CUDA_OK(cudaMemset(pc->d_flags, 0, 16 * sizeof(unsigned long long))); }
    if (!pc->d_timeout) { CHK(dev_alloc(&pc->d_timeout, 1)); CUDA_OK(cudaMemset(pc->d_timeout, 0, sizeof(int))); }
    cudaIpcMemHandle_t h[3];
    CUDA_OK(cudaIpcGetMemHandle(&h[0], pc->d_xg));
    CUDA_OK(cudaIpcGetMemHandle(&h[1], pc->d_yin));
    CUDA_OK(cudaIpcGetMemHandle(&h[2], pc->d_flags));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handles, h, 192);
    return 0;
}

extern "C" int SSCPatchPeerConnect(SSCPatchPC pc, const char *neighHandles, const int *slotAtNeigh, const SSCInt *remoteLeaf, const SSCInt *remoteRoot)
{
    CHK(set_device(pc));
    const size_t nn = pc->neigh.size();
    if (nn > 16) return ssc_set_error(SSC_ERR_SUP, "at most 16 neighbour ranks are supported by the peer exchange");
    if (!pc->d_flags) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "call SSCPatchPeerExport first");
    for (size_t i = 0; i < pc->selfLeaf.size(); ++i)
        if (pc->selfLeaf[i] != (i64)i || pc->selfRoot[i] != (i64)i) return ssc_set_error(SSC_ERR_SUP, "the peer exchange needs owned dofs numbered first and identically in the local and owned vectors");
    if ((i64)pc->selfLeaf.size() != pc->nowned) return ssc_set_error(SSC_ERR_SUP, "the peer exchange needs every owned dof in the local vector");
    const i64 nghost = pc->nlocal - pc->nowned;
    std::vector<int> ghostPeer((size_t)std::max<i64>(nghost, 1), -1), ghostRemote((size_t)std::max<i64>(nghost, 1), 0);
    std::vector<int> sendPeer(std::max<size_t>(pc->sendRoot.size(), 1), 0), rleaf(std::max<size_t>(pc->sendRoot.size(), 1), 0);
    for (size_t k = 0; k < nn; ++k) {
        cudaIpcMemHandle_t h[3];
        memcpy(h, neighHandles + 192 * k, 192);
        void *p[3] = {nullptr, nullptr, nullptr};
        for (int j = 0; j < 3; ++j) {
            cudaError_t err = cudaIpcOpenMemHandle(&p[j], h[j], cudaIpcMemLazyEnablePeerAccess);
            if (err != cudaSuccess) { cudaGetLastError(); return ssc_set_error(SSC_ERR_GPU, "cudaIpcOpenMemHandle failed for neighbour rank %d: %s", pc->neigh[k], cudaGetErrorString(err)); }
            pc->ipc_opened.push_back(p[j]);
        }
        pc->h_push.xl[k] = (double *)p[0];
        pc->h_ps.yl[k] = (double *)p[1];
        pc->h_flags.flags[k] = (unsigned long long *)p[2];
        pc->h_flags.slot[k] = slotAtNeigh[k];
        for (i64 i = pc->recvOff[k]; i < pc->recvOff[k + 1]; ++i) {
            const i64 s = pc->recvLeaf[i] - pc->nowned;
            if (s < 0 || s >= nghost) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "receive slot is not a ghost slot");
            ghostPeer[(size_t)s] = (int)k; ghostRemote[(size_t)s] = (int)remoteRoot[i];
        }
        for (i64 i = pc->sendOff[k]; i < pc->sendOff[k + 1]; ++i) {
            if (remoteLeaf[i] < 0) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "remoteLeaf must be the neighbour's ghost slot number (local index minus its owned count)");
            sendPeer[(size_t)i] = (int)k; rleaf[(size_t)i] = (int)remoteLeaf[i];
        }
    }
    for (i64 s = 0; s < nghost; ++s) if (ghostPeer[(size_t)s] < 0) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "ghost slot %lld has no owner among the neighbours", (long long)s);
    auto up = [&](const std::vector<int> &v, int **d) -> int {
        CHK(dev_alloc(d, (i64)v.size()));
        CUDA_OK(cudaMemcpy(*d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
        return 0;
    };
    CHK(up(ghostPeer, &pc->d_ghostPeer)); CHK(up(ghostRemote, &pc->d_ghostRemote));
    CHK(up(sendPeer, &pc->d_sendPeer)); CHK(up(rleaf, &pc->d_remoteLeaf));
    pc->h_ps.ghostPeer = pc->d_ghostPeer; pc->h_ps.ghostRemote = pc->d_ghostRemote; pc->h_ps.nowned = (int)pc->nowned;
    pc->h_ps.xg = pc->d_xg;
    {
        std::vector<i64> u(pc->sendRoot);
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
        pc->niface = (i64)u.size();
        CHK(upload_i32(pc, u, &pc->d_iface));
    }
    CHK(dev_alloc(&pc->d_ps, 1));
    CUDA_OK(cudaMemcpy(pc->d_ps, &pc->h_ps, sizeof(PeerScatter), cudaMemcpyHostToDevice));
    pc->p2p = true;
    return 0;
}

// halo fill through peer stores + barrier: neighbours' ghost arrays receive this rank's interface values
static int peer_bcast(PC *pc, const double *x)
{
    const i64 ns = (i64)pc->sendRoot.size();
    if (ns) { ssc_peer_push_kernel<<<grid_for(ns, 256), 256, 0, pc->stream>>>(ns, pc->d_sendRoot, pc->d_sendPeer, pc->d_remoteLeaf, x, pc->h_push); pc->launches++; }
    return peer_barrier(pc);
}
// after the patch kernels: wait for the neighbours, then fold their contributions to this rank's interface dofs into y
static int peer_fold(PC *pc, double *y)
{
    CHK(peer_barrier(pc));
    if (pc->niface) { ssc_peer_fold_kernel<<<grid_for(pc->niface, 256), 256, 0, pc->stream>>>(pc->niface, pc->d_iface, pc->d_yin, y); pc->launches++; }
    CUDA_OK(cudaGetLastError());
    return 0;
}
// overlap sum of a local (owned + ghost) vector that was filled WITHOUT the fused scatter (set-up only: dof multiplicities)
static int peer_reduce(PC *pc, const double *yl, double *y)
{
    CHK(peer_barrier(pc));
    const i64 ng = pc->nlocal - pc->nowned;
    if (ng) { ssc_peer_pushadd_kernel<<<grid_for(ng, 256), 256, 0, pc->stream>>>(ng, (int)pc->nowned, yl, pc->h_ps); pc->launches++; }
    CUDA_OK(cudaMemcpyAsync(y, yl, (size_t)pc->nowned * sizeof(double), cudaMemcpyDeviceToDevice, pc->stream));
    return peer_fold(pc, y);
}

// Host -> device copy of a large pageable array: worker threads copy slices into a small ring of pinned buffers while
// the previous buffer is in flight over PCIe (a plain cudaMemcpy from pageable memory is bound by one host thread).
static int staged_upload(PC *pc, void *dst, const void *src, size_t bytes)
{
    const size_t chunk = (size_t)128 << 20;
    if (bytes < 2 * chunk) { CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, pc->up_stream)); return 0; }
    const int R = 3;
    if (!pc->stage[0]) {
        for (int r = 0; r < R; ++r) {
            cudaError_t err = cudaMallocHost(&pc->stage[r], chunk);
            if (err != cudaSuccess) {       // no pinned memory to spare: fall back to the driver's own staging
                cudaGetLastError();
                for (int q = 0; q < r; ++q) { cudaFreeHost(pc->stage[q]); pc->stage[q] = nullptr; }
                CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, pc->up_stream));
                return 0;
            }
            CUDA_OK(cudaEventCreateWithFlags(&pc->stage_ev[r], cudaEventDisableTiming));
        }
    }
    const int T = (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
    size_t off = 0;
    for (int it = 0; off < bytes; ++it) {
        const int r = it % R;
        const size_t len = std::min(chunk, bytes - off);
        if (pc->stage_busy[r]) CUDA_OK(cudaEventSynchronize(pc->stage_ev[r]));   // also across calls: the previous array's tail may still be in flight
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) {
            const size_t a = len * t / T, b = len * (t + 1) / T;
            th.emplace_back([=]() { memcpy((char *)pc->stage[r] + a, (const char *)src + off + a, b - a); });
        }
        for (auto &x : th) x.join();
        CUDA_OK(cudaMemcpyAsync((char *)dst + off, pc->stage[r], len, cudaMemcpyHostToDevice, pc->up_stream));
        CUDA_OK(cudaEventRecord(pc->stage_ev[r], pc->up_stream));
        pc->stage_busy[r] = true;
        off += len;
    }
    return 0;
}

static int upload_csr(PC *pc)
{
    if (!pc->csr_set) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "Must call SSCPatchSetOperatorCSR() to set the operator");   // cf. ssc/libssc.c:1131-1133
    if (pc->csr_nrows < pc->nlocal) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "operator has %lld rows but the local vector has %lld dofs", (long long)pc->csr_nrows, (long long)pc->nlocal);
    if (pc->csr_space == SSC_MEM_DEVICE) {
        if (pc->csr_owned) { dev_free(pc->d_rowptr); dev_free(pc->d_colidx); dev_free(pc->d_vals); pc->csr_owned = false; }
        pc->d_rowptr = const_cast<i64 *>(pc->csr_rowptr); pc->d_colidx = const_cast<int *>(pc->csr_colidx); pc->d_vals = const_cast<double *>(pc->csr_vals);
        return 0;
    }
    const i64 nnz = pc->csr_rowptr[pc->csr_nrows];
    // the device copy of the previous set-up is reused when the shape is the same (PCSetUp after a change of values, ssc/patch.py:269-270)
    if (pc->csr_owned && (pc->csr_dev_rows != pc->csr_nrows || pc->csr_dev_nnz != nnz)) { dev_free(pc->d_rowptr); dev_free(pc->d_colidx); dev_free(pc->d_vals); pc->csr_owned = false; }
    pc->csr_nnz = nnz;
    if (!pc->csr_owned) {
        CHK(dev_alloc(&pc->d_rowptr, pc->csr_nrows + 1));
        CHK(dev_alloc(&pc->d_colidx, nnz));
        CHK(dev_alloc(&pc->d_vals, nnz));
        pc->csr_owned = true; pc->csr_dev_rows = pc->csr_nrows; pc->csr_dev_nnz = nnz;
    }
    CUDA_OK(cudaMemcpyAsync(pc->d_rowptr, pc->csr_rowptr, (size_t)(pc->csr_nrows + 1) * sizeof(i64), cudaMemcpyHostToDevice, pc->up_stream));
    CHK(staged_upload(pc, pc->d_colidx, pc->csr_colidx, (size_t)nnz * sizeof(int)));
    CHK(staged_upload(pc, pc->d_vals, pc->csr_vals, (size_t)nnz * sizeof(double)));
    return 0;
}

// device copies (int32) of the cell-node maps, `rows` cells each: independent of the patches, so the uploader thread
// sends them while the host still builds the patch lists
static int upload_cell_node_maps(PC *pc, i64 rows)
{
    const Discretisation &D = pc->disc;
    if (D.nsub > 8) return ssc_set_error(SSC_ERR_SUP, "element matrices: at most 8 subspaces");
    for (int *&q : pc->d_cnm) dev_free(q);
    pc->d_cnm.clear();
    ElemSpaces &E = pc->espaces;
    memset(&E, 0, sizeof(E));
    E.nsub = (int)D.nsub; E.t = (int)D.totalDofsPerCell;
    RawArray<int> narrow;
    for (i64 k = 0; k < D.nsub; ++k) {
        const i64 len = rows * D.nodesPerCell[k];
        narrow.resize((size_t)len);
        for (i64 i = 0; i < len; ++i) narrow[(size_t)i] = (int)D.cellNodeMap[k][i];
        int *d = nullptr;
        CHK(dev_alloc(&d, len));
        pc->d_cnm.push_back(d);
        CUDA_OK(cudaMemcpy(d, narrow.data(), (size_t)len * sizeof(int), cudaMemcpyHostToDevice));
        E.cnm[k] = d; E.bs[k] = (int)D.bs[k]; E.npc[k] = (int)D.nodesPerCell[k]; E.subOff[k] = (int)D.subspaceOffsets[k];
        E.tOff[k + 1] = E.tOff[k] + E.npc[k] * E.bs[k];
    }
    if (E.tOff[D.nsub] != E.t) return ssc_set_error(SSC_ERR_ARG_INCOMP, "subspaces add up to %d dofs per cell, expected %d", E.tOff[D.nsub], E.t);
    pc->elem_maps_ready = true; pc->elem_map_rows = rows;
    return 0;
}

static int upload_elements(PC *pc)
{
    if (pc->multiplicative) return ssc_set_error(SSC_ERR_SUP, "multiplicative sweeps update the residual with the assembled operator: use SSCPatchSetOperatorCSR()");
    if (!pc->elem_maps_ready && pc->disc.set && pc->disc.cn_set && pc->plex.set && !pc->direct_patches) {
        // rows of the cell-node maps = cells in the cell numbering
        const i64 cS = pc->plex.depthStart[pc->plex.dim], cE = pc->plex.depthEnd[pc->plex.dim];
        i64 rows = 0;
        for (i64 c = cS; c < cE; ++c) if (pc->disc.cnDof[(size_t)c] > 0) rows = std::max(rows, pc->disc.cnOff[(size_t)c] + 1);
        if (rows > 0) CHK(upload_cell_node_maps(pc, rows));
    }
    const i64 t2 = pc->elem_t * pc->elem_t;
    const i64 len = pc->elem_stride == 0 ? t2 : (pc->elem_ncells - 1) * pc->elem_stride + t2;
    if (pc->elem_owned && (pc->elem_space == SSC_MEM_DEVICE || pc->elem_dev_len != len)) { dev_free(pc->d_elemK); pc->elem_owned = false; }
    if (pc->elem_space == SSC_MEM_DEVICE) { pc->d_elemK = const_cast<double *>(pc->elem_K); return 0; }
    if (!pc->elem_owned) {                                      // else: the buffer of the previous set-up has the right size
        CHK(dev_alloc(&pc->d_elemK, len));
        pc->elem_owned = true; pc->elem_dev_len = len;
    }
    CHK(staged_upload(pc, pc->d_elemK, pc->elem_K, (size_t)len * sizeof(double)));
    return 0;
}

static int upload_operator(PC *pc)
{
    if (pc->op_released) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "the operator arrays were released (SSCPatchReleaseOperator): set the operator again before SSCPatchSetUp()");
    return pc->elem_set ? upload_elements(pc) : upload_csr(pc);
}

// cells of every patch in size-sorted order + the cell-node maps: what the assemble kernel turns into the `dofs` table on the fly
static int upload_element_lists(PC *pc)
{
    const PatchLists &L = pc->lists;
    const Discretisation &D = pc->disc;
    if (pc->direct_patches || !D.set || L.cells.empty()) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "element matrices need patches built from the mesh and the discretisation (cells per patch, cell-node maps)");
    if (D.nsub > 8) return ssc_set_error(SSC_ERR_SUP, "element matrices: at most 8 subspaces");
    if (pc->elem_t != D.totalDofsPerCell) return ssc_set_error(SSC_ERR_ARG_INCOMP, "element matrices have %lld dofs per cell but the discretisation has %lld", (long long)pc->elem_t, (long long)D.totalDofsPerCell);
    i64 maxcell = -1;
    for (i64 c : L.cells) maxcell = std::max(maxcell, c);
    if (pc->elem_stride != 0 && maxcell >= pc->elem_ncells) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "patches use cell %lld but only %lld element matrices were given", (long long)maxcell, (long long)pc->elem_ncells);
    if (pc->elem_lists_ready) return 0;
    const i64 ns = (i64)pc->sorted2orig.size();
    std::vector<i64> off((size_t)ns + 1, 0);
    std::vector<int> cells;
    cells.reserve(L.cells.size());
    for (i64 q = 0; q < ns; ++q) {
        const i64 p = pc->sorted2orig[(size_t)q];
        for (i64 k = L.cellOff[p]; k < L.cellOff[p + 1]; ++k) cells.push_back((int)L.cells[(size_t)k]);
        off[(size_t)q + 1] = (i64)cells.size();
    }
    CHK(dev_alloc(&pc->d_pcelloff, ns + 1));
    CHK(dev_alloc(&pc->d_pcells, std::max<i64>(1, (i64)cells.size())));
    CUDA_OK(cudaMemcpyAsync(pc->d_pcelloff, off.data(), (size_t)(ns + 1) * sizeof(i64), cudaMemcpyHostToDevice, pc->stream));
    CUDA_OK(cudaMemcpyAsync(pc->d_pcells, cells.data(), cells.size() * sizeof(int), cudaMemcpyHostToDevice, pc->stream));
    if (!pc->elem_maps_ready || pc->elem_map_rows <= maxcell) CHK(upload_cell_node_maps(pc, maxcell + 1));
    CUDA_OK(cudaStreamSynchronize(pc->stream));
    pc->elem_lists_ready = true;
    return 0;
}

static inline bool pivot_wanted(const PC *pc) { return pc->sub_pc == SSC_SUB_LU; }

static size_t invert_smem_bytes(int n, bool shared)
{
    const size_t ld = (size_t)(n | 1);
    return (2 * (size_t)n + 32 + ((34 + (size_t)n + 1) / 2 + 1) + (shared ? (size_t)n * ld : 0)) * sizeof(double);
}

// extract + factor for patches [q0, q1) of class c into `blocks` (class-relative addressing through moff)
static int extract_and_invert(PC *pc, const SizeClass &c, i64 q0, i64 q1, double *blocks, int *fail, bool time_it)
{
    const i64 cnt = q1 - q0;
    if (cnt <= 0) return 0;
    const int n = c.n;
    int H = 32;
    while (H < 2 * n) H <<= 1;
    const int bd = 128, nw = bd / 32;
    const size_t smem_e = (size_t)H * 2 * sizeof(int) + (size_t)n * sizeof(i64) + (size_t)((n + 2) & ~1) * sizeof(int) + (size_t)2 * nw * n * sizeof(double);
    if (smem_e > 200 * 1024) return ssc_set_error(SSC_ERR_SUP, "patch with %d dofs is too large for the extract kernel", n);
    if (smem_e > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(ssc_extract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e));
    if (time_it) CUDA_OK(cudaEventRecord(pc->ev0, pc->stream));
    if (pc->elem_set) {
        const int tpad = ((int)pc->elem_t + 1) & ~1;
        const size_t base = ((size_t)H * 2 + 2 * (size_t)tpad) * sizeof(int);
        const size_t accb = (size_t)n * c.ld * sizeof(double);
        const int acc_shared = pc->asm_shared && base + accb <= 200 * 1024;
        const size_t smem_a = base + (acc_shared ? accb : 0);
        if (smem_a > 200 * 1024) return ssc_set_error(SSC_ERR_SUP, "patch with %d dofs is too large for the assemble kernel", n);
        if (smem_a > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(ssc_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
        const int grid_a = (int)std::min<i64>(cnt, (i64)pc->sm_count * pc->asm_ctas_per_sm);
        ssc_assemble_kernel<<<grid_a, pc->asm_threads, smem_a, pc->stream>>>(pc->espaces, pc->d_elemK, (ll)pc->elem_stride, (const ll *)pc->d_pcelloff + q0, pc->d_pcells,
                                                               pc->d_gtol + c.goff + (q0 - c.first) * n, blocks, n, c.ld, c.stride, cnt, H, acc_shared);
    } else {
    const int grid_e = (int)std::min<i64>(cnt, (i64)pc->sm_count * 64);
    ssc_extract_kernel<<<grid_e, bd, smem_e, pc->stream>>>((const ll *)pc->d_rowptr, pc->d_colidx, pc->d_vals, pc->d_gtol + c.goff + (q0 - c.first) * n,
                                                          blocks, n, c.ld, c.stride, cnt, H);
    }
    pc->launches++;
    CUDA_OK(cudaGetLastError());
    if (time_it) {
        CUDA_OK(cudaEventRecord(pc->ev1, pc->stream));
        CUDA_OK(cudaEventSynchronize(pc->ev1));
        float ms = 0; CUDA_OK(cudaEventElapsedTime(&ms, pc->ev0, pc->ev1));
        pc->event_ms["PCPATCHComputeOp"] += ms;
    }
    if (pc->keep_operators && pc->d_blocks0 && blocks == pc->d_blocks + c.moff)
        CUDA_OK(cudaMemcpyAsync(pc->d_blocks0 + c.moff, blocks, (size_t)cnt * c.stride * sizeof(double), cudaMemcpyDeviceToDevice, pc->stream));
    // factor
    const bool shared = n <= 168;
    const size_t smem_f = invert_smem_bytes(n, shared);
    const int bdf = !shared ? 1024 : (n <= 32 ? 128 : (n <= 64 ? 256 : 512));
    const bool pivot = pc->sub_pc == SSC_SUB_LU;
    if (time_it) CUDA_OK(cudaEventRecord(pc->ev0, pc->stream));
    const int grid_f = (int)std::min<i64>(cnt, (i64)pc->sm_count * 16);
#define LAUNCH_INV(P, S)                                                                                              \
    do {                                                                                                              \
        if (smem_f > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(ssc_invert_kernel<P, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f)); \
        ssc_invert_kernel<P, S><<<grid_f, bdf, smem_f, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail);             \
    } while (0)
    const int fvar = pc->factor_variant >= 0 ? pc->factor_variant : ((n >= 100 && n <= 128) ? (pivot_wanted(pc) ? 4 : 3) : 1);
    if (n > 128 && fvar >= 1) {
        // blocked Gauss-Jordan with DMMA updates, block resident in L2
        const int npad = (n + 7) & ~7, WP = ((npad + 15) & ~15) + 4;
        auto smem_for = [&](int nb) { return ((size_t)npad * (nb + 1) + (size_t)nb * WP + nb + 32) * sizeof(double) + (size_t)(34 + 2 * n + 2) * sizeof(int); };
        const int nb = (pc->factor_panel == 32 && smem_for(32) <= 220 * 1024) ? 32 : 16;
        const size_t smem_b = smem_for(nb);
        if (smem_b > 220 * 1024) return ssc_set_error(SSC_ERR_SUP, "patch with %d dofs is too large for the blocked factor kernel", n);
        const int per_sm = smem_b <= 110 * 1024 ? 2 : 1;
        const int grid_b = (int)std::min<i64>(cnt, pc->factor_ctas > 0 ? (i64)pc->factor_ctas : (i64)pc->sm_count * per_sm);
#define LAUNCH_BLK(P, NBV)                                                                                                     \
    do {                                                                                                                       \
        CUDA_OK(cudaFuncSetAttribute(ssc_invert_blocked_kernel<P, NBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b)); \
        ssc_invert_blocked_kernel<P, NBV><<<grid_b, 512, smem_b, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail);            \
    } while (0)
        if (pivot && nb == 32) LAUNCH_BLK(true, 32);
        else if (pivot) LAUNCH_BLK(true, 16);
        else if (nb == 32) LAUNCH_BLK(false, 32);
        else LAUNCH_BLK(false, 16);
#undef LAUNCH_BLK
    }
    else if (n <= 128 && fvar == 1 && pc->factor_tile == 8) {
        // register-resident 8 x 8 tiles: (n/8)^2 threads; needs tid < n for the permutation arrays -> at least n threads
        const int TC = (n + 7) / 8;
        const int threads = std::max((TC * TC + 31) / 32 * 32, (n + 31) / 32 * 32);
        if (pivot) ssc_invert_reg8_kernel<true><<<grid_f, threads, 0, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail, TC);
        else ssc_invert_reg8_kernel<false><<<grid_f, threads, 0, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail, TC);
    }
    else if (n <= 128 && fvar == 4) {
        // blocked look-ahead Gauss-Jordan, searcher decoupled by two panels
        const int TC = (n + 4) / 5, TR = (n + 7) / 8;
        const int W = (TR * TC + 31) / 32 * 32;
        const int remap = (W == 416 && pc->factor_remap) ? 1 : 0;
        const int threads = remap ? 512 : W + 32;
        if (pivot) ssc_invert_panel2_kernel<true><<<grid_f, threads, 0, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail, TR, TC, W, remap);
        else ssc_invert_panel2_kernel<false><<<grid_f, threads, 0, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail, TR, TC, W, remap);
    }
    else if (n <= 128 && fvar == 3) {
        // blocked look-ahead Gauss-Jordan: panels of 5 columns, (n/8) x (n/5) worker threads + one searcher warp
        const int TC = (n + 4) / 5, TR = (n + 7) / 8;
        const int W = (TR * TC + 31) / 32 * 32;
        const int NS = pc->factor_searchers == 2 ? 2 : 1;
        const int remap = (NS == 1 && W == 416 && pc->factor_remap) ? 1 : 0;
        const int threads = remap ? 512 : W + 32 * NS;
        if (NS == 2) {
            if (pivot) ssc_invert_panel_kernel<true, 2><<<grid_f, threads, 0, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail, TR, TC, W, 0);
            else ssc_invert_panel_kernel<false, 2><<<grid_f, threads, 0, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail, TR, TC, W, 0);
        } else {
            if (pivot) ssc_invert_panel_kernel<true, 1><<<grid_f, threads, 0, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail, TR, TC, W, remap);
            else ssc_invert_panel_kernel<false, 1><<<grid_f, threads, 0, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail, TR, TC, W, remap);
        }
    }
    else if (n <= 128 && fvar == 2) {
        // register-resident tiles with a look-ahead pivot search: (n/8) x (n/5) worker threads + one searcher warp
        const int TC = (n + 4) / 5, TR = (n + 7) / 8;
        const int W = (TR * TC + 31) / 32 * 32;
        if (pivot) ssc_invert_la_kernel<true><<<grid_f, W + 32, 0, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail, TC, W);
        else ssc_invert_la_kernel<false><<<grid_f, W + 32, 0, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail, TC, W);
    }
    else if (n <= 128 && fvar == 1) {
        // register-resident tiles: (n/8) x (n/4) threads
        const int TC = (n + 3) / 4, TR = (n + 7) / 8;
        const int threads = (TR * TC + 31) / 32 * 32;
        if (pivot) ssc_invert_reg_kernel<true><<<grid_f, threads, 0, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail, TC);
        else ssc_invert_reg_kernel<false><<<grid_f, threads, 0, pc->stream>>>(blocks, n, c.ld, c.stride, cnt, fail, TC);
    }
    else if (pivot && shared) LAUNCH_INV(true, true);
    else if (pivot) LAUNCH_INV(true, false);
    else if (shared) LAUNCH_INV(false, true);
    else LAUNCH_INV(false, false);
#undef LAUNCH_INV
    pc->launches++;
    CUDA_OK(cudaGetLastError());
    if (time_it) {
        CUDA_OK(cudaEventRecord(pc->ev1, pc->stream));
        CUDA_OK(cudaEventSynchronize(pc->ev1));
        float ms = 0; CUDA_OK(cudaEventElapsedTime(&ms, pc->ev0, pc->ev1));
        pc->event_ms["PCPATCHSolve"] += ms;
    }
    return 0;
}

static int build_host_pipeline(PC *pc);
static int build_peer_pipeline(PC *pc);
static int compute_weights(PC *pc)
{
    // ssc/libssc.c:1254-1284: scatter ones through every patch, reduce to owners, reciprocal
    dev_free(pc->d_w);
    CHK(dev_alloc(&pc->d_w, pc->nowned));
    double *tmp_local = nullptr;
    if (pc->sf_set && pc->p2p) CHK(dev_alloc(&tmp_local, pc->nlocal));
    double *ml = pc->sf_set ? (pc->p2p ? tmp_local : pc->d_yl) : pc->d_w;
    CUDA_OK(cudaMemsetAsync(ml, 0, (size_t)pc->nlocal * sizeof(double), pc->stream));
    if (pc->sum_n) { ssc_count_kernel<<<grid_for(pc->sum_n, 256), 256, 0, pc->stream>>>(pc->d_gtol, pc->sum_n, ml); pc->launches++; }
    if (pc->sf_set && pc->p2p) CHK(peer_reduce(pc, ml, pc->d_w));
    else if (pc->sf_set) CHK(sf_reduce(pc, ml, pc->d_w));
    ssc_reciprocal_kernel<<<grid_for(pc->nowned, 256), 256, 0, pc->stream>>>(pc->nowned, pc->d_w);
    pc->launches++;
    CUDA_OK(cudaGetLastError());
    if (tmp_local) { CUDA_OK(cudaStreamSynchronize(pc->stream)); cudaFree(tmp_local); }
    return 0;
}

extern "C" int SSCPatchSetUp(SSCPatchPC pc)
{
    if (!pc) return ssc_set_error(SSC_ERR_ARG_NULL, "null PC");
    CHK(set_device(pc));
    if (pc->sub_ksp_type != "preonly") return ssc_set_error(SSC_ERR_SUP, "only preonly patch solves are supported");
    const bool first = !pc->device_ready;
    auto tick = std::chrono::steady_clock::now();
    auto lap = [&](const char *name) {
        auto now = std::chrono::steady_clock::now();
        pc->event_ms[name] = std::chrono::duration<double, std::milli>(now - tick).count();
        tick = now;
    };
    // the operator travels to the device on its own stream while the host builds the patches
    int up_rc = 0;
    std::string up_msg;
    std::thread uploader([&]() {
        cudaSetDevice(pc->device);
        up_rc = upload_operator(pc);
        if (up_rc) up_msg = SSCGetLastError();
    });
    int list_rc = first ? build_device_lists(pc) : 0;
    std::string list_msg = list_rc ? SSCGetLastError() : "";
    lap("SSCSetupLists");                                               // patch construction + size classes + dof-list upload
    uploader.join();
    if (list_rc) return ssc_set_error(list_rc, "%s", list_msg.c_str());
    if (up_rc) return ssc_set_error(up_rc, "%s", up_msg.c_str());
    CUDA_OK(cudaStreamSynchronize(pc->up_stream));
    if (pc->elem_set) CHK(upload_element_lists(pc));
    lap("SSCSetupUploadOperator");                                      // what is left of the upload after the patches are built
    if (first && pc->partition_of_unity) CHK(compute_weights(pc));      // only on the first setup, ssc/libssc.c:1254
    pc->event_ms["PCPATCHComputeOp"] = 0; pc->event_ms["PCPATCHSolve"] = 0;
    if (pc->save_operators) {
        if (!pc->d_blocks) CHK(dev_alloc(&pc->d_blocks, pc->block_elems));
        lap("SSCSetupAllocBlocks");
        if (pc->keep_operators && !pc->d_blocks0) CHK(dev_alloc(&pc->d_blocks0, pc->block_elems));
        for (const SizeClass &c : pc->classes)
            CHK(extract_and_invert(pc, c, c.first, c.first + c.count, pc->d_blocks + c.moff, pc->d_fail + c.first, true));
        if (pc->symmetric) {
            if (pc->sub_pc != SSC_SUB_CHOLESKY) return ssc_set_error(SSC_ERR_ARG_INCOMP, "ssc_symmetric_storage keeps half of a symmetric inverse: it needs SPD blocks, i.e. -sub_pc_type cholesky");
            if (!pc->d_packed) CHK(dev_alloc(&pc->d_packed, pc->packed_elems));
            for (const SizeClass &c : pc->classes) {
                ssc_pack_sym_kernel<<<(unsigned)std::min<i64>(c.count, (i64)pc->sm_count * 32), 256, 0, pc->stream>>>(pc->d_blocks + c.moff, c.n, c.ld, c.stride, pc->d_packed + c.poff, c.pstride, c.count);
                pc->launches++;
            }
            CUDA_OK(cudaGetLastError());
        }
        CUDA_OK(cudaStreamSynchronize(pc->stream));
        if (pc->symmetric) dev_free(pc->d_blocks);          // only the packed triangles stay resident
        // per-patch failure flags: the reference would surface KSP_DIVERGED_PCSETUP_FAILED (dead code ssc/libssc.c:1439-1442)
        std::vector<int> fail(pc->sorted2orig.size());
        if (!fail.empty()) CUDA_OK(cudaMemcpy(fail.data(), pc->d_fail, fail.size() * sizeof(int), cudaMemcpyDeviceToHost));
        pc->failed = 0;
        i64 firstbad = -1;
        pc->failed_list.clear();
        for (size_t q = 0; q < fail.size(); ++q) if (fail[q]) { if (firstbad < 0) firstbad = pc->sorted2orig[q]; pc->failed++; pc->failed_list.push_back(pc->sorted2orig[q]); }
        lap("SSCSetupKernels");                                     // extract / assemble + factor + failure flags
        // The device copy of the operator is only dropped when memory is tight: cudaFree of the 10.65 GB CSR of the 64^3 Q3
        // case takes 0.3 s (a third of the set-up), and the next set-up of the same shape reuses the buffers.
        {
            // (own bookkeeping instead of cudaMemGetInfo, which itself took up to 70 ms here)
            const double held = 8.0 * (double)pc->block_elems * (pc->keep_operators ? 2.0 : 1.0) + 12.0 * (double)pc->csr_nnz + 8.0 * (double)pc->elem_dev_len + 4.0 * (double)pc->sum_n;
            const bool tight = held > 0.6 * (double)pc->total_mem;
            if (tight && pc->csr_owned && !pc->multiplicative) { dev_free(pc->d_rowptr); dev_free(pc->d_colidx); dev_free(pc->d_vals); pc->csr_owned = false; }
            if (tight && pc->elem_owned) { dev_free(pc->d_elemK); pc->elem_owned = false; }
        }
        lap("SSCSetupFreeOperator");
        if (pc->failed) return ssc_set_error(SSC_ERR_MAT_LU_ZRPVT, "%lld patch operators could not be factorised (first: patch %lld)%s", (long long)pc->failed,
                                             (long long)firstbad, pc->sub_pc == SSC_SUB_CHOLESKY ? "; sub_pc_type cholesky needs SPD blocks" : "");
    } else {
        if (pc->symmetric) return ssc_set_error(SSC_ERR_ARG_INCOMP, "ssc_symmetric_storage needs saved patch operators");
        CUDA_OK(cudaStreamSynchronize(pc->stream));
    }
    pc->pipe_tried = false;      // the slab schedule of host-space applies is built on first use
    pc->setup_done = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// apply
// ---------------------------------------------------------------------------------------------
static int launch_apply_class(PC *pc, const SizeClass &c, i64 q0, i64 q1, const double *blocks_override, const double *xl, double *yl, double scale)
{
    const i64 cnt = q1 - q0;
    if (cnt <= 0) return 0;
    const int *gt = pc->d_gtol + c.goff + (q0 - c.first) * c.n;
    const double *mult = pc->d_mult ? pc->d_mult + q0 : nullptr;
    double *ypatch = (pc->deterministic && !pc->multiplicative) ? pc->d_ypatch + c.goff + (q0 - c.first) * c.n : nullptr;
    if (pc->symmetric && !blocks_override) {
        const int threads = std::min(1024, (c.n + 31) / 32 * 32);
        const i64 per_cta = cnt / std::max(1, pc->sm_count);
        int stages = (int)std::min<size_t>(4, ((size_t)220 * 1024 - 5 * (size_t)(c.n + 32) * 8) / std::max<size_t>(c.pbytes, 16));
        if (pc->sym_persistent && stages >= 2 && per_cta >= 8) {
            const int rp = (c.n + 31) / 32 * 32;
            const int Fp = std::max(1, std::min(4, 1024 / rp));
            const int threads_p = rp * Fp;
            const size_t smem_p = (size_t)stages * c.pbytes + (2 * (size_t)c.n + (size_t)(Fp - 1) * rp) * sizeof(double);
            CUDA_OK(cudaFuncSetAttribute(ssc_apply_sym_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
            ssc_apply_sym_persistent_kernel<<<pc->sm_count, threads_p, smem_p, pc->launch_stream>>>(pc->d_packed + c.poff + (q0 - c.first) * c.pstride, c.pstride, c.pbytes, stages, gt,
                                                                                        c.n, cnt, xl, yl, scale, mult, pc->p2p ? pc->d_ps : nullptr, ypatch);
            pc->launches++;
            CUDA_OK(cudaGetLastError());
            return 0;
        }
        if (pc->sym_bulk && pc->sym_two_chunk && 2 * ((c.n + 31) / 32 * 32) <= 1024) {
            const int rp = (c.n + 31) / 32 * 32;
            int j1 = c.n;                                       // split column: about half of the bytes, 16-byte aligned start
            unsigned bytes0 = c.pbytes;
            if (c.pbytes >= 8192) {
                for (int j = (int)(0.2929 * c.n); j < c.n; ++j) {
                    const i64 cs = (i64)j * c.n - (i64)j * (j - 1) / 2;
                    if ((cs & 1) == 0) { j1 = j; bytes0 = (unsigned)(cs * 8); break; }
                }
            }
            const size_t smem2 = (size_t)c.pbytes + (size_t)(c.n + rp) * sizeof(double);
            if (smem2 > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(ssc_apply_sym2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            ssc_apply_sym2_kernel<<<(unsigned)cnt, 2 * rp, smem2, pc->launch_stream>>>(pc->d_packed + c.poff + (q0 - c.first) * c.pstride, c.pstride, c.pbytes, j1, bytes0, gt, c.n, cnt,
                                                                           xl, yl, scale, mult, pc->p2p ? pc->d_ps : nullptr, ypatch);
            pc->launches++;
            CUDA_OK(cudaGetLastError());
            return 0;
        }
        const size_t smem = (size_t)c.pbytes + (size_t)c.n * sizeof(double);
        if (pc->sym_bulk) {
            if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(ssc_apply_sym_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ssc_apply_sym_kernel<true><<<(unsigned)cnt, threads, smem, pc->launch_stream>>>(pc->d_packed + c.poff + (q0 - c.first) * c.pstride, c.pstride, c.pbytes, gt, c.n, cnt,
                                                                                  xl, yl, scale, mult, pc->p2p ? pc->d_ps : nullptr, ypatch);
            pc->launches++;
            CUDA_OK(cudaGetLastError());
            return 0;
        }
        if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(ssc_apply_sym_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ssc_apply_sym_kernel<false><<<(unsigned)cnt, threads, smem, pc->launch_stream>>>(pc->d_packed + c.poff + (q0 - c.first) * c.pstride, c.pstride, c.pbytes, gt, c.n, cnt,
                                                                        xl, yl, scale, mult, pc->p2p ? pc->d_ps : nullptr, ypatch);
        pc->launches++;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    const double *blocks = blocks_override ? blocks_override : pc->d_blocks + c.moff + (q0 - c.first) * c.stride;
    const int threads = std::max(32, (c.tp * c.ppb + 31) / 32 * 32);
    const size_t smem = (size_t)c.ppb * c.n * sizeof(double);
    const i64 grid = (cnt + c.ppb - 1) / c.ppb;
#define LAUNCH_APPLY(K)                                                                                                  \
    do {                                                                                                                 \
        if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        K<<<(unsigned)grid, threads, smem, pc->launch_stream>>>(blocks, c.stride, c.ld, gt, c.n, cnt, c.tp, c.ppb, xl, yl, scale, mult, pc->p2p ? pc->d_ps : nullptr, ypatch); \
    } while (0)
    if (pc->apply_variant == 1 && pc->apply_stream_small && c.n <= pc->apply_stream_max_n && c.tp >= c.n && threads <= 256) {
        // persistent, gathers prefetched one and two patches ahead: groups walk the list in lockstep
        const int per_sm = std::max(1, 2048 / threads);
        const i64 ctas = pc->apply_stream_ctas > 0 ? std::min<i64>(grid, pc->apply_stream_ctas) : std::min<i64>(grid, (i64)pc->sm_count * per_sm);
        if (grid >= (i64)pc->apply_stream_min_waves * ctas) {
            const size_t smem2 = 2 * smem;
            if (smem2 > 48 * 1024) CUDA_OK(cudaFuncSetAttribute(ssc_apply_stream_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            ssc_apply_stream_kernel<8><<<(unsigned)ctas, threads, smem2, pc->launch_stream>>>(blocks, c.stride, c.ld, gt, c.n, cnt, c.tp, c.ppb, xl, yl, scale, mult, pc->p2p ? pc->d_ps : nullptr, ypatch);
            pc->launches++;
            CUDA_OK(cudaGetLastError());
            return 0;
        }
    }
    if (pc->apply_variant == 2 && threads <= 256) LAUNCH_APPLY((ssc_apply_pipe_kernel<2, 8>));
    else if (pc->apply_variant == 1 && threads <= 256) LAUNCH_APPLY((ssc_apply_pipe_kernel<1, 8>));
    else LAUNCH_APPLY(ssc_apply_kernel);
#undef LAUNCH_APPLY
    pc->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

// Coloured multiplicative Schwarz (no reference code path: ssc/libssc.c:1568-1572 removed it).
//   y = 0;  for each colour c:  r = x - A y;  y += sum_{p in c} R_p^T B_p^-1 R_p r
// Patches of one colour share no cell, so their corrections commute and one launch per size class applies them all;
// with symmetrise_sweep the colours are then visited in reverse order.
static int apply_multiplicative(PC *pc, const double *x, double *y)
{
    if (!pc->save_operators) return ssc_set_error(SSC_ERR_SUP, "multiplicative sweeps need saved patch operators");
    if (!pc->d_rowptr) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "the operator is not resident on the device");
    CUDA_OK(cudaMemsetAsync(y, 0, (size_t)pc->nowned * sizeof(double), pc->stream));
    const int nsweep = pc->symmetrise_sweep ? 2 : 1;
    bool first = true;
    for (int sweep = 0; sweep < nsweep; ++sweep) {
        for (i64 cc = 0; cc < pc->ncolours; ++cc) {
            const i64 col = sweep == 0 ? cc : pc->ncolours - 1 - cc;
            const double *r = x;
            if (!first) {
                const int bd = 256;
                ssc_residual_kernel<<<(unsigned)((pc->nlocal * 32 + bd - 1) / bd), bd, 0, pc->stream>>>(pc->nlocal, (const ll *)pc->d_rowptr, pc->d_colidx, pc->d_vals, x, y, pc->d_r);
                pc->launches++;
                r = pc->d_r;
            }
            first = false;
            for (size_t ci = 0; ci < pc->classes.size(); ++ci) {
                const SizeClass &c = pc->classes[ci];
                const i64 q0 = pc->colour_ranges[(size_t)col][ci].first, q1 = pc->colour_ranges[(size_t)col][ci].second;
                if (q1 > q0) CHK(launch_apply_class(pc, c, q0, q1, nullptr, r, y, 1.0));
            }
        }
    }
    if (pc->nbc) { ssc_bc_kernel<<<grid_for(pc->nbc, 256), 256, 0, pc->stream>>>(pc->d_bc, pc->nbc, x, y); pc->launches++; }
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int apply_device(PC *pc, const double *x, double *y)
{
    if (!pc->setup_done) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "SSCPatchSetUp() has not been called");
    if (pc->multiplicative) return apply_multiplicative(pc, x, y);
    const bool identity = !pc->sf_set;
    const double *xl = x;
    double *yl = y;
    if (!identity && pc->p2p) {
        // owned entries are read from x and added into y directly; ghosts live in d_xg (filled by the neighbours) and
        // their corrections go straight to the owners' d_yin over NVLink (PeerScatter)
        CUDA_OK(cudaMemsetAsync(y, 0, (size_t)pc->nowned * sizeof(double), pc->stream));
        CHK(peer_bcast(pc, x));
    } else {
        if (!identity) {                                    // ssc/libssc.c:1321-1326
            CHK(sf_bcast(pc, x, pc->d_xl));
            xl = pc->d_xl; yl = pc->d_yl;
        }
        CUDA_OK(cudaMemsetAsync(yl, 0, (size_t)pc->nlocal * sizeof(double), pc->stream));   // :1334
    }
    const double scale = pc->symmetrise_sweep ? 2.0 : 1.0;   // the second sweep re-adds the same corrections, :1336-1343
    cudaEvent_t k0 = nullptr, k1 = nullptr;
    if (pc->ktimer) {
        if (pc->kev_used == pc->kev.size()) {
            cudaEvent_t a, b;
            CUDA_OK(cudaEventCreate(&a)); CUDA_OK(cudaEventCreate(&b));
            pc->kev.push_back({a, b});
        }
        k0 = pc->kev[pc->kev_used].first; k1 = pc->kev[pc->kev_used].second; pc->kev_used++;
        CUDA_OK(cudaEventRecord(k0, pc->stream));
    }
    if (pc->save_operators) {
        // the size class with the most bytes goes first on the apply stream; the others run beside it on a second
        // stream (their launch gaps, ramps and tails disappear under the big kernel; the adds into y commute)
        size_t big = 0;
        for (size_t ci = 1; ci < pc->classes.size(); ++ci)
            if (pc->classes[ci].count * pc->classes[ci].stride > pc->classes[big].count * pc->classes[big].stride) big = ci;
        const bool fork = pc->overlap_classes && pc->classes.size() > 1 && !pc->deterministic;
        if (fork) {
            CUDA_OK(cudaEventRecord(pc->ev_fork, pc->stream));
            CUDA_OK(cudaStreamWaitEvent(pc->side_stream, pc->ev_fork, 0));
        }
        if (!pc->classes.empty()) CHK(launch_apply_class(pc, pc->classes[big], pc->classes[big].first, pc->classes[big].first + pc->classes[big].count, nullptr, xl, yl, scale));
        if (fork) pc->launch_stream = pc->side_stream;
        for (size_t ci = pc->classes.size(); ci-- > 0;) {
            if (ci == big) continue;
            const SizeClass &c = pc->classes[ci];
            const int rc = launch_apply_class(pc, c, c.first, c.first + c.count, nullptr, xl, yl, scale);
            if (rc) { pc->launch_stream = pc->stream; return rc; }
        }
        pc->launch_stream = pc->stream;
        if (fork) {
            CUDA_OK(cudaEventRecord(pc->ev_join, pc->side_stream));
            CUDA_OK(cudaStreamWaitEvent(pc->stream, pc->ev_join, 0));
        }
    } else {
        // -pc_patch_save_operators false: rebuild, factorise, use and drop every patch operator inside the apply
        // (ssc/libssc.c:1363-1382), in batches that fit a bounded scratch buffer
        const i64 budget = (i64)1 << 27;   // doubles (1 GiB)
        for (const SizeClass &c : pc->classes) {
            const i64 per = std::max<i64>(1, budget / c.stride);
            if (!pc->d_blocks) CHK(dev_alloc(&pc->d_blocks, std::max<i64>(budget, pc->max_n * pc->max_n + 16)));
            for (i64 q0 = c.first; q0 < c.first + c.count; q0 += per) {
                const i64 q1 = std::min(c.first + c.count, q0 + per);
                SizeClass sub = c; sub.first = q0; sub.goff = c.goff + (q0 - c.first) * c.n; sub.moff = 0;
                CHK(extract_and_invert(pc, sub, q0, q1, pc->d_blocks, pc->d_fail + q0, false));
                CHK(launch_apply_class(pc, sub, q0, q1, pc->d_blocks, xl, yl, scale));
            }
        }
    }
    if (pc->deterministic) {
        ssc_slot_sum_kernel<<<grid_for(pc->nlocal, 256), 256, 0, pc->stream>>>(pc->nlocal, (const ll *)pc->d_slotoff, pc->d_slots, pc->d_ypatch, yl);
        pc->launches++;
    }
    if (pc->ktimer) CUDA_OK(cudaEventRecord(k1, pc->stream));
    if (!identity && pc->p2p) CHK(peer_fold(pc, y));         // neighbours' contributions to the interface dofs
    else if (!identity) CHK(sf_reduce(pc, yl, y));           // :1396-1401
    if (pc->nbc) { ssc_bc_kernel<<<grid_for(pc->nbc, 256), 256, 0, pc->stream>>>(pc->d_bc, pc->nbc, x, y); pc->launches++; }   // :1404-1413
    if (pc->partition_of_unity) {                            // :1418-1421
        if (!pc->d_w) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "partition of unity was switched on after the first setup");
        ssc_scale_kernel<<<grid_for(pc->nowned, 256), 256, 0, pc->stream>>>(pc->nowned, pc->d_w, y); pc->launches++;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

// Patch-number bounds of the slabs of a pipelined host-space apply.  Tapered (default): slab sizes grow geometrically
// from both ends (1, 2, 4, ... , 4, 2, 1), so the first kernels start after a sliver of x has arrived and only a sliver
// of y is left to copy back when the last kernel ends, while the middle of the mesh runs in a few large launches.
static i64 slab_bound(const PC *pc, int s, int S, i64 np)
{
    if (!pc->pipe_taper || S < 4) return np * s / S;
    double tot = 0, acc = 0;
    for (int k = 0; k < S; ++k) tot += (double)(1u << std::min(std::min(k, S - 1 - k), 5));
    for (int k = 0; k < s; ++k) acc += (double)(1u << std::min(std::min(k, S - 1 - k), 5));
    return s >= S ? np : (i64)((double)np * acc / tot);
}

// Slab schedule for host-space applies when neighbour ranks take part (peer exchange).  Same slabs by patch number as
// below, but (i) the pieces of x that hold interface values travel first, so the halo push and the first neighbour
// barrier happen a fraction of a millisecond into the apply, (ii) the slabs that touch ghost dofs run first and the
// rank announces the second barrier right after them -- the neighbours' contributions to this rank's interface dofs
// are complete long before the interior slabs are -- and (iii) the pieces of y that hold interface dofs leave last,
// after the fold; every other piece leaves as soon as the slabs that write into it are done.
static int build_peer_pipeline(PC *pc)
{
    const int S = pc->pipe_slabs;
    const PatchLists &L = pc->lists;
    const i64 np = L.npatch, N = pc->nowned;
    PC::PeerPipe &P = pc->pp;
    std::vector<i64> lo(S, N), hi(S, -1);
    P.slab_ghost.assign(S, 0);
    pc->pipe_ranges.assign(S, std::vector<std::pair<i64, i64>>(pc->classes.size()));
    for (int s = 0; s < S; ++s) {
        const i64 p0 = slab_bound(pc, s, S, np), p1 = slab_bound(pc, s + 1, S, np);
        for (size_t ci = 0; ci < pc->classes.size(); ++ci) {
            const SizeClass &c = pc->classes[ci];
            auto first = pc->sorted2orig.begin() + c.first, last = first + c.count;
            pc->pipe_ranges[s][ci] = {std::lower_bound(first, last, p0) - pc->sorted2orig.begin(), std::lower_bound(first, last, p1) - pc->sorted2orig.begin()};
        }
        for (i64 p = p0; p < p1; ++p)
            for (i64 k = L.gtolOff[p]; k < L.gtolOff[p + 1]; ++k) {
                const i64 g = L.gtol[k];
                if (g >= N) P.slab_ghost[s] = 1;
                else { lo[s] = std::min(lo[s], g); hi[s] = std::max(hi[s], g); }
            }
    }
    P.cut.assign(S + 1, 0);
    i64 run = 0;
    for (int s = 0; s < S; ++s) { run = std::max(run, hi[s] + 1); P.cut[s + 1] = s == S - 1 ? N : run; }
    auto piece_of = [&](i64 g) { return (int)(std::upper_bound(P.cut.begin() + 1, P.cut.end(), g) - (P.cut.begin() + 1)); };
    P.slab_pieces.assign(S, {0, -1});
    for (int s = 0; s < S; ++s) if (hi[s] >= 0) P.slab_pieces[s] = {std::min(piece_of(lo[s]), S - 1), std::min(piece_of(hi[s]), S - 1)};
    // launch order: ghost-touching slabs first
    P.slab_order.clear();
    for (int s = 0; s < S; ++s) if (P.slab_ghost[s]) P.slab_order.push_back(s);
    P.nghost_slabs = (int)P.slab_order.size();
    for (int s = 0; s < S; ++s) if (!P.slab_ghost[s]) P.slab_order.push_back(s);
    // copy order: pieces with interface values first, then the pieces in the order the slabs ask for them
    std::vector<char> has_send(S, 0), placed(S, 0);
    P.piece_late.assign(S, 0);
    for (i64 g : pc->sendRoot) if (g >= 0 && g < N) { const int q = std::min(piece_of(g), S - 1); has_send[q] = 1; P.piece_late[q] = 1; }
    P.copy_order.clear();
    for (int q = 0; q < S; ++q) if (has_send[q]) { P.copy_order.push_back(q); placed[q] = 1; }
    P.nsend_pieces = (int)P.copy_order.size();
    for (int s : P.slab_order)
        for (int q = P.slab_pieces[s].first; q <= P.slab_pieces[s].second; ++q) if (!placed[q]) { P.copy_order.push_back(q); placed[q] = 1; }
    for (int q = 0; q < S; ++q) if (!placed[q]) P.copy_order.push_back(q);
    // a piece of y is complete after the last (in launch order) slab that writes into it
    P.piece_ready.assign(S, -1);
    for (int pos = 0; pos < S; ++pos) {
        const int s = P.slab_order[pos];
        for (int q = P.slab_pieces[s].first; q <= P.slab_pieces[s].second; ++q) P.piece_ready[q] = std::max(P.piece_ready[q], pos);
    }
    std::vector<i64> bc;
    for (i64 d : pc->globalBc) if (d >= 0 && d < N) bc.push_back(d);
    if (!std::is_sorted(bc.begin(), bc.end())) return 0;
    P.bc_off.assign(S + 1, 0);
    for (int q = 0; q <= S; ++q) P.bc_off[q] = std::lower_bound(bc.begin(), bc.end(), P.cut[q]) - bc.begin();
    if (!pc->s_in) {
        CUDA_OK(cudaStreamCreateWithFlags(&pc->s_in, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&pc->s_out, cudaStreamNonBlocking));
    }
    while ((int)pc->ev_in.size() < S + 1) {
        cudaEvent_t a, b;
        CUDA_OK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); CUDA_OK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        pc->ev_in.push_back(a); pc->ev_k.push_back(b);
    }
    pc->pipe_ready = true;
    return 0;
}

// the classes of one slab of a pipelined host-space apply: biggest class on the apply stream, the rest beside it
static int launch_slab_classes(PC *pc, int slab, const double *dx, double *dy, double scale, const std::vector<int> &input_events)
{
    size_t big = 0;
    i64 best = -1;
    for (size_t ci = 0; ci < pc->classes.size(); ++ci) {
        const i64 w = (pc->pipe_ranges[slab][ci].second - pc->pipe_ranges[slab][ci].first) * pc->classes[ci].stride;
        if (w > best) { best = w; big = ci; }
    }
    const bool fork = pc->overlap_classes && pc->classes.size() > 1;
    if (fork) for (int e : input_events) CUDA_OK(cudaStreamWaitEvent(pc->side_stream, pc->ev_in[e], 0));
    if (fork) {                                               // and everything already on the apply stream (the memset of y, earlier slabs' Dirichlet rows)
        CUDA_OK(cudaEventRecord(pc->ev_fork, pc->stream));
        CUDA_OK(cudaStreamWaitEvent(pc->side_stream, pc->ev_fork, 0));
    }
    for (int pass = 0; pass < 2; ++pass) {
        for (size_t ci = 0; ci < pc->classes.size(); ++ci) {
            if ((pass == 0) != (ci == big)) continue;
            const SizeClass &c = pc->classes[ci];
            const i64 q0 = pc->pipe_ranges[slab][ci].first, q1 = pc->pipe_ranges[slab][ci].second;
            if (q1 <= q0) continue;
            pc->launch_stream = (fork && pass == 1) ? pc->side_stream : pc->stream;
            const int rc = launch_apply_class(pc, c, q0, q1, nullptr, dx, dy, scale);
            pc->launch_stream = pc->stream;
            if (rc) return rc;
        }
    }
    if (fork) {
        CUDA_OK(cudaEventRecord(pc->ev_join, pc->side_stream));
        CUDA_OK(cudaStreamWaitEvent(pc->stream, pc->ev_join, 0));
    }
    return 0;
}

static int apply_host_pipelined_peer(PC *pc, const double *x, double *y)
{
    const int S = pc->pipe_slabs, nn = (int)pc->neigh.size();
    const PC::PeerPipe &P = pc->pp;
    double *dx = pc->d_xs, *dy = pc->d_ys;
    const double scale = pc->symmetrise_sweep ? 2.0 : 1.0;
    // SSC_PIPE_TRACE=1: per-apply timeline on stderr (ms after the start: halo barrier passed, ghost slabs done, all
    // slabs done, neighbours' second barrier collected, last byte of y on the host)
    static const bool trace = getenv("SSC_PIPE_TRACE") != nullptr;
    cudaEvent_t *tev = pc->trace_ev;                          // per PC: events belong to the device they were created on
    if (trace && !tev[0]) for (int i = 0; i < 6; ++i) cudaEventCreate(&tev[i]);
    if (trace) cudaEventRecord(tev[0], pc->stream);
    CUDA_OK(cudaMemsetAsync(dy, 0, (size_t)pc->nowned * sizeof(double), pc->stream));
    CUDA_OK(cudaEventRecord(pc->ev_k[S], pc->stream));
    CUDA_OK(cudaStreamWaitEvent(pc->s_in, pc->ev_k[S], 0));        // do not overwrite x while an earlier apply still reads it
    for (int q : P.copy_order) {
        const i64 a = P.cut[q], b = P.cut[q + 1];
        if (b > a) CUDA_OK(cudaMemcpyAsync(dx + a, x + a, (size_t)(b - a) * sizeof(double), cudaMemcpyHostToDevice, pc->s_in));
        CUDA_OK(cudaEventRecord(pc->ev_in[q], pc->s_in));
    }
    // halo push as soon as the interface values are on the device, then the first neighbour barrier
    for (int i = 0; i < P.nsend_pieces; ++i) CUDA_OK(cudaStreamWaitEvent(pc->stream, pc->ev_in[P.copy_order[i]], 0));
    CHK(peer_bcast(pc, dx));
    if (trace) cudaEventRecord(tev[1], pc->stream);
    const unsigned long long e2 = ++pc->epoch;                    // the second barrier of this apply, in two halves
    auto emit_piece = [&](int q) -> int {
        const i64 a = P.cut[q], b = P.cut[q + 1];
        const i64 nb = P.bc_off[q + 1] - P.bc_off[q];
        CUDA_OK(cudaStreamWaitEvent(pc->stream, pc->ev_in[q], 0));      // the Dirichlet rows copy x: a piece no slab reads has not been waited for yet
        if (nb > 0) { ssc_bc_kernel<<<grid_for(nb, 256), 256, 0, pc->stream>>>(pc->d_bc + P.bc_off[q], nb, dx, dy); pc->launches++; }
        if (pc->partition_of_unity && b > a) { ssc_scale_kernel<<<grid_for(b - a, 256), 256, 0, pc->stream>>>(b - a, pc->d_w + a, dy + a); pc->launches++; }
        CUDA_OK(cudaEventRecord(pc->ev_k[q], pc->stream));
        CUDA_OK(cudaStreamWaitEvent(pc->s_out, pc->ev_k[q], 0));
        if (b > a) CUDA_OK(cudaMemcpyAsync(y + a, dy + a, (size_t)(b - a) * sizeof(double), cudaMemcpyDeviceToHost, pc->s_out));
        return 0;
    };
    auto signal = [&]() -> int {
        if (nn) { ssc_peer_signal_kernel<<<1, 32, 0, pc->stream>>>(nn, pc->h_flags, e2); pc->launches++; }
        return 0;
    };
    if (P.nghost_slabs == 0) CHK(signal());
    for (int pos = 0; pos < S; ++pos) {
        const int s = P.slab_order[pos];
        std::vector<int> need;
        for (int q = P.slab_pieces[s].first; q <= P.slab_pieces[s].second; ++q) { CUDA_OK(cudaStreamWaitEvent(pc->stream, pc->ev_in[q], 0)); need.push_back(q); }
        CHK(launch_slab_classes(pc, s, dx, dy, scale, need));
        if (pos == P.nghost_slabs - 1) { CHK(signal()); if (trace) cudaEventRecord(tev[2], pc->stream); }   // nothing after this point writes to a neighbour
        for (int q = 0; q < S; ++q) if (P.piece_ready[q] == pos && !P.piece_late[q]) CHK(emit_piece(q));
    }
    if (trace) cudaEventRecord(tev[3], pc->stream);
    if (nn) { ssc_peer_wait_kernel<<<1, 32, 0, pc->stream>>>(nn, pc->d_flags, e2, pc->d_timeout); pc->launches++; }
    if (trace) cudaEventRecord(tev[4], pc->stream);
    if (pc->niface) { ssc_peer_fold_kernel<<<grid_for(pc->niface, 256), 256, 0, pc->stream>>>(pc->niface, pc->d_iface, pc->d_yin, dy); pc->launches++; }
    for (int q = 0; q < S; ++q) if (P.piece_late[q] || P.piece_ready[q] < 0) CHK(emit_piece(q));
    CUDA_OK(cudaGetLastError());
    if (trace) cudaEventRecord(tev[5], pc->s_out);
    CUDA_OK(cudaStreamSynchronize(pc->s_out));
    CUDA_OK(cudaStreamSynchronize(pc->stream));
    if (trace) {
        float t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 1; i < 6; ++i) cudaEventElapsedTime(&t[i], tev[0], tev[i]);
        cudaGetLastError();      // an event that was never recorded (no ghost slabs on this rank) is not an error of the apply
        fprintf(stderr, "[ssc pipe rank %d] halo %.2f  ghost slabs %.2f  all slabs %.2f  neighbours %.2f  y home %.2f ms\n", getenv("RANK") ? atoi(getenv("RANK")) : 0, t[1], t[2], t[3], t[4], t[5]);
    }
    return 0;
}

// Slab schedule for host-space applies.  Patches keep their construction (mesh) order inside a size class, and dof
// numbers follow the mesh too, so consecutive patches touch a moving window of x and y.  The patch list is cut into
// S slabs by original patch number; slab s can start once x[0, in[s+1]) has arrived and, once it is done,
// y[out[s], out[s+1]) is final.  With an unrelated numbering the bounds collapse to "everything first" and the path
// degenerates to the serial one -- it is never wrong, only not overlapped.
static int build_host_pipeline(PC *pc)
{
    pc->pipe_ready = false;
    const int S = pc->pipe_slabs;
    if (S < 2 || (pc->sf_set && !pc->p2p) || pc->multiplicative || pc->deterministic || !pc->save_operators || (!pc->pipe_force && pc->sum_n2 * 8 < ((i64)64 << 20))) return 0;
    if (pc->lists.npatch < S) return 0;
    if (pc->sf_set) return build_peer_pipeline(pc);
    const PatchLists &L = pc->lists;
    const i64 np = L.npatch, N = pc->nowned;
    std::vector<i64> lo(S, N), hi(S, -1);
    pc->pipe_ranges.assign(S, std::vector<std::pair<i64, i64>>(pc->classes.size()));
    for (int s = 0; s < S; ++s) {
        const i64 p0 = slab_bound(pc, s, S, np), p1 = slab_bound(pc, s + 1, S, np);
        for (size_t ci = 0; ci < pc->classes.size(); ++ci) {
            const SizeClass &c = pc->classes[ci];
            auto first = pc->sorted2orig.begin() + c.first, last = first + c.count;
            const i64 q0 = std::lower_bound(first, last, p0) - pc->sorted2orig.begin();
            const i64 q1 = std::lower_bound(first, last, p1) - pc->sorted2orig.begin();
            pc->pipe_ranges[s][ci] = {q0, q1};
        }
        for (i64 p = p0; p < p1; ++p)
            for (i64 k = L.gtolOff[p]; k < L.gtolOff[p + 1]; ++k) { lo[s] = std::min(lo[s], L.gtol[k]); hi[s] = std::max(hi[s], L.gtol[k]); }
    }
    pc->pipe_in.assign(S + 1, 0); pc->pipe_out.assign(S + 1, 0);
    i64 run = 0;
    for (int s = 0; s < S; ++s) { run = std::max(run, hi[s] + 1); pc->pipe_in[s + 1] = s == S - 1 ? N : run; }
    i64 suf = N;
    pc->pipe_out[S] = N;
    for (int s = S - 1; s >= 1; --s) { suf = std::min(suf, lo[s]); pc->pipe_out[s] = suf; }
    for (int s = 1; s < S; ++s) pc->pipe_out[s] = std::min(pc->pipe_out[s], pc->pipe_in[s]);   // a piece of y never ends beyond the x that has arrived (Dirichlet rows copy x)
    for (int s = 1; s <= S; ++s) pc->pipe_out[s] = std::max(pc->pipe_out[s], pc->pipe_out[s - 1]);
    // Dirichlet list offsets per y piece (the list is uploaded in ascending order)
    std::vector<i64> bc;
    for (i64 d : pc->globalBc) if (d >= 0 && d < N) bc.push_back(d);
    if (!std::is_sorted(bc.begin(), bc.end())) return 0;
    pc->pipe_bc.assign(S + 1, 0);
    for (int s = 0; s <= S; ++s) pc->pipe_bc[s] = std::lower_bound(bc.begin(), bc.end(), pc->pipe_out[s]) - bc.begin();
    if (!pc->s_in) {
        CUDA_OK(cudaStreamCreateWithFlags(&pc->s_in, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&pc->s_out, cudaStreamNonBlocking));
    }
    while ((int)pc->ev_in.size() < S + 1) {
        cudaEvent_t a, b;
        CUDA_OK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); CUDA_OK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        pc->ev_in.push_back(a); pc->ev_k.push_back(b);
    }
    pc->pipe_ready = true;
    return 0;
}

static int apply_host_pipelined(PC *pc, const double *x, double *y)
{
    const int S = pc->pipe_slabs;
    double *dx = pc->d_xs, *dy = pc->d_ys;
    const double scale = pc->symmetrise_sweep ? 2.0 : 1.0;
    CUDA_OK(cudaMemsetAsync(dy, 0, (size_t)pc->nowned * sizeof(double), pc->stream));
    CUDA_OK(cudaEventRecord(pc->ev_k[S], pc->stream));
    CUDA_OK(cudaStreamWaitEvent(pc->s_in, pc->ev_k[S], 0));        // do not overwrite x while an earlier apply still reads it
    for (int s = 0; s < S; ++s) {
        const i64 a = pc->pipe_in[s], b = pc->pipe_in[s + 1];
        if (b > a) CUDA_OK(cudaMemcpyAsync(dx + a, x + a, (size_t)(b - a) * sizeof(double), cudaMemcpyHostToDevice, pc->s_in));
        CUDA_OK(cudaEventRecord(pc->ev_in[s], pc->s_in));
    }
    for (int s = 0; s < S; ++s) {
        CUDA_OK(cudaStreamWaitEvent(pc->stream, pc->ev_in[s], 0));
        CHK(launch_slab_classes(pc, s, dx, dy, scale, std::vector<int>()));      // the fork event already carries the wait on x
        const i64 a = pc->pipe_out[s], b = pc->pipe_out[s + 1];
        const i64 nb = pc->pipe_bc[s + 1] - pc->pipe_bc[s];
        if (nb > 0) { ssc_bc_kernel<<<grid_for(nb, 256), 256, 0, pc->stream>>>(pc->d_bc + pc->pipe_bc[s], nb, dx, dy); pc->launches++; }
        if (pc->partition_of_unity && b > a) { ssc_scale_kernel<<<grid_for(b - a, 256), 256, 0, pc->stream>>>(b - a, pc->d_w + a, dy + a); pc->launches++; }
        CUDA_OK(cudaEventRecord(pc->ev_k[s], pc->stream));
        CUDA_OK(cudaStreamWaitEvent(pc->s_out, pc->ev_k[s], 0));
        if (b > a) CUDA_OK(cudaMemcpyAsync(y + a, dy + a, (size_t)(b - a) * sizeof(double), cudaMemcpyDeviceToHost, pc->s_out));
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(pc->s_out));
    CUDA_OK(cudaStreamSynchronize(pc->stream));
    return 0;
}

extern "C" int SSCPatchApply(SSCPatchPC pc, const double *x, double *y, int memspace)
{
    if (!pc || !x || !y) return ssc_set_error(SSC_ERR_ARG_NULL, "null argument");
    CHK(set_device(pc));
    if (memspace == SSC_MEM_DEVICE) return apply_device(pc, x, y);
    if (!pc->setup_done) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "SSCPatchSetUp() has not been called");
    if (!pc->d_xs) { CHK(dev_alloc(&pc->d_xs, pc->nowned)); CHK(dev_alloc(&pc->d_ys, pc->nowned)); }
    if (!pc->pipe_ready && !pc->pipe_tried) { pc->pipe_tried = true; CHK(build_host_pipeline(pc)); }
    if (pc->pipe_ready && pc->partition_of_unity && !pc->d_w) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "partition of unity was switched on after the first setup");
    if (pc->pipe_ready) return pc->sf_set ? apply_host_pipelined_peer(pc, x, y) : apply_host_pipelined(pc, x, y);
    CUDA_OK(cudaMemcpyAsync(pc->d_xs, x, (size_t)pc->nowned * sizeof(double), cudaMemcpyHostToDevice, pc->stream));
    CHK(apply_device(pc, pc->d_xs, pc->d_ys));
    CUDA_OK(cudaMemcpyAsync(y, pc->d_ys, (size_t)pc->nowned * sizeof(double), cudaMemcpyDeviceToHost, pc->stream));
    CUDA_OK(cudaStreamSynchronize(pc->stream));
    return 0;
}

extern "C" int SSCPatchApplyTranspose(SSCPatchPC, const double *, double *, int)
{
    // pc->ops->applytranspose = 0 in the reference (ssc/libssc.c:1715) => PETSc raises PETSC_ERR_SUP
    return ssc_set_error(SSC_ERR_SUP, "PC does not have apply transpose");
}

extern "C" int SSCPatchSynchronize(SSCPatchPC pc)
{
    CHK(set_device(pc));
    CUDA_OK(cudaStreamSynchronize(pc->stream));
    return peer_check(pc);
}

extern "C" int SSCPatchView(SSCPatchPC pc, char *buf, SSCInt len)
{
    // the lines of PCView_PATCH ssc/libssc.c:1637-1667
    std::string s;
    char line[256];
    snprintf(line, sizeof line, "  Subspace Correction preconditioner with %lld patches\n", (long long)pc->lists.npatch); s += line;
    s += pc->multiplicative ? "  Schwarz type: multiplicative\n" : "  Schwarz type: additive\n";
    if (pc->multiplicative && pc->setup_done) { snprintf(line, sizeof line, "  %lld colours of mutually independent patches\n", (long long)pc->ncolours); s += line; }
    s += pc->partition_of_unity ? "  Weighting by partition of unity\n" : "  Not weighting by partition of unity\n";
    s += pc->symmetrise_sweep ? "  Symmetrising sweep (start->end, then end->start)\n" : "  Not symmetrising sweep\n";
    s += !pc->save_operators ? "  Not saving patch operators (rebuilt every PCApply)\n" : "  Saving patch operators (rebuilt every PCSetUp)\n";
    if (pc->bopt.construct_type == SSC_PATCH_STAR) s += "  Patch construction operator: star\n";
    else if (pc->bopt.construct_type == SSC_PATCH_VANKA) s += "  Patch construction operator: Vanka\n";
    else s += "  Patch construction operator: user-specified\n";
    s += "  KSP on patches (all same):\n";
    if (pc->setup_done) {
        snprintf(line, sizeof line, "    KSP Object: type preonly; PC Object: type %s (dense fp64 blocks, stored inverse, sm_100a)\n", pc->sub_pc == SSC_SUB_LU ? "lu" : "cholesky"); s += line;
        snprintf(line, sizeof line, "    %zu size classes, max patch size %lld, %lld patch dofs, %.3f GB of blocks on device %d\n", pc->classes.size(), (long long)pc->max_n,
                 (long long)pc->sum_n, pc->block_elems * 8.0 / 1e9, pc->device); s += line;
    } else {
        s += "    KSP not yet set.\n";
    }
    if (len > 0) { strncpy(buf, s.c_str(), (size_t)len - 1); buf[len - 1] = 0; }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------
extern "C" int SSCPatchGetSize(SSCPatchPC pc, const char *what, SSCInt *value)
{
    std::string w(what);
    i64 nnz = 0;
    if (w == "npatch") *value = pc->lists.npatch;
    else if (w == "nlocal") *value = pc->nlocal;
    else if (w == "nowned") *value = pc->nowned;
    else if (w == "sum_n") *value = pc->sum_n;
    else if (w == "sum_n2") *value = pc->sum_n2;
    else if (w == "max_n") *value = pc->max_n;
    else if (w == "num_cells") *value = (i64)pc->lists.cells.size();
    else if (w == "num_gtol") *value = (i64)pc->lists.gtol.size();
    else if (w == "num_dofs") *value = (i64)pc->lists.dofs.size();
    else if (w == "num_classes") *value = (i64)pc->classes.size();
    else if (w == "num_colours") *value = pc->ncolours;
    else if (w == "lists_from_device") *value = pc->lists_from_device ? 1 : 0;
    else if (w == "block_bytes") *value = (pc->symmetric ? pc->packed_elems : pc->block_elems) * 8;
    else if (w == "bytes_apply") *value = (pc->symmetric ? 4 * (pc->sum_n2 + pc->sum_n) : 8 * pc->sum_n2) + (4 + 8 + 8) * pc->sum_n + 16 * pc->nowned;   // SURVEY.md 8(d); packed: 4 sum n(n+1)
    else if (w == "bytes_extract") { nnz = pc->csr_nnz; *value = 12 * nnz + 8 * (pc->nlocal + 1) + 4 * pc->sum_n + 8 * pc->sum_n2; }   // nnz as of the last host-space SetUp
    else if (w == "flops_factor") { i64 f = 0; for (auto &c : pc->classes) f += 2 * c.count * (i64)c.n * c.n * c.n; *value = f; }
    else if (w == "failed_patches") *value = pc->failed;
    else if (w == "gpu_launches") *value = pc->launches;
    else return ssc_set_error(SSC_ERR_USER, "unknown size '%s'", what);
    return 0;
}

// N2: star patches built on the device.  Returns 0 when pc->lists was filled, 1 when the case is not covered (or a patch
// did not fit the kernel's shared-memory tables) and the host builder must run, an error code otherwise.
static int build_lists_on_device(PC *pc)
{
    using namespace devbuild;
    const HostPlex &P = pc->plex;
    const Discretisation &D = pc->disc;
    const BuildOptions &O = pc->bopt;
    if (pc->device < 0 || !pc->device_construct || !P.set || !D.set || !D.cn_set) return 1;
    if (O.construct_type != 0 || O.exclude_subspace >= 0 || O.keep_dofs || O.print_patches || D.nsub > kMaxSub) return 1;
    const i64 nlocal = D.subspaceOffsets[D.nsub];
    if (nlocal > 2147483647LL || P.pEnd > 2147483647LL || (i64)P.cones.size() > 2147483647LL) return 1;
    const i64 cStart = P.depthStart[P.dim], cEnd = P.depthEnd[P.dim];
    i64 vStart, vEnd;
    if (O.codim < 0) { const i64 d = O.dim < 0 ? 0 : O.dim; if (d > P.dim) return 1; vStart = P.depthStart[d]; vEnd = P.depthEnd[d]; }
    else { if (O.codim > P.dim) return 1; vStart = P.depthStart[P.dim - O.codim]; vEnd = P.depthEnd[P.dim - O.codim]; }
    const i64 np = vEnd - vStart;
    if (np < 4096) return 1;                                        // small meshes: the host builder is instant
    i64 ncellrows = 0;
    for (i64 c = cStart; c < cEnd; ++c) { if (D.cnDof[c] <= 0) return 1; ncellrows = std::max(ncellrows, D.cnOff[c] + 1); }
    for (i64 g : D.ghostBcNodes) if (g < 0 || g >= nlocal) return 1;
    CHK(set_device(pc));
    std::vector<void *> tmp;                                         // every device temporary, freed at the end
    auto cleanup = [&]() { for (void *q : tmp) cudaFree(q); };
    auto up64 = [&](const i64 *src, i64 n, int **dst) -> int {       // upload int64, narrow on the device
        i64 *d64 = nullptr; int *d32 = nullptr;
        if (dev_alloc(&d64, n) || dev_alloc(&d32, n)) return SSC_ERR_MEM;
        tmp.push_back(d64); tmp.push_back(d32);
        if (n) {
            if (cudaMemcpyAsync(d64, src, (size_t)n * sizeof(i64), cudaMemcpyHostToDevice, pc->stream) != cudaSuccess) return SSC_ERR_GPU;
            narrow_kernel<<<grid_for(n, 256), 256, 0, pc->stream>>>(n, (const long long *)d64, d32);
        }
        *dst = d32;
        return 0;
    };
#define UP(src, n, dst) do { int rc_ = up64(src, n, dst); if (rc_) { cleanup(); return ssc_set_error(rc_, "device patch construction: upload failed"); } } while (0)
    Topo T;
    memset(&T, 0, sizeof(T));
    int *d;
    UP(P.coneOff.data(), P.pEnd + 1, &d); T.coneOff = d;
    UP(P.cones.data(), (i64)P.cones.size(), &d); T.cones = d;
    UP(P.suppOff.data(), P.pEnd + 1, &d); T.suppOff = d;
    UP(P.supports.data(), (i64)P.supports.size(), &d); T.supports = d;
    UP(D.cnOff.data(), P.pEnd, &d); T.cnOff = d;
    unsigned char *d_ghost = nullptr, *d_bc = nullptr;
    if (!P.ghost.empty()) {
        if (dev_alloc(&d_ghost, P.pEnd)) { cleanup(); return SSC_ERR_MEM; }
        tmp.push_back(d_ghost);
        cudaMemcpyAsync(d_ghost, P.ghost.data(), (size_t)P.pEnd, cudaMemcpyHostToDevice, pc->stream);
    }
    T.ghost = d_ghost;
    {
        std::vector<unsigned char> isbc((size_t)nlocal, 0);
        for (i64 g : D.ghostBcNodes) isbc[(size_t)g] = 1;
        if (dev_alloc(&d_bc, nlocal)) { cleanup(); return SSC_ERR_MEM; }
        tmp.push_back(d_bc);
        cudaMemcpy(d_bc, isbc.data(), (size_t)nlocal, cudaMemcpyHostToDevice);
    }
    T.isBc = d_bc;
    T.cStart = (int)cStart; T.cEnd = (int)cEnd; T.vStart = (int)vStart; T.npatch = (int)np; T.nlocal = (int)nlocal; T.nsub = (int)D.nsub;
    for (i64 k = 0; k < D.nsub; ++k) {
        UP(D.secDof[k].data(), P.pEnd, &d); T.secDof[k] = d;
        UP(D.secOff[k].data(), P.pEnd, &d); T.secOff[k] = d;
        UP(D.cellNodeMap[k], ncellrows * D.nodesPerCell[k], &d); T.cnm[k] = d;
        T.bs[k] = (int)D.bs[k]; T.npc[k] = (int)D.nodesPerCell[k]; T.subOff[k] = (int)D.subspaceOffsets[k];
    }
#undef UP
    int *d_gc = nullptr, *d_cc = nullptr, *d_over = nullptr;
    if (dev_alloc(&d_gc, np) || dev_alloc(&d_cc, np) || dev_alloc(&d_over, 1)) { cleanup(); return SSC_ERR_MEM; }
    tmp.push_back(d_gc); tmp.push_back(d_cc); tmp.push_back(d_over);
    cudaMemsetAsync(d_over, 0, sizeof(int), pc->stream);
    cudaFuncSetAttribute(build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    cudaFuncSetAttribute(build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    const int grid = (int)std::min<i64>(np, (i64)pc->sm_count * 16);
    build_kernel<false><<<grid, kThreads, kSmemBytes, pc->stream>>>(T, d_gc, d_cc, nullptr, nullptr, nullptr, nullptr, d_over);
    pc->launches++;
    std::vector<int> gc((size_t)np), cc((size_t)np);
    int over = 0;
    if (cudaMemcpyAsync(gc.data(), d_gc, (size_t)np * sizeof(int), cudaMemcpyDeviceToHost, pc->stream) != cudaSuccess ||
        cudaMemcpyAsync(cc.data(), d_cc, (size_t)np * sizeof(int), cudaMemcpyDeviceToHost, pc->stream) != cudaSuccess ||
        cudaMemcpyAsync(&over, d_over, sizeof(int), cudaMemcpyDeviceToHost, pc->stream) != cudaSuccess ||
        cudaStreamSynchronize(pc->stream) != cudaSuccess) {
        cudaError_t err = cudaGetLastError();
        cleanup();
        return ssc_set_error(SSC_ERR_GPU, "device patch construction failed: %s", cudaGetErrorString(err));
    }
    if (over) { cleanup(); return 1; }
    PatchLists &L = pc->lists;
    L = PatchLists();
    L.npatch = np; L.vStart = vStart;
    L.gtolOff.assign((size_t)np + 1, 0); L.cellOff.assign((size_t)np + 1, 0);
    for (i64 p = 0; p < np; ++p) { L.gtolOff[(size_t)p + 1] = L.gtolOff[(size_t)p] + gc[(size_t)p]; L.cellOff[(size_t)p + 1] = L.cellOff[(size_t)p] + cc[(size_t)p]; }
    const i64 ng = L.gtolOff[(size_t)np], ncl = L.cellOff[(size_t)np];
    i64 *d_goff = nullptr, *d_coff = nullptr;
    int *d_g = nullptr, *d_c = nullptr;
    if (dev_alloc(&d_goff, np + 1) || dev_alloc(&d_coff, np + 1) || dev_alloc(&d_g, ng) || dev_alloc(&d_c, ncl)) { cleanup(); return SSC_ERR_MEM; }
    tmp.push_back(d_goff); tmp.push_back(d_coff); tmp.push_back(d_g); tmp.push_back(d_c);
    cudaMemcpyAsync(d_goff, L.gtolOff.data(), (size_t)(np + 1) * sizeof(i64), cudaMemcpyHostToDevice, pc->stream);
    cudaMemcpyAsync(d_coff, L.cellOff.data(), (size_t)(np + 1) * sizeof(i64), cudaMemcpyHostToDevice, pc->stream);
    build_kernel<true><<<grid, kThreads, kSmemBytes, pc->stream>>>(T, nullptr, nullptr, (const long long *)d_goff, (const long long *)d_coff, d_g, d_c, d_over);
    pc->launches++;
    std::vector<int> g32((size_t)std::max<i64>(ng, 1)), c32((size_t)std::max<i64>(ncl, 1));
    if (cudaMemcpyAsync(g32.data(), d_g, (size_t)ng * sizeof(int), cudaMemcpyDeviceToHost, pc->stream) != cudaSuccess ||
        cudaMemcpyAsync(c32.data(), d_c, (size_t)ncl * sizeof(int), cudaMemcpyDeviceToHost, pc->stream) != cudaSuccess ||
        cudaStreamSynchronize(pc->stream) != cudaSuccess) {
        cudaError_t err = cudaGetLastError();
        cleanup();
        return ssc_set_error(SSC_ERR_GPU, "device patch construction failed: %s", cudaGetErrorString(err));
    }
    cleanup();
    L.gtol.resize((size_t)ng); L.cells.resize((size_t)ncl);
    {   // widen to PetscInt with a few threads
        const int Tn = 8;
        std::vector<std::thread> th;
        for (int t = 0; t < Tn; ++t) th.emplace_back([&, t]() {
            for (i64 i = ng * t / Tn; i < ng * (t + 1) / Tn; ++i) L.gtol[(size_t)i] = g32[(size_t)i];
            for (i64 i = ncl * t / Tn; i < ncl * (t + 1) / Tn; ++i) L.cells[(size_t)i] = c32[(size_t)i];
        });
        for (auto &x : th) x.join();
    }
    pc->lists_from_device = true;
    return 0;
}

static int need_lists(PC *pc)
{
    if (pc->lists_valid) return 0;
    pc->lists_from_device = false;
    const int rc = build_lists_on_device(pc);
    if (rc == 0) { pc->lists_valid = true; return 0; }
    if (rc != 1) return rc;
    CHK(ssc_build_patches(pc->plex, pc->disc, pc->user, pc->bopt, pc->lists));
    pc->lists_valid = true;
    if (pc->print_patches) fputs(pc->lists.printout.c_str(), stdout);     // PetscSynchronizedPrintf of ssc/libssc.c:679-728
    return 0;
}
// text of -pc_patch_print_patches (also written to stdout when the patches are built); returns the full length
extern "C" int SSCPatchGetPrintout(SSCPatchPC pc, char *buf, SSCInt len, SSCInt *needed)
{
    CHK(need_lists(pc));
    if (needed) *needed = (i64)pc->lists.printout.size() + 1;
    if (buf && len > 0) { strncpy(buf, pc->lists.printout.c_str(), (size_t)len - 1); buf[len - 1] = 0; }
    return 0;
}
// host-only: build (if needed) the dof lists without touching the device
extern "C" int SSCPatchBuildPatches(SSCPatchPC pc) { return need_lists(pc); }

extern "C" int SSCPatchGetPatches(SSCPatchPC pc, SSCInt *gtolOffsets, SSCInt *gtol)
{
    CHK(need_lists(pc));
    if (gtolOffsets) std::copy(pc->lists.gtolOff.begin(), pc->lists.gtolOff.end(), gtolOffsets);
    if (gtol) std::copy(pc->lists.gtol.begin(), pc->lists.gtol.end(), gtol);
    return 0;
}
extern "C" int SSCPatchGetCells(SSCPatchPC pc, SSCInt *cellOffsets, SSCInt *cells)
{
    CHK(need_lists(pc));
    if (cellOffsets) std::copy(pc->lists.cellOff.begin(), pc->lists.cellOff.end(), cellOffsets);
    if (cells) std::copy(pc->lists.cells.begin(), pc->lists.cells.end(), cells);
    return 0;
}
extern "C" int SSCPatchGetDofs(SSCPatchPC pc, SSCInt *dofs)
{
    if (!pc->bopt.keep_dofs) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "option ssc_keep_dofs_table is off");
    CHK(need_lists(pc));
    std::copy(pc->lists.dofs.begin(), pc->lists.dofs.end(), dofs);
    return 0;
}

extern "C" int SSCPatchGetPatchMatrix(SSCPatchPC pc, SSCInt p, int which, double *out)
{
    CHK(set_device(pc));
    if (!pc->setup_done || !pc->save_operators) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "patch blocks are only held after SetUp with save_operators");
    if (p < 0 || p >= pc->lists.npatch) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "Asked for operator index is invalid");   // ssc/libssc.c:1069-1071
    const i64 q = pc->orig2sorted[p];
    if (q < 0) return 0;
    if (which != 0 && pc->symmetric) {
        for (const SizeClass &c : pc->classes) {
            if (q >= c.first && q < c.first + c.count) {
                const i64 n = c.n, psz = n * (n + 1) / 2;
                std::vector<double> tmp((size_t)psz);
                CUDA_OK(cudaStreamSynchronize(pc->stream));
                CUDA_OK(cudaMemcpy(tmp.data(), pc->d_packed + c.poff + (q - c.first) * c.pstride, (size_t)psz * sizeof(double), cudaMemcpyDeviceToHost));
                for (i64 j = 0; j < n; ++j) for (i64 i = j; i < n; ++i) out[i * n + j] = out[j * n + i] = tmp[(size_t)(j * n - j * (j - 1) / 2 + (i - j))];
                return 0;
            }
        }
    }
    const double *src = which == 0 ? pc->d_blocks0 : pc->d_blocks;
    if (!src) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "unfactored blocks are only kept with option ssc_keep_patch_operators");
    for (const SizeClass &c : pc->classes) {
        if (q >= c.first && q < c.first + c.count) {
            const i64 n = c.n;
            const i64 ld = c.ld;
            std::vector<double> tmp((size_t)(n * ld));
            CUDA_OK(cudaStreamSynchronize(pc->stream));
            CUDA_OK(cudaMemcpy(tmp.data(), src + c.moff + (q - c.first) * c.stride, (size_t)(n * ld) * sizeof(double), cudaMemcpyDeviceToHost));
            for (i64 i = 0; i < n; ++i) for (i64 j = 0; j < n; ++j)
                out[i * n + j] = which == 0 ? tmp[(size_t)(i * ld + j)] : tmp[(size_t)(j * ld + i)];   // the inverse is stored transposed
            return 0;
        }
    }
    return ssc_set_error(SSC_ERR_LIB, "patch %lld not found in any size class", (long long)p);
}

extern "C" int SSCPatchGetColours(SSCPatchPC pc, SSCInt *colours)
{
    if (!pc->multiplicative || pc->colour.empty()) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "colours exist after SetUp in multiplicative mode");
    std::copy(pc->colour.begin(), pc->colour.end(), colours);
    return 0;
}

// original numbers of the patches whose operator could not be factorised at the last SetUp (at most maxn are written)
extern "C" int SSCPatchGetFailedPatches(SSCPatchPC pc, SSCInt *list, SSCInt maxn, SSCInt *nfailed)
{
    if (nfailed) *nfailed = (i64)pc->failed_list.size();
    for (i64 i = 0; i < maxn && i < (i64)pc->failed_list.size(); ++i) list[i] = pc->failed_list[(size_t)i];
    return 0;
}

extern "C" int SSCPatchGetDofWeights(SSCPatchPC pc, double *w)
{
    CHK(set_device(pc));
    if (!pc->d_w) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "partition of unity is off");
    CUDA_OK(cudaStreamSynchronize(pc->stream));
    CUDA_OK(cudaMemcpy(w, pc->d_w, (size_t)pc->nowned * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int SSCPatchGetEventTime(SSCPatchPC pc, const char *event, double *ms)
{
    auto it = pc->event_ms.find(event);
    *ms = it == pc->event_ms.end() ? 0.0 : it->second;
    return 0;
}

extern "C" int SSCPatchKernelTimerReset(SSCPatchPC pc, int enable)
{
    CHK(set_device(pc));
    CUDA_OK(cudaStreamSynchronize(pc->stream));
    pc->ktimer = enable != 0; pc->kev_used = 0; pc->ktimer_ms = 0; pc->ktimer_n = 0;
    return 0;
}
extern "C" int SSCPatchKernelTimerRead(SSCPatchPC pc, double *total_ms, SSCInt *napply)
{
    CHK(set_device(pc));
    CUDA_OK(cudaStreamSynchronize(pc->stream));
    double tot = 0;
    for (size_t i = 0; i < pc->kev_used; ++i) { float ms = 0; CUDA_OK(cudaEventElapsedTime(&ms, pc->kev[i].first, pc->kev[i].second)); tot += ms; }
    *total_ms = tot; *napply = (i64)pc->kev_used;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// device plumbing
// ---------------------------------------------------------------------------------------------
extern "C" int SSCDeviceMalloc(SSCPatchPC pc, SSCInt bytes, void **ptr) { CHK(set_device(pc)); unsigned char *p = nullptr; CHK(dev_alloc(&p, bytes)); *ptr = p; return 0; }
extern "C" int SSCDeviceFree(SSCPatchPC pc, void *ptr) { CHK(set_device(pc)); CUDA_OK(cudaFree(ptr)); return 0; }
extern "C" int SSCHostAlloc(SSCPatchPC pc, SSCInt bytes, void **ptr) { CHK(set_device(pc)); CUDA_OK(cudaMallocHost(ptr, (size_t)std::max<i64>(bytes, 1))); return 0; }
extern "C" int SSCHostFree(SSCPatchPC pc, void *ptr) { CHK(set_device(pc)); CUDA_OK(cudaFreeHost(ptr)); return 0; }
extern "C" int SSCMemcpy(SSCPatchPC pc, void *dst, const void *src, SSCInt bytes, int dstSpace, int srcSpace)
{
    CHK(set_device(pc));
    cudaMemcpyKind kind = dstSpace == SSC_MEM_DEVICE ? (srcSpace == SSC_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice)
                                                     : (srcSpace == SSC_MEM_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost);
    CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)bytes, kind, pc->stream));
    CUDA_OK(cudaStreamSynchronize(pc->stream));
    return 0;
}
extern "C" int SSCDeviceFlushL2(SSCPatchPC pc)
{
    CHK(set_device(pc));
    const size_t bytes = (size_t)256 << 20;
    if (!pc->d_flush) CUDA_OK(cudaMalloc(&pc->d_flush, bytes));
    ssc_fill_kernel<<<148 * 8, 256, 0, pc->stream>>>((ll)(bytes / 8), (double *)pc->d_flush, 1.0);
    CUDA_OK(cudaGetLastError());
    return 0;
}
// Streaming-read bandwidth of this GPU in GB/s, measured by summing the resident patch blocks (or a 4 GiB scratch buffer
// when no blocks are resident) `reps` times: the read-only counterpart of the copy-measured peak in MEASURED_PEAKS.json.
extern "C" int SSCDeviceMeasureReadBandwidth(SSCPatchPC pc, int reps, double *gbs)
{
    CHK(set_device(pc));
    const double *buf = pc->symmetric ? pc->d_packed : pc->d_blocks;
    i64 elems = pc->symmetric ? pc->packed_elems : pc->block_elems;
    double *scratch = nullptr;
    if (!buf || elems < ((i64)1 << 27)) {
        elems = (i64)1 << 29;
        CHK(dev_alloc(&scratch, elems));
        CUDA_OK(cudaMemsetAsync(scratch, 0, (size_t)elems * sizeof(double), pc->stream));
        buf = scratch;
    }
    double *sink = nullptr;
    CHK(dev_alloc(&sink, 1));
    const ll n2 = elems / 2;
    float best = 1e30f;
    for (int r = 0; r < std::max(reps, 1) + 1; ++r) {
        CUDA_OK(cudaEventRecord(pc->ev0, pc->stream));
        ssc_stream_read_kernel<<<pc->sm_count * 32, 256, 0, pc->stream>>>(reinterpret_cast<const double2 *>(buf), n2, sink);
        CUDA_OK(cudaEventRecord(pc->ev1, pc->stream));
        CUDA_OK(cudaEventSynchronize(pc->ev1));
        float ms = 0; CUDA_OK(cudaEventElapsedTime(&ms, pc->ev0, pc->ev1));
        if (r > 0) best = std::min(best, ms);
    }
    *gbs = (double)n2 * 16.0 / (best * 1e-3) / 1e9;
    cudaFree(sink);
    if (scratch) cudaFree(scratch);
    return 0;
}
extern "C" int SSCEventCreate(SSCPatchPC pc, void **ev) { CHK(set_device(pc)); cudaEvent_t e; CUDA_OK(cudaEventCreate(&e)); *ev = e; return 0; }
extern "C" int SSCEventRecord(SSCPatchPC pc, void *ev) { CHK(set_device(pc)); CUDA_OK(cudaEventRecord((cudaEvent_t)ev, pc->stream)); return 0; }
extern "C" int SSCEventElapsed(SSCPatchPC pc, void *start, void *stop, double *ms)
{
    CHK(set_device(pc));
    CUDA_OK(cudaEventSynchronize((cudaEvent_t)stop));
    float f = 0; CUDA_OK(cudaEventElapsedTime(&f, (cudaEvent_t)start, (cudaEvent_t)stop));
    *ms = f;
    return 0;
}
extern "C" int SSCEventDestroy(SSCPatchPC pc, void *ev) { CHK(set_device(pc)); CUDA_OK(cudaEventDestroy((cudaEvent_t)ev)); return 0; }

// ---------------------------------------------------------------------------------------------
// Low-order (P1) coarse correction: y = R^T A1^-1 R x, Dirichlet values passed through (ssc/lo.py:105-198).
// SURVEY.md 8(f) N1.  The coarse solve (-lo_pc_type lu) reuses the patch engine: the coarse space is one patch
// holding every coarse dof, so "extract, invert, multiply" are the same kernels.  Dense, therefore bounded in size.
// ---------------------------------------------------------------------------------------------
struct DeviceCSR { i64 nrows = 0, ncols = 0, nnz = 0; i64 *rowptr = nullptr; int *colidx = nullptr; double *vals = nullptr; };
struct SSCLowOrderPC_s {
    int device = 0;
    PC *coarse = nullptr;          // single-patch PC on the coarse operator
    DeviceCSR R, RT;
    i64 nfine = 0, ncoarse = 0, nbc = 0;
    int *d_bc = nullptr;
    double *d_w = nullptr, *d_v = nullptr, *d_xs = nullptr, *d_ys = nullptr;
    std::vector<i64> a_rowptr; std::vector<int32_t> a_colidx; std::vector<double> a_vals;
    bool have_R = false, have_A = false, setup = false;
};
typedef SSCLowOrderPC_s LO;

static void free_csr(DeviceCSR &m) { dev_free(m.rowptr); dev_free(m.colidx); dev_free(m.vals); }
static int upload_csr_matrix(LO *lo, DeviceCSR &m, i64 nrows, i64 ncols, const i64 *rowptr, const int32_t *colidx, const double *vals)
{
    free_csr(m);
    m.nrows = nrows; m.ncols = ncols; m.nnz = rowptr[nrows];
    CHK(dev_alloc(&m.rowptr, nrows + 1)); CHK(dev_alloc(&m.colidx, m.nnz)); CHK(dev_alloc(&m.vals, m.nnz));
    CUDA_OK(cudaMemcpy(m.rowptr, rowptr, (size_t)(nrows + 1) * sizeof(i64), cudaMemcpyHostToDevice));
    if (m.nnz) {
        CUDA_OK(cudaMemcpy(m.colidx, colidx, (size_t)m.nnz * sizeof(int), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(m.vals, vals, (size_t)m.nnz * sizeof(double), cudaMemcpyHostToDevice));
    }
    (void)lo;
    return 0;
}

extern "C" int SSCLowOrderCreate(int device, SSCLowOrderPC *out)
{
    LO *lo = new LO();
    lo->device = device;
    int ierr = SSCPatchCreate(device, &lo->coarse);
    if (ierr) { delete lo; return ierr; }
    *out = lo;
    return 0;
}
extern "C" int SSCLowOrderDestroy(SSCLowOrderPC *plo)
{
    if (!plo || !*plo) return 0;
    LO *lo = *plo;
    cudaSetDevice(lo->device);
    free_csr(lo->R); free_csr(lo->RT);
    dev_free(lo->d_bc); dev_free(lo->d_w); dev_free(lo->d_v); dev_free(lo->d_xs); dev_free(lo->d_ys);
    SSCPatchDestroy(&lo->coarse);
    delete lo;
    *plo = nullptr;
    return 0;
}
extern "C" int SSCLowOrderSetOption(SSCLowOrderPC lo, const char *name_, const char *value)
{
    std::string name(name_ ? name_ : ""), v(value ? value : "");
    if (name == "lo_ksp_type") { if (v != "preonly") return ssc_set_error(SSC_ERR_SUP, "lo_ksp_type '%s' is not supported (preonly)", v.c_str()); return 0; }
    if (name == "lo_pc_type") {
        if (v == "lu") return SSCPatchSetOption(lo->coarse, "sub_pc_type", "lu");
        if (v == "cholesky") return SSCPatchSetOption(lo->coarse, "sub_pc_type", "cholesky");
        return ssc_set_error(SSC_ERR_SUP, "lo_pc_type '%s' is not supported (lu, cholesky)", v.c_str());
    }
    if (name == "lo_mat_type") return 0;
    return ssc_set_error(SSC_ERR_USER, "unknown option '%s'", name.c_str());
}
// restriction matrix of ssc/lo.py:83-102: ncoarse x nfine, rows/columns of Dirichlet nodes already dropped
extern "C" int SSCLowOrderSetRestriction(SSCLowOrderPC lo, SSCInt ncoarse, SSCInt nfine, const SSCInt *rowptr, const int32_t *colidx, const double *vals)
{
    CUDA_OK(cudaSetDevice(lo->device));
    CHK(upload_csr_matrix(lo, lo->R, ncoarse, nfine, rowptr, colidx, vals));
    // transpose on the host (set-up only)
    const i64 nnz = rowptr[ncoarse];
    std::vector<i64> tp((size_t)nfine + 1, 0);
    for (i64 k = 0; k < nnz; ++k) tp[colidx[k] + 1]++;
    for (i64 j = 0; j < nfine; ++j) tp[j + 1] += tp[j];
    std::vector<int32_t> tc((size_t)nnz);
    std::vector<double> tv((size_t)nnz);
    std::vector<i64> fill(tp.begin(), tp.end() - 1);
    for (i64 i = 0; i < ncoarse; ++i)
        for (i64 k = rowptr[i]; k < rowptr[i + 1]; ++k) { const i64 d = fill[colidx[k]]++; tc[d] = (int32_t)i; tv[d] = vals[k]; }
    CHK(upload_csr_matrix(lo, lo->RT, nfine, ncoarse, tp.data(), tc.data(), tv.data()));
    lo->nfine = nfine; lo->ncoarse = ncoarse; lo->have_R = true;
    dev_free(lo->d_w); dev_free(lo->d_v); dev_free(lo->d_xs); dev_free(lo->d_ys);
    CHK(dev_alloc(&lo->d_w, ncoarse)); CHK(dev_alloc(&lo->d_v, ncoarse)); CHK(dev_alloc(&lo->d_xs, nfine)); CHK(dev_alloc(&lo->d_ys, nfine));
    return 0;
}
// coarse operator of ssc/lo.py:142-144 (rows of Dirichlet nodes = identity); values are copied, re-settable before every SetUp
extern "C" int SSCLowOrderSetOperator(SSCLowOrderPC lo, SSCInt n, const SSCInt *rowptr, const int32_t *colidx, const double *vals)
{
    lo->a_rowptr.assign(rowptr, rowptr + n + 1);
    lo->a_colidx.assign(colidx, colidx + rowptr[n]);
    lo->a_vals.assign(vals, vals + rowptr[n]);
    lo->have_A = true;
    if (lo->ncoarse && lo->ncoarse != n) return ssc_set_error(SSC_ERR_ARG_INCOMP, "coarse operator has %lld rows, restriction has %lld", (long long)n, (long long)lo->ncoarse);
    return 0;
}
// owned Dirichlet dofs of the fine space (ssc/lo.py:175-183)
extern "C" int SSCLowOrderSetBcs(SSCLowOrderPC lo, SSCInt nbc, const SSCInt *bc)
{
    CUDA_OK(cudaSetDevice(lo->device));
    std::vector<int> b32((size_t)nbc);
    for (i64 i = 0; i < nbc; ++i) b32[(size_t)i] = (int)bc[i];
    dev_free(lo->d_bc);
    CHK(dev_alloc(&lo->d_bc, nbc));
    if (nbc) CUDA_OK(cudaMemcpy(lo->d_bc, b32.data(), (size_t)nbc * sizeof(int), cudaMemcpyHostToDevice));
    lo->nbc = nbc;
    return 0;
}
// P1PC.initialize / update (ssc/lo.py:112-186): factorise the coarse operator
extern "C" int SSCLowOrderSetUp(SSCLowOrderPC lo)
{
    if (!lo->have_R || !lo->have_A) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "set the restriction and the coarse operator before SetUp");
    const i64 n = lo->ncoarse;
    if (n > 3000) return ssc_set_error(SSC_ERR_SUP, "coarse space with %lld dofs: the dense coarse LU of this build is limited to 3000", (long long)n);
    if (!lo->setup) {
        std::vector<i64> off = {0, n}, gtol((size_t)n);
        for (i64 i = 0; i < n; ++i) gtol[(size_t)i] = i;
        CHK(SSCPatchSetPatches(lo->coarse, 1, off.data(), gtol.data(), n, n, 0, nullptr));
    }
    CHK(SSCPatchSetOperatorCSR(lo->coarse, n, lo->a_rowptr.data(), lo->a_colidx.data(), lo->a_vals.data(), SSC_MEM_HOST));
    CHK(SSCPatchSetUp(lo->coarse));
    lo->setup = true;
    return 0;
}
// P1PC.apply (ssc/lo.py:187-198)
extern "C" int SSCLowOrderApply(SSCLowOrderPC lo, const double *x, double *y, int memspace)
{
    if (!lo->setup) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "SSCLowOrderSetUp() has not been called");
    CUDA_OK(cudaSetDevice(lo->device));
    cudaStream_t st = lo->coarse->stream;
    const double *dx = x;
    double *dy = y;
    if (memspace == SSC_MEM_HOST) {
        CUDA_OK(cudaMemcpyAsync(lo->d_xs, x, (size_t)lo->nfine * sizeof(double), cudaMemcpyHostToDevice, st));
        dx = lo->d_xs; dy = lo->d_ys;
    }
    const int bd = 256;
    ssc_spmv_kernel<<<(unsigned)((lo->ncoarse * 32 + bd - 1) / bd), bd, 0, st>>>(lo->R.nrows, (const ll *)lo->R.rowptr, lo->R.colidx, lo->R.vals, dx, lo->d_w);
    CHK(SSCPatchApply(lo->coarse, lo->d_w, lo->d_v, SSC_MEM_DEVICE));
    ssc_spmv_kernel<<<(unsigned)((lo->nfine * 32 + bd - 1) / bd), bd, 0, st>>>(lo->RT.nrows, (const ll *)lo->RT.rowptr, lo->RT.colidx, lo->RT.vals, lo->d_v, dy);
    if (lo->nbc) ssc_bc_kernel<<<grid_for(lo->nbc, 256), 256, 0, st>>>(lo->d_bc, lo->nbc, dx, dy);
    lo->coarse->launches += 3;
    CUDA_OK(cudaGetLastError());
    if (memspace == SSC_MEM_HOST) {
        CUDA_OK(cudaMemcpyAsync(y, lo->d_ys, (size_t)lo->nfine * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
    }
    return 0;
}
