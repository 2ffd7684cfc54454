// Internal declarations shared by the host-side patch builder and the CUDA engine.
#pragma once
#include <cstdint>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <new>
#include <string>
#include <vector>

typedef int64_t i64;

// PETSc error codes the reference raises (petscerror.h); kept so callers see the same integers.
enum {
    SSC_ERR_MEM = 55, SSC_ERR_SUP = 56, SSC_ERR_ARG_OUTOFRANGE = 63, SSC_ERR_MAT_LU_ZRPVT = 71,
    SSC_ERR_ARG_WRONGSTATE = 73, SSC_ERR_ARG_INCOMP = 75, SSC_ERR_LIB = 76, SSC_ERR_USER = 83,
    SSC_ERR_ARG_NULL = 85, SSC_ERR_GPU = 97
};

int ssc_set_error(int code, const char *fmt, ...);

struct HostPlex {
    i64 dim = 0, pEnd = 0;
    std::vector<i64> coneOff, cones, suppOff, supports, depthStart, depthEnd;
    std::vector<unsigned char> ghost;
    bool set = false;
};

struct Discretisation {
    i64 nsub = 0, totalDofsPerCell = 0;
    std::vector<std::vector<i64>> secDof, secOff;
    std::vector<i64> bs, nodesPerCell, subspaceOffsets;
    std::vector<const i64 *> cellNodeMap;      // borrowed, ssc/libssc.c:268
    std::vector<i64> ghostBcNodes, globalBcNodes;
    std::vector<i64> cnDof, cnOff;             // cell numbering section
    bool set = false, cn_set = false;
};

struct BuildOptions {
    int construct_type = 0;                    // PC_PATCH_STAR
    i64 codim = -1, dim = -1, exclude_subspace = -1, vankadim = -1;
    bool keep_dofs = false;
    bool print_patches = false;                // -pc_patch_print_patches, ssc/libssc.c:674-730
    int nthreads = 0;
};

struct UserPatches {
    std::vector<i64> off, pts, iterationSet;
    bool set = false;
};

// A plain array of trivially copyable T whose resize() does not touch the new elements: the patch builder's worker
// threads fill disjoint ranges of the big result arrays themselves (a std::vector would first zero 260 MB on one thread).
template <class T> class RawArray {
    T *p_ = nullptr;
    size_t n_ = 0;
public:
    RawArray() {}
    RawArray(const RawArray &o) { assign(o.begin(), o.end()); }
    RawArray(RawArray &&o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    RawArray &operator=(const RawArray &o) { if (this != &o) assign(o.begin(), o.end()); return *this; }
    RawArray &operator=(RawArray &&o) noexcept { if (this != &o) { std::free(p_); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; } return *this; }
    ~RawArray() { std::free(p_); }
    void resize(size_t n) {                                 // contents up to min(old, new) size are kept, the rest is uninitialised
        if (n == n_) return;
        if (n == 0) { std::free(p_); p_ = nullptr; n_ = 0; return; }
        T *q = static_cast<T *>(std::realloc(p_, n * sizeof(T)));
        if (!q) throw std::bad_alloc();
        p_ = q; n_ = n;
    }
    template <class It> void assign(It first, It last) { resize((size_t)(last - first)); std::copy(first, last, p_); }
    void clear() { resize(0); }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    T *data() { return p_; }
    const T *data() const { return p_; }
    T *begin() { return p_; }
    T *end() { return p_ + n_; }
    const T *begin() const { return p_; }
    const T *end() const { return p_ + n_; }
    T &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }
};

struct PatchLists {
    i64 npatch = 0, vStart = 0;
    std::vector<i64> cellOff, gtolOff;         // offsets of cellCounts (ssc/libssc.c:452-580) and gtolCounts (:779-837)
    RawArray<i64> cells;                       // cells of every patch, renumbered (:887)
    RawArray<i64> gtol;                        // patch -> local dof lists (:779-837)
    RawArray<i64> dofs;                        // per-cell patch-local index table (:846-889), optional
    std::string printout;                      // text of -pc_patch_print_patches
};

// Host-side restatement-free implementation of A1-A5 (see patch_builder.cpp)
int ssc_build_patches(const HostPlex &plex, const Discretisation &disc, const UserPatches &user,
                      const BuildOptions &opt, PatchLists &out);
