// Internal declarations shared by the host-side patch builder and the CUDA engine.
#pragma once
#include <cstdint>
#include <cstdarg>
#include <cstdio>
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

struct PatchLists {
    i64 npatch = 0, vStart = 0;
    std::vector<i64> cellOff, cells;           // cellCounts / cells   (ssc/libssc.c:452-580, renumbered :887)
    std::vector<i64> gtolOff, gtol;            // gtolCounts / gtol    (ssc/libssc.c:779-837)
    std::vector<i64> dofs;                     // per-cell patch-local index table (:846-889), optional
    std::string printout;                      // text of -pc_patch_print_patches
};

// Host-side restatement-free implementation of A1-A5 (see patch_builder.cpp)
int ssc_build_patches(const HostPlex &plex, const Discretisation &disc, const UserPatches &user,
                      const BuildOptions &opt, PatchLists &out);
