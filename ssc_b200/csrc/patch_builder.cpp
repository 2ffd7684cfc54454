// Host-side patch construction: which cells and which dofs make up each patch.
//
// Produces what PCPatchCreateCellPatches (ssc/libssc.c:452-580) and
// PCPatchCreateCellPatchDiscretisationInfo (ssc/libssc.c:598-900) produce -- cellCounts/cells,
// gtolCounts/gtol and the per-cell `dofs` table -- for the star (:1447), Vanka (:1471) and user (:1522)
// constructions.  north_star keeps this stage on the host ("PatchPC.pyx and the patch.py setup logic").
//
// The reference walks the mesh three times per patch with PETSc hash tables.  Here each worker thread owns
// a set of stamp arrays over mesh points and local dofs (stamp == current patch means "member"), so set
// union, membership and difference are O(1) array probes, every patch is visited once, and contiguous
// ranges of patches are built in parallel and concatenated.
//
// Ordering contract (DESIGN.md "dof-list order"): cells of a patch ascend by mesh point number; local dof
// numbers are handed out in first-encounter order walking subspaces -> patch cells -> cell nodes ->
// components (ssc/libssc.c:731-773).
#include "common.h"

#include <algorithm>
#include <thread>
#include <atomic>
#include <cstdlib>

namespace {

// the four per-dof marks share a cache line: a patch touches each of its ~n dofs in all four roles
struct DofMark { int32_t owned, seen, local, idx; };

// calloc, not vector::assign: a worker only ever touches the marks near its own range of patches, and untouched
// zero pages are never faulted in
template <class T> struct ZeroArray {
    T *p = nullptr;
    void init(i64 n) { std::free(p); p = static_cast<T *>(std::calloc((size_t)std::max<i64>(n, 1), sizeof(T))); }
    ~ZeroArray() { std::free(p); }
    T &operator[](i64 i) { return p[i]; }
};

struct Scratch {
    ZeroArray<int32_t> ptOwned, ptStar, ptSeen;        // stamps over mesh points
    ZeroArray<DofMark> dof;                            // stamps over local dofs
    std::vector<i64> owned, star, seen, cells;
    void init(i64 pEnd, i64 nlocal) {
        ptOwned.init(pEnd); ptStar.init(pEnd); ptSeen.init(pEnd);
        dof.init(nlocal);
    }
};

struct Chunk {
    std::vector<i64> cellCount, gtolCount, cells, gtol, dofs;
    std::string printout;
    int err = 0;
    std::string msg;
};

// grow `list` by everything reachable from list[from..] through `off/adj`, marking with `stamp`
inline void bfs(const std::vector<i64> &off, const std::vector<i64> &adj, ZeroArray<int32_t> &mark,
                int32_t stamp, std::vector<i64> &list, size_t from) {
    for (size_t head = from; head < list.size(); ++head) {
        i64 q = list[head];
        for (i64 k = off[q]; k < off[q + 1]; ++k) {
            i64 r = adj[k];
            if (mark[r] != stamp) { mark[r] = stamp; list.push_back(r); }
        }
    }
}

inline void push(ZeroArray<int32_t> &mark, int32_t stamp, std::vector<i64> &list, i64 p) {
    if (mark[p] != stamp) { mark[p] = stamp; list.push_back(p); }
}

struct Builder {
    const HostPlex &plex;
    const Discretisation &disc;
    const UserPatches &user;
    const BuildOptions &opt;
    i64 vStart, vEnd, cStart, cEnd, nlocal;
    std::vector<unsigned char> isGlobalBc;

    // owned points of patch `v` -> s.owned (marks ptOwned)
    int owned_points(i64 v, int32_t stamp, Scratch &s, std::string &msg) const {
        s.owned.clear();
        if (opt.construct_type == 0) {                       // star, ssc/libssc.c:1447-1469
            push(s.ptOwned, stamp, s.owned, v);
            bfs(plex.suppOff, plex.supports, s.ptOwned, stamp, s.owned, 0);
        } else if (opt.construct_type == 1) {                // Vanka, ssc/libssc.c:1471-1519
            push(s.ptOwned, stamp, s.owned, v);
            i64 iStart = 0, iEnd = 0;
            bool ignore = opt.vankadim >= 0;
            if (ignore) { iStart = plex.depthStart[opt.vankadim]; iEnd = plex.depthEnd[opt.vankadim]; }
            // star(v) and the closure of its cells, both through the ptStar stamps (reset below by restamping)
            s.star.clear();
            push(s.ptStar, stamp, s.star, v);
            bfs(plex.suppOff, plex.supports, s.ptStar, stamp, s.star, 0);
            s.seen.clear();
            for (i64 c : s.star) if (c >= cStart && c < cEnd) push(s.ptSeen, stamp, s.seen, c);
            bfs(plex.coneOff, plex.cones, s.ptSeen, stamp, s.seen, 0);
            for (i64 e : s.seen) {
                if (ignore && iStart <= e && e < iEnd) continue;
                push(s.ptOwned, stamp, s.owned, e);
            }
        } else {                                             // user, ssc/libssc.c:1522-1547
            for (i64 i = user.off[v]; i < user.off[v + 1]; ++i) {
                i64 e = user.pts[i];
                if (e < 0 || e >= plex.pEnd) { msg = "Entities need to be between the bounds of DMPlexGetChart()"; return SSC_ERR_USER; }
                push(s.ptOwned, stamp, s.owned, e);
            }
        }
        return 0;
    }

    template <class F> void for_point_dofs(i64 k, i64 p, F f) const {
        i64 ldof = disc.secDof[k][p], loff = disc.secOff[k][p], bs = disc.bs[k], so = disc.subspaceOffsets[k];
        for (i64 j = loff; j < loff + ldof; ++j)
            for (i64 l = 0; l < bs; ++l) f(bs * j + l + so);
    }

    int build_one(i64 v, int32_t stamp, int32_t stamp2, Scratch &s, Chunk &out) const {
        int ierr = owned_points(v, stamp, s, out.msg);
        if (ierr) return ierr;
        // PCPatchCompleteCellPatch ssc/libssc.c:325-363: union over owned points of closure(star(.)).
        // Two multi-source sweeps: supports from every owned point, then cones from everything reached.
        // (stamp2 keeps the Vanka helper marks above from leaking into these sets.)
        s.star.clear();
        for (i64 e : s.owned) push(s.ptStar, stamp2, s.star, e);
        bfs(plex.suppOff, plex.supports, s.ptStar, stamp2, s.star, 0);
        s.seen.clear();
        for (i64 e : s.star) push(s.ptSeen, stamp2, s.seen, e);
        bfs(plex.coneOff, plex.cones, s.ptSeen, stamp2, s.seen, 0);
        s.cells.clear();
        for (i64 e : s.seen) if (cStart <= e && e < cEnd) s.cells.push_back(e);
        std::sort(s.cells.begin(), s.cells.end());
        i64 ncell = (i64)s.cells.size();
        out.cellCount.push_back(ncell);
        if (ncell == 0) { out.gtolCount.push_back(0); return 0; }
        // owned / seen dofs: PCPatchGetPointDofs ssc/libssc.c:370-416; artificial BCs = seen \ owned (:420-440)
        for (i64 k = 0; k < disc.nsub; ++k) {
            if (k == opt.exclude_subspace) {
                for_point_dofs(k, v, [&](i64 g) { s.dof[g].owned = stamp; });
            } else {
                for (i64 p : s.owned) for_point_dofs(k, p, [&](i64 g) { s.dof[g].owned = stamp; });
            }
            for (i64 p : s.seen) for_point_dofs(k, p, [&](i64 g) { s.dof[g].seen = stamp; });
        }
        if (opt.print_patches) {
            // ssc/libssc.c:674-730 (the reference prints in hash order; here every list ascends)
            std::vector<i64> od, sd, gb, ab;
            for (i64 k = 0; k < disc.nsub; ++k)
                for (i64 p : s.seen) for_point_dofs(k, p, [&](i64 g) {
                    sd.push_back(g);
                    if (s.dof[g].owned == stamp) od.push_back(g);
                    else ab.push_back(g);
                    if (isGlobalBc[g]) gb.push_back(g);
                });
            for (i64 k = 0; k < disc.nsub; ++k)        // owned dofs on points outside the seen set (cannot happen for star patches)
                for (i64 p : s.owned) if (s.ptSeen[p] != stamp2) for_point_dofs(k, p, [&](i64 g) { od.push_back(g); });
            auto line = [&](const char *title, std::vector<i64> &v_) {
                std::sort(v_.begin(), v_.end());
                v_.erase(std::unique(v_.begin(), v_.end()), v_.end());
                out.printout += "Patch " + std::to_string(v) + ": " + title + ":\n";
                for (i64 g : v_) out.printout += std::to_string(g) + " ";
                out.printout += "\n";
            };
            line("owned dofs", od); line("seen dofs", sd); line("global BCs", gb); line("artificial BCs", ab);
            out.printout += "\n";
        }
        // first-encounter numbering, ssc/libssc.c:731-778
        size_t g0 = out.gtol.size(), d0 = out.dofs.size();
        int32_t next = 0;
        for (i64 k = 0; k < disc.nsub; ++k) {
            const i64 npc = disc.nodesPerCell[k], so = disc.subspaceOffsets[k], bs = disc.bs[k];
            const i64 *cnm = disc.cellNodeMap[k];
            for (i64 c : s.cells) {
                if (disc.cnDof[c] <= 0) { out.msg = "Cell doesn't appear in cell numbering map"; return SSC_ERR_ARG_OUTOFRANGE; }
                const i64 cell = disc.cnOff[c];
                for (i64 j = 0; j < npc; ++j) {
                    const i64 base = cnm[cell * npc + j] * bs + so;
                    for (i64 l = 0; l < bs; ++l) {
                        const i64 g = base + l;
                        if (g < 0 || g >= nlocal) { out.msg = "cell node map entry outside the local dof range"; return SSC_ERR_ARG_OUTOFRANGE; }
                        i64 loc;
                        DofMark &m = s.dof[g];
                        if (m.local == stamp) loc = m.idx;               // met before in this patch (kept or dropped): one load
                        else {
                            m.local = stamp;
                            if (isGlobalBc[g] || (m.seen == stamp && m.owned != stamp)) m.idx = -1;   // global or artificial BC
                            else { m.idx = next++; out.gtol.push_back(g); }
                            loc = m.idx;
                        }
                        if (opt.keep_dofs && disc.nsub == 1) out.dofs.push_back(loc);
                    }
                }
            }
        }
        out.gtolCount.push_back((i64)(out.gtol.size() - g0));
        if (opt.keep_dofs && disc.nsub > 1) {                // cell-major table, ssc/libssc.c:846-872
            (void)d0;
            for (i64 c : s.cells) {
                const i64 cell = disc.cnOff[c];
                for (i64 k = 0; k < disc.nsub; ++k) {
                    const i64 npc = disc.nodesPerCell[k], so = disc.subspaceOffsets[k], bs = disc.bs[k];
                    const i64 *cnm = disc.cellNodeMap[k];
                    for (i64 j = 0; j < npc; ++j) for (i64 l = 0; l < bs; ++l) {
                        const i64 g = cnm[cell * npc + j] * bs + so + l;
                        out.dofs.push_back(s.dof[g].local == stamp ? s.dof[g].idx : -1);
                    }
                }
            }
        }
        for (i64 c : s.cells) out.cells.push_back(disc.cnOff[c]);   // renumbered, ssc/libssc.c:747, :887
        return 0;
    }
};

}  // namespace

int ssc_build_patches(const HostPlex &plex, const Discretisation &disc, const UserPatches &user,
                      const BuildOptions &opt, PatchLists &out) {
    if (!plex.set) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "DM not yet set on patch PC");      // ssc/libssc.c:474-476
    if (!disc.set) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "PCPatchSetDiscretisationInfo() has not been called");
    if (!disc.cn_set) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "PCPatchSetCellNumbering() has not been called");
    Builder b{plex, disc, user, opt, 0, 0, 0, 0, 0, {}};
    b.cStart = plex.depthStart[plex.dim]; b.cEnd = plex.depthEnd[plex.dim];
    if (opt.construct_type >= 2) {                                                                    // ssc/libssc.c:485-489
        if (!user.set) return ssc_set_error(SSC_ERR_ARG_WRONGSTATE, "user patch construction requested but no patches were supplied");
        b.vStart = 0; b.vEnd = (i64)user.off.size() - 1;
    } else if (opt.codim < 0) {                                                                       // :490-495
        i64 d = opt.dim < 0 ? 0 : opt.dim;
        if (d > plex.dim) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "construction dim %lld exceeds mesh dimension", (long long)d);
        b.vStart = plex.depthStart[d]; b.vEnd = plex.depthEnd[d];
    } else {                                                                                          // :496-498
        if (opt.codim > plex.dim) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "construction codim %lld exceeds mesh dimension", (long long)opt.codim);
        b.vStart = plex.depthStart[plex.dim - opt.codim]; b.vEnd = plex.depthEnd[plex.dim - opt.codim];
    }
    b.nlocal = disc.subspaceOffsets[disc.nsub];
    b.isGlobalBc.assign(b.nlocal, 0);
    for (i64 g : disc.ghostBcNodes) {                                                                 // ssc/libssc.c:642-648
        if (g < 0 || g >= b.nlocal) return ssc_set_error(SSC_ERR_ARG_OUTOFRANGE, "ghost BC node %lld outside the local dof range", (long long)g);
        b.isGlobalBc[g] = 1;
    }
    const i64 np = b.vEnd - b.vStart;
    int T = opt.nthreads > 0 ? opt.nthreads : (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    if (np < 4096) T = 1;
    T = (int)std::min<i64>(T, std::max<i64>(np, 1));
    std::vector<Chunk> chunks(T);
    auto worker = [&](int t) {
        Scratch s; s.init(plex.pEnd, b.nlocal);
        i64 lo = b.vStart + np * t / T, hi = b.vStart + np * (t + 1) / T;
        Chunk &c = chunks[t];
        int32_t stamp = 0;
        for (i64 v = lo; v < hi; ++v) {
            stamp += 2;
            bool ghost = opt.construct_type < 2 && !plex.ghost.empty() && plex.ghost[v];             // :515-521
            if (ghost) { c.cellCount.push_back(0); c.gtolCount.push_back(0); continue; }
            int ierr = b.build_one(v, stamp, stamp + 1, s, c);
            if (ierr) { c.err = ierr; return; }
            if (v - lo == 255 && hi - lo > 1024) {
                // the first 256 patches of the range size the output arrays (20 % headroom): no regrowth copies of a
                // dof list that ends up at tens of megabytes per thread
                const double per = 1.2 * (double)(hi - lo) / 256.0;
                c.gtol.reserve((size_t)(per * (double)c.gtol.size()));
                c.cells.reserve((size_t)(per * (double)c.cells.size()));
                c.dofs.reserve((size_t)(per * (double)c.dofs.size()));
                c.cellCount.reserve((size_t)(hi - lo)); c.gtolCount.reserve((size_t)(hi - lo));
            }
        }
    };
    if (T == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(worker, t);
        for (auto &x : th) x.join();
    }
    for (auto &c : chunks) if (c.err) return ssc_set_error(c.err, "%s", c.msg.c_str());
    out.npatch = np; out.vStart = b.vStart;
    out.cellOff.assign(np + 1, 0); out.gtolOff.assign(np + 1, 0);
    i64 i = 0, ncells = 0, ngtol = 0, ndofs = 0;
    for (auto &c : chunks) {
        for (size_t k = 0; k < c.cellCount.size(); ++k, ++i) {
            out.cellOff[i + 1] = out.cellOff[i] + c.cellCount[k];
            out.gtolOff[i + 1] = out.gtolOff[i] + c.gtolCount[k];
        }
        ncells += (i64)c.cells.size(); ngtol += (i64)c.gtol.size(); ndofs += (i64)c.dofs.size();
    }
    out.printout.clear();
    for (auto &c : chunks) out.printout += c.printout;
    out.cells.resize(ncells); out.gtol.resize(ngtol); out.dofs.resize(ndofs);
    // every chunk copies itself to its place (in parallel when there are several)
    std::vector<i64> ca(chunks.size() + 1, 0), ga(chunks.size() + 1, 0), da(chunks.size() + 1, 0);
    for (size_t t = 0; t < chunks.size(); ++t) {
        ca[t + 1] = ca[t] + (i64)chunks[t].cells.size(); ga[t + 1] = ga[t] + (i64)chunks[t].gtol.size(); da[t + 1] = da[t] + (i64)chunks[t].dofs.size();
    }
    auto place = [&](int t) {
        Chunk &c = chunks[t];
        std::copy(c.cells.begin(), c.cells.end(), out.cells.begin() + ca[t]);
        std::copy(c.gtol.begin(), c.gtol.end(), out.gtol.begin() + ga[t]);
        std::copy(c.dofs.begin(), c.dofs.end(), out.dofs.begin() + da[t]);
        Chunk().cells.swap(c.cells); Chunk().gtol.swap(c.gtol); Chunk().dofs.swap(c.dofs);
    };
    if (T == 1) place(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back(place, t);
        for (auto &x : th) x.join();
    }
    return 0;
}
