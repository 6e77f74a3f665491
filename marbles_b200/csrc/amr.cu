// amr.cu -- multi-box / multi-level ("AMR-exact") levels behind the C ABI.
//
// The reference keeps its AMR hierarchy in AMReX (BoxArray, DistributionMapping, regrid); what crosses the C ABI
// is the box list of a level.  This file owns the device state of such a level (one AMReX-shaped FAB per box,
// library memory or bound AMReX memory), turns the box lists into device copy-tag lists on the host --
//   FabArray::FillBoundary(periodicity)                 AMReX_FabArrayCommI.H:8-253
//   average_down_with_ghosts' two ParallelCopy calls    Source/Utilities.cpp:17, 26 (order of overlapping tags:
//                                                        FabArrayBase::CPC::define, AMReX_FabArrayBase.cpp:328-472)
//   FillPatchTwoLevels' coarse patch + fine ghost regions  AMReX_FillPatchUtil_I.H:564-609, FPinfo
//                                                        (AMReX_FabArrayBase.cpp:1797-1851)
// -- and sequences the kernels of patch.cu in the reference's order.  Box calculus is plain host C++.
#include <algorithm>
#include <array>
#include <cstring>
#include <initializer_list>
#include <map>

#include "internal.cuh"
#include "patch.cuh"

using namespace mbl;

namespace mbl {

namespace {

struct HBox {
    int lo[3], hi[3];
    bool ok() const { return lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]; }
    long long pts() const { return ok() ? (long long)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1) : 0; }
};
HBox isect(const HBox& a, const HBox& b)
{
    HBox r;
    for (int d = 0; d < 3; ++d) {
        r.lo[d] = std::max(a.lo[d], b.lo[d]);
        r.hi[d] = std::min(a.hi[d], b.hi[d]);
    }
    return r;
}
HBox grow(HBox b, int n)
{
    for (int d = 0; d < 3; ++d) b.lo[d] -= n, b.hi[d] += n;
    return b;
}
HBox shifted(HBox b, const int s[3])
{
    for (int d = 0; d < 3; ++d) b.lo[d] += s[d], b.hi[d] += s[d];
    return b;
}
int floor_div2(int a) { return a >> 1; }
HBox coarsen2(HBox b)
{
    for (int d = 0; d < 3; ++d) b.lo[d] = floor_div2(b.lo[d]), b.hi[d] = floor_div2(b.hi[d]);
    return b;
}
// a minus b as at most six disjoint boxes
void box_diff(const HBox& a, const HBox& b, std::vector<HBox>& out)
{
    const HBox c = isect(a, b);
    if (!c.ok()) {
        out.push_back(a);
        return;
    }
    HBox rest = a;
    for (int d = 2; d >= 0; --d) {
        if (rest.lo[d] < c.lo[d]) {
            HBox p = rest;
            p.hi[d] = c.lo[d] - 1;
            out.push_back(p);
            rest.lo[d] = c.lo[d];
        }
        if (rest.hi[d] > c.hi[d]) {
            HBox p = rest;
            p.lo[d] = c.hi[d] + 1;
            out.push_back(p);
            rest.hi[d] = c.hi[d];
        }
    }
}

// Periodicity::shiftIntVect (AMReX_Periodicity.cpp:8-33): x outermost, z innermost
std::vector<std::array<int, 3>> periodic_shifts(const PGeom& G, int nghost)
{
    int per[3] = {0, 0, 0}, jmp[3] = {1, 1, 1};
    for (int d = 0; d < 3; ++d)
        if (G.periodic[d]) {
            const int len = G.dhi[d] - G.dlo[d] + 1;
            per[d] = jmp[d] = len;
            while (per[d] < nghost) per[d] += len;
        }
    std::vector<std::array<int, 3>> r;
    for (int i = -per[0]; i <= per[0]; i += jmp[0])
        for (int j = -per[1]; j <= per[1]; j += jmp[1])
            for (int k = -per[2]; k <= per[2]; k += jmp[2]) r.push_back(std::array<int, 3>{{i, j, k}});
    return r;
}

// order in which BoxArray::intersections visits boxes (AMReX_BoxArray.cpp:1219-1310, hash bins of getHashMap :1547-1598)
std::vector<int> hash_order(const std::vector<HBox>& b)
{
    int maxext[3] = {1, 1, 1};
    for (const HBox& x : b)
        for (int d = 0; d < 3; ++d) maxext[d] = std::max(maxext[d], x.hi[d] - x.lo[d] + 1);
    auto fdiv = [](int a, int m) { return a >= 0 ? a / m : -((-a + m - 1) / m); };
    std::vector<int> idx(b.size());
    for (size_t n = 0; n < b.size(); ++n) idx[n] = (int)n;
    std::stable_sort(idx.begin(), idx.end(), [&](int p, int q) {
        for (int d = 2; d >= 0; --d) {
            const int kp = fdiv(b[p].lo[d], maxext[d]), kq = fdiv(b[q].lo[d], maxext[d]);
            if (kp != kq) return kp < kq;
        }
        return p < q;
    });
    return idx;
}

struct BoxSet {  // a table of FABs with its own memory (a level, or the auxiliary coarse boxes between two levels)
    std::vector<PBox> h;
    PBox* d = nullptr;   // every box of the level, indexed by box id (copy tags, regions); boxes of other ranks are empty
    PBox* dl = nullptr;  // the boxes that live here, packed: what the one-launch-per-level kernels walk (blockIdx.y)
    int nl = 0;
    long long max_face = 0;  // cells of the largest 3-cell-thick face slab of a local box (grid size of the BC kernels)
    double* pool_f = nullptr;  // aux sets: one f and one g buffer per box
    double* pool_g = nullptr;
    long long max_cells = 0;
    void free_all()
    {
        if (d) cudaFree(d);
        if (dl) cudaFree(dl);
        if (pool_f) cudaFree(pool_f);
        if (pool_g) cudaFree(pool_g);
        d = dl = nullptr, pool_f = pool_g = nullptr, nl = 0;
        h.clear();
    }
};

// the tags of one peer's messages: what this rank sends (its boxes are the sources) and receives (destinations)
struct PeerPart {
    int peer = -1;
    CopyTag* d_send = nullptr;
    CopyTag* d_recv = nullptr;
    int nsend = 0, nrecv = 0;
    long long send_cells = 0, recv_cells = 0, send_max = 0, recv_max = 0;
};
struct Tags {
    CopyTag* d = nullptr;  // tags whose two boxes live on this rank
    int n = 0;
    long long max_cells = 0;
    std::vector<PeerPart> peers;
    void free_all()
    {
        if (d) cudaFree(d);
        d = nullptr, n = 0;
        for (PeerPart& p : peers) {
            if (p.d_send) cudaFree(p.d_send);
            if (p.d_recv) cudaFree(p.d_recv);
        }
        peers.clear();
    }
};

PBox make_pbox(const HBox& valid, int ng)
{
    PBox b;
    memset(&b, 0, sizeof(b));
    for (int d = 0; d < 3; ++d) {
        b.lo[d] = valid.lo[d], b.hi[d] = valid.hi[d];
        b.glo[d] = valid.lo[d] - ng;
        b.n[d] = valid.hi[d] - valid.lo[d] + 1 + 2 * ng;
    }
    b.sy = b.n[0];
    b.sz = (long long)b.n[0] * b.n[1];
    b.sq = b.sz * b.n[2];
    return b;
}

int to_device(const std::vector<CopyTag>& v, CopyTag*& d)
{
    d = nullptr;
    if (v.empty()) return 0;
    CU(cudaMalloc(&d, v.size() * sizeof(CopyTag)));
    CU(cudaMemcpy(d, v.data(), v.size() * sizeof(CopyTag), cudaMemcpyHostToDevice));
    return 0;
}

// Splits a tag list (global box ids, the same list in the same order on every rank) by where the two boxes live:
// both here -> local copies; source here -> a piece of the message to the destination's rank; destination here -> a
// piece of the message from the source's rank.  Sender and receiver walk the same sub-sequence of the list, so the
// offsets inside a message agree without any handshake.  downer / sowner == nullptr: every box lives here.
int upload_tags(const std::vector<CopyTag>& v, Tags& t, const std::vector<int>* downer = nullptr,
                const std::vector<int>* sowner = nullptr, int rank = 0)
{
    t.free_all();
    std::vector<CopyTag> local;
    std::map<int, std::pair<std::vector<CopyTag>, std::vector<CopyTag>>> remote;
    std::map<int, std::pair<long long, long long>> cells;
    t.max_cells = 0;
    for (CopyTag c : v) {
        const int od = downer ? (*downer)[c.dbox] : rank, os = sowner ? (*sowner)[c.sbox] : rank;
        const long long nc = (long long)c.n[0] * c.n[1] * c.n[2];
        c.boff = 0;
        if (od == rank && os == rank) {
            local.push_back(c);
            t.max_cells = std::max(t.max_cells, nc);
        } else if (os == rank) {
            c.boff = cells[od].first;
            cells[od].first += nc;
            remote[od].first.push_back(c);
        } else if (od == rank) {
            c.boff = cells[os].second;
            cells[os].second += nc;
            remote[os].second.push_back(c);
        }
    }
    t.n = (int)local.size();
    if (to_device(local, t.d)) return 1;
    for (auto& kv : remote) {
        PeerPart p;
        p.peer = kv.first;
        p.nsend = (int)kv.second.first.size(), p.nrecv = (int)kv.second.second.size();
        p.send_cells = cells[kv.first].first, p.recv_cells = cells[kv.first].second;
        for (const CopyTag& c : kv.second.first) p.send_max = std::max(p.send_max, (long long)c.n[0] * c.n[1] * c.n[2]);
        for (const CopyTag& c : kv.second.second) p.recv_max = std::max(p.recv_max, (long long)c.n[0] * c.n[1] * c.n[2]);
        if (to_device(kv.second.first, p.d_send) || to_device(kv.second.second, p.d_recv)) return 1;
        t.peers.push_back(p);
    }
    return 0;
}

CopyTag make_tag(int dbox, int sbox, const HBox& dst_region, const int shift[3])
{
    CopyTag t;
    t.dbox = dbox, t.sbox = sbox;
    for (int d = 0; d < 3; ++d) {
        t.d[d] = dst_region.lo[d];
        t.s[d] = dst_region.lo[d] - shift[d];
        t.n[d] = dst_region.hi[d] - dst_region.lo[d] + 1;
    }
    return t;
}

// auxiliary coarse boxes between a fine level and the next coarser one
struct Inter {
    bool valid = false;
    BoxSet avg[2];        // [ng]: coarsened fine boxes grown by ng (masked average-down, ng = 0 and 1)
    Tags a2c[2];          // aux (grown) -> coarse valid, overlaps resolved to "last tag wins"
    BoxSet cpatch;        // coarsen(grown fine box) grown by 1 (coarse-fine interpolation)
    Tags c2p;
    RegionTag* d_regs = nullptr;
    int nregs = 0;
    long long reg_max = 0;
    void free_all()
    {
        for (int n = 0; n < 2; ++n) avg[n].free_all(), a2c[n].free_all();
        cpatch.free_all();
        c2p.free_all();
        if (d_regs) cudaFree(d_regs);
        d_regs = nullptr, nregs = 0, valid = false;
    }
};

}  // namespace

struct PatchLevel {
    int lev = 0;
    PGeom G;
    Phys P;
    BcInfo B;
    mbl_level_geom geom;
    std::vector<HBox> boxes;  // ALL boxes of the level (every rank holds the list)
    std::vector<int> owner;   // rank that holds each box (AMReX DistributionMapping)
    int rank = 0;
    bool local(int n) const { return owner[n] == rank; }
    BoxSet set;          // the level's FABs (set.pool_* unused: separate pools below); boxes of other ranks have no memory
    int cur = 0;
    double* pool_f[2] = {nullptr, nullptr};
    double* pool_g[2] = {nullptr, nullptr};
    double* pool_qc = nullptr;
    double* pool_macro = nullptr;
    int32_t* pool_isfl = nullptr;
    double* d_red = nullptr;  // 3 doubles: reduction target of compute_eb_forces
    std::vector<long long> cell_off;  // first cell of each box in the pools
    long long total_cells = 0;
    bool any_bound = false;
    bool dq_from_macro = false;
    Tags fb3, fb1;       // FillBoundary over 3 ghost cells (f, g) and 1 (QCorr, macrodata)
    Inter inter;         // this level as the FINE level of (lev - 1, lev)
    int generation = 0, inter_coarse_generation = -1;
};

namespace {

// device copy of a box table: a box that lives on another rank has no memory and is EMPTY there (no cells to loop
// over), so the kernels, which take the whole table, skip it without knowing about ranks
PBox device_view(const PBox& b)
{
    if (b.f[0]) return b;
    PBox e = b;
    for (int d = 0; d < 3; ++d) e.lo[d] = 0, e.hi[d] = -1, e.n[d] = 0;
    e.sy = e.sz = e.sq = 0;
    return e;
}
int upload_table(BoxSet& S)
{
    std::vector<PBox> v(S.h.size()), loc;
    for (size_t n = 0; n < S.h.size(); ++n) {
        v[n] = device_view(S.h[n]);
        if (S.h[n].f[0]) {
            loc.push_back(S.h[n]);
            const long long a = S.h[n].n[0], b = S.h[n].n[1], c = S.h[n].n[2];
            S.max_face = std::max(S.max_face, PNG * std::max(a * b, std::max(a * c, b * c)));
        }
    }
    if (!S.d) CU(cudaMalloc(&S.d, std::max<size_t>(v.size(), 1) * sizeof(PBox)));
    if (!S.dl) CU(cudaMalloc(&S.dl, std::max<size_t>(v.size(), 1) * sizeof(PBox)));
    CU(cudaMemcpy(S.d, v.data(), v.size() * sizeof(PBox), cudaMemcpyHostToDevice));
    if (!loc.empty()) CU(cudaMemcpy(S.dl, loc.data(), loc.size() * sizeof(PBox), cudaMemcpyHostToDevice));
    S.nl = (int)loc.size();
    return 0;
}
int sync_table(PatchLevel& L) { return upload_table(L.set); }

void build_fill_boundary(const PatchLevel& L, int ng, std::vector<CopyTag>& out)
{
    const auto shifts = periodic_shifts(L.G, ng);
    const int nb = (int)L.boxes.size();
    for (int i = 0; i < nb; ++i) {
        const HBox gb = grow(L.boxes[i], ng);
        for (const auto& s : shifts) {
            const int sh[3] = {s[0], s[1], s[2]};
            for (int j = 0; j < nb; ++j) {
                if (i == j && !sh[0] && !sh[1] && !sh[2]) continue;
                const HBox r = isect(gb, shifted(L.boxes[j], sh));
                if (!r.ok()) continue;
                // the ghost cells only: a periodic image may fall on valid cells of the box itself when the domain is
                // narrower than the box is wide -- those keep their values
                std::vector<HBox> pieces;
                box_diff(r, L.boxes[i], pieces);
                for (const HBox& p : pieces) out.push_back(make_tag(i, j, p, sh));
            }
        }
    }
}

// auxiliary boxes (one per fine box, living where the fine box lives)
int alloc_aux(BoxSet& S, const std::vector<HBox>& valid, int ng, const std::vector<int>& owner, int rank)
{
    S.free_all();
    long long total = 0;
    for (size_t n = 0; n < valid.size(); ++n) {
        PBox b = make_pbox(valid[n], ng);
        if (owner[n] == rank) {
            S.max_cells = std::max(S.max_cells, b.sq);
            total += b.sq;
        }
        S.h.push_back(b);
    }
    CU(cudaMalloc(&S.pool_f, (size_t)std::max(total, 1LL) * NQ * sizeof(double)));
    CU(cudaMalloc(&S.pool_g, (size_t)std::max(total, 1LL) * NQ * sizeof(double)));
    CU(cudaMemset(S.pool_f, 0, (size_t)std::max(total, 1LL) * NQ * sizeof(double)));
    CU(cudaMemset(S.pool_g, 0, (size_t)std::max(total, 1LL) * NQ * sizeof(double)));
    long long off = 0;
    for (size_t n = 0; n < S.h.size(); ++n) {
        if (owner[n] != rank) continue;
        PBox& b = S.h[n];
        b.f[0] = b.f[1] = S.pool_f + off * NQ;
        b.g[0] = b.g[1] = S.pool_g + off * NQ;
        off += b.sq;
    }
    return upload_table(S);
}

// coarse patches + fine regions of the coarse-fine interpolation into the boxes `fine` (grown by the ghost width):
// every cell inside dstdomain that no box of `cover` contains.  cover = the level's own valid boxes for a ghost fill
// (FillPatchTwoLevels, mf == fine level), the OLD level's boxes when a re-made level is filled (RemakeLevel).
int build_interp(Inter& I, const std::vector<HBox>& fine, const std::vector<int>& fine_owner, const PGeom& FG, int fine_lev,
                 const PatchLevel& Cl, const std::vector<HBox>& cover)
{
    const int nf = (int)fine.size(), nc = (int)Cl.boxes.size();
    const int rank = Cl.rank;
    // coarse patches for the interpolation: coarsen(grown fine box) grown by 1 (CellConservativeLinear::CoarseBox)
    std::vector<HBox> cp(nf);
    for (int n = 0; n < nf; ++n) cp[n] = grow(coarsen2(grow(fine[n], PNG)), 1);
    if (alloc_aux(I.cpatch, cp, 0, fine_owner, rank)) return 1;
    // fine ghost regions the interpolation fills: grown box inside dstdomain (the domain grown by the ghost width in
    // periodic directions) minus the fine level's valid boxes, NOT periodically shifted (FPinfo: complementIn)
    std::vector<RegionTag> regs;
    std::vector<std::vector<HBox>> needed(nf);  // coarse cells the interpolation of box n reads (disjoint pieces)
    const auto shifts_f = periodic_shifts(FG, PNG);
    for (int n = 0; n < nf; ++n) {
        HBox r0 = grow(fine[n], PNG);
        for (int d = 0; d < 3; ++d)
            if (!FG.periodic[d]) {
                r0.lo[d] = std::max(r0.lo[d], FG.dlo[d]);
                r0.hi[d] = std::min(r0.hi[d], FG.dhi[d]);
            }
        std::vector<HBox> list{r0};
        for (const HBox& v : cover) {
            std::vector<HBox> next;
            for (const HBox& r : list) box_diff(r, v, next);
            list.swap(next);
        }
        for (const HBox& r : list) {
            // Slopes of coarse cells on a non-periodic domain face are one-sided and depend on the extents of AMReX's
            // internal coarse patch and on what its BCFill put beyond the face (AMReX_MFInterp_C.H:15-33): not
            // reproduced.  Only cells that SURVIVE matter -- the copy from the fine level that follows the interpolation
            // (FillPatchSingleLevel, periodic images included) overwrites every cell a box of `cover` lies on -- so a
            // fine box may touch a non-periodic face as long as no coarse-fine interface cell has its parent there.
            bool near_face = false;
            for (int d = 0; d < 3; ++d)
                if (!FG.periodic[d] && (floor_div2(r.lo[d]) <= Cl.G.dlo[d] || floor_div2(r.hi[d]) >= Cl.G.dhi[d])) near_face = true;
            if (near_face) {
                std::vector<HBox> surv{r};
                for (const auto& s : shifts_f) {
                    const int sh[3] = {s[0], s[1], s[2]};
                    for (const HBox& v : cover) {
                        std::vector<HBox> next;
                        for (const HBox& p : surv) box_diff(p, shifted(v, sh), next);
                        surv.swap(next);
                    }
                }
                for (const HBox& p : surv)
                    for (int d = 0; d < 3; ++d)
                        if (!FG.periodic[d] && (floor_div2(p.lo[d]) <= Cl.G.dlo[d] || floor_div2(p.hi[d]) >= Cl.G.dhi[d]))
                            return fail("level %d: a coarse-fine interface lies within one coarse cell of a non-periodic "
                                        "domain face; the interpolation there is not supported", fine_lev);
            }
            {   // parents of the region's cells and their 26 neighbours (slopes, min / max limiting)
                std::vector<HBox> add{isect(grow(coarsen2(r), 1), cp[n])};
                for (const HBox& have : needed[n]) {
                    std::vector<HBox> next;
                    for (const HBox& a : add) box_diff(a, have, next);
                    add.swap(next);
                }
                for (const HBox& a : add) needed[n].push_back(a);
            }
            if (fine_owner[n] != rank) continue;  // (the check above runs for every box on every rank: same verdict everywhere)
            RegionTag t;
            t.box = n;
            for (int d = 0; d < 3; ++d) t.lo[d] = r.lo[d], t.n[d] = r.hi[d] - r.lo[d] + 1;
            regs.push_back(t);
            I.reg_max = std::max(I.reg_max, r.pts());
        }
    }
    // the coarse level's valid cells (periodic images included) -> the pieces of the coarse patches that are read; a
    // fine box in the interior of its level has no region and gets nothing
    std::vector<CopyTag> c2p;
    const auto shifts_p = periodic_shifts(Cl.G, PNG + 2);
    for (int n = 0; n < nf; ++n)
        for (const HBox& piece : needed[n])
            for (const auto& s : shifts_p) {
                const int sh[3] = {s[0], s[1], s[2]};
                for (int j = 0; j < nc; ++j) {
                    const HBox r = isect(piece, shifted(Cl.boxes[j], sh));
                    if (r.ok()) c2p.push_back(make_tag(n, j, r, sh));
                }
            }
    if (upload_tags(c2p, I.c2p, &fine_owner, &Cl.owner, rank)) return 1;
    if (I.d_regs) cudaFree(I.d_regs);
    I.d_regs = nullptr;
    I.nregs = (int)regs.size();
    if (I.nregs) {
        CU(cudaMalloc(&I.d_regs, regs.size() * sizeof(RegionTag)));
        CU(cudaMemcpy(I.d_regs, regs.data(), regs.size() * sizeof(RegionTag), cudaMemcpyHostToDevice));
    }
    return 0;
}

// everything between level F.lev - 1 (coarse) and F.lev (fine), ratio 2
int build_inter(PatchLevel& F, const PatchLevel& Cl)
{
    Inter& I = F.inter;
    I.free_all();
    const int nf = (int)F.boxes.size(), nc = (int)Cl.boxes.size();
    for (const HBox& b : F.boxes)
        for (int d = 0; d < 3; ++d)
            if ((b.lo[d] & 1) || !(b.hi[d] & 1)) return fail("fine boxes must be aligned to the refinement ratio 2");
    std::vector<HBox> cfine(nf);
    for (int n = 0; n < nf; ++n) cfine[n] = coarsen2(F.boxes[n]);
    const std::vector<int> order = hash_order(F.boxes);
    for (int ng = 0; ng < 2; ++ng) {
        if (alloc_aux(I.avg[ng], cfine, ng, F.owner, F.rank)) return 1;
        // (1) cfine.ParallelCopy(crse, src ng 0, dst ng) is not made: its only effect is that a cell whose eight fine
        // values are all masked hands the coarse cell its own value back (k_patch_avgdown's KEEP marker does that).  The
        // reference's copy is not periodic, and ring cells outside the domain stay uninitialised there
        // (oracle/amr_oracle.py); here they keep the coarse value as well
        // (2) crse.ParallelCopy(cfine, src ng, dst ng 0, periodicity): tags in CPC order; where the rings of two fine
        // boxes overlap the LAST tag wins, so earlier tags are cut back to what later ones leave
        std::vector<CopyTag> a2c;
        const auto shifts_c = periodic_shifts(Cl.G, 0);
        for (int j = 0; j < nc; ++j) {
            std::vector<std::pair<HBox, CopyTag>> live;  // disjoint destination regions of box j
            for (const auto& s : shifts_c) {
                const int sh[3] = {s[0], s[1], s[2]};
                for (int n : order) {
                    const HBox r = isect(Cl.boxes[j], shifted(grow(cfine[n], ng), sh));
                    if (!r.ok()) continue;
                    std::vector<std::pair<HBox, CopyTag>> next;
                    for (auto& e : live) {
                        std::vector<HBox> pieces;
                        box_diff(e.first, r, pieces);
                        for (const HBox& p : pieces) {
                            const int esh[3] = {e.second.d[0] - e.second.s[0], e.second.d[1] - e.second.s[1],
                                                e.second.d[2] - e.second.s[2]};
                            next.push_back({p, make_tag(e.second.dbox, e.second.sbox, p, esh)});
                        }
                    }
                    next.push_back({r, make_tag(j, n, r, sh)});
                    live.swap(next);
                }
            }
            for (auto& e : live) a2c.push_back(e.second);
        }
        if (upload_tags(a2c, I.a2c[ng], &Cl.owner, &F.owner, F.rank)) return 1;
    }
    if (build_interp(I, F.boxes, F.owner, F.G, F.lev, Cl, F.boxes)) return 1;
    I.valid = true;
    F.inter_coarse_generation = Cl.generation;
    return 0;
}

int ensure_inter(mbl_ctx* ctx, int fine_lev)
{
    if (fine_lev < 1 || !ctx->plev[fine_lev] || !ctx->plev[fine_lev - 1])
        return fail("levels %d and %d must both be multi-box levels (mbl_level_define_boxes)", fine_lev - 1, fine_lev);
    PatchLevel& F = *ctx->plev[fine_lev];
    const PatchLevel& Cl = *ctx->plev[fine_lev - 1];
    if (F.inter.valid && F.inter_coarse_generation == Cl.generation) return 0;
    CU(cudaStreamSynchronize(ctx->stream));
    return build_inter(F, Cl);
}

// One tag list applied to one or more arrays: the local tags as copies, the remote ones as one message per peer that
// holds all the arrays (pack, the caller's exchange, unpack)
struct ArrPair {
    int darr, sarr, ncomp;
};
int run_copy(mbl_ctx* ctx, const PBox* dtab, int dcur, const PBox* stab, int scur, const Tags& t,
             std::initializer_list<ArrPair> arrs, bool skip_keep = false)
{
    cudaStream_t st = ctx->stream;
    for (const ArrPair& a : arrs)
        ctx->launches += launch_patch_copy(dtab, dcur, stab, scur, t.d, t.n, a.darr, a.sarr, a.ncomp, t.max_cells, st, skip_keep);
    if (t.peers.empty()) return 0;
    if (!ctx->exchange) return fail("a distributed level needs mbl_set_exchange");
    int ncomp = 0;
    for (const ArrPair& a : arrs) ncomp += a.ncomp;
    if ((int)ctx->peer_buf.size() < ctx->world) ctx->peer_buf.resize(ctx->world);
    std::vector<int> peers;
    std::vector<double*> send, recv;
    std::vector<int64_t> nsend, nrecv;
    for (const PeerPart& p : t.peers) {
        if (p.peer < 0 || p.peer >= ctx->world) return fail("owner rank %d outside the world of %d ranks", p.peer, ctx->world);
        mbl_ctx::PeerBuf& b = ctx->peer_buf[p.peer];
        const long long ns = p.send_cells * ncomp, nr = p.recv_cells * ncomp;
        if (ns > b.cap_send) {
            CU(cudaStreamSynchronize(st));
            if (b.send) cudaFree(b.send);
            CU(cudaMalloc(&b.send, (size_t)ns * sizeof(double)));
            b.cap_send = ns;
        }
        if (nr > b.cap_recv) {
            CU(cudaStreamSynchronize(st));
            if (b.recv) cudaFree(b.recv);
            CU(cudaMalloc(&b.recv, (size_t)nr * sizeof(double)));
            b.cap_recv = nr;
        }
        long long base = 0;
        for (const ArrPair& a : arrs) {
            ctx->launches += launch_patch_pack(stab, scur, p.d_send, p.nsend, a.sarr, a.ncomp, p.send_max, b.send, base, true, st);
            base += p.send_cells * a.ncomp;
        }
        peers.push_back(p.peer);
        send.push_back(b.send), recv.push_back(b.recv);
        nsend.push_back(ns), nrecv.push_back(nr);
    }
    if (ctx->exchange(ctx->exchange_user, (int)peers.size(), peers.data(), send.data(), nsend.data(), recv.data(), nrecv.data(),
                      (void*)st))
        return fail("the caller's exchange function failed");
    for (const PeerPart& p : t.peers) {
        long long base = 0;
        for (const ArrPair& a : arrs) {
            ctx->launches += launch_patch_pack(dtab, dcur, p.d_recv, p.nrecv, a.darr, a.ncomp, p.recv_max, ctx->peer_buf[p.peer].recv,
                                               base, false, st, skip_keep);
            base += p.recv_cells * a.ncomp;
        }
    }
    return 0;
}

// FabArray::FillBoundary(periodicity) of one or more arrays of a level over ng ghost cells
int fill_boundary(mbl_ctx* ctx, PatchLevel& L, std::initializer_list<ArrPair> arrs, int ng)
{
    return run_copy(ctx, L.set.d, L.cur, L.set.d, L.cur, ng == 1 ? L.fb1 : L.fb3, arrs);
}
int fill_boundary(mbl_ctx* ctx, PatchLevel& L, int arr, int ncomp, int ng) { return fill_boundary(ctx, L, {{arr, arr, ncomp}}, ng); }
int fill_boundary_fg(mbl_ctx* ctx, PatchLevel& L) { return fill_boundary(ctx, L, {{PA_F, PA_F, NQ}, {PA_G, PA_G, NQ}}, PNG); }

int ensure_macro(mbl_ctx* ctx, PatchLevel& L)
{
    if (L.pool_macro) return 0;
    const size_t tc = (size_t)std::max(L.total_cells, 1LL);
    CU(cudaMalloc(&L.pool_macro, tc * NMACRO_ALL * sizeof(double)));
    CU(cudaMemsetAsync(L.pool_macro, 0, tc * NMACRO_ALL * sizeof(double), ctx->stream));
    for (size_t n = 0; n < L.set.h.size(); ++n)
        if (L.local((int)n)) L.set.h[n].macro = L.pool_macro + L.cell_off[n] * NMACRO_ALL;
    CU(cudaStreamSynchronize(ctx->stream));
    return sync_table(L);
}

PatchLevel* plevel(mbl_ctx* ctx, int lev)
{
    if (!ctx || lev < 0 || lev >= MAX_LEVELS || !ctx->plev[lev]) {
        fail("level %d is not a multi-box level (mbl_level_define_boxes)", lev);
        return nullptr;
    }
    return ctx->plev[lev];
}

// FAB <-> FAB copy of ncomp components between a caller's array (ghost width ng_fab) and a device box (ghost
// width PNG): the cells both have, one pitched DMA per component
int copy_fab(const PBox& b, double* dev, int ncomp, double* fab, int ng_fab, bool to_device, cudaStream_t st)
{
    const int g = std::min(ng_fab, PNG);
    const int nx = b.hi[0] - b.lo[0] + 1, ny = b.hi[1] - b.lo[1] + 1, nz = b.hi[2] - b.lo[2] + 1;
    const size_t fx = nx + 2 * ng_fab, fy = ny + 2 * ng_fab, fn = fx * fy * (nz + 2 * ng_fab);
    for (int q = 0; q < ncomp; ++q) {
        cudaMemcpy3DParms p;
        memset(&p, 0, sizeof(p));
        cudaPitchedPtr host = make_cudaPitchedPtr(fab + (size_t)q * fn, fx * sizeof(double), fx, fy);
        cudaPitchedPtr devp = make_cudaPitchedPtr(dev + (size_t)q * b.sq, b.n[0] * sizeof(double), b.n[0], b.n[1]);
        const cudaPos hpos = make_cudaPos((size_t)(ng_fab - g) * sizeof(double), ng_fab - g, ng_fab - g);
        const cudaPos dpos = make_cudaPos((size_t)(PNG - g) * sizeof(double), PNG - g, PNG - g);
        p.srcPtr = to_device ? host : devp;
        p.srcPos = to_device ? hpos : dpos;
        p.dstPtr = to_device ? devp : host;
        p.dstPos = to_device ? dpos : hpos;
        p.extent = make_cudaExtent((size_t)(nx + 2 * g) * sizeof(double), ny + 2 * g, nz + 2 * g);
        p.kind = cudaMemcpyDefault;
        CU(cudaMemcpy3DAsync(&p, st));
    }
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// patch-level counterparts of the per-level entry points (called from api.cu)
// ---------------------------------------------------------------------------------------------------------
int patch_clear(mbl_ctx* ctx, int lev)
{
    PatchLevel* L = ctx->plev[lev];
    if (!L) return 0;
    cudaStreamSynchronize(ctx->stream);
    for (int n = 0; n < 2; ++n) {
        if (L->pool_f[n]) cudaFree(L->pool_f[n]);
        if (L->pool_g[n]) cudaFree(L->pool_g[n]);
    }
    if (L->pool_qc) cudaFree(L->pool_qc);
    if (L->pool_macro) cudaFree(L->pool_macro);
    if (L->pool_isfl) cudaFree(L->pool_isfl);
    if (L->d_red) cudaFree(L->d_red);
    L->set.free_all();
    L->fb3.free_all();
    L->fb1.free_all();
    L->inter.free_all();
    delete L;
    ctx->plev[lev] = nullptr;
    return 0;
}

int patch_initialize(mbl_ctx* ctx, int lev, const IcInfo& I)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    ctx->launches += launch_patch_initialize(L->set.dl, L->set.nl, L->set.max_cells, L->cur, L->B, I, ctx->stream);
    // initialize_f ends with FillBoundary of f and g (LBM.cpp:1209-1210)
    if (fill_boundary_fg(ctx, *L)) return 1;
    CU(cudaGetLastError());
    return 0;
}

int patch_physbc(mbl_ctx* ctx, int lev, double /*time*/)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    ctx->launches += launch_patch_physbc(L->set.dl, L->set.nl, L->set.max_face, L->cur, L->G, L->B, ctx->stream);
    CU(cudaGetLastError());
    return 0;
}

// FillPatchOps::fillpatch(lev, time, m_f[lev]) and the same for g (FillPatchOps.H:75-132)
int patch_fillpatch(mbl_ctx* ctx, int lev, double time)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    const int nb = (int)L->boxes.size();
    cudaStream_t st = ctx->stream;
    const bool all_periodic = L->G.periodic[0] && L->G.periodic[1] && L->G.periodic[2];
    ctx->launches += launch_patch_prepass(L->set.dl, L->set.nl, L->set.max_cells, L->cur, L->G, st);  // K6
    if (lev > 0) {
        // FillPatchTwoLevels: coarse patch from the coarse level's valid cells, CellConservativeLinear into the fine
        // ghost cells no fine valid cell covers
        if (ensure_inter(ctx, lev)) return 1;
        PatchLevel& Cl = *ctx->plev[lev - 1];
        Inter& I = L->inter;
        if (run_copy(ctx, I.cpatch.d, 0, Cl.set.d, Cl.cur, I.c2p, {{PA_F, PA_F, NQ}, {PA_G, PA_G, NQ}})) return 1;
        ctx->launches += launch_patch_interp(L->set.d, L->cur, I.cpatch.d, I.d_regs, I.nregs, I.reg_max, st);
    }
    if (fill_boundary_fg(ctx, *L)) return 1;
    (void)all_periodic;
    CU(cudaGetLastError());
    return patch_physbc(ctx, lev, time);
}

// LBM::stream(lev, m_f); LBM::stream(lev, m_g) on the grown boxes (LBM.cpp:558-604)
int patch_stream(mbl_ctx* ctx, int lev)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    cudaStream_t st = ctx->stream;
    ctx->launches += launch_patch_stream(L->set.dl, L->set.nl, L->set.max_cells, L->cur, st);
    if (L->any_bound) {
        // bound FABs stay the current buffers: copy the streamed state back (MultiFab::Copy, LBM.cpp:601)
        for (const PBox& b : L->set.h) {
            if (!b.f[0]) continue;
            CU(cudaMemcpyAsync(b.f[L->cur], b.f[1 - L->cur], (size_t)b.sq * NQ * sizeof(double), cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(b.g[L->cur], b.g[1 - L->cur], (size_t)b.sq * NQ * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
    } else {
        L->cur = 1 - L->cur;
    }
    if (fill_boundary_fg(ctx, *L)) return 1;  // LBM.cpp:603
    CU(cudaGetLastError());
    return 0;
}

static int patch_macro_pass(mbl_ctx* ctx, PatchLevel& L, int want_macro, bool pull = false)
{
    const int nb = (int)L.boxes.size();
    if (want_macro && ensure_macro(ctx, L)) return 1;
    ctx->launches += launch_patch_qcorr(L.set.dl, L.set.nl, L.set.max_cells, L.cur, L.P, want_macro, ctx->stream, pull);
    // m_macrodata.FillBoundary, LBM.cpp:905 (the comps the gradient reads; all 19 when they are stored)
    if (want_macro) return fill_boundary(ctx, L, {{PA_QC, PA_QC, 3}, {PA_MACRO, PA_MACRO, MBL_NMACRO}}, 1);
    return fill_boundary(ctx, L, PA_QC, 3, 1);
}

// LBM::collide(lev) (LBM.cpp:607-618) on the streamed state, in place
int patch_collide(mbl_ctx* ctx, int lev, int want_macro)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    if (patch_macro_pass(ctx, *L, want_macro)) return 1;
    if (want_macro) L->dq_from_macro = false;
    ctx->launches += launch_patch_collide(L->set.dl, L->set.nl, L->set.max_cells, L->cur, L->G, L->P, want_macro,
                                          ctx->stream);
    if (fill_boundary_fg(ctx, *L)) return 1;  // LBM.cpp:805-806
    CU(cudaGetLastError());
    return 0;
}

// LBM::advance(lev) (LBM.cpp:523-544): stream; average_down_to(lev, 1 ghost ring) when a finer level exists; collide.
// On a finest level nothing happens between the stream and the collide, and the two are one pass: the q-corrections
// of the streamed state are computed by pulling (k_patch_qcorr<., true>), then k_patch_advance pulls again and
// collides.  The FillBoundary between the two passes of the reference (LBM.cpp:603) writes ghost cells that nothing
// reads before the FillBoundary after the collide writes them again, so it is dropped.  Bit-identical to the un-fused
// sequence (MBL_AMR_FUSED=0 selects it).
int patch_advance(mbl_ctx* ctx, int lev, int want_macro)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    const bool finest = lev + 1 >= MAX_LEVELS || !ctx->plev[lev + 1];
    static const bool fused_ok = !(getenv("MBL_AMR_FUSED") && atoi(getenv("MBL_AMR_FUSED")) == 0);
    bool fits = true;  // the fused kernels address a FAB with 32-bit element offsets (27 components)
    for (const PBox& b : L->set.h) fits = fits && b.sq * NQ < (1LL << 31);  // (every rank sees every box: same verdict)
    if (!finest || !fused_ok || !fits) {
        if (patch_stream(ctx, lev)) return 1;
        if (!finest && mbl_average_down(ctx, lev, 1)) return 1;
        return patch_collide(ctx, lev, want_macro);
    }
    cudaStream_t st = ctx->stream;
    if (patch_macro_pass(ctx, *L, want_macro, true)) return 1;
    if (want_macro) L->dq_from_macro = false;
    ctx->launches += launch_patch_advance(L->set.dl, L->set.nl, L->set.max_cells, L->cur, L->G, L->P, want_macro, st);
    if (L->any_bound) {
        // bound FABs stay the current buffers (MultiFab::Copy, LBM.cpp:601)
        for (const PBox& b : L->set.h) {
            if (!b.f[0]) continue;
            CU(cudaMemcpyAsync(b.f[L->cur], b.f[1 - L->cur], (size_t)b.sq * NQ * sizeof(double), cudaMemcpyDeviceToDevice, st));
            CU(cudaMemcpyAsync(b.g[L->cur], b.g[1 - L->cur], (size_t)b.sq * NQ * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
    } else {
        L->cur = 1 - L->cur;
    }
    if (fill_boundary_fg(ctx, *L)) return 1;  // LBM.cpp:805-806
    CU(cudaGetLastError());
    return 0;
}

int patch_f_to_macrodata(mbl_ctx* ctx, int lev)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    if (patch_macro_pass(ctx, *L, 1)) return 1;
    L->dq_from_macro = true;
    CU(cudaGetLastError());
    return 0;
}

// this level's part of LBM::compute_eb_forces (LBM.cpp:994-1044) over the boxes that live here.  m_mask is empty in
// every run the reference completes (tests/golden/make_golden.py, amr3_chcyl_forces): no cell is excluded
int patch_eb_forces(mbl_ctx* ctx, int lev, double out[3])
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    if (!L->d_red) CU(cudaMalloc(&L->d_red, 3 * sizeof(double)));
    ctx->launches += launch_patch_eb_forces(L->set.dl, L->set.nl, L->set.max_cells, L->cur, L->d_red, ctx->stream);
    CU(cudaMemcpyAsync(out, L->d_red, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int patch_compute_derived(mbl_ctx* ctx, int lev)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    if (!L->pool_macro) return fail("mbl_compute_derived needs macrodata");
    ctx->launches += launch_patch_derived(L->set.dl, L->set.nl, L->set.max_cells, L->G, L->P, L->dq_from_macro ? 1 : 0,
                                          ctx->stream);
    CU(cudaGetLastError());
    return 0;
}

}  // namespace mbl

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int mbl_set_exchange(mbl_ctx* ctx, int rank, int world, mbl_exchange_fn fn, void* user)
{
    if (!ctx) return fail("null context");
    if (world < 1 || rank < 0 || rank >= world) return fail("mbl_set_exchange: rank %d of %d", rank, world);
    for (int l = 0; l < MAX_LEVELS; ++l)
        if (ctx->plev[l]) return fail("mbl_set_exchange: call it before the levels are defined");
    ctx->rank = rank, ctx->world = world;
    ctx->exchange = fn, ctx->exchange_user = user;
    return 0;
}

// host only (no device, no context): how many cells the FillBoundary over ng ghost cells of a distributed level moves
// between `rank` and every other rank -- the split of the tag list that mbl_level_define_boxes_on makes
int mbl_fill_boundary_plan(int nboxes, const int* lo, const int* hi, const int* owner, const int dom_lo[3], const int dom_hi[3],
                           const int periodic[3], int ng, int rank, int world, int64_t* send_cells, int64_t* recv_cells,
                           int64_t* local_cells)
{
    if (!lo || !hi || !owner || !send_cells || !recv_cells || nboxes < 1 || world < 1) return fail("mbl_fill_boundary_plan: bad argument");
    PatchLevel L;
    for (int d = 0; d < 3; ++d) L.G.dlo[d] = dom_lo[d], L.G.dhi[d] = dom_hi[d], L.G.periodic[d] = periodic[d];
    for (int n = 0; n < nboxes; ++n) {
        HBox b;
        for (int d = 0; d < 3; ++d) b.lo[d] = lo[3 * n + d], b.hi[d] = hi[3 * n + d];
        if (owner[n] < 0 || owner[n] >= world) return fail("box %d: owner %d outside the world of %d ranks", n, owner[n], world);
        L.boxes.push_back(b);
    }
    std::vector<CopyTag> tags;
    build_fill_boundary(L, ng, tags);
    for (int r = 0; r < world; ++r) send_cells[r] = recv_cells[r] = 0;
    int64_t local = 0;
    for (const CopyTag& c : tags) {
        const int64_t nc = (int64_t)c.n[0] * c.n[1] * c.n[2];
        const int od = owner[c.dbox], os = owner[c.sbox];
        if (od == rank && os == rank) local += nc;
        else if (os == rank) send_cells[od] += nc;
        else if (od == rank) recv_cells[os] += nc;
    }
    if (local_cells) *local_cells = local;
    return 0;
}

int mbl_level_define_boxes(mbl_ctx* ctx, int lev, const mbl_level_geom* g, int nboxes, const int* lo, const int* hi)
{
    return mbl_level_define_boxes_on(ctx, lev, g, nboxes, lo, hi, nullptr);
}

int mbl_level_box_owner(mbl_ctx* ctx, int lev, int ibox)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L || ibox < 0 || ibox >= (int)L->boxes.size()) return -1;
    return L->owner[ibox];
}

int mbl_level_define_boxes_on(mbl_ctx* ctx, int lev, const mbl_level_geom* g, int nboxes, const int* lo, const int* hi,
                              const int* owner)
{
    if (!ctx || !g || !lo || !hi) return fail("null argument");
    if (lev < 0 || lev >= MAX_LEVELS) return fail("level %d out of range", lev);
    if (nboxes < 1) return fail("a level needs at least one box");
    CU(cudaSetDevice(ctx->device));
    mbl_level_clear(ctx, lev);
    PatchLevel* L = new PatchLevel();
    L->lev = lev;
    L->rank = ctx->rank;
    L->geom = *g;
    for (int d = 0; d < 3; ++d) {
        L->G.dlo[d] = g->dom_lo[d], L->G.dhi[d] = g->dom_hi[d];
        L->G.periodic[d] = ctx->prm.periodic[d];
    }
    level_phys_bc(ctx->prm, *g, L->P, L->B);
    for (int n = 0; n < nboxes; ++n) {
        HBox b;
        for (int d = 0; d < 3; ++d) b.lo[d] = lo[3 * n + d], b.hi[d] = hi[3 * n + d];
        if (!b.ok()) {
            delete L;
            return fail("box %d is empty", n);
        }
        for (int d = 0; d < 3; ++d)
            if (b.lo[d] < g->dom_lo[d] || b.hi[d] > g->dom_hi[d]) {
                delete L;
                return fail("box %d leaves the level domain in direction %d", n, d);
            }
        for (const HBox& o : L->boxes)
            if (isect(o, b).ok()) {
                delete L;
                return fail("box %d overlaps another box of the level", n);
            }
        const int own = owner ? owner[n] : ctx->rank;
        if (own < 0 || own >= ctx->world) {
            delete L;
            return fail("box %d: owner %d outside the world of %d ranks (mbl_set_exchange)", n, own, ctx->world);
        }
        L->boxes.push_back(b);
        L->owner.push_back(own);
        PBox p = make_pbox(b, PNG);
        L->cell_off.push_back(own == ctx->rank ? L->total_cells : -1);
        if (own == ctx->rank) {
            L->total_cells += p.sq;
            L->set.max_cells = std::max(L->set.max_cells, p.sq);
        }
        L->set.h.push_back(p);
    }
    ctx->plev[lev] = L;
    const size_t lat = (size_t)std::max(L->total_cells, 1LL) * NQ * sizeof(double);
    for (int n = 0; n < 2; ++n) {
        CU(cudaMalloc(&L->pool_f[n], lat));
        CU(cudaMalloc(&L->pool_g[n], lat));
        CU(cudaMemsetAsync(L->pool_f[n], 0, lat, ctx->stream));
        CU(cudaMemsetAsync(L->pool_g[n], 0, lat, ctx->stream));
    }
    const size_t tc = (size_t)std::max(L->total_cells, 1LL);
    CU(cudaMalloc(&L->pool_qc, tc * 3 * sizeof(double)));
    CU(cudaMemsetAsync(L->pool_qc, 0, tc * 3 * sizeof(double), ctx->stream));
    CU(cudaMalloc(&L->pool_isfl, tc * sizeof(int32_t)));
    {
        std::vector<int32_t> ones(tc, 1);  // all fluid until mbl_box_set_is_fluid says otherwise
        CU(cudaMemcpy(L->pool_isfl, ones.data(), ones.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    for (size_t n = 0; n < L->set.h.size(); ++n) {
        if (!L->local((int)n)) continue;  // a box of another rank: no memory, empty on the device
        PBox& p = L->set.h[n];
        const long long o = L->cell_off[n];
        for (int c = 0; c < 2; ++c) p.f[c] = L->pool_f[c] + o * NQ, p.g[c] = L->pool_g[c] + o * NQ;
        p.qc = L->pool_qc + o * 3;
        p.isfl = L->pool_isfl + o;
    }
    if (sync_table(*L)) return 1;
    std::vector<CopyTag> t3, t1;
    build_fill_boundary(*L, PNG, t3);
    build_fill_boundary(*L, 1, t1);
    if (upload_tags(t3, L->fb3, &L->owner, &L->owner, L->rank) || upload_tags(t1, L->fb1, &L->owner, &L->owner, L->rank)) return 1;
    static int generation = 0;
    L->generation = ++generation;
    patch_init_tables();
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return 0;
}

int mbl_level_num_boxes(mbl_ctx* ctx, int lev)
{
    if (!ctx || lev < 0 || lev >= MAX_LEVELS) return -1;
    if (ctx->plev[lev]) return (int)ctx->plev[lev]->boxes.size();
    return ctx->lev[lev].defined ? 1 : 0;
}

int mbl_level_bind(mbl_ctx* ctx, int lev, int ibox, int which, double* device_fab)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    if (ibox < 0 || ibox >= (int)L->boxes.size() || !device_fab) return fail("mbl_level_bind: bad argument");
    if (!L->local(ibox)) return fail("mbl_level_bind: box %d lives on rank %d", ibox, L->owner[ibox]);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (L->cur != 0) {
        // bring the library-owned state into buffer 0 before the table changes: bound FABs are always buffer 0
        for (PBox& b : L->set.h) {
            std::swap(b.f[0], b.f[1]);
            std::swap(b.g[0], b.g[1]);
        }  // (boxes of other ranks: null either way)
        L->cur = 0;
    }
    PBox& b = L->set.h[ibox];
    (which == MBL_G ? b.g[0] : b.f[0]) = device_fab;
    L->any_bound = true;
    return sync_table(*L);
}

int mbl_box_set_is_fluid(mbl_ctx* ctx, int lev, int ibox, const int32_t* fab, int ng)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    if (ibox < 0 || ibox >= (int)L->boxes.size() || !fab) return fail("mbl_box_set_is_fluid: bad argument");
    if (!L->local(ibox)) return fail("mbl_box_set_is_fluid: box %d lives on rank %d", ibox, L->owner[ibox]);
    if (ng < PNG) return fail("is_fluid needs %d ghost cells (got %d): out-of-domain values come from the geometry", PNG, ng);
    CU(cudaSetDevice(ctx->device));
    const PBox& b = L->set.h[ibox];
    const int nx = b.hi[0] - b.lo[0] + 1, ny = b.hi[1] - b.lo[1] + 1;
    const size_t fx = nx + 2 * ng, fy = ny + 2 * ng;
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    p.srcPtr = make_cudaPitchedPtr(const_cast<int32_t*>(fab), fx * sizeof(int32_t), fx, fy);
    p.srcPos = make_cudaPos((size_t)(ng - PNG) * sizeof(int32_t), ng - PNG, ng - PNG);
    p.dstPtr = make_cudaPitchedPtr(b.isfl, b.n[0] * sizeof(int32_t), b.n[0], b.n[1]);
    p.extent = make_cudaExtent((size_t)b.n[0] * sizeof(int32_t), b.n[1], b.n[2]);
    p.kind = cudaMemcpyDefault;
    CU(cudaMemcpy3DAsync(&p, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mbl_box_upload(mbl_ctx* ctx, int lev, int ibox, int which, const double* fab, int ng)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    if (ibox < 0 || ibox >= (int)L->boxes.size() || !fab || ng < 0) return fail("mbl_box_upload: bad argument");
    if (!L->local(ibox)) return fail("mbl_box_upload: box %d lives on rank %d", ibox, L->owner[ibox]);
    CU(cudaSetDevice(ctx->device));
    const PBox& b = L->set.h[ibox];
    if (copy_fab(b, which == MBL_G ? b.g[L->cur] : b.f[L->cur], NQ, const_cast<double*>(fab), ng, true, ctx->stream)) return 1;
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mbl_box_download(mbl_ctx* ctx, int lev, int ibox, int which, double* fab, int ng)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    if (ibox < 0 || ibox >= (int)L->boxes.size() || !fab || ng < 0) return fail("mbl_box_download: bad argument");
    if (!L->local(ibox)) return fail("mbl_box_download: box %d lives on rank %d", ibox, L->owner[ibox]);
    CU(cudaSetDevice(ctx->device));
    const PBox& b = L->set.h[ibox];
    if (copy_fab(b, which == MBL_G ? b.g[L->cur] : b.f[L->cur], NQ, fab, ng, false, ctx->stream)) return 1;
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mbl_box_download_macrodata(mbl_ctx* ctx, int lev, int ibox, double* fab, int ng, int derived)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    if (ibox < 0 || ibox >= (int)L->boxes.size() || !fab || ng < 0) return fail("mbl_box_download_macrodata: bad argument");
    if (!L->local(ibox)) return fail("mbl_box_download_macrodata: box %d lives on rank %d", ibox, L->owner[ibox]);
    if (!L->pool_macro) return fail("no macrodata yet: call mbl_collide with want_macrodata or mbl_f_to_macrodata");
    CU(cudaSetDevice(ctx->device));
    const PBox& b = L->set.h[ibox];
    double* src = derived ? b.macro + (size_t)MBL_NMACRO * b.sq : b.macro;
    if (copy_fab(b, src, derived ? MBL_NDERIVED : MBL_NMACRO, fab, ng, false, ctx->stream)) return 1;
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// LBM::RemakeLevel (LBM.cpp:1302-1364) after AmrCore::regrid changed the box list of level lev >= 1: the new level's
// f, g by FillPatchOps::fillpatch into NEW FABs -- K6 pre-pass on the old level, cells no OLD valid box covers by
// CellConservativeLinear from level lev - 1, everything the old level covers copied from it (periodic images
// included), BCFill.  The caller then hands over the new is_fluid (mbl_box_set_is_fluid) and calls
// mbl_fill_f_inside_eb, as RemakeLevel does.
int mbl_level_regrid(mbl_ctx* ctx, int lev, int nboxes, const int* lo, const int* hi)
{
    return mbl_level_regrid_on(ctx, lev, nboxes, lo, hi, nullptr);
}

int mbl_level_regrid_on(mbl_ctx* ctx, int lev, int nboxes, const int* lo, const int* hi, const int* owner)
{
    if (!ctx || !lo || !hi) return fail("null argument");
    if (lev < 1 || lev >= MAX_LEVELS || !ctx->plev[lev] || !ctx->plev[lev - 1])
        return fail("mbl_level_regrid: levels %d and %d must be multi-box levels", lev - 1, lev);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    PatchLevel* old = ctx->plev[lev];
    ctx->launches += launch_patch_prepass(old->set.dl, old->set.nl, old->set.max_cells, old->cur, old->G, st);
    CU(cudaStreamSynchronize(st));
    ctx->plev[lev] = nullptr;  // keep the old level alive while the new one is filled from it
    const mbl_level_geom geom = old->geom;
    int rc = mbl_level_define_boxes_on(ctx, lev, &geom, nboxes, lo, hi, owner);
    PatchLevel* L = ctx->plev[lev];
    if (!rc) {
        PatchLevel& Cl = *ctx->plev[lev - 1];
        Inter tmp;
        rc = build_interp(tmp, L->boxes, L->owner, L->G, lev, Cl, old->boxes);
        if (!rc) {
            rc = run_copy(ctx, tmp.cpatch.d, 0, Cl.set.d, Cl.cur, tmp.c2p, {{PA_F, PA_F, NQ}, {PA_G, PA_G, NQ}});
            ctx->launches += launch_patch_interp(L->set.d, L->cur, tmp.cpatch.d, tmp.d_regs, tmp.nregs, tmp.reg_max, st);
            // FillPatchSingleLevel(new, {old}): valid and ghost cells of the new boxes that lie on old valid cells
            std::vector<CopyTag> tags;
            const auto shifts = periodic_shifts(L->G, PNG);
            for (int i = 0; i < (int)L->boxes.size(); ++i)
                for (const auto& s : shifts) {
                    const int sh[3] = {s[0], s[1], s[2]};
                    for (int j = 0; j < (int)old->boxes.size(); ++j) {
                        const HBox r = isect(grow(L->boxes[i], PNG), shifted(old->boxes[j], sh));
                        if (r.ok()) tags.push_back(make_tag(i, j, r, sh));
                    }
                }
            Tags t;
            if (!rc) rc = upload_tags(tags, t, &L->owner, &old->owner, L->rank);
            if (!rc) rc = run_copy(ctx, L->set.d, L->cur, old->set.d, old->cur, t, {{PA_F, PA_F, NQ}, {PA_G, PA_G, NQ}});
            cudaStreamSynchronize(st);
            t.free_all();
        }
        cudaStreamSynchronize(st);
        tmp.free_all();
    }
    // release the old level
    PatchLevel* keep = ctx->plev[lev];
    ctx->plev[lev] = old;
    patch_clear(ctx, lev);
    ctx->plev[lev] = keep;
    if (rc) return 1;
    CU(cudaGetLastError());
    return patch_physbc(ctx, lev, 0.0);
}

// LBM::MakeNewLevelFromCoarse (LBM.cpp:1088-1144): f, g of a level that did not exist, every cell of the grown boxes
// inside the periodically grown domain interpolated from level lev-1 (FillPatchOps::fillpatch_from_coarse =
// InterpFromCoarseLevel), then the fine BCFill.  No K6 pre-pass, no zeroing of solid cells, no FillBoundary.
int mbl_level_make_from_coarse(mbl_ctx* ctx, int lev, const mbl_level_geom* geom, int nboxes, const int* lo, const int* hi)
{
    return mbl_level_make_from_coarse_on(ctx, lev, geom, nboxes, lo, hi, nullptr);
}

int mbl_level_make_from_coarse_on(mbl_ctx* ctx, int lev, const mbl_level_geom* geom, int nboxes, const int* lo, const int* hi,
                                  const int* owner)
{
    if (!ctx || !geom || !lo || !hi) return fail("null argument");
    if (lev < 1 || lev >= MAX_LEVELS || !ctx->plev[lev - 1])
        return fail("mbl_level_make_from_coarse: level %d must be a multi-box level", lev - 1);
    if (ctx->plev[lev] || ctx->lev[lev].defined)
        return fail("mbl_level_make_from_coarse: level %d exists (mbl_level_regrid re-makes it)", lev);
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (mbl_level_define_boxes_on(ctx, lev, geom, nboxes, lo, hi, owner)) return 1;
    PatchLevel* L = ctx->plev[lev];
    PatchLevel& Cl = *ctx->plev[lev - 1];
    Inter tmp;
    int rc = build_interp(tmp, L->boxes, L->owner, L->G, lev, Cl, std::vector<HBox>());
    if (!rc) {
        rc = run_copy(ctx, tmp.cpatch.d, 0, Cl.set.d, Cl.cur, tmp.c2p, {{PA_F, PA_F, NQ}, {PA_G, PA_G, NQ}});
        ctx->launches += launch_patch_interp(L->set.d, L->cur, tmp.cpatch.d, tmp.d_regs, tmp.nregs, tmp.reg_max, st);
    }
    cudaStreamSynchronize(st);
    tmp.free_all();
    if (rc) {
        patch_clear(ctx, lev);
        return 1;
    }
    CU(cudaGetLastError());
    return patch_physbc(ctx, lev, 0.0);
}

// fill_f_inside_eb (LBM.cpp:1278-1298) + the FillBoundary that follows it in RemakeLevel (LBM.cpp:1347-1348)
int mbl_fill_f_inside_eb(mbl_ctx* ctx, int lev)
{
    PatchLevel* L = plevel(ctx, lev);
    if (!L) return 1;
    CU(cudaSetDevice(ctx->device));
    ctx->launches += launch_patch_zero_solid(L->set.dl, L->set.nl, L->set.max_cells, L->cur, ctx->stream);
    if (fill_boundary_fg(ctx, *L)) return 1;
    CU(cudaGetLastError());
    return 0;
}

// LBM::average_down_to(crse_lev, ng) (LBM.cpp:1571-1584): f and g of level crse_lev + 1 onto level crse_lev
int mbl_average_down(mbl_ctx* ctx, int crse_lev, int ng)
{
    if (!ctx) return fail("null context");
    if (ng < 0 || ng > 1) return fail("mbl_average_down: ng must be 0 or 1");
    CU(cudaSetDevice(ctx->device));
    if (ensure_inter(ctx, crse_lev + 1)) return 1;
    PatchLevel& F = *ctx->plev[crse_lev + 1];
    PatchLevel& Cl = *ctx->plev[crse_lev];
    Inter& I = F.inter;
    cudaStream_t st = ctx->stream;
    const int nf = (int)F.boxes.size();
    // (the reference first copies the coarse level into the coarsened fine boxes so that a cell whose eight fine values
    // are all masked keeps it; here such a cell carries a marker that the copy back does not store)
    ctx->launches += launch_patch_avgdown(F.set.dl, F.cur, I.avg[ng].dl, F.set.nl, I.avg[ng].max_cells, ng, st);
    if (run_copy(ctx, Cl.set.d, Cl.cur, I.avg[ng].d, 0, I.a2c[ng], {{PA_F, PA_F, NQ}, {PA_G, PA_G, NQ}}, true)) return 1;
    CU(cudaGetLastError());
    return 0;
}

}  // extern "C"
