// One-time symbolic pass (host): block structure of H, edge -> slot map, sliced jagged storage
// and the structure of the aggregation-AMG hierarchy.  The reference never materialises this
// structure explicitly -- it is implied by the 36 (pose-pose) / 25 (pose-landmark) `put` calls
// per edge of update_linear_system / set_matrix (pose_graph_optimization.rs:165-206), re-derived
// inside russell_sparse's COO->CSC conversion on every iteration.  Here it is computed once.
#include "pgo_internal.h"

#include <algorithm>
#include <cstring>
#include <numeric>

namespace pgo {

static const int KIND_DIM[3] = {3, 2, 6};   // lut stride, g2o.rs:61,68,77
static const int KIND_NVAL[3] = {3, 2, 7};

// ------------------------------------------------------------------------------------------------
// row ordering: inside windows of `window` rows (multiple of 32) sort by decreasing degree (stable)
static void order_rows(const std::vector<int32_t> &deg, int64_t n, int window, std::vector<int32_t> &perm) {
    perm.resize(n);
    std::iota(perm.begin(), perm.end(), 0);
    if (window < 32) window = 32;
    window = (window / 32) * 32;
    for (int64_t w = 0; w < n; w += window) {
        int64_t e = std::min<int64_t>(n, w + window);
        std::stable_sort(perm.begin() + w, perm.begin() + e, [&](int32_t a, int32_t b) { return deg[a] > deg[b]; });
    }
}

// sliced jagged storage from per-row entry counts (rows already in storage order)
static void build_jds(HostLevel &L, int64_t n, const std::vector<int64_t> &ptr) {
    L.n = n;
    L.n_pad = (n + 31) / 32 * 32;
    if (L.n_pad == 0) L.n_pad = 32;
    L.n_slices = L.n_pad / 32;
    L.deg.assign(L.n_pad, 0);
    for (int64_t r = 0; r < n; r++) L.deg[r] = (int32_t)(ptr[r + 1] - ptr[r]);
    L.slice_ptr.assign(L.n_slices + 1, 0);
    L.adj_ptr = ptr;
    int64_t nent = ptr[n];
    L.adj_slot.assign(nent, 0);
    L.adj_cnt.assign(nent, 0);
    int64_t base = 0;
    for (int64_t s = 0; s < L.n_slices; s++) {
        L.slice_ptr[s] = base;
        const int32_t *d = &L.deg[32 * s];
        int maxdeg = d[0];
        int64_t off = 0;
        for (int k = 0; k < maxdeg; k++) {
            int cnt = 0;
            while (cnt < 32 && d[cnt] > k) cnt++;         // rows are sorted by decreasing degree
            for (int l = 0; l < cnt; l++) {
                int64_t row = 32 * s + l;
                L.adj_slot[ptr[row] + k] = base + off + l;
                L.adj_cnt[ptr[row] + k] = cnt;
            }
            off += cnt;
        }
        base += (off + 1) & ~int64_t(1);                   // even slot count: blobs stay 16-byte aligned
    }
    L.slice_ptr[L.n_slices] = base;
    L.n_slots = base;
}

// ------------------------------------------------------------------------------------------------
bool build_canonical(Symbolic &S) {
    if (!S.brow_ptr.empty()) return true;
    const int64_t n = S.n, ne = S.n_edges;
    std::vector<int64_t> cnt(n + 1, 0);
    for (int64_t k = 0; k < ne; k++) { cnt[S.efrom[k] + 1]++; cnt[S.eto[k] + 1]++; }
    for (int64_t v = 0; v < n; v++) cnt[v + 1] += cnt[v] + 1;      // +1: the diagonal block
    std::vector<int32_t> nb(cnt[n]);
    std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
    for (int64_t v = 0; v < n; v++) nb[pos[v]++] = (int32_t)v;
    for (int64_t k = 0; k < ne; k++) { nb[pos[S.efrom[k]]++] = S.eto[k]; nb[pos[S.eto[k]]++] = S.efrom[k]; }
    S.brow_ptr.assign(n + 1, 0);
    S.bcol.clear();
    S.bcol.reserve(cnt[n]);
    for (int64_t v = 0; v < n; v++) {
        auto b = nb.begin() + cnt[v], e = nb.begin() + cnt[v + 1];
        std::sort(b, e);
        e = std::unique(b, e);
        S.bcol.insert(S.bcol.end(), b, e);
        S.brow_ptr[v + 1] = (int64_t)S.bcol.size();
    }
    auto find = [&](int32_t r, int32_t c) -> int64_t {
        auto b = S.bcol.begin() + S.brow_ptr[r], e = S.bcol.begin() + S.brow_ptr[r + 1];
        return std::lower_bound(b, e, c) - S.bcol.begin();
    };
    S.edge_slots.resize(4 * ne);
    for (int64_t k = 0; k < ne; k++) {
        int32_t i = S.efrom[k], j = S.eto[k];
        S.edge_slots[4 * k + 0] = find(i, i); S.edge_slots[4 * k + 1] = find(i, j);
        S.edge_slots[4 * k + 2] = find(j, i); S.edge_slots[4 * k + 3] = find(j, j);
    }
    return true;
}

bool build_csc_pattern(Symbolic &S) {
    if (!S.csc_ptr.empty()) return true;
    build_canonical(S);
    // H is structurally symmetric: block column v has the block rows of block row v
    int64_t nnz = 0;
    for (int64_t v = 0; v < S.n; v++) {
        int64_t rows = 0;
        for (int64_t p = S.brow_ptr[v]; p < S.brow_ptr[v + 1]; p++) rows += KIND_DIM[S.vkind[S.bcol[p]]];
        nnz += rows * KIND_DIM[S.vkind[v]];
    }
    if (nnz > 0x7fffffffll || S.len > 0x7ffffffell) { S.error = "pattern exceeds 32-bit CSC indices"; return false; }
    S.csc_ptr.assign(S.len + 1, 0);
    S.csc_row.resize(nnz);
    int64_t o = 0;
    for (int64_t v = 0; v < S.n; v++)
        for (int c = 0; c < KIND_DIM[S.vkind[v]]; c++) {
            for (int64_t p = S.brow_ptr[v]; p < S.brow_ptr[v + 1]; p++) {
                int32_t u = S.bcol[p];
                for (int r = 0; r < KIND_DIM[S.vkind[u]]; r++) S.csc_row[o++] = (int32_t)(S.voffset[u] + r);
            }
            S.csc_ptr[S.voffset[v] + c + 1] = (int32_t)o;
        }
    return true;
}

// ------------------------------------------------------------------------------------------------
// unique-neighbour adjacency of a level (storage order), from its stored entries
static void unique_adjacency(const HostLevel &L, std::vector<int64_t> &ptr, std::vector<int32_t> &nbr) {
    ptr.assign(L.n + 1, 0);
    nbr.clear();
    nbr.reserve(L.adj_nbr.size());
    for (int64_t r = 0; r < L.n; r++) {
        size_t start = nbr.size();
        for (int64_t p = L.adj_ptr[r]; p < L.adj_ptr[r + 1]; p++) nbr.push_back(L.adj_nbr[p]);
        std::sort(nbr.begin() + start, nbr.end());
        nbr.erase(std::unique(nbr.begin() + start, nbr.end()), nbr.end());
        ptr[r + 1] = (int64_t)nbr.size();
    }
}

// root + neighbours aggregation (Vanek-style, three passes) on a unique-neighbour adjacency
static int32_t aggregate_graph(int64_t n, const std::vector<int64_t> &ptr, const std::vector<int32_t> &nbr,
                               int max_size, std::vector<int32_t> &agg) {
    agg.assign(n, -1);
    int32_t nc = 0;
    for (int64_t i = 0; i < n; i++) {
        if (agg[i] >= 0) continue;
        bool free_nb = true;
        for (int64_t p = ptr[i]; p < ptr[i + 1] && free_nb; p++) free_nb = agg[nbr[p]] < 0;
        if (!free_nb) continue;
        agg[i] = nc;
        int sz = 1;
        for (int64_t p = ptr[i]; p < ptr[i + 1] && sz < max_size; p++) { agg[nbr[p]] = nc; sz++; }
        nc++;
    }
    std::vector<int32_t> snap(agg);
    std::vector<int32_t> cand;
    for (int64_t i = 0; i < n; i++) {
        if (agg[i] >= 0) continue;
        cand.clear();
        for (int64_t p = ptr[i]; p < ptr[i + 1]; p++) if (snap[nbr[p]] >= 0) cand.push_back(snap[nbr[p]]);
        if (cand.empty()) continue;
        std::sort(cand.begin(), cand.end());
        int32_t best = cand[0]; int bc = 0, run = 0;
        for (size_t q = 0; q < cand.size(); q++) {
            run = (q > 0 && cand[q] == cand[q - 1]) ? run + 1 : 1;
            if (run > bc) { bc = run; best = cand[q]; }
        }
        agg[i] = best;
    }
    for (int64_t i = 0; i < n; i++) {
        if (agg[i] >= 0) continue;
        agg[i] = nc;
        int sz = 1;
        for (int64_t p = ptr[i]; p < ptr[i + 1] && sz < max_size; p++) if (agg[nbr[p]] < 0) { agg[nbr[p]] = nc; sz++; }
        nc++;
    }
    return nc;
}

// Level 0: runs of up to `run` consecutive (lut-order) vertices that are chained by edges -- a short
// odometry segment moves almost rigidly, which is exactly what the coarse basis can represent.
// Vertices left alone (landmarks, chain breaks) join the neighbouring aggregate they touch most.
static int32_t aggregate_chain(const Symbolic &S, const HostLevel &L, const std::vector<int64_t> &ptr,
                               const std::vector<int32_t> &nbr, int run, int max_size, std::vector<int32_t> &agg) {
    const int64_t n = L.n;
    agg.assign(n, -1);
    std::vector<int32_t> size;
    int32_t nc = 0; int cur = 0;
    auto adjacent = [&](int32_t a, int32_t b) {            // internal rows
        return std::binary_search(nbr.begin() + ptr[a], nbr.begin() + ptr[a + 1], b);
    };
    for (int64_t v = 0; v < n; v++) {                       // lut order
        int32_t r = S.iperm[v];
        bool join = v > 0 && cur > 0 && cur < run && adjacent(r, S.iperm[v - 1]) && agg[S.iperm[v - 1]] == nc - 1;
        if (join) { agg[r] = nc - 1; cur++; size[nc - 1]++; }
        else { agg[r] = nc++; cur = 1; size.push_back(1); }
    }
    // merge singletons into the most-connected neighbouring aggregate with room
    std::vector<int32_t> cand;
    for (int64_t r = 0; r < n; r++) {
        if (size[agg[r]] != 1) continue;
        cand.clear();
        for (int64_t p = ptr[r]; p < ptr[r + 1]; p++) { int32_t a = agg[nbr[p]]; if (a != agg[r] && size[a] < max_size && size[a] > 1) cand.push_back(a); }
        if (cand.empty()) continue;
        std::sort(cand.begin(), cand.end());
        int32_t best = cand[0]; int bc = 0, rn = 0;
        for (size_t q = 0; q < cand.size(); q++) { rn = (q > 0 && cand[q] == cand[q - 1]) ? rn + 1 : 1; if (rn > bc) { bc = rn; best = cand[q]; } }
        size[agg[r]] = 0; agg[r] = best; size[best]++;
    }
    // compact ids
    std::vector<int32_t> remap(nc, -1);
    int32_t m = 0;
    for (int64_t v = 0; v < n; v++) { int32_t &a = agg[S.iperm[v]]; if (remap[a] < 0) remap[a] = m++; a = remap[a]; }
    return m;
}

// Given a fine level and an aggregation (ids 0..nc-1 in creation order), build the coarse level
// (storage-ordered), relabel agg to coarse storage rows, and fill the Galerkin targets.
static void build_coarse_level(HostLevel &F, HostLevel &C, std::vector<int32_t> &agg, int32_t nc, int window, int DD) {
    const int64_t n = F.n;
    // coarse unique adjacency in creation ids
    std::vector<int64_t> cnt(nc + 1, 0);
    for (int64_t r = 0; r < n; r++) cnt[agg[r] + 1] += F.adj_ptr[r + 1] - F.adj_ptr[r];
    for (int32_t a = 0; a < nc; a++) cnt[a + 1] += cnt[a];
    std::vector<int32_t> tmp(cnt[nc]);
    std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
    for (int64_t r = 0; r < n; r++)
        for (int64_t p = F.adj_ptr[r]; p < F.adj_ptr[r + 1]; p++) tmp[pos[agg[r]]++] = agg[F.adj_nbr[p]];
    std::vector<int64_t> cptr(nc + 1, 0);
    std::vector<int32_t> cnbr;
    cnbr.reserve(tmp.size() / 2);
    std::vector<int32_t> cdeg(nc);
    for (int32_t a = 0; a < nc; a++) {
        auto b = tmp.begin() + cnt[a], e = tmp.begin() + cnt[a + 1];
        std::sort(b, e);
        e = std::unique(b, e);
        for (auto it = b; it != e; ++it) if (*it != a) cnbr.push_back(*it);
        cptr[a + 1] = (int64_t)cnbr.size();
        cdeg[a] = (int32_t)(cptr[a + 1] - cptr[a]);
    }
    // storage order of the coarse level
    std::vector<int32_t> cperm, ciperm(nc);
    order_rows(cdeg, nc, window, cperm);
    for (int32_t s = 0; s < nc; s++) ciperm[cperm[s]] = s;
    std::vector<int64_t> sptr(nc + 1, 0);
    for (int32_t s = 0; s < nc; s++) sptr[s + 1] = sptr[s] + cdeg[cperm[s]];
    build_jds(C, nc, sptr);
    C.adj_nbr.resize(sptr[nc]);
    C.col.assign(C.n_slots, 0);
    for (int32_t s = 0; s < nc; s++) {
        int32_t a = cperm[s];
        int64_t o = sptr[s];
        for (int64_t p = cptr[a]; p < cptr[a + 1]; p++) C.adj_nbr[o++] = ciperm[cnbr[p]];
        std::sort(C.adj_nbr.begin() + sptr[s], C.adj_nbr.begin() + sptr[s + 1]);
        for (int64_t q = sptr[s]; q < sptr[s + 1]; q++) C.col[C.adj_slot[q]] = (uint32_t)C.adj_nbr[q];
    }
    // relabel agg to storage rows; padding rows -> -1
    F.agg.assign(F.n_pad, -1);
    for (int64_t r = 0; r < n; r++) F.agg[r] = ciperm[agg[r]];
    // members
    C.mem_ptr.assign(nc + 1, 0);
    for (int64_t r = 0; r < n; r++) C.mem_ptr[F.agg[r] + 1]++;
    for (int32_t s = 0; s < nc; s++) C.mem_ptr[s + 1] += C.mem_ptr[s];
    C.mem_idx.resize(n);
    std::vector<int64_t> mp(C.mem_ptr.begin(), C.mem_ptr.end() - 1);
    for (int64_t r = 0; r < n; r++) C.mem_idx[mp[F.agg[r]]++] = (int32_t)r;
    // Galerkin targets of every fine slot
    F.ctgt.assign(F.n_slots, CTGT_DIAG);
    F.cstr.assign(F.n_slots, 0);
    for (int64_t r = 0; r < n; r++) {
        int32_t I = F.agg[r];
        for (int64_t p = F.adj_ptr[r]; p < F.adj_ptr[r + 1]; p++) {
            int32_t J = F.agg[F.adj_nbr[p]];
            int64_t slot = F.adj_slot[p];
            if (I == J) { F.ctgt[slot] = CTGT_DIAG | I; F.cstr[slot] = (int32_t)C.n_pad; continue; }
            auto b = C.adj_nbr.begin() + C.adj_ptr[I], e = C.adj_nbr.begin() + C.adj_ptr[I + 1];
            int64_t q = std::lower_bound(b, e, J) - C.adj_nbr.begin();
            int64_t cs = C.adj_slot[q];
            int lane = I & 31;
            F.ctgt[slot] = (cs - lane) * DD + lane;
            F.cstr[slot] = C.adj_cnt[q];
        }
    }
}

// ------------------------------------------------------------------------------------------------
bool build_symbolic(Symbolic &S, const SymbolicOptions &opt,
                    int64_t nv, const uint32_t *vid, const uint8_t *vkind,
                    int64_t ne, const uint8_t *ekind, const uint32_t *efrom, const uint32_t *eto) {
    if (nv <= 0 || nv > 0x3fffffffll || ne < 0 || ne > 0x7fffffffll) { S.error = "vertex/edge count out of range"; return false; }
    S.n = nv; S.n_edges = ne;
    S.vkind.assign(vkind, vkind + nv);
    S.voffset.resize(nv); S.vvalofs.resize(nv);
    int has2 = 0, has3 = 0;
    int64_t off = 0, vo = 0;
    for (int64_t v = 0; v < nv; v++) {
        if (vkind[v] > 2) { S.error = "unknown vertex kind"; return false; }
        (vkind[v] == 2 ? has3 : has2) = 1;
        S.voffset[v] = off; S.vvalofs[v] = vo;
        off += KIND_DIM[vkind[v]]; vo += KIND_NVAL[vkind[v]];
    }
    if (has2 && has3) { S.error = "SE2/XY and SE3 vertices cannot be mixed in one graph"; return false; }
    S.D = has3 ? 6 : 3;
    S.len = off; S.n_values = vo;
    // id -> index (the reference's lut, g2o.rs:60)
    std::vector<std::pair<uint32_t, int32_t>> ids(nv);
    for (int64_t v = 0; v < nv; v++) ids[v] = {vid[v], (int32_t)v};
    std::sort(ids.begin(), ids.end());
    for (int64_t v = 1; v < nv; v++) if (ids[v].first == ids[v - 1].first) { S.error = "duplicate vertex id " + std::to_string(ids[v].first); return false; }
    auto lookup = [&](uint32_t id) -> int32_t {
        auto it = std::lower_bound(ids.begin(), ids.end(), std::make_pair(id, (int32_t)-1));
        return (it != ids.end() && it->first == id) ? it->second : -1;
    };
    S.efrom.resize(ne); S.eto.resize(ne); S.ekind.assign(ekind, ekind + ne);
    S.anchor = -1;
    for (int64_t k = 0; k < ne; k++) {
        int32_t i = lookup(efrom[k]), j = lookup(eto[k]);
        if (i < 0 || j < 0) { S.error = "edge " + std::to_string(k) + " references an unknown vertex id"; return false; }
        if (i == j) { S.error = "edge " + std::to_string(k) + " is a self loop"; return false; }
        bool ok = (ekind[k] == 0 && vkind[i] == 0 && vkind[j] == 0) || (ekind[k] == 1 && vkind[i] == 0 && vkind[j] == 1) ||
                  (ekind[k] == 2 && vkind[i] == 2 && vkind[j] == 2);
        if (!ok) { S.error = "edge " + std::to_string(k) + ": vertex kinds do not match the edge kind"; return false; }
        S.efrom[k] = i; S.eto[k] = j;
        if (S.anchor < 0 && ekind[k] != 1) S.anchor = i;   // first pose-pose edge's `from` (:330-336)
    }
    // degrees, storage order
    std::vector<int32_t> deg(nv, 0);
    for (int64_t k = 0; k < ne; k++) { deg[S.efrom[k]]++; deg[S.eto[k]]++; }
    order_rows(deg, nv, opt.sort_window, S.perm);
    S.iperm.resize(nv);
    for (int64_t r = 0; r < nv; r++) S.iperm[S.perm[r]] = (int32_t)r;
    // half edges per storage row, sorted by neighbour row
    std::vector<int64_t> ptr(nv + 1, 0);
    for (int64_t r = 0; r < nv; r++) ptr[r + 1] = ptr[r] + deg[S.perm[r]];
    struct HE { int32_t nbr, edge; uint32_t flags; };
    std::vector<HE> he(ptr[nv]);
    {
        std::vector<int64_t> pos(ptr.begin(), ptr.end() - 1);
        for (int64_t k = 0; k < ne; k++) {
            int32_t ri = S.iperm[S.efrom[k]], rj = S.iperm[S.eto[k]];
            uint32_t xy = ekind[k] == 1 ? COL_EDGE_XY : 0u;
            he[pos[ri]++] = {rj, (int32_t)k, xy};
            he[pos[rj]++] = {ri, (int32_t)k, xy | COL_ROLE_TO};
        }
        for (int64_t r = 0; r < nv; r++)
            std::sort(he.begin() + ptr[r], he.begin() + ptr[r + 1], [](const HE &a, const HE &b) { return a.nbr != b.nbr ? a.nbr < b.nbr : a.edge < b.edge; });
    }
    S.levels.clear();
    S.levels.emplace_back();
    HostLevel &L0 = S.levels[0];
    build_jds(L0, nv, ptr);
    L0.col.assign(L0.n_slots, 0);
    L0.adj_nbr.resize(ptr[nv]);
    S.slot_edge.assign(L0.n_slots, -1);
    for (int64_t q = 0; q < ptr[nv]; q++) {
        L0.adj_nbr[q] = he[q].nbr;
        L0.col[L0.adj_slot[q]] = (uint32_t)he[q].nbr | he[q].flags;
        S.slot_edge[L0.adj_slot[q]] = he[q].edge;
    }
    he.clear(); he.shrink_to_fit();
    if (!opt.build_amg) return true;

    // ---- aggregation hierarchy
    const int DD = S.D * S.D;
    for (int lvl = 0; lvl + 1 < opt.amg_max_levels; lvl++) {
        HostLevel &F = S.levels[lvl];
        if (F.n <= opt.coarsest_max) break;
        std::vector<int64_t> uptr; std::vector<int32_t> unbr;
        unique_adjacency(F, uptr, unbr);
        std::vector<int32_t> agg;
        int32_t nc = 0;
        if (lvl == 0) {
            nc = aggregate_chain(S, F, uptr, unbr, 4, 8, agg);
            if (nc > 0.6 * F.n) nc = aggregate_graph(F.n, uptr, unbr, 16, agg);
        } else {
            nc = aggregate_graph(F.n, uptr, unbr, 16, agg);
        }
        if (nc >= F.n || nc > 0.9 * F.n) break;                 // coarsening stalled
        S.levels.emplace_back();
        build_coarse_level(S.levels[lvl], S.levels[lvl + 1], agg, nc, opt.sort_window, DD);
    }
    return true;
}

} // namespace pgo
