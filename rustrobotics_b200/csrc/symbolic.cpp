// One-time symbolic pass (host): block structure of H, edge -> slot map, sliced jagged storage of the
// Gauss-Newton system, contiguous vertex-range partition across ranks, and the structure of the
// aggregation-AMG hierarchy.  The reference never materialises this structure explicitly -- it is implied
// by the 36 (pose-pose) / 25 (pose-landmark) `put` calls per edge of update_linear_system / set_matrix
// (pose_graph_optimization.rs:165-206), re-derived inside russell_sparse's COO->CSC conversion on every
// iteration.  Here it is computed once.
#include "pgo_internal.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <numeric>
#include <thread>

namespace pgo {

static const int KIND_DIM[3] = {3, 2, 6};   // lut stride, g2o.rs:61,68,77
static const int KIND_NVAL[3] = {3, 2, 7};

static inline int64_t pad32(int64_t x) { return (x + 31) / 32 * 32; }

// partitions laid out back to back, each padded to a multiple of 32 rows
static void layout_partitions(HostLevel &L, const std::vector<int64_t> &counts) {
    const int world = (int)counts.size();
    L.part_real = counts;
    L.part_off.assign(world + 1, 0);
    L.n = 0;
    for (int k = 0; k < world; k++) { L.part_off[k + 1] = L.part_off[k] + pad32(counts[k]); L.n += counts[k]; }
    if (L.part_off[world] == 0) L.part_off[world] = 32;
    L.n_pad = L.part_off[world];
    L.real.assign(L.n_pad, 0);
    for (int k = 0; k < world; k++) std::fill(L.real.begin() + L.part_off[k], L.real.begin() + L.part_off[k] + counts[k], 1);
}

// sliced jagged storage from L.adj_ptr (rows of a slice already ordered by decreasing entry count)
static void build_jds(HostLevel &L) {
    L.jds = true;
    L.n_slices = L.n_pad / 32;
    L.deg.assign(L.n_pad, 0);
    for (int64_t r = 0; r < L.n_pad; r++) L.deg[r] = (int32_t)(L.adj_ptr[r + 1] - L.adj_ptr[r]);
    L.slice_ptr.assign(L.n_slices + 1, 0);
    const int64_t nent = L.adj_ptr[L.n_pad];
    L.adj_slot.assign(nent, 0);
    L.adj_cnt.assign(nent, 0);
    // slots per slice (even: blobs stay 16-byte aligned) -> prefix sum -> fill
    parallel_for(L.n_slices, 1024, [&](int64_t s0, int64_t s1) {
        for (int64_t s = s0; s < s1; s++) {
            int64_t tot = 0;
            for (int l = 0; l < 32; l++) tot += L.deg[32 * s + l];
            L.slice_ptr[s + 1] = (tot + 1) & ~int64_t(1);
        }
    });
    for (int64_t s = 0; s < L.n_slices; s++) L.slice_ptr[s + 1] += L.slice_ptr[s];
    parallel_for(L.n_slices, 1024, [&](int64_t s0, int64_t s1) {
        for (int64_t s = s0; s < s1; s++) {
            const int64_t base = L.slice_ptr[s];
            const int32_t *d = &L.deg[32 * s];
            const int maxdeg = d[0];
            int64_t off = 0;
            for (int k = 0; k < maxdeg; k++) {
                int cnt = 0;
                while (cnt < 32 && d[cnt] > k) cnt++;
                for (int l = 0; l < cnt; l++) {
                    const int64_t row = 32 * s + l;
                    L.adj_slot[L.adj_ptr[row] + k] = base + off + l;
                    L.adj_cnt[L.adj_ptr[row] + k] = cnt;
                }
                off += cnt;
            }
        }
    });
    const int64_t base = L.slice_ptr[L.n_slices];
    L.slice_ptr[L.n_slices] = base;
    L.n_slots = base;
    L.part_slot.assign(L.part_off.size(), 0);
    for (size_t k = 0; k < L.part_off.size(); k++) L.part_slot[k] = L.slice_ptr[L.part_off[k] / 32];
}

static void build_csr(HostLevel &L) {
    L.jds = false;
    const int64_t nent = L.adj_ptr[L.n_pad];
    L.adj_slot.resize(nent);
    std::iota(L.adj_slot.begin(), L.adj_slot.end(), int64_t(0));
    L.adj_cnt.assign(nent, 1);
    L.n_slots = nent;
    L.part_slot.assign(L.part_off.size(), 0);
    for (size_t k = 0; k < L.part_off.size(); k++) L.part_slot[k] = L.adj_ptr[L.part_off[k]];
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
// Root + neighbours aggregation (Vanek-style, three passes) of the rows of ONE partition of a level, restricted to
// edges inside the partition (aggregates never straddle ranks, so restriction, prolongation and the Galerkin
// product need no communication).  agg[row - r0] = aggregate id in creation order.
//
// Graphs with landmarks (level 0 only; landmark[row - r0] != 0 marks a VERTEX_XY row): a landmark is a star -- observed from
// dozens of poses, with no edge to another landmark -- and a star centre as aggregation root collects poses from all over the
// trajectory (dlr.g2o: 250 PCG iterations).  So only the POSES are aggregated, over the pose-pose edges; every landmark then
// joins the aggregate that holds most of its observing poses, and when the pose graph is chain-like (odometry only: aggregates
// of ~3 poses) neighbouring aggregates are paired once more.  Same hierarchy, 6x fewer iterations on dlr (tools/research/).
static int32_t aggregate_partition(const HostLevel &F, int k, int max_size, std::vector<int32_t> &agg, const uint8_t *landmark = nullptr) {
    const int64_t r0 = F.part_off[k], n = F.part_real[k];
    agg.assign(n, -1);
    if (n == 0) return 0;
    bool mixed = false;
    if (landmark) for (int64_t i = 0; i < n && !mixed; i++) mixed = landmark[i] != 0;
    if (!mixed) landmark = nullptr;
    // sorted unique in-partition neighbours of every row: row i's list is nbr[ptr[i] .. end[i])
    const int64_t a0 = F.adj_ptr[r0];
    std::vector<int64_t> ptr(n), end(n);
    std::vector<int32_t> nbr(std::max<int64_t>(F.adj_ptr[r0 + n] - a0, 1));
    parallel_for(n, 4096, [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; i++) {
            int32_t *b = nbr.data() + (F.adj_ptr[r0 + i] - a0), *e = b;
            if (!(landmark && landmark[i]))
                for (int64_t p = F.adj_ptr[r0 + i]; p < F.adj_ptr[r0 + i + 1]; p++) {
                    const int64_t j = F.adj_nbr[p] - r0;
                    if (j >= 0 && j < n && j != i && !(landmark && landmark[j])) *e++ = (int32_t)j;
                }
            std::sort(b, e);
            e = std::unique(b, e);
            ptr[i] = b - nbr.data(); end[i] = e - nbr.data();
        }
    });
    if (landmark) for (int64_t i = 0; i < n; i++) if (landmark[i]) agg[i] = -2;      // not a candidate in the three passes below
    int32_t nc = 0;
    for (int64_t i = 0; i < n; i++) {
        if (agg[i] != -1) continue;
        bool free_nb = true;
        for (int64_t p = ptr[i]; p < end[i] && free_nb; p++) free_nb = agg[nbr[p]] < 0;
        if (!free_nb) continue;
        agg[i] = nc;
        int sz = 1;
        for (int64_t p = ptr[i]; p < end[i] && sz < max_size; p++) { agg[nbr[p]] = nc; sz++; }
        nc++;
    }
    // a free row joins the root aggregate most of its neighbours are in (the state after the root pass: rows are independent)
    std::vector<int32_t> snap(agg), cand;
    parallel_for(n, 8192, [&](int64_t i0, int64_t i1) {
        std::vector<int32_t> cand;
        for (int64_t i = i0; i < i1; i++) {
            if (snap[i] != -1) continue;
            cand.clear();
            for (int64_t p = ptr[i]; p < end[i]; p++) if (snap[nbr[p]] >= 0) cand.push_back(snap[nbr[p]]);
            if (cand.empty()) continue;
            std::sort(cand.begin(), cand.end());
            int32_t best = cand[0]; int bc = 0, run = 0;
            for (size_t q = 0; q < cand.size(); q++) {
                run = (q > 0 && cand[q] == cand[q - 1]) ? run + 1 : 1;
                if (run > bc) { bc = run; best = cand[q]; }
            }
            agg[i] = best;
        }
    });
    for (int64_t i = 0; i < n; i++) {
        if (agg[i] >= 0 || agg[i] == -2) continue;
        agg[i] = nc;
        int sz = 1;
        for (int64_t p = ptr[i]; p < end[i] && sz < max_size; p++) if (agg[nbr[p]] == -1) { agg[nbr[p]] = nc; sz++; }
        nc++;
    }
    if (!landmark) return nc;
    // ---- landmark graphs: pair up the aggregates of a chain-like pose graph, then attach the landmarks
    int64_t n_pose = 0;
    for (int64_t i = 0; i < n; i++) n_pose += !landmark[i];
    if (nc > 1 && (int64_t)nc * 4 > n_pose) {
        std::vector<std::vector<int32_t>> cadj(nc);
        for (int64_t i = 0; i < n; i++) {
            if (landmark[i]) continue;
            for (int64_t p = ptr[i]; p < end[i]; p++) if (agg[nbr[p]] != agg[i]) cadj[agg[i]].push_back(agg[nbr[p]]);
        }
        std::vector<int32_t> pair(nc, -1);
        int32_t np = 0;
        for (int32_t a = 0; a < nc; a++) {
            if (pair[a] >= 0) continue;
            pair[a] = np;
            std::sort(cadj[a].begin(), cadj[a].end());
            for (int32_t c : cadj[a]) if (pair[c] < 0) { pair[c] = np; break; }
            np++;
        }
        for (int64_t i = 0; i < n; i++) if (agg[i] >= 0) agg[i] = pair[agg[i]];
        nc = np;
    }
    for (int64_t i = 0; i < n; i++) {
        if (!landmark[i]) continue;
        cand.clear();
        for (int64_t p = F.adj_ptr[r0 + i]; p < F.adj_ptr[r0 + i + 1]; p++) {
            const int64_t j = F.adj_nbr[p] - r0;
            if (j >= 0 && j < n && agg[j] >= 0 && !landmark[j]) cand.push_back(agg[j]);
        }
        if (cand.empty()) { agg[i] = nc++; continue; }        // all its poses live on other ranks (or none at all)
        std::sort(cand.begin(), cand.end());
        int32_t best = cand[0]; int bc = 0, run = 0;
        for (size_t q = 0; q < cand.size(); q++) {
            run = (q > 0 && cand[q] == cand[q - 1]) ? run + 1 : 1;
            if (run > bc) { bc = run; best = cand[q]; }
        }
        agg[i] = best;
    }
    return nc;
}

// Build the coarse level (block CSR, global padded numbering) from per-partition aggregate ids; fills F.agg, F.ctgt.
// merge: F is sharded over pnc.size() partitions but C becomes one replicated partition (rows of rank k's aggregates at src_off[k])
static void build_coarse_level(HostLevel &F, HostLevel &C, const std::vector<std::vector<int32_t>> &pagg,
                               const std::vector<int64_t> &pnc, bool jds, bool merge, int DD) {
    double t_last = now_s();
    const int world = (int)pnc.size();
    std::vector<int64_t> cbase(world, 0);       // coarse row of partition k's aggregate 0
    if (merge) {
        int64_t tot = 0;
        C.src_off.assign(world + 1, 0);
        for (int k = 0; k < world; k++) { cbase[k] = tot; tot += pnc[k]; C.src_off[k + 1] = tot; }
        layout_partitions(C, std::vector<int64_t>{tot});
    } else {
        layout_partitions(C, pnc);
        for (int k = 0; k < world; k++) cbase[k] = C.part_off[k];
    }
    C.repl = merge || F.repl;
    F.agg.assign(F.n_pad, -1);
    for (int k = 0; k < world; k++)
        for (int64_t i = 0; i < F.part_real[k]; i++) F.agg[F.part_off[k] + i] = (int32_t)(cbase[k] + pagg[k][i]);
    // members of every coarse row
    C.mem_ptr.assign(C.n_pad + 1, 0);
    for (int64_t r = 0; r < F.n_pad; r++) if (F.agg[r] >= 0) C.mem_ptr[F.agg[r] + 1]++;
    for (int64_t I = 0; I < C.n_pad; I++) C.mem_ptr[I + 1] += C.mem_ptr[I];
    C.mem_idx.resize(C.mem_ptr[C.n_pad]);
    {
        std::vector<int64_t> mp(C.mem_ptr.begin(), C.mem_ptr.end() - 1);
        for (int64_t r = 0; r < F.n_pad; r++) if (F.agg[r] >= 0) C.mem_idx[mp[F.agg[r]]++] = (int32_t)r;
    }
    TICK("  bcl: members");
    // coarse adjacency: I ~ J when any member of I has a stored block towards a member of J
    C.adj_ptr.assign(C.n_pad + 1, 0);
    C.adj_nbr.clear();
    {
        std::vector<std::vector<int32_t>> lists(C.n_pad);
        parallel_for(C.n_pad, 256, [&](int64_t I0, int64_t I1) {
            std::vector<int32_t> tmp;
            for (int64_t I = I0; I < I1; I++) {
                tmp.clear();
                for (int64_t m = C.mem_ptr[I]; m < C.mem_ptr[I + 1]; m++) {
                    const int64_t r = C.mem_idx[m];
                    for (int64_t p = F.adj_ptr[r]; p < F.adj_ptr[r + 1]; p++) { const int32_t J = F.agg[F.adj_nbr[p]]; if (J != I) tmp.push_back(J); }
                }
                std::sort(tmp.begin(), tmp.end());
                tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
                lists[I] = tmp;
            }
        });
        for (int64_t I = 0; I < C.n_pad; I++) C.adj_ptr[I + 1] = C.adj_ptr[I] + (int64_t)lists[I].size();
        C.adj_nbr.resize(C.adj_ptr[C.n_pad]);
        parallel_for(C.n_pad, 256, [&](int64_t I0, int64_t I1) {
            for (int64_t I = I0; I < I1; I++) std::copy(lists[I].begin(), lists[I].end(), C.adj_nbr.begin() + C.adj_ptr[I]);
        });
    }
    TICK("  bcl: adjacency");
    if (jds) build_jds(C); else build_csr(C);
    TICK("  bcl: storage");
    // Galerkin targets of every fine block, local to the owning partition: element offset of component 0 of the coarse
    // block inside the partition's val array + the stride between components (CSR: 9 s, 1 ; JDS: component-major).
    // Contributor lists of the deterministic (atomics-free) Galerkin product (block CSR coarse levels): every fine block is first
    // projected into a staging buffer in storage order (coalesced), then ONE group of lanes per coarse block sums its contributors
    // in ascending stage order (stored blocks by slot, then the diagonal blocks by row).
    // Both come out of walks over the COARSE rows: every contributor of the blocks of coarse row I (its diagonal block and its
    // stored blocks) is a block of one of I's members, so a coarse row is counted, filled and sorted by one thread, without atomics.
    F.ctgt.assign(F.n_slots, 0);
    F.cstr.assign(F.n_slots, 1);
    F.gal_ptr.assign(world, {});
    F.gal_src.assign(world, {});
    for (int k = 0; k < world; k++) {
        const int kc = merge ? 0 : k;
        const int64_t c0 = C.part_off[kc], crows = C.part_off[kc + 1] - c0, cs0 = C.part_slot[kc], cslots = C.part_slot[kc + 1] - cs0;
        const int64_t I0 = merge ? C.src_off[k] : c0, I1 = merge ? C.src_off[k + 1] : c0 + C.part_real[kc];   // coarse rows built from partition k
        const int64_t fr0 = F.part_off[k], fs0 = F.part_slot[k], fslots = F.part_slot[k + 1] - fs0;
        const int64_t nblk = crows + cslots;
        std::vector<int32_t> &ptr = F.gal_ptr[k], &src = F.gal_src[k];
        if (!jds) ptr.assign(nblk + 1, 0);
        // coarse block id (local to the coarse partition) of a target: [0, crows) diagonal blocks, crows + s stored block s
        auto block_of = [&](int32_t ct) -> int64_t { return ct < 0 ? (int64_t)(-1 - ct) : crows + ct / DD; };
        parallel_for(I1 - I0, 512, [&](int64_t a, int64_t b2) {
            for (int64_t I = I0 + a; I < I0 + b2; I++) {
                const auto cb = C.adj_nbr.begin() + C.adj_ptr[I], ce = C.adj_nbr.begin() + C.adj_ptr[I + 1];
                for (int64_t m = C.mem_ptr[I]; m < C.mem_ptr[I + 1]; m++) {
                    const int64_t r = C.mem_idx[m];
                    if (!jds) ptr[(I - c0) + 1]++;                                      // the member's diagonal block
                    for (int64_t p = F.adj_ptr[r]; p < F.adj_ptr[r + 1]; p++) {
                        const int32_t J = F.agg[F.adj_nbr[p]];
                        const int64_t slot = F.adj_slot[p];
                        if (I == J) { F.ctgt[slot] = (int32_t)(-1 - (I - c0)); if (!jds) ptr[(I - c0) + 1]++; continue; }
                        const int64_t q = std::lower_bound(cb, ce, J) - C.adj_nbr.begin();
                        const int64_t cs = C.adj_slot[q] - cs0;
                        if (jds) {
                            const int64_t lane = I & 31;
                            F.ctgt[slot] = (int32_t)((cs - lane) * DD + lane);
                            F.cstr[slot] = C.adj_cnt[q];
                        } else { F.ctgt[slot] = (int32_t)(cs * DD); ptr[crows + cs + 1]++; }
                    }
                }
            }
        });
        if (jds) continue;
        for (int64_t q = 0; q < nblk; q++) ptr[q + 1] += ptr[q];
        src.resize(ptr[nblk]);
        parallel_for(I1 - I0, 512, [&](int64_t a, int64_t b2) {
            std::vector<int32_t> cur;                                                   // fill position of row I's blocks: [0] diagonal, [1 + j] stored block j
            for (int64_t I = I0 + a; I < I0 + b2; I++) {
                const int64_t ns = C.adj_ptr[I + 1] - C.adj_ptr[I], sb = ns > 0 ? C.adj_slot[C.adj_ptr[I]] - cs0 : 0;      // block CSR: the row's slots are contiguous
                cur.assign(1 + ns, 0);
                cur[0] = ptr[I - c0];
                for (int64_t j = 0; j < ns; j++) cur[1 + j] = ptr[crows + sb + j];
                for (int64_t m = C.mem_ptr[I]; m < C.mem_ptr[I + 1]; m++) {
                    const int64_t r = C.mem_idx[m];
                    for (int64_t p = F.adj_ptr[r]; p < F.adj_ptr[r + 1]; p++) {
                        const int64_t slot = F.adj_slot[p], t = block_of(F.ctgt[slot]);
                        src[cur[t < crows ? 0 : 1 + (t - crows - sb)]++] = (int32_t)(slot - fs0);
                    }
                }
                for (int64_t m = C.mem_ptr[I]; m < C.mem_ptr[I + 1]; m++) src[cur[0]++] = (int32_t)(fslots + (C.mem_idx[m] - fr0));
                std::sort(src.begin() + ptr[I - c0], src.begin() + ptr[I - c0 + 1]);
                for (int64_t j = 0; j < ns; j++) std::sort(src.begin() + ptr[crows + sb + j], src.begin() + ptr[crows + sb + j + 1]);
            }
        });
    }
    TICK("  bcl: galerkin contributor lists");
}

// renumber the aggregates of every partition by decreasing coarse degree inside windows (what the sliced storage needs)
static void sort_aggregates_by_degree(const HostLevel &C, std::vector<std::vector<int32_t>> &pagg, const std::vector<int64_t> &pnc, int window) {
    const int world = (int)pnc.size();
    for (int k = 0; k < world; k++) {
        const int64_t n = pnc[k], r0 = C.src_off.empty() ? C.part_off[k] : C.src_off[k];
        std::vector<int32_t> order(n), newid(n);
        std::iota(order.begin(), order.end(), 0);
        for (int64_t w = 0; w < n; w += window) {
            const int64_t e = std::min<int64_t>(n, w + window);
            std::stable_sort(order.begin() + w, order.begin() + e, [&](int32_t a, int32_t b) {
                return C.adj_ptr[r0 + a + 1] - C.adj_ptr[r0 + a] > C.adj_ptr[r0 + b + 1] - C.adj_ptr[r0 + b];
            });
        }
        for (int64_t i = 0; i < n; i++) newid[order[i]] = (int32_t)i;
        for (auto &a : pagg[k]) a = newid[a];
    }
}

// ------------------------------------------------------------------------------------------------
bool build_symbolic(Symbolic &S, const SymbolicOptions &opt,
                    int64_t nv, const uint32_t *vid, const uint8_t *vkind,
                    int64_t ne, const uint8_t *ekind, const uint32_t *efrom, const uint32_t *eto) {
    double t_last = now_s();
    if (nv <= 0 || nv > (int64_t)COL_LOCAL_MASK || ne < 0 || ne > 0x7fffffffll) { S.error = "vertex/edge count out of range"; return false; }
    const int world = opt.world;
    if (world < 1 || world > MAX_RANKS) { S.error = "world size must be 1.." + std::to_string(MAX_RANKS); return false; }
    if (world > 1 && nv < 64 * world) { S.error = "graph too small to shard: needs at least 64 vertices per rank"; return false; }
    S.world = world;
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
    TICK("vertex scan");
    // id -> index (the reference's lut, g2o.rs:60)
    // dense ids (the common case: 0..n-1 in some order) resolve through a direct table, anything else by binary search
    uint32_t max_id = 0;
    for (int64_t v = 0; v < nv; v++) max_id = std::max(max_id, vid[v]);
    const bool dense_ids = (uint64_t)max_id < (uint64_t)nv * 4 + 1024;
    std::vector<int32_t> table;
    std::vector<std::pair<uint32_t, int32_t>> ids;
    if (dense_ids) {
        table.assign((size_t)max_id + 1, -1);
        for (int64_t v = 0; v < nv; v++) {
            if (table[vid[v]] >= 0) { S.error = "duplicate vertex id " + std::to_string(vid[v]); return false; }
            table[vid[v]] = (int32_t)v;
        }
    } else {
        ids.resize(nv);
        for (int64_t v = 0; v < nv; v++) ids[v] = {vid[v], (int32_t)v};
        std::sort(ids.begin(), ids.end());
        for (int64_t v = 1; v < nv; v++) if (ids[v].first == ids[v - 1].first) { S.error = "duplicate vertex id " + std::to_string(ids[v].first); return false; }
    }
    auto lookup = [&](uint32_t id) -> int32_t {
        if (dense_ids) return id <= max_id ? table[id] : -1;
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
    TICK("edge lookup");
    std::vector<int32_t> deg(nv, 0);
    for (int64_t k = 0; k < ne; k++) { deg[S.efrom[k]]++; deg[S.eto[k]]++; }

    // ---- contiguous vertex (lut) ranges per rank, boundaries on multiples of 32, balanced by stored blocks
    S.vrange.assign(world + 1, 0);
    S.vrange[world] = nv;
    if (world > 1) {
        std::vector<int64_t> pre(nv + 1, 0);
        for (int64_t v = 0; v < nv; v++) pre[v + 1] = pre[v] + 1 + deg[v];
        for (int k = 1; k < world; k++) {
            const int64_t target = pre[nv] * k / world;
            int64_t v = std::lower_bound(pre.begin(), pre.end(), target) - pre.begin();
            v = (v + 16) / 32 * 32;
            v = std::max(v, S.vrange[k - 1] + 32);
            v = std::min(v, (nv - 32 * (int64_t)(world - k)) / 32 * 32);
            S.vrange[k] = v;
        }
    }
    // ---- storage order: inside every rank's range, windows of `sort_window` rows sorted by decreasing degree (stable)
    S.levels.clear();
    S.levels.emplace_back();
    {
        HostLevel &L0 = S.levels[0];
        std::vector<int64_t> counts(world);
        for (int k = 0; k < world; k++) counts[k] = S.vrange[k + 1] - S.vrange[k];
        layout_partitions(L0, counts);
        S.perm.assign(L0.n_pad, -1);
        S.iperm.assign(nv, -1);
        int window = std::max(32, opt.sort_window / 32 * 32);
        std::vector<int32_t> idv;
        for (int k = 0; k < world; k++) {
            idv.resize(counts[k]);
            std::iota(idv.begin(), idv.end(), (int32_t)S.vrange[k]);
            parallel_for((counts[k] + window - 1) / window, 8, [&](int64_t w0, int64_t w1) {
                for (int64_t w = w0 * window; w < std::min<int64_t>(w1 * window, counts[k]); w += window) {
                    const int64_t e = std::min<int64_t>(counts[k], w + window);
                    std::stable_sort(idv.begin() + w, idv.begin() + e, [&](int32_t a, int32_t b) { return deg[a] > deg[b]; });
                    for (int64_t i = w; i < e; i++) { S.perm[L0.part_off[k] + i] = idv[i]; S.iperm[idv[i]] = (int32_t)(L0.part_off[k] + i); }
                }
            });
        }
        TICK("storage order");
        // half edges per storage row, sorted by neighbour row
        L0.adj_ptr.assign(L0.n_pad + 1, 0);
        for (int64_t r = 0; r < L0.n_pad; r++) L0.adj_ptr[r + 1] = L0.adj_ptr[r] + (S.perm[r] >= 0 ? deg[S.perm[r]] : 0);
        struct HE { int32_t nbr, edge; uint32_t flags; };
        std::vector<HE> he(L0.adj_ptr[L0.n_pad]);
        {
            // filled by all cores: the order inside a row does not matter here, the row sort below fixes it (thread-count independent)
            std::vector<std::atomic<int32_t>> fill(L0.n_pad);
            parallel_for(L0.n_pad, 65536, [&](int64_t a, int64_t b) { for (int64_t r = a; r < b; r++) fill[r].store(0, std::memory_order_relaxed); });
            parallel_for(ne, 65536, [&](int64_t k0, int64_t k1) {
                for (int64_t k = k0; k < k1; k++) {
                    const int32_t ri = S.iperm[S.efrom[k]], rj = S.iperm[S.eto[k]];
                    const uint32_t xy = ekind[k] == 1 ? COL_EDGE_XY : 0u;
                    he[L0.adj_ptr[ri] + fill[ri].fetch_add(1, std::memory_order_relaxed)] = {rj, (int32_t)k, xy};
                    he[L0.adj_ptr[rj] + fill[rj].fetch_add(1, std::memory_order_relaxed)] = {ri, (int32_t)k, xy | COL_ROLE_TO};
                }
            });
            TICK("half-edge fill");
            parallel_for(L0.n_pad, 4096, [&](int64_t r0, int64_t r1) {
                for (int64_t r = r0; r < r1; r++)
                    std::sort(he.begin() + L0.adj_ptr[r], he.begin() + L0.adj_ptr[r + 1],
                              [](const HE &a, const HE &b) { return a.nbr != b.nbr ? a.nbr < b.nbr : a.edge < b.edge; });
            });
        }
        TICK("row sorts");
        build_jds(L0);
        TICK("build_jds");
        const int64_t nent = L0.adj_ptr[L0.n_pad];
        L0.adj_nbr.resize(nent); L0.adj_flags.resize(nent);
        S.slot_edge.assign(L0.n_slots, -1);
        parallel_for(nent, 65536, [&](int64_t q0, int64_t q1) {
            for (int64_t q = q0; q < q1; q++) {
                L0.adj_nbr[q] = he[q].nbr;
                L0.adj_flags[q] = he[q].flags;
                S.slot_edge[L0.adj_slot[q]] = he[q].edge;
            }
        });
    }
    TICK("level-0 arrays");
    S.dense_coarsest = false;
    if (!opt.build_amg) return true;

    // ---- aggregation hierarchy
    const int64_t repl_max = std::max<int64_t>(opt.repl_max_rows, opt.dense_max);
    while (true) {
        const size_t lvl = S.levels.size() - 1;
        HostLevel &F = S.levels[lvl];
        // a sharded level is never solved directly: in sharded handles the dense coarsest level is always a replicated one
        if (F.n <= opt.dense_max && (world == 1 || F.repl)) { S.dense_coarsest = true; break; }
        if ((int)S.levels.size() >= opt.max_levels) break;
        const int fworld = (int)F.part_real.size();
        std::vector<std::vector<int32_t>> pagg(fworld);
        std::vector<int64_t> pnc(fworld, 0);
        int64_t nc = 0;
        std::vector<uint8_t> lm;                                  // level 0 of an SE2 graph with landmarks: VERTEX_XY rows
        if (lvl == 0 && S.D == 3) {
            bool any = false;
            for (int64_t v = 0; v < nv && !any; v++) any = vkind[v] == 1;
            if (any) { lm.assign(F.n_pad, 0); for (int64_t r = 0; r < F.n_pad; r++) if (S.perm[r] >= 0) lm[r] = vkind[S.perm[r]] == 1; }
        }
        for (int k = 0; k < fworld; k++) {
            pnc[k] = aggregate_partition(F, k, opt.agg_size, pagg[k], lm.empty() ? nullptr : lm.data() + F.part_off[k]);
            nc += pnc[k];
        }
        TICK("aggregate");
        if (nc > 0.8 * F.n) break;                               // coarsening stalled
        const bool merge = fworld > 1 && nc <= repl_max;
        const bool jds = nc >= opt.jds_min_rows && !merge && !F.repl;
        S.levels.emplace_back();
        build_coarse_level(S.levels[lvl], S.levels[lvl + 1], pagg, pnc, false, merge, S.D * S.D);
        TICK("build_coarse_level");
        if (jds) {
            // a large coarse level is streamed like level 0: one thread per row over the sliced storage
            sort_aggregates_by_degree(S.levels[lvl + 1], pagg, pnc, std::max(32, opt.sort_window / 32 * 32));
            S.levels[lvl + 1] = HostLevel();
            build_coarse_level(S.levels[lvl], S.levels[lvl + 1], pagg, pnc, true, merge, S.D * S.D);
        }
        if (S.levels[lvl + 1].n_slots * S.D * S.D > 0x7fffffffll) { S.error = "coarse level exceeds 32-bit Galerkin targets"; return false; }
    }
    return true;
}

} // namespace pgo
