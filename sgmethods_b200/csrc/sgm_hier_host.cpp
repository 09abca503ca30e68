// sgm_hier_host.cpp -- host tables of the grouped hierarchical-surplus kernel ("H2",
// eval_hier_groups_kernel in sgm_kernels.cuh).  No GPU state touched.
//
// The work-reduced form of the pw-linear interpolant on nested knots is
//     I[u](x) = sum_{nu in Lambda} alpha_{nu, j(x)} * prod_{d in supp nu} phi_{nu_d}(x_d)
// (replaces the term loop of sgmethods/sparse_grid_interpolant.py:268-279, see hierarchical.py).
// A multi-index is a path in the trie of (direction, level) ENTRIES in ascending direction;
// by downward closure every prefix of a path is itself in Lambda.  Writing nu = c + e (c = nu
// without its last entry e),
//     I(x) = alpha_0 + sum_{c internal} w_c(x) * sum_{e : c+e in Lambda} phi_e(x) alpha_{c+e}[off_c(x) + S_c j_e(x)],
// where w_c, off_c are the weight / table-offset prefixes of c and S_c the size of c's block.
//
// A GROUP is up to 7 internal nodes c_0..c_{n-1} ("members") with the SAME parent and the same
// level of their last entry, ordered so that their child-entry sets are nested,
// E_0 >= E_1 >= ...: the group sweeps E_0 once, and at step s (entry e_s) exactly the first
// nv(s) members have the child c_g + e_s.  One read of (phi_e, j_e) from the per-point bracket
// tables per step serves nv gathers.
//
// The surpluses are re-laid out for this kernel (the kernel's value table is filled by a
// gather through `hmap`): the blocks (c_g + e_s), g < nv(s), of a step are adjacent slots of
// equal size `slot(s)` starting at `sbase(s)`; slots of at most 16 doubles are padded to a
// power of two and aligned to their size, so that a 32-lane gather into a slot touches ONE
// 128-byte line.  Nothing per (step, member) is stored.  Tables (all int32):
//   group (12 ints): [0] first step, [1] steps | n_mem << 16 | n_par << 20, [2] S_par (size of
//                    the parent's block), [3] S_mem (size of a member's block = stride of the
//                    leaf direction), [4..7] member entries * 64 (8 x u16; unused slots and the
//                    root group's "member = the parent itself" name the DUMMY entry n_entries,
//                    whose hat value is 1 and rank 0), [8..10] 6 x u16: the steps are listed by
//                    decreasing nv (then by slot size), entry i = first step (relative) with
//                    fewer than 7 - i members present; [1] bits 24..29 and [11]: see below;
//   parent (8 ints): word j = entry_j * 64 | extent_j << 16 along the parent's path;
//   step (2 ints):   [0] sbase, [1] entry * 64 | slot << 16.
// EXECUTION is dynamic -- the groups are stored by decreasing cost and the warps of a block take
// them from a shared counter, so that all warps reach the end-of-tile barrier together -- but
// the SUMMATION order is static: the groups are dealt to `n_lists` lists (heaviest first, each
// to the list with the least load; group word [1] bits 24..29 = list, word [11] = position in
// the list), the kernel adds a group's result to its list's accumulator in list order
// (a ticket per list) and adds the lists up in order.  The result bits therefore do not
// depend on which warp ran which group.  split[w] = number of groups in lists < w.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "sgm_internal.h"

namespace {
constexpr int kMaxMembers = 7;
constexpr int kMaxParent = 8;
constexpr int kMaxEntries = 1022;           // entry * 64 (and the dummy entry) in 16 bits
constexpr int64_t kMaxSlot = 0xffff;        // 16 bits in a step record

struct Group {
    int g[12];
    int par[8];
    std::vector<int> step_words;             // 2 per step, sbase relative to the group's base
    int64_t vals = 0;                        // value slots of the group (after alignment)
    double cost = 0.0;
    std::vector<std::pair<int64_t, std::pair<int, int64_t>>> blocks;   // (rel. slot start, (term, size))
};
}

// term_fam: T x N family index per direction (-1 inactive); fam_extent: table entries per
// family (new knots); val_off: T offsets of the terms' surplus blocks in the caller's layout
// (first active direction fastest).  Entry ids are assigned in order of first use over the
// terms (the numbering sgm_plan_create uses).  hmap[j] = index in the caller's layout of value
// slot j of the kernel's layout, or -1 (padding).  Returns false (tables left empty) when the
// set cannot be grouped (a parent multi-index is missing, i.e. the set is not downward closed,
// a path is longer than 9 entries, or a table field overflows).
bool sgm_build_hier_groups(int64_t T, int N, const int32_t* term_fam, int n_fam,
                           const int* fam_extent, const int64_t* val_off,
                           std::vector<int>& groups, std::vector<int>& parents,
                           std::vector<int>& steps, std::vector<int>& hmap, int* root_slot,
                           int n_lists, std::vector<int>& split) {
    split.clear();
    groups.clear();
    parents.clear();
    steps.clear();
    hmap.clear();
    *root_slot = -1;
    if (T <= 0 || n_fam <= 0 || n_lists < 1 || n_lists > 64) return false;
    std::vector<int> entry_of((size_t)N * n_fam, -1);
    std::vector<int> ent_fam;                                 // family of every entry
    std::vector<std::vector<int>> path((size_t)T);
    for (int64_t t = 0; t < T; ++t)
        for (int n = 0; n < N; ++n) {
            const int fam = term_fam[t * N + n];
            if (fam < 0) continue;
            int& e = entry_of[(size_t)n * n_fam + fam];
            if (e < 0) {
                e = (int)ent_fam.size();
                ent_fam.push_back(fam);
            }
            path[t].push_back(e);
        }
    if ((int)ent_fam.size() > kMaxEntries) return false;
    std::map<std::vector<int>, int> node_of;
    for (int64_t t = 0; t < T; ++t) {
        if ((int)path[t].size() > kMaxParent + 1) return false;
        if (!node_of.emplace(path[t], (int)t).second) return false;       // duplicate multi-index
    }
    const auto root_it = node_of.find(std::vector<int>());
    if (root_it == node_of.end()) return false;
    const int root = root_it->second;
    // children of every node: last entry -> child term
    std::vector<std::map<int, int>> kids((size_t)T);
    for (int64_t t = 0; t < T; ++t) {
        if (path[t].empty()) continue;
        std::vector<int> par(path[t].begin(), path[t].end() - 1);
        const auto it = node_of.find(par);
        if (it == node_of.end()) return false;
        kids[it->second][path[t].back()] = (int)t;
    }
    auto extent = [&](int e) { return (int64_t)fam_extent[ent_fam[e]]; };
    auto subset = [&](int a, int b) {                          // E_a subset of E_b
        for (const auto& kv : kids[a])
            if (!kids[b].count(kv.first)) return false;
        return true;
    };
    std::vector<Group> all;
    bool ok = true;
    double cost_w[4] = {800.0, 60.0, 250.0, 10.0};
    if (const char* env = std::getenv("SGM_H2_COST"))
        std::sscanf(env, "%lf,%lf,%lf,%lf", &cost_w[0], &cost_w[1], &cost_w[2], &cost_w[3]);
    auto emit_group = [&](int parent, const std::vector<int>& members) {
        // members: internal nodes whose parent is `parent`, child sets nested in this order;
        // parent < 0: the root itself is the only member
        Group G;
        std::memset(G.g, 0, sizeof G.g);
        std::memset(G.par, 0, sizeof G.par);
        const int n_par = parent >= 0 ? (int)path[parent].size() : 0;
        int64_t S_par = 1;
        const int dummy64 = (int)ent_fam.size() * 64;
        for (int j = 0; j < n_par; ++j) {
            const int e = path[parent][j];
            if (extent(e) > 0xffff) { ok = false; return; }
            G.par[j] = (e * 64) | ((int)extent(e) << 16);
            S_par *= extent(e);
            if (S_par > 0x3fffffff) { ok = false; return; }
        }
        unsigned short* me = reinterpret_cast<unsigned short*>(G.g + 4);
        for (int j = 0; j < 8; ++j) me[j] = (unsigned short)dummy64;
        int64_t S_mem = S_par;
        if (parent >= 0) {
            for (size_t m = 0; m < members.size(); ++m) me[m] = (unsigned short)(path[members[m]].back() * 64);
            S_mem = S_par * extent(path[members[0]].back());
            if (S_mem > 0x3fffffff) { ok = false; return; }
        }
        int64_t leaves = 0;
        // the sweep: entries of the largest set, by decreasing number of members that have them
        std::vector<std::pair<std::pair<int, int64_t>, int>> sweep;   // ((-nv, extent), entry)
        for (const auto& kv : kids[members[0]]) {
            int nv = 0;
            while (nv < (int)members.size() && kids[members[nv]].count(kv.first)) ++nv;
            sweep.push_back({{-nv, extent(kv.first)}, kv.first});
        }
        std::sort(sweep.begin(), sweep.end());
        if (sweep.size() > 0xffff) { ok = false; return; }
        unsigned short* bounds = reinterpret_cast<unsigned short*>(G.g + 8);
        for (int i = 0; i < 6; ++i) {                          // first step with nv < 7 - i
            size_t sidx = 0;
            while (sidx < sweep.size() && -sweep[sidx].first.first >= 7 - i) ++sidx;
            bounds[i] = (unsigned short)sidx;
        }
        for (const auto& sw : sweep) {
            const int e = sw.second;
            const int nv = -sw.first.first;
            const int64_t need = S_mem * extent(e);
            int64_t slot = need;
            if (need <= 16) { slot = 1; while (slot < need) slot <<= 1; }
            else slot = (need + 15) / 16 * 16;
            if (slot > kMaxSlot) { ok = false; return; }
            const int64_t align = std::min<int64_t>(slot, 16);
            G.vals = (G.vals + align - 1) / align * align;
            G.step_words.push_back((int)G.vals);
            G.step_words.push_back((e * 64) | (int)((unsigned)slot << 16));
            for (int g = 0; g < nv; ++g)
                G.blocks.push_back({G.vals + g * slot, {kids[members[g]].at(e), need}});
            G.vals += nv * slot;
            leaves += nv;
        }
        G.g[1] = (int)(G.step_words.size() / 2) | ((int)members.size() << 16) | (n_par << 20);
        G.g[2] = (int)S_par;
        G.g[3] = (int)S_mem;
        // time of a group on one warp, in cycles: a warp is latency-bound, so a step costs
        // about one dependent shared-memory read plus one gather round trip whatever the number
        // of members, and every group pays its header / prefix chain (dev knob SGM_H2_COST =
        // "group,parent entry,step,leaf")
        G.cost = cost_w[0] + cost_w[1] * n_par + cost_w[2] * (G.step_words.size() / 2) + cost_w[3] * leaves;
        all.push_back(std::move(G));
    };
    if (!kids[root].empty()) emit_group(-1, {root});
    std::vector<int> stack{root};
    while (!stack.empty() && ok) {
        const int P = stack.back();
        stack.pop_back();
        std::vector<int> internal;
        for (const auto& kv : kids[P])
            if (!kids[kv.second].empty()) internal.push_back(kv.second);
        // siblings of one level (family) together, larger child sets first
        std::sort(internal.begin(), internal.end(), [&](int a, int b) {
            const int fa = ent_fam[path[a].back()], fb = ent_fam[path[b].back()];
            if (fa != fb) return fa < fb;
            if (kids[a].size() != kids[b].size()) return kids[a].size() > kids[b].size();
            return path[a].back() < path[b].back();
        });
        std::vector<char> used(internal.size(), 0);
        for (size_t i = 0; i < internal.size() && ok; ++i) {
            if (used[i]) continue;
            std::vector<int> members{internal[i]};
            used[i] = 1;
            const int fam = ent_fam[path[internal[i]].back()];
            for (size_t j = i + 1; j < internal.size() && (int)members.size() < kMaxMembers; ++j) {
                if (used[j] || ent_fam[path[internal[j]].back()] != fam) continue;
                if (subset(internal[j], members.back())) {      // keeps the chain nested
                    members.push_back(internal[j]);
                    used[j] = 1;
                }
            }
            emit_group(P, members);
        }
        for (int c : internal) stack.push_back(c);
    }
    if (!ok) return false;
    // deal the groups to the lists: heaviest first, each to the list with the least load
    std::stable_sort(all.begin(), all.end(), [](const Group& a, const Group& b) { return a.cost > b.cost; });
    {
        std::vector<double> load((size_t)n_lists, 0.0);
        std::vector<int> list_of(all.size());
        for (size_t i = 0; i < all.size(); ++i) {
            const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            list_of[i] = w;
            load[w] += all[i].cost;
        }
        split.assign((size_t)n_lists + 1, 0);
        for (size_t i = 0; i < all.size(); ++i) {
            all[i].g[1] |= list_of[i] << 24;
            all[i].g[11] = split[(size_t)list_of[i] + 1]++;        // position in its list
        }
        for (int w = 0; w < n_lists; ++w) split[w + 1] += split[w];
    }
    // value layout: slot 0 = the root's surplus, then the groups (16-aligned) in table order
    int64_t total = 16;
    for (const Group& G : all) total += (G.vals + 15) / 16 * 16;
    if (total > 0x7fffffff) return false;
    hmap.assign((size_t)total, -1);
    hmap[0] = (int)val_off[root];
    *root_slot = 0;
    int64_t base = 16;
    for (Group& G : all) {
        G.g[0] = (int)(steps.size() / 2);
        for (size_t i = 0; i < G.step_words.size(); i += 2) {
            steps.push_back((int)(base + G.step_words[i]));
            steps.push_back(G.step_words[i + 1]);
        }
        for (const auto& b : G.blocks)
            for (int64_t i = 0; i < b.second.second; ++i)
                hmap[(size_t)(base + b.first + i)] = (int)(val_off[b.second.first] + i);
        groups.insert(groups.end(), G.g, G.g + 12);
        parents.insert(parents.end(), G.par, G.par + 8);
        base += (G.vals + 15) / 16 * 16;
    }
    return true;
}

extern "C" {

int sgm_host_hier_groups(int64_t T, int32_t N, const int32_t* term_fam, int32_t n_fam,
                         const int32_t* fam_extent, const int64_t* val_off,
                         int32_t n_lists, int64_t* n_groups, int64_t* n_steps, int64_t* n_slots,
                         int32_t* root_slot, int32_t* groups, int32_t* parents, int32_t* steps,
                         int32_t* hmap, int32_t* split) {
    if (T < 0 || N <= 0 || n_fam < 0 || n_lists < 1 || !n_groups || !n_steps || !n_slots || !root_slot ||
        (T > 0 && (!term_fam || !val_off)) || (n_fam > 0 && !fam_extent))
        return sgm_fail(SGM_ERR_INVALID, "hier_groups: bad argument");
    std::vector<int> g, p, s, h, sp;
    int root = -1;
    if (!sgm_build_hier_groups(T, N, term_fam, n_fam, fam_extent, val_off, g, p, s, h, &root, n_lists, sp))
        return sgm_fail(SGM_ERR_INVALID, "hier_groups: the multi-index set cannot be grouped "
                                         "(not downward closed, or more than 9 active directions)");
    *n_groups = (int64_t)g.size() / 12;
    *n_steps = (int64_t)s.size() / 2;
    *n_slots = (int64_t)h.size();
    *root_slot = root;
    if (groups) std::memcpy(groups, g.data(), g.size() * sizeof(int));
    if (parents) std::memcpy(parents, p.data(), p.size() * sizeof(int));
    if (steps) std::memcpy(steps, s.data(), s.size() * sizeof(int));
    if (hmap) std::memcpy(hmap, h.data(), h.size() * sizeof(int));
    if (split) std::memcpy(split, sp.data(), sp.size() * sizeof(int));
    return SGM_OK;
}

}  // extern "C"
