// sgm_host.cpp -- host-side table construction for libsgm_b200.so (no GPU state touched).
//
// Replaces, with O(work) hashing algorithms and bit-identical results, the integer table
// construction of the reference's SGInterpolant constructor:
//   * sgm_host_combination_coeffs : sgmethods/sparse_grid_interpolant.py:77-92
//   * sgm_grid_build              : sgmethods/sparse_grid_interpolant.py:109-139
//   * sgm_host_match_rows         : sgmethods/sparse_grid_interpolant.py:214-231
// The reference searches linearly (window search for the coefficients, an L1 scan over all
// nodes found so far for every tensor-grid node); here multi-indices and nodes are hashed.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "sgm_internal.h"

namespace {

inline uint64_t mix64(uint64_t x) {            // splitmix64 finaliser
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
inline uint64_t pair_hash(uint64_t a, uint64_t b) { return mix64(a * 0x100000001b3ull + mix64(b)); }

// Open-addressing table of int64 payloads keyed by a 64-bit hash; equality is decided by
// the caller (several keys may share a hash, every candidate is offered to `same`).
struct HashIndex {
    std::vector<int64_t> slot;
    std::vector<uint64_t> tag;
    uint64_t mask = 0;
    int64_t used = 0;
    void init(int64_t expected) {
        uint64_t cap = 64;
        while (cap < (uint64_t)expected * 2 + 8) cap <<= 1;
        slot.assign(cap, -1);
        tag.assign(cap, 0);
        mask = cap - 1;
        used = 0;
    }
    void grow() {
        std::vector<int64_t> os;
        std::vector<uint64_t> ot;
        os.swap(slot);
        ot.swap(tag);
        uint64_t cap = (mask + 1) << 1;
        slot.assign(cap, -1);
        tag.assign(cap, 0);
        mask = cap - 1;
        for (size_t i = 0; i < os.size(); ++i) {
            if (os[i] < 0) continue;
            uint64_t p = ot[i] & mask;
            while (slot[p] >= 0) p = (p + 1) & mask;
            slot[p] = os[i];
            tag[p] = ot[i];
        }
    }
    template <class Same>
    int64_t find(uint64_t h, Same same) const {
        uint64_t p = h & mask;
        while (slot[p] >= 0) {
            if (tag[p] == h && same(slot[p])) return slot[p];
            p = (p + 1) & mask;
        }
        return -1;
    }
    void insert(uint64_t h, int64_t value) {
        if ((uint64_t)(used + 1) * 2 > mask + 1) grow();
        uint64_t p = h & mask;
        while (slot[p] >= 0) p = (p + 1) & mask;
        slot[p] = value;
        tag[p] = h;
        ++used;
    }
};

inline uint64_t level_hash(int32_t dim, int64_t level) {
    return level == 0 ? 0ull : pair_hash((uint64_t)dim + 1, (uint64_t)level);
}

// Single-linkage clustering of 1-D values at distance `tol`.  Returns false if a cluster is
// so wide that "same cluster in every direction" would not imply an L1 distance < tol.
struct Clusters {
    std::vector<double> sorted;       // distinct sorted values
    std::vector<int32_t> id_of_sorted;
    int32_t count = 0;
    bool build(std::vector<double> values, double tol, int32_t N) {
        std::sort(values.begin(), values.end());
        values.erase(std::unique(values.begin(), values.end()), values.end());
        sorted.swap(values);
        id_of_sorted.resize(sorted.size());
        double start = 0.0;
        for (size_t i = 0; i < sorted.size(); ++i) {
            if (i == 0 || !(sorted[i] - sorted[i - 1] < tol)) {
                start = sorted[i];
                id_of_sorted[i] = count++;
            } else {
                id_of_sorted[i] = count - 1;
            }
            if ((sorted[i] - start) * (double)N >= tol) return false;
        }
        return true;
    }
    int32_t id(double v) const {
        size_t i = std::lower_bound(sorted.begin(), sorted.end(), v) - sorted.begin();
        return id_of_sorted[i];
    }
};

}  // namespace

struct sgm_grid {
    int64_t T = 0;
    int32_t N = 0;
    int64_t M = 0;
    std::vector<int64_t> map_off;     // T+1
    std::vector<int32_t> maps;        // node id per tensor-grid node
    std::vector<int64_t> row_off;     // M+1  (first-seen non-zero coordinates, CSR)
    std::vector<int32_t> cols;
    std::vector<double> vals;
};

extern "C" {

int sgm_host_combination_coeffs(const int64_t* mid_set, int64_t card, int32_t N,
                                int64_t* coeffs) {
    if (!mid_set || !coeffs || card <= 0 || N <= 0)
        return sgm_fail(SGM_ERR_INVALID, "combination_coeffs: empty multi-index set");
    // validation: non-negative, strictly increasing lexicographically, leading entries
    // contiguous from 0 (the reference's bookmark table indexes by that value, :77-79,84)
    for (int64_t r = 0; r < card; ++r) {
        const int64_t* row = mid_set + r * N;
        for (int32_t n = 0; n < N; ++n)
            if (row[n] < 0) return sgm_fail(SGM_ERR_INVALID, "multi-index with negative entry");
        if (r == 0) {
            if (row[0] != 0)
                return sgm_fail(SGM_ERR_INVALID, "multi-index set must start at leading entry 0");
            continue;
        }
        const int64_t* prev = row - N;
        int32_t n = 0;
        while (n < N && prev[n] == row[n]) ++n;
        if (n == N || prev[n] > row[n])
            return sgm_fail(SGM_ERR_INVALID,
                            "multi-index set must be lexicographically sorted without repeats");
        if (row[0] - prev[0] > 1)
            return sgm_fail(SGM_ERR_INVALID, "leading entries of the set must be contiguous");
    }
    std::vector<uint64_t> hashes(card);
    HashIndex index;
    index.init(card);
    for (int64_t r = 0; r < card; ++r) {
        const int64_t* row = mid_set + r * N;
        uint64_t h = 0;
        for (int32_t n = 0; n < N; ++n) h += level_hash(n, row[n]);
        hashes[r] = h;
        index.insert(h, r);
    }
    std::fill(coeffs, coeffs + card, (int64_t)0);
    std::vector<int64_t> probe(N);
    std::vector<int32_t> supp;
    std::vector<uint64_t> delta;      // hash change when direction supp[i] is lowered by one
    for (int64_t r = 0; r < card; ++r) {
        const int64_t* mu = mid_set + r * N;
        supp.clear();
        delta.clear();
        for (int32_t n = 0; n < N; ++n)
            if (mu[n] > 0) {
                supp.push_back(n);
                delta.push_back(level_hash(n, mu[n] - 1) - level_hash(n, mu[n]));
            }
        const int k = (int)supp.size();
        if (k > 30) return sgm_fail(SGM_ERR_OVERFLOW, "multi-index with more than 30 non-zero entries");
        std::memcpy(probe.data(), mu, sizeof(int64_t) * N);
        // every nu = mu - e, e subset of supp(mu): mu contributes (-1)^{|e|} to c_nu.
        // In a downward-closed set every such nu is present; in a general set absent rows
        // simply receive nothing, which is what the reference's pair search yields too.
        const uint64_t n_sub = 1ull << k;
        for (uint64_t e = 0; e < n_sub; ++e) {
            uint64_t h = hashes[r];
            int bits = 0;
            for (int i = 0; i < k; ++i)
                if (e >> i & 1) {
                    h += delta[i];
                    probe[supp[i]] = mu[supp[i]] - 1;
                    ++bits;
                }
            int64_t hit = index.find(h, [&](int64_t cand) {
                return std::memcmp(mid_set + cand * N, probe.data(), sizeof(int64_t) * N) == 0;
            });
            if (hit >= 0) coeffs[hit] += (bits & 1) ? -1 : 1;
            for (int i = 0; i < k; ++i)
                if (e >> i & 1) probe[supp[i]] = mu[supp[i]];
        }
    }
    return SGM_OK;
}

// Shared implementation: every (term, direction) names a knot family (-1 = single node at 0).
static int grid_build_impl(int64_t T, int32_t N, const int32_t* term_fam, int32_t n_fam,
                           const int64_t* fam_off, const double* fam_knots, double tol,
                           sgm_grid** out) {
    // 1-D coordinate clusters (knots of every family and the 0.0 of inactive directions)
    std::vector<double> pool(fam_knots, fam_knots + (n_fam ? fam_off[n_fam] : 0));
    for (double v : pool)
        if (!std::isfinite(v)) return sgm_fail(SGM_ERR_INVALID, "grid_build: non-finite knot");
    pool.push_back(0.0);
    Clusters cl;
    if (!cl.build(pool, tol, N))
        return sgm_fail(SGM_ERR_NODE_TOL,
                        "1-D knots are not separable at the node tolerance (clusters of knots "
                        "closer than tol are wider than tol / N)");
    const int32_t zero_id = cl.id(0.0);
    std::vector<int32_t> knot_cluster(pool.size() - 1);
    for (size_t i = 0; i + 1 < pool.size(); ++i) knot_cluster[i] = cl.id(fam_knots[i]);

    sgm_grid* g = new sgm_grid();
    g->T = T;
    g->N = N;
    g->map_off.assign(T + 1, 0);
    int64_t total = 0;
    for (int64_t t = 0; t < T; ++t) {
        int64_t S = 1;
        for (int32_t n = 0; n < N; ++n) {
            const int32_t f = term_fam[t * N + n];
            if (f >= n_fam) {
                delete g;
                return sgm_fail(SGM_ERR_INVALID, "grid_build: knot family index out of range");
            }
            if (f >= 0) S *= fam_off[f + 1] - fam_off[f];
            if (S > (int64_t)1 << 40) {
                delete g;
                return sgm_fail(SGM_ERR_OVERFLOW, "grid_build: tensor grid too large");
            }
        }
        total += S;
        g->map_off[t + 1] = total;
    }
    g->maps.resize(total);

    // node table: key = sorted list of (direction, cluster id != cluster of 0)
    std::vector<int64_t> key_off(1, 0);
    std::vector<uint64_t> key_pool;
    HashIndex index;
    index.init(std::max<int64_t>(total / 2, 1024));
    g->row_off.assign(1, 0);

    std::vector<int32_t> adims, asize, afam, digit;
    std::vector<uint64_t> key;
    for (int64_t t = 0; t < T; ++t) {
        adims.clear(); asize.clear(); afam.clear();
        for (int32_t n = 0; n < N; ++n) {
            const int32_t f = term_fam[t * N + n];
            if (f >= 0) {
                adims.push_back(n);
                asize.push_back((int32_t)(fam_off[f + 1] - fam_off[f]));
                afam.push_back(f);
            }
        }
        const int k = (int)adims.size();
        digit.assign(k, 0);
        const int64_t S = g->map_off[t + 1] - g->map_off[t];
        int32_t* tmap = g->maps.data() + g->map_off[t];
        for (int64_t s = 0; s < S; ++s) {
            key.clear();
            uint64_t h = 0x51ed270b27b4f3cfull;
            for (int j = 0; j < k; ++j) {
                int32_t c = knot_cluster[fam_off[afam[j]] + digit[j]];
                if (c != zero_id) {
                    uint64_t kv = ((uint64_t)adims[j] << 32) | (uint32_t)c;
                    key.push_back(kv);
                    h += mix64(kv);
                }
            }
            const size_t klen = key.size();
            int64_t node = index.find(h, [&](int64_t cand) {
                return (size_t)(key_off[cand + 1] - key_off[cand]) == klen &&
                       std::memcmp(key_pool.data() + key_off[cand], key.data(), klen * 8) == 0;
            });
            if (node < 0) {
                node = g->M++;
                if (node > 0x7ffffffe) {
                    delete g;
                    return sgm_fail(SGM_ERR_OVERFLOW, "grid_build: more than 2^31 nodes");
                }
                key_pool.insert(key_pool.end(), key.begin(), key.end());
                key_off.push_back((int64_t)key_pool.size());
                index.insert(h, node);
                // first-seen coordinates (may be e.g. -6e-17 for a centre knot)
                for (int j = 0; j < k; ++j) {
                    double v = fam_knots[fam_off[afam[j]] + digit[j]];
                    uint64_t bits;
                    std::memcpy(&bits, &v, 8);
                    if (bits != 0) { g->cols.push_back(adims[j]); g->vals.push_back(v); }
                }
                g->row_off.push_back((int64_t)g->cols.size());
            }
            tmap[s] = (int32_t)node;
            for (int j = k - 1; j >= 0; --j) {      // C order: last active direction fastest
                if (++digit[j] < asize[j]) break;
                digit[j] = 0;
            }
        }
    }
    *out = g;
    return SGM_OK;
}

int sgm_grid_build(int64_t T, int32_t N, const int32_t* num_nodes, int32_t n_fam,
                   const int32_t* fam_m, const int64_t* fam_off, const double* fam_knots,
                   double tol, sgm_grid** out) {
    if (!out) return sgm_fail(SGM_ERR_INVALID, "grid_build: null output");
    *out = nullptr;
    if (T < 0 || N <= 0 || (T > 0 && !num_nodes) || n_fam < 0)
        return sgm_fail(SGM_ERR_INVALID, "grid_build: bad sizes");
    // family lookup by node count
    int32_t max_m = 1;
    for (int32_t f = 0; f < n_fam; ++f) {
        if (fam_m[f] < 2 || fam_off[f + 1] - fam_off[f] != fam_m[f])
            return sgm_fail(SGM_ERR_INVALID, "grid_build: family table inconsistent");
        max_m = std::max(max_m, fam_m[f]);
    }
    std::vector<int32_t> fam_of_m(max_m + 1, -1);
    for (int32_t f = 0; f < n_fam; ++f) fam_of_m[fam_m[f]] = f;
    std::vector<int32_t> term_fam((size_t)T * N, -1);
    for (int64_t i = 0; i < T * N; ++i) {
        const int32_t m = num_nodes[i];
        if (m < 1 || (m > 1 && (m > max_m || fam_of_m[m] < 0)))
            return sgm_fail(SGM_ERR_INVALID, "grid_build: node count without knot family");
        term_fam[i] = m > 1 ? fam_of_m[m] : -1;
    }
    return grid_build_impl(T, N, term_fam.data(), n_fam, fam_off, fam_knots, tol, out);
}

int sgm_grid_build_families(int64_t T, int32_t N, const int32_t* term_fam, int32_t n_fam,
                            const int64_t* fam_off, const double* fam_knots, double tol,
                            sgm_grid** out) {
    if (!out) return sgm_fail(SGM_ERR_INVALID, "grid_build: null output");
    *out = nullptr;
    if (T < 0 || N <= 0 || (T > 0 && !term_fam) || n_fam < 0 || (n_fam > 0 && (!fam_off || !fam_knots)))
        return sgm_fail(SGM_ERR_INVALID, "grid_build: bad sizes");
    for (int32_t f = 0; f < n_fam; ++f)
        if (fam_off[f + 1] - fam_off[f] < 1)
            return sgm_fail(SGM_ERR_INVALID, "grid_build: empty knot family");
    return grid_build_impl(T, N, term_fam, n_fam, fam_off, fam_knots, tol, out);
}

int64_t sgm_grid_num_nodes(const sgm_grid* g) { return g ? g->M : 0; }
int64_t sgm_grid_map_size(const sgm_grid* g) { return g ? (int64_t)g->maps.size() : 0; }
int64_t sgm_grid_nnz(const sgm_grid* g) { return g ? (int64_t)g->cols.size() : 0; }

int sgm_grid_copy_maps(const sgm_grid* g, int64_t* map_off, int32_t* maps) {
    if (!g || !map_off || (!maps && !g->maps.empty()))
        return sgm_fail(SGM_ERR_INVALID, "grid_copy_maps: null argument");
    std::memcpy(map_off, g->map_off.data(), sizeof(int64_t) * (g->T + 1));
    if (!g->maps.empty()) std::memcpy(maps, g->maps.data(), sizeof(int32_t) * g->maps.size());
    return SGM_OK;
}

int sgm_grid_copy_nodes_dense(const sgm_grid* g, double* SG) {
    if (!g || (!SG && g->M)) return sgm_fail(SGM_ERR_INVALID, "grid_copy_nodes_dense: null argument");
    std::memset(SG, 0, sizeof(double) * (size_t)g->M * g->N);
    for (int64_t r = 0; r < g->M; ++r)
        for (int64_t q = g->row_off[r]; q < g->row_off[r + 1]; ++q)
            SG[r * g->N + g->cols[q]] = g->vals[q];
    return SGM_OK;
}

int sgm_grid_copy_nodes_csr(const sgm_grid* g, int64_t* row_off, int32_t* cols, double* vals) {
    if (!g || !row_off) return sgm_fail(SGM_ERR_INVALID, "grid_copy_nodes_csr: null argument");
    std::memcpy(row_off, g->row_off.data(), sizeof(int64_t) * (g->M + 1));
    if (!g->cols.empty()) {
        if (!cols || !vals) return sgm_fail(SGM_ERR_INVALID, "grid_copy_nodes_csr: null argument");
        std::memcpy(cols, g->cols.data(), sizeof(int32_t) * g->cols.size());
        std::memcpy(vals, g->vals.data(), sizeof(double) * g->vals.size());
    }
    return SGM_OK;
}

void sgm_grid_destroy(sgm_grid* g) { delete g; }

int sgm_host_match_rows(const double* xx, int64_t n, const double* old_xx, int64_t n_old,
                        int32_t N, double tol, int64_t* match) {
    if (n < 0 || n_old < 0 || N <= 0 || (n && (!xx || !match)) || (n_old && !old_xx))
        return sgm_fail(SGM_ERR_INVALID, "match_rows: bad arguments");
    for (int64_t i = 0; i < n; ++i) match[i] = -1;
    if (n == 0 || n_old == 0) return SGM_OK;
    auto l1_close = [&](const double* a, const double* b) {
        double s = 0.0;
        for (int32_t d = 0; d < N; ++d) s += std::fabs(a[d] - b[d]);
        return s < tol;
    };
    std::vector<double> pool;
    pool.reserve((size_t)(n + n_old) * N);
    bool finite = true;
    for (int64_t i = 0; i < n * N; ++i) { pool.push_back(xx[i]); finite &= std::isfinite(xx[i]); }
    for (int64_t i = 0; i < n_old * N; ++i) { pool.push_back(old_xx[i]); finite &= std::isfinite(old_xx[i]); }
    Clusters cl;
    if (!finite || !cl.build(pool, tol, N)) {
        // exact quadratic search, the reference's own algorithm
        for (int64_t i = 0; i < n; ++i) {
            for (int64_t j = 0; j < n_old; ++j)
                if (l1_close(xx + i * N, old_xx + j * N)) {
                    if (match[i] >= 0) return sgm_fail(SGM_ERR_INVALID, "match_rows: several old rows match one node");
                    match[i] = j;
                }
        }
        return SGM_OK;
    }
    auto row_hash = [&](const double* row) {
        uint64_t h = 0;
        for (int32_t d = 0; d < N; ++d) h += pair_hash((uint64_t)d, (uint64_t)cl.id(row[d]));
        return h;
    };
    // several old rows may share a hash (or even a cluster vector): chain them
    HashIndex index;
    index.init(n_old);
    std::vector<int64_t> next(n_old, -1);
    for (int64_t j = 0; j < n_old; ++j) {
        uint64_t h = row_hash(old_xx + j * N);
        int64_t head = index.find(h, [](int64_t) { return true; });
        if (head < 0) {
            index.insert(h, j);
        } else {
            int64_t tail = head;
            while (next[tail] >= 0) tail = next[tail];
            next[tail] = j;
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        uint64_t h = row_hash(xx + i * N);
        int64_t j = index.find(h, [](int64_t) { return true; });
        for (; j >= 0; j = next[j])
            if (l1_close(xx + i * N, old_xx + j * N)) {
                if (match[i] >= 0) return sgm_fail(SGM_ERR_INVALID, "match_rows: several old rows match one node");
                match[i] = j;
            }
    }
    return SGM_OK;
}

}  // extern "C"
