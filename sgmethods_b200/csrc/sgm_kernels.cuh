// sgm_kernels.cuh -- fp64 evaluation kernels for combination-technique sparse-grid
// interpolants on sm_100a.  No tensor cores: the path is gather / FP64-issue bound (see
// DESIGN.md).  The file is compiled with -fmad=false so that every product and sum is
// rounded separately, in the order of the reference's numpy / scipy arithmetic:
//
//   per term t (reference order):   tp = 0
//       per corner / node in C order: w = ((1*b_0)*b_1)*...   (directions ascending)
//                                     tp = tp + value * w
//   out = out + c_t * tp
//
// Reference semantics implemented here (paths under /root/reference/sgmethods):
//   fill_basis<LINEAR>     scipy find_indices + norm distances, called from
//                          tp_inteprolants.py:49-53 (RegularGridInterpolator, extrapolating)
//   fill_basis<QUADRATIC>  tp_inteprolants.py:100-136
//   fill_basis<CUBIC>      tp_inteprolants.py:206-253
//   fill_basis<LAGRANGE>   tp_inteprolants.py:339-346 (lambda from :311-320, host-computed)
//   term walk              sparse_grid_interpolant.py:268-279, tp_interpolant_wrapper.py:54-64
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "../../include/sgm_b200.h"

namespace sgm {

struct __align__(16) TermHdr { int k; int dim_off; int val_off; int ncorner; };  // one per term
struct __align__(16) TermDim { int boff; int stride; int radix; int ent; };      // one per active dir
struct __align__(16) Entry   { int dim; int knot_off; int n; int boff; };        // (direction, knots)
// Packed 32-byte term record of the split kernel: everything a term with <= 4 active
// directions needs, in two 16-byte uniform loads.  `info` = k | PACKED_FLAG when the
// entry ids and strides fit in 16 bits; otherwise the term goes through hdr/tdim.
struct __align__(16) TermRec {
    double coeff;
    int val_off;
    int info;
    unsigned short ent[4];
    unsigned short stride[4];
};
constexpr int TERMREC_PACKED = 0x100;
// internal kernel flavour: pw-linear plan that has terms with more than 4 active directions;
// every term goes through the nested walk (the 4-direction fast path is compiled out so that
// it keeps its register budget in the common flavour)
constexpr int SGM_LINEAR_WIDE = 4;
__host__ __device__ constexpr int base_kind(int kind) {
    return kind == SGM_LINEAR_WIDE ? (int)SGM_PW_LINEAR : kind;
}
// corner contributions per active direction for the fixed-radix kinds
__host__ __device__ constexpr int fixed_radix(int kind) {
    return kind == SGM_HIER_PW_LINEAR ? 1 : (base_kind(kind) == SGM_PW_LINEAR ? 2 : (kind == SGM_PW_QUADRATIC ? 3 : 4));
}

// Device view of a plan (passed by value to the kernels).
struct PlanDev {
    int kind, N, E, B;                 // B = basis doubles per point
    int T;
    int n_knots;                       // doubles in `knots` (and `lambda`)
    const TermHdr* hdr;
    const TermDim* tdim;
    const TermRec* rec;
    const double* coeff;
    const Entry* ent;
    const int* maps;
    const TermRec* prec;               // prefix kernel: records with first-direction-fastest strides
    const int* gmaps;                  // prefix kernel: node map in that layout (null if not built)
    const unsigned char* run_start;    // prefix kernel <8 warps, 64-term chunks>: [chunk][9] cost-balanced run bounds
    const int* corner_term;           // term of every entry of the global corner list
    const int* corner_off;             // T+1 prefix sums of the terms' corner counts
    int n_corners;                     // sum_t C_t (0 if the list is not built)
    const double* knots;
    const double* lambda;              // Lagrange: barycentric weights; SGM_HIER: per-family new-knot ranks / basis id / denominator
    const TermRec* hrec;               // SGM_HIER_PW_LINEAR: records grouped by number of directions
    int hgroup[6];                     // hrec[hgroup[k] .. hgroup[k+1]) = terms with k directions
    // grouped hierarchical kernel (H2, sgm_hier_host.cpp): group headers, parent paths, 8-byte
    // step records and the kernel's own value layout; null when the set cannot be grouped
    const int4* h2_groups;             // 3 x int4 per group
    const int* h2_parents;             // 8 ints per group
    const int* h2_steps;               // 2 ints per step
    const int* h2_maps;                // value slot of the H2 layout -> row of f (-1: padding)
    const int* h2_split;               // H2_LISTS + 1 group bounds: list w = groups [split[w], split[w+1])
    int h2_n_groups, h2_n_vals;
    int h2_root_off;                   // slot of the root's surplus in the H2 value table (-1: none)
    unsigned* flags;
    // segmented plans (sgm_eval_segments): segment of every term (consecutive, ascending) and
    // the distance in doubles between the outputs of consecutive segments; null otherwise
    const int* seg_id;
    int64_t seg_stride;
    // locality key of a point (<= 16 bits): group i contributes kg_bits[i] bits, the leading bits
    // of a binary search of x[kg_dim[i]] in the finest knot family used in that direction (kg_n
    // knots at knots + kg_off): the coarse end of the bracket index on every level.  Groups are
    // listed most significant first.
    int* tile_counter;                 // grouped kernel: global tile counter for dynamic tile scheduling (null: static)
    int dev_flags;                     // development switches: 1 = strided entry fill, 2 = read all 7 member slots, 4 = no record prefetch
    int n_key;                         // total key bits
    int n_kgrp;
    int kg_dim[16];
    int kg_off[16];
    unsigned short kg_n[16];
    unsigned char kg_bits[16];
};

// ---------------------------------------------------------------------------------------
// 1-D basis tables.  `bas[b * bs]` / `idx[e * bs]` address the slot of this point.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int count_le(const double* g, int n, double x) {
    int lo = 0, hi = n;                          // #{i : g[i] <= x}; NaN -> 0
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (g[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int count_lt_strided(const double* g, int nh, int q, double x) {
    int lo = 0, hi = nh;                         // #{i : g[i*q] < x}; NaN -> 0
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (g[mid * q] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int KIND, typename IdxT>
__device__ __forceinline__ void fill_basis(const Entry e, double x, const double* __restrict__ knots,
                                           const double* __restrict__ lambda, double* bas,
                                           IdxT* idx, int ent_id, int bs, unsigned* flags) {
    const double* g = knots + e.knot_off;
    const int n = e.n;
    double* b = bas + (size_t)e.boff * bs;
    if (KIND == SGM_HIER_PW_LINEAR) {
        // ONE non-zero hierarchical basis function per entry; which one is a property of the
        // knot family (sgm_hier_basis).  `lambda` carries, per family, the new-knot ranks as
        // shorts, the basis id in short n and the one-knot denominator in the last double.
        const short* nr = reinterpret_cast<const short*>(lambda + e.knot_off);
        const int hb = nr[n];
        if (hb == SGM_HB_HAT) {
            // bracket as for pw-linear; exactly one end of the bracket is a knot that is new on
            // this level: its hat value at x
            int i = count_le(g, n, x) - 1;
            i = max(0, min(i, n - 2));
            const double y = (x - g[i]) / (g[i + 1] - g[i]);
            const int r_lo = nr[i];
            b[0] = r_lo >= 0 ? 1.0 - y : y;
            idx[(size_t)ent_id * bs] = (IdxT)(r_lo >= 0 ? r_lo : nr[i + 1]);
        } else if (hb == SGM_HB_QUAD_MID) {
            // patch search of the pw-quadratic operator; the new knot is the patch's middle knot
            const int H = (n + 1) >> 1;
            const int j = max(count_lt_strided(g, H, 2, x) - 1, 0);
            if (j > H - 2 || x != x) {
                atomicOr(flags, SGM_FLAG_QUADRATIC_RIGHT);
                b[0] = __longlong_as_double(0x7ff8000000000000ll);
                idx[(size_t)ent_id * bs] = (IdxT)0;
                return;
            }
            const double x0 = g[2 * j], x1 = g[2 * j + 1], x2 = g[2 * j + 2];
            b[0] = ((x - x0) * (x - x2)) / ((x1 - x0) * (x1 - x2));
            idx[(size_t)ent_id * bs] = (IdxT)nr[2 * j + 1];
        } else {
            // cardinal polynomial of the family's one new knot over all knots of the family
            double num = 1.0;
            bool hit = false;
            for (int i = 0; i < n; ++i) {
                const double dx = x - g[i];
                hit = hit || dx == 0.0;
                if (nr[i] < 0) num = num * dx;
            }
            double v = num / reinterpret_cast<const double*>(lambda + e.knot_off)[n - 1];
            if (hb == SGM_HB_QUAD_END) {
                if (x > g[n - 1] || x != x) {
                    atomicOr(flags, SGM_FLAG_QUADRATIC_RIGHT);
                    v = __longlong_as_double(0x7ff8000000000000ll);
                }
            } else if (hit) {
                v = __longlong_as_double(0x7ff8000000000000ll);   // the reference's 0/0 on a knot
            }
            b[0] = v;
            idx[(size_t)ent_id * bs] = (IdxT)0;
        }
    } else if (base_kind(KIND) == SGM_PW_LINEAR) {
        int i = count_le(g, n, x) - 1;
        i = max(0, min(i, n - 2));
        b[0] = (x - g[i]) / (g[i + 1] - g[i]);
        idx[(size_t)ent_id * bs] = (IdxT)i;
    } else if (KIND == SGM_PW_QUADRATIC) {
        const int H = (n + 1) >> 1;
        int j = max(count_lt_strided(g, H, 2, x) - 1, 0);
        if (j > H - 2 || x != x) {
            // the reference indexes past the knot vector here -> IndexError
            atomicOr(flags, SGM_FLAG_QUADRATIC_RIGHT);
            const double qnan = __longlong_as_double(0x7ff8000000000000ll);
            b[0] = qnan; b[bs] = qnan; b[2 * bs] = qnan;
            idx[(size_t)ent_id * bs] = (IdxT)0;
            return;
        }
        const double x0 = g[2 * j], x1 = g[2 * j + 1], x2 = g[2 * j + 2];
        b[0]      = ((x - x1) * (x - x2)) / ((x0 - x1) * (x0 - x2));
        b[bs]     = ((x - x0) * (x - x2)) / ((x1 - x0) * (x1 - x2));
        b[2 * bs] = ((x - x0) * (x - x1)) / ((x2 - x0) * (x2 - x1));
        idx[(size_t)ent_id * bs] = (IdxT)(2 * j);
    } else if (KIND == SGM_PW_CUBIC) {
        const int H = (n + 2) / 3;
        int j = count_lt_strided(g, H, 3, x) - 1;
        if (x != x) j = H - 2;                    // NaN sorts last in the reference
        j = max(0, min(j, H - 2));
        const double x0 = g[3 * j], x1 = g[3 * j + 1], x2 = g[3 * j + 2], x3 = g[3 * j + 3];
        b[0]      = ((x - x1) * (x - x2) * (x - x3)) / ((x0 - x1) * (x0 - x2) * (x0 - x3));
        b[bs]     = ((x - x0) * (x - x2) * (x - x3)) / ((x1 - x0) * (x1 - x2) * (x1 - x3));
        b[2 * bs] = ((x - x0) * (x - x1) * (x - x3)) / ((x2 - x0) * (x2 - x1) * (x2 - x3));
        b[3 * bs] = ((x - x0) * (x - x1) * (x - x2)) / ((x3 - x0) * (x3 - x1) * (x3 - x2));
        idx[(size_t)ent_id * bs] = (IdxT)(3 * j);
    } else {
        const double* lam = lambda + e.knot_off;
        double ell = x - g[0];
        for (int r = 1; r < n; ++r) ell = ell * (x - g[r]);
        for (int i = 0; i < n; ++i) b[(size_t)i * bs] = (ell * lam[i]) / (x - g[i]);
        idx[(size_t)ent_id * bs] = (IdxT)0;
    }
}

// weight of digit `a` in the direction described by `d`
template <int KIND>
__device__ __forceinline__ double digit_weight(const double* bas, int bs, const TermDim d, int a) {
    if (base_kind(KIND) == SGM_PW_LINEAR) {
        const double y = bas[(size_t)d.boff * bs];
        return a ? y : 1.0 - y;
    }
    return bas[(size_t)(d.boff + a) * bs];
}

// ---------------------------------------------------------------------------------------
// Kernel A: one thread per point, scalar output (NF == 1).  G holds the node values already
// gathered per term (G[val_off + c_order_index]), so a corner is one dependent load.
// Basis tables live in shared memory ([b][thread], conflict-free) or, when they do not fit,
// in a per-block global scratch with the same layout (coalesced, L2 resident).
// ---------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ double linear_term(const TermDim* __restrict__ td, const double* bas,
                                              const unsigned short* idx, int bs,
                                              const double* __restrict__ Gt) {
    double wl[K], wh[K];
    int st[K];
    int base = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const TermDim d = td[j];
        const double y = bas[(size_t)d.boff * bs];
        wh[j] = y;
        wl[j] = 1.0 - y;
        st[j] = d.stride;
        base += (int)idx[(size_t)d.ent * bs] * d.stride;
    }
    double tp = 0.0;
#pragma unroll
    for (int c = 0; c < (1 << K); ++c) {
        double w = 1.0;
        int off = base;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int bit = (c >> (K - 1 - j)) & 1;
            w = w * (bit ? wh[j] : wl[j]);
            off += bit ? st[j] : 0;
        }
        tp = tp + Gt[off] * w;
    }
    return tp;
}

template <int KIND>
__device__ __noinline__ double generic_term(const TermHdr h, const TermDim* __restrict__ td,
                                            const double* bas, const unsigned short* idx, int bs,
                                            const double* __restrict__ Gt) {
    double tp = 0.0;
    for (int c = 0; c < h.ncorner; ++c) {
        double w = 1.0;
        int off = 0;
        int cs = h.ncorner;
        for (int j = 0; j < h.k; ++j) {
            const TermDim d = td[j];
            cs /= d.radix;
            const int a = (c / cs) % d.radix;
            w = w * digit_weight<KIND>(bas, bs, d, a);
            off += ((int)idx[(size_t)d.ent * bs] + a) * d.stride;
        }
        tp = tp + Gt[off] * w;
    }
    return tp;
}

// Nested-loop walk of ALL corner / node contributions of a term with K active directions
// (K compile-time, trip counts run-time): the weight prefix ((1*b_0)*b_1)*... is carried
// down the recursion in registers, so a contribution costs one weight multiply, one value
// multiply and one add -- in C order, i.e. the reference's order (tp_inteprolants.py:146-160,
// :263-277, :339-360).  Used for quadratic / cubic / Lagrange and for pw-linear terms with
// more than 4 active directions.  For Lagrange every lane reads the same node value
// (broadcast load), for the local operators the start index differs per lane.
template <int KIND, int J, int K>
__device__ __forceinline__ void nested_walk(const TermDim* __restrict__ td, const double* bas,
                                            const unsigned short* idx, int bs,
                                            const double* __restrict__ Gt, double w, int off,
                                            double& tp) {
    const TermDim d = td[J];
    const int start = (KIND == SGM_LAGRANGE) ? 0 : (int)idx[(size_t)d.ent * bs];
    const int base = off + start * d.stride;
    for (int a = 0; a < d.radix; ++a) {
        const double wa = w * digit_weight<KIND>(bas, bs, d, a);
        if constexpr (J + 1 == K) {
            tp = tp + Gt[base + a * d.stride] * wa;
        } else {
            nested_walk<KIND, J + 1, K>(td, bas, idx, bs, Gt, wa, base + a * d.stride, tp);
        }
    }
}

template <int KIND>
__device__ __forceinline__ double walk_term_inl(const TermHdr h, const TermDim* __restrict__ td,
                                                const double* bas, const unsigned short* idx,
                                                int bs, const double* __restrict__ Gt) {
    double tp = 0.0;
    switch (h.k) {
        case 1: nested_walk<KIND, 0, 1>(td, bas, idx, bs, Gt, 1.0, 0, tp); break;
        case 2: nested_walk<KIND, 0, 2>(td, bas, idx, bs, Gt, 1.0, 0, tp); break;
        case 3: nested_walk<KIND, 0, 3>(td, bas, idx, bs, Gt, 1.0, 0, tp); break;
        case 4: nested_walk<KIND, 0, 4>(td, bas, idx, bs, Gt, 1.0, 0, tp); break;
        case 5: nested_walk<KIND, 0, 5>(td, bas, idx, bs, Gt, 1.0, 0, tp); break;
        case 6: nested_walk<KIND, 0, 6>(td, bas, idx, bs, Gt, 1.0, 0, tp); break;
        case 7: nested_walk<KIND, 0, 7>(td, bas, idx, bs, Gt, 1.0, 0, tp); break;
        case 8: nested_walk<KIND, 0, 8>(td, bas, idx, bs, Gt, 1.0, 0, tp); break;
        default: tp = generic_term<KIND>(h, td, bas, idx, bs, Gt); break;
    }
    return tp;
}
// pw-linear reaches this only for terms with more than 4 active directions: it keeps the
// simple out-of-line corner loop so that the rare path does not inflate the registers of the
// unrolled fast path; the other kinds use the nested walk.
template <int KIND>
__device__ __forceinline__ double walk_term(const TermHdr h, const TermDim* __restrict__ td,
                                            const double* bas, const unsigned short* idx, int bs,
                                            const double* __restrict__ Gt) {
    if constexpr (KIND == SGM_PW_LINEAR) return generic_term<KIND>(h, td, bas, idx, bs, Gt);
    else return walk_term_inl<KIND>(h, td, bas, idx, bs, Gt);
}

template <int KIND, bool GLOBAL_BASIS>
__global__ void eval_points_kernel(const PlanDev pl, const double* __restrict__ x, int64_t P,
                                   int64_t ldx, const double* __restrict__ G,
                                   double* __restrict__ out, int64_t ldo, double scale,
                                   int accumulate, unsigned char* __restrict__ gscratch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int bs = blockDim.x;
    double* sknots = reinterpret_cast<double*>(smem_raw);
    double* slambda = sknots + pl.n_knots;
    for (int i = tid; i < pl.n_knots; i += bs) {
        sknots[i] = pl.knots[i];
        if (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR) slambda[i] = pl.lambda[i];
    }
    unsigned char* after = reinterpret_cast<unsigned char*>(
        sknots + (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR ? 2 : 1) * (size_t)pl.n_knots);
    double* bas;
    unsigned short* idx;
    if (GLOBAL_BASIS) {
        const size_t per_block = ((size_t)pl.B * 8 + (size_t)pl.E * 2) * bs;
        unsigned char* mine = gscratch + (size_t)blockIdx.x * per_block;
        bas = reinterpret_cast<double*>(mine) + tid;
        idx = reinterpret_cast<unsigned short*>(mine + (size_t)pl.B * 8 * bs) + tid;
    } else {
        bas = reinterpret_cast<double*>(after) + tid;
        idx = reinterpret_cast<unsigned short*>(after + (size_t)pl.B * 8 * bs) + tid;
    }
    __syncthreads();

    const int64_t n_tiles = (P + bs - 1) / bs;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t p = tile * bs + tid;
        if (p >= P) continue;                        // no block-level sync below
        const double* xp = x + p * ldx;
        for (int e = 0; e < pl.E; ++e) {
            const Entry en = pl.ent[e];
            fill_basis<KIND, unsigned short>(en, xp[en.dim], sknots, slambda, bas, idx, e, bs,
                                             pl.flags);
        }
        double acc = 0.0;
        for (int t = 0; t < pl.T; ++t) {
            const TermHdr h = pl.hdr[t];
            const double c = pl.coeff[t];
            const TermDim* td = pl.tdim + h.dim_off;
            const double* Gt = G + h.val_off;
            double tp;
            if (h.k == 0) {
                tp = Gt[0];
            } else if (KIND == SGM_PW_LINEAR && h.k <= 4) {
                switch (h.k) {
                    case 1: tp = linear_term<1>(td, bas, idx, bs, Gt); break;
                    case 2: tp = linear_term<2>(td, bas, idx, bs, Gt); break;
                    case 3: tp = linear_term<3>(td, bas, idx, bs, Gt); break;
                    default: tp = linear_term<4>(td, bas, idx, bs, Gt); break;
                }
            } else {
                tp = walk_term<KIND>(h, td, bas, idx, bs, Gt);
            }
            acc = acc + c * tp;
        }
        double* o = out + p * ldo;
        if (accumulate) *o = *o + scale * acc;
        else *o = (scale == 1.0) ? acc : scale * acc;
    }
}

// ---------------------------------------------------------------------------------------
// Kernel A0 ("small"): pw-linear plans with FEW terms evaluated at MANY points (example/1 as
// shipped: 7 terms, 22 corners per point) -- the regime where the interpolant is HBM-bound:
// 8 (N + 1) algorithmic bytes per point against a few hundred instructions.  One thread per
// point.  What makes it cheap:
//   * all plan metadata (entries, packed term records) travels in the KERNEL PARAMETERS, i.e.
//     through the constant bank into uniform registers: no warp-wide loads of uniform data
//     (a 16-byte uniform LDG costs the L1 data pipe as much as a 512-byte gather);
//   * one binary search per DIRECTION, on its finest knot family; the brackets of the coarser
//     families of that direction are read from a table indexed by the fine count when their
//     knots are bit-for-bit a subset of the fine ones (nested families; else they search);
//   * the bracket coordinate keeps the reference's IEEE division y = (x - g_i) / (g_{i+1} - g_i).
// Same operations per point and output as kernel A: bit-identical.
// ---------------------------------------------------------------------------------------
constexpr int SMALL_MAX_TERMS = 24, SMALL_MAX_ENTRIES = 16;
// dim | n << 16; knots of the family at knot_off (searched families: padded with +inf to
// 2^steps - 1 entries, `steps` = rounds of the branch-free search); from = finest entry of the
// direction (-1: this entry searches) and its bracket table at tab_off
struct __align__(16) SmallEntry { int dim_n, knot_off, from_steps, tab_off; };
struct SmallTerm { double coeff; int val_off, k; unsigned short ent[4], stride[4]; };
struct SmallPlan {
    int T, E, n_knots, n_tab;
    SmallEntry ent[SMALL_MAX_ENTRIES];
    SmallTerm term[SMALL_MAX_TERMS];
};

template <int K>
__device__ __forceinline__ double small_term(const SmallTerm& t, const double* sy, const unsigned short* si,
                                             int bs, const double* G) {
    double wl[K], wh[K];
    unsigned st[K];
    unsigned base = (unsigned)t.val_off;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const double y = sy[(int)t.ent[j] * bs];
        wh[j] = y;
        wl[j] = 1.0 - y;
        st[j] = t.stride[j];
        base += (unsigned)si[(int)t.ent[j] * bs] * st[j];
    }
    double tp = 0.0;
#pragma unroll
    for (int c = 0; c < (1 << K); ++c) {              // corners in itertools.product order
        double w = 1.0;
        unsigned off = base;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int bit = (c >> (K - 1 - j)) & 1;
            w = w * (bit ? wh[j] : wl[j]);
            off += bit ? st[j] : 0u;
        }
        tp = tp + G[off] * w;
    }
    return tp;
}

// PPT points per thread (p, p + THREADS, ...): the uniform bookkeeping (parameter loads, loop
// control, shared-memory addressing) is paid once for PPT points and their dependent chains
// (search, division, corner sums) interleave.
template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS)
eval_points_small_kernel(const __grid_constant__ SmallPlan sp, const double* __restrict__ knots,
                         const unsigned short* __restrict__ tabs, const double* __restrict__ x,
                         int64_t P, int64_t ldx, const int* __restrict__ maps, int n_vals,
                         const double* __restrict__ f, int64_t ldf,
                         double* __restrict__ out, int64_t ldo, double scale, int accumulate) {
    constexpr int BS = THREADS * PPT;                                       // table row length
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sknots = reinterpret_cast<double*>(smem_raw);
    double* sy_all = sknots + sp.n_knots;                                   // [E][BS]
    // the node values of all terms (f gathered through the node maps: the reference's per-term
    // fancy-index copy, sparse_grid_interpolant.py:274) live in shared memory: a corner gather
    // of 32 lanes into a table of a few KB touches most of its lines when it goes through L1
    double* G = sy_all + (size_t)sp.E * BS;                                 // [n_vals]
    unsigned short* si_all = reinterpret_cast<unsigned short*>(G + n_vals);
    unsigned short* stab = si_all + (size_t)sp.E * BS;                      // derived-bracket tables
    for (int i = threadIdx.x; i < sp.n_knots; i += THREADS) sknots[i] = knots[i];
    for (int i = threadIdx.x; i < sp.n_tab; i += THREADS) stab[i] = tabs[i];
    for (int i = threadIdx.x; i < n_vals; i += THREADS) G[i] = f[(int64_t)maps[i] * ldf];
    __syncthreads();
    double* sy = sy_all + threadIdx.x;
    unsigned short* si = si_all + threadIdx.x;
    const int64_t n_tiles = (P + BS - 1) / BS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int64_t p[PPT];
        bool ok[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            const int64_t pp = tile * BS + q * THREADS + threadIdx.x;
            ok[q] = pp < P;
            p[q] = ok[q] ? pp : P - 1;
        }
        // entries are listed direction by direction, the finest family of a direction first
        double xd[PPT];
        int cnt_fine[PPT], cur_dim = -1;
#pragma unroll
        for (int q = 0; q < PPT; ++q) { xd[q] = 0.0; cnt_fine[q] = 0; }
#pragma unroll 1
        for (int e = 0; e < sp.E; ++e) {
            const SmallEntry en = sp.ent[e];
            const int dim = en.dim_n & 0xffff, n = en.dim_n >> 16;
            if (dim != cur_dim) {                                            // warp-uniform
                cur_dim = dim;
#pragma unroll
                for (int q = 0; q < PPT; ++q) xd[q] = x[p[q] * ldx + dim];
            }
            const double* g = sknots + en.knot_off;
            int i[PPT];
            if (en.from_steps >= 0) {
                // #{g <= x} by a branch-free search over the +inf padded family: one load, one
                // compare, one select per round (NaN compares false everywhere: 0)
                int cnt[PPT];
#pragma unroll
                for (int q = 0; q < PPT; ++q) cnt[q] = 0;
                for (int step = 1 << (en.from_steps - 1); step > 0; step >>= 1) {
#pragma unroll
                    for (int q = 0; q < PPT; ++q) cnt[q] += (g[cnt[q] + step - 1] <= xd[q]) ? step : 0;
                }
#pragma unroll
                for (int q = 0; q < PPT; ++q) {
                    cnt_fine[q] = cnt[q];
                    i[q] = max(0, min(cnt[q] - 1, n - 2));
                }
            } else {
#pragma unroll
                for (int q = 0; q < PPT; ++q) i[q] = stab[en.tab_off + cnt_fine[q]];   // nested coarser family
            }
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                sy[e * BS + q * THREADS] = (xd[q] - g[i[q]]) / (g[i[q] + 1] - g[i[q]]);
                si[e * BS + q * THREADS] = (unsigned short)i[q];
            }
        }
        double acc[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) acc[q] = 0.0;
#pragma unroll 1
        for (int t = 0; t < sp.T; ++t) {
            const SmallTerm& tm = sp.term[t];
            double tp[PPT];
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                switch (tm.k) {
                    case 0: tp[q] = G[tm.val_off]; break;
                    case 1: tp[q] = small_term<1>(tm, sy + q * THREADS, si + q * THREADS, BS, G); break;
                    case 2: tp[q] = small_term<2>(tm, sy + q * THREADS, si + q * THREADS, BS, G); break;
                    case 3: tp[q] = small_term<3>(tm, sy + q * THREADS, si + q * THREADS, BS, G); break;
                    default: tp[q] = small_term<4>(tm, sy + q * THREADS, si + q * THREADS, BS, G); break;
                }
            }
#pragma unroll
            for (int q = 0; q < PPT; ++q) acc[q] = acc[q] + tm.coeff * tp[q];
        }
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            if (!ok[q]) continue;
            double* o = out + p[q] * ldo;
            if (accumulate) *o = *o + scale * acc[q];
            else *o = (scale == 1.0) ? acc[q] : scale * acc[q];
        }
    }
}

// ---------------------------------------------------------------------------------------
// Kernel A2 ("split"): scalar output, one block = 32 points x W warps.  The 32 lanes of every
// warp are the 32 points of the tile; the W warps split the TERMS of a chunk among them, so
// the per-point basis table is stored once per tile (not once per thread) and 16+ warps per
// SM hide the gather latency.  Bit-exactness is kept by separating the two roles of the
// reference's loop body  out = out + c_t * tp_t :
//   (1) p_t = c_t * tp_t is computed by any warp, in any order, and parked in shared memory;
//   (2) out = out + p_t is then applied in term order by ONE warp per chunk (rotating).
// ---------------------------------------------------------------------------------------
template <int K, typename IdxT = unsigned short>
__device__ __forceinline__ double linear_term_packed(const int4 r1, const double* bas,
                                                     const IdxT* idx,
                                                     const double* __restrict__ Gt) {
    // r1 = {ent0|ent1<<16, ent2|ent3<<16, stride0|stride1<<16, stride2|stride3<<16};
    // the last active direction always has stride 1 (C order).
    double wl[K], wh[K];
    unsigned st[K];
    unsigned base = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const unsigned e = (unsigned)((j < 2 ? r1.x : r1.y) >> (16 * (j & 1))) & 0xffffu;
        const unsigned sd = (unsigned)((j < 2 ? r1.z : r1.w) >> (16 * (j & 1))) & 0xffffu;
        const double y = bas[e * 32];
        wh[j] = y;
        wl[j] = 1.0 - y;
        st[j] = sd;
        base += (unsigned)idx[e * 32] * sd;
    }
    const double* gp = Gt + base;
    double tp = 0.0;
#pragma unroll
    for (int c = 0; c < (1 << K); c += 2) {
        double w = 1.0;
        unsigned off = 0;
#pragma unroll
        for (int j = 0; j < K - 1; ++j) {
            const int bit = (c >> (K - 1 - j)) & 1;
            w = w * (bit ? wh[j] : wl[j]);
            off += bit ? st[j] : 0u;
        }
        const double* q = gp + off;
        tp = tp + q[0] * (w * wl[K - 1]);
        tp = tp + q[1] * (w * wh[K - 1]);
    }
    return tp;
}

// hierarchical-surplus term: ONE table entry per term, weight = ((1*phi_0)*phi_1)*...
template <int K>
__device__ __forceinline__ double hier_term_packed(const int4 r1, const double* bas,
                                                   const unsigned short* idx,
                                                   const double* __restrict__ At) {
    double w = 1.0;
    unsigned off = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const unsigned e = (unsigned)((j < 2 ? r1.x : r1.y) >> (16 * (j & 1))) & 0xffffu;
        const unsigned sd = (unsigned)((j < 2 ? r1.z : r1.w) >> (16 * (j & 1))) & 0xffffu;
        w = w * bas[e * 32];
        off += (unsigned)idx[e * 32] * sd;
    }
    return At[off] * w;
}

// (an explicit minimum of one block per SM for the nested-walk kinds: without it ptxas held the
// Lagrange instantiations at 64 registers with 224 bytes of spills inside the loop nest)
__host__ __device__ constexpr int split_min_blocks(int kind) {
    return (kind == SGM_PW_QUADRATIC || kind == SGM_PW_CUBIC || kind == SGM_LAGRANGE) ? 1 : 0;
}
// IDX8 (pw-linear plans whose terms are all packed and whose knot families have <= 257 knots):
// bracket indices are stored as bytes, which brings the tables of config 5 (288 entries) under
// the size that lets TWO 16-warp blocks share an SM instead of one 32-warp block.
// SEG (segmented plans, sgm_eval_segments): the warp that adds up the products writes the
// running sum to the output slab of the term's segment whenever a segment ends, and restarts
// from 0 -- one launch evaluates many small independent interpolants (the hierarchical
// surpluses of all refinement candidates) on shared bracket tables.
template <int KIND, int W, int CHUNK, bool IDX8 = false, bool SEG = false>
__global__ void __launch_bounds__(W * 32, split_min_blocks(KIND))
eval_points_split_kernel(const PlanDev pl, const double* __restrict__ x, int64_t P, int64_t ldx,
                         const double* __restrict__ G, double* __restrict__ out, int64_t ldo,
                         double scale, int accumulate, const int* __restrict__ perm) {
    constexpr int chunk = CHUNK;
    static_assert(!IDX8 || KIND == SGM_PW_LINEAR, "byte indices: packed pw-linear plans only");
    using IdxT = typename std::conditional<IDX8, unsigned char, unsigned short>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    static_assert((W & (W - 1)) == 0, "W must be a power of two");
    double* sknots = reinterpret_cast<double*>(smem_raw);
    double* slambda = sknots + pl.n_knots;
    for (int i = threadIdx.x; i < pl.n_knots; i += blockDim.x) {
        sknots[i] = pl.knots[i];
        if (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR) slambda[i] = pl.lambda[i];
    }
    double* bas_all = sknots + (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR ? 2 : 1) * (size_t)pl.n_knots;
    double* sP = bas_all + (size_t)pl.B * 32;                 // [2][chunk][32]
    double* sOut = sP + 2 * (size_t)chunk * 32;               // [32]
    IdxT* idx_all = reinterpret_cast<IdxT*>(sOut + 32);
    const double* bas = bas_all + lane;
    const IdxT* idx = idx_all + lane;
    const int4* rec4 = reinterpret_cast<const int4*>(pl.rec);

    const int64_t n_tiles = (P + 31) >> 5;
    unsigned gc = 0;                                           // chunk counter (block-uniform)
    // dynamic tiles (pl.tile_counter, sorted batches): the next tile comes from a global counter,
    // requested one tile ahead by thread 0
    __shared__ int s_dyn_tile;
    const bool dyn_tiles = pl.tile_counter != nullptr;
    int next_dyn = 0;
    if (dyn_tiles && threadIdx.x == 0) next_dyn = atomicAdd(pl.tile_counter, 1);
    for (int64_t tile = blockIdx.x;; tile += gridDim.x) {
        if (dyn_tiles) {
            if (threadIdx.x == 0) {
                s_dyn_tile = next_dyn;
                if (next_dyn < n_tiles) next_dyn = atomicAdd(pl.tile_counter, 1);
            }
            __syncthreads();
            tile = s_dyn_tile;
        }
        if (tile >= n_tiles) break;
        // lanes of a tile = 32 consecutive points, or 32 consecutive entries of the locality
        // permutation (points whose coarse brackets agree gather from the same cache lines)
        const int64_t slot = (tile << 5) + lane;
        const bool valid = slot < P;
        const int64_t p = perm ? (int64_t)perm[valid ? slot : P - 1] : (valid ? slot : P - 1);
        const double* xp = x + p * ldx;
        __syncthreads();     // previous tile: nobody still reads the basis / knots are staged
        {   // every warp fills a run of CONSECUTIVE entries: the entries are numbered in order of
            // first use, so a run mostly stays in one direction and reads its coordinate once
            // (a per-lane row read costs the L1 data pipe 32 wavefronts)
            const bool strided = pl.dev_flags & 1;
            const int per = (pl.E + W - 1) / W;
            int last_dim = -1;
            double xv = 0.0;
            for (int e = strided ? warp : warp * per; e < (strided ? pl.E : min(pl.E, (warp + 1) * per));
                 e += strided ? W : 1) {
                const Entry en = pl.ent[e];
                if (en.dim != last_dim) { xv = xp[en.dim]; last_dim = en.dim; }
                fill_basis<KIND, IdxT>(en, xv, sknots, slambda, bas_all + lane,
                                       idx_all + lane, e, 32, pl.flags);
            }
        }
        __syncthreads();
        if (pl.T == 0 && warp == 0 && valid) {
            double* o = out + p * ldo;
            *o = accumulate ? *o + scale * 0.0 : scale * 0.0;
        }
        for (int t0 = 0; t0 < pl.T; t0 += chunk, ++gc) {
            double* sp = sP + (size_t)(gc & 1u) * chunk * 32 + lane;
            const int nt = min(chunk, pl.T - t0);
            // the record of a warp's NEXT term is requested before the current term is walked:
            // with few resident tiles (validation samples, small batches) a term is otherwise a
            // chain of two dependent L2 round trips (record, then gathers)
            int4 n0 = make_int4(0, 0, 0, 0), n1 = n0;
            if (warp < nt) {
                n0 = __ldg(rec4 + 2 * (t0 + warp));
                if (KIND == SGM_PW_LINEAR) n1 = __ldg(rec4 + 2 * (t0 + warp) + 1);
            }
            for (int tl = warp; tl < nt; tl += W) {
                const int t = t0 + tl;
                if (pl.dev_flags & 4) {                          // dev switch: no prefetch
                    n0 = __ldg(rec4 + 2 * t);
                    if (KIND == SGM_PW_LINEAR) n1 = __ldg(rec4 + 2 * t + 1);
                }
                const int4 r0 = n0;                              // coeff (x,y), val_off, info
                const int4 r1p = n1;
                if (!(pl.dev_flags & 4) && tl + W < nt) {
                    n0 = __ldg(rec4 + 2 * (t + W));
                    if (KIND == SGM_PW_LINEAR) n1 = __ldg(rec4 + 2 * (t + W) + 1);
                }
                const double coef = __hiloint2double(r0.y, r0.x);
                double tp;
                if (KIND == SGM_PW_LINEAR && (r0.w & TERMREC_PACKED)) {
                    const int4 r1 = r1p;
                    switch (r0.w & 0xff) {
                        case 0: tp = G[r0.z]; break;
                        case 1: tp = linear_term_packed<1, IdxT>(r1, bas, idx, G + r0.z); break;
                        case 2: tp = linear_term_packed<2, IdxT>(r1, bas, idx, G + r0.z); break;
                        case 3: tp = linear_term_packed<3, IdxT>(r1, bas, idx, G + r0.z); break;
                        default: tp = linear_term_packed<4, IdxT>(r1, bas, idx, G + r0.z); break;
                    }
                } else if constexpr (IDX8) {
                    tp = 0.0;                                    // not reached: every term is packed
                } else if (KIND == SGM_HIER_PW_LINEAR && (r0.w & TERMREC_PACKED)) {
                    const int4 r1 = __ldg(rec4 + 2 * t + 1);
                    switch (r0.w & 0xff) {
                        case 0: tp = G[r0.z]; break;
                        case 1: tp = hier_term_packed<1>(r1, bas, idx, G + r0.z); break;
                        case 2: tp = hier_term_packed<2>(r1, bas, idx, G + r0.z); break;
                        case 3: tp = hier_term_packed<3>(r1, bas, idx, G + r0.z); break;
                        default: tp = hier_term_packed<4>(r1, bas, idx, G + r0.z); break;
                    }
                } else {
                    const TermHdr h = pl.hdr[t];
                    tp = h.k == 0 ? G[h.val_off]
                                  : walk_term<KIND>(h, pl.tdim + h.dim_off, bas, idx, 32,
                                                    G + h.val_off);
                }
                sp[tl * 32] = coef * tp;
            }
            __syncthreads();
            if (SEG) {
                if (warp == (int)(gc & (unsigned)(W - 1))) {
                    double o = (t0 == 0) ? 0.0 : sOut[lane];
                    for (int tl = 0; tl < nt; ++tl) {
                        o = o + sp[tl * 32];
                        const int t = t0 + tl;
                        const int sg = __ldg(pl.seg_id + t);
                        if (t + 1 == pl.T || __ldg(pl.seg_id + t + 1) != sg) {      // warp-uniform
                            if (valid) {
                                double* dst = out + (int64_t)sg * pl.seg_stride + p * ldo;
                                *dst = accumulate ? *dst + scale * o : scale * o;
                            }
                            o = 0.0;
                        }
                    }
                    sOut[lane] = o;
                }
            } else if (warp == (int)(gc & (unsigned)(W - 1))) {
                double o = (t0 == 0) ? 0.0 : sOut[lane];
                for (int tl = 0; tl < nt; ++tl) o = o + sp[tl * 32];
                if (t0 + chunk >= pl.T) {
                    if (valid) {
                        double* dst = out + p * ldo;
                        *dst = accumulate ? *dst + scale * o : scale * o;
                    }
                } else {
                    sOut[lane] = o;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Kernel A3 ("prefix"): the split kernel for pw-linear plans whose terms all have <= 4
// active directions -- same block shape (32 points x W warps), same two-role accumulation
// (p_t = c_t * tp_t parked in shared memory, out = out + p_t applied in term order by one
// warp), same roundings.  What changes is who recomputes what:
//   * every warp takes R CONSECUTIVE terms of a chunk.  In a lexicographically sorted index
//     set consecutive terms share their leading (direction, level) pairs, so the weight
//     prefixes  P2[c] = (1*a_0)*a_1,  P3[c] = P2*a_2  and the offset prefixes of the previous
//     term stay in registers: a term reads the bracket tables of its NEW directions only and
//     multiplies each corner weight once (the products are the same floating-point
//     operations on the same operands as when recomputed, hence the same bits);
//   * the value tables G are laid out with the FIRST active direction fastest (gmaps), which
//     makes the offset prefix independent of the directions that follow;
//   * the 32-byte term records of a chunk are staged in shared memory by a coalesced copy
//     (double buffered, prefetched one chunk ahead) and read back as broadcasts, instead of
//     two 16-byte warp-uniform global loads per term;
//   * PPL = 2 points per lane share the decoding of a term and interleave their chains.
// `info` of a record: k | cp << 4, cp = number of leading entries shared with the previous
// term (0 <= cp <= min(k_prev, k - 1), at most 3).
// Tried and measured slower on config 2 (B200): a split chunk barrier (producers arrive on a
// named barrier, only the adding warp waits: 24.96 -> 27.8 ms, the warps drift apart and lose
// their shared L1 lines) and prefetching the next term's value table into L1 (no gain).
// ---------------------------------------------------------------------------------------
template <int W, int R, int BPS, bool BALANCED, int PPL = 1>
__global__ void __launch_bounds__(W * 32, BPS)
eval_points_prefix_kernel(const PlanDev pl, const double* __restrict__ x, int64_t P, int64_t ldx,
                          const double* __restrict__ G, double* __restrict__ out, int64_t ldo,
                          double scale, int accumulate, const int* __restrict__ perm) {
    // PPL = points per lane: a tile is 32 * PPL points, lane l owns points l, l + 32, ...  With
    // PPL = 2 every term is decoded once for two points and the two (independent) summation
    // chains of a lane interleave, which hides the FP64 / gather latency that bounds PPL = 1.
    constexpr int CHUNK = W * R;
    constexpr int TPTS = 32 * PPL;
    static_assert((W & (W - 1)) == 0, "W must be a power of two");
    static_assert(2 * CHUNK <= W * 32, "one thread per 16-byte half record");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int4* sRec = reinterpret_cast<int4*>(smem_raw);                    // [2][CHUNK][2]
    double* sknots = reinterpret_cast<double*>(sRec + 4 * CHUNK);
    double* bas_all = sknots + pl.n_knots;                             // [B][TPTS]
    double* sP = bas_all + (size_t)pl.B * TPTS;                        // [2][CHUNK][TPTS]
    double* sOut = sP + 2 * (size_t)CHUNK * TPTS;                      // [TPTS]
    unsigned short* idx_all = reinterpret_cast<unsigned short*>(sOut + TPTS);
    for (int i = threadIdx.x; i < pl.n_knots; i += blockDim.x) sknots[i] = pl.knots[i];
    const unsigned char* ybase = reinterpret_cast<const unsigned char*>(bas_all + lane);
    const unsigned char* ibase = reinterpret_cast<const unsigned char*>(idx_all + lane);
    const int4* rec4 = reinterpret_cast<const int4*>(pl.prec);
    const int T = pl.T;
    const int n_chunks = (T + CHUNK - 1) / CHUNK;
    if ((int)threadIdx.x < 2 * min(CHUNK, T)) sRec[threadIdx.x] = __ldg(rec4 + threadIdx.x);

    const int64_t n_tiles = (P + TPTS - 1) / TPTS;
    unsigned gc = 0;                                           // chunk counter (block-uniform)
    // dynamic tiles (pl.tile_counter, sorted batches): the next tile comes from a global counter,
    // requested one tile ahead by thread 0
    __shared__ int s_dyn_tile;
    const bool dyn_tiles = pl.tile_counter != nullptr;
    int next_dyn = 0;
    if (dyn_tiles && threadIdx.x == 0) next_dyn = atomicAdd(pl.tile_counter, 1);
    for (int64_t tile = blockIdx.x;; tile += gridDim.x) {
        if (dyn_tiles) {
            if (threadIdx.x == 0) {
                s_dyn_tile = next_dyn;
                if (next_dyn < n_tiles) next_dyn = atomicAdd(pl.tile_counter, 1);
            }
            __syncthreads();
            tile = s_dyn_tile;
        }
        if (tile >= n_tiles) break;
        bool valid[PPL];
        int64_t p[PPL];
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const int64_t slot = tile * TPTS + 32 * j + lane;
            valid[j] = slot < P;
            p[j] = perm ? (int64_t)perm[valid[j] ? slot : P - 1] : (valid[j] ? slot : P - 1);
        }
        __syncthreads();     // previous tile: nobody still reads the basis / knots are staged
        {   // runs of consecutive entries per warp: one coordinate read per direction of the run
            const bool strided = pl.dev_flags & 1;
            const int per = (pl.E + W - 1) / W;
            int last_dim = -1;
            double xv[PPL];
#pragma unroll
            for (int j = 0; j < PPL; ++j) xv[j] = 0.0;
            for (int e = strided ? warp : warp * per; e < (strided ? pl.E : min(pl.E, (warp + 1) * per));
                 e += strided ? W : 1) {
                const Entry en = pl.ent[e];
                if (en.dim != last_dim) {
#pragma unroll
                    for (int j = 0; j < PPL; ++j) xv[j] = x[p[j] * ldx + en.dim];
                    last_dim = en.dim;
                }
#pragma unroll
                for (int j = 0; j < PPL; ++j)
                    fill_basis<SGM_PW_LINEAR, unsigned short>(en, xv[j], sknots, nullptr,
                                                              bas_all + lane + 32 * j,
                                                              idx_all + lane + 32 * j, e, TPTS, pl.flags);
            }
        }
        __syncthreads();
        if (T == 0 && warp == 0) {
#pragma unroll
            for (int j = 0; j < PPL; ++j)
                if (valid[j]) {
                    double* o = out + p[j] * ldo;
                    *o = accumulate ? *o + scale * 0.0 : scale * 0.0;
                }
        }
        for (int c = 0; c < n_chunks; ++c, ++gc) {
            const int t0 = c * CHUNK;
            const int nt = min(CHUNK, T - t0);
            {   // records of the next chunk (cyclic: the last chunk prefetches chunk 0)
                const int tn0 = (c + 1 == n_chunks) ? 0 : t0 + CHUNK;
                const int ntn = min(CHUNK, T - tn0);
                if ((int)threadIdx.x < 2 * ntn)
                    sRec[((gc + 1) & 1u) * 2 * CHUNK + threadIdx.x] = __ldg(rec4 + 2 * tn0 + threadIdx.x);
            }
            const int4* myrec = sRec + (gc & 1u) * 2 * CHUNK;
            double* sp = sP + (size_t)(gc & 1u) * CHUNK * TPTS + lane;
            double P2[PPL][4], P3[PPL][8];
            unsigned B2[PPL], B3[PPL];
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                B2[j] = 0; B3[j] = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) P2[j][i] = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) P3[j][i] = 0.0;
            }
            // this warp's run of consecutive terms: cost-balanced bounds from the plan (the
            // warp that is still adding up the previous chunk gets fewer terms), else R each
            int tl_begin = warp * R, tl_end = min(nt, (warp + 1) * R);
            if (BALANCED) {
                tl_begin = pl.run_start[c * (W + 1) + warp];
                tl_end = pl.run_start[c * (W + 1) + warp + 1];
            }
#pragma unroll 1
            for (int tl = tl_begin; tl < tl_end; ++tl) {
                const int4 r0 = myrec[2 * tl];                  // coeff (x,y), val_off, info
                const int4 r1 = myrec[2 * tl + 1];              // 4 x (entry * 64) u16, 4 x stride u16
                const int k = r0.w & 0xf;
                const int cp = (tl == tl_begin) ? 0 : ((r0.w >> 4) & 0xf);
                // entry fields are pre-scaled byte offsets into a u16 index table [e][32];
                // times PPL they address this tile's index table [e][32 PPL], times 4 PPL the
                // fp64 bracket-coordinate table
                const unsigned f0 = (unsigned)r1.x & 0xffffu, f1 = (unsigned)r1.x >> 16,
                               f2 = (unsigned)r1.y & 0xffffu, f3 = (unsigned)r1.y >> 16;
                const unsigned s1 = (unsigned)r1.z >> 16, s2 = (unsigned)r1.w & 0xffffu,
                               s3 = (unsigned)r1.w >> 16;
#define SGM_Y(fe, j) (*reinterpret_cast<const double*>(ybase + (4u * PPL) * (fe) + 256u * (j)))
#define SGM_I(fe, j) ((unsigned)*reinterpret_cast<const unsigned short*>(ibase + PPL * (fe) + 64u * (j)))
                double tp[PPL];
#pragma unroll
                for (int j = 0; j < PPL; ++j) tp[j] = 0.0;
                if (k == 0) {
#pragma unroll
                    for (int j = 0; j < PPL; ++j) tp[j] = G[(unsigned)r0.z];
                } else if (k == 1) {
#pragma unroll
                    for (int j = 0; j < PPL; ++j) {
                        const double y = SGM_Y(f0, j);
                        const double* q = G + ((unsigned)r0.z + SGM_I(f0, j));
                        tp[j] = tp[j] + q[0] * (1.0 - y);
                        tp[j] = tp[j] + q[1] * y;
                    }
                } else {
                    if (cp < 2) {
#pragma unroll
                        for (int j = 0; j < PPL; ++j) {
                            const double y0 = SGM_Y(f0, j), y1 = SGM_Y(f1, j);
                            const double l0 = 1.0 - y0, l1 = 1.0 - y1;
                            P2[j][0] = l0 * l1; P2[j][1] = l0 * y1; P2[j][2] = y0 * l1; P2[j][3] = y0 * y1;
                            B2[j] = SGM_I(f0, j) + SGM_I(f1, j) * s1;
                        }
                    }
                    if (k == 2) {
#pragma unroll
                        for (int j = 0; j < PPL; ++j) {
                            const double* q = G + ((unsigned)r0.z + B2[j]);
                            const double* q1 = q + s1;
                            tp[j] = tp[j] + q[0] * P2[j][0];
                            tp[j] = tp[j] + q1[0] * P2[j][1];
                            tp[j] = tp[j] + q[1] * P2[j][2];
                            tp[j] = tp[j] + q1[1] * P2[j][3];
                        }
                    } else {
                        if (cp < 3) {
#pragma unroll
                            for (int j = 0; j < PPL; ++j) {
                                const double y2 = SGM_Y(f2, j);
                                const double l2 = 1.0 - y2;
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    P3[j][2 * i] = P2[j][i] * l2;
                                    P3[j][2 * i + 1] = P2[j][i] * y2;
                                }
                                B3[j] = B2[j] + SGM_I(f2, j) * s2;
                            }
                        }
                        if (k == 3) {
                            // corner (b0 b1 b2): address q + b0 + b1 s1 + b2 s2
#pragma unroll
                            for (int j = 0; j < PPL; ++j) {
                                const double* q = G + ((unsigned)r0.z + B3[j]);
                                const double* a[4] = {q, q + s2, q + s1, q + (s1 + s2)};
#pragma unroll
                                for (int i = 0; i < 8; ++i) tp[j] = tp[j] + a[i & 3][i >> 2] * P3[j][i];
                            }
                        } else {
                            const unsigned c12 = s1 + s2;
#pragma unroll
                            for (int j = 0; j < PPL; ++j) {
                                const double y3 = SGM_Y(f3, j);
                                const double l3 = 1.0 - y3;
                                const double* q = G + ((unsigned)r0.z + B3[j] + SGM_I(f3, j) * s3);
                                const double* a[8] = {q, q + s3, q + s2, q + (s2 + s3),
                                                      q + s1, q + (s1 + s3), q + c12, q + (c12 + s3)};
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    tp[j] = tp[j] + a[2 * (i & 3)][i >> 2] * (P3[j][i] * l3);
                                    tp[j] = tp[j] + a[2 * (i & 3) + 1][i >> 2] * (P3[j][i] * y3);
                                }
                            }
                        }
                    }
                }
#undef SGM_Y
#undef SGM_I
                const double coeff = __hiloint2double(r0.y, r0.x);
#pragma unroll
                for (int j = 0; j < PPL; ++j) sp[tl * TPTS + 32 * j] = coeff * tp[j];
            }
            __syncthreads();
            if (warp == (c & (W - 1))) {
                double o[PPL];
#pragma unroll
                for (int j = 0; j < PPL; ++j) o[j] = (c == 0) ? 0.0 : sOut[lane + 32 * j];
                for (int tl = 0; tl < nt; ++tl) {
#pragma unroll
                    for (int j = 0; j < PPL; ++j) o[j] = o[j] + sp[tl * TPTS + 32 * j];
                }
#pragma unroll
                for (int j = 0; j < PPL; ++j) {
                    if (c + 1 == n_chunks) {
                        if (valid[j]) {
                            double* dst = out + p[j] * ldo;
                            *dst = accumulate ? *dst + scale * o[j] : scale * o[j];
                        }
                    } else {
                        sOut[lane + 32 * j] = o[j];
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Kernel H1: dedicated scalar kernel of the opt-in hierarchical-surplus form.  Block = 32
// points x W warps; every warp takes a contiguous range of the (lexicographically sorted)
// multi-indices and keeps the running weight / offset prefixes of the leading directions in
// registers: consecutive multi-indices share all but their last active direction, so a
// term costs two shared-memory reads, one multiply, one table gather and one FMA.  Term
// tables are laid out with the FIRST active direction fastest, which makes the offset prefix
// reusable.  Partial sums of the W warps are added in warp order (deterministic); this form
// makes no bit-exactness claim.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int hrec4_val(const TermRec* hrec, int t) {
    return __ldg(reinterpret_cast<const int4*>(hrec) + 2 * t).z;
}
// all multi-indices with exactly K active directions, split in contiguous ranges over W warps
template <int K, int W>
__device__ __forceinline__ double hier_group(const PlanDev& pl, int warp, const double* bas,
                                             const unsigned short* idx,
                                             const double* __restrict__ A) {
    const int g0 = pl.hgroup[K], g1 = pl.hgroup[K + 1];
    const int per_warp = (g1 - g0 + W - 1) / W;
    const int t_begin = min(g1, g0 + warp * per_warp), t_end = min(g1, t_begin + per_warp);
    const int4* rec4 = reinterpret_cast<const int4*>(pl.hrec);
    double acc = 0.0;
    double w0 = 1.0, w1 = 1.0, w2 = 1.0;
    unsigned o0 = 0, o1 = 0, o2 = 0;
    for (int t = t_begin; t < t_end; ++t) {
        const int4 r0 = __ldg(rec4 + 2 * t);
        const int4 r1 = __ldg(rec4 + 2 * t + 1);
        const int cp = (t == t_begin) ? 0 : ((r0.w >> 12) & 7);
        if (K >= 2 && cp < 1) {
            const unsigned e = (unsigned)r1.x & 0xffffu;
            w0 = bas[e * 32];
            o0 = (unsigned)idx[e * 32] * ((unsigned)r1.z & 0xffffu);
        }
        if (K >= 3 && cp < 2) {
            const unsigned e = (unsigned)r1.x >> 16;
            w1 = w0 * bas[e * 32];
            o1 = o0 + (unsigned)idx[e * 32] * ((unsigned)r1.z >> 16);
        }
        if (K >= 4 && cp < 3) {
            const unsigned e = (unsigned)r1.y & 0xffffu;
            w2 = w1 * bas[e * 32];
            o2 = o1 + (unsigned)idx[e * 32] * ((unsigned)r1.w & 0xffffu);
        }
        // last direction: always evaluated
        const unsigned el = K == 1 ? ((unsigned)r1.x & 0xffffu) : K == 2 ? ((unsigned)r1.x >> 16)
                          : K == 3 ? ((unsigned)r1.y & 0xffffu) : ((unsigned)r1.y >> 16);
        const unsigned sl = K == 1 ? ((unsigned)r1.z & 0xffffu) : K == 2 ? ((unsigned)r1.z >> 16)
                          : K == 3 ? ((unsigned)r1.w & 0xffffu) : ((unsigned)r1.w >> 16);
        const double wp = K == 1 ? 1.0 : K == 2 ? w0 : K == 3 ? w1 : w2;
        const unsigned op = K == 1 ? 0u : K == 2 ? o0 : K == 3 ? o1 : o2;
        const double w = K == 1 ? bas[el * 32] : wp * bas[el * 32];
        const unsigned off = op + (unsigned)idx[el * 32] * sl;
        acc = fma(__ldg(A + r0.z + off), w, acc);
    }
    return acc;
}

template <int W>
__global__ void __launch_bounds__(W * 32)
eval_hier_points_kernel(const PlanDev pl, const double* __restrict__ x, int64_t P, int64_t ldx,
                        const double* __restrict__ A, double* __restrict__ out, int64_t ldo,
                        double scale, int accumulate, const int* __restrict__ perm) {
    constexpr int KIND = SGM_HIER_PW_LINEAR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double* sknots = reinterpret_cast<double*>(smem_raw);
    double* slambda = sknots + pl.n_knots;
    for (int i = threadIdx.x; i < pl.n_knots; i += blockDim.x) {
        sknots[i] = pl.knots[i];
        slambda[i] = pl.lambda[i];
    }
    double* bas_all = sknots + 2 * (size_t)pl.n_knots;
    double* sPart = bas_all + (size_t)pl.B * 32;              // [W][32]
    unsigned short* idx_all = reinterpret_cast<unsigned short*>(sPart + (size_t)W * 32);
    const double* bas = bas_all + lane;
    const unsigned short* idx = idx_all + lane;

    const int64_t n_tiles = (P + 31) >> 5;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t slot = (tile << 5) + lane;
        const bool valid = slot < P;
        const int64_t p = perm ? (int64_t)perm[valid ? slot : P - 1] : (valid ? slot : P - 1);
        const double* xp = x + p * ldx;
        __syncthreads();
        {   // runs of consecutive entries per warp: one coordinate read per direction of the run
            const bool strided = pl.dev_flags & 1;
            const int per = (pl.E + W - 1) / W;
            int last_dim = -1;
            double xv = 0.0;
            for (int e = strided ? warp : warp * per; e < (strided ? pl.E : min(pl.E, (warp + 1) * per));
                 e += strided ? W : 1) {
                const Entry en = pl.ent[e];
                if (en.dim != last_dim) { xv = xp[en.dim]; last_dim = en.dim; }
                fill_basis<KIND, unsigned short>(en, xv, sknots, slambda, bas_all + lane,
                                                 idx_all + lane, e, 32, pl.flags);
            }
        }
        __syncthreads();
        // the multi-indices are grouped by their number k of active directions (the order of
        // the sum is free in this form); inside a group they stay lexicographic, so that
        // consecutive ones share their leading directions.  One specialised loop per k.
        double acc = pl.hgroup[1] > pl.hgroup[0] && warp == 0 ? __ldg(A + hrec4_val(pl.hrec, pl.hgroup[0])) : 0.0;
        acc += hier_group<1, W>(pl, warp, bas, idx, A);
        acc += hier_group<2, W>(pl, warp, bas, idx, A);
        acc += hier_group<3, W>(pl, warp, bas, idx, A);
        acc += hier_group<4, W>(pl, warp, bas, idx, A);
        sPart[warp * 32 + lane] = acc;
        __syncthreads();
        if (warp == 0 && valid) {
            double o = 0.0;
#pragma unroll
            for (int ww = 0; ww < W; ++ww) o = o + sPart[ww * 32 + lane];
            double* dst = out + p * ldo;
            *dst = accumulate ? *dst + scale * o : scale * o;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Kernel H2: grouped hierarchical-surplus kernel (scalar output), the contract path of
// method="auto".  Block = 32 points x W warps, bracket tables (hat value phi_e and new-knot
// rank j_e per entry and point) in shared memory as in H1.  The multi-indices are walked as
// GROUPS of up to 7 trie nodes with a common parent (tables and value layout:
// sgm_hier_host.cpp): a group computes its members' weight / offset prefixes once
// (w_g = w_par * phi, off_g = off_par + S_par * j) and then sweeps the entries of its largest
// child set: ONE read of (phi_e, j_e) and one 8-byte record per step feed nv independent gathers
//     acc_g += alpha[sbase + g * slot + off_g + S_mem * j_e] * phi_e,     g < nv.
// L1 data-pipe cost per gather: the 8-byte surplus (one 128-byte line per warp: slots are
// aligned) plus 1/nv of the bracket-table and record reads.  H1 paid 3 table wavefronts and
// two 16-byte record loads per gather -- and a warp-wide load of UNIFORM data costs the pipe
// 32 x its size, which is why the records here are 8 bytes per step and nothing per gather.
// The warps of a block take groups from a shared counter (listed by decreasing cost), so the
// end-of-tile barrier waits for at most one small group.  The SUM is nevertheless formed in a
// fixed order: every group belongs to one of H2_LISTS lists with a position in it; its result
// is added to the list's shared-memory accumulator when the list's ticket reaches that position
// (the earlier groups of a list were taken earlier, by warps that are running: no deadlock),
// and the lists are added in order.  Same bits whichever warp ran which group.
// The order of the sum is not the reference's: this form makes no bit-exactness claim
// (DESIGN.md 5.1), it is proven <= 1e-12 against the reference.
// ---------------------------------------------------------------------------------------
// Table loads of the grouped kernel: read-only path, kept in L1 in preference to the surplus
// lines (the 0.5 MB value table streams through L1 once per tile and would evict them).
__device__ __forceinline__ int2 ldg_keep(const int2* p) {
    int2 v;
    asm volatile("ld.global.nc.L1::evict_last.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ldg_keep(const int4* p) {
    int4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldg_keep(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::evict_last.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// steps [s0, s1) of a group, all with exactly NV members present: NV independent gathers per
// step, issued before the first use (a warp issues in order: a load followed by its own FMA
// would expose the full L1/L2 latency once per gather).  Address of member g's surplus:
// A + 8 (sbase + S_mem j_e) + 8 (g slot + off_g): two integer multiply-adds, the load and the
// FMA per gather; no branch inside the loop, the next step's record is fetched a step ahead.
template <int NV, int UNROLL>
__device__ __forceinline__ void hier_sweep(const int2* __restrict__ st, int s0, int s1, int s_last,
                                           int2& r, const unsigned char* ybase,
                                           const unsigned char* ibase,
                                           const double* __restrict__ A, int S_mem,
                                           const int (&og)[7], double (&acc)[7]) {
    // r = record of step s0 (sbase, entry * 64 | slot << 16), fetched by the caller / the
    // previous sweep of the group; on return the record of step s1
#pragma unroll UNROLL
    for (int s = s0; s < s1; ++s) {
        const int2 rn = ldg_keep(st + min(s + 1, s_last));   // next record: off the dependency chain
        const unsigned e64 = (unsigned)r.y & 0xffffu;
        const unsigned slot = (unsigned)r.y >> 16;
        const double phi = *reinterpret_cast<const double*>(ybase + 4u * e64);
        const unsigned ix = *reinterpret_cast<const unsigned short*>(ibase + e64);
        const double* a0 = A + (unsigned)(r.x + S_mem * (int)ix);
        double v[NV];
#pragma unroll
        for (int g = 0; g < NV; ++g) {               // address = a0 + 8 * u32: one mad.wide.u32
            const unsigned u = (unsigned)g * slot + (unsigned)og[g];
            unsigned long long addr;
            asm("mad.wide.u32 %0, %1, 8, %2;" : "=l"(addr) : "r"(u), "l"(a0));
            asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v[g]) : "l"(addr));
        }
#pragma unroll
        for (int g = 0; g < NV; ++g) acc[g] = fma(v[g], phi, acc[g]);
        r = rn;
    }
}

constexpr int H2_LISTS = 64;           // summation lists the groups are dealt to
template <int W, int MINB, int UNROLL, bool CONTIG>
__global__ void __launch_bounds__(W * 32, MINB)
eval_hier_groups_kernel(const PlanDev pl, const double* __restrict__ x, int64_t P, int64_t ldx,
                        const double* __restrict__ A, double* __restrict__ out, int64_t ldo,
                        double scale, int accumulate, const int* __restrict__ perm) {
    constexpr int KIND = SGM_HIER_PW_LINEAR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double* sknots = reinterpret_cast<double*>(smem_raw);
    double* slambda = sknots + pl.n_knots;
    for (int i = threadIdx.x; i < pl.n_knots; i += blockDim.x) {
        sknots[i] = pl.knots[i];
        slambda[i] = pl.lambda[i];
    }
    // bracket tables [E + 1][32]: row E is the DUMMY entry (hat value 1, rank 0) that unused
    // member slots and the root group point to, so the member set-up needs no predicates
    double* bas_all = sknots + 2 * (size_t)pl.n_knots;
    unsigned short* idx_all = reinterpret_cast<unsigned short*>(bas_all + (size_t)(pl.B + 1) * 32);
    double* sAcc = reinterpret_cast<double*>(idx_all + (size_t)(pl.E + 1) * 32);   // [H2_LISTS][32]
    volatile int* sTicket = reinterpret_cast<volatile int*>(sAcc + H2_LISTS * 32);    // [H2_LISTS]
    int* sCounter = const_cast<int*>(sTicket) + H2_LISTS;
    const int n_groups = pl.h2_n_groups;
    if (threadIdx.x < 32) {
        bas_all[(size_t)pl.B * 32 + threadIdx.x] = 1.0;
        idx_all[(size_t)pl.E * 32 + threadIdx.x] = 0;
    }
    const unsigned char* ybase = reinterpret_cast<const unsigned char*>(bas_all + lane);
    const unsigned char* ibase = reinterpret_cast<const unsigned char*>(idx_all + lane);
    const int2* all_steps = reinterpret_cast<const int2*>(pl.h2_steps);

    const int64_t n_tiles = (P + 31) >> 5;
    // CONTIG: a block takes a contiguous range of tiles (neighbours in the locality order)
    const int64_t tile_begin = CONTIG ? n_tiles * blockIdx.x / gridDim.x : blockIdx.x;
    const int64_t tile_end = CONTIG ? n_tiles * (blockIdx.x + 1) / gridDim.x : n_tiles;
    const int64_t tile_step = CONTIG ? 1 : gridDim.x;
    // dynamic tiles (pl.tile_counter): a block takes its next tile from a global counter, so that
    // blocks which start late -- their SM was busy with another stream's small kernels, e.g. the
    // validation sample of method="auto" -- simply process fewer tiles
    const bool dyn = !CONTIG && pl.tile_counter != nullptr;
    int64_t tile = tile_begin;
    int next_dyn = 0;                    // thread 0: the tile after this one, requested a tile ahead
    if (dyn && threadIdx.x == 0) next_dyn = atomicAdd(pl.tile_counter, 1);
    for (;;) {
        if (dyn) {
            if (threadIdx.x == 0) {
                sCounter[1] = next_dyn;
                if (next_dyn < n_tiles) next_dyn = atomicAdd(pl.tile_counter, 1);   // used one tile later
            }
            __syncthreads();
            tile = sCounter[1];
            if (tile >= n_tiles) break;
        } else if (tile >= tile_end) {
            break;
        }
        const int64_t this_tile = tile;
        if (!dyn) tile += tile_step;
        const int64_t slot_p = (this_tile << 5) + lane;
        const bool valid = slot_p < P;
        const int64_t p = perm ? (int64_t)perm[valid ? slot_p : P - 1] : (valid ? slot_p : P - 1);
        const double* xp = x + p * ldx;
        __syncthreads();
        {   // runs of consecutive entries per warp: one coordinate read per direction of the run
            const bool strided = pl.dev_flags & 1;
            const int per = (pl.E + W - 1) / W;
            int last_dim = -1;
            double xv = 0.0;
            for (int e = strided ? warp : warp * per; e < (strided ? pl.E : min(pl.E, (warp + 1) * per));
                 e += strided ? W : 1) {
                const Entry en = pl.ent[e];
                if (en.dim != last_dim) { xv = xp[en.dim]; last_dim = en.dim; }
                fill_basis<KIND, unsigned short>(en, xv, sknots, slambda, bas_all + lane,
                                                 idx_all + lane, e, 32, pl.flags);
            }
        }
        for (int i = threadIdx.x; i < H2_LISTS * 32; i += blockDim.x) sAcc[i] = 0.0;
        if (threadIdx.x < H2_LISTS) sTicket[threadIdx.x] = 0;
        if (threadIdx.x == 0) *sCounter = 0;
        __syncthreads();
#pragma unroll 1
        for (;;) {
            int gi = 0;
            if (lane == 0) gi = atomicAdd(sCounter, 1);
            gi = __shfl_sync(0xffffffffu, gi, 0);
            if (gi >= n_groups) break;
            const int4* gh = pl.h2_groups + 3 * (size_t)gi;
            const int4 h0 = ldg_keep(gh);    // first step, steps | n_mem << 16 | n_par << 20, S_par, S_mem
            const int4 h1 = ldg_keep(gh + 1);// member entries * 64 (8 x u16; unused: the dummy entry)
            const int4 h2 = ldg_keep(gh + 2);// step bounds per member count (8 x u16)
            const int n_steps = h0.y & 0xffff, n_par = (h0.y >> 20) & 0xf;   // (bits 24..29: list)
            const int n_mem = (pl.dev_flags & 2) ? 7 : (h0.y >> 16) & 0xf;
            double w_par = 1.0;
            int off_par = 0;
            {                                // the parent's path: entry * 64 | extent << 16 per word
                const int* pw = pl.h2_parents + 8 * (size_t)gi;
                int stride = 1;
#pragma unroll 1
                for (int j = 0; j < n_par; ++j) {
                    const unsigned w = (unsigned)ldg_keep(pw + j);
                    const unsigned e64 = w & 0xffffu;
                    w_par = w_par * *reinterpret_cast<const double*>(ybase + 4u * e64);
                    off_par += (int)*reinterpret_cast<const unsigned short*>(ibase + e64) * stride;
                    stride *= (int)(w >> 16);
                }
            }
            double acc[7];
            int og[7];
            const unsigned me[4] = {(unsigned)h1.x, (unsigned)h1.y, (unsigned)h1.z, (unsigned)h1.w};
#pragma unroll
            for (int g = 0; g < 7; ++g) {
                const unsigned e64 = (g & 1) ? me[g >> 1] >> 16 : me[g >> 1] & 0xffffu;
                acc[g] = 0.0;
                // unused member slots (group-uniform) skip their table reads: 3 wavefronts each
                og[g] = g < n_mem ? off_par + h0.z * (int)*reinterpret_cast<const unsigned short*>(ibase + e64)
                                  : off_par;
            }
            // steps are listed by decreasing number of present members; bound i = first step
            // (relative) with fewer than 7 - i members
            const int2* st = all_steps + h0.x;
            const int S_mem = h0.w;
            const int e7 = h2.x & 0xffff, e6 = (unsigned)h2.x >> 16, e5 = h2.y & 0xffff,
                      e4 = (unsigned)h2.y >> 16, e3 = h2.z & 0xffff, e2 = (unsigned)h2.z >> 16;
            int2 r = ldg_keep(st);
            hier_sweep<7, UNROLL>(st, 0, e7, n_steps - 1, r, ybase, ibase, A, S_mem, og, acc);
            hier_sweep<6, UNROLL>(st, e7, e6, n_steps - 1, r, ybase, ibase, A, S_mem, og, acc);
            hier_sweep<5, UNROLL>(st, e6, e5, n_steps - 1, r, ybase, ibase, A, S_mem, og, acc);
            hier_sweep<4, UNROLL>(st, e5, e4, n_steps - 1, r, ybase, ibase, A, S_mem, og, acc);
            hier_sweep<3, UNROLL>(st, e4, e3, n_steps - 1, r, ybase, ibase, A, S_mem, og, acc);
            hier_sweep<2, UNROLL>(st, e3, e2, n_steps - 1, r, ybase, ibase, A, S_mem, og, acc);
            hier_sweep<1, UNROLL>(st, e2, n_steps, n_steps - 1, r, ybase, ibase, A, S_mem, og, acc);
            // total += w_par * sum_g phi(e_g) acc_g; the members' hat values are read again here
            // instead of living in 14 registers during the sweeps (unused slots: acc = 0)
            double gsum = 0.0;
#pragma unroll
            for (int g = 0; g < 7; ++g) {
                const unsigned e64 = (g & 1) ? me[g >> 1] >> 16 : me[g >> 1] & 0xffffu;
                if (g < n_mem) gsum = fma(*reinterpret_cast<const double*>(ybase + 4u * e64), acc[g], gsum);
            }
            {   // add to the group's list in list order (ticket = position in the list)
                const int list = (h0.y >> 24) & 0x3f, pos = h2.w;
                while (sTicket[list] != pos) {}
                volatile double* slot = sAcc + list * 32 + lane;
                *slot = *slot + w_par * gsum;
                __threadfence_block();
                __syncwarp();
                if (lane == 0) sTicket[list] = pos + 1;
            }
        }
        __syncthreads();
        if (warp == 0 && valid) {
            double o = pl.h2_root_off >= 0 ? __ldg(A + pl.h2_root_off) : 0.0;
#pragma unroll
            for (int ww = 0; ww < H2_LISTS; ++ww) o = o + sAcc[ww * 32 + lane];
            double* dst = out + p * ldo;
            *dst = accumulate ? *dst + scale * o : scale * o;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Locality sort of the evaluation points (counting sort on a <= 16-bit key).  Each point is
// evaluated independently, so the order only changes which points share a warp: with equal
// coarse brackets the 32 lanes of a gather hit 1-2 cache lines instead of 4-5.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned locality_key(const PlanDev& pl, const double* xp) {
    unsigned key = 0;
    for (int i = 0; i < pl.n_kgrp; ++i) {
        const double xv = xp[pl.kg_dim[i]];
        const double* g = pl.knots + pl.kg_off[i];
        int lo = 0, hi = pl.kg_n[i];
        for (int b = 0; b < pl.kg_bits[i]; ++b) {
            const int mid = (lo + hi) >> 1;
            const bool up = xv >= __ldg(g + min(mid, (int)pl.kg_n[i] - 1));
            key = (key << 1) | (up ? 1u : 0u);
            if (up) lo = mid + 1; else hi = mid;
        }
    }
    return key;
}
__global__ void sort_hist_kernel(const PlanDev pl, const double* __restrict__ x, int64_t P,
                                 int64_t ldx, unsigned* __restrict__ hist,
                                 unsigned short* __restrict__ keys, int* __restrict__ tile_counter) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *tile_counter = 0;   // for the kernel that follows the sort
    // warp-aggregated counting: lanes with the same key elect one lane that adds their number
    // (clustered points fall into a handful of bins: one atomic per lane serialised on them)
    for (int64_t p0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) & ~(int64_t)31; p0 < P;
         p0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = p0 + (threadIdx.x & 31);
        const bool act = p < P;
        const unsigned key = act ? locality_key(pl, x + p * ldx) : 0xffffffffu;
        if (act) keys[p] = (unsigned short)key;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (act && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hist + key, (unsigned)__popc(peers));
    }
}
// exclusive scan of the bins, two levels: every block scans 1024 consecutive bins in place
// (coalesced) and publishes its total; the scatter kernel adds the totals of the blocks
// before a bin's block (at most 64 of them).
__global__ void __launch_bounds__(1024) sort_scan_kernel(unsigned* __restrict__ hist, int n,
                                                         unsigned* __restrict__ block_total) {
    __shared__ unsigned wsum[32];
    const int i = blockIdx.x * 1024 + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned v = i < n ? hist[i] : 0u;
    unsigned inc = v;                                          // inclusive scan inside the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned w = wsum[lane], winc = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, winc, off);
            if (lane >= off) winc += u;
        }
        wsum[lane] = winc - w;                                 // exclusive prefix of the warps
        if (lane == 31) block_total[blockIdx.x] = winc;
    }
    __syncthreads();
    if (i < n) hist[i] = wsum[warp] + inc - v;
}
__global__ void sort_scatter_kernel(const unsigned short* __restrict__ keys, int64_t P,
                                    unsigned* __restrict__ offs,
                                    const unsigned* __restrict__ block_total, int n_blocks,
                                    int* __restrict__ perm) {
    __shared__ unsigned base[64];
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int b = 0; b < n_blocks; ++b) { base[b] = run; run += block_total[b]; }
    }
    __syncthreads();
    for (int64_t p0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) & ~(int64_t)31; p0 < P;
         p0 += (int64_t)gridDim.x * blockDim.x) {
        const int lane = threadIdx.x & 31;
        const int64_t p = p0 + lane;
        const bool act = p < P;
        const unsigned key = act ? keys[p] : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        unsigned first = 0;
        if (act && lane == leader) first = atomicAdd(offs + key, (unsigned)__popc(peers));
        first = __shfl_sync(0xffffffffu, first, leader);
        if (act) perm[base[key >> 10] + first + __popc(peers & ((1u << lane) - 1u))] = (int)p;
    }
}

// G[j, 0:NF] = f[maps[j], 0:NF]  (the reference's per-term fancy-index copy,
// sparse_grid_interpolant.py:274, done once for all terms)
__global__ void pregather_kernel(const int* __restrict__ maps, int64_t n_vals,
                                 const double* __restrict__ f, int64_t ldf, int64_t NF,
                                 double* __restrict__ G) {
    const int64_t total = n_vals * NF;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = i / NF, c = i - j * NF;
        const int row = maps[j];                       // -1: padding slot of an aligned layout
        G[i] = row >= 0 ? f[(int64_t)row * ldf + c] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------
// Kernel B: vector-valued output.  One warp per point, lanes across output columns
// (coalesced row-segment gathers of f_on_SG through the node map), CPL columns per lane.
// The 32 lanes first compute the weights and rows of 32 corners in parallel, then the
// warp walks them in reference order, sharing weight and row by shuffles.
// ---------------------------------------------------------------------------------------
template <int KIND, int CPL, bool SEG = false>
__global__ void eval_cols_kernel(const PlanDev pl, const double* __restrict__ x, int64_t P,
                                 int64_t ldx, const double* __restrict__ f, int64_t ldf,
                                 int64_t NF, double* __restrict__ out, int64_t ldo,
                                 double scale, int accumulate, int pts_per_block) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;
    double* sknots = reinterpret_cast<double*>(smem_raw);
    double* slambda = sknots + pl.n_knots;
    for (int i = threadIdx.x; i < pl.n_knots; i += blockDim.x) {
        sknots[i] = pl.knots[i];
        if (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR) slambda[i] = pl.lambda[i];
    }
    double* sbas_all = sknots + (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR ? 2 : 1) * (size_t)pl.n_knots;
    double* bas = sbas_all + (size_t)warp * pl.B;
    int* idx = reinterpret_cast<int*>(sbas_all + (size_t)n_warps * pl.B) + (size_t)warp * pl.E;
    __syncthreads();

    const int64_t col0 = (int64_t)blockIdx.y * (32 * CPL) + lane;
    const int64_t p_begin = (int64_t)blockIdx.x * pts_per_block;
    const int64_t p_end = min(P, p_begin + pts_per_block);
    for (int64_t p = p_begin + warp; p < p_end; p += n_warps) {
        const double* xp = x + p * ldx;
        __syncwarp();
        for (int e = lane; e < pl.E; e += 32) {
            const Entry en = pl.ent[e];
            fill_basis<KIND, int>(en, xp[en.dim], sknots, slambda, bas, idx, e, 1, pl.flags);
        }
        __syncwarp();
        double acc[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[c] = 0.0;
        for (int t = 0; t < pl.T; ++t) {
            const TermHdr h = pl.hdr[t];
            const double coef = pl.coeff[t];
            const TermDim* td = pl.tdim + h.dim_off;
            const int* tmap = pl.maps + h.val_off;
            double tp[CPL];
            if (h.k == 0) {
                const double* fr = f + (int64_t)tmap[0] * ldf;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int64_t col = col0 + 32 * c;
                    tp[c] = col < NF ? fr[col] : 0.0;
                }
            } else {
#pragma unroll
                for (int c = 0; c < CPL; ++c) tp[c] = 0.0;
                for (int c_base = 0; c_base < h.ncorner; c_base += 32) {
                    const int cid = c_base + lane;
                    double w = 1.0;
                    int row = 0;
                    if (cid < h.ncorner) {
                        int off = 0;
                        if (KIND == SGM_LAGRANGE) {
                            int cs = h.ncorner;
                            for (int j = 0; j < h.k; ++j) {
                                const TermDim d = td[j];
                                cs /= d.radix;
                                const int a = (cid / cs) % d.radix;
                                w = w * digit_weight<KIND>(bas, 1, d, a);
                                off += (idx[d.ent] + a) * d.stride;
                            }
                        } else {
                            // fixed radix 2 / 3 / 4: digits by constant division, most
                            // significant digit = first active direction
                            constexpr int R = fixed_radix(KIND);
                            int cs = h.ncorner;
                            int rem = cid;
                            for (int j = 0; j < h.k; ++j) {
                                const TermDim d = td[j];
                                cs /= R;
                                const int a = rem / cs;
                                rem -= a * cs;
                                w = w * digit_weight<KIND>(bas, 1, d, a);
                                off += (idx[d.ent] + a) * d.stride;
                            }
                        }
                        row = tmap[off];
                    }
                    const int n_in = min(32, h.ncorner - c_base);
                    for (int cc = 0; cc < n_in; ++cc) {
                        const double wv = __shfl_sync(0xffffffffu, w, cc);
                        const int r = __shfl_sync(0xffffffffu, row, cc);
                        const double* fr = f + (int64_t)r * ldf;
#pragma unroll
                        for (int c = 0; c < CPL; ++c) {
                            const int64_t col = col0 + 32 * c;
                            if (col < NF) tp[c] = tp[c] + __ldg(fr + col) * wv;
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CPL; ++c) acc[c] = acc[c] + coef * tp[c];
            if (SEG) {                              // segment ends: store its sum, restart
                const int sg = __ldg(pl.seg_id + t);
                if (t + 1 == pl.T || __ldg(pl.seg_id + t + 1) != sg) {
                    double* o = out + (int64_t)sg * pl.seg_stride + p * ldo;
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        const int64_t col = col0 + 32 * c;
                        if (col < NF) o[col] = accumulate ? o[col] + scale * acc[c] : scale * acc[c];
                        acc[c] = 0.0;
                    }
                }
            }
        }
        if (SEG) continue;
        double* o = out + p * ldo;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int64_t col = col0 + 32 * c;
            if (col < NF) {
                if (accumulate) o[col] = o[col] + scale * acc[c];
                else o[col] = (scale == 1.0) ? acc[c] : scale * acc[c];
            }
        }
    }
}

// sumsq[s] = sum over the P x NF values of segment s of v^2, in a fixed order (thread-strided
// partial sums, then a shared-memory tree): the mean-square norm of the Gerstner-Griebel
// indicators without a device -> host copy of the surpluses.
__global__ void __launch_bounds__(256)
segment_sumsq_kernel(const double* __restrict__ vals, int64_t P, int64_t NF, int64_t ld,
                     int64_t seg_stride, double* __restrict__ sumsq) {
    __shared__ double part[256];
    const double* v = vals + (int64_t)blockIdx.x * seg_stride;
    const int64_t total = P * NF;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < total; i += 256) {
        const int64_t r = i / NF, c = i - r * NF;
        const double a = v[r * ld + c];
        acc = acc + a * a;
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
        if ((int)threadIdx.x < h) part[threadIdx.x] = part[threadIdx.x] + part[threadIdx.x + h];
        __syncthreads();
    }
    if (threadIdx.x == 0) sumsq[blockIdx.x] = part[0];
}

// ---------------------------------------------------------------------------------------
// Kernel B3 ("wide"): vector-valued output, full column tiles only.  Same structure as
// eval_cols_kernel (one warp per point, 32 lanes compute weight + row of 32 corners, then
// the warp walks them in order) but every lane owns CPL2 pairs of adjacent columns and
// fetches them with 16-byte loads, without column guards: per corner 3 shuffles, one
// address and CPL2 LDG.128 feed 4*CPL2 FP64 operations, so the FP64 pipe (not instruction
// issue) becomes the limiter.  The ragged last columns go through eval_cols_kernel.
// Requires 16-byte aligned rows: f, out 16-byte aligned, ldf and ldo even.
// ---------------------------------------------------------------------------------------
template <int KIND, int CPL2>
__global__ void __launch_bounds__(256)
eval_cols_wide_kernel(const PlanDev pl, const double* __restrict__ x, int64_t P, int64_t ldx,
                      const double* __restrict__ f, int64_t ldf, double* __restrict__ out,
                      int64_t ldo, double scale, int accumulate, int pts_per_block) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;
    double* sknots = reinterpret_cast<double*>(smem_raw);
    double* slambda = sknots + pl.n_knots;
    for (int i = threadIdx.x; i < pl.n_knots; i += blockDim.x) {
        sknots[i] = pl.knots[i];
        if (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR) slambda[i] = pl.lambda[i];
    }
    double* sbas_all = sknots + (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR ? 2 : 1) * (size_t)pl.n_knots;
    double* bas = sbas_all + (size_t)warp * pl.B;
    int* idx = reinterpret_cast<int*>(sbas_all + (size_t)n_warps * pl.B) + (size_t)warp * pl.E;
    __syncthreads();

    // lane owns columns tile*64*CPL2 + 64*c + 2*lane + {0, 1}, c < CPL2
    const int64_t col0 = (int64_t)blockIdx.y * (64 * CPL2) + 2 * lane;
    const double* fcol = f + col0;
    const int64_t p_begin = (int64_t)blockIdx.x * pts_per_block;
    const int64_t p_end = min(P, p_begin + pts_per_block);
    for (int64_t p = p_begin + warp; p < p_end; p += n_warps) {
        const double* xp = x + p * ldx;
        __syncwarp();
        for (int e = lane; e < pl.E; e += 32) {
            const Entry en = pl.ent[e];
            fill_basis<KIND, int>(en, xp[en.dim], sknots, slambda, bas, idx, e, 1, pl.flags);
        }
        __syncwarp();
        double2 acc[CPL2], tp[CPL2];
#pragma unroll
        for (int c = 0; c < CPL2; ++c) {
            acc[c] = make_double2(0.0, 0.0);
            tp[c] = make_double2(0.0, 0.0);
        }
        // The corner contributions of ALL terms form one list (term order, C order inside a
        // term).  32 lanes set up 32 consecutive entries of the list at a time -- possibly
        // belonging to several small terms -- then the warp walks them in order; when the
        // last corner of a term has been added, out = out + c_t * tp is applied and tp reset.
        for (int g0 = 0; g0 < pl.n_corners; g0 += 32) {
            const int g = g0 + lane;
            double w = 1.0, coef = 0.0;
            int row = 0, last = 0;
            if (g < pl.n_corners) {
                const int t = pl.corner_term[g];
                const TermHdr h = pl.hdr[t];
                const int cid = g - pl.corner_off[t];
                const TermDim* td = pl.tdim + h.dim_off;
                int off = 0;
                if (KIND == SGM_LAGRANGE) {
                    int cs = h.ncorner;
                    for (int j = 0; j < h.k; ++j) {
                        const TermDim d = td[j];
                        cs /= d.radix;
                        const int a = (cid / cs) % d.radix;
                        w = w * digit_weight<KIND>(bas, 1, d, a);
                        off += (idx[d.ent] + a) * d.stride;
                    }
                } else {
                    constexpr int R = fixed_radix(KIND);
                    int cs = h.ncorner;
                    int rem = cid;
                    for (int j = 0; j < h.k; ++j) {
                        const TermDim d = td[j];
                        cs /= R;
                        const int a = rem / cs;
                        rem -= a * cs;
                        w = w * digit_weight<KIND>(bas, 1, d, a);
                        off += (idx[d.ent] + a) * d.stride;
                    }
                }
                row = pl.maps[h.val_off + off];
                last = (cid == h.ncorner - 1);
                if (last) coef = pl.coeff[t];
            }
            const int n_in = min(32, pl.n_corners - g0);
            for (int cc = 0; cc < n_in; ++cc) {
                const double wv = __shfl_sync(0xffffffffu, w, cc);
                const int r = __shfl_sync(0xffffffffu, row, cc);
                const int is_last = __shfl_sync(0xffffffffu, last, cc);
                const double* fr = fcol + (int64_t)r * ldf;
                double2 v[CPL2];
#pragma unroll
                for (int c = 0; c < CPL2; ++c)
                    v[c] = __ldg(reinterpret_cast<const double2*>(fr + 64 * c));
#pragma unroll
                for (int c = 0; c < CPL2; ++c) {
                    tp[c].x = tp[c].x + v[c].x * wv;
                    tp[c].y = tp[c].y + v[c].y * wv;
                }
                if (is_last) {
                    const double cf = __shfl_sync(0xffffffffu, coef, cc);
#pragma unroll
                    for (int c = 0; c < CPL2; ++c) {
                        acc[c].x = acc[c].x + cf * tp[c].x;
                        acc[c].y = acc[c].y + cf * tp[c].y;
                        tp[c] = make_double2(0.0, 0.0);
                    }
                }
            }
        }
        double* o = out + p * ldo + col0;
#pragma unroll
        for (int c = 0; c < CPL2; ++c) {
            double2* dst = reinterpret_cast<double2*>(o + 64 * c);
            double2 r = make_double2(scale * acc[c].x, scale * acc[c].y);
            if (accumulate) {
                const double2 old = *dst;
                r.x = old.x + r.x;
                r.y = old.y + r.y;
            }
            *dst = r;
        }
    }
}

// EXPERIMENT (BASELINE's "TMA-staged when NF is large", dev knob SGM_WIDE_BULK=1; rejected, see
// DESIGN.md 7.1 and profiles/r2_cols_bulk_ncu_full.txt): the one-point-per-warp wide kernel of
// the hierarchical kind with the 2 KB row segments staged in shared memory by bulk asynchronous
// copies (cp.async.bulk 1-D, one mbarrier per stage, a ring of WIDE_BULK_STAGES stages per warp,
// lane 0 of each warp is its producer) instead of per-lane LDG.128 through L1.  Same arithmetic
// as eval_cols_wide_kernel<SGM_HIER, 4> (bit-identical).
constexpr int WIDE_BULK_STAGES = 4;
__global__ void __launch_bounds__(256)
eval_cols_wide_bulk_kernel(const PlanDev pl, const double* __restrict__ x, int64_t P, int64_t ldx,
                           const double* __restrict__ f, int64_t ldf, double* __restrict__ out,
                           int64_t ldo, double scale, int accumulate, int pts_per_block, int64_t ncols) {
    constexpr int KIND = SGM_HIER_PW_LINEAR;
    constexpr int CPL2 = 4, D = WIDE_BULK_STAGES;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;
    // [ring: n_warps x D x 2 KB][mbarriers: n_warps x D][knots | lambda | basis tables | indices]
    double* ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * D * 256;
    const unsigned ring0 = (unsigned)__cvta_generic_to_shared(ring);
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(
        smem_raw + (size_t)n_warps * D * 2048 + (size_t)warp * D * 8);
    double* sknots = reinterpret_cast<double*>(smem_raw + (size_t)n_warps * D * 2048 + (size_t)n_warps * D * 8);
    double* slambda = sknots + pl.n_knots;
    for (int i = threadIdx.x; i < pl.n_knots; i += blockDim.x) {
        sknots[i] = pl.knots[i];
        slambda[i] = pl.lambda[i];
    }
    double* sbas_all = sknots + 2 * (size_t)pl.n_knots;
    double* bas = sbas_all + (size_t)warp * pl.B;
    int* idx = reinterpret_cast<int*>(sbas_all + (size_t)n_warps * pl.B) + (size_t)warp * pl.E;
    int* srow = idx + (size_t)(n_warps - warp) * pl.E + (size_t)warp * 32;      // [n_warps][32] after the indices
    if (lane == 0) {
        for (int s = 0; s < D; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t colbase = (int64_t)blockIdx.y * 256;
    const int tile_cols = (int)min((int64_t)256, ncols - colbase);           // even (launcher)
    const unsigned tile_bytes = (unsigned)tile_cols * 8u;
    const int64_t col0 = colbase + 2 * lane;
    bool okc[CPL2];
#pragma unroll
    for (int c = 0; c < CPL2; ++c) okc[c] = 2 * lane + 64 * c + 1 < tile_cols;
    unsigned n_issued = 0, n_used = 0;                    // entries issued / consumed by this warp so far
    const int64_t p_begin = (int64_t)blockIdx.x * pts_per_block;
    const int64_t p_end = min(P, p_begin + pts_per_block);
    for (int64_t p = p_begin + warp; p < p_end; p += n_warps) {
        const double* xp = x + p * ldx;
        __syncwarp();
        for (int e = lane; e < pl.E; e += 32) {
            const Entry en = pl.ent[e];
            fill_basis<KIND, int>(en, xp[en.dim], sknots, slambda, bas, idx, e, 1, pl.flags);
        }
        __syncwarp();
        double2 acc[CPL2];
#pragma unroll
        for (int c = 0; c < CPL2; ++c) acc[c] = make_double2(0.0, 0.0);
        for (int g0 = 0; g0 < pl.n_corners; g0 += 32) {
            const int g = g0 + lane;
            double w = 0.0;
            int row = 0;
            if (g < pl.n_corners) {
                const int t = pl.corner_term[g];
                const TermHdr h = pl.hdr[t];
                const TermDim* td = pl.tdim + h.dim_off;
                int off = 0;
                w = pl.coeff[t];
                for (int j = 0; j < h.k; ++j) {
                    const TermDim d = td[j];
                    w = w * bas[d.boff];
                    off += idx[d.ent] * d.stride;
                }
                row = pl.maps[h.val_off + off];
            }
            __syncwarp();
            srow[lane] = row;
            __syncwarp();
            const int n_in = min(32, pl.n_corners - g0);
            auto issue = [&](int cc) {                      // lane 0: row of entry cc -> next stage
                const unsigned s = n_issued % D;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s), "r"(tile_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(ring0 + 2048u * s), "l"(f + (int64_t)srow[cc] * ldf + colbase), "r"(tile_bytes),
                               "r"(bar0 + 8u * s) : "memory");
            };
            for (int i = 0; i < min(D - 1, n_in); ++i) {
                if (lane == 0) issue(i);
                ++n_issued;
            }
            for (int cc = 0; cc < n_in; ++cc) {
                if (cc + D - 1 < n_in) {
                    if (lane == 0) issue(cc + D - 1);
                    ++n_issued;
                }
                const unsigned s = n_used % D, parity = (n_used / D) & 1u;
                unsigned done = 0;
                while (!done)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(done) : "r"(bar0 + 8u * s), "r"(parity) : "memory");
                ++n_used;
                const double wv = __shfl_sync(0xffffffffu, w, cc);
                const double* seg = ring + (size_t)s * 256 + 2 * lane;
#pragma unroll
                for (int c = 0; c < CPL2; ++c) {
                    const double2 v = okc[c] ? *reinterpret_cast<const double2*>(seg + 64 * c) : make_double2(0.0, 0.0);
                    acc[c].x = fma(v.x, wv, acc[c].x);
                    acc[c].y = fma(v.y, wv, acc[c].y);
                }
                __syncwarp();                               // stage s may be refilled from the next iteration on
            }
        }
        double* o = out + p * ldo + col0;
#pragma unroll
        for (int c = 0; c < CPL2; ++c) {
            if (!okc[c]) continue;
            double2* dst = reinterpret_cast<double2*>(o + 64 * c);
            double2 r = make_double2(scale * acc[c].x, scale * acc[c].y);
            if (accumulate) {
                const double2 old = *dst;
                r.x = old.x + r.x;
                r.y = old.y + r.y;
            }
            *dst = r;
        }
    }
}

// Kernel C3 ("wide pairs"): the wide kernel with TWO points per warp.  The points come in the
// order of the locality permutation, so the two points of a warp mostly lie in the same
// brackets of the coarse levels and their corner contributions read the SAME rows of f: a row
// is then loaded once and used for both points (each with its own weight, accumulated in its
// own registers, in the same order as when evaluated alone -- bit-identical).  This halves
// the L1 -> register traffic that bounds the wide kernel for every shared row and doubles the
// FP64 work in flight per load.
template <int KIND, int CPL2, bool GUARD>
__device__ __forceinline__ void
cols_wide2_body(const PlanDev& pl, const double* __restrict__ x, int64_t P, int64_t ldx,
                const double* __restrict__ f, int64_t ldf, double* __restrict__ out,
                int64_t ldo, double scale, int accumulate, int pairs_per_block,
                const int* __restrict__ perm, int64_t ncols) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;
    double* sknots = reinterpret_cast<double*>(smem_raw);
    double* slambda = sknots + pl.n_knots;
    for (int i = threadIdx.x; i < pl.n_knots; i += blockDim.x) {
        sknots[i] = pl.knots[i];
        if (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR) slambda[i] = pl.lambda[i];
    }
    double* sbas_all = sknots + (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR ? 2 : 1) * (size_t)pl.n_knots;
    double* basA = sbas_all + (size_t)(2 * warp) * pl.B;
    double* basB = basA + pl.B;
    // per-warp staging of 32 list entries: (wA, wB) | (rowA + last flag, rowB) | coefficient;
    // written once per batch by the lane that set the entry up, read back as broadcasts
    // (3 shared-memory wavefronts per entry instead of 7 warp shuffles)
    unsigned char* stage = reinterpret_cast<unsigned char*>(sbas_all + (size_t)(2 * n_warps) * pl.B);
    stage += (16 - (reinterpret_cast<uintptr_t>(stage) & 15)) & 15;      // 16-byte entries
    double2* sW = reinterpret_cast<double2*>(stage) + (size_t)warp * 32;
    double* sC = reinterpret_cast<double*>(stage + (size_t)n_warps * 32 * 16) + (size_t)warp * 32;
    int2* sR = reinterpret_cast<int2*>(stage + (size_t)n_warps * 32 * 24) + (size_t)warp * 32;
    int* idxA = reinterpret_cast<int*>(stage + (size_t)n_warps * 32 * 32) + (size_t)(2 * warp) * pl.E;
    int* idxB = idxA + pl.E;
    __syncthreads();

    // lane owns columns tile*64*CPL2 + 64*c + 2*lane + {0, 1}, c < CPL2
    const int64_t col0 = (int64_t)blockIdx.y * (64 * CPL2) + 2 * lane;
    const double* fcol = f + col0;
    // GUARD (ragged last tile only): column pairs at or beyond ncols (even) are neither read
    // nor written; full tiles run the unguarded instantiation
    bool okc[CPL2];
#pragma unroll
    for (int c = 0; c < CPL2; ++c) okc[c] = !GUARD || col0 + 64 * c + 1 < ncols;
    const int64_t n_pairs = (P + 1) >> 1;
    const int64_t q_begin = (int64_t)blockIdx.x * pairs_per_block;
    const int64_t q_end = min(n_pairs, q_begin + pairs_per_block);
    for (int64_t q = q_begin + warp; q < q_end; q += n_warps) {
        const bool hasB = 2 * q + 1 < P;
        const int64_t pA = perm ? (int64_t)perm[2 * q] : 2 * q;
        const int64_t pB = hasB ? (perm ? (int64_t)perm[2 * q + 1] : 2 * q + 1) : pA;
        const double* xA = x + pA * ldx;
        const double* xB = x + pB * ldx;
        __syncwarp();
        for (int e = lane; e < pl.E; e += 32) {
            const Entry en = pl.ent[e];
            fill_basis<KIND, int>(en, xA[en.dim], sknots, slambda, basA, idxA, e, 1, pl.flags);
            fill_basis<KIND, int>(en, xB[en.dim], sknots, slambda, basB, idxB, e, 1, pl.flags);
        }
        __syncwarp();
        double2 accA[CPL2], tpA[CPL2], accB[CPL2], tpB[CPL2];
#pragma unroll
        for (int c = 0; c < CPL2; ++c) {
            accA[c] = make_double2(0.0, 0.0); tpA[c] = make_double2(0.0, 0.0);
            accB[c] = make_double2(0.0, 0.0); tpB[c] = make_double2(0.0, 0.0);
        }
        for (int g0 = 0; g0 < pl.n_corners; g0 += 32) {
            const int g = g0 + lane;
            double wA = 1.0, wB = 1.0, coef = 0.0;
            int rowA = 0, rowB = 0, last = 0;
            if (g < pl.n_corners) {
                const int t = pl.corner_term[g];
                const TermHdr h = pl.hdr[t];
                const int cid = g - pl.corner_off[t];
                const TermDim* td = pl.tdim + h.dim_off;
                int offA = 0, offB = 0;
                int cs = h.ncorner;
                int rem = cid;
                for (int j = 0; j < h.k; ++j) {
                    const TermDim d = td[j];
                    int a;
                    if (KIND == SGM_LAGRANGE) {
                        cs /= d.radix;
                        a = (cid / cs) % d.radix;
                    } else {
                        constexpr int R = fixed_radix(KIND);
                        cs /= R;
                        a = rem / cs;
                        rem -= a * cs;
                    }
                    wA = wA * digit_weight<KIND>(basA, 1, d, a);
                    wB = wB * digit_weight<KIND>(basB, 1, d, a);
                    offA += (idxA[d.ent] + a) * d.stride;
                    offB += (idxB[d.ent] + a) * d.stride;
                }
                rowA = pl.maps[h.val_off + offA];
                rowB = pl.maps[h.val_off + offB];
                last = (cid == h.ncorner - 1);
                if (last) coef = pl.coeff[t];
                if (KIND == SGM_HIER_PW_LINEAR) { wA = wA * coef; wB = wB * coef; }   // folded into the weight
            }
            __syncwarp();                                        // previous batch fully read
            sW[lane] = make_double2(wA, wB);
            sR[lane] = make_int2(rowA | (last << 31), rowB);
            sC[lane] = coef;
            __syncwarp();
            const int n_in = min(32, pl.n_corners - g0);
            for (int cc = 0; cc < n_in; ++cc) {
                const double2 wv = sW[cc];
                const int2 rr = sR[cc];
                const double wvA = wv.x, wvB = wv.y;
                const int rA = rr.x & 0x7fffffff, rB = rr.y;
                const int is_last = rr.x < 0;
                const double* frA = fcol + (int64_t)rA * ldf;
                double2 v[CPL2];
#pragma unroll
                for (int c = 0; c < CPL2; ++c)
                    v[c] = okc[c] ? __ldg(reinterpret_cast<const double2*>(frA + 64 * c)) : make_double2(0.0, 0.0);
                // hierarchical-surplus kind: every term is ONE table entry with coefficient 1, and
                // the form makes no bit-exactness claim: acc = fma(v, w, acc), one FP64 operation
                // per element instead of three (product, coefficient multiply, add)
#pragma unroll
                for (int c = 0; c < CPL2; ++c) {
                    if (KIND == SGM_HIER_PW_LINEAR) {
                        accA[c].x = fma(v[c].x, wvA, accA[c].x);
                        accA[c].y = fma(v[c].y, wvA, accA[c].y);
                    } else {
                        tpA[c].x = tpA[c].x + v[c].x * wvA;
                        tpA[c].y = tpA[c].y + v[c].y * wvA;
                    }
                }
                if (rA != rB) {                                   // warp-uniform
                    const double* frB = fcol + (int64_t)rB * ldf;
#pragma unroll
                    for (int c = 0; c < CPL2; ++c)
                        v[c] = okc[c] ? __ldg(reinterpret_cast<const double2*>(frB + 64 * c)) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int c = 0; c < CPL2; ++c) {
                    if (KIND == SGM_HIER_PW_LINEAR) {
                        accB[c].x = fma(v[c].x, wvB, accB[c].x);
                        accB[c].y = fma(v[c].y, wvB, accB[c].y);
                    } else {
                        tpB[c].x = tpB[c].x + v[c].x * wvB;
                        tpB[c].y = tpB[c].y + v[c].y * wvB;
                    }
                }
                if (KIND != SGM_HIER_PW_LINEAR && is_last) {
                    const double cf = sC[cc];
#pragma unroll
                    for (int c = 0; c < CPL2; ++c) {
                        accA[c].x = accA[c].x + cf * tpA[c].x;
                        accA[c].y = accA[c].y + cf * tpA[c].y;
                        tpA[c] = make_double2(0.0, 0.0);
                        accB[c].x = accB[c].x + cf * tpB[c].x;
                        accB[c].y = accB[c].y + cf * tpB[c].y;
                        tpB[c] = make_double2(0.0, 0.0);
                    }
                }
            }
        }
#pragma unroll
        for (int pt = 0; pt < 2; ++pt) {
            if (pt == 1 && !hasB) break;
            double* o = out + (pt == 0 ? pA : pB) * ldo + col0;
#pragma unroll
            for (int c = 0; c < CPL2; ++c) {
                if (!okc[c]) continue;
                double2* dst = reinterpret_cast<double2*>(o + 64 * c);
                const double2 a2 = pt == 0 ? accA[c] : accB[c];
                double2 r = make_double2(scale * a2.x, scale * a2.y);
                if (accumulate) {
                    const double2 old = *dst;
                    r.x = old.x + r.x;
                    r.y = old.y + r.y;
                }
                *dst = r;
            }
        }
    }
}

template <int KIND, int CPL2>
__global__ void __launch_bounds__(256, 2)
eval_cols_wide2_kernel(const PlanDev pl, const double* __restrict__ x, int64_t P, int64_t ldx,
                       const double* __restrict__ f, int64_t ldf, double* __restrict__ out,
                       int64_t ldo, double scale, int accumulate, int pairs_per_block,
                       const int* __restrict__ perm, int64_t ncols) {
    if (((int64_t)blockIdx.y + 1) * (64 * CPL2) <= ncols)        // block-uniform
        cols_wide2_body<KIND, CPL2, false>(pl, x, P, ldx, f, ldf, out, ldo, scale, accumulate,
                                           pairs_per_block, perm, ncols);
    else
        cols_wide2_body<KIND, CPL2, true>(pl, x, P, ldx, f, ldf, out, ldo, scale, accumulate,
                                          pairs_per_block, perm, ncols);
}

// Kernel C4 ("wide quads"): FOUR points per warp on 128-column tiles (CPL2 = 2: a lane owns two
// pairs of adjacent columns).  Same list walk as the pair kernel; the four points are consecutive
// in the locality order, and a row of f is loaded again only when it differs from the row of the
// previous point of the quad, so a row shared by the whole quad is fetched ONCE for 4 x 4 FP64
// multiply-adds per lane: the L1 -> register traffic per multiply-add falls to a quarter of the
// one-point kernel's.  Per point the operations and their order are those of the one-point kernel
// (bit-identical); the hierarchical-surplus kind, which makes no bit-exactness claim, contracts
// multiply and add into one FMA.  MEASURED SLOWER than the pair kernel on 256-column tiles
// (config 3: x0.81, config 4: x0.88, tools/exp_quads.py): the kernel is bound by load latency at 16
// warps per SM, not by L1 bandwidth, and 128-column tiles repeat the per-point set-up twice as
// often; kept behind SGM_WIDE_PTS=4 as a record of the experiment.
template <int KIND, int CPL2, bool GUARD>
__device__ __forceinline__ void
cols_wide4_body(const PlanDev& pl, const double* __restrict__ x, int64_t P, int64_t ldx,
                const double* __restrict__ f, int64_t ldf, double* __restrict__ out,
                int64_t ldo, double scale, int accumulate, int quads_per_block,
                const int* __restrict__ perm, int64_t ncols) {
    constexpr int NP = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;
    double* sknots = reinterpret_cast<double*>(smem_raw);
    double* slambda = sknots + pl.n_knots;
    for (int i = threadIdx.x; i < pl.n_knots; i += blockDim.x) {
        sknots[i] = pl.knots[i];
        if (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR) slambda[i] = pl.lambda[i];
    }
    double* sbas_all = sknots + (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR ? 2 : 1) * (size_t)pl.n_knots;
    double* bas0 = sbas_all + (size_t)(NP * warp) * pl.B;
    // per-warp staging of 32 list entries: 4 weights (32 B) | 4 rows, last flag in the sign bit of
    // the first (16 B) | coefficient (8 B)
    unsigned char* stage = reinterpret_cast<unsigned char*>(sbas_all + (size_t)(NP * n_warps) * pl.B);
    stage += (16 - (reinterpret_cast<uintptr_t>(stage) & 15)) & 15;
    double2* sW = reinterpret_cast<double2*>(stage) + (size_t)warp * 64;
    int4* sR = reinterpret_cast<int4*>(stage + (size_t)n_warps * 32 * 32) + (size_t)warp * 32;
    double* sC = reinterpret_cast<double*>(stage + (size_t)n_warps * 32 * 48) + (size_t)warp * 32;
    int* idx0 = reinterpret_cast<int*>(stage + (size_t)n_warps * 32 * 56) + (size_t)(NP * warp) * pl.E;
    __syncthreads();

    const int64_t col0 = (int64_t)blockIdx.y * (64 * CPL2) + 2 * lane;
    const double* fcol = f + col0;
    bool okc[CPL2];
#pragma unroll
    for (int c = 0; c < CPL2; ++c) okc[c] = !GUARD || col0 + 64 * c + 1 < ncols;
    const int64_t n_quads = (P + NP - 1) / NP;
    const int64_t q_begin = (int64_t)blockIdx.x * quads_per_block;
    const int64_t q_end = min(n_quads, q_begin + quads_per_block);
    for (int64_t q = q_begin + warp; q < q_end; q += n_warps) {
        int64_t pt[NP];
        bool has[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            has[j] = NP * q + j < P;
            const int64_t s = has[j] ? NP * q + j : NP * q;
            pt[j] = perm ? (int64_t)perm[s] : s;
        }
        __syncwarp();
        for (int e = lane; e < pl.E; e += 32) {
            const Entry en = pl.ent[e];
#pragma unroll
            for (int j = 0; j < NP; ++j)
                fill_basis<KIND, int>(en, x[pt[j] * ldx + en.dim], sknots, slambda, bas0 + (size_t)j * pl.B,
                                      idx0 + (size_t)j * pl.E, e, 1, pl.flags);
        }
        __syncwarp();
        double2 acc[NP][CPL2], tp[NP][CPL2];
#pragma unroll
        for (int j = 0; j < NP; ++j)
#pragma unroll
            for (int c = 0; c < CPL2; ++c) {
                acc[j][c] = make_double2(0.0, 0.0);
                tp[j][c] = make_double2(0.0, 0.0);
            }
        for (int g0 = 0; g0 < pl.n_corners; g0 += 32) {
            const int g = g0 + lane;
            double w[NP];
            int row[NP];
            double coef = 0.0;
            int last = 0;
#pragma unroll
            for (int j = 0; j < NP; ++j) { w[j] = 1.0; row[j] = 0; }
            if (g < pl.n_corners) {
                const int t = pl.corner_term[g];
                const TermHdr h = pl.hdr[t];
                const int cid = g - pl.corner_off[t];
                const TermDim* td = pl.tdim + h.dim_off;
                int off[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) off[j] = 0;
                int cs = h.ncorner;
                int rem = cid;
                for (int jd = 0; jd < h.k; ++jd) {
                    const TermDim d = td[jd];
                    int a;
                    if (KIND == SGM_LAGRANGE) {
                        cs /= d.radix;
                        a = (cid / cs) % d.radix;
                    } else {
                        constexpr int R = fixed_radix(KIND);
                        cs /= R;
                        a = rem / cs;
                        rem -= a * cs;
                    }
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        w[j] = w[j] * digit_weight<KIND>(bas0 + (size_t)j * pl.B, 1, d, a);
                        off[j] += (idx0[(size_t)j * pl.E + d.ent] + a) * d.stride;
                    }
                }
#pragma unroll
                for (int j = 0; j < NP; ++j) row[j] = pl.maps[h.val_off + off[j]];
                last = (cid == h.ncorner - 1);
                if (last) coef = pl.coeff[t];
                if (KIND == SGM_HIER_PW_LINEAR) {
#pragma unroll
                    for (int j = 0; j < NP; ++j) w[j] = w[j] * coef;                   // folded into the weight
                }
            }
            __syncwarp();                                        // previous batch fully read
            sW[2 * lane] = make_double2(w[0], w[1]);
            sW[2 * lane + 1] = make_double2(w[2], w[3]);
            sR[lane] = make_int4(row[0] | (last << 31), row[1], row[2], row[3]);
            sC[lane] = coef;
            __syncwarp();
            const int n_in = min(32, pl.n_corners - g0);
            for (int cc = 0; cc < n_in; ++cc) {
                const double2 w01 = sW[2 * cc], w23 = sW[2 * cc + 1];
                const int4 rr = sR[cc];
                const int r0 = rr.x & 0x7fffffff;
                const int is_last = rr.x < 0;
                const int rj[NP] = {r0, rr.y, rr.z, rr.w};
                const double wj[NP] = {w01.x, w01.y, w23.x, w23.y};
                double2 v[CPL2];
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    if (j == 0 || rj[j] != rj[j - 1]) {           // warp-uniform
                        const double* fr = fcol + (int64_t)rj[j] * ldf;
#pragma unroll
                        for (int c = 0; c < CPL2; ++c)
                            v[c] = okc[c] ? __ldg(reinterpret_cast<const double2*>(fr + 64 * c)) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int c = 0; c < CPL2; ++c) {
                        if (KIND == SGM_HIER_PW_LINEAR) {         // one entry per term, coefficient 1
                            acc[j][c].x = fma(v[c].x, wj[j], acc[j][c].x);
                            acc[j][c].y = fma(v[c].y, wj[j], acc[j][c].y);
                        } else {
                            tp[j][c].x = tp[j][c].x + v[c].x * wj[j];
                            tp[j][c].y = tp[j][c].y + v[c].y * wj[j];
                        }
                    }
                }
                if (KIND != SGM_HIER_PW_LINEAR && is_last) {
                    const double cf = sC[cc];
#pragma unroll
                    for (int j = 0; j < NP; ++j)
#pragma unroll
                        for (int c = 0; c < CPL2; ++c) {
                            acc[j][c].x = acc[j][c].x + cf * tp[j][c].x;
                            acc[j][c].y = acc[j][c].y + cf * tp[j][c].y;
                            tp[j][c] = make_double2(0.0, 0.0);
                        }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            if (!has[j]) continue;
            double* o = out + pt[j] * ldo + col0;
#pragma unroll
            for (int c = 0; c < CPL2; ++c) {
                if (!okc[c]) continue;
                double2* dst = reinterpret_cast<double2*>(o + 64 * c);
                double2 r = make_double2(scale * acc[j][c].x, scale * acc[j][c].y);
                if (accumulate) {
                    const double2 old = *dst;
                    r.x = old.x + r.x;
                    r.y = old.y + r.y;
                }
                *dst = r;
            }
        }
    }
}

template <int KIND, int CPL2>
__global__ void __launch_bounds__(256, 2)
eval_cols_wide4_kernel(const PlanDev pl, const double* __restrict__ x, int64_t P, int64_t ldx,
                       const double* __restrict__ f, int64_t ldf, double* __restrict__ out,
                       int64_t ldo, double scale, int accumulate, int quads_per_block,
                       const int* __restrict__ perm, int64_t ncols) {
    if (((int64_t)blockIdx.y + 1) * (64 * CPL2) <= ncols)        // block-uniform
        cols_wide4_body<KIND, CPL2, false>(pl, x, P, ldx, f, ldf, out, ldo, scale, accumulate,
                                           quads_per_block, perm, ncols);
    else
        cols_wide4_body<KIND, CPL2, true>(pl, x, P, ldx, f, ldf, out, ldo, scale, accumulate,
                                          quads_per_block, perm, ncols);
}

// max |a - b| / max |b| with the NaN rule of sgm_max_rel_deviation (one block).
__global__ void __launch_bounds__(1024)
max_rel_deviation_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                         double* __restrict__ out) {
    __shared__ double s_diff[32], s_abs[32];
    __shared__ int s_bad[32];
    double md = 0.0, ma = 0.0;
    int bad = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = a[i], y = b[i];
        const bool nx = x != x, ny = y != y;
        if (nx && ny) continue;
        if (nx || ny) { bad = 1; continue; }
        md = fmax(md, fabs(x - y));
        ma = fmax(ma, fabs(y));
    }
    for (int off = 16; off > 0; off >>= 1) {
        md = fmax(md, __shfl_xor_sync(0xffffffffu, md, off));
        ma = fmax(ma, __shfl_xor_sync(0xffffffffu, ma, off));
        bad |= __shfl_xor_sync(0xffffffffu, bad, off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_diff[warp] = md; s_abs[warp] = ma; s_bad[warp] = bad; }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        md = lane < nw ? s_diff[lane] : 0.0;
        ma = lane < nw ? s_abs[lane] : 0.0;
        bad = lane < nw ? s_bad[lane] : 0;
        for (int off = 16; off > 0; off >>= 1) {
            md = fmax(md, __shfl_xor_sync(0xffffffffu, md, off));
            ma = fmax(ma, __shfl_xor_sync(0xffffffffu, ma, off));
            bad |= __shfl_xor_sync(0xffffffffu, bad, off);
        }
        if (lane == 0)
            out[0] = bad ? __longlong_as_double(0x7ff8000000000000ll) : md / (ma > 0.0 ? ma : 1.0);
    }
}

// One batch of the dimension-wise hierarchisation (surplus = value - coarser interpolant).
__global__ void hier_apply_kernel(double* __restrict__ alpha, int64_t NF, int64_t ld,
                                  const int* __restrict__ child, const int* __restrict__ pa,
                                  const int* __restrict__ pb, const double* __restrict__ wa,
                                  const double* __restrict__ wb, int64_t n) {
    const int64_t total = n * NF;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t it = i / NF, c = i - it * NF;
        double coarse = wa[it] * alpha[(int64_t)pa[it] * ld + c];
        if (pb[it] >= 0) coarse = coarse + wb[it] * alpha[(int64_t)pb[it] * ld + c];
        alpha[(int64_t)child[it] * ld + c] -= coarse;
    }
}

// All hierarchisation batches in ONE launch of ONE thread-block cluster (small tables: the
// batches are sequential, a barrier separates them; 64 launches of a few microseconds of work
// each cost more than the work).  8 CTAs x 1024 threads share every batch and meet at the
// hardware cluster barrier (arrive.release / wait.acquire at cluster scope orders the global
// writes of a batch before the reads of the next): config 2's 64 batches 0.33 -> 0.06 ms
// against one block with __syncthreads.
constexpr int HIER_CLUSTER = 8;
__global__ void __cluster_dims__(HIER_CLUSTER, 1, 1) __launch_bounds__(1024)
hier_apply_all_kernel(double* __restrict__ alpha, int64_t NF, int64_t ld,
                      const int* __restrict__ child, const int* __restrict__ pa,
                      const int* __restrict__ pb, const double* __restrict__ wa,
                      const double* __restrict__ wb, const int64_t* __restrict__ batch_off,
                      int n_batches) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    // The index / weight records of a thread's FIRST element of the next batch do not depend on
    // alpha: they are requested before the barrier, so that after it a batch costs one round trip
    // (the alpha values) instead of two.
    int c_ch = 0, c_pa = 0, c_pb = -1;
    double c_wa = 0.0, c_wb = 0.0;
    int64_t c_col = 0;
    auto fetch = [&](int b) {
        if (b >= n_batches) return;
        const int64_t lo = batch_off[b], n = batch_off[b + 1] - lo;
        if (tid < n * NF) {
            const int64_t it = lo + tid / NF;
            c_col = tid % NF;
            c_ch = child[it]; c_pa = pa[it]; c_pb = pb[it]; c_wa = wa[it]; c_wb = wb[it];
        }
    };
    fetch(0);
    for (int b = 0; b < n_batches; ++b) {
        const int64_t lo = batch_off[b], n = batch_off[b + 1] - lo;
        const int64_t total = n * NF;
        const bool mine = tid < total;
        const int ch0 = c_ch, pa0 = c_pa, pb0 = c_pb;
        const double wa0 = c_wa, wb0 = c_wb;
        const int64_t col0 = c_col;
        fetch(b + 1);
        if (mine) {
            double coarse = wa0 * alpha[(int64_t)pa0 * ld + col0];
            if (pb0 >= 0) coarse = coarse + wb0 * alpha[(int64_t)pb0 * ld + col0];
            alpha[(int64_t)ch0 * ld + col0] -= coarse;
        }
        for (int64_t i = tid + nthreads; i < total; i += nthreads) {
            const int64_t it = lo + i / NF, c = i % NF;
            double coarse = wa[it] * alpha[(int64_t)pa[it] * ld + c];
            if (pb[it] >= 0) coarse = coarse + wb[it] * alpha[(int64_t)pb[it] * ld + c];
            alpha[(int64_t)child[it] * ld + c] -= coarse;
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\n"
                     "barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}

// ---------------------------------------------------------------------------------------
// Brownian-motion expansions (parametric_expansions_brownian_motion.py:17-91)
// ---------------------------------------------------------------------------------------
// One hat per level contains s = t/T; all other hats of the level contribute an exact 0 in
// the reference's O(d) sum, so accumulating only that hat, level by level, gives the same
// bits.
__global__ void lc_brownian_kernel(const double* __restrict__ tt, int64_t nt,
                                   const double* __restrict__ yy, int64_t P, int d, int64_t ldy,
                                   double T, double sqrtT, const double* __restrict__ level_scale,
                                   int L, double* __restrict__ W, int64_t ldw) {
    const int64_t total = P * nt;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / nt, it = i - p * nt;
        const double* y = yy + p * ldy;
        const double s = tt[it] / T;
        double w = y[0] * s;
        if (s >= 0.0 && s <= 1.0) {
            for (int l = 1; l <= L; ++l) {
                const int64_t half = (int64_t)1 << (l - 1);          // hats on this level
                const double inv = 1.0 / (double)(half << 1);        // 2^-l, exact
                int64_t j = (int64_t)floor(s * (double)half);
                if (j > half - 1) j = half - 1;
                const int64_t k = half + j;                          // coefficient index
                if (k >= d) continue;                                // zero padding
                const double left = (double)(2 * j) * inv;
                const double mid = (double)(2 * j + 1) * inv;
                const double right = (double)(2 * j + 2) * inv;
                const double eta = (s >= mid) ? (-s + right) : (s - left);
                w = w + (y[k] * level_scale[l - 1]) * eta;
            }
        }
        W[p * ldw + it] = w * sqrtT;
    }
}

// Fast path (L <= LMAX): a block owns a chunk of 256*TPT consecutive times and a range of
// rows.  Each thread computes the per-level hat index and hat value of its TPT times ONCE
// (they do not depend on the parameter vector) and then streams over the rows: per output
// L x (load y_k, two multiplies, one add) and one store; a block writes 2*TPT KB of every row
// contiguously, which keeps the HBM write stream page-friendly.
constexpr int LC_ROWS = 32;
// DC: number of coefficients when known at compile time (64 = config 3), 0 = run-time `d_rt`:
// with a constant row length all staging index arithmetic folds away and the coefficient
// reads of the unrolled row loop address as per-thread register + immediate.
template <int LMAX, int TPT, int BPS, int DC = 0, bool BULK = false>
__global__ void __launch_bounds__(256, BPS)
lc_brownian_tiles_kernel(const double* __restrict__ tt, int64_t nt,
                         const double* __restrict__ yy, int64_t P, int d_rt, int64_t ldy, double T,
                         double sqrtT, const double* __restrict__ level_scale, int L,
                         double* __restrict__ W, int64_t ldw, int rows_per_block) {
    const int d = DC ? DC : d_rt;
    const int64_t t_base = (int64_t)blockIdx.x * (256 * TPT) + threadIdx.x;
    int kidx[TPT][LMAX];
    double eta[TPT][LMAX], sv[TPT];
#pragma unroll
    for (int q = 0; q < TPT; ++q) {
        const int64_t it = t_base + q * 256;
        const double s = it < nt ? tt[it] / T : 0.0;
        sv[q] = s;
        const bool inside = (s >= 0.0 && s <= 1.0);
#pragma unroll
        for (int l = 1; l <= LMAX; ++l) {
            // inactive levels (outside [0,1], zero padding) read y[0] with eta = 0: they add
            // an exact zero, which is also what the reference's sum over all hats does
            kidx[q][l - 1] = 0;
            eta[q][l - 1] = 0.0;
            if (l <= L && inside) {
                const int64_t half = (int64_t)1 << (l - 1);
                const double inv = 1.0 / (double)(half << 1);
                int64_t j = (int64_t)floor(s * (double)half);
                if (j > half - 1) j = half - 1;
                const int64_t k = half + j;
                if (k < d) {
                    const double left = (double)(2 * j) * inv;
                    const double mid = (double)(2 * j + 1) * inv;
                    const double right = (double)(2 * j + 2) * inv;
                    kidx[q][l - 1] = (int)k;
                    eta[q][l - 1] = (s >= mid) ? (-s + right) : (s - left);
                }
            }
        }
    }
    // Rows are staged 32 at a time in shared memory so that the per-level coefficient reads
    // are shared-memory broadcasts instead of L2 round trips.  The copy of group g + 1
    // (cp.async, no registers) is in flight while group g is evaluated, so the global-load
    // latency is off the critical path; a block synchronises once per group.  After its own
    // copies have landed a thread replaces them by y_k * 2^((l-1)/2), the product of
    // coefficient k with its level scale: it does not depend on the time, so it is formed
    // once per (row, k) instead of once per output (same operands, same rounding; 14 instead
    // of 20 FP64 operations per output).
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sy = reinterpret_cast<double*>(smem_raw);          // [2][LC_ROWS][d]
    const int64_t p0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t p1 = min(P, p0 + rows_per_block);
    const int st_r = (int)threadIdx.x / d, st_c = (int)threadIdx.x - st_r * d;
    const int step_r = 256 / d, step_c = 256 - step_r * d;
    const int group_elems = LC_ROWS * d;
    auto level_scale_of = [&](int c) -> double {               // 1 for y_0 (never scaled)
        const int lv = 31 - __clz(c | 1);                      // level - 1 of coefficient c >= 1
        return c == 0 ? 1.0 : (lv < L ? __ldg(level_scale + lv) : 0.0);
    };
    const double my_scale = level_scale_of(st_c);              // column is fixed if step_c == 0
    // BULK (contiguous rows, ldy == d): a row group is ONE contiguous chunk of the input, so one
    // elected thread moves it with a single bulk asynchronous copy (cp.async.bulk, the 1-D TMA
    // path; completion counted in bytes on an mbarrier per buffer) instead of 256 threads
    // issuing 8-byte cp.async each: no per-element address arithmetic, no LSU issue slots.
    const unsigned mbar0 = (unsigned)__cvta_generic_to_shared(sy + 2 * (size_t)group_elems);
    if (BULK) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0 + 8u) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    auto issue_copy = [&](int64_t pb, int buf) {
        const int nrows = (int)min((int64_t)LC_ROWS, p1 - pb);
        const unsigned dst0 = (unsigned)__cvta_generic_to_shared(sy + (size_t)buf * group_elems);
        if (BULK) {
            if (threadIdx.x == 0) {
                const unsigned bytes = (unsigned)(nrows * d) * 8u;
                const unsigned bar = mbar0 + 8u * (unsigned)buf;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst0), "l"(yy + pb * ldy), "r"(bytes), "r"(bar) : "memory");
            }
            return;
        }
        // (row, column) of element i = threadIdx.x + 256 n, advanced without a division
        for (int i = threadIdx.x, r = st_r, c = st_c; i < nrows * d; i += 256) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst0 + 8u * i),
                         "l"(yy + (pb + r) * ldy + c) : "memory");
            r += step_r;
            c += step_c;
            if (c >= d) { c -= d; ++r; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (p0 < p1) issue_copy(p0, 0);
    int g = 0;
    for (int64_t pb = p0; pb < p1; pb += LC_ROWS, ++g) {
        const int nrows = (int)min((int64_t)LC_ROWS, p1 - pb);
        double* buf = sy + (size_t)(g & 1) * group_elems;
        if (BULK) {
            const unsigned bar = mbar0 + 8u * (unsigned)(g & 1), parity = (unsigned)(g >> 1) & 1u;
            unsigned done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        for (int i = threadIdx.x, c = st_c; i < nrows * d; i += 256) {
            const double sc = step_c == 0 ? my_scale : level_scale_of(c);
            if (c != 0) buf[i] = buf[i] * sc;                  // entry 0 stays y_0
            c += step_c;
            if (c >= d) c -= d;
        }
        // BULK: order this thread's generic-proxy accesses to the buffers (reads of group g - 1,
        // the scaling writes above) before the asynchronous-proxy write of the next copy
        if (BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();     // group g visible to all; everybody is done with group g - 1
        if (pb + LC_ROWS < p1) issue_copy(pb + LC_ROWS, (g + 1) & 1);
        // row pointers advance by uniform increments (shared-memory reads address as
        // per-thread offset + uniform row base, stores as per-thread column + uniform row)
        const double* y = buf;
        double* wrow = W + pb * ldw + t_base;
#pragma unroll 4
        for (int r = 0; r < nrows; ++r, y += d, wrow += ldw) {
            const double y0 = y[0];
#pragma unroll
            for (int q = 0; q < TPT; ++q) {
                double w = y0 * sv[q];
#pragma unroll
                for (int l = 0; l < LMAX; ++l) w = w + y[kidx[q][l]] * eta[q][l];
                if (t_base + q * 256 < nt) wrow[q * 256] = w * sqrtT;
            }
        }
    }
}

__global__ void kl_brownian_kernel(const double* __restrict__ tt, int64_t nt,
                                   const double* __restrict__ yy, int64_t P, int d, int64_t ldy,
                                   const double* __restrict__ factor,
                                   const double* __restrict__ freq, double* __restrict__ W,
                                   int64_t ldw) {
    const int64_t total = P * nt;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / nt, it = i - p * nt;
        const double* y = yy + p * ldy;
        const double t = tt[it];
        double w = 0.0 * t;
        for (int n = 0; n < d; ++n) w = w + (factor[n] * sin(freq[n] * t)) * y[n];
        W[p * ldw + it] = w;
    }
}

// Tiled form: the factor  a_n(t) = factor_n * sin(freq_n * t)  does not depend on the
// parameter vector, so a block computes it ONCE for its 64 times (shared memory, [n][64])
// and then streams over its rows, which are staged 32 at a time.  Thread = (time, row lane):
// 64 times x 4 row lanes; a thread evaluates two rows at once, so one read of a_n(t) serves
// two outputs.  Per output: d x (product, sum) in the order of the reference
// (parametric_expansions_brownian_motion.py:88-91), d sines per block instead of per output.
constexpr int KL_TT = 64, KL_ROWS = 32;
__global__ void __launch_bounds__(256)
kl_brownian_tiles_kernel(const double* __restrict__ tt, int64_t nt, const double* __restrict__ yy,
                         int64_t P, int d, int64_t ldy, const double* __restrict__ factor,
                         const double* __restrict__ freq, double* __restrict__ W, int64_t ldw,
                         int rows_per_block) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);          // [d][KL_TT]
    double* sy = sA + (size_t)d * KL_TT;                       // [KL_ROWS][d]
    const int tl = threadIdx.x & (KL_TT - 1), rl = threadIdx.x >> 6;
    const int64_t t0 = (int64_t)blockIdx.x * KL_TT;
    const int64_t it = t0 + tl;
    const double t = it < nt ? tt[it] : 0.0;
    for (int n = rl; n < d; n += 4) sA[n * KL_TT + tl] = factor[n] * sin(freq[n] * t);
    const int64_t p0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t p1 = min(P, p0 + rows_per_block);
    const double w0 = 0.0 * t;
    for (int64_t pb = p0; pb < p1; pb += KL_ROWS) {
        const int nrows = (int)min((int64_t)KL_ROWS, p1 - pb);
        __syncthreads();                                         // sA ready / previous rows consumed
        for (int i = threadIdx.x; i < nrows * d; i += 256) {
            const int r = i / d, c = i - r * d;
            sy[i] = __ldg(yy + (pb + r) * ldy + c);
        }
        __syncthreads();
        for (int r = rl; r < nrows; r += 8) {                    // rows r and r + 4
            const bool two = r + 4 < nrows;
            const double* ya = sy + (size_t)r * d;
            const double* yb = sy + (size_t)(two ? r + 4 : r) * d;
            double wa = w0, wb = w0;
            for (int n = 0; n < d; ++n) {
                const double a = sA[n * KL_TT + tl];
                wa = wa + a * ya[n];
                wb = wb + a * yb[n];
            }
            if (it < nt) {
                W[(pb + r) * ldw + it] = wa;
                if (two) W[(pb + r + 4) * ldw + it] = wb;
            }
        }
    }
}

}  // namespace sgm
