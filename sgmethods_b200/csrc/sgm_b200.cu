// sgm_b200.cu -- C ABI of libsgm_b200.so (see include/sgm_b200.h): plan packing / upload,
// launch heuristics and the entry points that enqueue the kernels of sgm_kernels.cuh.
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "sgm_internal.h"
#include "sgm_jit.h"
#include "sgm_kernels.cuh"

namespace {
thread_local std::string g_last_error;
}

// sgm_hier_host.cpp
bool sgm_build_hier_groups(int64_t T, int N, const int32_t* term_fam, int n_fam,
                           const int* fam_extent, const int64_t* val_off,
                           std::vector<int>& groups, std::vector<int>& parents,
                           std::vector<int>& steps, std::vector<int>& hmap, int* root_slot,
                           int n_lists, std::vector<int>& split);

int sgm_fail(int code, const char* msg) {
    g_last_error = msg ? msg : "";
    return code;
}
int sgm_failf(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define SGM_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t err__ = (call);                                                        \
        if (err__ != cudaSuccess)                                                          \
            return sgm_failf(SGM_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(err__)); \
    } while (0)

using sgm::Entry;
using sgm::PlanDev;
using sgm::TermDim;
using sgm::TermHdr;
using sgm::TermRec;

// Development switches, read ONCE when a plan is created -- never on the
// launch path.  All default to "off" = the measured-best choice; none changes results.
struct Knobs {
    int force_kernel = 0;        // SGM_FORCE_KERNEL: 1 = v1, 2 = split, 3 = prefix
    int prefix_run = 8;          // SGM_PREFIX_RUN
    bool prefix_unbalanced = false;
    int prefix_ppl = 0;          // SGM_PREFIX_PPL: 0 = automatic, 1 | 2 forced
    bool no_sort = false, no_wide = false, no_idx8 = false, no_h1 = false, no_h2 = false, no_small = false, wide_bulk = false;
    int dev_flags = 0;           // SGM_DEV_FLAGS: PlanDev::dev_flags
    int key_mode = 0;            // SGM_KEY_MODE: 1 = one key bit per direction first (round-1 key)
    bool no_jit = false;         // SGM_NO_JIT: never build the plan-specialised small-plan kernel
    bool jit_sync = false;       // SGM_JIT_SYNC: build it inside the triggering call (default: background thread)
    long long jit_min_points = 4000000;  // SGM_JIT_MIN_POINTS: points seen by a plan before it is built
    int wide_pairs = -1;         // SGM_WIDE_PAIRS: -1 = automatic, 0 | 1
    int wide_pts = 0;            // SGM_WIDE_PTS: points per warp of the wide column kernel, 0 = automatic, 2 | 4
    long long col_loop_max = -1; // SGM_COL_LOOP_MAX
    int h2_shape = 0;            // SGM_H2_SHAPE: first launch shape of the grouped hierarchical kernel
    static Knobs from_env() {
        Knobs k;
        if (const char* v = std::getenv("SGM_FORCE_KERNEL"))
            k.force_kernel = !std::strcmp(v, "v1") ? 1 : !std::strcmp(v, "split") ? 2 : !std::strcmp(v, "prefix") ? 3 : 0;
        if (const char* v = std::getenv("SGM_PREFIX_RUN")) k.prefix_run = std::atoi(v);
        k.prefix_unbalanced = std::getenv("SGM_PREFIX_UNBALANCED") != nullptr;
        if (const char* v = std::getenv("SGM_PREFIX_PPL")) k.prefix_ppl = std::atoi(v);
        k.no_sort = std::getenv("SGM_NO_SORT") != nullptr;
        k.no_wide = std::getenv("SGM_NO_WIDE") != nullptr;
        k.no_idx8 = std::getenv("SGM_NO_IDX8") != nullptr;
        k.no_h1 = std::getenv("SGM_NO_H1") != nullptr;
        k.no_h2 = std::getenv("SGM_NO_H2") != nullptr;
        k.wide_bulk = std::getenv("SGM_WIDE_BULK") != nullptr;
        if (const char* v = std::getenv("SGM_DEV_FLAGS")) k.dev_flags = std::atoi(v);
        if (const char* v = std::getenv("SGM_KEY_MODE")) k.key_mode = std::atoi(v);
        k.no_small = std::getenv("SGM_NO_SMALL") != nullptr;
        k.no_jit = std::getenv("SGM_NO_JIT") != nullptr;
        k.jit_sync = std::getenv("SGM_JIT_SYNC") != nullptr;
        if (const char* v = std::getenv("SGM_JIT_MIN_POINTS")) k.jit_min_points = std::atoll(v);
        if (const char* v = std::getenv("SGM_WIDE_PAIRS")) k.wide_pairs = v[0] == '0' ? 0 : 1;
        if (const char* v = std::getenv("SGM_WIDE_PTS")) k.wide_pts = std::atoi(v);
        if (const char* v = std::getenv("SGM_COL_LOOP_MAX")) k.col_loop_max = std::atoll(v);
        if (const char* v = std::getenv("SGM_H2_SHAPE")) k.h2_shape = std::max(0, std::atoi(v)) % 6;
        return k;
    }
};

struct sgm_plan {
    int device = 0;
    PlanDev dev{};
    Knobs knobs{};
    int64_t M = 0;               // rows of f the node maps were validated against
    int n_seg = 0;               // > 0: segmented plan (sgm_eval_segments only)
    int64_t n_vals = 0;          // sum_t S_t
    int64_t corners = 0;         // sum_t C_t
    bool all_packed = false;     // every term has a packed record (needed by the H1 kernel)
    bool prefix_ok = false;      // tables of the prefix kernel (pw-linear, all packed) are built
    int first_level_keys = 0;    // directions in the locality key (before second-level bits)
    int prefix_wins = 0;         // occupancy comparison prefix vs split kernel (plan creation)
    bool prefix_launch_ok = false;
    bool h2_ok = false;          // tables of the grouped hierarchical kernel are built
    bool small_ok = false;       // kernel A0: few pw-linear terms, metadata in kernel parameters
    sgm::SmallPlan small{};
    const double* small_knots = nullptr;       // knots re-ordered for kernel A0 (device)
    const unsigned short* small_tabs = nullptr;
    // plan-specialised variant of kernel A0 (sgm_jit.cpp), built once the plan has seen
    // knobs.jit_min_points points: state 0 = not tried, 2 = being built (background thread),
    // 1 = ready, -1 = unavailable (jit_why)
    SgmJitSmallDesc small_desc;
    SgmJitKernel small_jit;
    std::atomic<int> jit_state{0};
    std::atomic<long long> small_points{0};
    std::mutex jit_mutex;        // guards jit_why and jit_thread
    std::thread jit_thread;
    std::string jit_why;
    int cc_major = 10, cc_minor = 0;
    int64_t table_bytes = 0;
    int kmax = 0;
    int max_entry_n = 0;         // largest knot family among the table entries
    int sm_count = 148;
    int max_smem_optin = 0;
    std::vector<void*> allocations;
};

namespace {

// Entry points run on the plan's device and leave the caller's current device untouched.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device)
            switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

template <class T>
int upload(sgm_plan* plan, const std::vector<T>& host, const T** dev_ptr) {
    void* d = nullptr;
    size_t bytes = std::max<size_t>(host.size() * sizeof(T), 16);
    SGM_CUDA(cudaMalloc(&d, bytes));
    plan->allocations.push_back(d);
    if (!host.empty())
        SGM_CUDA(cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    plan->table_bytes += (int64_t)(host.size() * sizeof(T));
    *dev_ptr = static_cast<const T*>(d);
    return SGM_OK;
}

int radix_of(int kind, int m) {
    switch (kind) {
        case SGM_HIER_PW_LINEAR: return 1;
        case SGM_PW_LINEAR: return 2;
        case SGM_PW_QUADRATIC: return 3;
        case SGM_PW_CUBIC: return 4;
        default: return m;
    }
}
int basis_width(int kind, int m) {
    switch (kind) {
        case SGM_HIER_PW_LINEAR: return 1;  // value of the one non-zero hierarchical hat
        case SGM_PW_LINEAR: return 1;       // y; the lower weight 1 - y is formed on use
        case SGM_PW_QUADRATIC: return 3;
        case SGM_PW_CUBIC: return 4;
        default: return m;
    }
}

// Launch shape of kernel A (one thread per point).
struct PointLaunch { int threads; size_t smem; bool global_basis; int blocks_per_sm; };

PointLaunch choose_point_launch(const sgm_plan* plan) {
    const PlanDev& d = plan->dev;
    const size_t fixed = (size_t)d.n_knots * 8 * (d.kind == SGM_LAGRANGE || d.kind == SGM_HIER_PW_LINEAR ? 2 : 1);
    const size_t per_point = (size_t)d.B * 8 + (size_t)d.E * 2;
    const size_t cap = (size_t)plan->max_smem_optin;
    const size_t per_sm = 228 * 1024;
    PointLaunch best{0, 0, true, 0};
    size_t best_threads = 0;
    for (int threads : {128, 64, 32}) {
        const size_t smem = fixed + per_point * threads;
        if (smem > cap) continue;
        const int bps = (int)std::min<size_t>(16, per_sm / (smem + 1024));
        if (bps < 1) continue;
        const size_t resident = (size_t)bps * threads;
        if (resident > best_threads) {
            best_threads = resident;
            best = PointLaunch{threads, smem, false, bps};
        }
    }
    if (best.threads == 0 || best_threads < 128) {
        // basis tables do not fit next to enough threads: keep them in a global scratch
        if (fixed <= cap) return PointLaunch{128, fixed, true, 4};
    }
    return best;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// bytes of the pre-gathered value table at the head of the workspace (the grouped hierarchical
// kernel has its own, padded layout)
size_t g_bytes_of(const sgm_plan* plan) {
    return align_up((size_t)std::max<int64_t>(plan->n_vals, plan->dev.h2_n_vals) * 8, 256);
}
bool use_h2_kernel(const sgm_plan* plan) {
    return plan->dev.kind == SGM_HIER_PW_LINEAR && plan->h2_ok && !plan->knobs.no_h2;
}
constexpr int64_t kSortMinPoints = 32768;       // scalar kernels
constexpr int64_t kSortMinPointsCols = 4096;    // wide column kernel with two points per warp
size_t sort_scratch_bytes(int64_t P) {      // perm | bins | block totals | keys | tile counter
    return align_up((size_t)P * 4, 256) + 65536 * 4 + 64 * 4 + align_up((size_t)P * 2, 256) + 256;
}
int* sort_tile_counter(unsigned char* scratch, int64_t P) {
    return reinterpret_cast<int*>(scratch + sort_scratch_bytes(P) - 256);
}

// Launch shape of kernel A2 (32 points x W warps per block, terms split over the warps).
struct SplitLaunch { int warps; int chunk; size_t smem; int blocks_per_sm; bool idx8; };

using SplitKernel = void (*)(const PlanDev, const double*, int64_t, int64_t, const double*,
                             double*, int64_t, double, int, const int*);

template <int KIND>
SplitKernel split_kernel_for(int warps, int chunk, bool idx8 = false) {
    if constexpr (KIND == SGM_PW_LINEAR) {
        // byte bracket indices + 48-term chunks: two 16-warp blocks per SM where 16-bit indices
        // and 64-term chunks allow a single block (config 5: 288 table entries)
        if (idx8 && warps == 16 && chunk == 48) return sgm::eval_points_split_kernel<KIND, 16, 48, true>;
        if (idx8) return nullptr;
        // 1024-thread blocks (64 registers): for plans whose bracket tables leave room for one
        // block per SM only (128 directions, 92 KB of tables per 32 points)
        if (warps == 32 && chunk == 128) return sgm::eval_points_split_kernel<KIND, 32, 128>;
    }
    if (idx8) return nullptr;
    if (warps == 16 && chunk == 64) return sgm::eval_points_split_kernel<KIND, 16, 64>;
    if (warps == 8 && chunk == 32) return sgm::eval_points_split_kernel<KIND, 8, 32>;
    if (warps == 8 && chunk == 64) return sgm::eval_points_split_kernel<KIND, 8, 64>;
    return sgm::eval_points_split_kernel<KIND, 4, 32>;
}

// few_tiles: the batch has at most one tile per SM, so resident blocks per SM do not matter and
// the time is the latency of ONE tile's walk over all terms: most warps per block wins (the
// per-call validation sample of method="auto": config 2, 256 points: 0.83 -> ~0.25 ms).
template <int KIND>
SplitLaunch choose_split_launch(const sgm_plan* plan, bool few_tiles = false) {
    const PlanDev& d = plan->dev;
    const size_t fixed_no_idx = (size_t)d.n_knots * 8 * (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR ? 2 : 1) +
                                (size_t)d.B * 32 * 8 + 32 * 8 + 64;
    SplitLaunch best{0, 0, 0, 0, false};
    int best_warps = 0;
    // {warps, chunk, byte indices}; at equal resident warps the candidate with more blocks per
    // SM wins (independent blocks overlap each other's chunk barriers)
    const int cand[6][3] = {{16, 64, 0}, {8, 32, 0}, {8, 64, 0}, {4, 32, 0}, {32, 128, 0}, {16, 48, 1}};
    const bool idx8_ok = KIND == SGM_PW_LINEAR && plan->all_packed && plan->max_entry_n <= 257 &&
                         !plan->knobs.no_idx8;
    for (const auto& c : cand) {
        if (c[0] == 32 && KIND != SGM_PW_LINEAR) continue;
        if (c[2] && !idx8_ok) continue;
        const size_t smem = fixed_no_idx + (size_t)d.E * 32 * (c[2] ? 1 : 2) + 2 * (size_t)c[1] * 32 * 8;
        if (smem > (size_t)plan->max_smem_optin) continue;
        SplitKernel kern = split_kernel_for<KIND>(c[0], c[1], c[2] != 0);
        if (!kern) continue;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            continue;
        int bps = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, c[0] * 32, smem) != cudaSuccess)
            continue;
        const int score = few_tiles ? (bps > 0 ? c[0] : 0) : bps * c[0];
        if (score > best_warps || (score == best_warps && bps > best.blocks_per_sm)) {
            best_warps = score;
            best = SplitLaunch{c[0], c[1], smem, bps, c[2] != 0};
        }
    }
    return best;
}

// Locality sort of the points (scratch: perm P ints | 65536 bins | 64 block totals | keys P u16); *perm stays
// null when the batch is too small to benefit.
int locality_sort(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, unsigned char* scratch,
                  size_t scratch_bytes, cudaStream_t stream, const int** perm,
                  int64_t min_points = kSortMinPoints) {
    *perm = nullptr;
    if (!(P >= min_points && P < 0x7fffffff && plan->dev.n_key >= 4 && !plan->knobs.no_sort)) return SGM_OK;
    const size_t need = sort_scratch_bytes(P);
    if (scratch_bytes < need) return sgm_fail(SGM_ERR_WORKSPACE, "workspace too small for the point sort");
    int* perm_w = reinterpret_cast<int*>(scratch);
    unsigned* hist = reinterpret_cast<unsigned*>(scratch + align_up((size_t)P * 4, 256));
    unsigned* block_total = hist + 65536;
    unsigned short* keys = reinterpret_cast<unsigned short*>(
        scratch + align_up((size_t)P * 4, 256) + 65536 * 4 + 64 * 4);
    const int nbins = 1 << plan->dev.n_key;
    const int scan_blocks = (nbins + 1023) / 1024;
    SGM_CUDA(cudaMemsetAsync(hist, 0, (size_t)nbins * 4, stream));
    const int sblocks = (int)std::min<int64_t>((P + 255) / 256, (int64_t)plan->sm_count * 8);
    sgm::sort_hist_kernel<<<sblocks, 256, 0, stream>>>(plan->dev, x, P, ldx, hist, keys, sort_tile_counter(scratch, P));
    sgm::sort_scan_kernel<<<scan_blocks, 1024, 0, stream>>>(hist, nbins, block_total);
    sgm::sort_scatter_kernel<<<sblocks, 256, 0, stream>>>(keys, P, hist, block_total, scan_blocks, perm_w);
    SGM_CUDA(cudaGetLastError());
    *perm = perm_w;
    return SGM_OK;
}

// Launch shape of the prefix kernel (pw-linear, <= 4 active directions per term).
struct PrefixLaunch { int warps; int run; size_t smem; int blocks_per_sm; SplitKernel kern; int ppl; };

// ppl = points per lane (1 or 2).  Two points per lane need 128 registers and twice the shared
// memory per block: taken only if two such blocks still fit on an SM.
PrefixLaunch choose_prefix_launch(const sgm_plan* plan, int ppl = 1) {
    const PlanDev& d = plan->dev;
    PrefixLaunch none{0, 0, 0, 0, nullptr, 1};
    if (!plan->prefix_ok) return none;
    const int want = plan->knobs.prefix_run;                   // mean terms per warp run
    const int warps = 8;
    const bool balanced = !plan->knobs.prefix_unbalanced;
    const int cand[2] = {want, 4};
    for (int run : cand) {
        SplitKernel kern = nullptr;
        if (run == 8 && ppl == 2) kern = sgm::eval_points_prefix_kernel<8, 8, 2, true, 2>;
        else if (run == 8 && ppl == 1)
            kern = balanced ? sgm::eval_points_prefix_kernel<8, 8, 3, true> : sgm::eval_points_prefix_kernel<8, 8, 3, false>;
        else if (run == 4 && ppl == 1) kern = sgm::eval_points_prefix_kernel<8, 4, 3, false>;
        if (!kern) continue;
        const int chunk = warps * run;
        const size_t tpts = 32 * (size_t)ppl;
        const size_t smem = (size_t)chunk * 64 + (size_t)d.n_knots * 8 + (size_t)d.B * tpts * 8 +
                            2 * (size_t)chunk * tpts * 8 + tpts * 8 + (size_t)d.E * tpts * 2 + 64;
        if (smem > (size_t)plan->max_smem_optin) continue;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            continue;
        int bps = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, warps * 32, smem) != cudaSuccess || bps < 1)
            continue;
        if (ppl == 2 && bps < 2) continue;
        return PrefixLaunch{warps, run, smem, bps, kern, ppl};
    }
    return none;
}

// true when the scalar evaluation of this plan / batch goes through the prefix kernel (the
// value tables are then gathered in its first-direction-fastest layout)
bool use_prefix_kernel(const sgm_plan* plan, int64_t P) {
    if (!plan->prefix_ok || plan->dev.kind != SGM_PW_LINEAR) return false;
    if (plan->knobs.force_kernel) return plan->knobs.force_kernel == 3 && plan->prefix_launch_ok;
    if (plan->dev.T < 32 && P >= 16384) return false;         // few terms, many points: kernel A
    // at most one tile per SM: the time is ONE tile's latency over all terms, and the split kernel
    // puts 16-32 warps on a tile where the prefix kernel has 8 (validation sample of
    // method="auto", 256 points of config 2: 0.24 ms against 0.56 ms)
    if ((P + 31) / 32 <= plan->sm_count) return false;
    return plan->prefix_wins > 0;                             // decided at plan creation
}

// Large bracket tables (many directions x levels) limit the resident blocks: the split
// kernel's 16-warp blocks then keep more warps on the SM, which matters more than the prefix
// reuse (config 5, d=128: 8 vs 16 resident warps, 110 vs 85 ms per 10^5 points).  Called once,
// at the end of plan creation (the plan is read-only afterwards).
void decide_prefix(sgm_plan* plan) {
    plan->prefix_wins = 0;
    plan->prefix_launch_ok = false;
    if (!plan->prefix_ok || plan->dev.kind != SGM_PW_LINEAR) return;
    const PrefixLaunch L = choose_prefix_launch(plan);
    const SplitLaunch S = choose_split_launch<SGM_PW_LINEAR>(plan);
    plan->prefix_launch_ok = L.warps > 0;
    // measured: config 2 (64 table entries; prefix 3 x 8 warps, split 5 x 8 at 47 registers)
    // prefix 25 ms, split 34 ms per 10^6 points; config 5 (288 entries; prefix 1 x 8 warps, split
    // 2 x 16) prefix 110 ms, split 53 ms per 10^5.  The prefix kernel needs two resident blocks
    // and at least half the split kernel's resident warps.
    plan->prefix_wins = L.warps > 0 && L.blocks_per_sm >= 2 &&
                        2 * L.warps * L.blocks_per_sm >= S.warps * S.blocks_per_sm;
}

// Kernel A0: few packed pw-linear terms, many points (gathers its node values itself).
bool use_small_kernel(const sgm_plan* plan, int64_t P) {
    return plan->small_ok && !plan->knobs.no_small && !plan->knobs.force_kernel && P >= 16384;
}
int launch_points_small(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                        int64_t ldf, double* out, int64_t ldo, double scale, int accumulate,
                        cudaStream_t stream) {
    constexpr int TH = 128, PPT = 2;
    const sgm::SmallPlan& sp = plan->small;
    const size_t smem = (size_t)sp.n_knots * 8 + (size_t)sp.E * TH * PPT * 10 + (size_t)plan->n_vals * 8 +
                        (size_t)sp.n_tab * 2 + 16;
    auto kern = sgm::eval_points_small_kernel<TH, PPT>;
    SGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 0;
    SGM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, TH, smem));
    if (bps < 1) return sgm_fail(SGM_ERR_OVERFLOW, "small-plan kernel does not fit");
    const int64_t nb = std::min<int64_t>((P + TH * PPT - 1) / (TH * PPT), (int64_t)plan->sm_count * bps * 2);
    kern<<<(unsigned)nb, TH, smem, stream>>>(sp, plan->small_knots, plan->small_tabs, x, P, ldx, plan->dev.maps,
                                            (int)plan->n_vals, f, ldf, out, ldo, scale, accumulate);
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

// Plan-specialised variant of kernel A0 (sgm_jit.cpp).  Built once per plan, when the plan has
// been asked for knobs.jit_min_points points in total, by a BACKGROUND thread (NVRTC takes
// ~0.2 s, 2 s with a cold file cache): evaluations keep using the generic kernel A0 until the
// specialised one is loaded, so no call ever waits for the compiler.  Returns true when the
// specialised kernel is ready.
constexpr int kJitThreads = 256, kJitPointsPerThread = 1;
void small_jit_build(sgm_plan* plan) {
    cudaSetDevice(plan->device);
    cudaFree(nullptr);                                   // the runtime's primary context is current
    const std::string src = sgm_jit_small_source(plan->small_desc, kJitThreads, kJitPointsPerThread, "sgm_small_jit");
    std::string why;
    const bool ok = sgm_jit_build(src, "sgm_small_jit", plan->cc_major, plan->cc_minor, kJitThreads,
                                  &plan->small_jit, &why);
    plan->small_jit.ppt = kJitPointsPerThread;
    {
        std::lock_guard<std::mutex> lock(plan->jit_mutex);
        plan->jit_why = ok ? "ready: " + std::to_string(plan->small_jit.regs) + " registers, " +
                                 std::to_string(plan->small_jit.blocks_per_sm) + " blocks per SM"
                           : why;
    }
    plan->jit_state.store(ok ? 1 : -1, std::memory_order_release);
}
bool small_jit_ready(sgm_plan* plan, int64_t P) {
    if (plan->knobs.no_jit) return false;
    const int state = plan->jit_state.load(std::memory_order_acquire);
    if (state != 0) return state == 1;
    if (plan->small_points.fetch_add(P) + P < plan->knobs.jit_min_points) return false;
    int expected = 0;
    if (!plan->jit_state.compare_exchange_strong(expected, 2)) return expected == 1;
    if (plan->knobs.jit_sync) {
        DeviceGuard guard(plan->device);
        small_jit_build(plan);
        return plan->jit_state.load(std::memory_order_acquire) == 1;
    }
    std::lock_guard<std::mutex> lock(plan->jit_mutex);
    plan->jit_thread = std::thread(small_jit_build, plan);
    return false;
}
void small_jit_join(sgm_plan* plan) {
    std::thread t;
    {
        std::lock_guard<std::mutex> lock(plan->jit_mutex);
        t.swap(plan->jit_thread);
    }
    if (t.joinable()) t.join();
}

int launch_points_small_jit(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                            int64_t ldf, double* out, int64_t ldo, double scale, int accumulate,
                            cudaStream_t stream) {
    const SgmJitKernel& k = plan->small_jit;
    const int64_t per_tile = (int64_t)k.threads * k.ppt;
    // three waves of resident blocks: enough tiles per block for the software pipeline, enough
    // blocks for an even tail
    const int64_t nb = std::min<int64_t>((P + per_tile - 1) / per_tile, (int64_t)plan->sm_count * k.blocks_per_sm * 3);
    long long Pll = P, ldxll = ldx, ldfll = ldf, ldoll = ldo;
    const int* maps = plan->dev.maps;
    void* args[] = {(void*)&plan->small_knots, (void*)&plan->small_tabs, (void*)&x, &Pll, &ldxll, (void*)&maps,
                    (void*)&f, &ldfll, (void*)&out, &ldoll, &scale, &accumulate};
    const int rc = sgm_jit_launch(k, (unsigned)nb, args, stream);
    if (rc != 0) return sgm_failf(SGM_ERR_CUDA, "launch of the plan-specialised kernel failed (CUresult %d)", rc);
    return SGM_OK;
}

int launch_points_prefix(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* G,
                         double* out, int64_t ldo, double scale, int accumulate,
                         unsigned char* scratch, size_t scratch_bytes, cudaStream_t stream) {
    PrefixLaunch L = choose_prefix_launch(plan);
    if (L.warps == 0) return sgm_fail(SGM_ERR_INVALID, "prefix kernel not available for this plan");
    {   // two points per lane once the batch fills every SM with such tiles (config 2: 27.7 ->
        // 25.0 ms); smaller batches keep 32-point tiles for the parallelism
        const int force_ppl = plan->knobs.prefix_ppl;
        const PrefixLaunch L2 = choose_prefix_launch(plan, 2);
        const bool fills = P >= (int64_t)64 * plan->sm_count * std::max(L2.blocks_per_sm, 1);
        if (L2.warps > 0 && (force_ppl ? force_ppl == 2 : fills)) L = L2;
    }
    const int64_t tiles = (P + 32 * L.ppl - 1) / (32 * L.ppl);
    const int64_t blocks = std::min<int64_t>(tiles, (int64_t)plan->sm_count * L.blocks_per_sm);
    const int* perm = nullptr;
    int rc = locality_sort(plan, x, P, ldx, scratch, scratch_bytes, stream, &perm);
    if (rc) return rc;
    PlanDev pd = plan->dev;                                   // sorted batches: dynamic tiles
    pd.tile_counter = (perm && !(plan->knobs.dev_flags & 8)) ? sort_tile_counter(scratch, P) : nullptr;
    L.kern<<<(unsigned)blocks, L.warps * 32, L.smem, stream>>>(pd, x, P, ldx, G, out, ldo,
                                                              scale, accumulate, perm);
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

template <int KIND>
int launch_points(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* G,
                  double* out, int64_t ldo, double scale, int accumulate, unsigned char* scratch,
                  size_t scratch_bytes, cudaStream_t stream);

template <int KIND>
int launch_points_split(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* G,
                        double* out, int64_t ldo, double scale, int accumulate,
                        unsigned char* scratch, size_t scratch_bytes, cudaStream_t stream) {
    const SplitLaunch L = choose_split_launch<KIND>(plan, (P + 31) / 32 <= plan->sm_count);
    // few terms, many points: one thread per point -- except for hierarchical plans whose value
    // table was gathered in the grouped kernel's own layout (eval_one pregathers through h2_maps
    // whenever the plan has group tables; found by tools/fuzz_auto.py: plans with < 32 tuples at
    // >= 16384 points read that table with term offsets)
    const bool h2_layout = KIND == SGM_HIER_PW_LINEAR && use_h2_kernel(plan);
    const bool want_v1 = !h2_layout && (plan->knobs.force_kernel ? plan->knobs.force_kernel == 1
                                                                 : (plan->dev.T < 32 && P >= 16384));
    if (!h2_layout && (want_v1 || L.warps == 0 || L.blocks_per_sm * L.warps < 8))   // few terms / huge tables
        return launch_points<KIND>(plan, x, P, ldx, G, out, ldo, scale, accumulate, scratch,
                                   scratch_bytes, stream);
    const int64_t tiles = (P + 31) / 32;
    int64_t blocks = std::min<int64_t>(tiles, (int64_t)plan->sm_count * L.blocks_per_sm);
    const int* perm = nullptr;
    int* unsorted_counter = nullptr;     // grouped kernel: tile counter of an unsorted batch
    // The grouped kernel does not sort: its value slots are size-aligned, so a gather touches one
    // line whatever the lanes' brackets are, and the sort costs more than it saves (config 2
    // 4.23 -> 4.08 ms, clustered points 5.54 -> 4.06 ms; tools/exp_nosort.py).  It still hands out
    // its tiles from a global counter (dev switch 16 = sort anyway).
    if (KIND == SGM_HIER_PW_LINEAR && use_h2_kernel(plan) && !(plan->knobs.dev_flags & 16) &&
        P >= kSortMinPoints && scratch_bytes >= sort_scratch_bytes(P)) {
        unsorted_counter = sort_tile_counter(scratch, P);
        SGM_CUDA(cudaMemsetAsync(unsorted_counter, 0, 4, stream));
    } else {
        int rc = locality_sort(plan, x, P, ldx, scratch, scratch_bytes, stream, &perm);
        if (rc) return rc;
    }
    if (KIND == SGM_HIER_PW_LINEAR && use_h2_kernel(plan)) {
        // launch shapes of the grouped kernel, best first (dev knob SGM_H2_SHAPE picks one)
        struct Shape { SplitKernel kern; int warps; };
        constexpr int NS = 6;
        const Shape shapes[NS] = {{sgm::eval_hier_groups_kernel<16, 2, 2, false>, 16},
                                  {sgm::eval_hier_groups_kernel<16, 2, 2, true>, 16},
                                  {sgm::eval_hier_groups_kernel<8, 4, 2, false>, 8},
                                  {sgm::eval_hier_groups_kernel<8, 3, 2, false>, 8},
                                  {sgm::eval_hier_groups_kernel<16, 1, 2, false>, 16},
                                  {sgm::eval_hier_groups_kernel<16, 2, 4, false>, 16}};
        const PlanDev& d = plan->dev;
        for (int si = 0; si < NS; ++si) {
            const Shape& sh = shapes[(si + plan->knobs.h2_shape) % NS];
            const size_t smem = (size_t)d.n_knots * 16 + (size_t)(d.B + 1) * 32 * 8 +
                                (size_t)sgm::H2_LISTS * (32 * 8 + 4) + (size_t)(d.E + 1) * 32 * 2 + 64 + 64;
            if (smem > (size_t)plan->max_smem_optin ||
                cudaFuncSetAttribute(sh.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                continue;
            int bps = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, sh.kern, sh.warps * 32, smem);
            if (bps < 1) continue;
            blocks = std::min<int64_t>(tiles, (int64_t)plan->sm_count * bps);
            // sorted batches: tiles are handed out from a global counter (last word of the sort
            // scratch, zeroed by the sort's first kernel), so that the blocks of an
            // SM that was busy when the kernel started take fewer tiles (dev knob SGM_DEV_FLAGS
            // bit 3 = static striding)
            PlanDev pd = plan->dev;
            pd.tile_counter = nullptr;
            if (perm && !(plan->knobs.dev_flags & 8)) pd.tile_counter = sort_tile_counter(scratch, P);   // zeroed by the sort
            if (unsorted_counter) pd.tile_counter = unsorted_counter;
            sh.kern<<<(unsigned)blocks, sh.warps * 32, smem, stream>>>(pd, x, P, ldx, G, out, ldo,
                                                                     scale, accumulate, perm);
            SGM_CUDA(cudaGetLastError());
            return SGM_OK;
        }
        // G is in the grouped kernel's layout: no other kernel may read it
        return sgm_fail(SGM_ERR_OVERFLOW, "eval: the grouped hierarchical kernel does not fit this device");
    }
    if (KIND == SGM_HIER_PW_LINEAR && plan->all_packed && !plan->knobs.no_h1) {
        constexpr int HW = 16;
        const PlanDev& d = plan->dev;
        const size_t smem = (size_t)d.n_knots * 16 + (size_t)d.B * 32 * 8 + (size_t)HW * 32 * 8 +
                            (size_t)d.E * 32 * 2 + 64;
        auto hk = sgm::eval_hier_points_kernel<HW>;
        if (smem <= (size_t)plan->max_smem_optin &&
            cudaFuncSetAttribute(hk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
            int bps = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, hk, HW * 32, smem);
            if (bps >= 1) {
                blocks = std::min<int64_t>(tiles, (int64_t)plan->sm_count * bps);
                hk<<<(unsigned)blocks, HW * 32, smem, stream>>>(plan->dev, x, P, ldx, G, out, ldo, scale,
                                                            accumulate, perm);
                SGM_CUDA(cudaGetLastError());
                return SGM_OK;
            }
        }
    }
    SplitKernel kern = split_kernel_for<KIND>(L.warps, L.chunk, L.idx8);
    SGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    PlanDev pd = plan->dev;                                   // sorted batches: dynamic tiles
    pd.tile_counter = (perm && !(plan->knobs.dev_flags & 8)) ? sort_tile_counter(scratch, P) : nullptr;
    kern<<<(unsigned)blocks, L.warps * 32, L.smem, stream>>>(pd, x, P, ldx, G, out, ldo,
                                                            scale, accumulate, perm);
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

template <int KIND>
int launch_points(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* G,
                  double* out, int64_t ldo, double scale, int accumulate, unsigned char* scratch,
                  size_t scratch_bytes, cudaStream_t stream) {
    const PointLaunch L = choose_point_launch(plan);
    if (L.threads == 0) return sgm_fail(SGM_ERR_OVERFLOW, "knot tables exceed shared memory");
    const int64_t tiles = (P + L.threads - 1) / L.threads;
    int64_t blocks = std::min<int64_t>(tiles, (int64_t)plan->sm_count * L.blocks_per_sm * 8);
    if (L.global_basis) {
        blocks = std::min<int64_t>(tiles, (int64_t)plan->sm_count * L.blocks_per_sm);
        const size_t need = (size_t)blocks * L.threads * ((size_t)plan->dev.B * 8 + (size_t)plan->dev.E * 2);
        if (need > scratch_bytes) return sgm_fail(SGM_ERR_WORKSPACE, "workspace too small for basis scratch");
        auto kern = sgm::eval_points_kernel<KIND, true>;
        SGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
        kern<<<(unsigned)blocks, L.threads, L.smem, stream>>>(plan->dev, x, P, ldx, G, out, ldo,
                                                              scale, accumulate, scratch);
    } else {
        auto kern = sgm::eval_points_kernel<KIND, false>;
        SGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
        kern<<<(unsigned)blocks, L.threads, L.smem, stream>>>(plan->dev, x, P, ldx, G, out, ldo,
                                                              scale, accumulate, nullptr);
    }
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

// Points per warp of the one-point-per-warp column kernels: 8 for large batches; small batches
// take fewer so that the grid still covers every SM a few times (a 256-point validation sample
// with 10^3 columns ran on 12 blocks: 11.7 ms; per-point arithmetic and results are unchanged).
static int cols_points_per_warp(const sgm_plan* plan, int64_t P, int64_t col_tiles, int warps) {
    const int64_t target_blocks = (int64_t)plan->sm_count * 4;
    int ppw = 8;
    while (ppw > 1 && ((P + warps * ppw - 1) / (warps * ppw)) * col_tiles < target_blocks) ppw >>= 1;
    return ppw;
}

template <int KIND, int CPL>
int launch_cols_cpl(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                    int64_t ldf, int64_t NF, double* out, int64_t ldo, double scale,
                    int accumulate, cudaStream_t stream) {
    const PlanDev& d = plan->dev;
    const size_t fixed = (size_t)d.n_knots * 8 * (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR ? 2 : 1);
    const size_t per_warp = (size_t)d.B * 8 + (size_t)d.E * 4;
    int warps = 8;
    while (warps > 1 && fixed + per_warp * warps > (size_t)plan->max_smem_optin) warps >>= 1;
    const size_t smem = fixed + per_warp * warps;
    if (smem > (size_t)plan->max_smem_optin)
        return sgm_fail(SGM_ERR_OVERFLOW, "per-point basis table exceeds shared memory");
    const int64_t gy = (NF + 32 * CPL - 1) / (32 * CPL);
    const int pts_per_block = warps * cols_points_per_warp(plan, P, gy, warps);
    const int64_t gx = (P + pts_per_block - 1) / pts_per_block;
    if (gx > 0x7fffffff || gy > 65535) return sgm_fail(SGM_ERR_OVERFLOW, "grid too large");
    auto kern = sgm::eval_cols_kernel<KIND, CPL>;
    SGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)gx, (unsigned)gy), warps * 32, smem, stream>>>(
        plan->dev, x, P, ldx, f, ldf, NF, out, ldo, scale, accumulate, pts_per_block);
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

template <int KIND, int CPL2>
int launch_cols_wide(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                     int64_t ldf, int64_t n_tiles, double* out, int64_t ldo, double scale,
                     int accumulate, cudaStream_t stream) {
    const PlanDev& d = plan->dev;
    const size_t fixed = (size_t)d.n_knots * 8 * (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR ? 2 : 1);
    const size_t per_warp = (size_t)d.B * 8 + (size_t)d.E * 4;
    int warps = 8;
    while (warps > 1 && fixed + per_warp * warps > (size_t)plan->max_smem_optin) warps >>= 1;
    const size_t smem = fixed + per_warp * warps;
    if (smem > (size_t)plan->max_smem_optin)
        return sgm_fail(SGM_ERR_OVERFLOW, "per-point basis table exceeds shared memory");
    const int pts_per_block = warps * cols_points_per_warp(plan, P, n_tiles, warps);
    const int64_t gx = (P + pts_per_block - 1) / pts_per_block;
    if (gx > 0x7fffffff || n_tiles > 65535) return sgm_fail(SGM_ERR_OVERFLOW, "grid too large");
    auto kern = sgm::eval_cols_wide_kernel<KIND, CPL2>;
    SGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)gx, (unsigned)n_tiles), warps * 32, smem, stream>>>(
        plan->dev, x, P, ldx, f, ldf, out, ldo, scale, accumulate, pts_per_block);
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

// Two points per warp (kernel C3): needs the locality permutation to pair points that read
// the same rows of f.  Returns SGM_OK with *launched = false when the shape does not fit.
template <int KIND, int CPL2>
int launch_cols_wide2(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                      int64_t ldf, int64_t n_tiles, int64_t ncols, double* out, int64_t ldo,
                      double scale, int accumulate, const int* perm, cudaStream_t stream,
                      bool* launched) {
    *launched = false;
    const PlanDev& d = plan->dev;
    const size_t fixed = (size_t)d.n_knots * 8 * (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR ? 2 : 1);
    const size_t per_warp = 2 * ((size_t)d.B * 8 + (size_t)d.E * 4) + 32 * (16 + 8 + 8);   // + entry staging
    const int warps = 8;
    const size_t smem = fixed + per_warp * warps + 16;              // + alignment of the staging area
    if (smem > (size_t)plan->max_smem_optin / 2) return SGM_OK;    // keep two blocks per SM
    const int pairs_per_block = warps * 4;
    const int64_t gx = ((P + 1) / 2 + pairs_per_block - 1) / pairs_per_block;
    if (gx > 0x7fffffff || n_tiles > 65535) return SGM_OK;
    auto kern = sgm::eval_cols_wide2_kernel<KIND, CPL2>;
    SGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)gx, (unsigned)n_tiles), warps * 32, smem, stream>>>(
        plan->dev, x, P, ldx, f, ldf, out, ldo, scale, accumulate, pairs_per_block, perm, ncols);
    SGM_CUDA(cudaGetLastError());
    *launched = true;
    return SGM_OK;
}

// Four points per warp (kernel C4) on 64 * CPL2-column tiles; same contract as launch_cols_wide2.
template <int KIND, int CPL2 = 2>
int launch_cols_wide4(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                      int64_t ldf, int64_t ncols, double* out, int64_t ldo, double scale,
                      int accumulate, const int* perm, cudaStream_t stream, bool* launched) {
    *launched = false;
    const PlanDev& d = plan->dev;
    const size_t fixed = (size_t)d.n_knots * 8 * (KIND == SGM_LAGRANGE || KIND == SGM_HIER_PW_LINEAR ? 2 : 1);
    const size_t per_warp = 4 * ((size_t)d.B * 8 + (size_t)d.E * 4) + 32 * (32 + 16 + 8);
    const int warps = 8;
    const size_t smem = fixed + per_warp * warps + 16;
    if (smem > (size_t)plan->max_smem_optin / 2) return SGM_OK;    // keep two blocks per SM
    const int quads_per_block = warps * 4;
    const int64_t n_tiles = (ncols + 64 * CPL2 - 1) / (64 * CPL2);
    const int64_t gx = ((P + 3) / 4 + quads_per_block - 1) / quads_per_block;
    if (gx > 0x7fffffff || n_tiles > 65535) return SGM_OK;
    auto kern = sgm::eval_cols_wide4_kernel<KIND, CPL2>;
    SGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)gx, (unsigned)n_tiles), warps * 32, smem, stream>>>(
        plan->dev, x, P, ldx, f, ldf, out, ldo, scale, accumulate, quads_per_block, perm, ncols);
    SGM_CUDA(cudaGetLastError());
    *launched = true;
    return SGM_OK;
}

template <int KIND>
int launch_cols_narrow(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                       int64_t ldf, int64_t NF, double* out, int64_t ldo, double scale,
                       int accumulate, cudaStream_t stream);

// EXPERIMENT (dev knob SGM_WIDE_BULK=1): hierarchical kind, row segments staged by bulk copies.
int launch_cols_wide_bulk(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                          int64_t ldf, int64_t ncols, double* out, int64_t ldo, double scale,
                          int accumulate, cudaStream_t stream, bool* launched) {
    *launched = false;
    const PlanDev& d = plan->dev;
    const int warps = 8;
    const size_t smem = (size_t)warps * sgm::WIDE_BULK_STAGES * (2048 + 8) + (size_t)d.n_knots * 16 +
                        (size_t)warps * ((size_t)d.B * 8 + (size_t)d.E * 4 + 32 * 4);
    if (smem > (size_t)plan->max_smem_optin / 2) return SGM_OK;
    const int64_t n_tiles = (ncols + 255) / 256;
    const int pts_per_block = warps * cols_points_per_warp(plan, P, n_tiles, warps);
    const int64_t gx = (P + pts_per_block - 1) / pts_per_block;
    if (gx > 0x7fffffff || n_tiles > 65535) return SGM_OK;
    auto kern = sgm::eval_cols_wide_bulk_kernel;
    SGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)gx, (unsigned)n_tiles), warps * 32, smem, stream>>>(
        plan->dev, x, P, ldx, f, ldf, out, ldo, scale, accumulate, pts_per_block, ncols);
    SGM_CUDA(cudaGetLastError());
    *launched = true;
    return SGM_OK;
}

template <int KIND>
int launch_cols(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                int64_t ldf, int64_t NF, double* out, int64_t ldo, double scale, int accumulate,
                const int* perm, cudaStream_t stream) {
    // full 256-column (or 128-column) tiles through the wide kernel, the ragged rest through
    // the guarded kernel; needs 16-byte aligned rows
    const bool aligned = ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                         (ldf % 2 == 0) && (ldo % 2 == 0);
    if (aligned && !plan->knobs.no_wide && NF >= 64 && plan->dev.n_corners > 0) {
        const int tile = NF >= 512 ? 256 : (NF >= 128 ? 128 : 64);
        const int64_t n_tiles = NF / tile;
        bool launched = false;
        if constexpr (KIND == SGM_HIER_PW_LINEAR) {
            if (plan->knobs.wide_bulk && NF >= 256 && plan->dev.n_corners > 0) {
                const int64_t even = NF & ~(int64_t)1;
                int rcb = launch_cols_wide_bulk(plan, x, P, ldx, f, ldf, even, out, ldo, scale, accumulate, stream, &launched);
                if (rcb) return rcb;
                if (launched) {
                    if (even == NF) return SGM_OK;
                    return launch_cols_narrow<KIND>(plan, x, P, ldx, f + even, ldf, 1, out + even, ldo,
                                                    scale, accumulate, stream);
                }
            }
        }
        if (perm && plan->knobs.wide_pairs != 0) {
            // the pair kernel guards its last tile: all (even many) columns in one launch,
            // an odd last column through the guarded narrow kernel
            const int64_t even = NF & ~(int64_t)1;
            const int64_t nt2 = (even + tile - 1) / tile;
            if (plan->knobs.wide_pts == 4 && NF >= 128) {
                int rc4 = launch_cols_wide4<KIND>(plan, x, P, ldx, f, ldf, even, out, ldo, scale, accumulate, perm, stream, &launched);
                if (rc4) return rc4;
            }
            if constexpr (KIND == SGM_HIER_PW_LINEAR) {
                // hierarchical kind: one accumulator set per point (no per-term partial sums), so
                // four points fit on 256-column tiles (default for plans with at most 8 keyed directions: config 4 x1.07-1.16, config 3 (d = 64) x1.0; dev knob SGM_WIDE_PTS=8 forces it, =2 the pair kernel)
                const bool low_dim = plan->first_level_keys <= 8;     // neighbours share nearly all rows
                if ((plan->knobs.wide_pts == 8 || (plan->knobs.wide_pts == 0 && low_dim)) && NF >= 256) {
                    int rc4 = launch_cols_wide4<KIND, 4>(plan, x, P, ldx, f, ldf, even, out, ldo, scale, accumulate, perm, stream, &launched);
                    if (rc4) return rc4;
                }
            }
            int rc2 = launched ? SGM_OK : tile == 256
                ? launch_cols_wide2<KIND, 4>(plan, x, P, ldx, f, ldf, nt2, even, out, ldo, scale, accumulate, perm, stream, &launched)
                : tile == 128
                ? launch_cols_wide2<KIND, 2>(plan, x, P, ldx, f, ldf, nt2, even, out, ldo, scale, accumulate, perm, stream, &launched)
                : launch_cols_wide2<KIND, 1>(plan, x, P, ldx, f, ldf, nt2, even, out, ldo, scale, accumulate, perm, stream, &launched);
            if (rc2) return rc2;
            if (launched) {
                if (even == NF) return SGM_OK;
                return launch_cols_narrow<KIND>(plan, x, P, ldx, f + even, ldf, 1, out + even, ldo,
                                                scale, accumulate, stream);
            }
        }
        int rc = launched ? SGM_OK : tile == 256
            ? launch_cols_wide<KIND, 4>(plan, x, P, ldx, f, ldf, n_tiles, out, ldo, scale, accumulate, stream)
            : tile == 128
            ? launch_cols_wide<KIND, 2>(plan, x, P, ldx, f, ldf, n_tiles, out, ldo, scale, accumulate, stream)
            : launch_cols_wide<KIND, 1>(plan, x, P, ldx, f, ldf, n_tiles, out, ldo, scale, accumulate, stream);
        if (rc) return rc;
        const int64_t done = n_tiles * tile;
        if (done == NF) return SGM_OK;
        return launch_cols_narrow<KIND>(plan, x, P, ldx, f + done, ldf, NF - done, out + done, ldo,
                                        scale, accumulate, stream);
    }
    return launch_cols_narrow<KIND>(plan, x, P, ldx, f, ldf, NF, out, ldo, scale, accumulate, stream);
}

template <int KIND>
int launch_cols_narrow(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                       int64_t ldf, int64_t NF, double* out, int64_t ldo, double scale,
                       int accumulate, cudaStream_t stream) {
    if (NF <= 32) return launch_cols_cpl<KIND, 1>(plan, x, P, ldx, f, ldf, NF, out, ldo, scale, accumulate, stream);
    if (NF <= 64) return launch_cols_cpl<KIND, 2>(plan, x, P, ldx, f, ldf, NF, out, ldo, scale, accumulate, stream);
    return launch_cols_cpl<KIND, 4>(plan, x, P, ldx, f, ldf, NF, out, ldo, scale, accumulate, stream);
}


int eval_one(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f, int64_t M,
             int64_t NF, int64_t ldf, double* out, int64_t ldo, double scale, int accumulate,
             void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (!plan) return sgm_fail(SGM_ERR_INVALID, "eval: null plan");
    if (P < 0 || NF < 1 || ldx < plan->dev.N || ldf < NF || ldo < NF)
        return sgm_fail(SGM_ERR_INVALID, "eval: bad shapes / leading dimensions");
    if (M < plan->M)
        return sgm_failf(SGM_ERR_INVALID, "eval: f has %lld rows, the plan's node maps address %lld",
                         (long long)M, (long long)plan->M);
    if (P == 0) return SGM_OK;
    if (!x || !f || !out) return sgm_fail(SGM_ERR_INVALID, "eval: null buffer");
    DeviceGuard guard(plan->device);
    const int kind = plan->dev.kind;
    // one scalar evaluation (kernel A family) of column `col` of f into column `col` of out
    auto scalar_column = [&](int64_t col) -> int {
        const size_t g_bytes = g_bytes_of(plan);
        if (workspace_bytes < g_bytes || !workspace)
            return sgm_fail(SGM_ERR_WORKSPACE, "eval: workspace too small");
        if (kind == SGM_PW_LINEAR && use_small_kernel(plan, P)) {
            if (small_jit_ready(plan, P))
                return launch_points_small_jit(plan, x, P, ldx, f + col, ldf, out + col, ldo, scale, accumulate, stream);
            return launch_points_small(plan, x, P, ldx, f + col, ldf, out + col, ldo, scale, accumulate, stream);
        }
        double* G = static_cast<double*>(workspace);
        unsigned char* scratch = static_cast<unsigned char*>(workspace) + g_bytes;
        const bool h2 = use_h2_kernel(plan);
        const int64_t n = h2 ? plan->dev.h2_n_vals : plan->n_vals;
        const int gblocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)plan->sm_count * 8);
        const bool prefix = use_prefix_kernel(plan, P);
        sgm::pregather_kernel<<<std::max(gblocks, 1), 256, 0, stream>>>(
            h2 ? plan->dev.h2_maps : prefix ? plan->dev.gmaps : plan->dev.maps, n, f + col, ldf, 1, G);
        SGM_CUDA(cudaGetLastError());
        const size_t sb = workspace_bytes - g_bytes;
        double* o = out + col;
        if (prefix)
            return launch_points_prefix(plan, x, P, ldx, G, o, ldo, scale, accumulate, scratch, sb, stream);
        switch (kind) {
            case SGM_PW_LINEAR:
                if (plan->kmax > 4)
                    return launch_points_split<sgm::SGM_LINEAR_WIDE>(plan, x, P, ldx, G, o, ldo, scale, accumulate, scratch, sb, stream);
                return launch_points_split<SGM_PW_LINEAR>(plan, x, P, ldx, G, o, ldo, scale, accumulate, scratch, sb, stream);
            case SGM_PW_QUADRATIC: return launch_points_split<SGM_PW_QUADRATIC>(plan, x, P, ldx, G, o, ldo, scale, accumulate, scratch, sb, stream);
            case SGM_PW_CUBIC: return launch_points_split<SGM_PW_CUBIC>(plan, x, P, ldx, G, o, ldo, scale, accumulate, scratch, sb, stream);
            case SGM_HIER_PW_LINEAR: return launch_points_split<SGM_HIER_PW_LINEAR>(plan, x, P, ldx, G, o, ldo, scale, accumulate, scratch, sb, stream);
            default: return launch_points_split<SGM_LAGRANGE>(plan, x, P, ldx, G, o, ldo, scale, accumulate, scratch, sb, stream);
        }
    };
    auto column_kernels = [&](const double* fc, double* oc, int64_t ncols) -> int {
        // locality permutation of the points for the two-points-per-warp wide kernel (scratch
        // behind the G area, shared with the scalar path: launches are ordered on the stream)
        const int* perm = nullptr;
        const size_t g_bytes = g_bytes_of(plan);
        if (ncols >= 64 && workspace && workspace_bytes > g_bytes + sort_scratch_bytes(P)) {
            int rc = locality_sort(plan, x, P, ldx, static_cast<unsigned char*>(workspace) + g_bytes,
                                   workspace_bytes - g_bytes, stream, &perm, kSortMinPointsCols);
            if (rc) return rc;
        }
        switch (kind) {
            case SGM_PW_LINEAR: return launch_cols<SGM_PW_LINEAR>(plan, x, P, ldx, fc, ldf, ncols, oc, ldo, scale, accumulate, perm, stream);
            case SGM_PW_QUADRATIC: return launch_cols<SGM_PW_QUADRATIC>(plan, x, P, ldx, fc, ldf, ncols, oc, ldo, scale, accumulate, perm, stream);
            case SGM_PW_CUBIC: return launch_cols<SGM_PW_CUBIC>(plan, x, P, ldx, fc, ldf, ncols, oc, ldo, scale, accumulate, perm, stream);
            case SGM_HIER_PW_LINEAR: return launch_cols<SGM_HIER_PW_LINEAR>(plan, x, P, ldx, fc, ldf, ncols, oc, ldo, scale, accumulate, perm, stream);
            default: return launch_cols<SGM_LAGRANGE>(plan, x, P, ldx, fc, ldf, ncols, oc, ldo, scale, accumulate, perm, stream);
        }
    };
    if (NF == 1) return scalar_column(0);
    // Few output columns: the lanes-across-columns kernels leave most lanes idle, so every
    // column goes through the scalar kernel (one launch per column, the tables stay in L2).
    // Wide outputs: full 128/256-column tiles through the column kernels, the ragged rest
    // column by column.
    // measured (tools/exp_nf*.py): the packed pw-linear scalar kernel beats every column
    // kernel below 128 columns; for the nested-walk kinds the crossover is 16..25 columns
    const bool fast_scalar = (kind == SGM_PW_LINEAR && plan->kmax <= 4) || kind == SGM_HIER_PW_LINEAR;
    const int64_t col_loop_max = plan->knobs.col_loop_max >= 0 ? plan->knobs.col_loop_max : (fast_scalar ? 127 : 16);
    const int64_t wide_cols = NF >= 512 ? NF / 256 * 256
                              : (NF >= 128 ? NF / 128 * 128 : (fast_scalar ? 0 : NF / 64 * 64));
    const bool aligned = ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                         (ldf % 2 == 0) && (ldo % 2 == 0);
    int64_t done = 0;
    {   // Wide outputs on enough points: ONE launch of the two-points-per-warp kernel covers
        // all columns (its last 256-column tile is guarded), instead of full tiles + a ragged
        // remainder through the narrower kernels (config 4, NF = 1000: 84 -> 67 ms).  A short
        // tail of a packed pw-linear plan still goes column by column through the scalar kernel.
        const size_t g_bytes = g_bytes_of(plan);
        const bool pairs_ok = aligned && plan->dev.n_corners > 0 && NF >= 512 && P >= kSortMinPointsCols &&
                              P < 0x7fffffff && plan->dev.n_key >= 4 && !plan->knobs.no_sort &&
                              plan->knobs.wide_pairs != 0 && workspace &&
                              workspace_bytes > g_bytes + sort_scratch_bytes(P);
        if (pairs_ok) {
            const int64_t tail = NF % 256;
            const int64_t main_cols = (fast_scalar && tail <= 48) ? NF - tail : NF;
            int rc = column_kernels(f, out, main_cols);
            if (rc) return rc;
            for (int64_t c = main_cols; c < NF; ++c) {
                rc = scalar_column(c);
                if (rc) return rc;
            }
            return SGM_OK;
        }
    }
    if (wide_cols && aligned && plan->dev.n_corners > 0) {
        int rc = column_kernels(f, out, wide_cols);
        if (rc) return rc;
        done = wide_cols;
    }
    if (NF - done > col_loop_max) return column_kernels(f + done, out + done, NF - done);
    for (int64_t c = done; c < NF; ++c) {
        int rc = scalar_column(c);
        if (rc) return rc;
    }
    return SGM_OK;
}

// Segmented evaluation: out[s, p, :] = sum over the terms of segment s.  Scalar outputs (and
// few columns, one at a time) through the split kernel's SEG flavour on pre-gathered values,
// wider outputs through the guarded lanes-across-columns kernel.
template <int KIND>
SplitKernel split_seg_kernel_for(int warps, int chunk) {
    if (warps == 16 && chunk == 64) return sgm::eval_points_split_kernel<KIND, 16, 64, false, true>;
    if (warps == 8 && chunk == 32) return sgm::eval_points_split_kernel<KIND, 8, 32, false, true>;
    return sgm::eval_points_split_kernel<KIND, 4, 32, false, true>;
}

template <int KIND>
int launch_segments_scalar(sgm_plan* plan, const PlanDev& dev, const double* x, int64_t P, int64_t ldx,
                           const double* G, double* out, int64_t ldo, cudaStream_t stream) {
    const PlanDev& d = plan->dev;
    const size_t fixed = (size_t)d.n_knots * 8 * (KIND == SGM_LAGRANGE ? 2 : 1) + (size_t)d.B * 32 * 8 +
                         32 * 8 + 64 + (size_t)d.E * 32 * 2;
    const int cand[3][2] = {{16, 64}, {8, 32}, {4, 32}};
    for (const auto& c : cand) {
        const size_t smem = fixed + 2 * (size_t)c[1] * 32 * 8;
        if (smem > (size_t)plan->max_smem_optin) continue;
        SplitKernel kern = split_seg_kernel_for<KIND>(c[0], c[1]);
        SGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int bps = 0;
        SGM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, c[0] * 32, smem));
        if (bps < 1) continue;
        const int64_t tiles = (P + 31) / 32;
        const int64_t blocks = std::min<int64_t>(tiles, (int64_t)plan->sm_count * bps);
        kern<<<(unsigned)blocks, c[0] * 32, smem, stream>>>(dev, x, P, ldx, G, out, ldo, 1.0, 0, nullptr);
        SGM_CUDA(cudaGetLastError());
        return SGM_OK;
    }
    return sgm_fail(SGM_ERR_OVERFLOW, "eval_segments: bracket tables exceed shared memory");
}

template <int KIND, int CPL>
int launch_segments_cols(sgm_plan* plan, const PlanDev& dev, const double* x, int64_t P, int64_t ldx,
                         const double* f, int64_t ldf, int64_t NF, double* out, int64_t ldo,
                         cudaStream_t stream) {
    const PlanDev& d = plan->dev;
    const size_t fixed = (size_t)d.n_knots * 8 * (KIND == SGM_LAGRANGE ? 2 : 1);
    const size_t per_warp = (size_t)d.B * 8 + (size_t)d.E * 4;
    int warps = 8;
    while (warps > 1 && fixed + per_warp * warps > (size_t)plan->max_smem_optin) warps >>= 1;
    const size_t smem = fixed + per_warp * warps;
    if (smem > (size_t)plan->max_smem_optin)
        return sgm_fail(SGM_ERR_OVERFLOW, "eval_segments: per-point basis table exceeds shared memory");
    const int64_t gy = (NF + 32 * CPL - 1) / (32 * CPL);
    const int pts_per_block = warps * cols_points_per_warp(plan, P, gy, warps);
    const int64_t gx = (P + pts_per_block - 1) / pts_per_block;
    if (gx > 0x7fffffff || gy > 65535) return sgm_fail(SGM_ERR_OVERFLOW, "grid too large");
    auto kern = sgm::eval_cols_kernel<KIND, CPL, true>;
    SGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)gx, (unsigned)gy), warps * 32, smem, stream>>>(
        dev, x, P, ldx, f, ldf, NF, out, ldo, 1.0, 0, pts_per_block);
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

template <int KIND>
int eval_segments_kind(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                       int64_t NF, int64_t ldf, double* out, int64_t ldo, int64_t seg_stride,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    PlanDev dev = plan->dev;                   // per-launch copy carrying the output stride
    dev.seg_stride = seg_stride;
    if (NF >= 24) {
        constexpr int CK = sgm::base_kind(KIND);
        if (NF <= 32) return launch_segments_cols<CK, 1>(plan, dev, x, P, ldx, f, ldf, NF, out, ldo, stream);
        if (NF <= 64) return launch_segments_cols<CK, 2>(plan, dev, x, P, ldx, f, ldf, NF, out, ldo, stream);
        return launch_segments_cols<CK, 4>(plan, dev, x, P, ldx, f, ldf, NF, out, ldo, stream);
    }
    const size_t g_bytes = align_up((size_t)plan->n_vals * 8, 256);
    if (workspace_bytes < g_bytes || !workspace)
        return sgm_fail(SGM_ERR_WORKSPACE, "eval_segments: workspace too small");
    double* G = static_cast<double*>(workspace);
    const int64_t n = plan->n_vals;
    const int gblocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)plan->sm_count * 8);
    for (int64_t col = 0; col < NF; ++col) {
        sgm::pregather_kernel<<<std::max(gblocks, 1), 256, 0, stream>>>(plan->dev.maps, n, f + col, ldf, 1, G);
        SGM_CUDA(cudaGetLastError());
        int rc = launch_segments_scalar<KIND>(plan, dev, x, P, ldx, G, out + col, ldo, stream);
        if (rc) return rc;
    }
    return SGM_OK;
}

int eval_segments(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f, int64_t NF,
                  int64_t ldf, double* out, int64_t ldo, int64_t seg_stride, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream) {
    switch (plan->dev.kind) {
        case SGM_PW_LINEAR:
            if (plan->kmax > 4)
                return eval_segments_kind<sgm::SGM_LINEAR_WIDE>(plan, x, P, ldx, f, NF, ldf, out, ldo, seg_stride, workspace, workspace_bytes, stream);
            return eval_segments_kind<SGM_PW_LINEAR>(plan, x, P, ldx, f, NF, ldf, out, ldo, seg_stride, workspace, workspace_bytes, stream);
        case SGM_PW_QUADRATIC: return eval_segments_kind<SGM_PW_QUADRATIC>(plan, x, P, ldx, f, NF, ldf, out, ldo, seg_stride, workspace, workspace_bytes, stream);
        case SGM_PW_CUBIC: return eval_segments_kind<SGM_PW_CUBIC>(plan, x, P, ldx, f, NF, ldf, out, ldo, seg_stride, workspace, workspace_bytes, stream);
        case SGM_LAGRANGE: return eval_segments_kind<SGM_LAGRANGE>(plan, x, P, ldx, f, NF, ldf, out, ldo, seg_stride, workspace, workspace_bytes, stream);
        default: return sgm_fail(SGM_ERR_INVALID, "eval_segments: unsupported plan kind");
    }
}

}  // namespace

extern "C" {

const char* sgm_last_error(void) { return g_last_error.c_str(); }
int sgm_version(void) { return 100; }

int sgm_plan_create(const sgm_plan_desc* desc, int device, sgm_plan** out) {
    if (!out) return sgm_fail(SGM_ERR_INVALID, "plan_create: null output");
    *out = nullptr;
    if (!desc || desc->N <= 0 || desc->T < 0 || desc->kind < 0 ||
        (desc->kind > 3 && desc->kind != SGM_HIER_PW_LINEAR))
        return sgm_fail(SGM_ERR_INVALID, "plan_create: bad descriptor");
    if (desc->kind == SGM_HIER_PW_LINEAR && desc->n_fam > 0 && !desc->fam_newrank)
        return sgm_fail(SGM_ERR_INVALID, "plan_create: hierarchical plan without new-knot ranks");
    if (desc->T > 0 && (!desc->coeffs || !desc->term_fam || !desc->map_off || !desc->maps))
        return sgm_fail(SGM_ERR_INVALID, "plan_create: null table");
    if (desc->T > 0x7fffffff) return sgm_fail(SGM_ERR_OVERFLOW, "plan_create: too many terms");
    if (desc->n_seg < 0 || (desc->n_seg > 0 && desc->kind == SGM_HIER_PW_LINEAR))
        return sgm_fail(SGM_ERR_INVALID, "plan_create: bad segment count / segmented hierarchical plan");
    const int kind = desc->kind;
    const int N = desc->N;
    const int64_t T = desc->T;
    if (desc->n_fam < 0 || (desc->n_fam > 0 && (!desc->fam_off || !desc->fam_knots)))
        return sgm_fail(SGM_ERR_INVALID, "plan_create: bad family table");
    if (kind == SGM_LAGRANGE && desc->n_fam > 0 && !desc->fam_lambda)
        return sgm_fail(SGM_ERR_INVALID, "plan_create: Lagrange plan without lambda table");
    for (int f = 0; f < desc->n_fam; ++f) {
        const int64_t m = desc->fam_off[f + 1] - desc->fam_off[f];
        if (m < 2) return sgm_fail(SGM_ERR_INVALID, "plan_create: knot family with fewer than 2 knots");
        if (m > 65535) return sgm_fail(SGM_ERR_OVERFLOW, "plan_create: more than 65535 knots in one direction");
    }
    const int64_t n_knots = desc->n_fam ? desc->fam_off[desc->n_fam] : 0;
    if (n_knots > 0x3fffffff) return sgm_fail(SGM_ERR_OVERFLOW, "plan_create: knot table too large");

    // entries of a term table per direction: all knots, or (hierarchical form) the new knots
    std::vector<int> fam_extent(std::max(desc->n_fam, 1), 0);
    for (int f = 0; f < desc->n_fam; ++f) {
        const int m = (int)(desc->fam_off[f + 1] - desc->fam_off[f]);
        if (kind == SGM_HIER_PW_LINEAR) {
            const int basis = desc->fam_basis ? desc->fam_basis[f] : (int)SGM_HB_HAT;
            if (basis < SGM_HB_HAT || basis > SGM_HB_QUAD_END)
                return sgm_fail(SGM_ERR_INVALID, "plan_create: unknown hierarchical basis");
            int n_new = 0;
            for (int i = 0; i < m; ++i) {
                const int r = desc->fam_newrank[desc->fam_off[f] + i];
                if (r >= 0) { if (r != n_new) return sgm_fail(SGM_ERR_INVALID, "plan_create: new-knot ranks must count up"); ++n_new; }
            }
            if (basis == SGM_HB_HAT)
                for (int i = 0; i + 1 < m; ++i) {
                    const bool a = desc->fam_newrank[desc->fam_off[f] + i] >= 0;
                    const bool b = desc->fam_newrank[desc->fam_off[f] + i + 1] >= 0;
                    if (a == b) return sgm_fail(SGM_ERR_INVALID, "plan_create: every bracket needs exactly one new knot");
                }
            if (basis == SGM_HB_QUAD_MID) {
                if (m < 3 || m % 2 == 0 || n_new != (m - 1) / 2)
                    return sgm_fail(SGM_ERR_INVALID, "plan_create: pw-quadratic hierarchical family needs an odd knot count with the odd positions new");
                for (int i = 0; i < m; ++i)
                    if ((desc->fam_newrank[desc->fam_off[f] + i] >= 0) != (i % 2 == 1))
                        return sgm_fail(SGM_ERR_INVALID, "plan_create: pw-quadratic hierarchical family needs an odd knot count with the odd positions new");
            }
            if ((basis == SGM_HB_LAG_ONE || basis == SGM_HB_QUAD_END) && n_new != 1)
                return sgm_fail(SGM_ERR_INVALID, "plan_create: a one-knot hierarchical family needs exactly one new knot");
            if (basis == SGM_HB_QUAD_END && m != 3)
                return sgm_fail(SGM_ERR_INVALID, "plan_create: SGM_HB_QUAD_END needs a 3-knot family");
            fam_extent[f] = n_new;
        } else {
            fam_extent[f] = m;
        }
    }
    // entries = distinct (direction, family) pairs in order of first use
    std::vector<int> entry_of((size_t)N * std::max(desc->n_fam, 1), -1);
    std::vector<Entry> entries;
    std::vector<TermHdr> hdr((size_t)T);
    std::vector<TermRec> rec((size_t)T);
    int64_t n_packed = 0;
    std::vector<TermDim> tdim;
    int B = 0;
    int64_t corners = 0;
    int kmax = 0;
    for (int64_t t = 0; t < T; ++t) {
        const int32_t* fam_row = desc->term_fam + t * N;
        const int64_t off = desc->map_off[t];
        const int64_t S = desc->map_off[t + 1] - off;
        if (off > 0x7fffffff || S > 0x7fffffff)
            return sgm_fail(SGM_ERR_OVERFLOW, "plan_create: node maps exceed the 32-bit packed layout");
        TermHdr h{0, (int)tdim.size(), (int)off, 1};
        int64_t check = 1, ncorner = 1;
        for (int n = 0; n < N; ++n) {
            const int fam = fam_row[n];
            if (fam < 0) continue;
            if (fam >= desc->n_fam)
                return sgm_fail(SGM_ERR_INVALID, "plan_create: family index out of range");
            const int m = (int)(desc->fam_off[fam + 1] - desc->fam_off[fam]);
            int& e = entry_of[(size_t)n * desc->n_fam + fam];
            if (e < 0) {
                e = (int)entries.size();
                entries.push_back(Entry{n, (int)desc->fam_off[fam], m, B});
                B += basis_width(kind, m);
            }
            tdim.push_back(TermDim{entries[e].boff, 0, radix_of(kind, m), e});
            check *= fam_extent[fam];
            ncorner *= radix_of(kind, m);
            if (ncorner > 0x3fffffff || check > 0x7fffffff)
                return sgm_fail(SGM_ERR_OVERFLOW, "plan_create: term with more than 2^30 corner contributions");
            ++h.k;
        }
        if (check != S) return sgm_fail(SGM_ERR_INVALID, "plan_create: map size does not match node counts");
        if (kind == SGM_HIER_PW_LINEAR) {
            int stride = 1;                                // surplus tables: FIRST direction fastest
            for (int j = 0; j < h.k; ++j) {
                TermDim& d = tdim[h.dim_off + j];
                d.stride = stride;
                stride *= fam_extent[desc->term_fam[t * N + entries[d.ent].dim]];
            }
        } else {
            int stride = 1;                                // C order: last active direction fastest
            for (int j = h.k - 1; j >= 0; --j) {
                TermDim& d = tdim[h.dim_off + j];
                d.stride = stride;
                stride *= fam_extent[desc->term_fam[t * N + entries[d.ent].dim]];
            }
        }
        h.ncorner = (int)ncorner;
        hdr[t] = h;
        TermRec& r = rec[t];
        std::memset(&r, 0, sizeof r);
        r.coeff = desc->coeffs[t];
        r.val_off = h.val_off;
        r.info = h.k & 0xff;
        bool packable = (kind == SGM_PW_LINEAR || kind == SGM_HIER_PW_LINEAR) && h.k <= 4 &&
                        (h.k == 0 || kind == SGM_HIER_PW_LINEAR || tdim[h.dim_off + h.k - 1].stride == 1);
        for (int j = 0; j < h.k && packable; ++j) {
            const TermDim& d = tdim[h.dim_off + j];
            if (d.ent > 0xffff || d.stride > 0xffff) { packable = false; break; }
            r.ent[j] = (unsigned short)d.ent;
            r.stride[j] = (unsigned short)d.stride;
        }
        if (packable) {                   // counted only when the record really is packed
            r.info |= sgm::TERMREC_PACKED;
            ++n_packed;
        }
        corners += ncorner;
        kmax = std::max(kmax, h.k);
    }
    // common prefix (leading (direction, family) entries) with the previous term, bits 12..14
    for (int64_t t = 1; t < T; ++t) {
        TermRec& r = rec[t];
        const TermRec& q = rec[t - 1];
        if (!(r.info & sgm::TERMREC_PACKED) || !(q.info & sgm::TERMREC_PACKED)) continue;
        const int kk = std::min(r.info & 0xff, q.info & 0xff);
        int cp = 0;
        while (cp < kk && cp < 3 && r.ent[cp] == q.ent[cp]) ++cp;
        r.info |= cp << 12;
    }
    // hierarchical form: a second copy of the records grouped by the number of directions
    // (stable, so every group stays lexicographic), with the common prefixes recomputed
    std::vector<TermRec> hrec;
    int hgroup[6] = {0, 0, 0, 0, 0, 0};
    if (kind == SGM_HIER_PW_LINEAR && n_packed == T) {
        for (int kk = 0; kk <= 4; ++kk) {
            hgroup[kk] = (int)hrec.size();
            for (int64_t t = 0; t < T; ++t)
                if ((rec[t].info & 0xff) == kk) {
                    TermRec r = rec[t];
                    r.info &= ~(7 << 12);
                    if ((int)hrec.size() > hgroup[kk]) {
                        const TermRec& q = hrec.back();
                        int cp = 0;
                        while (cp < kk && cp < 3 && r.ent[cp] == q.ent[cp]) ++cp;
                        r.info |= cp << 12;
                    }
                    hrec.push_back(r);
                }
        }
        hgroup[5] = (int)hrec.size();
    }
    for (int64_t i = 0; i < desc->map_off[T]; ++i)
        if (desc->maps[i] < 0 || desc->maps[i] >= desc->M)
            return sgm_fail(SGM_ERR_INVALID, "plan_create: map entry outside [0, M)");
    // prefix kernel (pw-linear, every term packed): records and node map for value tables laid
    // out with the FIRST active direction fastest, plus the number of leading entries each
    // term shares with its predecessor
    std::vector<TermRec> prec;
    std::vector<int> gmaps;
    if (kind == SGM_PW_LINEAR && n_packed == T && T > 0) {
        prec.resize((size_t)T);
        gmaps.resize((size_t)desc->map_off[T]);
        bool ok = true;
        for (int64_t t = 0; t < T && ok; ++t) {
            const TermHdr& h = hdr[t];
            TermRec r = rec[t];
            if (!(r.info & sgm::TERMREC_PACKED)) { ok = false; break; }
            int ext[4] = {1, 1, 1, 1}, cs[4] = {0, 0, 0, 0}, fs[4] = {0, 0, 0, 0};
            int stride = 1;
            for (int j = 0; j < h.k; ++j) {
                const TermDim& d = tdim[h.dim_off + j];
                ext[j] = entries[d.ent].n;
                cs[j] = d.stride;
                fs[j] = stride;
                if (stride > 0xffff || d.ent > 1023) ok = false;
                r.ent[j] = (unsigned short)(d.ent * 64);   // byte offset into the [E][32] u16 table
                r.stride[j] = (unsigned short)stride;
                stride *= ext[j];
            }
            int cp = 0;
            if (t > 0) {
                const int kq = prec[t - 1].info & 0xf;
                while (cp < kq && cp < h.k - 1 && cp < 3 && r.ent[cp] == prec[t - 1].ent[cp]) ++cp;
            }
            r.info = h.k | (cp << 4);
            prec[t] = r;
            const int* src = desc->maps + h.val_off;
            int* dst = gmaps.data() + h.val_off;
            for (int i0 = 0; i0 < ext[0]; ++i0)
                for (int i1 = 0; i1 < ext[1]; ++i1)
                    for (int i2 = 0; i2 < ext[2]; ++i2)
                        for (int i3 = 0; i3 < ext[3]; ++i3)
                            dst[i0 * fs[0] + i1 * fs[1] + i2 * fs[2] + i3 * fs[3]] =
                                src[i0 * cs[0] + i1 * cs[1] + i2 * cs[2] + i3 * cs[3]];
        }
        if (!ok) { prec.clear(); gmaps.clear(); }
    }
    // cost-balanced split of every 64-term chunk into 8 runs of consecutive terms (one per
    // warp of the prefix kernel).  Cost model in issued instructions (from the SASS, then
    // scanned on a B200 with tools/exp_cost.py -- the optimum is flat): a term costs ~45 + 4.5
    // per corner, +35 for the first term of a run (no prefix to reuse); the warp that adds up
    // the previous chunk's products (chunk index - 1 mod 8) carries ~2.5 per term of that
    // chunk on top.
    std::vector<unsigned char> run_start;
    if (!prec.empty()) {
        constexpr int PW = 8, PCHUNK = 64;
        auto knob = [](const char* name, double dflt) {             // dev knobs for the cost model
            const char* v = std::getenv(name);
            return v ? std::atof(v) : dflt;
        };
        const double c_term = knob("SGM_COST_TERM", 45.0), c_corner = knob("SGM_COST_CORNER", 4.5),
                     c_first = knob("SGM_COST_FIRST", 35.0), c_sum = knob("SGM_COST_SUM", 2.5);
        const int64_t n_chunks = (T + PCHUNK - 1) / PCHUNK;
        run_start.resize((size_t)n_chunks * (PW + 1));
        for (int64_t c = 0; c < n_chunks; ++c) {
            const int64_t t0 = c * PCHUNK;
            const int nt = (int)std::min<int64_t>(PCHUNK, T - t0);
            const int prev_nt = c == 0 ? (int)(T - (n_chunks - 1) * PCHUNK) : PCHUNK;
            const int summer = (int)((c + n_chunks - 1) % n_chunks) & (PW - 1);
            double cost[PCHUNK], total = c_first * PW + c_sum * prev_nt;
            for (int i = 0; i < nt; ++i) {
                cost[i] = c_term + c_corner * hdr[t0 + i].ncorner;
                total += cost[i];
            }
            unsigned char* rs = run_start.data() + c * (PW + 1);
            int pos = 0;
            double remaining = total;
            for (int w = 0; w < PW; ++w) {
                rs[w] = (unsigned char)pos;
                const double target = remaining / (PW - w);
                double load = c_first + (w == summer ? c_sum * prev_nt : 0.0);
                // take terms while that brings the load closer to the target
                while (pos < nt && (w == PW - 1 || std::fabs(load + cost[pos] - target) <= std::fabs(load - target))) {
                    load += cost[pos];
                    ++pos;
                }
                remaining -= load;
            }
            rs[PW] = (unsigned char)nt;
        }
    }

    int dev_count = 0;
    SGM_CUDA(cudaGetDeviceCount(&dev_count));
    if (device < 0 || device >= dev_count) return sgm_fail(SGM_ERR_INVALID, "plan_create: bad device");
    DeviceGuard guard(device);
    sgm_plan* plan = new sgm_plan();
    plan->device = device;
    plan->knobs = Knobs::from_env();
    plan->dev.dev_flags = plan->knobs.dev_flags;
    plan->M = desc->M;
    cudaDeviceGetAttribute(&plan->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&plan->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&plan->cc_major, cudaDevAttrComputeCapabilityMajor, device);
    cudaDeviceGetAttribute(&plan->cc_minor, cudaDevAttrComputeCapabilityMinor, device);
    plan->n_vals = desc->map_off[T];
    plan->corners = corners;
    plan->all_packed = (n_packed == T);
    plan->kmax = kmax;
    for (const Entry& e : entries) plan->max_entry_n = std::max(plan->max_entry_n, e.n);
    PlanDev& d = plan->dev;
    d.kind = kind; d.N = N; d.E = (int)entries.size(); d.B = B; d.T = (int)T;
    d.n_knots = (int)n_knots;

    {   // locality key (<= 16 bits): which bracket bits of which directions decide the most rows.
        // A use of an entry with n knots depends on the leading ceil(log2(n + 1)) - 1 bits of the
        // bracket search in its direction; benefit(d, b) = uses of direction d that depend on bit
        // b.  The 16 most beneficial (direction, bit) pairs form the key (benefit is monotone in
        // b, so a direction always gets a prefix of its bits); config 2 / 5: one bit for each of
        // the 16 most used directions, config 3 (anisotropic): several bits for the leading
        // directions, config 4 (d = 8): two bits per direction.
        // SGM_KEY_MODE=1 (dev knob): one bit per direction first, spare bits as second level.
        std::vector<int64_t> usage(N, 0);
        std::vector<int> fine_fam(N, -1);
        std::vector<std::array<int64_t, 9>> dep(N);
        for (auto& a : dep) a.fill(0);
        auto bits_of = [](int n) { int b = 0; while ((1 << (b + 1)) < n + 1) ++b; return b; };   // ceil(log2(n+1)) - 1
        for (const TermDim& td : tdim) {
            const Entry& e = entries[td.ent];
            ++usage[e.dim];
            if (e.n >= 3 && (fine_fam[e.dim] < 0 || e.n > entries[fine_fam[e.dim]].n)) fine_fam[e.dim] = td.ent;
            const int nb = std::min(bits_of(e.n), 8);
            for (int b = 0; b < nb; ++b) ++dep[e.dim][b];
        }
        struct Cand { int64_t benefit; int dim; int bit; };
        std::vector<Cand> cands;
        for (int n = 0; n < N; ++n) {
            if (fine_fam[n] < 0) continue;
            const int maxb = std::min(bits_of(entries[fine_fam[n]].n), 8);
            for (int b = 0; b < maxb; ++b) {
                int64_t ben = dep[n][b];
                if (plan->knobs.key_mode == 1) ben = b == 0 ? ((int64_t)1 << 40) + usage[n] : (b == 1 ? usage[n] : 0);
                if (ben > 0) cands.push_back(Cand{ben, n, b});
            }
        }
        // most beneficial first; ties: lower bit first, then the HIGHER direction (a differing
        // bracket in a low direction is cheap for the kernels whose value tables are stored with
        // the first active direction fastest: neighbouring brackets = neighbouring addresses)
        std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) {
            if (a.benefit != b.benefit) return a.benefit > b.benefit;
            if (a.bit != b.bit) return a.bit < b.bit;
            return a.dim > b.dim;
        });
        if (cands.size() > 16) cands.resize(16);
        std::vector<int> nbits(N, 0), first_pos(N, -1);
        for (size_t i = 0; i < cands.size(); ++i) {
            ++nbits[cands[i].dim];
            if (first_pos[cands[i].dim] < 0) first_pos[cands[i].dim] = (int)i;
        }
        std::vector<int> gdims;
        for (int n = 0; n < N; ++n) if (nbits[n] > 0) gdims.push_back(n);
        std::sort(gdims.begin(), gdims.end(), [&](int a, int b) { return first_pos[a] < first_pos[b]; });
        d.n_kgrp = (int)gdims.size();
        d.n_key = 0;
        for (int i = 0; i < d.n_kgrp; ++i) {
            const Entry& e = entries[fine_fam[gdims[i]]];
            d.kg_dim[i] = e.dim;
            d.kg_off[i] = e.knot_off;
            d.kg_n[i] = (unsigned short)e.n;
            d.kg_bits[i] = (unsigned char)nbits[gdims[i]];
            d.n_key += nbits[gdims[i]];
        }
        plan->first_level_keys = d.n_kgrp;
    }
    std::vector<double> coeff(desc->coeffs, desc->coeffs + T);
    std::vector<int> maps(desc->maps, desc->maps + desc->map_off[T]);
    std::vector<double> knots(desc->fam_knots, desc->fam_knots + n_knots);
    std::vector<double> lambda;
    if (kind == SGM_LAGRANGE) lambda.assign(desc->fam_lambda, desc->fam_lambda + n_knots);
    // global corner list for the wide column kernel (skipped when absurdly long)
    std::vector<int> corner_term, corner_off;
    if (corners > 0 && corners < ((int64_t)1 << 27)) {
        corner_off.resize(T + 1);
        corner_term.reserve((size_t)corners);
        int64_t run = 0;
        for (int64_t t = 0; t < T; ++t) {
            corner_off[t] = (int)run;
            for (int c = 0; c < hdr[t].ncorner; ++c) corner_term.push_back((int)t);
            run += hdr[t].ncorner;
        }
        corner_off[T] = (int)run;
        d.n_corners = (int)run;
    } else {
        d.n_corners = 0;
    }
    // hierarchical kinds: the `lambda` table carries, per family (n knots at offset o, slot of n
    // doubles): n shorts = new-knot ranks, one short = sgm_hier_basis, and in the LAST double of
    // the slot the denominator prod_{i != p} (g_p - g_i) of the one-knot bases (2(n+1) <= 8(n-1)
    // for every n >= 2, so the fields never overlap)
    if (kind == SGM_HIER_PW_LINEAR) {
        lambda.assign((size_t)n_knots, 0.0);
        for (int f = 0; f < desc->n_fam; ++f) {
            const int64_t o = desc->fam_off[f];
            const int m = (int)(desc->fam_off[f + 1] - o);
            const int basis = desc->fam_basis ? desc->fam_basis[f] : (int)SGM_HB_HAT;
            short* nr = reinterpret_cast<short*>(lambda.data() + o);
            int p = -1;
            for (int i = 0; i < m; ++i) {
                nr[i] = (short)desc->fam_newrank[o + i];
                if (nr[i] >= 0 && p < 0) p = i;
            }
            nr[m] = (short)basis;
            if (basis == SGM_HB_LAG_ONE || basis == SGM_HB_QUAD_END) {
                double den = 1.0;
                for (int i = 0; i < m; ++i)
                    if (i != p) den = den * (desc->fam_knots[o + p] - desc->fam_knots[o + i]);
                lambda[(size_t)(o + m - 1)] = den;
            }
        }
    }
    // grouped hierarchical kernel: group / step tables over the trie of the multi-indices
    std::vector<int> h2_groups, h2_parents, h2_steps, h2_maps, h2_split;
    int h2_root = -1;
    if (kind == SGM_HIER_PW_LINEAR && T > 0) {
        if (sgm_build_hier_groups(T, N, desc->term_fam, desc->n_fam, fam_extent.data(), desc->map_off,
                                  h2_groups, h2_parents, h2_steps, h2_maps, &h2_root, sgm::H2_LISTS,
                                  h2_split)) {
            for (int& v : h2_maps)                      // slot -> row of f through the caller's map
                if (v >= 0) v = desc->maps[v];
        } else {
            h2_groups.clear(); h2_parents.clear(); h2_steps.clear(); h2_maps.clear(); h2_split.clear();
        }
    }
    std::vector<unsigned> flags(4, 0u);
    int rc = SGM_OK;
    const unsigned* cflags = nullptr;
    const int* h2g = nullptr; const int* h2p = nullptr;
    if ((rc = upload(plan, hdr, &d.hdr)) || (rc = upload(plan, tdim, &d.tdim)) ||
        (rc = upload(plan, rec, &d.rec)) || (rc = upload(plan, hrec, &d.hrec)) ||
        (rc = upload(plan, coeff, &d.coeff)) || (rc = upload(plan, entries, &d.ent)) ||
        (rc = upload(plan, maps, &d.maps)) || (rc = upload(plan, prec, &d.prec)) ||
        (rc = upload(plan, gmaps, &d.gmaps)) || (rc = upload(plan, run_start, &d.run_start)) ||
        (rc = upload(plan, corner_term, &d.corner_term)) ||
        (rc = upload(plan, corner_off, &d.corner_off)) || (rc = upload(plan, knots, &d.knots)) ||
        (rc = upload(plan, lambda, &d.lambda)) ||
        (rc = upload(plan, h2_groups, &h2g)) || (rc = upload(plan, h2_parents, &h2p)) ||
        (rc = upload(plan, h2_steps, &d.h2_steps)) || (rc = upload(plan, h2_maps, &d.h2_maps)) ||
        (rc = upload(plan, h2_split, &d.h2_split)) ||
        (rc = upload(plan, flags, &cflags))) {
        sgm_plan_destroy(plan);
        return rc;
    }
    d.flags = const_cast<unsigned*>(cflags);
    // kernel A0: few packed pw-linear terms -> metadata into a parameter block; entries listed
    // direction by direction, finest family first, coarser families of a direction derived from
    // the fine bracket when their knots are bit-for-bit among the fine ones (sgm_small_plan_build);
    // the same tables feed the plan-specialised kernel, generated on demand
    if (kind == SGM_PW_LINEAR && n_packed == T && sgm_small_plan_build(desc, &plan->small_desc)) {
        const SgmJitSmallDesc& jd = plan->small_desc;
        sgm::SmallPlan& sp = plan->small;
        sp.T = (int)T; sp.E = (int)jd.ents.size();
        sp.n_knots = (int)jd.knots.size();
        sp.n_tab = (int)jd.tabs.size();
        for (int q = 0; q < sp.E; ++q) {
            const SgmJitSmallDesc::Ent& e = jd.ents[q];
            sp.ent[q] = sgm::SmallEntry{e.dim | (e.n << 16), e.knot_off, e.steps, e.tab_off};
        }
        for (int64_t t = 0; t < T; ++t) {
            const SgmJitSmallDesc::Term& jt = jd.terms[t];
            sgm::SmallTerm& st = sp.term[t];
            st.coeff = jt.coeff; st.val_off = jt.val_off; st.k = jt.k;
            for (int j = 0; j < 4; ++j) {
                st.ent[j] = (unsigned short)jt.ent[j];
                st.stride[j] = (unsigned short)jt.stride[j];
            }
        }
        const unsigned short* dtabs = nullptr;
        if ((rc = upload(plan, jd.tabs, &dtabs))) { sgm_plan_destroy(plan); return rc; }
        const double* dsk = nullptr;
        if ((rc = upload(plan, jd.knots, &dsk))) { sgm_plan_destroy(plan); return rc; }
        plan->small_tabs = dtabs;
        plan->small_knots = dsk;
        plan->small_ok = true;
    }
    d.h2_groups = reinterpret_cast<const int4*>(h2g);
    d.h2_parents = h2p;
    d.h2_n_groups = (int)(h2_groups.size() / 12);
    d.h2_n_vals = (int)h2_maps.size();
    d.h2_root_off = h2_root;
    {   // the grouped kernel must be launchable, because sgm_eval gathers the value table in its
        // layout whenever h2_ok is set
        const size_t h2_smem = (size_t)d.n_knots * 16 + (size_t)(d.B + 1) * 32 * 8 +
                               (size_t)sgm::H2_LISTS * (32 * 8 + 4) + (size_t)(d.E + 1) * 32 * 2 + 64 + 64;
        plan->h2_ok = !h2_maps.empty() && h2_smem <= (size_t)plan->max_smem_optin;
    }
    d.seg_id = nullptr;
    d.seg_stride = 0;
    if (desc->n_seg > 0) {
        if (!desc->seg_off || desc->seg_off[0] != 0 || desc->seg_off[desc->n_seg] != T) {
            sgm_plan_destroy(plan);
            return sgm_fail(SGM_ERR_INVALID, "plan_create: segment offsets must run from 0 to T");
        }
        std::vector<int> seg_id((size_t)T);
        for (int sgi = 0; sgi < desc->n_seg; ++sgi) {
            if (desc->seg_off[sgi + 1] <= desc->seg_off[sgi]) {
                sgm_plan_destroy(plan);
                return sgm_fail(SGM_ERR_INVALID, "plan_create: empty or unordered segment");
            }
            for (int64_t t = desc->seg_off[sgi]; t < desc->seg_off[sgi + 1]; ++t) seg_id[t] = sgi;
        }
        if ((rc = upload(plan, seg_id, &d.seg_id))) {
            sgm_plan_destroy(plan);
            return rc;
        }
        plan->n_seg = desc->n_seg;
    }
    plan->prefix_ok = !prec.empty();
    if (!plan->prefix_ok) { d.prec = nullptr; d.gmaps = nullptr; d.run_start = nullptr; }
    for (int i = 0; i < 6; ++i) d.hgroup[i] = hgroup[i];
    decide_prefix(plan);
    *out = plan;
    return SGM_OK;
}

int sgm_plan_reload_knobs(sgm_plan* plan) {
    if (!plan) return sgm_fail(SGM_ERR_INVALID, "reload_knobs: null plan");
    DeviceGuard guard(plan->device);
    plan->knobs = Knobs::from_env();
    plan->dev.dev_flags = plan->knobs.dev_flags;
    decide_prefix(plan);
    return SGM_OK;
}

void sgm_plan_destroy(sgm_plan* plan) {
    if (!plan) return;
    DeviceGuard guard(plan->device);
    small_jit_join(plan);
    for (void* p : plan->allocations) cudaFree(p);
    sgm_jit_unload(&plan->small_jit);
    delete plan;
}

int sgm_plan_jit_wait(sgm_plan* plan) {
    if (!plan) return 0;
    small_jit_join(plan);
    return plan->jit_state.load(std::memory_order_acquire);
}

int sgm_host_small_kernel_source(const sgm_plan_desc* desc, int32_t threads, int32_t points_per_thread,
                                 int32_t compile_cc, char* buf, size_t buf_bytes, size_t* needed) {
    SgmJitSmallDesc jd;
    if (threads < 32 || threads > 1024 || threads % 32 || points_per_thread < 1 || points_per_thread > 8)
        return sgm_fail(SGM_ERR_INVALID, "small_kernel_source: bad launch shape");
    if (!sgm_small_plan_build(desc, &jd))
        return sgm_fail(SGM_ERR_INVALID, "small_kernel_source: the plan does not qualify for the small-plan kernels");
    const std::string src = sgm_jit_small_source(jd, threads, points_per_thread, "sgm_small_jit");
    if (needed) *needed = src.size() + 1;
    if (buf && buf_bytes) std::snprintf(buf, buf_bytes, "%s", src.c_str());
    if (compile_cc > 0) {
        std::vector<char> cubin;
        std::string why;
        if (!sgm_jit_compile(src, compile_cc / 10, compile_cc % 10, &cubin, &why))
            return sgm_failf(SGM_ERR_CUDA, "small_kernel_source: %s", why.c_str());
    }
    return SGM_OK;
}

int sgm_plan_jit_state(sgm_plan* plan, char* why, size_t why_bytes) {
    if (!plan) return 0;
    std::lock_guard<std::mutex> lock(plan->jit_mutex);
    if (why && why_bytes) {
        std::snprintf(why, why_bytes, "%s", plan->jit_why.c_str());
    }
    return plan->jit_state.load();
}

int64_t sgm_plan_table_bytes(const sgm_plan* plan) { return plan ? plan->table_bytes : 0; }
int64_t sgm_plan_corners_per_point(const sgm_plan* plan) { return plan ? plan->corners : 0; }

int sgm_plan_take_flags(sgm_plan* plan, void* stream, uint32_t* flags) {
    if (!plan || !flags) return sgm_fail(SGM_ERR_INVALID, "take_flags: null argument");
    DeviceGuard guard(plan->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned host = 0;
    SGM_CUDA(cudaMemcpyAsync(&host, plan->dev.flags, sizeof host, cudaMemcpyDeviceToHost, s));
    SGM_CUDA(cudaMemsetAsync(plan->dev.flags, 0, sizeof host, s));
    SGM_CUDA(cudaStreamSynchronize(s));
    *flags = host;
    return SGM_OK;
}

size_t sgm_eval_workspace_bytes(const sgm_plan* plan, int64_t P, int64_t NF) {
    if (!plan) return 256;
    size_t bytes = g_bytes_of(plan);
    const PointLaunch L = choose_point_launch(plan);
    size_t scratch = 0;
    if (L.global_basis) {
        const int64_t tiles = (P + L.threads - 1) / std::max(L.threads, 1);
        const int64_t blocks = std::min<int64_t>(tiles, (int64_t)plan->sm_count * L.blocks_per_sm);
        scratch = (size_t)std::max<int64_t>(blocks, 1) * L.threads *
                  ((size_t)plan->dev.B * 8 + (size_t)plan->dev.E * 2);
    }
    if (P >= (NF >= 64 ? kSortMinPointsCols : kSortMinPoints)) scratch = std::max(scratch, sort_scratch_bytes(P));
    return bytes + scratch + 256;
}

int sgm_eval(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f, int64_t M,
             int64_t NF, int64_t ldf, double* out, int64_t ldo, void* workspace,
             size_t workspace_bytes, void* stream) {
    if (plan && plan->n_seg > 0)
        return sgm_fail(SGM_ERR_INVALID, "eval: segmented plan (use sgm_eval_segments)");
    return eval_one(plan, x, P, ldx, f, M, NF, ldf, out, ldo, 1.0, 0, workspace, workspace_bytes,
                    static_cast<cudaStream_t>(stream));
}

int sgm_eval_segments(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                      int64_t M, int64_t NF, int64_t ldf, double* out, int64_t ldo,
                      int64_t seg_stride, void* workspace, size_t workspace_bytes, void* stream) {
    if (!plan || plan->n_seg < 1) return sgm_fail(SGM_ERR_INVALID, "eval_segments: plan has no segments");
    if (P < 0 || NF < 1 || ldx < plan->dev.N || ldf < NF || ldo < NF || seg_stride < (P > 0 ? (P - 1) * ldo + NF : 0))
        return sgm_fail(SGM_ERR_INVALID, "eval_segments: bad shapes / leading dimensions");
    if (M < plan->M) return sgm_fail(SGM_ERR_INVALID, "eval_segments: f has fewer rows than the plan addresses");
    if (P == 0) return SGM_OK;
    if (!x || !f || !out) return sgm_fail(SGM_ERR_INVALID, "eval_segments: null buffer");
    DeviceGuard guard(plan->device);
    return eval_segments(plan, x, P, ldx, f, NF, ldf, out, ldo, seg_stride, workspace, workspace_bytes,
                         static_cast<cudaStream_t>(stream));
}

int sgm_segment_sumsq(const double* vals, int32_t n_seg, int64_t P, int64_t NF, int64_t ld,
                      int64_t seg_stride, double* sumsq, void* stream) {
    if (n_seg < 0 || P < 0 || NF < 1 || ld < NF) return sgm_fail(SGM_ERR_INVALID, "segment_sumsq: bad shapes");
    if (n_seg == 0) return SGM_OK;
    if (!vals || !sumsq) return sgm_fail(SGM_ERR_INVALID, "segment_sumsq: null buffer");
    sgm::segment_sumsq_kernel<<<(unsigned)n_seg, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        vals, P, NF, ld, seg_stride, sumsq);
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

int sgm_surplus_norms(sgm_plan* plan, const double* x, int64_t P, int64_t ldx, const double* f,
                      int64_t M, int64_t NF, int64_t ldf, double* sumsq, void* workspace,
                      size_t workspace_bytes, void* stream) {
    if (!plan || plan->n_seg < 1) return sgm_fail(SGM_ERR_INVALID, "surplus_norms: plan has no segments");
    if (P < 0 || NF < 1) return sgm_fail(SGM_ERR_INVALID, "surplus_norms: bad shapes");
    const size_t eval_ws = sgm_eval_workspace_bytes(plan, P, NF);
    const size_t slab = align_up((size_t)P * (size_t)NF * 8, 256);
    if (!workspace || workspace_bytes < eval_ws + slab * (size_t)plan->n_seg)
        return sgm_fail(SGM_ERR_WORKSPACE, "surplus_norms: workspace too small (sgm_surplus_norms_workspace_bytes)");
    double* deltas = reinterpret_cast<double*>(static_cast<unsigned char*>(workspace) + align_up(eval_ws, 256));
    int rc = sgm_eval_segments(plan, x, P, ldx, f, M, NF, ldf, deltas, NF, (int64_t)(slab / 8), workspace,
                               eval_ws, stream);
    if (rc) return rc;
    DeviceGuard guard(plan->device);
    return sgm_segment_sumsq(deltas, plan->n_seg, P, NF, NF, (int64_t)(slab / 8), sumsq, stream);
}

size_t sgm_surplus_norms_workspace_bytes(const sgm_plan* plan, int64_t P, int64_t NF) {
    if (!plan) return 256;
    return align_up(sgm_eval_workspace_bytes(plan, P, NF), 256) +
           align_up((size_t)P * (size_t)NF * 8, 256) * (size_t)std::max(plan->n_seg, 1) + 256;
}

int sgm_eval_signed_sum(sgm_plan* const* plans, const double* signs, int32_t n_parts,
                        const double* x, int64_t P, int64_t ldx, const double* const* f,
                        const int64_t* M, const int64_t* ldf, int64_t NF, double* out, int64_t ldo,
                        void* workspace, size_t workspace_bytes, void* stream) {
    if (!plans || !signs || !f || !M || !ldf || n_parts < 1)
        return sgm_fail(SGM_ERR_INVALID, "eval_signed_sum: null argument");
    for (int i = 0; i < n_parts; ++i) {
        int rc = eval_one(plans[i], x, P, ldx, f[i], M[i], NF, ldf[i], out, ldo, signs[i], i > 0,
                          workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
        if (rc) return rc;
    }
    return SGM_OK;
}

int sgm_hier_apply(double* alpha, int64_t NF, int64_t ld, const int32_t* child, const int32_t* pa,
                   const int32_t* pb, const double* wa, const double* wb, int64_t n, void* stream) {
    if (n < 0 || NF < 1 || ld < NF) return sgm_fail(SGM_ERR_INVALID, "hier_apply: bad shapes");
    if (n == 0) return SGM_OK;
    if (!alpha || !child || !pa || !pb || !wa || !wb)
        return sgm_fail(SGM_ERR_INVALID, "hier_apply: null buffer");
    const int64_t total = n * NF;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    sgm::hier_apply_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        alpha, NF, ld, child, pa, pb, wa, wb, n);
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

int sgm_max_rel_deviation(const double* a, const double* b, int64_t n, double* out, void* stream) {
    if (n < 0 || !out || (n > 0 && (!a || !b))) return sgm_fail(SGM_ERR_INVALID, "max_rel_deviation: bad argument");
    sgm::max_rel_deviation_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n, out);
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

int sgm_hier_apply_batches(double* alpha, int64_t NF, int64_t ld, const int32_t* child,
                           const int32_t* pa, const int32_t* pb, const double* wa, const double* wb,
                           const int64_t* batch_off_host, const int64_t* batch_off_dev,
                           int32_t n_batches, void* stream) {
    if (n_batches < 0 || NF < 1 || ld < NF) return sgm_fail(SGM_ERR_INVALID, "hier_apply_batches: bad shapes");
    if (n_batches == 0) return SGM_OK;
    if (!alpha || !batch_off_host || !batch_off_dev)
        return sgm_fail(SGM_ERR_INVALID, "hier_apply_batches: null buffer");
    const int64_t total = batch_off_host[n_batches];
    if (total == 0) return SGM_OK;
    if (!child || !pa || !pb || !wa || !wb) return sgm_fail(SGM_ERR_INVALID, "hier_apply_batches: null buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (total * NF <= ((int64_t)1 << 22)) {        // one thread-block cluster walks all batches
        sgm::hier_apply_all_kernel<<<sgm::HIER_CLUSTER, 1024, 0, st>>>(alpha, NF, ld, child, pa, pb, wa, wb,
                                                                       batch_off_dev, n_batches);
        SGM_CUDA(cudaGetLastError());
        return SGM_OK;
    }
    for (int b = 0; b < n_batches; ++b) {           // large tables: one launch per batch
        const int64_t lo = batch_off_host[b], n = batch_off_host[b + 1] - lo;
        if (n <= 0) continue;
        const int blocks = (int)std::min<int64_t>((n * NF + 255) / 256, 148 * 16);
        sgm::hier_apply_kernel<<<blocks, 256, 0, st>>>(alpha, NF, ld, child + lo, pa + lo, pb + lo,
                                                       wa + lo, wb + lo, n);
    }
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

int sgm_lc_brownian(const double* tt, int64_t nt, const double* yy, int64_t P, int32_t d,
                    int64_t ldy, double T, double sqrtT, const double* level_scale, int32_t L,
                    double* W, int64_t ldw, void* stream) {
    if (nt < 0 || P < 0 || d < 1 || ldy < d || ldw < nt || L < 0 || L > 40)
        return sgm_fail(SGM_ERR_INVALID, "lc_brownian: bad shapes");
    if (nt == 0 || P == 0) return SGM_OK;
    if (!tt || !yy || !W || (L > 0 && !level_scale))
        return sgm_fail(SGM_ERR_INVALID, "lc_brownian: null buffer");
    int dev = 0, sms = 148;
    SGM_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t lc_smem = 2 * (size_t)sgm::LC_ROWS * d * 8;  // double-buffered row groups
    if (L <= 12 && nt >= 32 && lc_smem <= 96 * 1024) {
        const int tpt = 2;
        const int64_t gx = (nt + 256 * tpt - 1) / (256 * tpt);
        // 8 blocks per SM in the grid (3 resident: 80 registers), at least 32 rows per block:
        // measured best of 4 / 8 / 16 / 32 (tools/exp_lc2.py)
        const int64_t per_sm = 8;
        int64_t rows = std::max<int64_t>(sgm::LC_ROWS, (P * gx + (int64_t)sms * per_sm - 1) / ((int64_t)sms * per_sm));
        int64_t gy = (P + rows - 1) / rows;
        if (gy > 65535) { rows = (P + 65534) / 65535; gy = (P + rows - 1) / rows; }
        using LcKernel = void (*)(const double*, int64_t, const double*, int64_t, int, int64_t, double,
                                  double, const double*, int, double*, int64_t, int);
        // contiguous, 16-byte aligned rows: the row groups travel by bulk asynchronous copies
        // (dev knob SGM_LC_BULK=0: the cp.async flavour, for comparison)
        static const bool bulk_off = [] { const char* v = std::getenv("SGM_LC_BULK"); return v && v[0] == '0'; }();
        const bool bulk = !bulk_off && ldy == d && d % 2 == 0 && (reinterpret_cast<uintptr_t>(yy) & 15) == 0;
        // compile-time row length for the full dyadic sizes (d = 2^L <= 64)
        LcKernel kern = bulk
            ? (L > 8 ? sgm::lc_brownian_tiles_kernel<12, 2, 2, 0, true>
               : L > 6 ? sgm::lc_brownian_tiles_kernel<8, 2, 2, 0, true>
               : d == 64 ? sgm::lc_brownian_tiles_kernel<6, 2, 3, 64, true>
               : d == 32 ? sgm::lc_brownian_tiles_kernel<6, 2, 3, 32, true>
               : d == 16 ? sgm::lc_brownian_tiles_kernel<6, 2, 3, 16, true>
                         : sgm::lc_brownian_tiles_kernel<6, 2, 3, 0, true>)
            : (L > 8 ? sgm::lc_brownian_tiles_kernel<12, 2, 2>
               : L > 6 ? sgm::lc_brownian_tiles_kernel<8, 2, 2>
               : d == 64 ? sgm::lc_brownian_tiles_kernel<6, 2, 3, 64>
               : d == 32 ? sgm::lc_brownian_tiles_kernel<6, 2, 3, 32>
               : d == 16 ? sgm::lc_brownian_tiles_kernel<6, 2, 3, 16>
                         : sgm::lc_brownian_tiles_kernel<6, 2, 3>);
        const size_t lc_smem_total = lc_smem + (bulk ? 16 : 0);       // + two mbarriers
        if (lc_smem_total > 48 * 1024)
            SGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc_smem_total));
        kern<<<dim3((unsigned)gx, (unsigned)gy), 256, lc_smem_total, st>>>(tt, nt, yy, P, d, ldy, T, sqrtT,
                                                                     level_scale, L, W, ldw, (int)rows);
    } else {
        const int64_t total = P * nt;
        const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sms * 16);
        sgm::lc_brownian_kernel<<<blocks, 256, 0, st>>>(tt, nt, yy, P, d, ldy, T, sqrtT,
                                                        level_scale, L, W, ldw);
    }
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

int sgm_kl_brownian(const double* tt, int64_t nt, const double* yy, int64_t P, int32_t d,
                    int64_t ldy, const double* factor, const double* freq, double* W,
                    int64_t ldw, void* stream) {
    if (nt < 0 || P < 0 || d < 0 || ldy < d || ldw < nt)
        return sgm_fail(SGM_ERR_INVALID, "kl_brownian: bad shapes");
    if (nt == 0 || P == 0) return SGM_OK;
    if (!tt || !yy || !W || (d > 0 && (!factor || !freq)))
        return sgm_fail(SGM_ERR_INVALID, "kl_brownian: null buffer");
    int dev = 0, sms = 148;
    SGM_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t kl_smem = ((size_t)d * sgm::KL_TT + (size_t)sgm::KL_ROWS * d) * 8;
    if (d >= 1 && P >= 64 && kl_smem <= 96 * 1024 && !std::getenv("SGM_KL_PER_OUTPUT")) {   // (dev knob)
        // tiled kernel: sines once per (block, time), rows streamed through shared memory
        const int64_t gx = (nt + sgm::KL_TT - 1) / sgm::KL_TT;
        int64_t rows = std::max<int64_t>(256, (P * gx + (int64_t)sms * 8 - 1) / ((int64_t)sms * 8));
        int64_t gy = (P + rows - 1) / rows;
        if (gy > 65535) { rows = (P + 65534) / 65535; gy = (P + rows - 1) / rows; }
        if (kl_smem > 48 * 1024)
            SGM_CUDA(cudaFuncSetAttribute(sgm::kl_brownian_tiles_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kl_smem));
        sgm::kl_brownian_tiles_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, kl_smem,
                                        static_cast<cudaStream_t>(stream)>>>(
            tt, nt, yy, P, d, ldy, factor, freq, W, ldw, (int)rows);
    } else {
        const int64_t total = P * nt;
        const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sms * 16);
        sgm::kl_brownian_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            tt, nt, yy, P, d, ldy, factor, freq, W, ldw);
    }
    SGM_CUDA(cudaGetLastError());
    return SGM_OK;
}

}  // extern "C"
