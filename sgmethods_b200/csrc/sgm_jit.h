// sgm_jit.h -- plan-specialised kernels: source generator, NVRTC compilation and module
// loading (sgm_jit.cpp).  Used for few-term pw-linear plans evaluated at many points, the one
// regime where the interpolant is HBM-bound and the interpretive overhead of a generic kernel
// (term records, entry tables, loop control) is what keeps it off the memory roofline.
#ifndef SGM_JIT_H
#define SGM_JIT_H

#include <string>
#include <vector>

#include "../../include/sgm_b200.h"

constexpr int SGM_SMALL_MAX_TERMS = 24, SGM_SMALL_MAX_ENTRIES = 16;

struct SgmJitSmallDesc {
    // (direction, knot family) pairs, listed direction by direction, finest family first
    struct Ent {
        int dim, n, knot_off;
        int steps;      // >= 0: rounds of the branch-free search over the +inf padded family;
                        // -1: nested in the finest family of its direction (bracket from tab_off)
        int tab_off;
    };
    struct Term { double coeff; int val_off, k; int ent[4]; int stride[4]; };
    std::vector<Ent> ents;
    std::vector<Term> terms;
    std::vector<double> knots;            // as uploaded to the device (padded)
    std::vector<unsigned short> tabs;     // as uploaded to the device
    int n_vals = 0;
};

// CUDA source of the kernel `name` specialised to the plan.  Signature of the kernel:
//   (const double* knots, const unsigned short* tabs, const double* x, long long P, long long ldx,
//    const int* maps, const double* f, long long ldf, double* out, long long ldo, double scale,
//    int accumulate)
std::string sgm_jit_small_source(const SgmJitSmallDesc& d, int threads, int ppt, const char* name);

// The tables of the small-plan kernels (A0 and its specialised variant) from a plan descriptor.
// False when the plan does not qualify: not pw-linear, more than SGM_SMALL_MAX_TERMS terms or
// SGM_SMALL_MAX_ENTRIES (direction, family) pairs, a term with more than 4 active directions,
// more than 4096 node values, segmented.
bool sgm_small_plan_build(const sgm_plan_desc* desc, SgmJitSmallDesc* out);

// NVRTC only (no driver, no GPU needed): cubin of `src` for sm_<major><minor>[a], --fmad=false.
bool sgm_jit_compile(const std::string& src, int cc_major, int cc_minor, std::vector<char>* cubin,
                     std::string* why);

// A loaded specialised kernel (driver-API handles kept opaque).
struct SgmJitKernel {
    void* module = nullptr;     // CUmodule
    void* function = nullptr;   // CUfunction
    int threads = 0, ppt = 0, blocks_per_sm = 0, regs = 0;
};

// Compiles `src` for the device's architecture (NVRTC, --fmad=false) and loads `name` into the
// CURRENT context.  Returns false with a reason in `why` when NVRTC / the driver library is not
// available or compilation fails; the caller then keeps the generic kernel.
bool sgm_jit_build(const std::string& src, const char* name, int cc_major, int cc_minor,
                   int threads, SgmJitKernel* out, std::string* why);
// cuLaunchKernel on `stream` (a cudaStream_t); returns the CUresult (0 = success).
int sgm_jit_launch(const SgmJitKernel& k, unsigned grid, void** args, void* stream);
void sgm_jit_unload(SgmJitKernel* k);

#endif
