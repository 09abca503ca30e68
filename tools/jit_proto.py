"""Prototype of the plan-specialised small-plan kernel (generator in Python, for experiments;
the product generator lives in csrc/sgm_jit.cpp).  Usage:
    python tools/jit_proto.py gen  [PPT] [TH]   -> writes /tmp/sgm_jit_proto.cu (+ cubin with nvcc)
    python tools/jit_proto.py run  [PPT] [TH]   -> GPU: NVRTC-compiles, times, compares with plan.eval
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def example1():
    from sgmethods_b200 import multi_index_sets as mis, nodes_1d
    from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
    S = mis.aniso_smolyak_mid_set(5, 10, 2 ** (np.linspace(0, 9, 10)))
    return SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)


def small_plan(it):
    """Entries (direction by direction, finest family first), packed knots / bracket tables, terms."""
    tf, off, kn = it._term_fam, it._fam_off, it._fam_knots
    T, N = tf.shape
    pairs = sorted({(d, int(tf[t, d])) for t in range(T) for d in range(N) if tf[t, d] >= 0},
                   key=lambda p: (p[0], -(off[p[1] + 1] - off[p[1]])))
    sk, tabs, ents = [], [], []
    fine = None
    for q, (d, fam) in enumerate(pairs):
        g = kn[off[fam]:off[fam + 1]]
        n = g.size
        e = dict(dim=d, n=n, koff=len(sk), fam=fam)
        nested = False
        if q > 0 and pairs[q - 1][0] == d:
            gf = fine["g"]
            tab, jc = [0], 0
            for i in range(gf.size):
                if jc < n and gf[i].tobytes() == g[jc].tobytes():
                    jc += 1
                tab.append(max(0, min(jc - 1, n - 2)))
            nested = jc == n
            if nested:
                e.update(steps=-1, toff=len(tabs))
                tabs += tab
                sk += list(g)
        if not nested:
            steps = 1
            while (1 << steps) - 1 < n:
                steps += 1
            e.update(steps=steps, toff=0, g=g)
            sk += list(g) + [np.inf] * ((1 << steps) - 1 - n)
            fine = e
        ents.append(e)
    eid = {p: q for q, p in enumerate(pairs)}
    terms = []
    for t in range(T):
        dims = [d for d in range(N) if tf[t, d] >= 0]
        ns = [int(off[tf[t, d] + 1] - off[tf[t, d]]) for d in dims]
        strides = [int(np.prod(ns[j + 1:])) for j in range(len(dims))]
        terms.append(dict(coeff=float(it._coeffs[t]), val_off=int(it._map_off[t]),
                          ents=[eid[(d, int(tf[t, d]))] for d in dims], strides=strides))
    return dict(ents=ents, sk=np.array(sk, dtype=np.float64), tabs=np.array(tabs, dtype=np.uint16),
                terms=terms, n_vals=int(it._map_off[-1]))


def generate(sp, PPT=2, TH=128, name="sgm_small_jit"):
    E, NK, NT, NV = len(sp["ents"]), sp["sk"].size, max(sp["tabs"].size, 1), sp["n_vals"]
    L = []
    A = L.append
    A(f"// generated: {E} entries, {len(sp['terms'])} terms, PPT={PPT}, TH={TH}")
    A(f'extern "C" __global__ void __launch_bounds__({TH}) {name}(')
    A("    const double* __restrict__ knots, const unsigned short* __restrict__ tabs,")
    A("    const double* __restrict__ x, long long P, long long ldx, const int* __restrict__ maps,")
    A("    const double* __restrict__ f, long long ldf, double* __restrict__ out, long long ldo,")
    A("    double scale, int accumulate) {")
    A(f"  __shared__ double sk[{NK}]; __shared__ double G[{NV}]; __shared__ unsigned short stab[{NT}];")
    A(f"  for (int i = threadIdx.x; i < {NK}; i += {TH}) sk[i] = knots[i];")
    A(f"  for (int i = threadIdx.x; i < {sp['tabs'].size}; i += {TH}) stab[i] = tabs[i];")
    A(f"  for (int i = threadIdx.x; i < {NV}; i += {TH}) G[i] = f[(long long)maps[i] * ldf];")
    A("  __syncthreads();")
    A(f"  const long long n_tiles = (P + {TH * PPT - 1}) / {TH * PPT};")
    A("  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {")
    A(f"    long long p[{PPT}]; bool ok[{PPT}];")
    A("#pragma unroll")
    A(f"    for (int q = 0; q < {PPT}; ++q) {{ const long long pp = tile * {TH * PPT} + q * {TH} + threadIdx.x; "
      "ok[q] = pp < P; p[q] = ok[q] ? pp : P - 1; }")
    A(f"    double xd[{PPT}]; int cnt[{PPT}];")
    for e, en in enumerate(sp["ents"]):
        A(f"    double y{e}[{PPT}], z{e}[{PPT}]; unsigned i{e}[{PPT}];")
    cur = -1
    for e, en in enumerate(sp["ents"]):
        if en["dim"] != cur:
            cur = en["dim"]
            A("#pragma unroll")
            A(f"    for (int q = 0; q < {PPT}; ++q) xd[q] = x[p[q] * ldx + {cur}];")
        if en["steps"] >= 0:
            A("#pragma unroll")
            A(f"    for (int q = 0; q < {PPT}; ++q) cnt[q] = 0;")
            for s in range(en["steps"] - 1, -1, -1):
                st = 1 << s
                A("#pragma unroll")
                A(f"    for (int q = 0; q < {PPT}; ++q) cnt[q] += (sk[{en['koff'] + st - 1} + cnt[q]] <= xd[q]) ? {st} : 0;")
            A("#pragma unroll")
            A(f"    for (int q = 0; q < {PPT}; ++q) i{e}[q] = (unsigned)max(0, min(cnt[q] - 1, {en['n'] - 2}));")
        else:
            A("#pragma unroll")
            A(f"    for (int q = 0; q < {PPT}; ++q) i{e}[q] = stab[{en['toff']} + cnt[q]];")
        A("#pragma unroll")
        A(f"    for (int q = 0; q < {PPT}; ++q) {{ const double g0 = sk[{en['koff']} + i{e}[q]], g1 = sk[{en['koff'] + 1} + i{e}[q]]; "
          f"y{e}[q] = (xd[q] - g0) / (g1 - g0); z{e}[q] = 1.0 - y{e}[q]; }}")
    A(f"    double acc[{PPT}];")
    A("#pragma unroll")
    A(f"    for (int q = 0; q < {PPT}; ++q) acc[q] = 0.0;")
    for t, tm in enumerate(sp["terms"]):
        k = len(tm["ents"])
        A(f"    // term {t}: k={k}")
        A("#pragma unroll")
        A(f"    for (int q = 0; q < {PPT}; ++q) {{")
        if k == 0:
            A(f"      acc[q] = acc[q] + {tm['coeff']!r} * G[{tm['val_off']}];")
        else:
            base = " + ".join(f"i{e}[q] * {s}u" for e, s in zip(tm["ents"], tm["strides"]))
            A(f"      const double* g = G + ({tm['val_off']}u + {base});")
            A("      double tp = 0.0;")
            for c in range(1 << k):
                w, o = None, 0
                for j, (e, s) in enumerate(zip(tm["ents"], tm["strides"])):
                    bit = (c >> (k - 1 - j)) & 1
                    fac = f"{'y' if bit else 'z'}{e}[q]"
                    w = fac if w is None else f"({w} * {fac})"
                    o += s if bit else 0
                A(f"      tp = tp + g[{o}] * {w};")
            A(f"      acc[q] = acc[q] + {tm['coeff']!r} * tp;")
        A("    }")
    A("#pragma unroll")
    A(f"    for (int q = 0; q < {PPT}; ++q) {{ if (!ok[q]) continue; double* o = out + p[q] * ldo; "
      "if (accumulate) *o = *o + scale * acc[q]; else *o = (scale == 1.0) ? acc[q] : scale * acc[q]; }")
    A("  }")
    A("}")
    return "\n".join(L) + "\n"


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "gen"
    PPT = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    TH = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    it = example1()
    sp = small_plan(it)
    if mode == "sweep":
        for ppt, th in ((1, 128), (1, 256), (2, 64), (2, 128), (2, 256), (3, 128), (4, 64), (4, 128)):
            os.system(f"{sys.executable} {os.path.abspath(__file__)} run {ppt} {th}")
        return
    src = generate(sp, PPT, TH)
    if mode == "gen":
        with open("/tmp/sgm_jit_proto.cu", "w") as fh:
            fh.write(src)
        os.system("nvcc -O3 -std=c++17 -fmad=false -gencode arch=compute_100a,code=sm_100a -Xptxas -v "
                  "-cubin -o /tmp/sgm_jit_proto.cubin /tmp/sgm_jit_proto.cu")
        return
    import torch
    from cuda.bindings import driver, nvrtc

    def chk(r):
        if int(r[0]) != 0:
            raise RuntimeError(str(r))
        return r[1:] if len(r) > 2 else (r[1] if len(r) == 2 else None)
    prog = chk(nvrtc.nvrtcCreateProgram(src.encode(), b"small.cu", 0, [], []))
    opts = [b"--gpu-architecture=sm_100a", b"--fmad=false", b"-std=c++17"]
    r = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if int(r[0]) != 0:
        n = chk(nvrtc.nvrtcGetProgramLogSize(prog)); log = b" " * n
        nvrtc.nvrtcGetProgramLog(prog, log); print(log.decode()); raise SystemExit(1)
    n = chk(nvrtc.nvrtcGetCUBINSize(prog)); cubin = b" " * n
    chk(nvrtc.nvrtcGetCUBIN(prog, cubin))
    dev = torch.device("cuda")
    torch.zeros(1, device=dev)
    mod = chk(driver.cuModuleLoadData(cubin))
    fn = chk(driver.cuModuleGetFunction(mod, b"sgm_small_jit"))
    P = 10 ** 7
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
    f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
    plan = it.device_plan(0)
    ref = torch.empty((P, 1), dtype=torch.float64, device=dev)
    plan.eval(x, f, ref)
    out = torch.zeros((P, 1), dtype=torch.float64, device=dev)
    knots = torch.from_numpy(sp["sk"]).to(dev)
    tabs = torch.from_numpy(sp["tabs"].astype(np.int16)).to(dev)
    maps = torch.from_numpy(it._maps.astype(np.int32)).to(dev)
    import ctypes
    args = [ctypes.c_void_p(knots.data_ptr()), ctypes.c_void_p(tabs.data_ptr()), ctypes.c_void_p(x.data_ptr()),
            ctypes.c_longlong(P), ctypes.c_longlong(it.N), ctypes.c_void_p(maps.data_ptr()),
            ctypes.c_void_p(f.data_ptr()), ctypes.c_longlong(1), ctypes.c_void_p(out.data_ptr()),
            ctypes.c_longlong(1), ctypes.c_double(1.0), ctypes.c_int(0)]
    argv = (ctypes.c_void_p * len(args))(*[ctypes.cast(ctypes.pointer(a), ctypes.c_void_p) for a in args])
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for bps in (4, 6, 8, 12, 16):
        grid = min((P + TH * PPT - 1) // (TH * PPT), 148 * bps)
        ts = []
        for i in range(8):
            flush.fill_(float(i))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            chk(driver.cuLaunchKernel(fn, grid, 1, 1, TH, 1, 1, 0, stream, ctypes.addressof(argv), 0))
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts[1:]))
        print(f"PPT={PPT} TH={TH} blocks/SM={bps:2d}: {ms:.4f} ms  {88 * P / ms / 1e6:7.0f} GB/s "
              f"({88 * P / ms / 1e6 / 6563.2:.3f} of HBM peak)  bit-identical: {torch.equal(out, ref)}")


if __name__ == "__main__":
    main()
