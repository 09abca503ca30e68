"""Sparse-grid interpolant (combination technique) evaluated by fused fp64 CUDA kernels.

Drop-in for the reference's ``sgmethods.sparse_grid_interpolant.SGInterpolant``
(sgmethods/sparse_grid_interpolant.py:12-279): same constructor signature, same public
attributes (``mid_set, card_mid_set, N, knots, lev2knots, tp_interpolant,
combination_coeffs, active_mids, active_tp_nodes_list, active_tp_dims, map_tp_to_SG, SG,
num_nodes, n_parallel, verbose``) and methods (``setup_interpolant, setup_SG, sample_on_SG,
interpolate``).

What changed underneath:
  * ``setup_interpolant`` / ``setup_SG`` (reference :65-139) build packed integer tables with
    hashing in the native library (``sgm_host_combination_coeffs``, ``sgm_grid_build``) --
    O(sum of tensor-grid sizes) instead of O(M^2) -- with bit-identical results; the list
    attributes of the reference are materialised lazily from the packed tables;
  * ``interpolate`` (reference :252-279, the per-term Python loop over
    TPInterpolatorWrapper / scipy RegularGridInterpolator) is ONE kernel launch per call
    (``sgm_eval``) that walks all terms on the GPU;
  * ``sample_on_SG`` stays a host callback (reference :141-250); only the search for
    recyclable samples is hash-accelerated (``sgm_host_match_rows``).

There is no CPU fallback: ``interpolate`` raises when no CUDA device / library is available.
"""
from multiprocessing import Pool

import numpy as np

from . import _lib
from .tp_inteprolants import TPPwLinearInterpolator, resolve_kind

NODE_TOL = 1.e-10     # the reference's node-identification tolerance (:129, :220)


class SGInterpolant:
    """Sparse grid interpolant: multi-index set + 1-D knots + level-to-knots map, evaluated
    with the combination technique  I[u](x) = sum_nu c_nu TP_nu[u](x)."""

    def __init__(self, mid_set, knots, lev2knots, tp_interpolant=TPPwLinearInterpolator,
                 n_parallel=1, verbose=True):
        """Args (as in the reference):
            mid_set (numpy.ndarray[int]): downward-closed multi-index set, one index per
                row, lexicographically sorted.
            knots (Callable[[int], numpy.ndarray[float]]): knots(n) = the n 1-D nodes
                (knots(1) must be 0 for dimension growth to be consistent).
            lev2knots (Callable[[numpy.ndarray[int]], numpy.ndarray[int]]): level -> number
                of nodes, applied to a whole multi-index (lev2knots(0) must be 1).
            tp_interpolant: one of the classes in ``tp_inteprolants`` (or a factory that
                returns one); selects the kernel.  Defaults to piecewise linear.
            n_parallel (int): processes used by ``sample_on_SG``.
            verbose (bool): print recycling statistics in ``sample_on_SG``.
        """
        self.verbose = verbose
        self.mid_set = mid_set
        self.card_mid_set = mid_set.shape[0]
        self.N = mid_set.shape[1]
        self.knots = knots
        self.lev2knots = lev2knots
        self.tp_interpolant = tp_interpolant
        self._kind = resolve_kind(tp_interpolant)
        self._lists = {}
        self._device_plans = {}
        self.setup_interpolant()
        self.num_nodes = 0
        self.setup_SG()
        self.n_parallel = n_parallel

    # ------------------------------------------------------------------ table construction
    def setup_interpolant(self):
        """Combination coefficients, active multi-indices and per-term node counts
        (reference :65-107).  ``lev2knots`` is applied to each active multi-index as a whole
        vector and ``knots`` once per distinct node count, exactly the calls the reference
        makes (the latter deduplicated)."""
        coeffs = _lib.combination_coeffs(self.mid_set)
        self._active_rows = np.flatnonzero(coeffs)
        self._coeffs = coeffs[self._active_rows]
        T = self._active_rows.size
        num_nodes = np.ones((T, self.N), dtype=np.int32)
        for t, row in enumerate(self._active_rows):
            num_nodes[t] = self.lev2knots(self.mid_set[row]).astype(int)
        if T and num_nodes.min() < 1:
            raise ValueError("lev2knots returned a node count < 1")
        self._num_nodes = num_nodes
        fam_m = np.unique(num_nodes[num_nodes > 1]).astype(np.int32)
        self._fam_m = fam_m
        self._fam_knots_list = []
        for m in fam_m:
            g = np.atleast_1d(np.asarray(self.knots(np.int64(m)), dtype=np.float64))
            if g.shape != (int(m),):
                raise ValueError(f"knots({m}) returned {g.shape[0]} nodes")
            self._fam_knots_list.append(np.ascontiguousarray(g))
        self._fam_off = np.concatenate(([0], np.cumsum(fam_m, dtype=np.int64))).astype(np.int64)
        self._fam_knots = (np.concatenate(self._fam_knots_list) if fam_m.size
                           else np.zeros(0))
        # family index of every (term, direction); -1 where the term has a single node
        lut = np.full(int(fam_m.max()) + 1 if fam_m.size else 2, -1, dtype=np.int32)
        lut[fam_m] = np.arange(fam_m.size, dtype=np.int32)
        self._term_fam = np.where(num_nodes > 1, lut[np.minimum(num_nodes, lut.size - 1)],
                                  -1).astype(np.int32)
        self._lists.clear()

    def setup_SG(self):
        """Sparse grid nodes and the maps tensor grid -> sparse grid (reference :109-139):
        first-seen numbering over terms in order and nodes in C order, nodes closer than
        1e-10 in the 1-norm identified, first-seen coordinates kept."""
        self._grid = _lib.HostGrid(self._num_nodes, self._fam_m, self._fam_off,
                                   self._fam_knots, NODE_TOL)
        self._map_off, self._maps = self._grid.maps()
        self.num_nodes = self._grid.num_nodes
        self._lists.pop("SG", None)
        self._lists.pop("map_tp_to_SG", None)
        self._device_plans.clear()

    # ------------------------------------------------- lazily materialised list attributes
    def _lazy(self, name, builder):
        if name not in self._lists:
            self._lists[name] = builder()
        return self._lists[name]

    @property
    def combination_coeffs(self):
        return self._lazy("combination_coeffs", lambda: list(self._coeffs))

    @property
    def active_mids(self):
        return self._lazy("active_mids", lambda: [self.mid_set[r, :] for r in self._active_rows])

    @property
    def active_tp_dims(self):
        return self._lazy("active_tp_dims",
                          lambda: [np.where(row > 1)[0] for row in self._num_nodes])

    @property
    def active_tp_nodes_list(self):
        def build():
            by_m = {int(m): g for m, g in zip(self._fam_m, self._fam_knots_list)}
            one = None
            out = []
            for row in self._num_nodes:
                act = row[row > 1]
                if act.size == 0:
                    if one is None:
                        one = self.knots(np.int64(1))
                    out.append((one,))
                else:
                    out.append(tuple(by_m[int(m)] for m in act))
            return out
        return self._lazy("active_tp_nodes_list", build)

    @property
    def map_tp_to_SG(self):
        def build():
            out = []
            for t, row in enumerate(self._num_nodes):
                shape = tuple(int(m) for m in row[row > 1]) or (1,)
                out.append(self._maps[self._map_off[t]:self._map_off[t + 1]]
                           .astype(np.int64).reshape(shape))
            return out
        return self._lazy("map_tp_to_SG", build)

    @property
    def SG(self):
        return self._lazy("SG", self._grid.nodes_dense)

    # ------------------------------------------------------------------------ host sampling
    def sample_on_SG(self, f, dim_f=None, old_xx=None, old_samples=None):
        """Sample ``f`` on the sparse grid, recycling ``old_samples`` given on ``old_xx``
        where the nodes coincide (1-norm distance < 1e-10 after zero-padding / truncating
        ``old_xx`` to the current dimension).  Host callback, semantics of the reference
        (:141-250) including: ``dim_f`` is ignored, the output width is taken from
        ``old_samples`` or from one evaluation at the first node, ``n_parallel > 1`` maps
        ``f`` over a process pool.  Returns an array of shape (num_nodes, dim_f)."""
        SG = self.SG
        assert SG.shape[0] > 0
        if old_samples is not None:
            old_samples = np.atleast_1d(old_samples)
            dim_f = old_samples.shape[1]
            assert dim_f >= 1
        else:
            dim_f = np.atleast_1d(f(SG[0])).size
        f_on_SG = np.zeros((self.num_nodes, dim_f))

        if (old_xx is None) or (old_samples is None):
            old_xx = np.zeros((0, self.N))
            old_samples = np.zeros((0, dim_f))
        assert old_xx.shape[0] == old_samples.shape[0]
        assert old_samples.shape[1] == dim_f

        if old_xx.shape[1] < self.N:
            pad = np.zeros((old_xx.shape[0], self.N - old_xx.shape[1]))
            old_xx = np.hstack((old_xx, pad))
        elif old_xx.shape[1] > self.N:
            # as in the reference (:206-212) the tail norm is truncated to int before the test
            tail = np.linalg.norm(old_xx[:, self.N:], ord=1, axis=1).astype(int)
            keep = np.where(tail == 0)[0]
            old_xx = old_xx[keep, 0:self.N]
            old_samples = old_samples[keep]

        match = _lib.match_rows(SG, old_xx, NODE_TOL)
        found = match >= 0
        f_on_SG[found] = old_samples[match[found]]
        to_compute = np.flatnonzero(~found)
        n_recycle = int(found.sum())

        if to_compute.size > 0:
            if self.n_parallel == 1:
                for idx in to_compute:
                    f_on_SG[idx, :] = f(SG[idx, :])
            elif self.n_parallel > 1:
                with Pool(self.n_parallel) as pool:
                    tmp = np.array(pool.map(f, [SG[idx, :] for idx in to_compute]))
                if len(tmp.shape) == 1:
                    tmp = tmp.reshape((-1, 1))
                f_on_SG[to_compute, :] = tmp
            else:
                raise ValueError('self.NParallel not int >= 1"')
        if self.verbose:
            print("Recycled", n_recycle, "; Discarted", old_xx.shape[0] - n_recycle,
                  "; Sampled", SG.shape[0] - n_recycle)
        return f_on_SG

    # -------------------------------------------------------------------------- evaluation
    def device_plan(self, device=None):
        """The packed tables resident on a GPU (created on first use, cached per device)."""
        from ._device import DevicePlan, torch_cuda
        torch = torch_cuda()
        device = torch.cuda.current_device() if device is None else int(device)
        if device not in self._device_plans:
            self._device_plans[device] = DevicePlan(
                self._kind, self.N, self.num_nodes, self._coeffs.astype(np.float64),
                self._term_fam, self._map_off, self._maps, self._fam_off, self._fam_knots,
                device=device)
        return self._device_plans[device]

    def _pipeline(self, torch, x_host, NF, device, after=None):
        """Host points (pinned torch tensor, or numpy array staged through two pinned
        buffers): the chunks are copied on a side stream while the previous chunk is being
        evaluated; the kernels stay on the current stream, in order.  A generator in two phases:
        ``next()`` enqueues the copy of the first chunk and returns -- the caller uploads the
        nodal values, hierarchises and launches the validation sample meanwhile --, then
        ``send((plan, values))`` runs the rest of the pipeline and returns the (P, NF) result."""
        P = x_host.shape[0]
        plan = f = None
        out = torch.empty((P, NF), dtype=torch.float64, device=device)
        compute = torch.cuda.current_stream(device)
        is_numpy = isinstance(x_host, np.ndarray)
        # Everything the pipeline keeps between calls is PER DEVICE: the copy stream, the two
        # pinned staging buffers (numpy input) with the events of their last H2D copy, and two
        # persistent device chunk buffers (double buffering; no allocator traffic inside the
        # loop -- chunk tensors allocated on the copy stream and handed to the compute stream
        # made the caching allocator fall back to cudaMalloc on some runs: end-to-end time
        # bimodal, 26 / 42 ms on config 2).
        pipes = getattr(self, "_pipes", None)
        if pipes is None:
            pipes = self._pipes = {}
        st = pipes.get(device.index)
        if st is None:
            st = pipes[device.index] = {
                "copier": torch.cuda.Stream(device),
                "dev_buf": [torch.empty((_H2D_CHUNK, self.N), dtype=torch.float64, device=device)
                            for _ in range(2)],
                "dev_free": [None, None],        # event: last evaluation reading the buffer
                "staging": None, "staging_free": [None, None]}
        if is_numpy and st["staging"] is None:
            st["staging"] = [torch.empty((_H2D_CHUNK, self.N), dtype=torch.float64).pin_memory()
                             for _ in range(2)]
        copier, dev_buf, dev_free = st["copier"], st["dev_buf"], st["dev_free"]
        staging, staging_free = st["staging"], st["staging_free"]
        # the copies wait for the work that was enqueued BEFORE this call (``after``, recorded at
        # the top of interpolate()), not for this call's own validation launches
        if after is not None:
            copier.wait_event(after)
        else:
            copier.wait_stream(compute)
        # a short first chunk (nothing overlaps its transfer), then the steady size
        bounds, lo = [], 0
        first = min(_H2D_FIRST, _H2D_CHUNK)
        while lo < P:
            hi = min(P, lo + (first if not bounds else _H2D_CHUNK))
            bounds.append((lo, hi))
            lo = hi
        for i, (lo, hi) in enumerate(bounds):
            b = i % 2
            if is_numpy:
                buf = staging[b]
                if staging_free[b] is not None:
                    staging_free[b].synchronize()
                src = buf[:hi - lo]
                _parallel_copy(src.numpy(), x_host[lo:hi, :self.N])
            else:
                src = x_host[lo:hi, :self.N]
            xd = dev_buf[b][:hi - lo]
            with torch.cuda.stream(copier):
                if dev_free[b] is not None:
                    copier.wait_event(dev_free[b])
                xd.copy_(src, non_blocking=True)
                done = torch.cuda.Event()
                done.record(copier)
            if is_numpy:
                staging_free[b] = done
            # enqueue the evaluation of this chunk right away so that it overlaps the host
            # side staging copy of the next one
            if plan is None:
                plan, f = yield              # the first transfer is in flight
            compute.wait_event(done)
            plan.eval(xd, f, out[lo:hi])
            dev_free[b] = torch.cuda.Event()
            dev_free[b].record(compute)
        yield out

    def _interpolate_pipelined(self, torch, plan, f, x_host, device, after=None):
        pipe = self._pipeline(torch, x_host, f.shape[1], device, after)
        next(pipe)
        return pipe.send((plan, f))

    def _check_operator_constraints(self):
        if self._kind == _lib.KIND_CUBIC and self._fam_m.size and self._fam_m.min() < 4:
            # reference: UnboundLocalError in TPPwCubicInterpolator.__call__ (:232)
            raise UnboundLocalError("pw-cubic interpolation needs at least 4 knots in every "
                                    "active direction")

    def hierarchical(self):
        """The work-reduced evaluator (pw-linear, pw-quadratic or Lagrange operator on nested
        knots): see hierarchical.py."""
        if getattr(self, "_hier", None) is None:
            from .hierarchical import HierarchicalPlan
            self._hier = HierarchicalPlan(self)
        return self._hier

    # --------------------------------------------------------------- method selection
    def _auto_candidate(self, host_rows):
        """The hierarchical plan if ``method="auto"`` may use it for a batch of ``host_rows``
        points, else None: the form exists for the pw-linear, pw-quadratic and Lagrange
        operators on nested knots (hierarchical.py; not for pw-cubic), small batches stay on the exact kernels (they are latency-bound
        either way and the exact path is bit-identical), and an interpolant on which the form
        once failed its validation keeps the exact kernels."""
        if self._kind == _lib.KIND_CUBIC or host_rows < AUTO_MIN_POINTS or \
                getattr(self, "_auto_choice", None) == "combination":
            return None
        try:
            return self.hierarchical()
        except NotImplementedError as exc:
            self._auto_choice = "combination"
            self.auto_validation = {"enabled": False, "reason": str(exc)}
            return None

    def _auto_validation_launch(self, torch, hp, x_new, f, alpha, device, host_rows):
        """Enqueue the validation of the hierarchical form for THIS call: both kernels evaluate
        a sample of the batch (up to AUTO_SAMPLE rows spread over it) with the caller's nodal
        values; max|hier - exact| / max|exact| goes to a pinned host scalar.  Returns
        (pinned scalar, event recorded after its copy)."""
        # The sample runs on a SIDE stream, enqueued before the batch: its few blocks take some SMs
        # for a fraction of a millisecond while the batch kernel (dynamic tile scheduling) already
        # runs on the others, instead of the batch waiting for two latency-bound small launches.
        main = torch.cuda.current_stream(device)
        sides = self.__dict__.setdefault("_auto_streams", {})
        pair = sides.get(device.index)
        if pair is None:
            pair = sides[device.index] = (torch.cuda.Stream(device), torch.cuda.Stream(device))
        side, side2 = pair if AUTO_SIDE_STREAM else (main, main)
        ready = torch.cuda.Event()
        ready.record(main)                   # nodal values, surpluses and points are ready here
        side.wait_event(ready)
        step = max(1, host_rows // AUTO_SAMPLE)
        with torch.cuda.stream(side):
            xs = x_new[::step][:AUTO_SAMPLE]
            if isinstance(xs, np.ndarray):
                xs = np.ascontiguousarray(xs[:, :self.N])
            xs = _to_device(torch, xs, device)
            if xs.shape[1] > self.N:
                xs = xs[:, :self.N].contiguous()
            # the two sample evaluations start at once, each on its own stream (a kernel queued
            # behind the other one would find every SM taken by the batch kernel and run last)
            have_xs = torch.cuda.Event()
            have_xs.record(side)
            side2.wait_event(have_xs)
            with torch.cuda.stream(side2):
                fast = hp._dev(device.index)[0].eval(xs, alpha)
                have_fast = torch.cuda.Event()
                have_fast.record(side2)
            exact = self.device_plan(device.index).eval(xs, f)
            side.wait_event(have_fast)
            if side2 is not side:
                fast.record_stream(side)         # allocated on side2, read by the comparison on side
            # NaN where the reference is NaN (Lagrange on a knot) counts as agreement; NaN in one
            # of the two makes the deviation NaN and fails the validation (sgm_max_rel_deviation)
            dev = _max_rel_deviation(torch, fast, exact)
            # one pinned scalar per device: the side stream serialises the calls that use it
            hosts = self.__dict__.setdefault("_auto_hosts", {})
            host = hosts.get(device.index)
            if host is None:
                host = hosts[device.index] = torch.empty(1, dtype=torch.float64).pin_memory()
            host.copy_(dev.reshape(1), non_blocking=True)
            done = torch.cuda.Event()
            done.record(side)
            # tensors of this block belong to the side stream's allocator pool; the inputs
            # (x_new, f, alpha) outlive the sample because the verdict is always read (a host
            # synchronisation on `done`) before interpolate() returns
        return host, done, int(xs.shape[0])

    def _auto_verdict(self, host, done, n_sample):
        done.synchronize()                   # waits for the small validation launches only
        dev = float(host[0])
        ok = bool(np.isfinite(dev)) and dev <= AUTO_TOL
        self.auto_validation = {"enabled": ok, "max_rel_deviation": dev, "tolerance": AUTO_TOL,
                                "sample_points": n_sample}
        self._auto_choice = "hierarchical" if ok else "combination"     # a failure is sticky
        return ok

    def interpolate(self, x_new, f_on_SG, method=None):
        """Evaluate the interpolant at the rows of ``x_new`` (only the first N columns are
        read) for the nodal values ``f_on_SG`` (num_nodes, dim_f).

        numpy in -> numpy out (``np.squeeze``d like the reference, :279); CUDA torch tensors
        in -> CUDA tensor out without host round trips.  One fused kernel launch.

        ``method``:
          * ``"combination"`` -- the reference's combination formula, bit for bit
            (sparse_grid_interpolant.py:268-279 with the operators' own arithmetic order);
          * ``"hierarchical"`` -- the work-reduced hierarchical-surplus form (pw-linear,
            pw-quadratic, Lagrange; nested knots; hierarchical.py): the same function with ~10x
            fewer gathers, agreeing with the reference to about the reference's own rounding
            error;
          * ``"auto"`` (the default, ``DEFAULT_METHOD``) -- the hierarchical form when it exists,
            the batch is large and THIS call's validation passes: a sample of the batch is
            evaluated by both kernels with the caller's nodal values and must agree to
            ``AUTO_TOL`` = 2.5e-13 (4x inside the 1e-12 contract; the deviation is the
            reference's own cancellation error, which depends on the data -- config 5 with
            smooth values reaches 1e-11 and therefore stays on the exact kernel).  Otherwise
            the exact kernel.  The last outcome is in ``auto_validation``."""
        method = DEFAULT_METHOD if method is None else method
        if method not in ("auto", "combination", "hierarchical"):
            raise ValueError("method must be 'auto', 'combination' or 'hierarchical'")
        from ._device import torch_cuda
        torch = torch_cuda()
        self._check_operator_constraints()
        as_torch = isinstance(x_new, torch.Tensor)
        if as_torch:
            device = x_new.device if x_new.is_cuda else torch.device("cuda", torch.cuda.current_device())
        else:
            device = torch.device("cuda", torch.cuda.current_device())
        n_used = getattr(self, "_n_used", None)      # columns of x_new the interpolant reads
        if n_used is None:                           # (a (T, N) reduction: once, not per call --
            act = (self._num_nodes > 1).any(axis=0)  #  config 5: 10 ms of host time)
            n_used = self._n_used = int(act.nonzero()[0].max() + 1) if act.any() else 0
        if not hasattr(f_on_SG, "shape"):
            f_on_SG = np.asarray(f_on_SG, dtype=np.float64)
        f_shape = tuple(f_on_SG.shape)
        if len(f_shape) != 2:
            raise IndexError("f_on_SG must be 2-D (num_nodes, dim_f)")     # reference: tuple index out of range
        if f_shape[0] < self.num_nodes:
            raise IndexError(f"f_on_SG has {f_shape[0]} rows, the sparse grid {self.num_nodes}")
        host_rows = x_new.shape[0] if hasattr(x_new, "shape") and len(x_new.shape) == 2 else 0
        if host_rows and x_new.shape[1] < n_used:
            raise IndexError(f"x_new has {x_new.shape[1]} columns, the interpolant uses {n_used}")
        pinned = as_torch and not x_new.is_cuda and x_new.is_pinned() and \
            x_new.dtype == torch.float64 and x_new.dim() == 2 and x_new.shape[1] >= self.N
        big_numpy = (not as_torch) and isinstance(x_new, np.ndarray) and x_new.ndim == 2 and \
            x_new.dtype == np.float64 and x_new.shape[1] >= self.N
        pipelined = (pinned or big_numpy) and host_rows >= _H2D_MIN_ROWS
        pipe = None
        if pipelined:
            # host points: the transfer of the first chunk starts NOW, on the copy stream, and runs
            # under the upload of the nodal values, the hierarchisation and the validation sample
            entry = torch.cuda.Event()
            entry.record(torch.cuda.current_stream(device))
            pipe = self._pipeline(torch, x_new, f_shape[1], device, after=entry)
            next(pipe)
        f = _to_device(torch, f_on_SG, device)
        exact_plan = self.device_plan(device.index)
        plan, f_eval, pending = exact_plan, f, None
        hp = self.hierarchical() if method == "hierarchical" else \
            (self._auto_candidate(host_rows) if method == "auto" else None)
        if hp is not None:
            try:
                plan = hp._dev(device.index)[0]
                f_eval = hp.surpluses(f[:hp.M])      # nodal values -> hierarchical surpluses (GPU)
                if method == "auto":
                    pending = self._auto_validation_launch(torch, hp, x_new, f, f_eval, device, host_rows)
            except OverflowError as exc:
                # tables of the work-reduced form do not fit the device's packed layouts / shared
                # memory (very deep 1-D levels): "auto" keeps the exact kernels
                if method != "auto":
                    raise
                self._auto_choice = "combination"
                self.auto_validation = {"enabled": False, "reason": str(exc)}
                hp, plan, f_eval, pending = None, exact_plan, f, None
        if pipelined:
            # optimistic: the chunks are enqueued with the fast plan while the validation sample
            # runs on its side stream; a failed verdict (rare, and sticky for the interpolant)
            # repeats the pipeline with the exact plan
            if pending is not None and not AUTO_OPTIMISTIC_PIPELINE and not self._auto_verdict(*pending):
                plan, f_eval, pending = exact_plan, f, None
            out = pipe.send((plan, f_eval))
            if pending is not None and AUTO_OPTIMISTIC_PIPELINE and not self._auto_verdict(*pending):
                out = self._interpolate_pipelined(torch, exact_plan, f, x_new, device)
            pending = None
        else:
            if not as_torch and host_rows and x_new.shape[1] > max(self.N, 1):
                x_new = x_new[:, :self.N]        # only the first N columns travel to the GPU
            x = _to_device(torch, x_new, device)
            if x.dim() == 1:
                x = x.reshape(-1, 1)
            if x.shape[1] < n_used:
                raise IndexError(f"x_new has {x.shape[1]} columns, the interpolant uses {n_used}")
            if x.shape[1] < self.N:      # unused trailing directions: pad so that ldx >= N
                x = torch.nn.functional.pad(x, (0, self.N - x.shape[1]))
            out = plan.eval(x, f_eval)           # enqueued before the verdict is read: it overlaps
            if pending is not None and not self._auto_verdict(*pending):
                out = exact_plan.eval(x, f, out)             # validation failed: the exact kernel
        flags = 0
        if self._kind == _lib.KIND_QUADRATIC:
            flags = plan.take_flags()
            if hp is not None:               # the validation sample went through the exact plan
                flags |= exact_plan.take_flags()
        if flags & _lib.FLAG_QUADRATIC_RIGHT:
            raise IndexError("pw-quadratic interpolation: a point lies right of the last knot "
                             "(or is NaN); the reference raises IndexError here "
                             "(tp_inteprolants.py:131-133)")
        if as_torch:
            return out.squeeze()
        return np.squeeze(out.cpu().numpy())


# ``interpolate(..., method=None)`` uses this; "auto" = fastest form that is validated to meet
# the 1e-12 contract, "combination" = always the reference's formula bit for bit.
DEFAULT_METHOD = "auto"
AUTO_MIN_POINTS = 8192      # smaller batches stay on the exact kernels
AUTO_SAMPLE = 256           # rows of the first large batch evaluated by both kernels
AUTO_TOL = 2.5e-13          # max |hier - exact| / max |exact| that enables the form
AUTO_SIDE_STREAM = True     # the validation sample runs on a side stream, next to the batch kernel
AUTO_OPTIMISTIC_PIPELINE = True   # host inputs: enqueue the chunks before the verdict is read


# rows per host->device chunk of the pipelined path (tools/exp_e2e_ab.py, config 2 from pinned
# memory): a first chunk of 113664 rows (a whole number of waves of the scalar kernels; nothing
# overlaps its transfer) and steady chunks of twice that -- the transfer of a chunk (0.5 ms) then
# hides under the evaluation of the previous one (0.9 ms) with half the per-chunk launches of the
# 113664-row pipeline: 5.43 -> 5.18 ms per 10^6 points end to end; larger chunks lose (5.4-5.6 ms)
_H2D_CHUNK = 227328
_H2D_FIRST = 113664          # rows of the first chunk
_H2D_MIN_ROWS = 4 * 113664    # smaller host batches are copied in one piece


_STAGING_THREADS = 4


def _parallel_copy(dst, src):
    """dst[...] = src for the pageable -> pinned staging copy of the numpy path, row blocks on a
    few short-lived threads (numpy releases the GIL inside the copy; one thread moves ~10 GB/s,
    which made the staging copy the bottleneck of the numpy-in / numpy-out call).  No persistent
    pool: ``sample_on_SG(n_parallel > 1)`` forks, and forking a process with live threads is
    discouraged."""
    import threading
    rows = dst.shape[0]
    if rows < 8192 or _STAGING_THREADS <= 1:
        np.copyto(dst, src, casting="same_kind")
        return
    step = -(-rows // _STAGING_THREADS)
    threads = [threading.Thread(target=np.copyto, args=(dst[a:a + step], src[a:a + step], "same_kind"))
               for a in range(0, rows, step)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()


def _max_rel_deviation(torch, fast, exact):
    """Device scalar max|fast - exact| / max|exact| (one launch; NaN rules in include/sgm_b200.h)."""
    out = torch.empty(1, dtype=torch.float64, device=exact.device)
    fast, exact = fast.contiguous(), exact.contiguous()
    _lib.check(_lib.load().sgm_max_rel_deviation(
        fast.data_ptr(), exact.data_ptr(), exact.numel(), out.data_ptr(),
        torch.cuda.current_stream(exact.device).cuda_stream))
    return out


def _to_device(torch, a, device):
    if isinstance(a, torch.Tensor):
        t = a.to(device=device, dtype=torch.float64, non_blocking=True)
    else:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)
    if t.dim() == 2 and t.stride(1) != 1:
        t = t.contiguous()
    return t
