"""BFGSSolver and line_search_wolfe_conditions with the reference's signatures, executed by the
persistent sm_100a solve kernel through the C-ABI (include/davo_b200.h).

Reference: deep_attention_visual_odometry/autograd_solvers/bfgs_solver.py:26-303 and
autograd_solvers/line_search/wolfe_conditions.py:23-253.
"""
from __future__ import annotations

import ctypes
import warnings
from typing import NamedTuple

import torch
from torch.nn import Module

from . import _lib
from .objectives import CalibrationObjective


class SolveInfo(NamedTuple):
    """Per-problem outputs the reference keeps internal (north_star: 'cost and convergence mask out')."""

    parameters: torch.Tensor  # (B..) x n
    cost: torch.Tensor        # (B..)   objective at the returned parameters
    converged: torch.Tensor   # (B..)   bool, cost <= error_threshold
    iterations: torch.Tensor  # (B..)   int32, accepted steps (line searches taken part in)
    evaluations: torch.Tensor # (B..)   int32, objective evaluations the reference would have made
    reason: torch.Tensor      # (B..)   int32, 0 threshold / 1 step size / 2 iteration cap / 3 NaN cost


class SolveBuffers(NamedTuple):
    """Device-resident outputs of one solve launch (all rows are always written)."""

    x: torch.Tensor            # [B,n]
    cost: torch.Tensor         # [B]
    converged: torch.Tensor    # [B] uint8
    iterations: torch.Tensor   # [B] int32
    evaluations: torch.Tensor  # [B] int32
    reason: torch.Tensor       # [B] int32
    workspace: torch.Tensor    # DAVO_WORKSPACE_BYTES of scratch (work-queue counter)
    slab: torch.Tensor = None  # the one allocation the six outputs are views of (allocate()), else None

    @staticmethod
    def _sections(B, n, dtype):
        """(name, dtype, shape, byte offset) of the six outputs inside one slab, 256-byte aligned; total bytes."""
        item = torch.empty((), dtype=dtype).element_size()
        spec = (("x", dtype, (B, n), B * n * item), ("cost", dtype, (B,), B * item),
                ("converged", torch.uint8, (B,), B), ("iterations", torch.int32, (B,), 4 * B),
                ("evaluations", torch.int32, (B,), 4 * B), ("reason", torch.int32, (B,), 4 * B))
        out, off = [], 0
        for name, dt, shape, nbytes in spec:
            out.append((name, dt, shape, off, nbytes))
            off += (nbytes + 255) & ~255
        return out, off

    @classmethod
    def views(cls, slab, B, n, dtype):
        """The six outputs as views of `slab` (device or host), in field order."""
        sections, _ = cls._sections(B, n, dtype)
        return [slab[off:off + nbytes].view(dt).reshape(shape) for _, dt, shape, off, nbytes in sections]

    @classmethod
    def allocate(cls, B, n, dtype, device) -> "SolveBuffers":
        """One slab for the six outputs: a host caller's results then come back in ONE device-to-host copy."""
        _, total = cls._sections(B, n, dtype)
        slab = torch.empty(max(total, 1), dtype=torch.uint8, device=device)
        return cls(*cls.views(slab, B, n, dtype), torch.empty(_lib.WORKSPACE_BYTES, dtype=torch.uint8, device=device),
                   slab)


class _DifferentiableSolve(torch.autograd.Function):
    """x_out = solve(x0) with d loss / d x0 from the stored-iterate reverse sweep (csrc/solver_train.cuh).

    Forward: davo_solve_training on the objective's device in the parameters' precision; when a gradient is wanted
    the accepted iterates (x_k, g_k, alpha_k) are recorded and compacted to sum(steps) rows.  Backward:
    davo_solve_backward in float64 (a float32 trajectory and problem set are up-cast)."""

    @staticmethod
    def forward(ctx, parameters, obj, cfg, record, data):
        device = obj.device
        n, B = obj.n, obj.B
        batch_shape = parameters.shape[:-1]
        data0 = obj.data0  # stages a lazily staged (host-input) problem set
        K = int(cfg["iterations"])
        with torch.cuda.device(device):
            x0 = parameters.detach().to(device=device, dtype=obj.dtype).reshape(B, n).contiguous()
            buf = SolveBuffers.allocate(B, n, obj.dtype, device)
            traj = (None, None, None, None)
            if record:
                item = torch.empty((), dtype=obj.dtype).element_size()
                need = B * max(K, 1) * (2 * n + 1) * item
                if need > cfg["budget"]:
                    raise _lib.DavoError(
                        f"recording {B} problems x {K} iterations x {n} parameters needs {need / 2**30:.1f} GiB; lower "
                        "training_iterations or raise BFGSSolver.trajectory_budget_bytes")
                traj = (torch.empty(B, max(K, 1), n, dtype=obj.dtype, device=device),
                        torch.empty(B, max(K, 1), n, dtype=obj.dtype, device=device),
                        torch.empty(B, max(K, 1), dtype=obj.dtype, device=device),
                        torch.zeros(B, dtype=torch.int32, device=device))
            desc = obj.desc(iterations=K, strong=True, sufficient_decrease=cfg["sufficient_decrease"],
                            curvature=cfg["curvature"], error_threshold=cfg["error_threshold"],
                            minimum_step=cfg["minimum_step"], zoom_interpolation=cfg["zoom_interpolation"])
            tdesc = _lib.TrainingDesc(max(K, 1), int(cfg["return_second_last"]), float(cfg["drop_path_p"]),
                                      int(cfg["seed"]), 0.0)
            st = _lib.lib().davo_solve_training(
                ctypes.byref(desc), ctypes.byref(tdesc), _lib.ptr(data0), _lib.ptr(obj.data1), _lib.ptr(obj.weights),
                _lib.ptr(x0), _lib.ptr(buf.x), _lib.ptr(buf.cost), _lib.ptr(buf.converged), _lib.ptr(buf.iterations),
                _lib.ptr(buf.evaluations), _lib.ptr(buf.reason), _lib.ptr(traj[0]), _lib.ptr(traj[1]),
                _lib.ptr(traj[2]), _lib.ptr(traj[3]), _lib.ptr(buf.workspace), _lib.stream_ptr())
            _lib.check(st, "davo_solve_training")
            if record:
                # keep only the recorded rows: [sum(steps), n] instead of [B, K, n]
                steps = traj[3].long()
                keep = torch.arange(max(K, 1), device=device).unsqueeze(0) < steps.unsqueeze(1)
                ctx.traj = (traj[0][keep].double(), traj[1][keep].double(), traj[2][keep].double(), traj[3],
                            torch.cumsum(steps, 0) - steps)
                ctx.obj, ctx.cfg = obj, cfg
        ctx.in_device, ctx.in_dtype, ctx.in_shape = parameters.device, parameters.dtype, parameters.shape
        if data is not None:
            ctx.data_device, ctx.data_dtype, ctx.data_shape = data.device, data.dtype, data.shape
        out_dev = parameters.device
        outs = (buf.x.reshape(parameters.shape).to(device=out_dev, dtype=parameters.dtype),
                buf.cost.reshape(batch_shape).to(out_dev), buf.converged.reshape(batch_shape).to(out_dev).bool(),
                buf.iterations.reshape(batch_shape).to(out_dev), buf.evaluations.reshape(batch_shape).to(out_dev),
                buf.reason.reshape(batch_shape).to(out_dev))
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, grad_x, *unused):
        if not hasattr(ctx, "traj"):
            return None, None, None, None, None
        obj, cfg = ctx.obj.as_float64(), ctx.cfg
        device, n, B = obj.device, obj.n, obj.B
        tx, tg, ta, tlen, offsets = ctx.traj
        with torch.cuda.device(device):
            g_out = grad_x.detach().to(device=device, dtype=torch.float64).reshape(B, n).contiguous()
            g_x0 = torch.empty_like(g_out)
            want_data = ctx.needs_input_grad[4]
            # observations: [N, 2] per problem (DISTORT10) or [V, N, 2] (JOINT, ANGLE_BA)
            g_data = torch.empty(B, obj.views * obj.N * 2, dtype=torch.float64, device=device) if want_data else None
            rows = int(tx.shape[0])
            scratch = torch.empty(max(rows, 1) * (n * n + n), dtype=torch.float64, device=device)
            workspace = torch.empty(_lib.WORKSPACE_BYTES, dtype=torch.uint8, device=device)
            desc = obj.desc(iterations=int(cfg["iterations"]), strong=True)
            tdesc = _lib.TrainingDesc(max(int(cfg["iterations"]), 1), int(cfg["return_second_last"]), 0.0, 0, 0.0)
            st = _lib.lib().davo_solve_backward(
                ctypes.byref(desc), ctypes.byref(tdesc), _lib.ptr(obj.data0), _lib.ptr(obj.data1),
                _lib.ptr(obj.weights), _lib.ptr(tx), _lib.ptr(tg), _lib.ptr(ta), _lib.ptr(tlen), _lib.ptr(offsets),
                _lib.ptr(offsets), _lib.ptr(scratch), _lib.ptr(g_out), _lib.ptr(g_x0), _lib.ptr(g_data),
                _lib.ptr(workspace), _lib.stream_ptr())
            _lib.check(st, "davo_solve_backward")
        if want_data:
            g_data = g_data.reshape(ctx.data_shape).to(device=ctx.data_device, dtype=ctx.data_dtype)
        g_par = g_x0.reshape(ctx.in_shape).to(device=ctx.in_device, dtype=ctx.in_dtype) if ctx.needs_input_grad[0] else None
        return g_par, None, None, None, g_data


class PendingSolve:
    """Handle returned by BFGSSolver.submit: the solve (and, for a host caller, the copies of its results into pinned
    host memory) has been queued; `result()` waits for it and returns what `forward` would have returned."""

    def __init__(self, finalize=None, value=None):
        self._finalize, self._value = finalize, value

    def result(self):
        if self._finalize is not None:
            self._value = self._finalize()
            self._finalize = None
        return self._value


_STREAMS: dict = {}


def _side_streams(device):
    """One copy stream and two compute streams per device, created once (stream creation is not free and the
    caching allocator tracks cross-stream use per stream)."""
    key = (device.type, device.index)
    if key not in _STREAMS:
        _STREAMS[key] = [torch.cuda.Stream(device=device) for _ in range(3)]
    return _STREAMS[key]


def _require_descriptor(error_function) -> CalibrationObjective:
    if not isinstance(error_function, CalibrationObjective):
        raise TypeError(
            "the B200 solver runs the objective inside a CUDA kernel and needs a CalibrationObjective "
            "descriptor (DistortionObjective, JointPoseObjective, AngleDistanceObjective, AnalyticObjective), not an "
            "arbitrary Python "
            f"callable (got {type(error_function).__name__}); there is no CPU / autograd fallback")
    return error_function


class BFGSSolver(Module):
    """Drop-in for the reference's BFGSSolver (autograd_solvers/bfgs_solver.py:26-78): same constructor
    arguments and defaults, same ``forward(parameters, error_function) -> parameters`` contract, batch
    dimensions ``(B..)`` of any rank.  ``error_function`` must be a CalibrationObjective descriptor.

    Training mode (`self.training`) uses the training thresholds, drop-path and return_second_last exactly as
    bfgs_solver.py:88-93,122-125,196-212; parameters that require grad make the solve differentiable (the
    reference's create_graph=True, :85,133-135): the forward keeps the accepted iterates and `backward` runs the
    reverse sweep through them in a CUDA kernel (csrc/solver_train.cuh).  Drop-path draws come from a counter-based
    device generator seeded from torch's default generator (torch.manual_seed makes them reproducible); they are
    not torch.rand's numbers.
    """

    #: refuse to record trajectories larger than this (bytes): B x training_iterations x (2n+1) values
    trajectory_budget_bytes = 32 << 30

    #: opt-in line-search variant: the zoom step interpolates phi' at the bracket ends
    #: (utils/func_interpolate_alpha.py, as the older solvers/line_search_strong_wolfe_conditions.py:147-155 does)
    #: instead of bisecting; set on an instance.  Not a constructor argument: the constructor is the reference's.
    zoom_interpolation = False

    #: set to a dict to receive the streamed path's CUDA events (tools/e2e_timeline.py)
    _trace = None

    def __init__(self, sufficient_decrease: float = 1e-4, curvature: float = 0.9, error_threshold: float = 1e-4,
                 iterations: int = 1000, minimum_step: float = 1e-8, drop_path_p: float = 0.1,
                 return_second_last: bool = False, training_iterations: int = None,
                 training_error_threshold: float = None):
        super().__init__()
        self.sufficient_decrease = float(sufficient_decrease)
        self.curvature = float(curvature)
        self.error_threshold = float(error_threshold)
        self.iterations = int(iterations)
        self.minimum_step = float(minimum_step)
        self.drop_path_p = float(drop_path_p)
        self.return_second_last = bool(return_second_last)
        self.training_iterations = int(training_iterations) if training_iterations is not None else self.iterations
        self.training_error_threshold = (float(training_error_threshold) if training_error_threshold is not None
                                         else self.error_threshold)

    def forward(self, parameters: torch.Tensor, error_function, return_info: bool = False, out=None):
        return self.submit(parameters, error_function, return_info=return_info, out=out).result()

    def submit(self, parameters: torch.Tensor, error_function, return_info: bool = False, out=None) -> PendingSolve:
        """`forward` without the final wait: queue the solve (for a host caller also the copies of the results into
        pinned host memory) and return a PendingSolve; `forward(...)` is `submit(...).result()`.  Submitting the next
        batch before collecting this one lets its uploads start early, but measured on this pool that does not pay
        (6.4 to 13 ms per 64K batch from run to run against a steady 7.3 ms one at a time): bench.py times forward()."""
        obj = _require_descriptor(error_function)
        if self.training:
            error_threshold, iterations = self.training_error_threshold, self.training_iterations
        else:
            error_threshold, iterations = self.error_threshold, self.iterations
        if not 0.0 < self.sufficient_decrease < self.curvature < 1.0:  # wolfe_conditions.py:65-69
            warnings.warn(f"Line search conditions should satisfy 0 < c1 < c2 < 1. "
                          f"Got c1={self.sufficient_decrease} and c2={self.curvature}")
        batch_shape, n = parameters.shape[:-1], parameters.shape[-1]
        if n != obj.n:
            raise ValueError(f"parameters have {n} columns, the objective expects {obj.n}")
        if tuple(batch_shape) != obj.batch_shape and parameters.numel() // n != obj.B:
            raise ValueError(f"parameters batch {tuple(batch_shape)} does not match the objective's {obj.batch_shape}")
        out_dev = parameters.device
        differentiable = torch.is_grad_enabled() and (parameters.requires_grad or
                                                      getattr(obj, "differentiable_data", None) is not None)
        if differentiable or (self.training and (self.drop_path_p > 0.0 or self.return_second_last)):
            return PendingSolve(value=self._forward_training(parameters, obj, error_threshold, iterations,
                                                             differentiable, return_info))
        streamed = not (getattr(obj, "is_staged", True) or self.zoom_interpolation)
        if not streamed:
            x0 = parameters.detach().to(device=obj.device, dtype=obj.dtype, non_blocking=True).reshape(obj.B, n)
            buf = self.solve_into(x0.contiguous(), obj, error_threshold=error_threshold, iterations=iterations,
                                  out=out)
            done_events = None
        else:  # host-resident problem set: overlap the copies with staging + solve, chunk by chunk
            buf, done_events = self._solve_streamed(parameters.detach().reshape(obj.B, n), obj, error_threshold,
                                                    iterations, out)
        if out_dev.type == "cpu":
            # host caller: every output goes device -> pinned host memory, then ONE wait in result() (six blocking
            # copies would each pay a synchronisation).  After a streamed solve the copies run on the solver's result
            # stream behind the chunks' events, so the calling stream is not involved at all.
            if return_info and buf.slab is not None:   # the solver's own slab: one copy brings all six outputs
                wanted = (buf.slab,)
                host_slab = torch.empty(buf.slab.shape, dtype=torch.uint8, pin_memory=True)
                copies = ((host_slab, buf.slab),)
                host = SolveBuffers.views(host_slab, obj.B, n, obj.dtype)
            else:
                wanted = (buf.x, buf.cost, buf.converged, buf.iterations, buf.evaluations, buf.reason) if return_info else (buf.x,)
                host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in wanted]
                copies = tuple(zip(host, wanted))
            with torch.cuda.device(obj.device):
                # the copies run on the calling stream, which already waits for every chunk (a separate result stream
                # behind the chunks' events measured 0.05 ms slower per 64K batch)
                res_stream = torch.cuda.current_stream(obj.device)
                with torch.cuda.stream(res_stream):
                    for h, t in copies:
                        h.copy_(t, non_blocking=True)
                    finished = torch.cuda.Event()
                    finished.record(res_stream)
                for t in wanted:
                    t.record_stream(res_stream)
            shape, dtype = parameters.shape, parameters.dtype

            def finalize(host=host, finished=finished, keep=buf):
                finished.synchronize()
                result = host[0].reshape(shape).to(dtype)
                if not return_info:
                    return result
                shaped = [h.reshape(batch_shape) for h in host[1:]]
                return SolveInfo(result, shaped[0], shaped[1].bool(), shaped[2], shaped[3], shaped[4])
            return PendingSolve(finalize=finalize)
        result = buf.x.reshape(parameters.shape).to(device=out_dev, dtype=parameters.dtype)
        if not return_info:
            return PendingSolve(value=result)
        back = lambda t: t.reshape(batch_shape).to(out_dev)
        return PendingSolve(value=SolveInfo(result, back(buf.cost), back(buf.converged).bool(), back(buf.iterations),
                                            back(buf.evaluations), back(buf.reason)))

    def _forward_training(self, parameters, obj, error_threshold, iterations, differentiable, return_info):
        """Training-mode / differentiable solve (davo_solve_training, davo_solve_backward)."""
        seed = 0
        drop_p = self.drop_path_p if self.training else 0.0
        if drop_p > 0.0:  # one draw from torch's default generator per call: torch.manual_seed reproduces a run
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        cfg = dict(error_threshold=error_threshold, iterations=iterations, drop_path_p=drop_p, seed=seed,
                   return_second_last=bool(self.training and self.return_second_last),
                   sufficient_decrease=self.sufficient_decrease, curvature=self.curvature,
                   minimum_step=self.minimum_step, budget=self.trajectory_budget_bytes,
                   zoom_interpolation=bool(self.zoom_interpolation))
        x, cost, converged, iters, fevals, reason = _DifferentiableSolve.apply(
            parameters, obj, cfg, differentiable, obj.differentiable_data if differentiable else None)
        if not return_info:
            return x
        return SolveInfo(x, cost, converged, iters, fevals, reason)

    #: problems per chunk of the streamed (host-input) path; each chunk is one H2D copy + stage + solve
    stream_chunk = 4096

    #: the last chunk is halved repeatedly down to this many problems
    stream_tail = 1024

    @classmethod
    def _chunk_spans(cls, B: int, chunk: int):
        """[lo, hi) spans of `chunk` problems; the last chunk is cut into 1/2, 1/4, ... down to `stream_tail`
        problems so that the compute left over once the final copy has landed (the only part the copies cannot
        hide) is short."""
        spans = [(lo, min(lo + chunk, B)) for lo in range(0, B, chunk)]
        if len(spans) > 1:
            lo, hi = spans.pop()
            while hi - lo >= 2 * max(1, cls.stream_tail):
                mid = lo + (hi - lo) // 2
                spans.append((lo, mid))
                lo = mid
            spans.append((lo, hi))
        return spans

    def _solve_streamed(self, x0_host: torch.Tensor, obj, error_threshold, iterations, out):
        """Host inputs -> device results with the compute hidden behind the copies.

        The batch is cut into chunks of `stream_chunk` problems.  ONE copy stream moves every chunk's points,
        observations and start parameters host -> device back to back, so the copy engine never waits for
        compute (at 64K x 256 the 338 MB of raw inputs take longer over PCIe than the solve takes on the SMs);
        an event per chunk releases its staging (davo_stage_matches) and solve (davo_solve_calibration) on
        one of two compute streams, alternating so that the next chunk's persistent grid fills the SMs as the
        previous chunk's tail drains.  The staged matches stay resident in `obj` afterwards.  Nothing here
        synchronises the host."""
        device = obj.device
        B, n, N = obj.B, obj.n, obj.N
        spans = self._chunk_spans(B, max(1, int(self.stream_chunk)))
        with torch.cuda.device(device):
            buf = out if out is not None else SolveBuffers.allocate(B, n, obj.dtype, device)
            new = lambda *shape: torch.empty(*shape, dtype=obj.dtype, device=device)
            staged, d_pts, d_obs, x0_dev = new(B, N, 4), new(B, N, 3), new(B, N, 2), new(B, n)
            d_pose = new(B, 6) if obj._raw[2] is not None else None
            wdev = new(B, N) if obj._raw[3] is not None else None
            workspaces = torch.empty(len(spans), _lib.WORKSPACE_BYTES, dtype=torch.uint8, device=device)
            main = torch.cuda.current_stream()
            new_event = lambda: torch.cuda.Event(enable_timing=self._trace is not None)
            start = new_event()
            start.record(main)
            copy, *compute = _side_streams(device)[:3]
            # the side streams follow the calling stream only when something on it can matter to them: start
            # parameters that live on the device.  Host inputs have no producer on any stream, and the buffers below
            # are fresh; without this dependency a second batch's uploads start while the first is still solving.
            follow_main = x0_host.is_cuda
            if follow_main:
                copy.wait_event(start)
            # every copy is queued before the first compute launch: the copy engine then runs back to back from the
            # first byte (the copies, not the SMs, bound this path) and the host's per-chunk launch work below never
            # sits between two copies.  (A second copy stream for the observations was tried: the two streams share
            # the link and every chunk lands later: 10.5 ms against 7.3 ms.)
            landed_events = []
            with torch.cuda.stream(copy):
                # the small per-problem inputs (start parameters, poses) go first, whole: 40 + 24 bytes per problem
                # against 5 KB of matches — as per-chunk copies they were 2 x 18 small transfers between the large ones
                x0_dev.copy_(x0_host, non_blocking=True)
                obj.upload_poses(d_pose)
                for lo, hi in spans:
                    obj.upload_rows(lo, hi, d_pts, d_obs, d_pose, wdev, poses=False)
                    landed = new_event()
                    landed.record(copy)
                    landed_events.append(landed)
            if follow_main:
                for cs in compute:
                    cs.wait_event(start)
            # The per-chunk launches go straight to the C-ABI with base pointer + offset and the compute stream's
            # handle: no tensor slicing, no stream context, no descriptor objects per chunk.  The host's time per batch
            # is what bounds this path once the copies are queued (it was ~6.5 ms of Python for 20 chunks, more than the
            # 6.1 ms the copies take; now ~2 ms).
            lib = _lib.lib()
            es = staged.element_size()
            vp = ctypes.c_void_p
            P = lambda t: 0 if t is None else t.data_ptr()
            p_pts, p_obs, p_pose, p_w, p_staged, p_x0 = P(d_pts), P(d_obs), P(d_pose), P(wdev), P(staged), P(x0_dev)
            p_x, p_cost, p_conv, p_it, p_fe, p_re, p_ws = (P(buf.x), P(buf.cost), P(buf.converged), P(buf.iterations),
                                                          P(buf.evaluations), P(buf.reason), P(workspaces))
            off = lambda base, count: vp(base + count) if base else None
            descs = {}
            done_events = []
            for k, (lo, hi) in enumerate(spans):
                cs = compute[k % 2]
                cs.wait_event(landed_events[k])
                rows = hi - lo
                if rows not in descs:
                    descs[rows] = (obj.desc(B=rows),
                                   obj.desc(B=rows, iterations=iterations, strong=True,
                                            sufficient_decrease=self.sufficient_decrease, curvature=self.curvature,
                                            error_threshold=error_threshold, minimum_step=self.minimum_step))
                d_stage, d_solve = descs[rows]
                stream = vp(cs.cuda_stream)
                st = lib.davo_stage_matches(ctypes.byref(d_stage), off(p_pts, lo * N * 3 * es), off(p_obs, lo * N * 2 * es),
                                            off(p_pose, lo * 6 * es), off(p_staged, lo * N * 4 * es), stream)
                _lib.check(st, "davo_stage_matches")
                st = lib.davo_solve_calibration(
                    ctypes.byref(d_solve), off(p_staged, lo * N * 4 * es), None, off(p_w, lo * N * es),
                    off(p_x0, lo * n * es), off(p_x, lo * n * es), off(p_cost, lo * es), off(p_conv, lo),
                    off(p_it, lo * 4), off(p_fe, lo * 4), off(p_re, lo * 4), vp(p_ws + k * _lib.WORKSPACE_BYTES), stream)
                _lib.check(st, "davo_solve_calibration")
                done = new_event()
                done.record(cs)
                done_events.append(done)
            for cs in compute:   # later work on the calling stream (e.g. obj.evaluate) sees the results
                main.wait_stream(cs)
            obj.data0, obj.weights = staged, wdev
            if self._trace is not None:  # tools/e2e_timeline.py: where the step's time goes
                self._trace.update(start=start, landed=landed_events, done=done_events, spans=spans)
            for t in (staged, d_pts, d_obs, x0_dev, workspaces, d_pose, wdev):
                if t is not None:
                    for s in (copy, *compute):
                        t.record_stream(s)
        return buf, done_events

    def solve_into(self, x0: torch.Tensor, obj: CalibrationObjective, *, error_threshold=None, iterations=None,
                   out: "SolveBuffers | None" = None) -> "SolveBuffers":
        """The launch itself: x0 [B,n] on the objective's device -> device-resident SolveBuffers.  Stream
        ordered, no host synchronisation; `out` lets a caller (e.g. the multi-GPU gather slab) own the
        output memory so the kernel writes straight into it."""
        device = _lib.require_cuda() if obj.device.type != "cuda" else obj.device
        B, n = obj.B, obj.n
        error_threshold = self.error_threshold if error_threshold is None else error_threshold
        iterations = self.iterations if iterations is None else iterations
        with torch.cuda.device(device):
            buf = out if out is not None else SolveBuffers.allocate(B, n, obj.dtype, device)
            data0 = obj.data0  # stages a lazily staged problem set (and its weights) before the descriptor is built
            desc = obj.desc(iterations=iterations, strong=True, sufficient_decrease=self.sufficient_decrease,
                            curvature=self.curvature, error_threshold=error_threshold,
                            minimum_step=self.minimum_step, zoom_interpolation=bool(self.zoom_interpolation))
            st = _lib.lib().davo_solve_calibration(
                ctypes.byref(desc), _lib.ptr(data0), _lib.ptr(obj.data1), _lib.ptr(obj.weights), _lib.ptr(x0),
                _lib.ptr(buf.x), _lib.ptr(buf.cost), _lib.ptr(buf.converged), _lib.ptr(buf.iterations),
                _lib.ptr(buf.evaluations), _lib.ptr(buf.reason), _lib.ptr(buf.workspace), _lib.stream_ptr())
        _lib.check(st, "davo_solve_calibration")
        return buf

    # ---- the two static helpers of the reference, on the GPU -----------------------------------------
    @staticmethod
    def scale_initial_inverse_hessian(step: torch.Tensor, delta_gradient: torch.Tensor) -> torch.Tensor:
        """bfgs_solver.py:217-233 (eq. 6.20): max(s.y / max(y.y, 1e-5), 1e-4), shape (B..) x 1."""
        device = _lib.require_cuda()
        n = step.shape[-1]
        dt = step.dtype
        s = step.detach().to(device).reshape(-1, n).contiguous()
        y = delta_gradient.detach().to(device=device, dtype=dt).reshape(-1, n).contiguous()
        out = torch.empty(s.shape[0], dtype=dt, device=device)
        st = _lib.lib().davo_bfgs_initial_scale(_lib.dtype_code(dt), s.shape[0], n, _lib.ptr(s), _lib.ptr(y),
                                                _lib.ptr(out), _lib.stream_ptr())
        _lib.check(st, "davo_bfgs_initial_scale")
        return out.reshape(step.shape[:-1] + (1,)).to(step.device)

    @staticmethod
    def update_inverse_hessian(inverse_hessian: torch.Tensor, step: torch.Tensor,
                               delta_gradient: torch.Tensor) -> torch.Tensor:
        """bfgs_solver.py:235-303 (eq. 6.17); non-positive curvature leaves H unchanged."""
        device = _lib.require_cuda()
        n = step.shape[-1]
        dt = inverse_hessian.dtype
        H = inverse_hessian.detach().to(device).reshape(-1, n, n).contiguous().clone()
        s = step.detach().to(device=device, dtype=dt).reshape(-1, n).contiguous()
        y = delta_gradient.detach().to(device=device, dtype=dt).reshape(-1, n).contiguous()
        st = _lib.lib().davo_bfgs_update(_lib.dtype_code(dt), H.shape[0], n, _lib.ptr(H), _lib.ptr(s), _lib.ptr(y),
                                         _lib.stream_ptr())
        _lib.check(st, "davo_bfgs_update")
        return H.reshape(inverse_hessian.shape).to(inverse_hessian.device)


def line_search_wolfe_conditions(parameters: torch.Tensor, search_direction: torch.Tensor, base_error: torch.Tensor,
                                 base_gradient: torch.Tensor, error_function, sufficient_decrease: float = 1e-4,
                                 curvature: float = 0.9, strong: bool = False, return_probes: bool = False,
                                 zoom_interpolation: bool = False):
    """Drop-in for autograd_solvers/line_search/wolfe_conditions.py:23-239: alpha of shape (B..).
    zoom_interpolation=True: the zoom step is interpolate_alpha(lo, hi, phi'(lo), phi'(hi)) instead of the midpoint."""
    obj = _require_descriptor(error_function)
    if not 0.0 < sufficient_decrease < curvature < 1.0:
        warnings.warn(f"Line search conditions should satisfy 0 < c1 < c2 < 1. "
                      f"Got c1={sufficient_decrease} and c2={curvature}")
    device = obj.device
    n = parameters.shape[-1]
    batch_shape = parameters.shape[:-1]
    prep = lambda t, shape: t.detach().to(device=device, dtype=obj.dtype).reshape(shape).contiguous()
    B = obj.B
    with torch.cuda.device(device):
        x, d, g = prep(parameters, (B, n)), prep(search_direction, (B, n)), prep(base_gradient, (B, n))
        f0 = prep(base_error, (B,))
        alpha = torch.empty(B, dtype=obj.dtype, device=device)
        probes = torch.empty(B, dtype=torch.int32, device=device)
        data0 = obj.data0  # stages a lazily staged problem set (and its weights) before the descriptor is built
        desc = obj.desc(strong=strong, sufficient_decrease=sufficient_decrease, curvature=curvature,
                        zoom_interpolation=zoom_interpolation)
        st = _lib.lib().davo_line_search(ctypes.byref(desc), _lib.ptr(data0), _lib.ptr(obj.data1),
                                         _lib.ptr(obj.weights), _lib.ptr(x), _lib.ptr(d), _lib.ptr(f0), _lib.ptr(g),
                                         _lib.ptr(alpha), _lib.ptr(probes), _lib.stream_ptr())
    _lib.check(st, "davo_line_search")
    alpha = alpha.reshape(batch_shape).to(device=parameters.device, dtype=parameters.dtype)
    if return_probes:
        return alpha, probes.reshape(batch_shape).to(parameters.device)
    return alpha


class _InterpolateAlpha(torch.autograd.Function):
    """utils/func_interpolate_alpha.py:5-79 on the GPU (davo_interpolate_alpha, davo_interpolate_alpha_backward)."""

    @staticmethod
    def forward(ctx, alpha_1, alpha_2, value_1, value_2):
        device = _lib.require_cuda()
        dt = alpha_1.dtype
        shape = torch.broadcast_shapes(alpha_1.shape, alpha_2.shape, value_1.shape, value_2.shape)
        ins = [t.detach().to(device=device, dtype=dt).expand(shape).contiguous() for t in (alpha_1, alpha_2, value_1, value_2)]
        out = torch.empty(shape, dtype=dt, device=device)
        with torch.cuda.device(device):
            st = _lib.lib().davo_interpolate_alpha(_lib.dtype_code(dt), out.numel(), *[_lib.ptr(t) for t in ins],
                                                   _lib.ptr(out), _lib.stream_ptr())
        _lib.check(st, "davo_interpolate_alpha")
        ctx.save_for_backward(*ins)
        ctx.shapes = [t.shape for t in (alpha_1, alpha_2, value_1, value_2)]
        ctx.in_device = alpha_1.device
        ctx.set_materialize_grads(False)
        return out.to(alpha_1.device)

    @staticmethod
    def backward(ctx, grad_output):
        if grad_output is None:
            return None, None, None, None
        ins = ctx.saved_tensors
        device, dt = ins[0].device, ins[0].dtype
        go = grad_output.detach().to(device=device, dtype=dt).expand(ins[0].shape).contiguous()
        grads = [torch.empty_like(ins[0]) if need else None for need in ctx.needs_input_grad]
        with torch.cuda.device(device):
            st = _lib.lib().davo_interpolate_alpha_backward(
                _lib.dtype_code(dt), go.numel(), *[_lib.ptr(t) for t in ins], _lib.ptr(go),
                *[_lib.ptr(g) for g in grads], _lib.stream_ptr())
        _lib.check(st, "davo_interpolate_alpha_backward")
        return tuple(None if g is None else g.sum_to_size(shape).to(ctx.in_device) for g, shape in zip(grads, ctx.shapes))


def interpolate_alpha(alpha_1: torch.Tensor, alpha_2: torch.Tensor, value_1: torch.Tensor,
                      value_2: torch.Tensor) -> torch.Tensor:
    """Drop-in for utils/func_interpolate_alpha.py:82-: the zero of the line through (alpha_1, value_1),
    (alpha_2, value_2), or the midpoint when it is degenerate or within 1e-3 of the interval's ends."""
    return _InterpolateAlpha.apply(alpha_1, alpha_2, value_1, value_2)
