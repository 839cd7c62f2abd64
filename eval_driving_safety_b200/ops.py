"""autograd.Function / functional layer over the C ABI (include/b2attack.h).

Host-side mirror of the operator interface the reference's model reaches:
``BuildCostVolume`` (upstream dsgn.layers.BuildCostVolume), ``F.grid_sample``
lifting, ``Conv3d``/``ConvTranspose3d`` + ``GroupNorm`` of the hourglass stacks
(all behind attack/DSGN/pgd_attack.py:308 forward and :336 backward) and
``ROIAlign`` (attack/Stereo-RCNN/stereo_rcnn.py:44-45, 110-141).

Every op runs on the CURRENT torch CUDA stream, takes fp32 CUDA tensors, and
raises RuntimeError on anything else: there is no CPU or PyTorch fallback.
Volumes are logical NCDHW tensors in ``torch.channels_last_3d`` memory format,
so they remain drop-in arguments for any stock torch op.
"""
import ctypes
import os
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import check, c_vp

CL3 = torch.channels_last_3d
CL2 = torch.channels_last

# conv implementation: 0 = tcgen05 (TF32 in, fp32 accumulate), 1 = fp32 SIMT (verification)
CONV_IMPL = int(os.environ.get("B2_CONV_IMPL", "0"))


# optionally round packed weights to nearest TF32 (measured: no visible effect on full-size parity, so off;
# it would also make the fp32 verification kernel see rounded weights)
TF32_ROUND_WEIGHTS = os.environ.get("B2_TF32_ROUND_WEIGHTS", "0") != "0"


def set_conv_impl(impl):
    """0 = the benchmarked path (tcgen05 TF32 3-D convs; 2-D convs: 3xTF32 forward, plain-TF32 data gradient);
    1 = verification mode: fp32 SIMT 3-D convs and the 3xTF32 split in the 2-D data gradients too."""
    global CONV_IMPL, CONV2D_SPLIT_BWD
    CONV_IMPL = int(impl)
    CONV2D_SPLIT_BWD = 1 if CONV_IMPL == 1 else int(os.environ.get("B2_CONV2D_SPLIT_BWD", "0"))


def _stream():
    return c_vp(torch.cuda.current_stream().cuda_stream)


# -- launch accounting / live kernel timing (bench.py) -------------------------
LAUNCH_COUNT = 0          # kernels of libb2attack.so launched so far
PROFILE_SPIN_CYCLES = int(os.environ.get("B2_PROFILE_SPIN", "200000"))   # ~0.1 ms device-side spin before each timed launch
_PROFILE = None           # name -> [(start_event, end_event, work)] while profiling


class _op:
    """Counts the kernels an entry point launches and, while ``profile()`` is active,
    brackets it with CUDA events on the launching (= current torch) stream.  ``work`` is
    the ALGORITHMIC byte (or flop) count of the call, see DESIGN.md."""

    def __init__(self, name, nlaunch, work=0):
        self.name, self.nlaunch, self.work = name, nlaunch, work

    def __enter__(self):
        if _PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            # The bracketed launch must already be QUEUED when the stream reaches e0, otherwise the pair also
            # times the host (Python + ctypes + tensor-map encode, ~30 us) -- a cost the graph-replayed timed path
            # does not have.  A device-side spin ahead of e0 lets the host run ahead of the GPU.
            if PROFILE_SPIN_CYCLES:
                torch.cuda._sleep(PROFILE_SPIN_CYCLES)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        global LAUNCH_COUNT
        LAUNCH_COUNT += self.nlaunch
        if _PROFILE is not None and exc[0] is None:
            self.e1.record()
            _PROFILE.setdefault(self.name, []).append((self.e0, self.e1, self.work))
        return False


class profile:
    """with ops.profile() as prof: ... ; prof.summary() -> {name: dict(calls, ms, work, per_s)}"""

    def __enter__(self):
        global _PROFILE
        _PROFILE = self.records = {}
        return self

    def __exit__(self, *exc):
        global _PROFILE
        _PROFILE = None
        return False

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, recs in self.records.items():
            ms = sum(e0.elapsed_time(e1) for e0, e1, _ in recs)
            work = sum(w for _, _, w in recs)
            out[name] = dict(calls=len(recs), ms=ms, work=work, per_s=(work / (ms * 1e-3) if ms > 0 else 0.0))
        return out


def _p(t):
    return c_vp(t.data_ptr()) if t is not None else c_vp(0)


def _need_cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("eval_driving_safety_b200 ops need CUDA tensors (no CPU fallback); got %s" % t.device)
        if t.dtype != torch.float32:
            raise RuntimeError("eval_driving_safety_b200 ops are fp32; got %s" % t.dtype)


def cl3(x):
    """NCDHW logical tensor -> channels-last-3d memory (no copy if already)."""
    if x.permute(0, 2, 3, 4, 1).is_contiguous():
        return x
    return x.contiguous(memory_format=CL3)


def cl2(x):
    if x.permute(0, 2, 3, 1).is_contiguous():
        return x
    return x.contiguous(memory_format=CL2)


def empty_cl3(n, c, d, h, w, device):
    # allocated WITH channels-last strides rather than as a permuted view: autograd treats the output of a custom
    # Function that is a view specially (an in-place op on it, e.g. upstream's nn.ReLU(inplace=True) after a swapped
    # norm, is refused)
    return torch.empty((n, c, d, h, w), device=device, dtype=torch.float32, memory_format=CL3)


def empty_cl2(n, c, h, w, device):
    return torch.empty((n, c, h, w), device=device, dtype=torch.float32, memory_format=CL2)


# ---------------------------------------------------------------------------
# (1) cost volume
# ---------------------------------------------------------------------------
class BuildCostVolumeFn(Function):
    """left,right [N,C,H,W], shifts [N,D] -> cost [N,2C,D,H,W] (channels-last-3d
    memory when ``channels_last``; plain NCDHW, the upstream layout, otherwise)."""

    @staticmethod
    def forward(ctx, left, right, shifts, channels_last=True):
        _need_cuda(left, right, shifts)
        lib = _lib.load()
        n, c, h, w = left.shape
        d = shifts.shape[1]
        shifts = shifts.contiguous()
        if channels_last:
            l, r = cl2(left), cl2(right)
            cost = empty_cl3(n, 2 * c, d, h, w, left.device)
        else:
            l, r = left.contiguous(), right.contiguous()
            cost = torch.empty((n, 2 * c, d, h, w), device=left.device, dtype=torch.float32)
        with _op("cost_volume_fwd", 1, 4 * (2 * n * c * h * w + cost.numel())):
            check(lib.b2_cost_volume_fwd(_p(l), _p(r), _p(shifts), _p(cost), n, c, d, h, w,
                                         1 if channels_last else 0, _stream()), "cost_volume_fwd")
        ctx.save_for_backward(shifts)
        ctx.dims = (n, c, d, h, w, channels_last)
        return cost

    @staticmethod
    @once_differentiable
    def backward(ctx, gcost):
        (shifts,) = ctx.saved_tensors
        n, c, d, h, w, channels_last = ctx.dims
        lib = _lib.load()
        if channels_last:
            g = cl3(gcost)
            gl, gr = empty_cl2(n, c, h, w, g.device), empty_cl2(n, c, h, w, g.device)
        else:
            g = gcost.contiguous()
            gl = torch.empty((n, c, h, w), device=g.device, dtype=torch.float32)
            gr = torch.empty_like(gl)
        ws = None
        if channels_last:
            ws = torch.empty(lib.b2_cost_volume_bwd_workspace_bytes(n, c, h, w), device=g.device, dtype=torch.uint8)
        with _op("cost_volume_bwd", 2 if channels_last else 1, 4 * (2 * n * c * h * w + g.numel())):
            check(lib.b2_cost_volume_bwd(_p(g), _p(shifts), _p(gl), _p(gr), n, c, d, h, w,
                                         1 if channels_last else 0, _p(ws), _stream()), "cost_volume_bwd")
        return gl, gr, None, None


def build_cost_volume(left, right, shifts, channels_last=True):
    return BuildCostVolumeFn.apply(left, right, shifts, channels_last)


# ---------------------------------------------------------------------------
# (2) grid_sample lifting
# ---------------------------------------------------------------------------
class GridPlan:
    """CSR 'input cell -> (output voxel, weight)' plan for the deterministic
    backward; built once per sampling grid (b2_grid_plan_*)."""

    def __init__(self, grid, in_spatial, align_corners):
        _need_cuda(grid)
        lib = _lib.load()
        ndim = grid.shape[-1]
        assert ndim in (2, 3) and len(in_spatial) == ndim
        n = grid.shape[0]
        d, h, w = (1,) + tuple(in_spatial) if ndim == 2 else tuple(in_spatial)
        grid = grid.contiguous()
        nvox_per_n = grid[0].numel() // ndim
        ncell = n * d * h * w
        counts = torch.zeros(ncell, device=grid.device, dtype=torch.int32)
        st = _stream()
        check(lib.b2_grid_plan_count(_p(grid), _p(counts), ndim, n, d, h, w, nvox_per_n,
                                     int(align_corners), st), "grid_plan_count")
        row_ptr = torch.zeros(ncell + 1, device=grid.device, dtype=torch.int32)
        row_ptr[1:] = torch.cumsum(counts, 0, dtype=torch.int32)
        nnz = int(row_ptr[-1].item())
        entries = torch.empty((max(nnz, 1), 2), device=grid.device, dtype=torch.int32)
        counts.zero_()
        check(lib.b2_grid_plan_fill(_p(grid), _p(row_ptr), _p(counts), _p(entries), ndim, n, d, h, w,
                                    nvox_per_n, int(align_corners), st), "grid_plan_fill")
        check(lib.b2_grid_plan_sort(_p(row_ptr), _p(entries), ncell, st), "grid_plan_sort")
        self.row_ptr, self.entries, self.ncell, self.nnz = row_ptr, entries, ncell, nnz
        self.ndim, self.in_spatial, self.n = ndim, tuple(in_spatial), n
        self.long_rows = int(nnz > 32 * ncell)       # mean CSR row length decides the backward's mapping


def _gs_fwd(lib, inp, grid, out, out_c, coff, align):
    n, c = inp.shape[:2]
    nvox_per_n = grid[0].numel() // grid.shape[-1]
    # algorithmic bytes: read input once + grid, write the sampled channels
    work = 4 * (inp.numel() + grid.numel() + n * nvox_per_n * c)
    if grid.shape[-1] == 3:
        d, h, w = inp.shape[2:]
        with _op("grid_sample3d_fwd", 1, work):
            check(lib.b2_grid_sample3d_fwd(_p(inp), _p(grid), _p(out), n, c, d, h, w, nvox_per_n, out_c, coff,
                                           int(align), _stream()), "grid_sample3d_fwd")
    else:
        h, w = inp.shape[2:]
        with _op("grid_sample2d_fwd", 1, work):
            check(lib.b2_grid_sample2d_fwd(_p(inp), _p(grid), _p(out), n, c, h, w, nvox_per_n, out_c, coff,
                                           int(align), _stream()), "grid_sample2d_fwd")


class GridSampleFn(Function):
    """F.grid_sample(input, grid, 'bilinear', 'zeros', align_corners) for 4-D/5-D
    inputs; gradient w.r.t. ``input`` only (the grid is calibration, not data)."""

    @staticmethod
    def forward(ctx, inp, grid, align_corners, plan):
        _need_cuda(inp, grid)
        lib = _lib.load()
        nd = grid.shape[-1]
        inp = cl3(inp) if nd == 3 else cl2(inp)
        grid = grid.contiguous()
        n, c = inp.shape[:2]
        if nd == 3:
            out = empty_cl3(n, c, grid.shape[1], grid.shape[2], grid.shape[3], inp.device)
        else:
            out = empty_cl2(n, c, grid.shape[1], grid.shape[2], inp.device)
        _gs_fwd(lib, inp, grid, out, c, 0, align_corners)
        ctx.plan, ctx.grid, ctx.align, ctx.in_shape = plan, grid, align_corners, tuple(inp.shape)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        lib = _lib.load()
        plan = ctx.plan or GridPlan(ctx.grid, ctx.in_shape[2:], ctx.align)
        nd = plan.ndim
        g = cl3(gout) if nd == 3 else cl2(gout)
        c = ctx.in_shape[1]
        gin = empty_cl3(*ctx.in_shape, g.device) if nd == 3 else empty_cl2(*ctx.in_shape, g.device)
        with _op("grid_sample%dd_bwd" % nd, 1, 4 * (g.numel() + gin.numel()) + 8 * plan.nnz + 4 * plan.ncell):
            check(lib.b2_grid_sample_bwd(_p(g), _p(plan.row_ptr), _p(plan.entries), _p(gin), plan.ncell, c, c, 0,
                                         plan.long_rows, _stream()), "grid_sample_bwd")
        return gin, None, None, None


def grid_sample(inp, grid, align_corners=False, plan=None):
    return GridSampleFn.apply(inp, grid, align_corners, plan)


LIFT_FUSED = os.environ.get("B2_LIFT_FUSED", "1") != "0"


class LiftFn(Function):
    """Frustum -> voxel lifting of DSGN: trilinear sample of the PSV feature and
    bilinear sample of the image feature written side by side into ONE
    channels-last voxel tensor [N, C3+C2, Z, Y, X] (no torch.cat round trip)."""

    @staticmethod
    def forward(ctx, psv, img, grid3, plan3, plan2, align_corners):
        _need_cuda(psv, img, grid3)
        lib = _lib.load()
        psv, img, grid3 = cl3(psv), cl2(img), grid3.contiguous()
        n, c3 = psv.shape[:2]
        c2 = img.shape[1]
        z, y, x = grid3.shape[1:4]
        fused = c3 == 64 and c2 == 32 and LIFT_FUSED
        # the 2-channel copy of the grid is only needed by the un-fused 2-D sampler and to build a missing plan
        grid2 = None if (fused and plan2 is not None) else grid3[..., :2].contiguous().view(n, z * y, x, 2)
        out = empty_cl3(n, c3 + c2, z, y, x, psv.device)
        if fused:
            # one launch for both samplings, 8 lanes per voxel
            d, h, w = psv.shape[2:]
            work = 4 * (psv.numel() + img.numel() + grid3.numel() + out.numel())
            with _op("lift_fwd", 1, work):
                check(lib.b2_lift_fwd(_p(psv), _p(img), _p(grid3), _p(out), n, c3, c2, d, h, w, img.shape[2], img.shape[3],
                                      z * y * x, int(align_corners), _stream()), "lift_fwd")
        else:
            _gs_fwd(lib, psv, grid3, out, c3 + c2, 0, align_corners)
            _gs_fwd(lib, img, grid2, out, c3 + c2, c3, align_corners)
        ctx.plans = (plan3 or GridPlan(grid3, psv.shape[2:], align_corners),
                     plan2 or GridPlan(grid2, img.shape[2:], align_corners))
        ctx.shapes = (tuple(psv.shape), tuple(img.shape))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        lib = _lib.load()
        g = cl3(gout)
        (s3, s2), (p3, p2) = ctx.shapes, ctx.plans
        ctot = s3[1] + s2[1]
        g3, g2 = empty_cl3(*s3, g.device), empty_cl2(*s2, g.device)
        nv = g.numel() // ctot
        with _op("grid_sample3d_bwd", 1, 4 * (nv * s3[1] + g3.numel()) + 8 * p3.nnz + 4 * p3.ncell):
            check(lib.b2_grid_sample_bwd(_p(g), _p(p3.row_ptr), _p(p3.entries), _p(g3), p3.ncell, s3[1], ctot, 0,
                                         p3.long_rows, _stream()), "grid_sample_bwd(3d)")
        with _op("grid_sample2d_bwd", 1, 4 * (nv * s2[1] + g2.numel()) + 8 * p2.nnz + 4 * p2.ncell):
            check(lib.b2_grid_sample_bwd(_p(g), _p(p2.row_ptr), _p(p2.entries), _p(g2), p2.ncell, s2[1], ctot, s3[1],
                                         p2.long_rows, _stream()), "grid_sample_bwd(2d)")
        return g3, g2, None, None, None, None


def lift(psv, img, grid3, plan3=None, plan2=None, align_corners=True):
    return LiftFn.apply(psv, img, grid3, plan3, plan2, align_corners)


# ---------------------------------------------------------------------------
# (3) conv3d / deconv3d (+ data gradients) and GroupNorm
# ---------------------------------------------------------------------------
_PACK_CACHE = {}


def _packed(weight, kind):
    """Packed weights wp[27][Cout][Cin] for the gather modes of b2_conv3d.
    kind: conv_fwd, conv_dgrad_s1, conv_dgrad_s2, deconv_fwd, deconv_dgrad."""
    key = (id(weight), kind)
    hit = _PACK_CACHE.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == (weight._version, weight.data_ptr()):
        return hit[2]
    w = weight.detach()
    if kind == "conv_fwd":            # w [Co,Ci,k]: wp[k][co][ci]
        wp = w.permute(2, 3, 4, 0, 1)
    elif kind == "conv_dgrad_s1":     # CONV s1 on gout, wp[k][ci][co] = w[co,ci,flip k]
        wp = w.flip(2, 3, 4).permute(2, 3, 4, 1, 0)
    elif kind == "conv_dgrad_s2":     # DECONV on gout, wp[k][ci][co] = w[co,ci,k]
        wp = w.permute(2, 3, 4, 1, 0)
    elif kind == "deconv_fwd":        # wt [Ci,Co,k]: DECONV, wp[k][co][ci] = wt[ci,co,k]
        wp = w.permute(2, 3, 4, 1, 0)
    elif kind == "deconv_dgrad":      # CONV s2 on gout, wp[k][ci][co] = wt[ci,co,k]
        wp = w.permute(2, 3, 4, 0, 1)
    else:
        raise ValueError(kind)
    wp = wp.reshape(27, wp.shape[3], wp.shape[4]).contiguous()
    if TF32_ROUND_WEIGHTS and wp.is_cuda:
        # tcgen05 kind::tf32 TRUNCATES its operands to 10 mantissa bits; rounding the (frozen) weights to
        # nearest TF32 once at pack time removes the truncation bias of one operand for free
        wp = ((wp.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    _cache_put(key, weight, wp)
    return wp


def _cache_put(key, weight, packed):
    # entries are validated by identity (weakref) + version + storage address, so a freed
    # tensor whose id/address is reused can never produce a stale hit
    if len(_PACK_CACHE) > 1024:
        _PACK_CACHE.clear()
    _PACK_CACHE[key] = (weakref.ref(weight), (weight._version, weight.data_ptr()), packed)


FUSE_GN_STATS = os.environ.get("B2_FUSE_GN_STATS", "1") != "0"
FUSE_GRAD_ADD = os.environ.get("B2_FUSE_GRAD_ADD", "1") != "0"
FUSE_GN_BWD = os.environ.get("B2_FUSE_GN_BWD", "1") != "0"
BSTAT_ON_DECONV = os.environ.get("B2_BSTAT_ON_DECONV", "0") == "1"
_CAPS = {}


class GnLink:
    """Hand-off between a GroupNorm node and the conv that consumes its output.  The norm's forward
    fills in what its backward statistics need (x, stats, activation mode); the conv's DATA-GRADIENT
    launch -- which produces exactly the gradient w.r.t. the norm's output -- adds up the norm's
    backward sums in its epilogue and leaves them here together with the gradient tensor it wrote;
    the norm's backward uses them only if that very tensor, unmodified, arrives.
    The link holds the tensor itself, not just its address: when the norm's output has further consumers the autograd
    engine adds their gradients to the conv's -- in place if it holds the last reference to the buffered tensor, or
    into a new tensor after which the conv's tensor is freed and the caching allocator may hand its address to the
    NEXT sum.  Either way a tensor with the recorded address could arrive that the sums no longer describe (measured:
    2.8 % error in the extractor's input gradient).  While the link holds the conv's tensor neither can happen: the
    engine must add out of place, and the address cannot be reused until ``take``."""
    __slots__ = ("x", "stats", "groups", "mode", "bwd_partial", "gy", "gy_key")

    def __init__(self):
        self.x = self.stats = self.bwd_partial = self.gy = self.gy_key = None
        self.groups = self.mode = 0

    def offer(self, partial, gy):
        self.bwd_partial, self.gy = partial, gy
        self.gy_key = (gy.data_ptr(), gy._version, tuple(gy.shape))

    def take(self, gy):
        """The sums, if ``gy`` is the tensor they were added up from; clears the hand-off either way."""
        ok = self.bwd_partial is not None and self.gy_key == (gy.data_ptr(), gy._version, tuple(gy.shape))
        part = self.bwd_partial if ok else None
        self.bwd_partial = self.gy = self.gy_key = None
        return part


def _conv_caps(n, cin, cout, di, hi, wi, stride, mode):
    """(rows of the statistics table or 0, fused addend available) for the kernel serving this shape."""
    key = (n, cin, cout, di, hi, wi, stride, mode)
    hit = _CAPS.get(key)
    if hit is None:
        rows, aok = ctypes.c_int(0), ctypes.c_int(0)
        check(_lib.load().b2_conv3d_fusion_caps(n, cin, cout, di, hi, wi, stride, mode, ctypes.byref(rows),
                                                ctypes.byref(aok)), "conv3d_fusion_caps")
        hit = _CAPS[key] = (rows.value, bool(aok.value))
    return hit


def _conv_call(x, wp, stride, mode, impl, stats=False, addend=None, bstat=None):
    """``stats``: also return the conv epilogue's GroupNorm partial sums of the output
    ([rows, 2, Cout]) or None when this launch configuration cannot produce them.
    ``addend``: tensor of the output's shape added to the result -- in the epilogue where the kernel
    can (no extra pass), by a separate add otherwise.
    ``bstat``: GnLink of the norm whose output gradient this launch computes; its backward sums are
    added up in the epilogue when possible."""
    lib = _lib.load()
    n, cin, di, hi, wi = x.shape
    cout = wp.shape[1]
    assert wp.shape[2] == cin, (wp.shape, cin)
    if mode == 0:
        do, ho, wo = (di - 1) // stride + 1, (hi - 1) // stride + 1, (wi - 1) // stride + 1
    else:
        do, ho, wo = 2 * di, 2 * hi, 2 * wi
    out = empty_cl3(n, cout, do, ho, wo, x.device)
    # algorithmic flops: 2*Cin*Cout*27 per output voxel for CONV; a transposed conv touches
    # 27/8 taps per output voxel on average (= 2*Cin*Cout*27 per INPUT voxel)
    vox = n * do * ho * wo if mode == 0 else n * di * hi * wi
    rows, aok = _conv_caps(n, cin, cout, di, hi, wi, stride, mode) if impl == 0 else (0, False)
    rows = rows if (stats and FUSE_GN_STATS) else 0
    fused_add = addend is not None and aok and FUSE_GRAD_ADD
    if fused_add:
        addend = cl3(addend)
        assert addend.shape == out.shape, (addend.shape, out.shape)
    part = None
    brows = 0
    # (only on the stride-1 kernel: its 4-plane tiles take ~24 us of MMAs, which hides the epilogue's extra
    # row reads; the transposed kernel's 3-6 us units do not -- measured 0.203 -> 0.344 ms, more than the
    # 0.107 ms statistics pass it would replace.  tools/bench_epilogue.py)
    if (bstat is not None and FUSE_GN_BWD and impl == 0 and not stats and bstat.mode in (0, 2)
            and (mode == 0 and stride == 1 or BSTAT_ON_DECONV)
            and (addend is None or fused_add) and bstat.x is not None and tuple(bstat.x.shape) == tuple(out.shape)):
        brows = _conv_caps(n, cin, cout, di, hi, wi, stride, mode)[0]
    with _op("conv3d_tcgen05" if impl == 0 else "conv3d_simt", 1, 2 * cin * cout * 27 * vox):
        if brows > 0:
            bpart = torch.empty((brows, 2, cout), device=x.device, dtype=torch.float32)
            coef = ctypes.c_void_p(bstat.stats.data_ptr() + 4 * 2 * bstat.groups) if bstat.mode == 2 else None
            check(lib.b2_conv3d_fused(_p(x), _p(wp), _p(out), _p(addend) if fused_add else None, _p(bpart),
                                      3 if bstat.mode == 2 else 2, _p(bstat.x), coef, n, cin, cout,
                                      di, hi, wi, stride, mode, _stream()),
                  "conv3d_fused(mode=%d,stride=%d,bstat)" % (mode, stride))
            bstat.offer(bpart, out)
        elif rows > 0 or fused_add:
            part = torch.empty((rows, 2, cout), device=x.device, dtype=torch.float32) if rows > 0 else None
            check(lib.b2_conv3d_fused(_p(x), _p(wp), _p(out), _p(addend) if fused_add else None, _p(part),
                                      1 if rows > 0 else 0, None, None, n, cin, cout,
                                      di, hi, wi, stride, mode, _stream()),
                  "conv3d_fused(mode=%d,stride=%d)" % (mode, stride))
        else:
            check(lib.b2_conv3d(_p(x), _p(wp), _p(out), n, cin, cout, di, hi, wi, stride, mode, impl, _stream()),
                  "conv3d(mode=%d,stride=%d,impl=%d)" % (mode, stride, impl))
    if addend is not None and not fused_add:
        out = out + addend
    return (out, part) if stats else out


class Conv3dFn(Function):
    """3x3x3, padding 1, bias-free Conv3d (transposed=False, stride 1|2) or
    ConvTranspose3d(stride 2, output_padding 1) (transposed=True); forward and
    data gradient.  Weights are frozen in an attack: no weight gradient."""

    @staticmethod
    def forward(ctx, x, weight, stride, transposed, impl, stats=False, fork=False, gn_link=None):
        _need_cuda(x, weight)
        ctx.gn_link = gn_link
        if weight.requires_grad:
            raise RuntimeError("attack path: weights are frozen; call requires_grad_(False) on the model "
                               "(the reference wastes its wgrad, attack/DSGN/pgd_attack.py:333)")
        if tuple(weight.shape[2:]) != (3, 3, 3):
            raise RuntimeError("only 3x3x3 kernels are supported")
        x = cl3(x)
        if stride == 2 and any(s % 2 for s in x.shape[2:]) and not transposed:
            raise RuntimeError("stride-2 conv needs even spatial dims, got %s" % (tuple(x.shape[2:]),))
        impl = CONV_IMPL if impl is None else impl
        if transposed:
            assert stride == 2
            out = _conv_call(x, _packed(weight, "deconv_fwd"), 2, 1, impl, stats)
        else:
            out = _conv_call(x, _packed(weight, "conv_fwd"), stride, 0, impl, stats)
        ctx.weight, ctx.cfg = weight, (stride, transposed, impl)
        ctx.layout = (bool(stats), bool(fork))
        ctx.set_materialize_grads(False)              # an unused output's gradient stays None (no zero volumes)
        res = []
        if stats:
            out, part = out
            if part is None:
                part = out.new_empty(0)               # "no statistics": autograd outputs must be tensors
            ctx.mark_non_differentiable(part)
            res.append(part)
        if fork:
            # second handle on the input for its OTHER consumers: their gradient comes back to this
            # node, and the data-gradient kernel adds it in its epilogue instead of autograd's add pass
            res.append(x.view_as(x))
        return (out, *res) if res else out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout, *extra):
        stride, transposed, impl = ctx.cfg
        stats, fork = ctx.layout
        g_other = extra[int(stats)] if fork else None          # gradient of the forked input handle
        if gout is None:                                       # only the forked handle was used downstream
            return g_other, None, None, None, None, None, None, None
        g = cl3(gout)
        link = ctx.gn_link
        if transposed:
            gin = _conv_call(g, _packed(ctx.weight, "deconv_dgrad"), 2, 0, impl, addend=g_other, bstat=link)
        elif stride == 1:
            gin = _conv_call(g, _packed(ctx.weight, "conv_dgrad_s1"), 1, 0, impl, addend=g_other, bstat=link)
        else:
            gin = _conv_call(g, _packed(ctx.weight, "conv_dgrad_s2"), 2, 1, impl, addend=g_other, bstat=link)
        return gin, None, None, None, None, None, None, None


def conv3d(x, weight, stride=1, transposed=False, impl=None):
    return Conv3dFn.apply(x, weight, stride, transposed, impl, False, False, getattr(x, "_b2_gn", None))


def conv3d_with_stats(x, weight, stride=1, transposed=False, impl=None):
    """conv3d whose epilogue also adds up the GroupNorm statistics of its output.  Returns
    (y, partial): ``partial`` [rows, 2, Cout] goes to ``groupnorm_act(..., partial=)``; it is None
    when the kernel that serves this shape has no statistics epilogue (GroupNorm then runs its own
    statistics pass)."""
    y, part = Conv3dFn.apply(x, weight, stride, transposed, impl, True, False, getattr(x, "_b2_gn", None))
    return y, (part if part.numel() else None)


def conv3d_fork(x, weight, stride=1, transposed=False, impl=None):
    """For an input with SEVERAL consumers.  Returns (y, partial, x2): y, partial as
    ``conv3d_with_stats``; x2 is the same data as x and must be what every other consumer of x reads.
    In the backward the gradient those consumers send to x2 arrives at this node and is added by the
    data-gradient kernel's epilogue (out = dgrad + other) -- autograd's separate accumulation pass
    (two reads and a write of the full volume) disappears."""
    y, part, x2 = Conv3dFn.apply(x, weight, stride, transposed, impl, True, True, getattr(x, "_b2_gn", None))
    return y, (part if part.numel() else None), x2


# ---------------------------------------------------------------------------
# 2-D convolutions (feature extractor, BEV head): tcgen05 with in-kernel 3xTF32
# ---------------------------------------------------------------------------
# True: error-compensated 3xTF32 (fp32-class accuracy -- the reference computes in fp32); False: plain TF32
CONV2D_SPLIT = int(os.environ.get("B2_CONV2D_SPLIT", "1"))
# Operand split of the DATA-GRADIENT launches.  Default 0 = plain TF32: the backward pass is LINEAR in the gradient it
# propagates (the linearisation point -- activations, GroupNorm statistics, ReLU masks -- comes from the forward, which
# keeps the split), so a 2^-11 operand error perturbs a gradient entry by ~1e-3 relative per layer and cannot flip its
# sign unless the entry is already ~0; the 3-D convs work the same way in both directions.  Measured on the timed path
# (tests/test_gpu_fullsize.py, 10 iterations x 2 pairs): sign agreement mean 99.944 % / min 99.901 %, identical pixels
# mean 98.338 % / min 98.167 % -- the same to four digits as with the split in both directions (98.337 / 98.182), at
# 43.1 instead of 40.2 pair-iterations/s.  B2_CONV2D_SPLIT_BWD=1 splits the data gradients too.
CONV2D_SPLIT_BWD = int(os.environ.get("B2_CONV2D_SPLIT_BWD", "0"))


def set_conv2d_split(flag):
    """False / 0: plain TF32; True / 1: 3xTF32; 2: 3xTF32 with the activation tile rewritten in place as its
    truncated value (verification of the tensor core's truncation, same results)."""
    global CONV2D_SPLIT
    CONV2D_SPLIT = int(flag)


def tf32_split(t):
    """t = hi + lo exactly; hi has its low 13 mantissa bits cleared (what kind::tf32 reads), lo = t - hi."""
    hi = (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    return hi, t - hi


def _packed2d(weight, kind, split):
    """Packed weights [S][k*k][Nout][K] for b2_conv2d (S = 2: w_hi, w_lo when ``split``).
    kind: fwd (Nout=Co, K=Ci), dgrad_s1 (CONV on gout, flipped taps, Nout=Ci, K=Co), dgrad_s2 (DECONV on gout)."""
    key = (id(weight), kind, bool(split))
    hit = _PACK_CACHE.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == (weight._version, weight.data_ptr()):
        return hit[2]
    w = weight.detach()
    if kind == "fwd":
        wp = w.permute(2, 3, 0, 1)
    elif kind == "dgrad_s1":
        wp = w.flip(2, 3).permute(2, 3, 1, 0)
    elif kind == "dgrad_s2":
        wp = w.permute(2, 3, 1, 0)
    else:
        raise ValueError(kind)
    wp = wp.reshape(wp.shape[0] * wp.shape[1], wp.shape[2], wp.shape[3]).contiguous()
    if split:
        hi, lo = tf32_split(wp)
        wp = torch.stack([hi, lo]).contiguous()
    _cache_put(key, weight, wp)
    return wp


def _plain_weight(weight):
    """Contiguous [Co,Ci,kh,kw] copy of a (possibly channels-last) frozen weight, cached."""
    if weight.is_contiguous():
        return weight.detach()
    key = (id(weight), "plain")
    hit = _PACK_CACHE.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == (weight._version, weight.data_ptr()):
        return hit[2]
    w = weight.detach().contiguous()
    _cache_put(key, weight, w)
    return w


_CAPS2D = {}


def _conv2d_stat_rows(n, cin, cout, hi, wi, ks, stride, dil, mode, split):
    """Rows per sample of the statistics table the conv2d kernel serving this shape writes (0: none)."""
    key = (n, cin, cout, hi, wi, ks, stride, dil, mode, int(split))
    hit = _CAPS2D.get(key)
    if hit is None:
        rows = ctypes.c_int(0)
        check(_lib.load().b2_conv2d_stat_rows(n, cin, cout, hi, wi, ks, stride, dil, mode, int(split), ctypes.byref(rows)),
              "conv2d_stat_rows")
        hit = _CAPS2D[key] = rows.value
    return hit


def _conv2d_call(x, wp, bias, addend, n, cin, cout, hi, wi, ks, stride, dil, mode, split, stats=False, bstat=None):
    """``stats``: also return the GroupNorm partial sums of the output ([n, rows, 2, cout]) added up by the epilogue,
    or None when the kernel serving this shape has none.  ``bstat``: GnLink of the norm whose output gradient this
    launch computes; its backward sums are added up in the epilogue when possible (as ``_conv_call``)."""
    lib = _lib.load()
    if mode == 0:
        ho, wo = (hi - 1) // stride + 1, (wi - 1) // stride + 1
    else:
        ho, wo = 2 * hi, 2 * wi
    out = empty_cl2(n, cout, ho, wo, x.device)
    if addend is not None:
        addend = cl2(addend)
        assert addend.shape == out.shape, (addend.shape, out.shape)
    pix = n * ho * wo if mode == 0 else n * hi * wi
    what = "conv2d(k=%d,stride=%d,dil=%d,mode=%d,%d->%d)" % (ks, stride, dil, mode, cin, cout)
    part, stat_mode, gn_x, gn_coef, coef_stride = None, 0, None, None, 0
    if bstat is not None and not (FUSE_GN_BWD and bstat.mode in (0, 2) and bstat.x is not None
                                  and tuple(bstat.x.shape) == tuple(out.shape)):
        bstat = None
    if (stats and FUSE_GN_STATS) or bstat is not None:
        rows = _conv2d_stat_rows(n, cin, cout, hi, wi, ks, stride, dil, mode, split)
        if rows:
            part = torch.empty((n, rows, 2, cout), device=x.device, dtype=torch.float32)
            if bstat is not None:
                stat_mode = 3 if bstat.mode == 2 else 2
                gn_x = bstat.x
                coef_stride = bstat.stats.shape[1]
                gn_coef = bstat.stats[:, 2 * bstat.groups:]           # scale[C], shift[C] of sample 0
            else:
                stat_mode = 1
    with _op("conv2d_tcgen05", 1, 2 * cin * cout * ks * ks * pix):
        if stat_mode:
            check(lib.b2_conv2d_fused(_p(x), _p(wp), _p(bias), _p(addend), _p(out), n, cin, cout, hi, wi, ks, stride,
                                      dil, mode, int(split), stat_mode, _p(part), _p(gn_x), _p(gn_coef), coef_stride,
                                      _stream()), what)
        else:
            check(lib.b2_conv2d(_p(x), _p(wp), _p(bias), _p(addend), _p(out), n, cin, cout, hi, wi, ks, stride, dil,
                                mode, int(split), _stream()), what)
    if bstat is not None and stat_mode >= 2:
        bstat.offer(part, out)
    if stats:
        return out, (part if stat_mode == 1 else None)
    return out


class Conv2dFn(Function):
    """nn.Conv2d(k 1|3, stride 1|2, padding = dilation*(k//2), dilation 1|2, groups 1) forward and data
    gradient on the sm_100a kernels; channels-last maps.  The 3-channel first layer (k3, s2) takes the exact
    fp32 SIMT pair and reads / writes the NCHW image tensors of the attack directly.
    ``stats``: also return the epilogue's GroupNorm partial sums of the output (empty tensor: none available).
    ``fork``: also return a second handle on x for its other consumers (see ``conv3d_fork``).
    ``gn_link``: GnLink of the norm that produced x (its backward sums ride this conv's data-gradient epilogue)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, dilation, fork, stats=False, gn_link=None):
        _need_cuda(x, weight, bias)
        if weight.requires_grad or (bias is not None and bias.requires_grad):
            raise RuntimeError("attack path: weights are frozen; call requires_grad_(False) on the model")
        lib = _lib.load()
        co, ci, kh, kw = weight.shape
        n, _, hi, wi = x.shape
        if kh != kw or kh not in (1, 3) or stride not in (1, 2) or dilation not in (1, 2) or x.shape[1] != ci:
            raise RuntimeError("conv2d: unsupported configuration k=%dx%d stride=%d dilation=%d" % (kh, kw, stride, dilation))
        first = ci == 3
        split = CONV2D_SPLIT
        part = None
        if first:
            if not (kh == 3 and stride == 2 and dilation == 1 and bias is None):
                raise RuntimeError("conv2d: the 3-channel layer must be k3/s2/p1/bias-free")
            x = x.contiguous()
            ho, wo = (hi - 1) // 2 + 1, (wi - 1) // 2 + 1
            out = empty_cl2(n, co, ho, wo, x.device)
            w = _plain_weight(weight)
            with _op("conv2d_first_fwd", 1, 4 * (x.numel() + out.numel())):
                check(lib.b2_conv2d_first_fwd(_p(x), _p(w), _p(out), n, co, hi, wi, _stream()), "conv2d_first_fwd")
        else:
            x = cl2(x)
            if stride == 2 and (hi % 2 or wi % 2):
                raise RuntimeError("conv2d: stride 2 needs even spatial dims, got %dx%d" % (hi, wi))
            b = bias.detach() if bias is not None else None
            out = _conv2d_call(x, _packed2d(weight, "fwd", split), b, None, n, ci, co, hi, wi, kh, stride, dilation, 0,
                               split, stats=stats)
            if stats:
                out, part = out
        ctx.weight = weight
        ctx.gn_link = gn_link
        ctx.cfg = (n, ci, co, hi, wi, kh, stride, dilation, first, (split if CONV2D_SPLIT_BWD else 0),
                   bool(fork), bool(stats))
        ctx.set_materialize_grads(False)
        res = []
        if stats:
            if part is None:
                part = out.new_empty(0)                   # "no statistics": autograd outputs must be tensors
            ctx.mark_non_differentiable(part)
            res.append(part)
        if fork:
            res.append(x.view_as(x))
        return (out, *res) if res else out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout, *extra):
        n, ci, co, hi, wi, ks, stride, dil, first, split, fork, stats = ctx.cfg
        g_other = extra[int(stats)] if fork else None
        if gout is None:
            return g_other, None, None, None, None, None, None, None
        lib = _lib.load()
        g = cl2(gout)
        if first:
            gin = torch.empty((n, 3, hi, wi), device=g.device, dtype=torch.float32)
            w = _plain_weight(ctx.weight)
            with _op("conv2d_first_dgrad", 1, 4 * (g.numel() + gin.numel())):
                check(lib.b2_conv2d_first_dgrad(_p(g), _p(w), _p(gin), n, co, hi, wi, _stream()), "conv2d_first_dgrad")
            if g_other is not None:
                gin = gin + g_other
            return gin, None, None, None, None, None, None, None
        ho, wo = g.shape[2:]
        link = ctx.gn_link
        if stride == 1:
            gin = _conv2d_call(g, _packed2d(ctx.weight, "dgrad_s1", split), None, g_other, n, co, ci, ho, wo, ks, 1, dil,
                               0, split, bstat=link)
        else:
            gin = _conv2d_call(g, _packed2d(ctx.weight, "dgrad_s2", split), None, g_other, n, co, ci, ho, wo, ks, 2, 1,
                               1, split, bstat=link)
        return gin, None, None, None, None, None, None, None


def _padded_out_channels(weight, bias):
    """Weight / bias zero-padded along the OUTPUT channels to a multiple of 32 (the tensor-core N tile and the data
    gradient's K chunk), cached per frozen weight: e.g. the 4- and 28-channel detection heads."""
    key = (id(weight), "padout")
    hit = _PACK_CACHE.get(key)
    bver = None if bias is None else (bias._version, bias.data_ptr())
    if hit is not None and hit[0]() is weight and hit[1] == (weight._version, weight.data_ptr()) and hit[2][2] == bver:
        return hit[2][0], hit[2][1]
    co = weight.shape[0]
    pad = (-co) % 32
    w = torch.cat([weight.detach(), weight.new_zeros((pad,) + tuple(weight.shape[1:]))], 0).contiguous()
    b = None if bias is None else torch.cat([bias.detach(), bias.new_zeros(pad)], 0).contiguous()
    _cache_put(key, weight, (w, b, bver))
    return w, b


def conv2d(x, weight, bias=None, stride=1, dilation=1):
    co = weight.shape[0]
    if weight.shape[1] != 3 and co % 32:
        if weight.requires_grad:
            raise RuntimeError("attack path: weights are frozen; call requires_grad_(False) on the model")
        w, b = _padded_out_channels(weight, bias)
        return Conv2dFn.apply(x, w, b, stride, dilation, False, False, getattr(x, "_b2_gn", None))[:, :co]
    return Conv2dFn.apply(x, weight, bias, stride, dilation, False, False, getattr(x, "_b2_gn", None))


def conv2d_fork(x, weight, bias=None, stride=1, dilation=1):
    """(y, x2): x2 is x for its OTHER consumers; their gradient is added inside this conv's data-gradient
    kernel (no separate accumulation pass), exactly as ``conv3d_fork``."""
    return Conv2dFn.apply(x, weight, bias, stride, dilation, True, False, getattr(x, "_b2_gn", None))


def conv2d_with_stats(x, weight, bias=None, stride=1, dilation=1, fork=False):
    """conv2d whose epilogue also adds up the GroupNorm statistics of its output: (y, partial[, x2]); ``partial``
    [N, rows, 2, Cout] goes to ``groupnorm_act(..., partial=)`` and is None when the kernel serving this shape has
    no statistics epilogue (the 3-channel first layer, odd widths)."""
    if weight.shape[1] != 3 and weight.shape[0] % 32:
        if fork:
            raise RuntimeError("conv2d_with_stats: a forked input needs an output width that is a multiple of 32 (got %d)"
                               % weight.shape[0])
        return conv2d(x, weight, bias, stride, dilation), None
    out = Conv2dFn.apply(x, weight, bias, stride, dilation, bool(fork), True, getattr(x, "_b2_gn", None))
    y, part = out[0], out[1]
    part = part if part.numel() else None
    return (y, part, out[2]) if fork else (y, part)


class Conv3dC1Fn(Function):
    """Conv3d(Cin -> 1, k3, p1, bias-free): bandwidth-bound SIMT head."""

    @staticmethod
    def forward(ctx, x, weight):
        _need_cuda(x, weight)
        lib = _lib.load()
        x = cl3(x)
        n, cin, d, h, w = x.shape
        key = (id(weight), "c1")
        hit = _PACK_CACHE.get(key)
        if hit is not None and hit[0]() is weight and hit[1] == (weight._version, weight.data_ptr()):
            w1 = hit[2]
        else:
            w1 = weight.detach()[0].permute(1, 2, 3, 0).reshape(27, cin).contiguous()
            _cache_put(key, weight, w1)
        out = torch.empty((n, 1, d, h, w), device=x.device, dtype=torch.float32)
        ws = torch.empty(lib.b2_conv3d_c1_workspace_bytes(n, d, h, w), device=x.device, dtype=torch.uint8)
        with _op("conv3d_c1_fwd", 2, 4 * (x.numel() + out.numel())):
            check(lib.b2_conv3d_c1_fwd(_p(x), _p(w1), _p(out), n, cin, d, h, w, _p(ws), _stream()), "conv3d_c1_fwd")
        ctx.w1, ctx.dims = w1, (n, cin, d, h, w)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        lib = _lib.load()
        n, cin, d, h, w = ctx.dims
        g = gout.contiguous()
        gin = empty_cl3(n, cin, d, h, w, g.device)
        with _op("conv3d_c1_dgrad", 1, 4 * (g.numel() + gin.numel())):
            check(lib.b2_conv3d_c1_dgrad(_p(g), _p(ctx.w1), _p(gin), n, cin, d, h, w, _stream()), "conv3d_c1_dgrad")
        return gin, None


def conv3d_c1(x, weight):
    return Conv3dC1Fn.apply(x, weight)


_WS = {}


def _workspace(nbytes, device):
    key = (device, torch.cuda.current_stream().cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), device=device, dtype=torch.uint8)
        _WS[key] = ws
    return ws


def _cl(x):
    return cl3(x) if x.dim() == 5 else cl2(x)


def _empty_cl_like(x):
    return empty_cl3(*x.shape, x.device) if x.dim() == 5 else empty_cl2(*x.shape, x.device)


class GroupNormActFn(Function):
    """y = act(GroupNorm(x) (+ res)) on channels-last 3-D volumes or 2-D maps; data gradient only."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, groups, eps, relu, partial=None, link=None):
        _need_cuda(x, res, gamma, beta)
        lib = _lib.load()
        x = _cl(x)
        res = _cl(res) if res is not None else None
        n, c = x.shape[:2]
        s = x[0, 0].numel()
        y = _empty_cl_like(x)
        stats = torch.empty((n, 2 * groups + 2 * c), device=x.device, dtype=torch.float32)
        ws = _workspace(lib.b2_groupnorm_workspace_bytes(n, c), x.device)
        gamma, beta = gamma.detach().contiguous(), beta.detach().contiguous()
        if partial is not None:
            # [rows, 2, C] (single sample) or [N, rows, 2, C]
            assert (partial.dim() == 3 and n == 1) or (partial.dim() == 4 and partial.shape[0] == n), (partial.shape, n)
            assert tuple(partial.shape[-2:]) == (2, c) and partial.is_contiguous(), (partial.shape, c)
            with _op("groupnorm_fwd", 2, 4 * x.numel() * (2 + (res is not None))):
                check(lib.b2_groupnorm_fwd_ext(_p(x), _p(res), _p(gamma), _p(beta), _p(y), _p(stats), n, c, s, groups,
                                               float(eps), int(relu), _p(partial), partial.shape[-3], _p(ws),
                                               _stream()), "groupnorm_fwd_ext")
        else:
            with _op("groupnorm_fwd", 3, 4 * x.numel() * (3 + (res is not None))):
                check(lib.b2_groupnorm_fwd(_p(x), _p(res), _p(gamma), _p(beta), _p(y), _p(stats), n, c, s, groups,
                                           float(eps), int(relu), _p(ws), _stream()), "groupnorm_fwd")
        # ReLU without residual: the backward recomputes the mask from x (relu mode 2), so y is not
        # kept alive by this node (and is never re-read)
        keep_y = relu and res is not None
        ctx.save_for_backward(x, y if keep_y else None, gamma, stats)
        ctx.cfg = (groups, relu, res is not None)
        ctx.link = link
        if link is not None:
            link.x, link.stats, link.groups = x, stats, groups
            link.mode = 0 if not relu else (1 if res is not None else 2)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        lib = _lib.load()
        x, y, gamma, stats = ctx.saved_tensors
        groups, relu, has_res = ctx.cfg
        gy = _cl(gy)
        n, c = x.shape[:2]
        s = x[0, 0].numel()
        gx = _empty_cl_like(x)
        gres = _empty_cl_like(x) if (has_res and relu) else None
        ws = _workspace(lib.b2_groupnorm_workspace_bytes(n, c), x.device)
        mode = 0 if not relu else (1 if has_res else 2)
        ext = None
        link = ctx.link
        if link is not None:
            # the conv that consumed y added up (sum gz*x, sum gz) while writing gy: valid only if that
            # very tensor arrives here untouched (no other consumer's gradient was accumulated into it)
            ext = link.take(gy)
            if mode == 1:
                ext = None
        if ext is not None:
            with _op("groupnorm_bwd", 2, 4 * x.numel() * (3 + (gres is not None))):
                check(lib.b2_groupnorm_bwd_ext(_p(gy), _p(x), _p(y), _p(gamma), _p(stats), _p(gx), _p(gres), n, c, s,
                                               groups, mode, _p(ext), ext.shape[-3], _p(ws), _stream()),
                      "groupnorm_bwd_ext")
        else:
            with _op("groupnorm_bwd", 3, 4 * x.numel() * (5 + 2 * int(mode == 1) + (gres is not None))):
                check(lib.b2_groupnorm_bwd(_p(gy), _p(x), _p(y), _p(gamma), _p(stats), _p(gx), _p(gres), n, c, s,
                                           groups, mode, _p(ws), _stream()), "groupnorm_bwd")
        if has_res and not relu:
            gres = gy
        return gx, gres, None, None, None, None, None, None, None


def groupnorm_act(x, gamma, beta, groups, eps=1e-5, relu=False, res=None, partial=None):
    """``partial``: per-channel partial sums of x from the producing conv's epilogue
    (``conv3d_with_stats``); the statistics pass over x is skipped.
    A single-sample 3-D output carries a ``GnLink`` so that a conv consuming it can add up this
    norm's backward sums while it writes the gradient (see ``GnLink``)."""
    link = GnLink() if (FUSE_GN_BWD and ((x.dim() == 5 and x.shape[0] == 1) or x.dim() == 4)) else None
    y = GroupNormActFn.apply(x, res, gamma, beta, groups, eps, relu, partial, link)
    if link is not None:
        y._b2_gn = link
    return y


# ---------------------------------------------------------------------------
# concat / split glue of the extractor (one launch per direction)
# ---------------------------------------------------------------------------
def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def _int_array(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def _channel_concat(pieces, widths, n, h, w, device):
    """[N, sum(widths), H, W] channels-last from channels-last pieces (None = zeros), one launch."""
    lib = _lib.load()
    out = empty_cl2(n, sum(widths), h, w, device)
    with _op("channel_concat", 1, 8 * out.numel()):
        check(lib.b2_channel_concat(_ptr_array(pieces), _int_array(widths), len(widths), _p(out), n * h * w, _stream()),
              "channel_concat")
    return out


def _channel_split(wide, widths, want):
    """Contiguous channels-last pieces of ``wide`` (only those with want[k]), one launch."""
    lib = _lib.load()
    n, _, h, w = wide.shape
    outs = [empty_cl2(n, wk, h, w, wide.device) if wk and k < len(want) and want[k] else None for k, wk in enumerate(widths)]
    with _op("channel_split", 1, 8 * wide.numel()):
        check(lib.b2_channel_split(_p(wide), _ptr_array(outs), _int_array(widths), len(widths), n * h * w, _stream()),
              "channel_split")
    return outs


class CatChannelsFn(Function):
    """torch.cat(xs, 1) of channels-last maps; the backward hands every input its contiguous gradient slice from ONE
    launch (stock autograd: strided views of the wide gradient that each consumer copies)."""

    @staticmethod
    def forward(ctx, *xs):
        _need_cuda(*xs)
        xs = [cl2(x) for x in xs]
        n, _, h, w = xs[0].shape
        widths = [x.shape[1] for x in xs]
        if any(x.shape[0] != n or tuple(x.shape[2:]) != (h, w) for x in xs) or len(xs) > 8 or any(c % 4 for c in widths):
            raise RuntimeError("cat_channels: up to 8 maps of one size with channel counts that are multiples of 4")
        ctx.widths = widths
        return _channel_concat(xs, widths, n, h, w, xs[0].device)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return tuple(_channel_split(cl2(g), ctx.widths, ctx.needs_input_grad))


def cat_channels(xs):
    return CatChannelsFn.apply(*xs)


class SplitChannelsFn(Function):
    """y -> views y[:, 0:w0], y[:, w0:w0+w1], ... (trailing channels beyond sum(widths) are padding); the backward
    assembles the gradient of y from whichever pieces received one in ONE launch (zeros elsewhere)."""

    @staticmethod
    def forward(ctx, y, *widths):
        _need_cuda(y)
        y = cl2(y)
        if any(wk % 4 for wk in widths) or y.shape[1] % 4 or sum(widths) > y.shape[1] or len(widths) > 7:
            raise RuntimeError("split_channels: up to 7 widths, multiples of 4, within the %d channels" % y.shape[1])
        ctx.widths, ctx.shape = list(widths), tuple(y.shape)
        ctx.set_materialize_grads(False)
        outs, o = [], 0
        for wk in widths:
            outs.append(y[:, o:o + wk])
            o += wk
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *gs):
        n, c, h, w = ctx.shape
        widths = list(ctx.widths)
        pieces = [None if g is None else cl2(g) for g in gs]
        if sum(widths) < c:
            widths.append(c - sum(widths)); pieces.append(None)
        dev = next(g.device for g in gs if g is not None)
        return (_channel_concat(pieces, widths, n, h, w, dev),) + (None,) * len(ctx.widths)


def split_channels(y, widths):
    return SplitChannelsFn.apply(y, *widths)


class PrefixForkFn(Function):
    """x -> (x, x[:n]) for a tensor consumed whole AND through its first n samples; the backward merges the two
    gradients in ONE launch (stock autograd: zero-filled full-size buffer + copy + add)."""

    @staticmethod
    def forward(ctx, x, n):
        _need_cuda(x)
        x = cl2(x)
        ctx.n, ctx.shape = int(n), tuple(x.shape)
        ctx.set_materialize_grads(False)
        return x.view_as(x), x[:n]

    @staticmethod
    @once_differentiable
    def backward(ctx, g_full, g_pre):
        if g_pre is None:
            return g_full, None
        lib = _lib.load()
        g_pre = cl2(g_pre)
        if g_full is None:                                        # only the prefix view was used downstream
            out = torch.zeros(ctx.shape, device=g_pre.device, dtype=g_pre.dtype).contiguous(memory_format=CL2)
            out[:ctx.n] = g_pre
            return out, None
        g_full = cl2(g_full)
        out = empty_cl2(*g_full.shape, g_full.device)
        with _op("add_prefix", 1, 4 * (2 * g_full.numel() + g_pre.numel())):
            check(lib.b2_add_prefix(_p(g_full), _p(g_pre), _p(out), g_full.numel(), g_pre.numel(), _stream()), "add_prefix")
        return out, None


def prefix_fork(x, n):
    return PrefixForkFn.apply(x, n)


class SplitBatchFn(Function):
    """x -> (x[:n], x[n:]); backward = one concatenation of the two gradients."""

    @staticmethod
    def forward(ctx, x, n):
        ctx.n, ctx.shape = int(n), tuple(x.shape)
        ctx.set_materialize_grads(False)
        return x[:n], x[n:]

    @staticmethod
    @once_differentiable
    def backward(ctx, ga, gb):
        n, shape = ctx.n, ctx.shape
        ref = ga if ga is not None else gb
        if ga is None:
            ga = ref.new_zeros((n,) + shape[1:])
        if gb is None:
            gb = ref.new_zeros((shape[0] - n,) + shape[1:])
        fmt = CL2 if len(shape) == 4 else torch.contiguous_format
        return torch.cat([ga.contiguous(memory_format=fmt), gb.contiguous(memory_format=fmt)], 0), None


def split_batch(x, n):
    return SplitBatchFn.apply(x, n)


class BevPoolFn(Function):
    """[N,C,Z,Y,X] volume -> [N, C*(Y/p), Z, X] BEV map (channel = c*(Y/p)+yy), both channels-last:
    F.avg_pool3d(v, (1,p,1)).permute(0,1,3,2,4).reshape(n, c*yy, z, x) in one pass."""

    @staticmethod
    def forward(ctx, v, p):
        _need_cuda(v)
        lib = _lib.load()
        v = cl3(v)
        n, c, z, y, x = v.shape
        bev = empty_cl2(n, c * (y // p), z, x, v.device)
        with _op("bev_pool_fwd", 1, 4 * (v.numel() + bev.numel())):
            check(lib.b2_bev_pool_fwd(_p(v), _p(bev), n, c, z, y, x, p, _stream()), "bev_pool_fwd")
        ctx.cfg = (n, c, z, y, x, p)
        return bev

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        n, c, z, y, x, p = ctx.cfg
        g = cl2(g)
        gv = empty_cl3(n, c, z, y, x, g.device)
        with _op("bev_pool_bwd", 1, 4 * (g.numel() + gv.numel())):
            check(lib.b2_bev_pool_bwd(_p(g), _p(gv), n, c, z, y, x, p, _stream()), "bev_pool_bwd")
        return gv, None


def bev_pool(v, p):
    return BevPoolFn.apply(v, p)


class DepthHeadFn(Function):
    """cost1 [N,1,D,Hc,Wc] -> depth [N,H,W]: fused trilinear upsample + softmax + expectation."""

    @staticmethod
    def forward(ctx, cost1, size, z0, dz):
        _need_cuda(cost1)
        lib = _lib.load()
        cost1 = cost1.contiguous()
        n, _, d, hc, wc = cost1.shape
        j, h, w = size
        depth = torch.empty((n, h, w), device=cost1.device, dtype=torch.float32)
        # softmax (max, sum) per pixel, kept for the backward (which then skips its own softmax pass)
        sm = torch.empty((n, h, w, 2), device=cost1.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        with _op("depth_head_fwd", 1, 4 * (cost1.numel() + depth.numel())):
            check(lib.b2_depth_head_fwd(_p(cost1), _p(depth), _p(sm), n, d, hc, wc, h, w, j, float(z0), float(dz),
                                        _stream()), "depth_head_fwd")
        ctx.save_for_backward(cost1, sm, depth if sm is not None else None)
        ctx.cfg = (n, d, hc, wc, h, w, j, float(z0), float(dz))
        return depth

    @staticmethod
    @once_differentiable
    def backward(ctx, gdepth):
        lib = _lib.load()
        cost1, sm, depth = ctx.saved_tensors
        n, d, hc, wc, h, w, j, z0, dz = ctx.cfg
        g = gdepth.contiguous()
        gcost = torch.empty_like(cost1)
        ws = torch.empty(lib.b2_depth_head_workspace_bytes(n, d, h, w), device=g.device, dtype=torch.uint8)
        with _op("depth_head_bwd", 2, 4 * (cost1.numel() * 2 + g.numel()) + 2 * ws.numel()):
            check(lib.b2_depth_head_bwd(_p(cost1), _p(g), _p(gcost), _p(sm), _p(depth), n, d, hc, wc, h, w, j, z0, dz,
                                        _p(ws), _stream()), "depth_head_bwd")
        return gcost, None, None, None


def depth_head(cost1, size, z0, dz):
    """size = (maxdisp, H, W); plane depths z_j = z0 + (j + 0.5) * dz."""
    return DepthHeadFn.apply(cost1, tuple(size), z0, dz)


# ---------------------------------------------------------------------------
# RoIAlign (Stereo R-CNN, config 5)
# ---------------------------------------------------------------------------
def _roi_gout(gout, c):
    """(gradient of the pooled features as the backward kernel reads it, its layout code): [R,P,P,C] memory (one copy
    unless the producer already wrote channels-last) for the FPN widths, the upstream [R,C,P,P] otherwise."""
    if c % 32 == 0 and c <= 256:
        return gout.contiguous(memory_format=CL2), 1
    return gout.contiguous(), 0


class RoIAlignFn(Function):
    @staticmethod
    def forward(ctx, feat, rois, pooled, scale):
        _need_cuda(feat, rois)
        lib = _lib.load()
        if feat.shape[0] != 1:
            raise RuntimeError("roi_align: batch size 1 (as the attack scripts run)")
        feat, rois = feat.contiguous(), rois.contiguous()
        r, (_, c, h, w) = rois.shape[0], feat.shape
        out = torch.empty((r, c, pooled, pooled), device=feat.device, dtype=torch.float32)
        with _op("roi_align_fwd", 1, 4 * (feat.numel() + out.numel())):
            check(lib.b2_roi_align_fwd(_p(feat), _p(rois), _p(out), r, c, h, w, pooled, float(scale), _stream()),
                  "roi_align_fwd")
        ctx.save_for_backward(rois)
        ctx.cfg = (r, c, h, w, pooled, float(scale))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        lib = _lib.load()
        (rois,) = ctx.saved_tensors
        r, c, h, w, pooled, scale = ctx.cfg
        g, layout = _roi_gout(gout, c)
        gfeat = torch.empty((1, c, h, w), device=g.device, dtype=torch.float32)
        with _op("roi_align_bwd", 1, 4 * (g.numel() + gfeat.numel())):
            check(lib.b2_roi_align_bwd(_p(g), _p(rois), _p(gfeat), r, c, h, w, pooled, scale, layout, _stream()),
                  "roi_align_bwd")
        return gfeat, None, None, None


def roi_align(feat, rois, pooled, scale):
    return RoIAlignFn.apply(feat, rois, pooled, scale)


class PyramidRoIAlignFn(Function):
    """``_StereoRCNN.PyramidRoI_Feat`` (attack/Stereo-RCNN/stereo_rcnn.py:110-141) in one launch per direction: FPN level
    per RoI evaluated in the kernel, output rows in the original RoI order, no host synchronisation."""

    @staticmethod
    def forward(ctx, rois, pooled, im_h, *feats):
        _need_cuda(rois, *feats)
        lib = _lib.load()
        if len(feats) != 4 or any(f.shape[0] != 1 or f.shape[1] != feats[0].shape[1] for f in feats):
            raise RuntimeError("pyramid_roi_align: four [1,C,H,W] maps (levels 2..5) expected")
        feats = [f.contiguous() for f in feats]
        rois = rois.contiguous()
        r, c = rois.shape[0], feats[0].shape[1]
        hs = (ctypes.c_int * 4)(*[f.shape[2] for f in feats])
        ws = (ctypes.c_int * 4)(*[f.shape[3] for f in feats])
        out = torch.empty((r, c, pooled, pooled), device=rois.device, dtype=torch.float32)
        ptrs = ctypes.cast(_lib.ptr_array(feats), ctypes.POINTER(ctypes.c_void_p))
        with _op("roi_align_fwd", 1, 4 * (sum(f.numel() for f in feats) + out.numel())):
            check(lib.b2_roi_align_pyramid_fwd(ptrs, hs, ws, _p(rois), _p(out), r, c, pooled, float(im_h), _stream()),
                  "roi_align_pyramid_fwd")
        ctx.save_for_backward(rois)
        ctx.cfg = (r, c, pooled, float(im_h), [tuple(f.shape) for f in feats])
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        lib = _lib.load()
        (rois,) = ctx.saved_tensors
        r, c, pooled, im_h, shapes = ctx.cfg
        g, layout = _roi_gout(gout, c)
        gfeats = [torch.empty(s, device=g.device, dtype=torch.float32) for s in shapes]
        hs = (ctypes.c_int * 4)(*[s[2] for s in shapes])
        ws = (ctypes.c_int * 4)(*[s[3] for s in shapes])
        ptrs = ctypes.cast(_lib.ptr_array(gfeats), ctypes.POINTER(ctypes.c_void_p))
        with _op("roi_align_bwd", 1, 4 * (g.numel() + sum(t.numel() for t in gfeats))):
            check(lib.b2_roi_align_pyramid_bwd(_p(g), _p(rois), ptrs, hs, ws, r, c, pooled, im_h, layout, _stream()),
                  "roi_align_pyramid_bwd")
        return (None, None, None) + tuple(gfeats)


def pyramid_roi_align(feat_maps, rois, im_h, pooled):
    return PyramidRoIAlignFn.apply(rois, pooled, im_h, *feat_maps)
