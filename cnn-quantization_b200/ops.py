"""Tensor-level wrappers over the C ABI: device pointers and the current stream come from torch, the
arithmetic all happens in libfqb200.so.  Nothing here synchronises with the host."""
import ctypes

import torch

from . import _lib as L

_workspaces = {}  # (device index, stream handle) -> uint8 tensor

# ---- optional per-launch timing (bench.py): CUDA events recorded on the launching stream around each call --------
_prof = {"on": False, "records": [], "launches": 0}


def profile_reset(enable):
    _prof["on"] = bool(enable)
    _prof["records"] = []
    _prof["launches"] = 0


def profile_collect():
    """Synchronise and return {'launches': n, 'modes': {mode: {launches, elems, bytes, ms}}}.  Modes by algorithmic
    traffic: 'D' two statistics passes + apply (16 B/elem), 'B' one statistics pass + apply (12), 'A' apply only (8),
    'S' statistics only."""
    torch.cuda.synchronize()
    modes, shapes = {}, {}
    for mode, elems, nbytes, e0, e1, tag in _prof["records"]:
        ms = e0.elapsed_time(e1)
        for table, key in ((modes, mode), (shapes, "%s %s" % (mode, tag))):
            m = table.setdefault(key, {"launches": 0, "elems": 0, "bytes": 0, "ms": 0.0})
            m["launches"] += 1
            m["elems"] += elems
            m["bytes"] += nbytes
            m["ms"] += ms
    return {"launches": _prof["launches"], "modes": modes, "shapes": shapes}


class _Timed(object):
    def __init__(self, mode, elems, bytes_per_elem, tag=""):
        self.args = (mode, elems, elems * bytes_per_elem)
        self.tag = tag

    def __enter__(self):
        _prof["launches"] += 1
        if _prof["on"]:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _prof["on"]:
            self.e1.record()
            _prof["records"].append(self.args + (self.e0, self.e1, self.tag))


def _stream_handle(device):
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise L.FqError("%s must live on a CUDA device: the fake-quantization path has no CPU implementation" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))


def _workspace(device, stream, nbytes):
    key = (device.index, stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        size = max(int(nbytes), 1 << 20)
        ws = torch.empty(size, dtype=torch.uint8, device=device)
        L.check(L.load().fqb200_workspace_init(ws.data_ptr(), ws.numel(), stream))
        _workspaces[key] = ws
    return ws


def _dense(t):
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


def cl_eligible(x, layout=None):
    """True when ``x`` is an NCHW-shaped activation stored channels-last that the flat-stream kernels take as it is:
    C % 4 == 0, C <= 2048 (``layout``, when given, must be its (N, C, H*W) view)."""
    if x.dim() != 4 or x.is_contiguous() or not x.is_contiguous(memory_format=torch.channels_last):
        return False
    n, c = x.shape[0], x.shape[1]
    if layout is not None and tuple(int(v) for v in layout) != (n, c, x.numel() // (n * c)):
        return False
    return c % 4 == 0 and 4 <= c <= 2048 and x.data_ptr() % 16 == 0


def _resolve_out(x, out):
    """(tensor the kernel writes, tensor the caller gets back).  ``x`` is the tensor handed to the kernel, i.e. AFTER any
    ``.contiguous()`` re-layout; ``out`` is the caller's output tensor or None.  The kernels write linearly in the memory
    order of ``x``, so the caller's tensor is used directly only when it has exactly that memory order (same strides);
    otherwise (e.g. an NHWC-strided ``out`` while the kernel runs on an NCHW copy) the result is produced in a scratch
    tensor and copied back element for element."""
    if out is None:
        return torch.empty_like(x), None
    if out.shape != x.shape:
        raise ValueError("out must have the shape of the input, got %r vs %r" % (tuple(out.shape), tuple(x.shape)))
    _require_cuda_f32(out, "out")
    if out.stride() == x.stride():
        return out, None
    return torch.empty_like(x), out


def _finish_out(kernel_out, user_out):
    if user_out is None:
        return kernel_out
    user_out.copy_(kernel_out)
    return user_out


def resident_ctas():
    return L.load().fqb200_resident_ctas()


def float2gemmlowp(x, range_, offset, num_bits, int_exp, enforce_true_zero, noise=None, out=None):
    """C ABI fqb200_float2gemmlowp on torch tensors (scalars by value, like the reference's pybind call)."""
    _require_cuda_f32(x, "in")
    lib = L.load()
    if noise is not None:
        _require_cuda_f32(noise, "noise")
        if noise.shape != x.shape:
            raise ValueError("noise must have the shape of the input")
    # one parameter set for the whole tensor: any dense memory order will do (no copy for channels-last activations)
    if not _dense(x) or (noise is not None and noise.stride() != x.stride()):
        x = x.contiguous()
        noise = noise.contiguous() if noise is not None else None
    kout, uout = _resolve_out(x, out)
    with torch.cuda.device(x.device), _Timed("A", x.numel(), 8):
        L.check(lib.fqb200_float2gemmlowp(x.data_ptr(), kout.data_ptr(), x.numel(), float(range_), float(offset),
                                          int(num_bits), int(bool(int_exp)), int(bool(enforce_true_zero)),
                                          noise.data_ptr() if noise is not None else None, _stream_handle(x.device)))
    return _finish_out(kout, uout)


def quantize1(x, delta, offset, num_bits, bits=None, layout=None, want_grid=False, out=None, bias=None):
    """C ABI fqb200_quantize1.  ``layout`` = (outer, groups, inner); default: [R, K] rows with per-row
    parameters when ``delta`` has R elements, else one parameter set for the whole tensor."""
    _require_cuda_f32(x, "tensor")
    lib = L.load()
    dev = x.device
    delta = torch.as_tensor(delta, dtype=torch.float32, device=dev).contiguous()
    offset = torch.as_tensor(offset, dtype=torch.float32, device=dev).contiguous()
    per_group = delta.numel() > 1 or (bits is not None)
    # channels-last activations with per-channel parameters run on the NHWC memory as it is (layout = (N, C, H*W));
    # one parameter set for the whole tensor does not care about the memory order at all
    cl = bool(per_group and layout is not None and cl_eligible(x, layout))
    if not cl and (per_group or not _dense(x)):
        x = x.contiguous()
    if layout is None:
        layout = (1, x.shape[0], x.numel() // x.shape[0]) if per_group else (1, 1, x.numel())
    outer, groups, inner = layout
    if per_group:
        if delta.numel() == 1:
            delta = delta.reshape(1).expand(groups).contiguous()
        if offset.numel() == 1:
            offset = offset.reshape(1).expand(groups).contiguous()
        if delta.numel() != groups or offset.numel() != groups:
            raise ValueError("per-group parameters must have %d elements" % groups)
    if bits is not None:
        bits = torch.as_tensor(bits, dtype=torch.float32, device=dev).contiguous()
        if bits.numel() != groups:
            raise ValueError("bit_alloc must have %d elements" % groups)
    if bias is not None:
        _require_cuda_f32(bias, "bias")
        bias = bias.contiguous()
        if bias.numel() != groups:
            raise ValueError("bias must have one element per group (%d)" % groups)
    kout, uout = _resolve_out(x, out)
    grid = torch.empty_like(x) if want_grid else None
    with torch.cuda.device(dev), _Timed("A", x.numel(), 8, "%dx%dx%d" % (outer, groups, inner)):
        L.check(lib.fqb200_quantize1(x.data_ptr(), kout.data_ptr(), grid.data_ptr() if want_grid else None,
                                     outer, groups, inner, delta.data_ptr(), offset.data_ptr(),
                                     bits.data_ptr() if bits is not None else None, int(per_group), int(num_bits),
                                     bias.data_ptr() if bias is not None else None, int(cl), _stream_handle(dev)))
    out = _finish_out(kout, uout)
    return (out, grid) if want_grid else out


def quantize1_bca(x, delta, offset, num_bits, bits=None, bias=None, relu_first=False, out=None, want_qbias=False):
    """C ABI fqb200_quantize1_bca: given-parameter quantization of a channels-last [N, C, H, W] activation with the
    reference's activation bias correction (`-bca`, inference_quantization_manager.py:180-196) in the same launch.
    ``x`` must satisfy ``cl_eligible``.  Returns the corrected quantized tensor (and the [C] corrections)."""
    _require_cuda_f32(x, "tensor")
    if not cl_eligible(x):
        raise ValueError("quantize1_bca needs a channels-last activation with C % 4 == 0, C <= 2048")
    lib = L.load()
    dev = x.device
    n, c = x.shape[0], x.shape[1]
    inner = x.numel() // (n * c)
    delta = torch.as_tensor(delta, dtype=torch.float32, device=dev).contiguous()
    offset = torch.as_tensor(offset, dtype=torch.float32, device=dev).contiguous()
    per_group = delta.numel() > 1 or bits is not None
    if per_group:
        delta = delta.reshape(-1).expand(c).contiguous() if delta.numel() == 1 else delta
        offset = offset.reshape(-1).expand(c).contiguous() if offset.numel() == 1 else offset
    if bits is not None:
        bits = torch.as_tensor(bits, dtype=torch.float32, device=dev).contiguous()
    if bias is not None:
        _require_cuda_f32(bias, "bias")
        bias = bias.contiguous()
    kout, uout = _resolve_out(x, out)
    qb = torch.empty(c, dtype=torch.float32, device=dev) if want_qbias else None
    d = L.Desc()
    d.outer, d.groups, d.inner, d.num_bits, d.channels_last = n, c, inner, 8, 1
    with torch.cuda.device(dev):
        stream = _stream_handle(dev)
        ws = _workspace(dev, stream, lib.fqb200_workspace_bytes(ctypes.byref(d)))
        with _Timed("C", x.numel(), 12, "%dx%dx%d" % (n, c, inner)):
            L.check(lib.fqb200_quantize1_bca(x.data_ptr(), kout.data_ptr(), n, c, inner, delta.data_ptr(), offset.data_ptr(),
                                             bits.data_ptr() if bits is not None else None, int(per_group), int(num_bits),
                                             bias.data_ptr() if bias is not None else None, int(bool(relu_first)),
                                             qb.data_ptr() if qb is not None else None, ws.data_ptr(), ws.numel(), stream))
    res = _finish_out(kout, uout)
    return (res, qb) if want_qbias else res


def fused(x, layout, *, scope=L.SCOPE_GROUP, range_mode=L.RANGE_MINMAX, leaf=L.LEAF_TORCH, num_bits=8,
          positive=False, solve_f64=False, clip_k=0.0, bit_alloc=False, bit_alloc_prior=L.PRIOR_STD,
          bit_alloc_round=True, bit_alloc_target=None, mt_target=0.0, mt_clip=False, bias_corr=False,
          var_corr=False, stats_only=False, want_stats=False, out=None, bias=None, bias_period=0, hist=None,
          channels_last=False, any_dense_format=False, debug_stamps=None, relu_passthrough=False, hist_offset=0,
          hist_clamped=None, residual=None, residual_relu=False, residual_stats=None, residual_bias=None, pool=None, given=None):
    """C ABI fqb200_fused: statistics -> parameters -> quantize/dequantize in one launch.

    Returns ``out`` (or ``(out, stats)`` with ``want_stats``; ``stats`` alone with ``stats_only``), where
    ``stats`` is a [groups, 12] tensor with columns ``_lib.STAT_COLUMNS``."""
    _require_cuda_f32(x, "tensor")
    lib = L.load()
    is_cl = x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
    if channels_last:
        # layout = (N, C, H*W) of a tensor stored [N][H*W][C]
        if not is_cl:
            raise ValueError("channels_last=True needs a channels-last contiguous 4-D tensor")
    elif not (any_dense_format and is_cl):
        # any_dense_format: the layout does not care about the order inside a sample (per-tensor / per-sample min-max)
        x = x.contiguous()
    dev = x.device
    outer, groups, inner = (int(v) for v in layout)
    if outer * groups * inner != x.numel():
        raise ValueError("layout %r does not cover %d elements" % (layout, x.numel()))
    d = L.Desc()
    d.outer, d.groups, d.inner = outer, groups, inner
    d.scope, d.range_mode, d.leaf = scope, range_mode, leaf
    d.num_bits, d.positive, d.solve_f64 = int(num_bits), int(bool(positive)), int(bool(solve_f64))
    d.clip_k = float(clip_k)
    d.bit_alloc, d.bit_alloc_prior, d.bit_alloc_round = int(bool(bit_alloc)), bit_alloc_prior, int(bool(bit_alloc_round))
    d.bit_alloc_target = float(bit_alloc_target if bit_alloc_target is not None else num_bits)
    d.mt_target, d.mt_clip = float(mt_target), int(bool(mt_clip))
    d.bias_corr, d.var_corr, d.stats_only = int(bool(bias_corr)), int(bool(var_corr)), int(bool(stats_only))
    d.channels_last = int(bool(channels_last))
    if bias is not None:
        _require_cuda_f32(bias, "bias")
        bias = bias.contiguous()
        want = (inner // bias_period if bias_period > 0 else -bias_period) if bias_period else groups
        if bias.numel() != want:
            raise ValueError("bias must have %d elements" % want)
        d.bias = bias.data_ptr()
        d.bias_period = int(bias_period)
    else:
        d.bias = None
        d.bias_period = 0
    d.hist_bins, d.hist_offset, d.out_hist_clamped = 0, 0, None
    if hist is not None:
        if hist.dtype != torch.int64 or not hist.is_cuda or not hist.is_contiguous() or not (1 <= hist.numel() <= 8192):
            raise ValueError("hist must be a contiguous CUDA int64 tensor of at most 8192 counters")
        d.out_hist = hist.data_ptr()
        d.hist_bins, d.hist_offset = hist.numel(), int(hist_offset)
        if hist_clamped is not None:
            if hist_clamped.dtype != torch.int64 or not hist_clamped.is_cuda or hist_clamped.numel() != 2 * groups:
                raise ValueError("hist_clamped must be a CUDA int64 tensor [groups, 2]")
            d.out_hist_clamped = hist_clamped.data_ptr()
    else:
        d.out_hist = None
    d.debug_stamps = debug_stamps.data_ptr() if debug_stamps is not None else None
    d.relu_passthrough = int(bool(relu_passthrough))
    d.residual, d.residual_relu = None, 0
    if residual is not None:
        _require_cuda_f32(residual, "residual")
        if residual.shape != x.shape or residual.stride() != x.stride():
            raise ValueError("residual needs a tensor with the input's shape and strides")
        d.residual, d.residual_relu = residual.data_ptr(), int(bool(residual_relu))
    d.residual_stats, d.residual_bias = None, None
    if residual_stats is not None:
        # the [groups, 12] table a stats_only launch exported for the residual tensor: it is quantized on the fly
        if residual is None or residual_stats.dtype != torch.float32 or not residual_stats.is_cuda or not residual_stats.is_contiguous() \
                or residual_stats.dim() != 2 or residual_stats.shape[1] != L.STATS_STRIDE:
            raise ValueError("residual_stats: the contiguous CUDA [groups, %d] table of a stats_only launch, with a residual" % L.STATS_STRIDE)
        d.residual_stats = residual_stats.data_ptr()
        if residual_bias is not None:
            _require_cuda_f32(residual_bias, "residual_bias")
            residual_bias = residual_bias.contiguous()
            if bias is None or residual_bias.numel() != bias.numel():
                raise ValueError("residual_bias must have the form of `bias`")
            d.residual_bias = residual_bias.data_ptr()
    elif residual_bias is not None:
        raise ValueError("residual_bias needs residual_stats")
    d.given_delta, d.given_offset, d.given_bits = None, None, None
    if range_mode == L.RANGE_GIVEN:
        # parameters from the caller (`-sm use`): per-group delta / offset (/ bits) device vectors, no statistics phases
        if given is None or not channels_last or stats_only or want_stats:
            raise ValueError("RANGE_GIVEN needs given=(delta, offset, bits), channels_last=True and no statistics outputs")
        gd, go, gb = given
        keep = []
        for name, v in (("delta", gd), ("offset", go), ("bits", gb)):
            if v is None:
                keep.append(None)
                continue
            _require_cuda_f32(v, "given " + name)
            v = v.contiguous()
            if v.numel() != groups:
                raise ValueError("given %s must have %d elements" % (name, groups))
            keep.append(v)
        if keep[0] is None or keep[1] is None:
            raise ValueError("RANGE_GIVEN needs delta and offset")
        d.given_delta, d.given_offset = keep[0].data_ptr(), keep[1].data_ptr()
        d.given_bits = keep[2].data_ptr() if keep[2] is not None else None
    d.pool, d.pool_h, d.pool_w, d.pool_out = 0, 0, 0, None
    pooled = None
    if pool is not None:
        # a 2x2 / stride-2 max pooling (floor mode) follows and is the only consumer: computed inside the apply phase
        # ((3, 3): stride 2, padding 1, H and W even - the ResNet stem)
        # (the per-sample / per-tensor min-max launches take it on channels-last memory when a channel-fastest bias tells
        # them the channel count)
        rows_cl = any_dense_format and is_cl and bias is not None and bias_period < 0
        if (tuple(pool) not in ((2, 2), (3, 3)) or not (channels_last or rows_cl) or stats_only or residual is not None or hist is not None
                or x.shape[3] % 2 or (tuple(pool) == (3, 3) and x.shape[2] % 2)):
            raise ValueError("pool=(2, 2) / (3, 3) needs a channels-last launch with an even W (3x3: and H) and no residual / histogram")
        n_, c_, h_, w_ = x.shape
        pooled = torch.empty((n_, c_, h_ // 2, w_ // 2), dtype=x.dtype, device=dev, memory_format=torch.channels_last)
        d.pool, d.pool_h, d.pool_w, d.pool_out = int(pool[0]), h_, w_, pooled.data_ptr()
    stats = None
    if want_stats or stats_only:
        stats = torch.zeros((groups, L.STATS_STRIDE), dtype=torch.float32, device=dev)
        d.out_stats = stats.data_ptr()
    else:
        d.out_stats = None
    if x.numel() == 0:
        res = pooled if pooled is not None else x.clone()
        return stats if stats_only else ((res, stats) if want_stats else res)
    kout, uout = (None, None) if (stats_only or pooled is not None) else _resolve_out(x, out)
    with torch.cuda.device(dev):
        stream = _stream_handle(dev)
        need = lib.fqb200_workspace_bytes(ctypes.byref(d))
        if need == 0:
            L.check(L.ERR_INVALID if not lib.fqb200_last_error() else L.ERR_INVALID)
        ws = _workspace(dev, stream, need)
        two_pass = (range_mode != L.RANGE_MINMAX or leaf == L.LEAF_MIDTREAD or var_corr or stats_only or
                    (bit_alloc and num_bits <= 4 and scope == L.SCOPE_GROUP))
        mode = "S" if stats_only else ("D" if two_pass else "B")
        bpe = (8 if two_pass else 4) + (0 if stats_only else 8)
        if range_mode == L.RANGE_GIVEN:
            mode, bpe = "A", 8
        if residual is not None:   # + the residual read of the fused block epilogue (the write is the apply's own)
            mode, bpe = mode + "r", bpe + 4
        if pooled is not None:     # the apply phase reads x (3x3: rows twice, the second time mostly out of L2) and writes a quarter
            mode, bpe = mode + "p", bpe - 3
        with _Timed(mode, x.numel(), bpe, "%dx%dx%d" % (outer, groups, inner)):
            L.check(lib.fqb200_fused(ctypes.byref(d), x.data_ptr(), kout.data_ptr() if kout is not None else None,
                                     ws.data_ptr(), ws.numel(), stream))
    if stats_only:
        return stats
    if pooled is not None:
        return (pooled, stats) if want_stats else pooled
    out = _finish_out(kout, uout)
    return (out, stats) if want_stats else out


def add_relu_(a, b):
    """``a += b; relu_(a)`` in one pass (C ABI fqb200_add_relu, 12 instead of 20 bytes per element).  Both tensors must be
    dense with identical strides (any memory format); returns ``a``.  Bit-identical to the two torch ops."""
    _require_cuda_f32(a, "a")
    _require_cuda_f32(b, "b")
    if a.shape != b.shape or a.stride() != b.stride() or not _dense(a):
        raise ValueError("add_relu_ needs two dense tensors of identical shape and strides")
    with torch.cuda.device(a.device), _Timed("E", a.numel(), 12):
        L.check(L.load().fqb200_add_relu(a.data_ptr(), b.data_ptr(), a.data_ptr(), a.numel(), _stream_handle(a.device)))
    return a


def maxpool2d_cl(x, kernel_size, stride, padding):
    """``F.max_pool2d(x, kernel_size, stride, padding)`` for a channels-last fp32 activation with C % 4 == 0 (C ABI
    fqb200_maxpool2d_nhwc); the result is channels-last as well.  Bit-identical to torch."""
    _require_cuda_f32(x, "input")
    if x.dim() != 4 or not x.is_contiguous(memory_format=torch.channels_last) or x.shape[1] % 4 != 0:
        raise ValueError("maxpool2d_cl needs a channels-last [N, C, H, W] tensor with C % 4 == 0")
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    ph, pw = (padding, padding) if isinstance(padding, int) else padding
    n, c, h, w = x.shape
    oh, ow = (h + 2 * ph - kh) // sh + 1, (w + 2 * pw - kw) // sw + 1
    out = torch.empty((n, c, oh, ow), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    with torch.cuda.device(x.device), _Timed("P", out.numel(), 4 + 4 * sh * sw):
        L.check(L.load().fqb200_maxpool2d_nhwc(x.data_ptr(), out.data_ptr(), n, h, w, c, kh, kw, sh, sw, ph, pw,
                                               _stream_handle(x.device)))
    return out


def _test_division(a, b):
    """(fast, ieee) quotients from the device: the 3-instruction exact division next to __fdiv_rn."""
    lib = L.load()
    fast, ieee = torch.empty_like(a), torch.empty_like(a)
    with torch.cuda.device(a.device):
        L.check(lib.fqb200_selftest_division(a.data_ptr(), b.data_ptr(), fast.data_ptr(), ieee.data_ptr(), a.numel(),
                                         _stream_handle(a.device)))
    return fast, ieee
