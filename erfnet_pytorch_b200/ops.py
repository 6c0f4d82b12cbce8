"""Thin torch-tensor wrappers over the C ABI.  torch is used for device memory and streams only; every
arithmetic op below is a kernel in liberfnet_b200.so.  Internal activations are NHWC tensors [N,H,W,C]
(contiguous) in fp32 (check mode) or bf16 (fast path)."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ERF_BF16, ERF_F32, ConvEpilogue, ConvGeom, check

DETAIL = False    # tools/profile_step.py: per-shape kernel names in the timing table
USE_TC = True     # tests flip this to compare the tcgen05 kernels with the generic CUDA-core path
EMULATED = False  # set by tests/emu.py only (CPU emulation of the kernel contracts for host-logic tests)
LAUNCHES = 0  # number of kernel-launching library calls made by this process (bench.py reports it)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


_TIMING = None  # list of (name, flops, start_event, end_event) while bench.py profiles one step


def start_kernel_timing():
    global _TIMING
    _TIMING = []


def stop_kernel_timing():
    """-> {kernel name: {"ms": total, "n": launches, "flops": total algorithmic flops}} (synchronises)."""
    global _TIMING
    recs, _TIMING = _TIMING or [], None
    torch.cuda.synchronize()
    out = {}
    for name, flops, e0, e1 in recs:
        d = out.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0})
        d["ms"] += e0.elapsed_time(e1)
        d["n"] += 1
        d["flops"] += flops
    return out


class _timed:
    def __init__(self, name, flops=0.0):
        self.name, self.flops = name, flops

    def __enter__(self):
        if _TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if _TIMING is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _TIMING.append((self.name, self.flops, self.e0, e1))


def conv_flops(spec, n, ho, wo):
    taps = spec.kh * spec.kw / (spec.stride ** 2 if spec.transposed else 1)
    return 2.0 * n * ho * wo * spec.cout * spec.cin * taps


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return ERF_F32
    if t.dtype == torch.bfloat16:
        return ERF_BF16
    raise RuntimeError(f"erfnet_pytorch_b200: unsupported activation dtype {t.dtype}")


def _stream(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError("erfnet_pytorch_b200 kernels need CUDA tensors (there is no CPU fallback)")
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def round_up(a, b):
    return (a + b - 1) // b * b


# ----------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ConvSpec:
    """Geometry of one convolution of the net in its FORWARD orientation."""
    cin: int
    cout: int
    kh: int
    kw: int
    stride: int = 1
    pad_h: int = 0
    pad_w: int = 0
    dil_h: int = 1
    dil_w: int = 1
    transposed: bool = False
    out_pad: int = 0

    def out_hw(self, h, w):
        if self.transposed:
            return ((h - 1) * self.stride - 2 * self.pad_h + self.dil_h * (self.kh - 1) + self.out_pad + 1,
                    (w - 1) * self.stride - 2 * self.pad_w + self.dil_w * (self.kw - 1) + self.out_pad + 1)
        return ((h + 2 * self.pad_h - self.dil_h * (self.kh - 1) - 1) // self.stride + 1,
                (w + 2 * self.pad_w - self.dil_w * (self.kw - 1) - 1) // self.stride + 1)

    @property
    def cin_p(self):  # padded input channels the kernels see
        return round_up(self.cin, 4)

    @property
    def w_ld(self):
        return round_up(self.cout, 8)


def pack_weight_fwd(w: torch.Tensor, spec: ConvSpec, dtype) -> torch.Tensor:
    """torch layout -> [(ky*KW+kx)*cin_p + ci][w_ld] for the forward gather (Conv2d [Co,Ci,kh,kw]; ConvT [Ci,Co,kh,kw])."""
    t = w.permute(2, 3, 0, 1) if spec.transposed else w.permute(2, 3, 1, 0)  # [kh,kw,ci,co]
    out = torch.zeros(spec.kh, spec.kw, spec.cin_p, spec.w_ld, device=w.device, dtype=dtype)
    out[:, :, :spec.cin, :spec.cout] = t
    return out.reshape(-1, spec.w_ld)


def dgrad_spec(spec: ConvSpec) -> ConvSpec:
    """The data gradient of a conv is the transposed gather with the same (stride, pad, dilation) and vice versa."""
    return ConvSpec(spec.cout, spec.cin_p, spec.kh, spec.kw, spec.stride, spec.pad_h, spec.pad_w, spec.dil_h, spec.dil_w,
                    not spec.transposed, 0)


def pack_weight_dgrad(w: torch.Tensor, spec: ConvSpec, dtype) -> torch.Tensor:
    """[(ky*KW+kx)*round4(cout) + co][round8(cin_p)] operand of the dgrad gather."""
    t = w.permute(2, 3, 1, 0) if spec.transposed else w.permute(2, 3, 0, 1)  # [kh,kw,co,ci]
    d = dgrad_spec(spec)
    out = torch.zeros(spec.kh, spec.kw, d.cin_p, d.w_ld, device=w.device, dtype=dtype)
    out[:, :, :spec.cout, :spec.cin] = t
    return out.reshape(-1, d.w_ld)


def unpack_wgrad(dw: torch.Tensor, spec: ConvSpec) -> torch.Tensor:
    t = dw.view(spec.kh, spec.kw, spec.cin_p, spec.w_ld)[:, :, :spec.cin, :spec.cout]
    return (t.permute(2, 3, 0, 1) if spec.transposed else t.permute(3, 2, 0, 1)).contiguous()


def _geom(spec: ConvSpec, x: torch.Tensor, out: torch.Tensor, out_coff: int, cin_used: Optional[int] = None) -> ConvGeom:
    n, h, w, ld = x.shape
    _, ho, wo, old = out.shape
    return ConvGeom(n, h, w, cin_used if cin_used is not None else spec.cin_p, ld, ho, wo, spec.cout, old, out_coff,
                    spec.kh, spec.kw, spec.stride, spec.pad_h, spec.pad_w, spec.dil_h, spec.dil_w, int(spec.transposed), spec.w_ld)


def conv_fwd(x: torch.Tensor, wp: torch.Tensor, spec: ConvSpec, bias: Optional[torch.Tensor] = None, relu=False,
             out: Optional[torch.Tensor] = None, out_coff=0, mask: Optional[torch.Tensor] = None,
             add: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None, out_hw: Optional[Tuple[int, int]] = None):
    """Generic gather conv (include/erfnet_b200.h erf_conv_fwd).  x: [N,H,W,ld>=cin_p]; returns/fills out [N,Ho,Wo,ldo]."""
    n, h, w, _ = x.shape
    if out is None:
        ho, wo = out_hw if out_hw is not None else spec.out_hw(h, w)
        out = torch.empty(n, ho, wo, spec.cout, device=x.device, dtype=x.dtype)
    g = _geom(spec, x, out, out_coff)
    e = ConvEpilogue(_p(bias), int(relu), _p(mask), 0 if mask is None else mask.shape[-1], _p(add), 0 if add is None else add.shape[-1],
                     _p(stats), 0 if stats is None else stats.numel() // 2)
    with _timed(f"conv_gather_kernel[{spec.cin}->{spec.cout} {spec.kh}x{spec.kw} s{spec.stride}{'T' if spec.transposed else ''}]" if DETAIL else "conv_gather_kernel",
                conv_flops(spec, n, out.shape[1], out.shape[2])):
        check(_lib.lib().erf_conv_fwd(_p(x), _p(wp), _p(out), C.byref(g), C.byref(e), dt(x), _stream(x)))
    _count()
    return out


def conv_wgrad(x: torch.Tensor, dout: torch.Tensor, spec: ConvSpec, dout_coff=0, dw_out=None, db_out=None, want_bias=True):
    """Weight / bias gradient in the TORCH parameter layout, accumulated into dw_out / db_out (e.g. views of a flat
    gradient buffer) or into fresh zero tensors.  Returns (dw, db)."""
    if dw_out is None:
        shape = (spec.cin, spec.cout, spec.kh, spec.kw) if spec.transposed else (spec.cout, spec.cin, spec.kh, spec.kw)
        dw_out = torch.zeros(shape, device=x.device, dtype=torch.float32)
    if db_out is None and want_bias:
        db_out = torch.zeros(spec.cout, device=x.device, dtype=torch.float32)
    g = _geom(spec, x, dout, dout_coff)
    name = f"conv_wgrad_kernel[{spec.cin}->{spec.cout} {spec.kh}x{spec.kw} s{spec.stride}{'T' if spec.transposed else ''}]" if DETAIL else "conv_wgrad_kernel"
    with _timed(name, conv_flops(spec, x.shape[0], dout.shape[1], dout.shape[2])):
        check(_lib.lib().erf_conv_wgrad_torch(_p(x), _p(dout), _p(dw_out), _p(db_out) if want_bias else None, C.byref(g), spec.cin, dt(x), _stream(x)))
    _count(2)
    return dw_out, db_out


def conv1d_tc_supported(c: int, w: int, dil: int, dtype) -> bool:
    return dtype == torch.bfloat16 and not EMULATED and bool(_lib.lib().erf_conv1d_tc_supported(c, w, dil))


def pack_conv1d_tc(w: torch.Tensor, dgrad: bool = False) -> torch.Tensor:
    """Conv2d weight [Co,Ci,3,1] or [Co,Ci,1,3] (fp32, BN already folded) -> bf16 shared-memory image
    [3][C/64][C][64] for erf_conv1d_tc: tile (t,kc) row n holds Wt[t][n][kc*64 .. +64] with its eight 16-byte chunks
    XOR-swizzled by (n & 7).  dgrad=True packs the data-gradient operand (taps flipped, channels transposed)."""
    co, ci = w.shape[:2]
    wt = w.reshape(co, ci, 3).permute(2, 0, 1)  # [t][co][ci]
    if dgrad:
        wt = wt.flip(0).transpose(1, 2)         # [t'][ci][co] = W[co,ci,2-t']
    c = wt.shape[1]
    ch = min(c, 64)          # channels per K chunk = one swizzle row (128 B, or 32 B for C = 16)
    cpr = ch // 8            # 16-byte chunks per row
    img = wt.reshape(3, c, c // ch, cpr, 8).permute(0, 2, 1, 3, 4)  # [t][kc][n][j][e]
    n = torch.arange(c, device=w.device)
    key = (n & 7) if cpr == 8 else ((n >> 2) & 1)
    perm = torch.arange(cpr, device=w.device)[None, :] ^ key[:, None]  # [n][jj] -> source chunk
    idx = perm[None, None, :, :, None].expand(3, c // ch, c, cpr, 8)
    return img.gather(3, idx).contiguous().to(torch.bfloat16)


def pack_conv1d_tc_pair(w: torch.Tensor):
    """(forward image, data-gradient image) of a Conv2d weight [C,C,3,1]/[C,C,1,3] in one kernel launch."""
    c = w.shape[0]
    w = w.detach().contiguous().float()
    imgs = torch.empty(2, 3 * c * c, device=w.device, dtype=torch.bfloat16)
    check(_lib.lib().erf_pack_conv1d_tc(_p(w), c, _p(imgs[0]), _p(imgs[1]), _stream(w)))
    _count()
    return imgs[0], imgs[1]


def conv1d_pair_tc_supported(c: int, w: int, dil: int, dtype) -> bool:
    return dtype == torch.bfloat16 and not EMULATED and USE_TC and USE_PAIR and bool(_lib.lib().erf_conv1d_pair_tc_supported(c, w, dil))


FORK_JOIN = os.environ.get("ERF_FORK_JOIN", "1") != "0"
_FORK_STREAMS = {}


def side_stream(device):
    s = _FORK_STREAMS.get(device)
    if s is None:
        s = _FORK_STREAMS[device] = torch.cuda.Stream(device=device)
    return s


def fork_join(side_fn, main_fn):
    """main_fn() and side_fn() launch kernels that are independent of each other (they write disjoint memory).  While a CUDA graph is being captured
    side_fn's launches go to a second stream, forked from and joined to the capturing stream, so that the replayed graph runs them side by side (the
    max-pool half of a downsampler next to its convolution half: two kernels that fit on an SM together).  Eagerly the two simply run in order."""
    if not FORK_JOIN or EMULATED or not torch.cuda.is_available() or not torch.cuda.is_current_stream_capturing():
        main_fn()
        side_fn()
        return
    main = torch.cuda.current_stream()
    s = side_stream(main.device)
    s.wait_stream(main)
    with torch.cuda.stream(s):
        side_fn()
    main_fn()
    main.wait_stream(s)


USE_PAIR = True   # fused 3x1 -> ReLU -> 1x3 pairs in eval mode (C = 16 / 64)
FOLD_32 = True          # 32-channel BN-free blocks run as [W/2, 64]-folded views on the 64-channel tcgen05 kernels
FUSE_NCHW_HEAD = True   # inference: the decoder head writes the NCHW fp32 logits from its epilogue (no NHWC logits, no layout pass)


STATIC_WEIGHT_PREFETCH = True   # inference: cached weight images are fetched before the PDL wait (bit 1 of the kernels' relu argument)


def conv1d_pair_tc(x, w1_img, b1, w2_img, b2, dil, relu=True, res=None, static_w=False):
    """act(conv1x3_d(relu(conv3x1_d(x) + b1)) + b2 [+ res]) in one launch; bit-identical to two conv1d_tc calls.
    static_w: no kernel in flight writes the weight images (cached eval-mode images), so their copy may overlap the previous kernel's tail."""
    n, h, w, c = x.shape
    out = torch.empty_like(x)
    with _timed("conv1d_pair_tc_kernel", 2.0 * 2.0 * n * h * w * c * c * 3):
        check(_lib.lib().erf_conv1d_pair_tc(_p(x), _p(w1_img), _p(b1), _p(w2_img), _p(b2), _p(out), _p(res), n, h, w, c, dil,
                                            int(relu) | (2 if static_w else 0), _stream(x)))
    _count()
    return out


class Conv1dPackTable:
    """Device table for erf_pack_conv1d_tc_batched: every listed Conv2d weight [C,C,3,1]/[C,C,1,3] (fp32, contiguous, fixed
    address -- e.g. views of a flat parameter buffer) is re-packed into persistent (forward, data-gradient) images by ONE launch."""

    def __init__(self, weights):
        import struct
        dev = weights[0].device
        sizes = [3 * w.shape[0] * w.shape[0] for w in weights]
        self.images = torch.empty(2 * sum(sizes), device=dev, dtype=torch.bfloat16)
        self.pairs, rows, off = [], b"", 0
        for w, n in zip(weights, sizes):
            if w.dtype != torch.float32 or not w.is_contiguous():
                raise RuntimeError("Conv1dPackTable: weights must be contiguous fp32")
            f, d = self.images[off:off + n], self.images[off + n:off + 2 * n]
            off += 2 * n
            self.pairs.append((f, d))
            rows += struct.pack("<QQQii", w.data_ptr(), f.data_ptr(), d.data_ptr(), w.shape[0], 0)
        self._ptrs = [w.data_ptr() for w in weights]
        self._weights = weights
        self.table = torch.frombuffer(bytearray(rows), dtype=torch.uint8).to(dev)
        self.n, self.max_c = len(weights), max(w.shape[0] for w in weights)

    def valid(self):
        return all(w.data_ptr() == p for w, p in zip(self._weights, self._ptrs))

    def run(self):
        check(_lib.lib().erf_pack_conv1d_tc_batched(_p(self.table), self.n, self.max_c, _stream(self.table)))
        _count()


CONV1D_CALL_LOG = None   # bench.py: a list that receives the arguments of every conv1d_tc call (to re-issue the step's launches back to back)


def conv1d_tc(x: torch.Tensor, w_img: torch.Tensor, bias: Optional[torch.Tensor], horizontal: bool, dil: int, relu=False,
              res: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
              colsum: Optional[torch.Tensor] = None, static_w: bool = False) -> torch.Tensor:
    """stats (fp64 [2C]) / colsum (fp32 [C]) are accumulated by the kernel itself (Gram / ones MMAs over the staged output tile).
    static_w: see conv1d_pair_tc."""
    n, h, w, c = x.shape
    if CONV1D_CALL_LOG is not None:
        CONV1D_CALL_LOG.append((x, w_img, bias, horizontal, dil, relu, res, mask, stats, colsum, static_w))
    out = torch.empty_like(x)
    fused = True
    name = (f"conv1d_tc[C={c} {'H' if horizontal else 'V'} d={dil}{' +res' if res is not None else ''}{' +mask' if mask is not None else ''}"
            f"{' +stats' if stats is not None else ''}{' +colsum' if colsum is not None else ''}]") if DETAIL else "conv1d_tc_kernel"
    with _timed(name, 2.0 * n * h * w * c * c * 3):
        check(_lib.lib().erf_conv1d_tc(_p(x), _p(w_img), _p(bias), _p(out), _p(res), _p(mask), n, h, w, c, int(horizontal), dil, int(relu) | (2 if static_w else 0),
                                       _p(stats) if fused else None, _p(colsum) if fused else None, _stream(x)))
    _count()
    if not fused:
        if stats is not None:
            channel_stats(out, stats)
        if colsum is not None:
            channel_sum(out, colsum)
    return out


# ----------------------------------------------------------------------------------------------------------
# generic tcgen05 conv (software im2col): planning, weight images, launch
# ----------------------------------------------------------------------------------------------------------
def im2col_supported(spec_g: ConvSpec, dtype) -> bool:
    """spec_g: the GATHER that is executed (a forward spec, or dgrad_spec(forward spec))."""
    return (dtype == torch.bfloat16 and not EMULATED and USE_TC and spec_g.cin % 16 == 0 and spec_g.cin <= 128 and spec_g.cout <= 128
            and spec_g.stride in (1, 2) and spec_g.kh * spec_g.kw <= 9)


MERGE_ROW_PARITY = True   # stride-2 transposed gathers: fold the two column parities into N (whole-row writes, 2 launches not 4)


def im2col_plan(spec_g: ConvSpec, row_parity: Optional[bool] = None):
    """One dict per launch (class of output pixels sharing a tap list, and an N range).  row_parity (default: MERGE_ROW_PARITY)
    selects the whole-row form of stride-2 transposed gathers, which needs a dense output tensor of round8(cout) channels."""
    s = spec_g.stride
    classes = []
    if spec_g.transposed and s == 2 and (MERGE_ROW_PARITY if row_parity is None else row_parity):
        return _im2col_plan_row_parity(spec_g)
    if not spec_g.transposed:
        taps = [(ky, kx, ky * spec_g.dil_h - spec_g.pad_h, kx * spec_g.dil_w - spec_g.pad_w) for ky in range(spec_g.kh) for kx in range(spec_g.kw)]
        classes.append(dict(taps=taps, in_step=s, out_step=1, oy=0, ox=0, div=1))
    else:
        for a in range(s):
            for b in range(s):
                taps = []
                for ky in range(spec_g.kh):
                    ty = a + spec_g.pad_h - ky * spec_g.dil_h
                    if ty % s:
                        continue
                    for kx in range(spec_g.kw):
                        tx = b + spec_g.pad_w - kx * spec_g.dil_w
                        if tx % s:
                            continue
                        taps.append((ky, kx, ty // s, tx // s))
                if taps:
                    classes.append(dict(taps=taps, in_step=1, out_step=s, oy=a, ox=b, div=s))
    np_full = 16
    while np_full < spec_g.cout:
        np_full *= 2
    plan = []
    for c in classes:
        npad = np_full
        while len(c["taps"]) * spec_g.cin * npad * 2 > 96 * 1024 and npad > 16:
            npad //= 2
        for n0 in range(0, np_full, npad):
            if n0 >= spec_g.cout:
                break
            plan.append(dict(c, n0=n0, NP=npad, n_valid=min(round_up(spec_g.cout, 8) - n0, npad)))
    return plan


def _im2col_plan_row_parity(spec_g: ConvSpec, split: Optional[int] = None):
    """Stride-2 transposed gather in row-parity form: one class per output row parity a; its N axis is (b, c) = (column parity,
    channel), so pixel (y, x) of the input grid writes the 2*cout8 contiguous values of output row 2y+a at columns 2x, 2x+1.
    A tap is an input offset (dy, dx) plus, per column parity, the kernel element that multiplies it (None = no contribution).
    split: N offset of column parity 1 (default round8(cout); the NCHW epilogue wants a power of two = NP / 2)."""
    c8 = round_up(spec_g.cout, 8) if split is None else split
    plan = []
    for a in range(2):
        taps = {}
        for b in range(2):
            for ky in range(spec_g.kh):
                ty = a + spec_g.pad_h - ky * spec_g.dil_h
                if ty % 2:
                    continue
                for kx in range(spec_g.kw):
                    tx = b + spec_g.pad_w - kx * spec_g.dil_w
                    if tx % 2:
                        continue
                    taps.setdefault((ty // 2, tx // 2), [None, None])[b] = (ky, kx)
        if not taps:
            continue
        tl = [(k0, k1, dy, dx) for (dy, dx), (k0, k1) in sorted(taps.items())]
        n_total = 2 * c8
        np_full = 16
        while np_full < n_total:
            np_full *= 2
        npad = min(np_full, 128)
        while len(tl) * round_up(spec_g.cin, 16) * npad * 2 > 96 * 1024 and npad > 16:
            npad //= 2
        for n0 in range(0, n_total, npad):
            # a chunk that lies inside one column parity only needs the taps feeding that parity
            if n0 + npad <= c8:
                tc = [t for t in tl if t[0] is not None]
            elif n0 >= c8:
                tc = [t for t in tl if t[1] is not None]
            else:
                tc = tl
            plan.append(dict(taps=tc, in_step=1, out_step=2, oy=a, ox=0, div=2, n0=n0, NP=npad, n_valid=min(n_total - n0, npad), merged=c8))
    return plan


def im2col_pack(w: torch.Tensor, spec_fwd: ConvSpec, dgrad: bool, plan):
    """bf16 weight images (one per plan entry) from the fp32 torch weight of the FORWARD conv spec_fwd."""
    w = w.detach().contiguous().float()
    khw = spec_fwd.kh * spec_fwd.kw
    d0, d1 = w.shape[0], w.shape[1]          # Conv2d: [Co,Ci,..]; ConvTranspose2d: [Ci,Co,..]
    co_is_dim0 = not spec_fwd.transposed
    n_is_co = not dgrad                       # forward: n = out channel; data gradient: n = in channel
    n_dim0 = (n_is_co == co_is_dim0)
    stride_n, stride_k = (d1 * khw, khw) if n_dim0 else (khw, d1 * khw)
    n_src, k_src = (d0, d1) if n_dim0 else (d1, d0)
    cin_g = round_up(k_src, 16)
    L = _lib.lib()
    imgs = []
    for e in plan:
        taps = _lib.Im2colTaps()
        taps.ntaps = len(e["taps"])
        if e.get("merged"):
            taps.n_split = e["merged"]
            for i, (k0, k1, _, _) in enumerate(e["taps"]):
                taps.off[i] = -1 if k0 is None else k0[0] * spec_fwd.kw + k0[1]
                taps.off2[i] = -1 if k1 is None else k1[0] * spec_fwd.kw + k1[1]
        else:
            for i, (ky, kx, _, _) in enumerate(e["taps"]):
                taps.off[i] = ky * spec_fwd.kw + kx
        img = torch.empty(len(e["taps"]) * cin_g * e["NP"], device=w.device, dtype=torch.bfloat16)
        check(L.erf_pack_im2col_weights(_p(w), stride_n, stride_k, e["n0"], n_src, k_src, C.byref(taps), cin_g, e["NP"], _p(img), _stream(w)))
        _count()
        imgs.append(img)
    return imgs


def conv_im2col_tc(x: torch.Tensor, imgs, plan, spec_g: ConvSpec, bias: Optional[torch.Tensor] = None, relu=False,
                   out: Optional[torch.Tensor] = None, out_coff=0, out_hw: Optional[Tuple[int, int]] = None):
    """Run the gather spec_g on the tensor cores.  x: [N,H,W,ld >= cin]; fills out[..., out_coff : out_coff + round8(cout)]."""
    n, h, w, ld = x.shape
    if out is None:
        ho, wo = out_hw if out_hw is not None else spec_g.out_hw(h, w)
        out = torch.empty(n, ho, wo, round_up(spec_g.cout, 8), device=x.device, dtype=x.dtype)
    ho, wo = out.shape[1:3]
    if bias is not None and bias.numel() % 8:
        bias = torch.nn.functional.pad(bias.float(), (0, 8 - bias.numel() % 8))  # the kernel adds bias to the padded channels too
    L = _lib.lib()
    name = f"conv_im2col_tc[{spec_g.cin}->{spec_g.cout} {spec_g.kh}x{spec_g.kw} s{spec_g.stride}{'T' if spec_g.transposed else ''}]" if DETAIL else "conv_im2col_tc_kernel"
    with _timed(name, conv_flops(spec_g, n, ho, wo)):
        for e, img in zip(plan, imgs):
            d = _lib.Im2colDesc()
            d.in_, d.H, d.W, d.in_ld, d.Cin = x.data_ptr(), h, w, ld, round_up(spec_g.cin, 16)
            # a parity class (oy, ox) of a stride-2 transposed gather owns ceil((ho - oy) / 2) x ceil((wo - ox) / 2) output pixels (odd sizes: the
            # data gradient of a stride-2 conv over an odd-sized input)
            d.EN, d.EH, d.EW, d.in_step = n, (ho - e["oy"] + e["div"] - 1) // e["div"], (wo - e["ox"] + e["div"] - 1) // e["div"], e["in_step"]
            if d.EH <= 0 or d.EW <= 0:
                continue
            d.ntaps = len(e["taps"])
            for i, (_, _, dy, dx) in enumerate(e["taps"]):
                d.tap_dy[i], d.tap_dx[i] = dy, dx
            d.w_img = img.data_ptr()
            if e.get("merged"):
                if out.shape[-1] != e["merged"] or out_coff != 0 or wo % 2:
                    raise RuntimeError(f"conv_im2col_tc: row-parity form needs a dense output of {e['merged']} channels (got ld {out.shape[-1]}, offset {out_coff})")
                d.bias = None if bias is None else bias.data_ptr()
                d.bias_period = e["merged"]
                d.out, d.Ho, d.Wo, d.out_ld, d.out_c0 = out.data_ptr(), ho, wo // 2, 2 * out.shape[-1], e["n0"]
                d.out_step, d.out_step_x, d.out_oy, d.out_ox = 2, 1, e["oy"], 0
            else:
                d.bias = None if bias is None else bias.data_ptr() + 4 * e["n0"]
                d.out, d.Ho, d.Wo, d.out_ld, d.out_c0 = out.data_ptr(), ho, wo, out.shape[-1], out_coff + e["n0"]
                d.out_step, d.out_oy, d.out_ox = e["out_step"], e["oy"], e["ox"]
            d.NP, d.n_valid, d.relu = e["NP"], e["n_valid"], int(relu)
            check(L.erf_conv_im2col_tc(C.byref(d), _stream(x)))
            _count()
    return out


def im2col_nchw_supported(spec_g: ConvSpec, dtype) -> bool:
    """The fused NCHW fp32 epilogue: stride-2 transposed convs in row-parity form whose two column parities fit N <= 64."""
    return (im2col_supported(spec_g, dtype) and spec_g.transposed and spec_g.stride == 2 and spec_g.cout <= 32
            and len(_im2col_plan_row_parity(spec_g, split=32 if spec_g.cout > 16 else 16)) == 2)


def im2col_plan_nchw(spec_g: ConvSpec):
    return _im2col_plan_row_parity(spec_g, split=32 if spec_g.cout > 16 else 16)


def conv_im2col_tc_nchw(x: torch.Tensor, imgs, plan, spec_g: ConvSpec, bias: Optional[torch.Tensor] = None, relu=False) -> torch.Tensor:
    """Like conv_im2col_tc for a plan of im2col_plan_nchw, but the epilogue writes the reference's NCHW fp32 tensor [N,cout,Ho,Wo] directly
    (no NHWC bf16 logits, no layout pass): what eval/erfnet.py:132 returns."""
    n, h, w, ld = x.shape
    ho, wo = spec_g.out_hw(h, w)
    out = torch.empty(n, spec_g.cout, ho, wo, device=x.device, dtype=torch.float32)
    split = plan[0]["merged"]
    if bias is not None:
        bias = torch.nn.functional.pad(bias.float()[:spec_g.cout], (0, split - spec_g.cout)).contiguous()
    L = _lib.lib()
    with _timed("conv_im2col_tc_kernel", conv_flops(spec_g, n, ho, wo)):
        for e, img in zip(plan, imgs):
            d = _lib.Im2colDesc()
            d.in_, d.H, d.W, d.in_ld, d.Cin = x.data_ptr(), h, w, ld, round_up(spec_g.cin, 16)
            d.EN, d.EH, d.EW, d.in_step = n, ho // 2, wo // 2, 1
            d.ntaps = len(e["taps"])
            for i, (_, _, dy, dx) in enumerate(e["taps"]):
                d.tap_dy[i], d.tap_dx[i] = dy, dx
            d.w_img = img.data_ptr()
            d.bias = None if bias is None else bias.data_ptr()
            d.bias_period = split
            d.out, d.Ho, d.Wo, d.out_ld, d.out_c0 = out.data_ptr(), ho, wo // 2, 2 * split, 0
            d.out_step, d.out_step_x, d.out_oy, d.out_ox = 2, 1, e["oy"], 0
            d.NP, d.n_valid, d.relu = e["NP"], e["NP"], int(relu)
            d.nchw_out, d.nchw_C = out.data_ptr(), spec_g.cout
            check(L.erf_conv_im2col_tc(C.byref(d), _stream(x)))
            _count()
    return out


def convT2x2_dgrad_tc_supported(spec: ConvSpec, dy: torch.Tensor) -> bool:
    """Data gradient of a non-overlapping ConvTranspose2d(k=2, s=2, p=0) (the decoder head, eval/erfnet.py:124) whose output
    channel count is not a multiple of 16: possible on the tensor cores through the row-folded view below."""
    ld = dy.shape[-1]
    return (dy.dtype == torch.bfloat16 and not EMULATED and USE_TC and spec.transposed and spec.kh == 2 and spec.kw == 2 and spec.stride == 2
            and spec.pad_h == 0 and spec.pad_w == 0 and spec.cin % 8 == 0 and spec.cin <= 128 and ld == round_up(spec.cout, 8)
            and (2 * ld) % 16 == 0 and 2 * ld <= 128 and dy.shape[2] % 2 == 0)


def convT2x2_dgrad_tc(dy: torch.Tensor, w: torch.Tensor, spec: ConvSpec) -> torch.Tensor:
    """dx[i,j,ci] = sum_{a,b,c} dy[2i+a, 2j+b, c] * w[ci,c,a,b].  dy [N,Ho,Wo,ld] is read as [N,Ho,Wo/2,2*ld]: the pixel pair
    (2j, 2j+1) of a row is one pixel with channels (b, c), so the gather is a 2-tap conv (rows 2i, 2i+1; row step 2, column step 1)
    with K = 2 * 2*ld -- a multiple of 16 even though cout (20) is not."""
    n, ho, wo, ld = dy.shape
    k2 = 2 * ld
    # fp32 operand [a][(b,c)][ci], zero in the pad channels
    wk = torch.zeros(2, 2, ld, spec.cin, device=w.device, dtype=torch.float32)
    wk[:, :, :spec.cout] = w.detach().float().permute(2, 3, 1, 0)
    taps = _lib.Im2colTaps()
    taps.ntaps = 2
    taps.off[0], taps.off[1] = 0, k2 * spec.cin
    npad = 16
    while npad < spec.cin:
        npad *= 2
    cin_g = round_up(k2, 16)
    img = torch.empty(2 * cin_g * npad, device=w.device, dtype=torch.bfloat16)
    L = _lib.lib()
    check(L.erf_pack_im2col_weights(_p(wk), 1, spec.cin, 0, spec.cin, k2, C.byref(taps), cin_g, npad, _p(img), _stream(dy)))
    dx = torch.empty(n, ho // 2, wo // 2, round_up(spec.cin, 8), device=dy.device, dtype=dy.dtype)
    d = _lib.Im2colDesc()
    d.in_, d.H, d.W, d.in_ld, d.Cin = dy.data_ptr(), ho, wo // 2, k2, cin_g
    d.EN, d.EH, d.EW, d.in_step, d.in_step_x = n, ho // 2, wo // 2, 2, 1
    d.ntaps = 2
    d.tap_dy[0], d.tap_dx[0], d.tap_dy[1], d.tap_dx[1] = 0, 0, 1, 0
    d.w_img, d.bias = img.data_ptr(), None
    d.out, d.Ho, d.Wo, d.out_ld, d.out_c0 = dx.data_ptr(), ho // 2, wo // 2, dx.shape[-1], 0
    d.out_step, d.out_oy, d.out_ox = 1, 0, 0
    d.NP, d.n_valid, d.relu = npad, dx.shape[-1], 0
    with _timed("conv_im2col_tc_kernel", 2.0 * n * ho * wo * spec.cout * spec.cin):
        check(L.erf_conv_im2col_tc(C.byref(d), _stream(dy)))
    _count(2)
    return dx


def wgrad_im2col_supported(spec: ConvSpec, dtype) -> bool:
    """spec: the FORWARD conv.  The gathered side is x for a conv and dy for a transposed conv."""
    return (dtype == torch.bfloat16 and not EMULATED and USE_TC and spec.cin <= 128 and spec.cout <= 128 and spec.stride in (1, 2)
            and spec.kh * spec.kw <= 9 and spec.dil_h == 1 and spec.dil_w == 1)


def conv_wgrad_im2col_tc(x: torch.Tensor, dy: torch.Tensor, spec: ConvSpec, dy_coff=0, dw_out=None):
    """Weight gradient (torch parameter layout, fp32, ACCUMULATED into dw_out or a fresh zero tensor) on the tensor cores."""
    n, h, w, xld = x.shape
    _, ho, wo, dyld = dy.shape
    if dw_out is None:
        shape = (spec.cin, spec.cout, spec.kh, spec.kw) if spec.transposed else (spec.cout, spec.cin, spec.kh, spec.kw)
        dw_out = torch.zeros(shape, device=x.device, dtype=torch.float32)
    L = _lib.lib()
    nbytes = L.erf_conv_wgrad_im2col_tc_workspace()
    ws = _WS.get((x.device, nbytes))
    if ws is None:
        ws = _WS[(x.device, nbytes)] = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    if (FOLD_HEAD_WGRAD and spec.transposed and spec.kh == 2 and spec.kw == 2 and spec.stride == 2 and spec.pad_h == 0 and spec.pad_w == 0 and dy_coff == 0
            and wo % 2 == 0 and (2 * dyld) % 16 == 0 and 2 * dyld <= 128 and spec.cin % 16 == 0 and spec.cin <= 128 and ho == 2 * h and wo == 2 * w):
        # Non-overlapping ConvTranspose2d(k = 2, s = 2) (the decoder head, eval/erfnet.py:124): dW[ci,co,a,b] = sum x[i,j,ci] * dy[2i+a, 2j+b, co].
        # dy [N,Ho,Wo,ld] read as [N,Ho,Wo/2,2*ld] makes the column pair (2j, 2j+1) ONE pixel with channels (b, co): two taps (a = 0, 1; row step 2,
        # column step 1) of 2*ld contiguous channels instead of four taps of ld -- whole 96-byte rows per gathered pixel (the kernel's fast path).
        d = _lib.WgradIm2colDesc()
        k2 = 2 * dyld
        d.g, d.GH, d.GW, d.g_ld, d.Cg, d.CgP, d.g_copy_bytes, d.g_step, d.g_step_x = dy.data_ptr(), ho, wo // 2, k2, k2, round_up(k2, 16), k2 * 2, 2, 1
        d.d, d.EN, d.EH, d.EW, d.d_ld, d.d_c0, d.d_c_base, d.Cd = x.data_ptr(), n, h, w, xld, 0, 0, spec.cin
        d.NP = round_up(spec.cin, 16)
        d.ntaps = 2
        d.tap_dy[0], d.tap_dx[0], d.tap_off[0] = 0, 0, 0
        d.tap_dy[1], d.tap_dx[1], d.tap_off[1] = 1, 0, 2
        d.KHW, d.fold, d.fold_c = 4, dyld, spec.cout
        with _timed(f"conv_wgrad_im2col_tc[{spec.cin}->{spec.cout} 2x2 s2T folded]" if DETAIL else "conv_wgrad_im2col_tc_kernel", conv_flops(spec, n, ho, wo)):
            check(L.erf_conv_wgrad_im2col_tc(C.byref(d), _p(dw_out), _p(ws), nbytes, _stream(x)))
        _count(2)
        return dw_out
    if spec.transposed:   # gathered = dy (channels co), dense = x (channels ci)
        g, gh, gw, gld, cg, goff = dy, ho, wo, dyld, spec.cout, dy_coff
        dn, eh, ew, dld, cd, dbase = x, h, w, xld, spec.cin, 0
    else:                 # gathered = x (channels ci), dense = dy (channels co)
        g, gh, gw, gld, cg, goff = x, h, w, xld, spec.cin, 0
        dn, eh, ew, dld, cd, dbase = dy, ho, wo, dyld, spec.cout, dy_coff
    if goff:
        raise RuntimeError("conv_wgrad_im2col_tc: the gathered operand must start at channel 0")
    cgp = round_up(cg, 16)
    avail = gld * 2                                   # bytes of the gathered pixel row
    copy = 8 if avail < 16 else min(avail // 16 * 16, cgp * 2)
    np_full = round_up(cd, 16)
    all_taps = [(ky, kx) for ky in range(spec.kh) for kx in range(spec.kw)]
    chg = 64 if cgp % 64 == 0 else 16

    def fits(ntaps, npad):   # two pipeline stages in shared memory and all accumulator groups in TMEM
        groups = -(-(ntaps * (cgp // chg)) // (128 // chg))
        stage = -(-(ntaps * (cgp // chg) * 64 * chg * 2) // 1024) * 1024 + npad * 128
        return groups * npad <= 512 and stage * 2 <= 190 * 1024

    # taps per launch (a 3x3 conv over 128 gathered channels does not fit at once: one kernel row per launch), then the N split
    tpl = len(all_taps)
    while tpl > 1 and not fits(tpl, 16):
        tpl = max(1, tpl // 2) if tpl % spec.kw else (tpl - spec.kw if tpl > spec.kw else max(1, tpl // 2))
    npad = np_full
    while npad > 16 and not fits(tpl, npad):
        npad = round_up(npad // 2, 16)
    if not fits(tpl, npad):
        raise RuntimeError(f"conv_wgrad_im2col_tc: {cg} gathered channels do not fit the kernel even one tap at a time")
    name = f"conv_wgrad_im2col_tc[{spec.cin}->{spec.cout} {spec.kh}x{spec.kw} s{spec.stride}{'T' if spec.transposed else ''}]" if DETAIL else "conv_wgrad_im2col_tc_kernel"
    with _timed(name, conv_flops(spec, n, ho, wo)):
        for t0 in range(0, len(all_taps), tpl):
            taps = all_taps[t0:t0 + tpl]
            for n0 in range(0, np_full, npad):
                d = _lib.WgradIm2colDesc()
                d.g, d.GH, d.GW, d.g_ld, d.Cg, d.CgP, d.g_copy_bytes, d.g_step = g.data_ptr(), gh, gw, gld, cg, cgp, copy, spec.stride
                d.d, d.EN, d.EH, d.EW, d.d_ld, d.d_c0, d.d_c_base, d.Cd = dn.data_ptr(), n, eh, ew, dld, dbase + n0, dbase, cd
                d.NP = min(npad, np_full - n0)
                d.ntaps = len(taps)
                for i, (ky, kx) in enumerate(taps):
                    d.tap_dy[i], d.tap_dx[i], d.tap_off[i] = ky - spec.pad_h, kx - spec.pad_w, ky * spec.kw + kx
                d.KHW = spec.kh * spec.kw
                check(L.erf_conv_wgrad_im2col_tc(C.byref(d), _p(dw_out), _p(ws), nbytes, _stream(x)))
                _count(2)
    return dw_out


WGRAD_FUSED_REDUCE = False  # True: the 3-tap weight gradients reduce their per-CTA partial sums behind a grid barrier inside the kernel (one launch
                            # instead of two; measured SLOWER at the bench workload: 19.2 -> 22.4 us at C=64 -- the barrier costs the tail overlap PDL gives two kernels)
FOLD_HEAD_WGRAD = True   # the 2x2 stride-2 transposed head's weight gradient through the column-folded view of dy (tests flip it)
_WS = {}  # (device, bytes) -> workspace tensor for the tensor-core wgrad partial sums


def conv1d_wgrad_tc_supported(c: int, w: int, dil: int, dtype) -> bool:
    return dtype == torch.bfloat16 and not EMULATED and c in (16, 64, 128) and w >= 16 and 1 <= dil <= 32


def conv1d_wgrad_tc(x: torch.Tensor, dy: torch.Tensor, horizontal: bool, dil: int, dw_out=None, db_out=None, want_bias=True):
    """Weight (torch layout [Co,Ci,3,1] / [Co,Ci,1,3], fp32) and bias gradients of a 3-tap conv on the tensor cores,
    accumulated into dw_out / db_out (e.g. the parameters' .grad views) or into fresh zero tensors."""
    n, h, w, c = x.shape
    L = _lib.lib()
    nbytes = L.erf_conv1d_wgrad_tc_workspace(c)
    ws = _WS.get((x.device, nbytes))
    if ws is None:
        ws = _WS[(x.device, nbytes)] = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    if dw_out is None:
        dw_out = torch.zeros((c, c, 1, 3) if horizontal else (c, c, 3, 1), device=x.device, dtype=torch.float32)
    with _timed(f"conv1d_wgrad_tc[C={c} {'H' if horizontal else 'V'}]" if DETAIL else "conv1d_wgrad_tc_kernel", 2.0 * n * h * w * c * c * 3):
        if WGRAD_FUSED_REDUCE:
            ctr = ARENA.take(1) if (ARENA is not None and ARENA.buf.device == x.device) else None      # zeroed grid-barrier word
            if ctr is None:
                ctr = torch.zeros(2, device=x.device, dtype=torch.int32)
            check(L.erf_conv1d_wgrad_tc_fused(_p(x), _p(dy), _p(dw_out), _p(ws), nbytes, _p(ctr), n, h, w, c, int(horizontal), dil, _stream(x)))
            _count()
        else:
            check(L.erf_conv1d_wgrad_tc(_p(x), _p(dy), _p(dw_out), _p(ws), nbytes, n, h, w, c, int(horizontal), dil, _stream(x)))
            _count(1 if (WGRAD_BULK and dw_out.data_ptr() % 16 == 0) else 2)   # bulk-reduction epilogue: no reduce launch (csrc/conv1d_wgrad_tc.cu)
    if want_bias:
        if db_out is None:
            db_out = torch.zeros(c, device=x.device, dtype=torch.float32)
        channel_sum(dy, db_out)
    return dw_out, db_out


WGRAD_BULK = os.environ.get("ERF_WGRAD_BULK", "1") != "0"     # mirrors the library's switch (launch counting only)


def set_wgrad_bulk(flag: bool) -> None:
    """Reduction form of conv1d_wgrad_tc: True (default) = every CTA adds its accumulators to the gradient with cp.reduce.async.bulk (one launch);
    False = per-CTA partial slabs + reduce kernel (two launches).  The library reads ERF_WGRAD_BULK on every call."""
    global WGRAD_BULK
    WGRAD_BULK = bool(flag)
    os.environ["ERF_WGRAD_BULK"] = "1" if flag else "0"


WGRAD_GROUP = os.environ.get("ERF_WGRAD_GROUP", "0") != "0"   # the four 3-tap weight gradients of a block in ONE launch (+ one reduce launch).  Measured: the
# eager step gains (89 -> 72 us per 128-channel block: six launch gaps less) but the graph-replayed step LOSES (5.33 -> 5.42 ms): programmatic
# dependent launch already hides those gaps, and deferring the weight gradients to the end of the block costs the L2 residency of x / dy


def conv1d_wgrad_tc_group(jobs):
    """jobs: [(x, dy, horizontal, dil, dw_out)] (1..4) over tensors of ONE shape [N,H,W,C]: dw_out_j += weight gradient of job j, all in one
    kernel launch plus one reduce launch (csrc/conv1d_wgrad_group_tc.cu)."""
    x0 = jobs[0][0]
    n, h, w, c = x0.shape
    L = _lib.lib()
    nbytes = L.erf_conv1d_wgrad_group_tc_workspace(c)
    ws = _WS.get((x0.device, nbytes))
    if ws is None:
        ws = _WS[(x0.device, nbytes)] = torch.empty(nbytes, device=x0.device, dtype=torch.uint8)
    arr = (_lib.WgradJob * len(jobs))()
    for i, (x, dy, horizontal, dil, dw) in enumerate(jobs):
        if x.shape != x0.shape or dy.shape != x0.shape or x.dtype != torch.bfloat16 or dy.dtype != torch.bfloat16:
            raise RuntimeError("conv1d_wgrad_tc_group: all jobs must share one bf16 [N,H,W,C] shape")
        arr[i].x, arr[i].dy, arr[i].dw, arr[i].horizontal, arr[i].dil = x.data_ptr(), dy.data_ptr(), dw.data_ptr(), int(horizontal), int(dil)
    with _timed(f"conv1d_wgrad_group[C={c} x{len(jobs)}]" if DETAIL else "conv1d_wgrad_tc_kernel", 2.0 * n * h * w * c * c * 3 * len(jobs)):
        check(L.erf_conv1d_wgrad_group_tc(arr, len(jobs), _p(ws), nbytes, n, h, w, c, _stream(x0)))
    _count(2)


def stem_supported(x: torch.Tensor) -> bool:
    return not EMULATED and x.shape[-1] == 4 and x.shape[1] % 2 == 0 and x.shape[2] % 4 == 0 and x.shape[2] <= 4096


def stem_fwd(x, w, bias, scale=None, shift=None, relu=False, stats=None):
    """DownsamplerBlock(3, 16) forward on x NHWC [N,H,W,4]: y [N,H/2,W/2,16] = relu?((conv | maxpool) * scale + shift)."""
    n, h, wd, _ = x.shape
    y = torch.empty(n, h // 2, wd // 2, 16, device=x.device, dtype=x.dtype)
    with _timed("stem_fwd", 2.0 * n * (h // 2) * (wd // 2) * 13 * 27):
        check(_lib.lib().erf_stem_fwd(_p(x), n, h, wd, _p(w), _p(bias), _p(scale), _p(shift), int(relu), _p(y), _p(stats), dt(x), _stream(x)))
    _count()
    return y


STEM_WGRAD_TC = os.environ.get("ERF_STEM_WGRAD_TC", "0") != "0"   # stem weight gradient on the generic tcgen05 gather kernel instead of the CUDA-core kernel
STEM_NCHW = True   # inference: the stem reads the caller's NCHW fp32 image directly (no layout pass)


def stem_nchw_supported(x: torch.Tensor) -> bool:
    return (STEM_NCHW and not EMULATED and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3 and x.shape[2] % 2 == 0
            and x.shape[3] % 4 == 0)


def stem_fwd_nchw(x, w, bias, scale, shift, relu, dtype):
    """DownsamplerBlock(3, 16) forward straight from the NCHW fp32 image x [N,3,H,W] -> NHWC y [N,H/2,W/2,16] of `dtype`."""
    n, _, h, wd = x.shape
    x = x.contiguous()
    y = torch.empty(n, h // 2, wd // 2, 16, device=x.device, dtype=dtype)
    with _timed("stem_fwd", 2.0 * n * (h // 2) * (wd // 2) * 13 * 27):
        check(_lib.lib().erf_stem_fwd_nchw(_p(x), n, h, wd, _p(w), _p(bias), _p(scale), _p(shift), int(relu), _p(y), None, dt(y), _stream(x)))
    _count()
    return y


def stem_wgrad(x, dy, dw_out=None, db_out=None, want_bias=True):
    """(dw [13,3,3,3], db [13] or None), ACCUMULATED into dw_out / db_out when given."""
    n, h, wd, _ = x.shape
    dw = dw_out if dw_out is not None else torch.zeros(13, 3, 3, 3, device=x.device, dtype=torch.float32)
    db = None
    if want_bias:
        db = db_out if db_out is not None else torch.zeros(13, device=x.device, dtype=torch.float32)
    with _timed("stem_wgrad", 2.0 * n * (h // 2) * (wd // 2) * 13 * 27):
        check(_lib.lib().erf_stem_wgrad(_p(x), _p(dy), n, h, wd, _p(dw), _p(db), dt(x), _stream(x)))
    _count()
    return dw, db


def maxpool_fwd(x, c, out, out_coff, scale=None, shift=None, relu=False, stats=None):
    n, h, w, ld = x.shape
    check(_lib.lib().erf_maxpool2x2_fwd(_p(x), n, h, w, c, ld, _p(out), out.shape[-1], out_coff, _p(scale), _p(shift), int(relu),
                                        _p(stats), 0 if stats is None else stats.numel() // 2, dt(x), _stream(x)))
    _count()


def maxpool_bwd(x, c, dout, out_coff, dx):
    n, h, w, ld = x.shape
    check(_lib.lib().erf_maxpool2x2_bwd(_p(x), n, h, w, c, ld, _p(dout), dout.shape[-1], out_coff, _p(dx), dx.shape[-1], dt(x), _stream(x)))
    _count()


def nchw_to_nhwc(x: torch.Tensor, ld: int, dtype) -> torch.Tensor:
    n, c, h, w = x.shape
    x = x.contiguous().float()
    out = torch.empty(n, h, w, ld, device=x.device, dtype=dtype)
    check(_lib.lib().erf_nchw_f32_to_nhwc(_p(x), n, c, h, w, _p(out), ld, dt(out), _stream(x)))
    _count()
    return out


def nhwc_to_nchw(x: torch.Tensor, c: int) -> torch.Tensor:
    n, h, w, ld = x.shape
    out = torch.empty(n, c, h, w, device=x.device, dtype=torch.float32)
    check(_lib.lib().erf_nhwc_to_nchw_f32(_p(x), ld, n, c, h, w, _p(out), dt(x), _stream(x)))
    _count()
    return out


def affine_act(x, scale=None, shift=None, res=None, relu=True, out=None):
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    per_sample = int(scale is not None and scale.dim() == 2)
    check(_lib.lib().erf_affine_act(_p(x), c, _p(scale), _p(shift), per_sample, _p(res), 0 if res is None else res.shape[-1], _p(out),
                                    out.shape[-1], n, h * w, c, int(relu), dt(x), _stream(x)))
    _count()
    return out


def channel_stats(x, stats):
    n, h, w, c = x.shape
    check(_lib.lib().erf_channel_stats(_p(x), c, c, n * h * w, _p(stats), dt(x), _stream(x)))
    _count()


def channel_sum(x, out):
    """out[c] += sum over pixels of x[..., c]"""
    n, h, w, c = x.shape
    check(_lib.lib().erf_channel_sum(_p(x), c, 0, c, n * h * w, _p(out), dt(x), _stream(x)))
    _count()


def channel_sum_slice(x, c, out):
    """out[k] += sum over pixels of x[..., k] for k < c (x may be wider than c)."""
    n, h, w, ld = x.shape
    check(_lib.lib().erf_channel_sum(_p(x), ld, 0, c, n * h * w, _p(out), dt(x), _stream(x)))
    _count()


def relu_bwd(g, mask):
    n, h, w, c = g.shape
    out = torch.empty_like(g)
    check(_lib.lib().erf_relu_bwd(_p(g), c, _p(mask), mask.shape[-1], _p(out), c, n * h * w, c, dt(g), _stream(g)))
    _count()
    return out


def bn_finalize(stats, count, gamma, beta, rm, rv, momentum, eps, drop=None):
    c = gamma.numel()
    n = 1 if drop is None else drop.shape[0]
    dev = gamma.device
    mean = torch.empty(c, device=dev, dtype=torch.float32)
    invstd = torch.empty(c, device=dev, dtype=torch.float32)
    shape = (c,) if drop is None else (n, c)
    scale = torch.empty(shape, device=dev, dtype=torch.float32)
    shift = torch.empty(shape, device=dev, dtype=torch.float32)
    check(_lib.lib().erf_bn_finalize(_p(stats), float(count), _p(gamma), _p(beta), _p(rm), _p(rv), float(momentum), float(eps), _p(drop), n, c,
                                     _p(mean), _p(invstd), _p(scale), _p(shift), _stream(gamma)))
    _count()
    return mean, invstd, scale, shift


def bn_eval_affine(gamma, beta, rm, rv, eps):
    c = gamma.numel()
    scale = torch.empty(c, device=gamma.device, dtype=torch.float32)
    shift = torch.empty_like(scale)
    check(_lib.lib().erf_bn_eval_affine(_p(gamma), _p(beta), _p(rm), _p(rv), float(eps), c, _p(scale), _p(shift), _stream(gamma)))
    _count()
    return scale, shift


BN_FUSED = True   # one launch per train-mode BatchNorm and direction (csrc/bn_fused.cu); False = the round-1 kernels (finalize + apply, reduce + apply)


def bn_fused_supported(c: int) -> bool:
    return BN_FUSED and not EMULATED and bool(_lib.lib().erf_bn_fused_supported(c))


def bn_stats_buf(c: int, device) -> torch.Tensor:
    """Zeroed fp64 [2C] accumulator for the batch sums (sum, sum of squares) of one train-mode BatchNorm: a slice of the step's scratch arena
    (one memset per step) when a trainer provides one, else a fresh tensor."""
    b = ARENA.take(2 * c) if (ARENA is not None and ARENA.buf.device == torch.device(device)) else None
    return b if b is not None else torch.zeros(2 * c, device=device, dtype=torch.float64)


def bn_apply(x, stats, count, gamma, beta, rm, rv, momentum, eps, drop=None, res=None, relu=True):
    """nn.BatchNorm2d.forward in training mode on the pre-BN activation x, from the fp64 batch sums `stats` its producer accumulated:
    out = relu?(((x - mean) * invstd * gamma + beta) * drop[n,c] (+ res)); running statistics updated (momentum, unbiased variance).
    Returns (out, mean, invstd).  One launch (every CTA derives scale / shift itself) when the layout allows, else finalize + apply."""
    n, h, w, c = x.shape
    if bn_fused_supported(c):
        out = torch.empty_like(x)
        mean = torch.empty(c, device=x.device, dtype=torch.float32)
        invstd = torch.empty(c, device=x.device, dtype=torch.float32)
        with _timed(f"bn_apply[C={c}{' +drop' if drop is not None else ''}{' +res' if res is not None else ''}]" if DETAIL else "bn_apply"):
            check(_lib.lib().erf_bn_fwd_apply(_p(x), c, _p(stats), float(count), _p(gamma), _p(beta), _p(rm), _p(rv), float(momentum), float(eps), _p(drop),
                                              _p(res), 0 if res is None else res.shape[-1], _p(out), c, n, h * w, c, int(relu), _p(mean), _p(invstd), dt(x),
                                              _stream(x)))
        _count()
        return out, mean, invstd
    mean, invstd, scale, shift = bn_finalize(stats, count, gamma, beta, rm, rv, momentum, eps, drop=drop)
    return affine_act(x, scale, shift, res=res, relu=relu), mean, invstd


_BN_SUMS = {}  # (device, C) -> fp64 [2C] scratch for the BN-backward reductions


class ScratchArena:
    """Pre-zeroed fp64 scratch for one training step: reset() is ONE memset, take(n) hands out consecutive zeroed slices.  Used by
    FlatAdamTrainer so that the 39 BN-backward reductions of a step do not each launch their own fill kernel (which is also a full
    serialisation point between programmatically dependent kernels)."""

    def __init__(self, device, n=131072):
        self.buf = torch.zeros(n, device=device, dtype=torch.float64)
        self.cursor = 0

    def reset(self):
        self.buf.zero_()
        self.cursor = 0

    def take(self, n):
        if self.cursor + n > self.buf.numel():
            return None
        out = self.buf[self.cursor:self.cursor + n]
        self.cursor += n
        return out


ARENA = None   # set by the trainer for the duration of a step


class SideQueue:
    """Weight-gradient launches off the critical path.  In the backward pass only the data gradients feed the next layer; the weight gradients
    are needed by the optimiser alone.  Every hot kernel here is sized one CTA per SM, so two of them never share an SM -- but at batch 6 each
    launch leaves SMs idle while its last CTAs finish (384 units over 148 SMs) and during the next launch's ramp-up.  Weight-gradient
    kernels enqueued on a second stream fill those gaps.  run(fn, keep): fn's launches are ordered after everything enqueued so far on the
    current stream; `keep` holds the tensors they read until join(), so the caching allocator cannot hand their memory to a later
    main-stream allocation while the side stream still reads it (also true inside a CUDA-graph capture, where the two streams become two
    branches of the graph).  Never put a kernel that contains a grid barrier on this stream: two barrier kernels on two streams can each
    hold part of the SMs and wait for the rest forever."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.keep = []
        self.used = False

    def run(self, fn, keep=()):
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            out = fn()
        self.keep.extend(keep)
        self.used = True
        return out

    def join(self):
        if self.used:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        self.keep.clear()
        self.used = False


SIDE = None        # a SideQueue while a trainer runs backward with WGRAD_SIDE_STREAM
# Weight gradients of backward on a second stream, joined before the optimiser.  "auto" (default): while a CUDA graph is being captured -- the
# replayed step gains ~1 % (4.712 -> 4.666 ms; configs 3 / 4: +7 %) because the weight-gradient kernels fill the ramp-up / tail gaps of the
# data-gradient chain; eagerly the extra events only cost CPU time (config 4 eager: 10.0 -> 11.8 ms), so it stays off there.  ERF_WGRAD_SIDE=1 / 0 force it.
WGRAD_SIDE_STREAM = {"1": True, "0": False}.get(os.environ.get("ERF_WGRAD_SIDE", "auto"), "auto")


SIDE_CAP = [int(v) for v in os.environ.get("ERF_SIDE_CAP", "0,0").split(",")]   # (main, side) SMs while the weight gradients run on the side stream


def set_launch_cap(main_sms: int, side_sms: int) -> None:
    if not EMULATED and hasattr(_lib.lib(), "erf_set_launch_cap"):
        check(_lib.lib().erf_set_launch_cap(int(main_sms), int(side_sms)))


MASK_FROM_Y = True   # BN -> ReLU without dropout / residual: the backward kernel recomputes the ReLU mask from y (one tensor read less)


BN_SUM_COPIES = int(os.environ.get("ERF_BN_SUM_COPIES", "8"))   # zeroed copies of the BN-backward sums: CTA b adds into copy b % K (spreads same-address atomics)


def bn_bwd(g, y, gamma, mean, invstd, mask=None, drop=None, dgamma_out=None, dbeta_out=None, want_masked=False, beta_for_mask=None):
    """Returns (dy, dgamma, dbeta) for y_bn = gamma*(y-mean)*invstd+beta, upstream grad g*(mask>0)*drop[n,c].
    dgamma/dbeta are WRITTEN into dgamma_out/dbeta_out when given (one BN contributes once per step).
    want_masked: also return g*(mask>0) as a 4th value (written by the apply pass, saves a separate relu_bwd).
    beta_for_mask: this BN's beta, given when mask == relu(BN(y)) exactly (no dropout, no residual): the bf16 kernel may then recompute
    the mask from y instead of reading the mask tensor."""
    n, h, w, c = y.shape
    K = BN_SUM_COPIES
    sums = ARENA.take(K * 2 * c + 1) if (ARENA is not None and ARENA.buf.device == y.device) else None     # [K][2C] sums + the grid-barrier counter
    if sums is None:
        sums = _BN_SUMS.get((y.device, c, K))
        if sums is None:
            sums = _BN_SUMS[(y.device, c, K)] = torch.zeros(K * 2 * c + 1, device=y.device, dtype=torch.float64)
        else:
            sums.zero_()
    L = _lib.lib()
    if bn_fused_supported(c):
        dy = torch.empty_like(y)
        dgamma = dgamma_out if dgamma_out is not None else torch.empty(c, device=y.device, dtype=torch.float32)
        dbeta = dbeta_out if dbeta_out is not None else torch.empty(c, device=y.device, dtype=torch.float32)
        gm = torch.empty_like(y) if want_masked else None
        bfm = beta_for_mask if (MASK_FROM_Y and mask is not None and drop is None) else None
        with _timed(f"bn_bwd[C={c}{' +drop' if drop is not None else ''}{' +gm' if want_masked else ''}{' y-mask' if bfm is not None else ''}]" if DETAIL else "bn_bwd"):
            check(L.erf_bn_bwd_fused_spread(_p(g), g.shape[-1], _p(y), c, _p(mask), 0 if mask is None else mask.shape[-1], _p(bfm), _p(drop), _p(gamma), _p(mean),
                                            _p(invstd), _p(sums), K, C.c_void_p(sums.data_ptr() + 16 * c * K), n, h * w, c, _p(dy), c, _p(dgamma), _p(dbeta),
                                            _p(gm), c, dt(y), _stream(y)))
        _count()
        return (dy, dgamma, dbeta, gm) if want_masked else (dy, dgamma, dbeta)
    dy = torch.empty_like(y)
    dgamma = dgamma_out if dgamma_out is not None else torch.empty(c, device=y.device, dtype=torch.float32)
    dbeta = dbeta_out if dbeta_out is not None else torch.empty(c, device=y.device, dtype=torch.float32)
    gm = torch.empty_like(y) if want_masked else None
    with _timed("bn_bwd"):
        check(L.erf_bn_bwd_reduce(_p(g), g.shape[-1], _p(y), c, _p(mask), 0 if mask is None else mask.shape[-1], _p(drop), _p(mean), _p(invstd),
                                  n, h * w, c, _p(sums), dt(y), _stream(y)))
        check(L.erf_bn_bwd_apply(_p(g), g.shape[-1], _p(y), c, _p(mask), 0 if mask is None else mask.shape[-1], _p(drop), _p(gamma), _p(mean),
                                 _p(invstd), _p(sums), n, h * w, c, _p(dy), c, _p(dgamma), _p(dbeta), _p(gm), c, dt(y), _stream(y)))
    _count(2)
    return (dy, dgamma, dbeta, gm) if want_masked else (dy, dgamma, dbeta)


def _same_device(what: str, ref: torch.Tensor, **tensors):
    """The kernels take raw pointers: a tensor on another device (or on the host) would be dereferenced on `ref`'s GPU."""
    if not EMULATED and not ref.is_cuda:
        raise RuntimeError(f"{what}: erfnet_pytorch_b200 kernels need CUDA tensors (there is no CPU fallback)")
    for name, t in tensors.items():
        if t is not None and t.device != ref.device:
            raise RuntimeError(f"{what}: {name} is on {t.device}, expected {ref.device}")


def check_ce_shapes(logits: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor], c: int, nchw: bool):
    """Host-side contract of the CE kernels (they loop over target.numel() pixels and read / WRITE c values per pixel of the logits):
    NCHW logits [N,c,H,W] or NHWC logits [N,H,W,ld >= c] against a target [N,H,W] on the same device.  The reference's
    nn.CrossEntropyLoss raises the same kind of size-mismatch error (train/evaluation_functions.py:11-14)."""
    _same_device("CrossEntropyLoss2d", logits, target=target, weight=weight)
    if logits.dim() != 4:
        raise RuntimeError(f"CrossEntropyLoss2d: expected 4-D logits, got shape {tuple(logits.shape)}")
    if nchw:
        n, ch, h, w = logits.shape
        if ch != c:
            raise RuntimeError(f"CrossEntropyLoss2d: logits have {ch} channels, expected {c}")
    else:
        n, h, w, ld = logits.shape
        if ld < c:
            raise RuntimeError(f"CrossEntropyLoss2d: NHWC logits pitch {ld} is smaller than the {c} classes")
    if tuple(target.shape) != (n, h, w):
        raise RuntimeError(f"CrossEntropyLoss2d: size mismatch, target {tuple(target.shape)} vs logits of {n} x {h} x {w} pixels "
                           f"(expected target shape {(n, h, w)})")
    if weight is not None and weight.numel() != c:
        raise RuntimeError(f"CrossEntropyLoss2d: weight has {weight.numel()} entries, expected {c}")
    if not logits.is_contiguous():
        raise RuntimeError("CrossEntropyLoss2d: logits must be contiguous")


def check_masked_loss_shapes(kind: int, a, b, prob, target):
    """Host-side contract of the masked-loss kernels (they loop over target.numel() pixels): per-pixel kinds need prediction / variance
    with exactly one value per target pixel; class-probability kinds need prob [B,C,*target.shape[1:]] and C per-class values."""
    _same_device("masked loss", target, prediction=a, variance=b, prob=prob)
    if prob is None:
        for name, t in (("prediction", a), ("variance", b)):
            if t is not None and t.numel() != target.numel():
                raise RuntimeError(f"masked loss: size mismatch, {name} {tuple(t.shape)} vs target {tuple(target.shape)}")
    else:
        want = (target.shape[0], prob.shape[1]) + tuple(target.shape[1:]) if prob.dim() >= 2 else None
        if prob.dim() != 4 or tuple(prob.shape) != want:
            raise RuntimeError(f"masked loss: size mismatch, class probabilities {tuple(prob.shape)} vs target {tuple(target.shape)} "
                               f"(expected {want})")
        for name, t in (("per-class cost", a), ("per-class variance", b)):
            if t is not None and t.numel() != prob.shape[1]:
                raise RuntimeError(f"masked loss: {name} has {t.numel()} values, expected {prob.shape[1]}")


def ce_weight_sum(target: torch.Tensor, weight: Optional[torch.Tensor], c: int, ignore_index: int, wsum: torch.Tensor):
    """wsum[0] += sum over valid pixels of weight[target] (1 when weight is None)."""
    if target.dtype != torch.int64:
        raise RuntimeError("CrossEntropyLoss2d: target must be int64")
    _same_device("CrossEntropyLoss2d", target, weight=weight, wsum=wsum)
    if weight is not None and weight.numel() != c:
        raise RuntimeError(f"CrossEntropyLoss2d: weight has {weight.numel()} entries, expected {c}")
    check(_lib.lib().erf_ce_weight_sum(_p(target), target.numel(), _p(weight), c, ignore_index, _p(wsum), _stream(target)))
    _count()


def ce_colsum_supported(logits: torch.Tensor) -> bool:
    """ce_fwd_bwd(..., colsum=) is available: NHWC bf16 logits with a pitch of 24 or 32 on the real kernels."""
    return (not EMULATED and logits.is_cuda and logits.dtype == torch.bfloat16
            and bool(_lib.lib().erf_ce_fwd_bwd_colsum_supported(logits.shape[-1], 0, ERF_BF16)))


def ce_fwd_bwd(logits: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor], c: int, ignore_index: int,
               wsum: torch.Tensor, lsum: torch.Tensor, nchw: bool = False, want_grad: bool = True, colsum: Optional[torch.Tensor] = None):
    """lsum[0] += sum_valid w[y]*nll; returns dlogits = w[y]*(softmax-onehot)/wsum (same layout/dtype as logits) or None.
    colsum (fp32 [>= c], only where ce_colsum_supported): += per-class sums of the stored dlogits (the head conv's bias gradient)."""
    check_ce_shapes(logits, target, weight, c, nchw)
    _same_device("CrossEntropyLoss2d", logits, wsum=wsum, lsum=lsum, colsum=colsum)
    if colsum is not None:
        if nchw or not want_grad or not ce_colsum_supported(logits):
            raise RuntimeError("ce_fwd_bwd: colsum needs NHWC bf16 logits with a pitch of 24 or 32 and want_grad")
        dlog = torch.empty_like(logits)
        check(_lib.lib().erf_ce_fwd_bwd_colsum(_p(logits), logits.shape[-1], _p(target), target.numel(), _p(weight), c, ignore_index, _p(wsum), _p(lsum),
                                               _p(dlog), _p(colsum), _stream(logits)))
        _count()
        return dlog
    dlog = torch.empty_like(logits) if want_grad else None
    if nchw:
        hw, ld = logits.shape[2] * logits.shape[3], 0
    else:
        hw, ld = 0, logits.shape[-1]
        if want_grad and ld > c and not (logits.dtype == torch.bfloat16 and ld in (24, 32)):
            dlog.zero_()   # (the vectorised bf16 kernel writes the padding channels itself)
    check(_lib.lib().erf_ce_fwd_bwd(_p(logits), ld, hw, _p(target), target.numel(), _p(weight), c, ignore_index, _p(wsum), _p(lsum), _p(dlog),
                                    dt(logits), _stream(logits)))
    _count()
    return dlog


# ---- global-batch loss normalisation (data-parallel training through a criterion, parallel.FlatAdamTrainer(loss_fn=...)) --------------------
# The reference computes its criterion on the outputs nn.DataParallel GATHERED from all replicas (train/main.py:869-874), i.e. the mean runs over the
# global batch.  While LOSS_REDUCE is set (a callable that sum-all-reduces a tensor in place), every criterion of this package reduces its
# (sum, count) accumulators across ranks before dividing, so its value is the global mean and its gradient kernels divide by the global count:
# SUM-all-reduced parameter gradients are then exactly the reference's.  LOSS_REDUCED records that a criterion did so.
LOSS_REDUCE = None
LOSS_REDUCED = False


def loss_global_reduce(t: torch.Tensor):
    global LOSS_REDUCED
    if LOSS_REDUCE is not None:
        LOSS_REDUCE(t)
        LOSS_REDUCED = True


def cross_entropy(logits: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor], ignore_index: int, c: int,
                  nchw: bool, want_grad: bool):
    """Returns (loss_sum fp32[1], wsum fp32[1], dlogits or None); loss = loss_sum / wsum."""
    check_ce_shapes(logits, target, weight, c, nchw)     # before any kernel touches a pointer (also under the CPU emulation of the tests)
    target = target.contiguous()
    acc = torch.zeros(2, device=logits.device, dtype=torch.float32)
    wsum, lsum = acc[0:1], acc[1:2]
    ce_weight_sum(target, weight, c, ignore_index, wsum)
    loss_global_reduce(wsum)                       # the gradient is normalised by the global valid-pixel weight
    dlog = ce_fwd_bwd(logits, target, weight, c, ignore_index, wsum, lsum, nchw, want_grad)
    loss_global_reduce(lsum)
    return lsum, wsum, dlog


def confusion_hist(pred: torch.Tensor, gt: torch.Tensor, nc: int, cm: torch.Tensor):
    if pred.dtype != gt.dtype or pred.dtype not in (torch.uint8, torch.int64):
        raise RuntimeError("iouEval: predictions and labels must both be uint8 or both int64")
    if pred.numel() != gt.numel():
        raise RuntimeError(f"iouEval: size mismatch {tuple(pred.shape)} vs {tuple(gt.shape)}")
    if pred.numel() == 0:
        return
    pred, gt = pred.contiguous(), gt.contiguous()
    check(_lib.lib().erf_confusion_hist(_p(pred), _p(gt), pred.element_size(), pred.numel(), nc, _p(cm), _stream(pred)))
    _count()


def iou_counts(cm: torch.Tensor, nc: int, ignore: int, kept: int):
    out = torch.empty(3, kept, device=cm.device, dtype=torch.float64)
    check(_lib.lib().erf_iou_counts(_p(cm), nc, ignore, kept, _p(out[0]), _p(out[1]), _p(out[2]), _stream(cm)))
    _count()
    return out


def argmax(logits: torch.Tensor, c: int, nchw: bool, out_dtype=torch.int64) -> torch.Tensor:
    if nchw:
        n, _, h, w = logits.shape
        hw, ld = h * w, 0
    else:
        n, h, w, ld = logits.shape
        hw = 0
    out = torch.empty(n, h, w, device=logits.device, dtype=out_dtype)
    check(_lib.lib().erf_argmax(_p(logits), ld, hw, n * h * w, c, _p(out), out.element_size(), dt(logits), _stream(logits)))
    _count()
    return out


def softmax_nhwc_fwd(logits, c, drop=None):
    """NHWC logits [N,H,W,ld] -> NCHW fp32 probabilities [N,c,H,W] = softmax_c(logits * drop[n,c])."""
    n, h, w, ld = logits.shape
    prob = torch.empty(n, c, h, w, device=logits.device, dtype=torch.float32)
    check(_lib.lib().erf_softmax_nhwc_fwd(_p(logits), ld, _p(drop), n, h * w, c, _p(prob), dt(logits), _stream(logits)))
    _count()
    return prob


def softmax_nhwc_bwd(prob, dprob, drop, ld, dtype):
    n, c, h, w = prob.shape
    dlog = torch.empty(n, h, w, ld, device=prob.device, dtype=dtype)
    check(_lib.lib().erf_softmax_nhwc_bwd(_p(prob), _p(dprob.contiguous().float()), _p(drop), n, h * w, c, _p(dlog), ld, dt(dlog), _stream(prob)))
    _count()
    return dlog


LOSS_L1, LOSS_LOGLIK, LOSS_L1_CLASSPROB, LOSS_LOGLIK_CLASSPROB, LOSS_MSE_WEIGHTED = range(5)


def _f32c(t, name):
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise RuntimeError(f"masked loss: {name} must be float32, got {t.dtype}")
    return t.contiguous()


def masked_loss_fwd(kind: int, a, b, prob, target, eps: float, acc: torch.Tensor):
    """acc (fp64 [2], caller-zeroed) += (sum of per-pixel losses, #labelled pixels); see include/erfnet_b200.h."""
    n_px = target.numel()
    if prob is not None:
        hw, c = prob.shape[2] * prob.shape[3], prob.shape[1]
    else:
        hw, c = n_px, 0
    check(_lib.lib().erf_masked_loss_fwd(kind, _p(a), _p(b), _p(prob), _p(target), n_px, hw, c, eps, _p(acc), _stream(target)))
    _count()


def masked_loss_bwd(kind: int, a, b, prob, target, eps: float, scale: float, acc, go):
    """Gradients of scale * acc[0] / acc[1] times go: (d_a, d_b) for the per-pixel kinds, (d_prob, dclass fp64 [2C] UNSCALED)
    for the class-probability kinds."""
    n_px = target.numel()
    if prob is not None:
        hw, c = prob.shape[2] * prob.shape[3], prob.shape[1]
        d_prob = torch.empty_like(prob)
        dclass = torch.zeros(2 * c, device=prob.device, dtype=torch.float64)
        d_a = d_b = None
    else:
        hw, c = n_px, 0
        d_prob = dclass = None
        d_a = torch.empty_like(a)
        d_b = torch.empty_like(b) if kind == LOSS_LOGLIK else None
    check(_lib.lib().erf_masked_loss_bwd(kind, _p(a), _p(b), _p(prob), _p(target), n_px, hw, c, eps, scale, _p(acc), _p(go), _p(d_a), _p(d_b),
                                         _p(d_prob), _p(dclass), _stream(target)))
    _count()
    return (d_prob, dclass) if prob is not None else (d_a, d_b)


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    check(_lib.lib().erf_adam_step(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, weight_decay, step, grad_scale, _stream(p)))
    _count()


def adam_step_dev(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale=1.0):
    """Adam with the step counter in device memory (int32[1], incremented by the call) -- CUDA-graph capturable."""
    check(_lib.lib().erf_adam_step_dev(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, weight_decay, _p(step_dev), grad_scale,
                                       _stream(p)))
    _count(2)


def adam_step_graph(p, g, m, v, hyper_dev, beta1, beta2, eps, step_dev, grad_scale=1.0):
    """Adam with step counter, learning rate and weight decay in device memory (hyper_dev fp32 [2] = {lr, weight_decay}): the captured
    launch follows later changes of the learning rate."""
    check(_lib.lib().erf_adam_step_graph(_p(p), _p(g), _p(m), _p(v), p.numel(), _p(hyper_dev), beta1, beta2, eps, _p(step_dev), grad_scale, _stream(p)))
    _count(2)


def sgd_step_graph(p, g, buf, hyper_dev, momentum, grad_scale=1.0):
    check(_lib.lib().erf_sgd_step_graph(_p(p), _p(g), _p(buf), p.numel(), _p(hyper_dev), momentum, grad_scale, _stream(p)))
    _count()


# ---- ImageNet tail, mean teacher, I/O side (csrc/tail_io.cu) -------------------------------------------------------------------
def avgpool_fwd(x: torch.Tensor, c: int, k: int) -> torch.Tensor:
    """nn.AvgPool2d(k, 1, 0) over NHWC x [N,H,W,ld] -> fp32 NHWC [N,H-k+1,W-k+1,c]."""
    n, h, w, ld = x.shape
    if h < k or w < k:
        raise RuntimeError(f"AvgPool2d({k}): input {h}x{w} is smaller than the window")
    out = torch.empty(n, h - k + 1, w - k + 1, c, device=x.device, dtype=torch.float32)
    check(_lib.lib().erf_avgpool_fwd(_p(x), n, h, w, ld, c, k, _p(out), dt(x), _stream(x)))
    _count()
    return out


def avgpool_bwd(dout: torch.Tensor, h: int, w: int, ld: int, k: int, dtype) -> torch.Tensor:
    n, _, _, c = dout.shape
    dx = torch.empty(n, h, w, ld, device=dout.device, dtype=dtype)
    check(_lib.lib().erf_avgpool_bwd(_p(dout.contiguous()), n, h, w, ld, c, k, _p(dx), dt(dx), _stream(dout)))
    _count()
    return dx


def linear_fwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    m, k = x.shape
    n = w.shape[0]
    out = torch.empty(m, n, device=x.device, dtype=torch.float32)
    check(_lib.lib().erf_linear_fwd(_p(x), _p(w), _p(bias), _p(out), m, n, k, _stream(x)))
    _count()
    return out


def linear_bwd(x, w, dout, need_dx=True, dw_out=None, db_out=None, want_bias=True):
    """-> (dx or None, dw, db): dw / db are ACCUMULATED into dw_out / db_out (e.g. .grad views) or into fresh zero tensors."""
    m, k = x.shape
    n = w.shape[0]
    dx = torch.empty(m, k, device=x.device, dtype=torch.float32) if need_dx else None
    dw = torch.zeros(n, k, device=x.device, dtype=torch.float32) if dw_out is None else dw_out
    db = (torch.zeros(n, device=x.device, dtype=torch.float32) if db_out is None else db_out) if want_bias else None
    check(_lib.lib().erf_linear_bwd(_p(x), _p(w), _p(dout.contiguous()), _p(dx), _p(dw), _p(db), m, n, k, _stream(x)))
    _count(2 + int(need_dx))
    return dx, dw, db


def ce_rows_fwd(logits: torch.Tensor, target: torch.Tensor, ignore_index: int):
    """-> (acc fp32 [2] = (sum nll, #valid rows), lse fp32 [N], rank int32 [N])."""
    n, c = logits.shape
    acc = torch.zeros(2, device=logits.device, dtype=torch.float32)
    lse = torch.empty(n, device=logits.device, dtype=torch.float32)
    rank = torch.empty(n, device=logits.device, dtype=torch.int32)
    check(_lib.lib().erf_ce_rows_fwd(_p(logits), _p(target), n, c, ignore_index, _p(lse), _p(rank), _p(acc), _stream(logits)))
    _count()
    return acc, lse, rank


def ce_rows_bwd(logits, target, ignore_index, lse, acc, go):
    n, c = logits.shape
    dlog = torch.empty_like(logits)
    check(_lib.lib().erf_ce_rows_bwd(_p(logits), _p(target), n, c, ignore_index, _p(lse), _p(acc), _p(go), _p(dlog), _stream(logits)))
    _count()
    return dlog


def sgd_step(p, g, buf, lr, momentum, weight_decay, first_step, grad_scale=1.0):
    check(_lib.lib().erf_sgd_step(_p(p), _p(g), _p(buf), p.numel(), lr, momentum, weight_decay, int(first_step), grad_scale, _stream(p)))
    _count()


def ema_update(teacher, student, rate):
    check(_lib.lib().erf_ema_update(_p(teacher), _p(student), teacher.numel(), rate, _stream(teacher)))
    _count()


def ingest_u8(img: torch.Tensor, ld: int, dtype, mean=None, std=None) -> torch.Tensor:
    """uint8 [N,H,W,3] (HWC, as PIL / a camera delivers it) -> NHWC activations [N,H,W,ld] = ToTensor [+ Normalize] + layout, one pass."""
    if img.dtype != torch.uint8 or img.dim() != 4 or img.shape[-1] != 3:
        raise RuntimeError(f"ingest_u8: expected a uint8 [N,H,W,3] tensor, got {img.dtype} {tuple(img.shape)}")
    img = img.contiguous()
    n, h, w, _ = img.shape
    out = torch.empty(n, h, w, ld, device=img.device, dtype=dtype)
    check(_lib.lib().erf_ingest_u8(_p(img), n * h * w, ld, _p(mean), _p(std), _p(out), dt(out), _stream(img)))
    _count()
    return out


def label_lut(labels: torch.Tensor, lut: torch.Tensor) -> torch.Tensor:
    if labels.dtype not in (torch.uint8, torch.int64):
        raise RuntimeError("label_lut: labels must be uint8 or int64")
    labels = labels.contiguous()
    out = torch.empty(labels.shape, device=labels.device, dtype=torch.int64)
    if labels.numel():
        check(_lib.lib().erf_label_lut(_p(labels), labels.element_size(), labels.numel(), _p(lut), lut.numel(), _p(out), _stream(labels)))
        _count()
    return out


def colorize(labels: torch.Tensor, cmap: torch.Tensor) -> torch.Tensor:
    """labels uint8 / int64 [N,H,W] -> uint8 [N,3,H,W]; cmap uint8 [n_colors,3]."""
    if labels.dtype not in (torch.uint8, torch.int64):
        raise RuntimeError("colorize: labels must be uint8 or int64")
    labels = labels.contiguous()
    n, h, w = labels.shape
    out = torch.empty(n, 3, h, w, device=labels.device, dtype=torch.uint8)
    check(_lib.lib().erf_colorize(_p(labels), labels.element_size(), n, h * w, _p(cmap), cmap.shape[0], _p(out), _stream(labels)))
    _count()
    return out


def umma_probe(a: torch.Tensor, b: torch.Tensor, row_shift: int, layout: int, iters: int = 1):
    """a [rows,K] bf16, b [N,K] bf16 -> (D [128,N] fp32, cycles)."""
    rows, k = a.shape
    n = b.shape[0]
    d = torch.empty(128, n, device=a.device, dtype=torch.float32)
    cyc = torch.zeros(1, device=a.device, dtype=torch.int64)
    check(_lib.lib().erf_umma_probe(_p(a), rows, _p(b), n, k // 16, row_shift, layout, _p(d), iters, _p(cyc), _stream(a)))
    _count()
    return d, cyc


def _wrap_timed(name):
    fn = globals()[name]

    def wrapper(*a, **k):
        with _timed(name):
            return fn(*a, **k)
    wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
    globals()[name] = wrapper


for _n in ("channel_stats", "channel_sum", "channel_sum_slice", "maxpool_fwd", "maxpool_bwd", "nchw_to_nhwc", "nhwc_to_nchw", "affine_act", "relu_bwd", "bn_finalize", "bn_eval_affine",
           "ce_weight_sum", "ce_fwd_bwd", "confusion_hist", "iou_counts", "argmax", "adam_step", "pack_weight_fwd", "pack_weight_dgrad",
           "unpack_wgrad", "pack_conv1d_tc", "pack_conv1d_tc_pair", "avgpool_fwd", "avgpool_bwd", "linear_fwd", "linear_bwd", "ce_rows_fwd", "ce_rows_bwd",
           "sgd_step", "ema_update", "ingest_u8", "label_lut", "colorize"):
    _wrap_timed(_n)
