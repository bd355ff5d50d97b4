"""Tensor-level wrappers over the C ABI (include/damvs.h).

Everything here takes CUDA torch tensors, validates shape / dtype / contiguity,
allocates outputs with torch (so the caching allocator owns all memory) and
calls the library on torch's current stream.  PyTorch is plumbing only: device
memory and streams.  There is no fallback path.
"""
from __future__ import annotations

import contextlib
import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import AGG_ADAPTIVE, AGG_VARIANCE, BF16, CONV_DIRECT, CONV_TCGEN05, F16, F32, ConvDesc

# --------------------------------------------------------------------------
# precision policy
# --------------------------------------------------------------------------
# The default is the reference's own arithmetic width: fp32 (the <= 1e-4 parity mode).  The reduced-precision pipelines
# are opt-in -- bench.py, HotPathTrainer users and dropin.install(precision=...) ask for them explicitly:
#   'fp16'  fp16 features, fp16 cost volume / activations / weights, tcgen05 convolutions, fp32 accumulation.  Inference.
#           11 significand bits: 8x finer than bf16 at the same width and tensor-core rate.
#   'bf16'  the same with bf16 volumes / activations / weights: the exponent range training gradients need.
_POLICY = {"precision": "fp32", "conv_impl": "auto", "features": "auto", "fuse_prob_head": False}


def get_precision() -> str:
    return _POLICY["precision"]


def set_precision(precision: str, conv_impl: str = "auto", features: str = "auto") -> None:
    """precision: 'fp32' (default: everything fp32, direct convolutions -- the mode the <=1e-4 depth parity claim is
    made in), 'fp16' or 'bf16' (cost volume and CostRegNet activations / weights in that 2-byte type, fp32 accumulate;
    tensor-core convolutions; fp16 is inference-only).
    conv_impl: 'auto' | 'direct' | 'tcgen05'.  features: 'auto' (fp16 NHWC features in the bf16 pipeline) | 'fp32'
    (keep the gather in fp32 even when the cost volume is emitted in bf16: the ablation row of DESIGN.md section 5)."""
    assert precision in ("bf16", "fp16", "fp32") and conv_impl in ("auto", "direct", "tcgen05") and features in ("auto", "fp32")
    _POLICY["precision"] = precision
    _POLICY["conv_impl"] = conv_impl
    _POLICY["features"] = features


@contextlib.contextmanager
def precision(precision: str, conv_impl: str = "auto", features: str = "auto"):
    old = dict(_POLICY)
    set_precision(precision, conv_impl, features)
    try:
        yield
    finally:
        _POLICY.update(old)


def set_fusion(prob_head: bool) -> None:
    """Opt into the fused `prob` convolution + head launch (damvs_prob_head_fwd).  Off by default: measured on B200 it
    removes 0.28 GB of logits traffic and three launches per view but is no faster -- the head pass sits on the epilogue
    warps' critical path -- so the two-kernel path stays the default."""
    _POLICY["fuse_prob_head"] = bool(prob_head)


def fuse_prob_head() -> bool:
    return _POLICY["fuse_prob_head"]


def half_features() -> bool:
    """Whether the bf16 pipeline gathers fp16 features (its default) or stays on fp32 features (ablation)."""
    return _POLICY["features"] == "auto"


_VOLUME_DTYPES = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}
HALF_DTYPES = (torch.bfloat16, torch.float16)


def volume_dtype() -> torch.dtype:
    return _VOLUME_DTYPES[_POLICY["precision"]]


def conv_impl_for(cin: int, cout: int, stride: int, transposed: bool) -> int:
    choice = _POLICY["conv_impl"]
    if choice == "direct" or _POLICY["precision"] == "fp32":
        return CONV_DIRECT
    if choice == "tcgen05":
        return CONV_TCGEN05
    return CONV_TCGEN05 if tc_supported(cin, cout, stride, transposed) else CONV_DIRECT


def tc_supported(cin: int, cout: int, stride: int, transposed: bool) -> bool:
    """Layer shapes conv3d_tc.cu has kernels for (bf16 in/out): the CostRegNet layers with base_channels 8
    (reference models/module.py:513-530) -- see the dispatch table in conv3d_tc_launch.  Anything else runs
    on the direct kernel."""
    def cp(c):
        return 16 if c <= 16 else (32 if c <= 32 else 64)
    g = cin // 8
    if cin % 8 or cout > 64:
        return False
    if transposed:
        return (g, cp(cout)) in ((8, 32), (4, 16), (2, 16))
    if stride == 2:
        return (g, cp(cout)) in ((1, 16), (2, 32), (4, 64))
    if cin == 64 and cout == 64:
        return True                      # two launches of 32 output channels
    return (g, cp(cout)) in ((1, 16), (2, 16), (4, 16), (4, 32), (8, 32), (1, 32))


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _need(t: torch.Tensor, name: str, dtype=torch.float32, ndim: Optional[int] = None) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (this package has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dims, got shape {tuple(t.shape)}")


def _dt(dtype: torch.dtype) -> int:
    return BF16 if dtype == torch.bfloat16 else (F16 if dtype == torch.float16 else F32)


@dataclass
class G8Volume:
    """A logical [B,C,D,H,W] volume stored as [B, C/8, D, H, W, 8] (include/damvs.h)."""
    data: torch.Tensor

    @property
    def shape(self):
        b, g, d, h, w, _ = self.data.shape
        return (b, g * 8, d, h, w)

    @property
    def dtype(self):
        return self.data.dtype

    @staticmethod
    def empty(b, c, d, h, w, dtype, device) -> "G8Volume":
        assert c % 8 == 0
        return G8Volume(torch.empty((b, c // 8, d, h, w, 8), dtype=dtype, device=device))

    @staticmethod
    def from_ncdhw(x: torch.Tensor, dtype: torch.dtype) -> "G8Volume":
        _need(x, "x", torch.float32, 5)
        x = x.contiguous()
        b, c, d, h, w = x.shape
        if c % 8:
            raise ValueError(f"channel count {c} must be a multiple of 8")
        out = G8Volume.empty(b, c, d, h, w, dtype, x.device)
        with torch.cuda.device_of(x):
            _lib.check(_lib.load().damvs_ncdhw_to_g8(_p(x), _p(out.data), _dt(dtype), b, c, d, h, w, _stream()))
        return out

    def to_ncdhw(self) -> torch.Tensor:
        b, c, d, h, w = self.shape
        out = torch.empty((b, c, d, h, w), dtype=torch.float32, device=self.data.device)
        with torch.cuda.device_of(self.data):
            _lib.check(_lib.load().damvs_g8_to_ncdhw(_p(self.data), _dt(self.dtype), _p(out), b, c, d, h, w, _stream()))
        return out


# --------------------------------------------------------------------------
# features / projections
# --------------------------------------------------------------------------
def features_to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] fp32 -> contiguous [B,H,W,C].  Zero-copy when x is already channels_last."""
    if isinstance(x, torch.Tensor) and x.dtype == torch.float16:
        x = x.float()          # fp16 features handed to the fp32 pipeline: widened, never silently kept narrow
    _need(x, "feature", torch.float32, 4)
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return v
    x = x.contiguous()
    b, c, h, w = x.shape
    out = torch.empty((b, h, w, c), dtype=torch.float32, device=x.device)
    with torch.cuda.device_of(x), _timed("repack", bytes=float(x.numel() * 8)):
        _lib.check(_lib.load().damvs_nchw_to_nhwc_f32(_p(x), _p(out), b, c, h, w, _stream()))
    return out


def features_to_nhwc_half(x: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] fp32 (any memory format) -> contiguous [B,H,W,C] fp16, saturating (the bf16 pipeline's feature
    format, damvs_nchw_to_nhwc_f16)."""
    if isinstance(x, torch.Tensor) and x.dtype == torch.float16 and x.is_cuda and x.dim() == 4:
        # already the kernel's width: zero-copy when channels_last (what HotPathRunner uploads), one permute otherwise
        v = x.permute(0, 2, 3, 1)
        return v if v.is_contiguous() else v.contiguous()
    _need(x, "feature", torch.float32, 4)
    b, c, h, w = x.shape
    if c % 8:
        raise ValueError(f"feature channels {c} must be a multiple of 8")
    if x.permute(0, 2, 3, 1).is_contiguous() and not x.is_contiguous():
        # already channels_last: a cast is all that is left (rare path; the reference hands over NCHW)
        return x.permute(0, 2, 3, 1).clamp(-65504.0, 65504.0).to(torch.float16)
    x = x.contiguous()
    out = torch.empty((b, h, w, c), dtype=torch.float16, device=x.device)
    with torch.cuda.device_of(x), _timed("repack", bytes=float(x.numel() * 4 + out.numel() * 2)):
        _lib.check(_lib.load().damvs_nchw_to_nhwc_f16(_p(x), _p(out), b, c, h, w, _stream()))
    return out


def features_to_nhwc_half_multi(xs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """features_to_nhwc_half for the N views of a stage in ONE launch (equal shapes, plain NCHW); anything else goes
    through the per-tensor path."""
    x0 = xs[0]
    if len(xs) > 16 or any((not isinstance(x, torch.Tensor)) or x.shape != x0.shape or x.dtype != torch.float32 or not x.is_cuda
                           or not x.is_contiguous() for x in xs) or x0.dim() != 4 or x0.shape[1] % 8:
        return [features_to_nhwc_half(x) for x in xs]
    b, c, h, w = x0.shape
    outs = [torch.empty((b, h, w, c), dtype=torch.float16, device=x0.device) for _ in xs]
    ins_p = (ctypes.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
    outs_p = (ctypes.c_void_p * len(xs))(*[o.data_ptr() for o in outs])
    with torch.cuda.device_of(x0), _timed("repack", bytes=float(len(xs) * x0.numel() * 6)):
        _lib.check(_lib.load().damvs_nchw_to_nhwc_f16_multi(ins_p, outs_p, len(xs), b, c, h, w, _stream()))
    return outs


def compose_projection(proj_pair: torch.Tensor) -> torch.Tensor:
    """[...,2,4,4] -> [...,4,4]: rows 0-2 = K @ E[:3,:4] (reference models/cas_mvsnet.py:44-47)."""
    out = proj_pair[..., 0, :, :].clone()
    out[..., :3, :4] = torch.matmul(proj_pair[..., 1, :3, :3], proj_pair[..., 0, :3, :4])
    return out


def relative_rot_trans(src_proj: torch.Tensor, ref_proj: torch.Tensor) -> torch.Tensor:
    """[...,4,4] x [...,4,4] -> [...,12]: rows 0-2 of src_proj @ inverse(ref_proj), rot then trans
    (reference models/module.py:308-310).  inv_ex avoids torch.inverse's host sync."""
    inv = torch.linalg.inv_ex(ref_proj, check_errors=False).inverse
    proj = torch.matmul(src_proj, inv)
    return torch.cat([proj[..., :3, :3].reshape(*proj.shape[:-2], 9), proj[..., :3, 3]], dim=-1).contiguous()


# --------------------------------------------------------------------------
# kernels
# --------------------------------------------------------------------------
def _hyp_flags(depth_values: torch.Tensor, b: int, h: int, w: int):
    _need(depth_values, "depth_values", torch.float32)
    if depth_values.dim() == 2:
        if depth_values.shape[0] != b:
            raise ValueError("depth_values batch mismatch")
        return depth_values.contiguous(), 0, depth_values.shape[1]
    if depth_values.dim() == 4:
        if depth_values.shape[0] != b or depth_values.shape[2] != h or depth_values.shape[3] != w:
            raise ValueError(f"depth_values shape {tuple(depth_values.shape)} does not match [B={b},D,H={h},W={w}]")
        return depth_values.contiguous(), 1, depth_values.shape[1]
    raise ValueError("depth_values must be [B,D] or [B,D,H,W]")


def homo_warp(src_nhwc: torch.Tensor, rot_trans: torch.Tensor, depth_values: torch.Tensor) -> torch.Tensor:
    """Stand-alone warp: [B,H,W,C] -> [B,C,D,H,W] fp32 (reference models/module.py:297-332)."""
    _need(src_nhwc, "src_nhwc", torch.float32, 4)
    b, h, w, c = src_nhwc.shape
    _need(rot_trans, "rot_trans", torch.float32, 2)
    if tuple(rot_trans.shape) != (b, 12):
        raise ValueError("rot_trans must be [B,12]")
    dv, per_pixel, d = _hyp_flags(depth_values, b, h, w)
    out = torch.empty((b, c, d, h, w), dtype=torch.float32, device=src_nhwc.device)
    with torch.cuda.device_of(src_nhwc):
        _lib.check(_lib.load().damvs_homo_warp_fwd(_p(src_nhwc.contiguous()), _p(rot_trans.contiguous()), _p(dv), _p(out),
                                                   b, c, d, h, w, per_pixel, _stream()))
    return out


def warp_aggregate(ref_nhwc: torch.Tensor, src_nhwc: Sequence[torch.Tensor], rot_trans: torch.Tensor,
                   depth_values: torch.Tensor, wnet: Optional[torch.Tensor], mode: str,
                   out_dtype: torch.dtype) -> G8Volume:
    """Fused warp + aggregation (reference models/cas_mvsnet.py:30-87) -> G8 cost volume."""
    fdt = ref_nhwc.dtype if isinstance(ref_nhwc, torch.Tensor) else None
    if fdt not in (torch.float32, torch.float16):
        raise ValueError("features must be fp32 or fp16 NHWC tensors")
    _need(ref_nhwc, "ref_nhwc", fdt, 4)
    b, h, w, c = ref_nhwc.shape
    n_src = len(src_nhwc)
    for i, s in enumerate(src_nhwc):
        _need(s, f"src_nhwc[{i}]", fdt, 4)
        if s.shape != ref_nhwc.shape or not s.is_contiguous():
            raise ValueError("source features must be contiguous and shaped like the reference feature")
    _need(rot_trans, "rot_trans", torch.float32, 3)
    if tuple(rot_trans.shape) != (n_src, b, 12):
        raise ValueError(f"rot_trans must be [{n_src},{b},12], got {tuple(rot_trans.shape)}")
    dv, per_pixel, d = _hyp_flags(depth_values, b, h, w)
    if mode == "adaptive":
        _need(wnet, "wnet", torch.float32, 1)
        if wnet.numel() != c + 5:
            raise ValueError("wnet must hold C+5 floats")
        m = AGG_ADAPTIVE
    elif mode == "variance":
        m = AGG_VARIANCE
        wnet = None
    else:
        raise ValueError(f"Don't support {mode}!")
    out = G8Volume.empty(b, c, d, h, w, out_dtype, ref_nhwc.device)
    ptrs = (ctypes.c_void_p * n_src)(*[s.data_ptr() for s in src_nhwc])
    nbytes = (n_src + 1) * b * c * h * w * 4 + dv.numel() * 4 + out.data.numel() * out.data.element_size()
    fn = _lib.load().damvs_warp_agg_fwd_f16 if fdt == torch.float16 else _lib.load().damvs_warp_agg_fwd
    with torch.cuda.device_of(ref_nhwc), _timed("warp_agg", bytes=float(nbytes)):
        _lib.check(fn(_p(ref_nhwc.contiguous()), ptrs, n_src, _p(rot_trans.contiguous()), _p(dv),
                      _p(wnet), _p(out.data), b, c, d, h, w, m, per_pixel, _dt(out_dtype), _stream()))
    return out


def warp_groupwise(ref_nhwc: torch.Tensor, src_nhwc: Sequence[torch.Tensor], rot_trans: torch.Tensor,
                   depth_values: torch.Tensor, groups: int, out_dtype: torch.dtype) -> G8Volume:
    """Fused warp + group-wise correlation (NOT in the reference; BASELINE.json configs[4]) -> G8 volume of
    max(groups, 8) channels: cost[g] = mean_v mean_{c in g} ref[c] * warp_v[c] (csrc/warp_gwc.cu)."""
    fdt = ref_nhwc.dtype if isinstance(ref_nhwc, torch.Tensor) else None
    if fdt not in (torch.float32, torch.float16):
        raise ValueError("features must be fp32 or fp16 NHWC tensors")
    _need(ref_nhwc, "ref_nhwc", fdt, 4)
    b, h, w, c = ref_nhwc.shape
    if c not in (8, 16, 32) or groups not in (4, 8, 16, 32) or groups > c:
        raise ValueError(f"group-wise correlation needs C in (8,16,32), groups in (4,8,16,32) and groups <= C; got C={c}, groups={groups}")
    n_src = len(src_nhwc)
    for i, s in enumerate(src_nhwc):
        _need(s, f"src_nhwc[{i}]", fdt, 4)
        if s.shape != ref_nhwc.shape or not s.is_contiguous():
            raise ValueError("source features must be contiguous and shaped like the reference feature")
    _need(rot_trans, "rot_trans", torch.float32, 3)
    if tuple(rot_trans.shape) != (n_src, b, 12):
        raise ValueError(f"rot_trans must be [{n_src},{b},12], got {tuple(rot_trans.shape)}")
    dv, per_pixel, d = _hyp_flags(depth_values, b, h, w)
    out = G8Volume.empty(b, max(groups, 8), d, h, w, out_dtype, ref_nhwc.device)
    ptrs = (ctypes.c_void_p * n_src)(*[s.data_ptr() for s in src_nhwc])
    nbytes = (n_src + 1) * b * c * h * w * ref_nhwc.element_size() + dv.numel() * 4 + out.data.numel() * out.data.element_size()
    with torch.cuda.device_of(ref_nhwc), _timed("warp_gwc", bytes=float(nbytes)):
        _lib.check(_lib.load().damvs_warp_gwc_fwd(_p(ref_nhwc.contiguous()), ptrs, n_src, _p(rot_trans.contiguous()), _p(dv), _p(out.data),
                                                  b, c, groups, d, h, w, per_pixel, _dt(fdt), _dt(out_dtype), _stream()))
    return out


def conv_desc(b, cin, cout, din, hin, win, stride, transposed, relu, in_dtype, out_dtype, plain_out, impl) -> ConvDesc:
    return ConvDesc(B=b, Cin=cin, Cout=cout, Din=din, Hin=hin, Win=win, stride=stride, transposed=int(transposed),
                    relu=int(relu), in_dtype=_dt(in_dtype), out_dtype=_dt(out_dtype), plain_out=int(plain_out),
                    impl=impl)


def conv3d_pack_weight(weight: torch.Tensor, cin: int, cout: int, transposed: bool, impl: int, stride: int = 1,
                       dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """PyTorch-layout fp32 conv weight -> the packed buffer `impl` consumes (uint8 tensor).
    The tcgen05 buffer embeds the layer's MMA program, which depends on stride / transposed, and holds the weights in the
    2-byte type of the volumes the layer will run on (`dtype`: bf16 or fp16; the direct kernel's weights are fp32)."""
    _need(weight, "weight", torch.float32, 5)
    wdt = dtype if (impl == CONV_TCGEN05 and dtype in HALF_DTYPES) else (torch.bfloat16 if impl == CONV_TCGEN05 else torch.float32)
    desc = conv_desc(1, cin, cout, 1, 1, 1, stride, transposed, 0, wdt, wdt, cout == 1, impl)
    lib = _lib.load()
    nbytes = lib.damvs_conv3d_packed_weight_bytes(ctypes.byref(desc))
    if nbytes == 0:
        _lib.check(2)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    with torch.cuda.device_of(weight):
        _lib.check(lib.damvs_conv3d_pack_weight(ctypes.byref(desc), _p(weight.detach().contiguous()), _p(packed), _stream()))
    return packed


def conv3d(vol: G8Volume, packed: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
           cout: int, stride: int, transposed: bool, relu: bool, skip: Optional[G8Volume], out_dtype: torch.dtype,
           plain_out: bool, impl: int):
    """One fused conv block: out = skip + act(conv(vol) * scale + shift)."""
    b, cin, d, h, w = vol.shape
    if transposed:
        do, ho, wo = 2 * d, 2 * h, 2 * w
    else:
        do, ho, wo = (d - 1) // stride + 1, (h - 1) // stride + 1, (w - 1) // stride + 1
    desc = conv_desc(b, cin, cout, d, h, w, stride, transposed, relu, vol.dtype, out_dtype, plain_out, impl)
    dev = vol.data.device
    if plain_out:
        out_t = torch.empty((b, do, ho, wo), dtype=torch.float32, device=dev)
        result = out_t
    else:
        result = G8Volume.empty(b, cout, do, ho, wo, out_dtype, dev)
        out_t = result.data
        if skip is not None and (skip.data.shape != out_t.shape or skip.dtype != out_dtype):
            raise ValueError("skip volume must match the output volume")
    vout = b * do * ho * wo
    flops = 2.0 * 27 * cin * cout * (b * d * h * w if transposed else vout)
    nbytes = vol.data.numel() * vol.data.element_size() + out_t.numel() * out_t.element_size() * (2 if skip is not None else 1)
    tag = "conv3d_tc" if impl == CONV_TCGEN05 else "conv3d_direct"
    detail = f"{tag} {cin}->{cout} {'T' if transposed else 's%d' % stride} in {d}x{h}x{w}"
    with torch.cuda.device_of(vol.data), _timed(tag, bytes=float(nbytes), flops=flops, detail=detail):
        _lib.check(_lib.load().damvs_conv3d_fwd(ctypes.byref(desc), _p(vol.data), _p(packed), _p(scale), _p(shift),
                                                _p(None if skip is None else skip.data), _p(out_t), _stream()))
    return result


def prob_head_supported(vol: G8Volume, impl: int) -> bool:
    """Whether the `prob` layer + head of this volume can run as one launch (damvs_prob_head_fwd)."""
    b, cin, d, h, w = vol.shape
    if impl != CONV_TCGEN05 or cin != 8 or vol.dtype not in HALF_DTYPES:
        return False
    desc = conv_desc(b, cin, 1, d, h, w, 1, False, False, vol.dtype, torch.float32, True, impl)
    return bool(_lib.load().damvs_prob_head_supported(ctypes.byref(desc)))


def prob_head(vol: G8Volume, packed: torch.Tensor, depth_values: torch.Tensor, impl: int):
    """CostRegNet's `prob` convolution fused with the head: G8 volume [B,8,D,H,W] + per-pixel hypotheses [B,D,H,W] ->
    (prob_volume [B,D,H,W], depth, confidence, variance [B,H,W]); the logits never reach HBM."""
    b, cin, d, h, w = vol.shape
    _need(depth_values, "depth_values", torch.float32, 4)
    if tuple(depth_values.shape) != (b, d, h, w):
        raise ValueError(f"depth_values shape {tuple(depth_values.shape)} does not match [B={b},D={d},H={h},W={w}]")
    dv = depth_values.contiguous()
    desc = conv_desc(b, cin, 1, d, h, w, 1, False, False, vol.dtype, torch.float32, True, impl)
    dev = vol.data.device
    prob = torch.empty((b, d, h, w), dtype=torch.float32, device=dev)
    depth = torch.empty((b, h, w), dtype=torch.float32, device=dev)
    conf = torch.empty_like(depth)
    var = torch.empty_like(depth)
    flops = 2.0 * 27 * cin * b * d * h * w
    nbytes = vol.data.numel() * vol.data.element_size() + 2 * dv.numel() * 4 + 3 * b * h * w * 4
    detail = f"conv_head {cin}->1 s1 + head in {d}x{h}x{w}"
    with torch.cuda.device_of(vol.data), _timed("conv_head", bytes=float(nbytes), flops=flops, detail=detail):
        _lib.check(_lib.load().damvs_prob_head_fwd(ctypes.byref(desc), _p(vol.data), _p(packed), _p(dv), _p(prob), _p(depth), _p(conf),
                                                   _p(var), _stream()))
    return prob, depth, conf, var


def softmax_regress(logits: torch.Tensor, depth_values: torch.Tensor, want_prob: bool = True):
    """logits [B,D,H,W] -> (prob [B,D,H,W] | None, depth, confidence, variance [B,H,W])."""
    _need(logits, "logits", torch.float32, 4)
    logits = logits.contiguous()
    b, d, h, w = logits.shape
    dv, per_pixel, d2 = _hyp_flags(depth_values, b, h, w)
    if d2 != d:
        raise ValueError(f"depth_values.shape[1]:{d2}  num_depth:{d}")
    dev = logits.device
    prob = torch.empty_like(logits) if want_prob else None
    depth = torch.empty((b, h, w), dtype=torch.float32, device=dev)
    conf = torch.empty_like(depth)
    var = torch.empty_like(depth)
    nbytes = logits.numel() * 4 + dv.numel() * 4 + (logits.numel() * 4 if want_prob else 0) + 3 * b * h * w * 4
    with torch.cuda.device_of(logits), _timed("head", bytes=float(nbytes)):
        _lib.check(_lib.load().damvs_softmax_regress_fwd(_p(logits), _p(dv), _p(prob), _p(depth), _p(conf), _p(var),
                                                         b, d, h, w, per_pixel, _stream()))
    return prob, depth, conf, var


def stage_hypotheses(prev_depth: torch.Tensor, prev_var: torch.Tensor, ndepth: int, height: int, width: int,
                     scale: int) -> torch.Tensor:
    """Fused hypothesis sampling for cascade stages 2/3: previous-stage depth / variance [B,hp,wp] ->
    depth_values [B,D,height/scale,width/scale] (reference models/cas_mvsnet.py:250-253, 269-274, 293-296 +
    models/module.py:1012-1036 in one kernel)."""
    _need(prev_depth, "prev_depth", torch.float32, 3)
    _need(prev_var, "prev_var", torch.float32, 3)
    if prev_var.shape != prev_depth.shape:
        raise ValueError("prev_depth and prev_var must have the same shape")
    assert ndepth > 1
    b, hp, wp = prev_depth.shape
    out = torch.empty((b, ndepth, height // scale, width // scale), dtype=torch.float32, device=prev_depth.device)
    with torch.cuda.device_of(prev_depth):
        _lib.check(_lib.load().damvs_uncertainty_samples_fwd(_p(prev_depth.contiguous()), _p(prev_var.contiguous()), _p(out), b, hp, wp,
                                                             ndepth, height, width, scale, _stream()))
    return out


def depth_regression(p: torch.Tensor, depth_values: torch.Tensor) -> torch.Tensor:
    _need(p, "p", torch.float32, 4)
    p = p.contiguous()
    b, d, h, w = p.shape
    dv, per_pixel, d2 = _hyp_flags(depth_values, b, h, w)
    if d2 != d:
        raise ValueError("depth_values / probability depth mismatch")
    out = torch.empty((b, h, w), dtype=torch.float32, device=p.device)
    with torch.cuda.device_of(p):
        _lib.check(_lib.load().damvs_depth_regression_fwd(_p(p), _p(dv), _p(out), b, d, h, w, per_pixel, _stream()))
    return out


# --------------------------------------------------------------------------
# per-call device timing (bench.py's roofline leg)
# --------------------------------------------------------------------------
class CallTimer:
    """Records a CUDA-event pair around every library call made while active.

    Events are recorded on torch's current stream, which is the stream the
    kernels are launched on (``_stream()``), so the durations are device times.
    """
    active: Optional["CallTimer"] = None

    def __init__(self):
        self.records = []  # (tag, start_event, end_event, meta)

    def __enter__(self):
        CallTimer.active = self
        return self

    def __exit__(self, *exc):
        CallTimer.active = None

    def summary(self, detail: bool = False):
        torch.cuda.synchronize()
        out = {}
        for tag, s, e, meta in self.records:
            if detail:
                tag = meta.get("detail", tag)
            d = out.setdefault(tag, {"ms": 0.0, "calls": 0, "bytes": 0.0, "flops": 0.0})
            d["ms"] += s.elapsed_time(e)
            d["calls"] += 1
            d["bytes"] += meta.get("bytes", 0.0)
            d["flops"] += meta.get("flops", 0.0)
        return out


@contextlib.contextmanager
def _timed(tag: str, **meta):
    t = CallTimer.active
    if t is None:
        yield
        return
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    try:
        yield
    finally:
        e.record()
        t.records.append((tag, s, e, meta))
