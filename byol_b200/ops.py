"""Thin Python wrappers over the C-ABI kernels (``include/byol_b200.h``).

PyTorch is used only as the owner of device memory and of the CUDA stream: every function borrows
``tensor.data_ptr()`` for the duration of one asynchronous launch on ``torch.cuda.current_stream()``.
There is no CPU / ATen fallback; calling any of these without the CUDA extension raises at import time.

Layout conventions: activations NHWC bf16 with channels padded to a multiple of 8; weights for the
tensor-core kernels bf16 ``[Cout, KH*KW*Cpad]`` (fprop) / ``[Cin, KH*KW*Cout]`` (dgrad) made by
:func:`prep_weight` from the fp32 master in the reference's ``[Cout, Cin, KH, KW]`` layout.
"""
import torch

from ._lib import lib, check

BF16 = torch.bfloat16
F32 = torch.float32


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (byol_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)


def conv_out_size(h, k, s, p):
    return (h + 2 * p - k) // s + 1


# ------------------------------------------------------------------------------------------------
# layout / weights
# ------------------------------------------------------------------------------------------------
def nchw_to_nhwc8(x, out=None):
    """fp32 NCHW [N, C<=8, H, W] -> bf16 NHWC [N, H, W, 8] (zero padded channels)."""
    _chk(x, F32, "x")
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty((n, h, w, 8), dtype=BF16, device=x.device)
    check(lib.byol_nchw_to_nhwc8(_ptr(x), _ptr(out), n, c, h, w, _stream()), "byol_nchw_to_nhwc8")
    return out


def stem4_supported(cin, cout, h, w, k, stride, pad):
    """True if the dedicated stem kernels (7x7/s2/p3, <= 4 -> 64 channels, W <= 256) handle this geometry."""
    return bool(lib.byol_stem4_supported(cin, cout, h, w, k, stride, pad))


def nchw_to_stem4(x, out=None):
    """fp32 NCHW [N, C<=4, H, W] -> zero-padded bf16 NHWC4 [N, H+6, Wp, 4] (input format of the stem kernels)."""
    _chk(x, F32, "x")
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty((n, h + 6, lib.byol_stem4_row_pixels(), 4), dtype=BF16, device=x.device)
    check(lib.byol_nchw_to_stem4(_ptr(x), _ptr(out), n, c, h, w, _stream()), "byol_nchw_to_stem4")
    return out


def prep_weight_stem4(w, out=None):
    """fp32 [64, Cin, 7, 7] -> bf16 [7, 4, 64, 8] (stem kernel layout)."""
    _chk(w, F32, "w")
    if out is None:
        out = torch.empty(7 * 4 * 64 * 8, dtype=BF16, device=w.device)
    check(lib.byol_prep_weight_stem4(_ptr(w), _ptr(out), w.shape[1], _stream()), "byol_prep_weight_stem4")
    return out


def stem_conv_fprop(xs4, w_stem4, h, w, stats=None):
    """y[N, H/2, W/2, 64] = conv7x7/s2/p3 over the padded NHWC4 image; stats (fp32 [128], zeroed): += sum | sum sq."""
    _chk(xs4, BF16, "xs4"); _chk(w_stem4, BF16, "w_stem4")
    n = xs4.shape[0]
    y = torch.empty((n, h // 2, w // 2, 64), dtype=BF16, device=xs4.device)
    cs = _ptr(stats) if stats is not None else 0
    cq = _ptr(stats[64:]) if stats is not None else 0
    check(lib.byol_stem_conv_fprop(_ptr(xs4), _ptr(w_stem4), _ptr(y), cs, cq, n, h, w, _stream()),
          "byol_stem_conv_fprop")
    return y


def stem_conv_wgrad(xs4, dy, dw, h, w):
    """dw[64, Cin, 7, 7] (fp32) += dy^T * im2col(image) for the 7x7/s2/p3 stem; xs4 from nchw_to_stem4."""
    _chk(xs4, BF16, "xs4"); _chk(dy, BF16, "dy"); _chk(dw, F32, "dw")
    check(lib.byol_stem_conv_wgrad(_ptr(xs4), _ptr(dy), _ptr(dw), xs4.shape[0], dw.shape[1], h, w, _stream()),
          "byol_stem_conv_wgrad")
    return dw


def subsample2(x):
    """x[N,H,W,C] -> x[:, ::2, ::2, :] compacted (H, W even): the input of a 1x1 / stride-2 convolution."""
    _chk(x, BF16, "x")
    n, h, w, c = x.shape
    y = torch.empty((n, h // 2, w // 2, c), dtype=BF16, device=x.device)
    check(lib.byol_subsample2(_ptr(x), _ptr(y), n, h, w, c, _stream()), "byol_subsample2")
    return y


def prep_weight(w, cpad=None, want_dgrad=True, out_f=None, out_d=None):
    """fp32 [Cout, Cin, KH, KW] (or [out, in] for Linear) -> (w_fprop bf16 [Cout, taps*Cpad], w_dgrad bf16 [Cin, taps*Cout])."""
    _chk(w, F32, "w")
    if w.dim() == 2:
        cout, cin = w.shape
        kh = kw = 1
    else:
        cout, cin, kh, kw = w.shape
    if cpad is None:
        cpad = (cin + 7) // 8 * 8
    if out_f is None:
        out_f = torch.empty((cout, kh * kw * cpad), dtype=BF16, device=w.device)
    if want_dgrad and out_d is None:
        out_d = torch.empty((cin, kh * kw * cout), dtype=BF16, device=w.device)
    check(lib.byol_prep_weight(_ptr(w), _ptr(out_f), _ptr(out_d) if want_dgrad else 0, cout, cin, cpad, kh, kw,
                               _stream()), "byol_prep_weight")
    return out_f, (out_d if want_dgrad else None)


def prep_weight_fold(w, out_f=None):
    """Stem layout: fp32 [Cout, Cin<=8, KH, KW<=8] -> bf16 [Cout, KH*64] (column = kh*64 + kw*8 + c)."""
    _chk(w, F32, "w")
    cout, cin, kh, kw = w.shape
    if out_f is None:
        out_f = torch.empty((cout, kh * 64), dtype=BF16, device=w.device)
    check(lib.byol_prep_weight_fold(_ptr(w), _ptr(out_f), cout, cin, kh, kw, _stream()), "byol_prep_weight_fold")
    return out_f


def prep_weights_multi(flat, pool_f, pool_d, desc, num_blocks=None):
    """All weights of one parameter set (flat fp32 vector) -> bf16 tensor-core layouts, one launch.
    num_blocks: sum of prep_unit_blocks over the rows of `desc` (computed from the device table when not given)."""
    if num_blocks is None:
        num_blocks = prep_blocks(desc.cpu().tolist())
    check(lib.byol_prep_weights_multi(_ptr(flat), _ptr(pool_f), _ptr(pool_d), _ptr(desc), desc.shape[0], num_blocks,
                                      _stream()), "byol_prep_weights_multi")


def prep_blocks(rows):
    """Grid size for prep_weights_multi: rows = [[src, dstf, dstd, Cout, Cin, Cpad, taps, fold], ...]."""
    return sum(lib.byol_prep_unit_blocks(int(r[3]), int(r[4]), int(r[5]), int(r[6]), int(r[7])) for r in rows)


def cast_bf16_pitched(x2d, ldy):
    """fp32 [rows, cols] (any row stride) -> bf16 [rows, ldy] with zero padding columns (ldy >= cols)."""
    if x2d.dtype != F32 or not x2d.is_cuda or x2d.stride(1) != 1:
        raise ValueError("cast_bf16_pitched: need a CUDA fp32 matrix with unit column stride")
    rows, cols = x2d.shape
    out = torch.empty((rows, ldy), dtype=BF16, device=x2d.device)
    check(lib.byol_cast_f32_bf16_2d(_ptr(x2d), _ptr(out), rows, cols, x2d.stride(0), ldy, _stream()),
          "byol_cast_f32_bf16_2d")
    return out


def cast_bf16(x, out=None):
    _chk(x, F32, "x")
    if out is None:
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
    check(lib.byol_cast_f32_bf16(_ptr(x), _ptr(out), x.numel(), _stream()), "byol_cast_f32_bf16")
    return out


# ------------------------------------------------------------------------------------------------
# tensor-core convolution / linear
# ------------------------------------------------------------------------------------------------
def conv_fprop(x, w_f, kh, kw, stride, pad, bias=None, resid=None, stats=None, relu=False, out_fp32=False,
               out=None, force_gather=False):
    """y[N,Ho,Wo,Cout] = conv(x[N,H,W,C], w).  stats: optional zeroed fp32 [2*Cout] receiving column sum / sqsum."""
    _chk(x, BF16, "x"); _chk(w_f, BF16, "w_f"); _chk(bias, F32, "bias"); _chk(resid, BF16, "resid")
    _chk(stats, F32, "stats")
    n, h, w, c = x.shape
    cout, ldw = w_f.shape
    ho, wo = conv_out_size(h, kh, stride, pad), conv_out_size(w, kw, stride, pad)
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=F32 if out_fp32 else BF16, device=x.device)
    cs = _ptr(stats)
    cq = (stats.data_ptr() + 4 * cout) if stats is not None else 0
    check(lib.byol_conv_igemm(_ptr(x), _ptr(w_f), _ptr(out), _ptr(resid), 0, 0, _ptr(bias), cs, cq, n, h, w, c, ho, wo,
                              cout, kh, kw, stride, pad, 0, ldw, cout, int(out_fp32), int(relu), int(force_gather),
                              _stream()), "byol_conv_igemm(fprop)")
    return out


def conv_dgrad(dy, w_d, h, w, kh, kw, stride, pad, resid=None, out=None, force_gather=False, resid_mask=None,
               resid_up=False):
    """dx[N,H,W,Cin] = conv_transpose(dy[N,Ho,Wo,Cout], w) (+ resid, optionally only where the bits of resid_mask,
    the uint8 ReLU mask written by bn_apply, are set; resid_up: resid is the compact [N, h/2, w/2, Cin] gradient of a
    stride-2 branch, added to the even pixels);  w_d is the dgrad layout [Cin, taps*Cout]."""
    _chk(dy, BF16, "dy"); _chk(w_d, BF16, "w_d"); _chk(resid, BF16, "resid"); _chk(resid_mask, torch.uint8, "mask")
    n, ho, wo, cout = dy.shape
    cin, ldw = w_d.shape
    if out is None:
        out = torch.empty((n, h, w, cin), dtype=BF16, device=dy.device)
    check(lib.byol_conv_igemm(_ptr(dy), _ptr(w_d), _ptr(out), _ptr(resid), _ptr(resid_mask), int(resid_up), 0, 0, 0, n, ho, wo, cout, h, w, cin,
                              kh, kw, stride, pad, 1, ldw, cin, 0, 0, int(force_gather), _stream()),
          "byol_conv_igemm(dgrad)")
    return out


def conv_wgrad(x, dy, dw, kh, kw, stride, pad, force_gather=False):
    """dw[Cout,Cin,KH,KW] (fp32, reference layout) += dy^T * im2col(x).  x: [N,H,W,Cpad], dy: [N,Ho,Wo,ldy] whose
    first Cout = dw.shape[0] columns are the gradient (ldy > Cout: pitched rows, e.g. a 10-class classifier)."""
    _chk(x, BF16, "x"); _chk(dy, BF16, "dy"); _chk(dw, F32, "dw")
    n, h, w, c = x.shape
    _, ho, wo, ldy = dy.shape
    cout, cin_real = dw.shape[0], dw.shape[1]
    check(lib.byol_conv_wgrad(_ptr(x), _ptr(dy), _ptr(dw), n, h, w, c, cin_real, ho, wo, cout, ldy, kh, kw, stride,
                              pad, int(force_gather), _stream()), "byol_conv_wgrad")
    return dw


def gemm_fused(x2d, w_f, out=None, colscale=None, bias=None, resid=None, resid_mask=None, resid_colscale=None,
               relu=False, mask_out=None, stats=None, no_store=False, bwd_reduce=False):
    """1x1 / stride-1 convolution as a plain GEMM y[M, N] = x2d[M, K] @ w_f[N, K]^T with the fused BatchNorm epilogue
    of byol_conv_igemm_fused: t = y*colscale + bias; out = act(t + resid_colscale * masked(resid)).
    no_store: only `stats` (zeroed fp32 [2N]: column sum | sum of squares of bf16(y)) is produced.
    bwd_reduce: only `stats` is produced: [sum dz | sum dz*t] with dz = masked(resid)."""
    _chk(x2d, BF16, "x"); _chk(w_f, BF16, "w_f"); _chk(resid, BF16, "resid"); _chk(resid_mask, torch.uint8, "mask")
    _chk(stats, F32, "stats"); _chk(mask_out, torch.uint8, "mask_out")
    m, k = x2d.shape
    n, ldw = w_f.shape
    if out is None and not (no_store or bwd_reduce):
        out = torch.empty((m, n), dtype=BF16, device=x2d.device)
    cs = _ptr(stats)
    cq = (stats.data_ptr() + 4 * n) if stats is not None else 0
    check(lib.byol_conv_igemm_fused(_ptr(x2d), _ptr(w_f), _ptr(out), _ptr(resid), _ptr(resid_mask), _ptr(colscale),
                                    _ptr(bias), _ptr(resid_colscale), _ptr(mask_out), cs, cq, m, k, n, ldw, n,
                                    int(relu), int(no_store), int(bwd_reduce), _stream()), "byol_conv_igemm_fused")
    return out


def mlp_fused_supported(b, k1, h, o):
    return bool(lib.byol_mlp_fused_supported(b, k1, h, o))


def mlp_fused_fwd(x2d, w1, b1, gamma, beta, w2, b2, stats, running_mean, running_var, momentum, eps, count, coeffs,
                  grid_bar, train, save, peer=None):
    """Linear -> BatchNorm1d -> ReLU -> Linear of one lane in ONE cooperative kernel (csrc/mlp_fused.cu).
    stats: zeroed fp32 [2H] (train); coeffs: fp32 [4, H] out; grid_bar: persistent zeroed int32 [2] (one per stream);
    peer: comm.PeerExchange of this stream's channel under SyncBatchNorm (world > 1), else None.
    Returns (out fp32 [B, O], h bf16 [B, H] | None, a bf16 [B, H] | None)."""
    _chk(x2d, BF16, "x"); _chk(w1, BF16, "w1"); _chk(w2, BF16, "w2"); _chk(stats, F32, "stats")
    b, k1 = x2d.shape
    h, ldw1 = w1.shape
    o, ldw2 = w2.shape
    dev = x2d.device
    out = torch.zeros((b, o), dtype=F32, device=dev)
    hs = torch.empty((b, h), dtype=BF16, device=dev) if save else None
    as_ = torch.empty((b, h), dtype=BF16, device=dev) if save else None
    pp, world, rank, cap, ctr = 0, 1, 0, 0, 0
    if peer is not None:
        pp, world, rank, cap, ctr = peer.ptrs, peer.world, peer.rank, peer.cap_bytes, peer.counter.data_ptr()
    check(lib.byol_mlp_fused_fwd(_ptr(x2d), _ptr(w1), _ptr(b1), _ptr(gamma), _ptr(beta), _ptr(w2), _ptr(b2), _ptr(stats),
                                 _ptr(running_mean), _ptr(running_var), float(momentum), float(eps), float(count),
                                 _ptr(coeffs), _ptr(out), _ptr(hs), _ptr(as_), _ptr(grid_bar), b, k1, h, o, ldw1, ldw2,
                                 int(train), pp, world, rank, cap, ctr, _stream()), "byol_mlp_fused_fwd")
    return out, hs, as_


def bn_bwd_prep(mean, invstd, out=None):
    c = mean.numel()
    if out is None:
        out = torch.empty(2 * c, dtype=F32, device=mean.device)
    check(lib.byol_bn_bwd_prep(_ptr(mean), _ptr(invstd), _ptr(out), c, _stream()), "byol_bn_bwd_prep")
    return out


def bn_bwd_coeffs(s12, coeffs, gamma, count, s12_local=None, dgamma=None, dbeta=None, out=None):
    """[B | Cc | A] per channel so that dy = A*dz + B*y + Cc (see csrc/bn.cu); accumulates dgamma / dbeta."""
    c = gamma.numel()
    if out is None:
        out = torch.empty(3 * c, dtype=F32, device=gamma.device)
    check(lib.byol_bn_bwd_coeffs(_ptr(s12), _ptr(s12_local), _ptr(coeffs[2]), _ptr(coeffs[3]), _ptr(gamma),
                                 float(count), _ptr(out), _ptr(dgamma), _ptr(dbeta), c, _stream()),
          "byol_bn_bwd_coeffs")
    return out


def linear_fprop(x2d, w_f, bias=None, stats=None, relu=False, out_fp32=False, out=None):
    """y[M, out] = x[M, in] @ W^T (+bias): a 1x1 'convolution' over M pixels."""
    m, k = x2d.shape
    y = conv_fprop(x2d.view(m, 1, 1, k), w_f, 1, 1, 1, 0, bias=bias, stats=stats, relu=relu, out_fp32=out_fp32,
                   out=None if out is None else out.view(m, 1, 1, -1))
    return y.view(m, -1)


def linear_dgrad(dy2d, w_d, out=None):
    m, n = dy2d.shape
    dx = conv_dgrad(dy2d.view(m, 1, 1, n), w_d, 1, 1, 1, 1, 1, 0, out=None if out is None else out.view(m, 1, 1, -1))
    return dx.view(m, -1)


def linear_wgrad(x2d, dy2d, dw):
    m, k = x2d.shape
    n = dy2d.shape[1]
    conv_wgrad(x2d.view(m, 1, 1, k), dy2d.view(m, 1, 1, n), dw.view(n, dw.shape[1], 1, 1), 1, 1, 1, 0)
    return dw


# ------------------------------------------------------------------------------------------------
# batch norm
# ------------------------------------------------------------------------------------------------
def bn_stats(x2d, stats):
    """stats (zeroed fp32 [2C]) += [column sums, column sums of squares] of x2d [M, C] bf16."""
    _chk(x2d, BF16, "x"); _chk(stats, F32, "stats")
    m, c = x2d.shape
    check(lib.byol_bn_stats(_ptr(x2d), _ptr(stats), m, c, _stream()), "byol_bn_stats")
    return stats


def bn_finalize(stats, count, gamma, beta, running_mean, running_var, momentum, eps, coeffs):
    """Single-lane form of :func:`bn_finalize_lanes`; coeffs: fp32 [4, C] receiving scale, shift, mean, invstd."""
    bn_finalize_lanes(stats, count, [gamma], [beta], running_mean, running_var, momentum, eps, coeffs)
    return coeffs


def bn_finalize_lanes(stats, count, gammas, betas, running_mean, running_var, momentum, eps, coeffs):
    """One launch for all lock-step lanes: stats [L*2C], coeffs [L,4,C]; gammas/betas: per-lane fp32 [C] tensors."""
    L = len(gammas)
    c = gammas[0].numel()
    g = [_ptr(t) for t in gammas] + [0] * (4 - L)
    b = [_ptr(t) for t in betas] + [0] * (4 - L)
    check(lib.byol_bn_finalize_lanes(_ptr(stats), float(count), L, g[0], b[0], g[1], b[1], g[2], b[2], g[3], b[3],
                                     _ptr(running_mean), _ptr(running_var), float(momentum), float(eps),
                                     _ptr(coeffs), c, _stream()), "byol_bn_finalize_lanes")
    return coeffs


def bn_eval_coeffs(gamma, beta, running_mean, running_var, eps, coeffs):
    c = gamma.numel()
    check(lib.byol_bn_eval_coeffs(_ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), float(eps),
                                  _ptr(coeffs[0]), _ptr(coeffs[1]), c, _stream()), "byol_bn_eval_coeffs")
    return coeffs


def bn_apply(x2d, scale, shift, relu, resid=None, rscale=None, rshift=None, out=None, out_f32=None, mask_out=None):
    """y = act(x*scale + shift (+ resid | resid*rscale + rshift)); mask_out (uint8 [M*C/8]) receives the bits y > 0."""
    _chk(x2d, BF16, "x"); _chk(resid, BF16, "resid"); _chk(mask_out, torch.uint8, "mask_out")
    m, c = x2d.shape
    if out is None and out_f32 is None:
        out = torch.empty_like(x2d)
    check(lib.byol_bn_apply(_ptr(x2d), _ptr(scale), _ptr(shift), _ptr(resid), _ptr(rscale), _ptr(rshift), _ptr(out),
                            _ptr(out_f32), _ptr(mask_out), m, c, int(relu), _stream()), "byol_bn_apply")
    return out if out is not None else out_f32


def bn_bwd_reduce(g, x, coeffs, s12, mask_mode, act=None):
    """s12 (zeroed fp32 [2C]) += [sum dz, sum dz*xhat];  mask_mode 0 none / 1 relu(x*scale+shift) / 2 act>0 /
    3 mask bits (act = the uint8 mask written by bn_apply)."""
    _chk(g, BF16, "g"); _chk(x, BF16, "x"); _chk(act, torch.uint8 if mask_mode == 3 else BF16, "act")
    m, c = x.shape
    check(lib.byol_bn_bwd_reduce(_ptr(g), _ptr(x), _ptr(act), _ptr(coeffs[0]), _ptr(coeffs[1]), _ptr(coeffs[2]),
                                 _ptr(coeffs[3]), _ptr(s12), m, c, mask_mode, _stream()), "byol_bn_bwd_reduce")
    return s12


def bn_bwd_apply(g, x, coeffs, gamma, s12, count, mask_mode, act=None, dy=None, dz_out=None, s12_local=None,
                 dgamma=None, dbeta=None):
    """dy = gamma*invstd*(dz - s1/n - xhat*s2/n) with the (global) sums s12 over `count` rows; when dgamma/dbeta
    are given they are incremented by the rank-local sums (s12_local, default s12)."""
    _chk(g, BF16, "g"); _chk(x, BF16, "x"); _chk(act, torch.uint8 if mask_mode == 3 else BF16, "act")
    m, c = x.shape
    if dy is None:
        dy = torch.empty_like(x)
    check(lib.byol_bn_bwd_apply(_ptr(g), _ptr(x), _ptr(act), _ptr(coeffs[0]), _ptr(coeffs[1]), _ptr(coeffs[2]),
                                _ptr(coeffs[3]), _ptr(gamma), _ptr(s12), float(count), _ptr(dy), _ptr(dz_out), m, c,
                                mask_mode, _ptr(s12_local), _ptr(dgamma), _ptr(dbeta), _stream()),
          "byol_bn_bwd_apply")
    return dy


def col_sum(x2d, out):
    """out[c] (fp32) += sum_r x2d[r, c]."""
    m, c = x2d.shape
    check(lib.byol_col_sum(_ptr(x2d), _ptr(out), m, c, x2d.stride(0), int(x2d.dtype == F32), _stream()),
          "byol_col_sum")
    return out


# ------------------------------------------------------------------------------------------------
# pooling
# ------------------------------------------------------------------------------------------------
def maxpool_fwd(x, k=3, s=2, p=1, want_idx=True):
    _chk(x, BF16, "x")
    n, h, w, c = x.shape
    ho, wo = conv_out_size(h, k, s, p), conv_out_size(w, k, s, p)
    y = torch.empty((n, ho, wo, c), dtype=BF16, device=x.device)
    idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device) if want_idx else None
    check(lib.byol_maxpool_fwd(_ptr(x), _ptr(y), _ptr(idx), n, h, w, c, k, s, p, _stream()), "byol_maxpool_fwd")
    return y, idx


def bn_relu_maxpool_fwd(x, scale, shift, k=3, s=2, p=1, want_idx=True):
    """maxpool(relu(x*scale + shift)) in one pass (stem); bit-identical to bn_apply(relu) + maxpool_fwd."""
    _chk(x, BF16, "x")
    n, h, w, c = x.shape
    ho, wo = conv_out_size(h, k, s, p), conv_out_size(w, k, s, p)
    y = torch.empty((n, ho, wo, c), dtype=BF16, device=x.device)
    idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device) if want_idx else None
    check(lib.byol_bn_relu_maxpool_fwd(_ptr(x), _ptr(scale), _ptr(shift), _ptr(y), _ptr(idx), n, h, w, c, k, s, p,
                                       _stream()), "byol_bn_relu_maxpool_fwd")
    return y, idx


def maxpool_bwd(dy, idx, h, w, k=3, s=2, p=1):
    _chk(dy, BF16, "dy")
    n, ho, wo, c = dy.shape
    dx = torch.empty((n, h, w, c), dtype=BF16, device=dy.device)
    check(lib.byol_maxpool_bwd(_ptr(dy), _ptr(idx), _ptr(dx), n, h, w, c, k, s, p, _stream()), "byol_maxpool_bwd")
    return dx


def avgpool_fwd(x, want_f32=True, want_bf16=True):
    _chk(x, BF16, "x")
    n, h, w, c = x.shape
    yf = torch.empty((n, c), dtype=F32, device=x.device) if want_f32 else None
    yb = torch.empty((n, c), dtype=BF16, device=x.device) if want_bf16 else None
    check(lib.byol_avgpool_fwd(_ptr(x), _ptr(yf), _ptr(yb), n, h * w, c, _stream()), "byol_avgpool_fwd")
    return yf, yb


def avgpool_bwd(g_bf16, g_f32, n, h, w, c):
    _chk(g_bf16, BF16, "g_bf16"); _chk(g_f32, F32, "g_f32")
    dev = (g_bf16 if g_bf16 is not None else g_f32).device
    dx = torch.empty((n, h, w, c), dtype=BF16, device=dev)
    check(lib.byol_avgpool_bwd(_ptr(g_bf16), _ptr(g_f32), _ptr(dx), n, h * w, c, _stream()), "byol_avgpool_bwd")
    return dx


# ------------------------------------------------------------------------------------------------
# objective / EMA / LARS
# ------------------------------------------------------------------------------------------------
def loss_fwd(q1, q2, z1, z2, workspace, loss, saved):
    for t, nm in ((q1, "q1"), (q2, "q2"), (z1, "z1"), (z2, "z2")):
        _chk(t, F32, nm)
    rows, dim = q1.shape
    check(lib.byol_loss_fwd(_ptr(q1), _ptr(q2), _ptr(z1), _ptr(z2), rows, dim, _ptr(workspace), _ptr(loss),
                            _ptr(saved), _stream()), "byol_loss_fwd", kernels=2)
    return loss


def loss_bwd(q1, q2, z1, z2, saved, grad_out, dq1, dq2):
    rows, dim = q1.shape
    check(lib.byol_loss_bwd(_ptr(q1), _ptr(q2), _ptr(z1), _ptr(z2), _ptr(saved), _ptr(grad_out), _ptr(dq1),
                            _ptr(dq2), rows, dim, _stream()), "byol_loss_bwd")
    return dq1, dq2


def ema_update(x, mean, one_minus_decay, decay):
    """mean <- fl(fl(a*x) + fl(d*mean)) in place; a, d already rounded to fp32 by the caller."""
    _chk(x, F32, "x"); _chk(mean, F32, "mean")
    if x.numel() != mean.numel():
        raise ValueError("ema_update: size mismatch")
    check(lib.byol_ema_update(_ptr(x), _ptr(mean), float(one_minus_decay), float(decay), x.numel(), _stream()),
          "byol_ema_update")
    return mean


def lars_sgd_step(table, trust_coef, eps, momentum, first_step):
    """table: dict of device tensors p_ptrs/g_ptrs/m_ptrs (int64 pointer tables, m_ptrs may be None),
    chunk_start (int64), chunk_len / chunk_tensor / tensor_first_chunk (int32), wd / lr (fp32), ignore (int32),
    partial (fp64 [2 * chunks])."""
    check(lib.byol_lars_sgd_step(_ptr(table["p_ptrs"]), _ptr(table["g_ptrs"]), _ptr(table.get("m_ptrs")),
                                 _ptr(table["chunk_start"]), _ptr(table["chunk_len"]), _ptr(table["chunk_tensor"]),
                                 table["chunk_start"].numel(), _ptr(table["tensor_first_chunk"]), _ptr(table["wd"]),
                                 _ptr(table["lr"]), _ptr(table["ignore"]), table["wd"].numel(),
                                 _ptr(table["partial"]), float(trust_coef), float(eps), float(momentum),
                                 int(first_step), _stream()),
          "byol_lars_sgd_step", kernels=2)


def ce_topk_fwd(logits, labels, scratch=None):
    """Softmax cross-entropy (mean) + top-1 / top-5 accuracy (%) of fp32 logits [R, C] in one launch; `labels` has R
    entries or a divisor of R (row r uses labels[r % len]: both views of a sample share its label).
    Returns (out fp32 [3] = loss, top1, top5; row_lse fp32 [R] for the backward pass)."""
    _chk(logits, F32, "logits")
    if labels.dtype != torch.int64 or not labels.is_cuda or not labels.is_contiguous():
        raise ValueError("labels must be a contiguous CUDA int64 tensor")
    r, c = logits.shape
    if labels.numel() == 0 or r % labels.numel() != 0:
        raise ValueError("labels: %d entries do not tile %d rows" % (labels.numel(), r))
    dev = logits.device
    fl = torch.empty(2 * r + 3, dtype=F32, device=dev)            # row_lse | row_loss | out
    it = torch.zeros(r + 1, dtype=torch.int32, device=dev)        # row_rank | ticket (must start at 0)
    check(lib.byol_ce_topk_fwd(_ptr(logits), _ptr(labels), labels.numel(), r, c, logits.stride(0), _ptr(fl), _ptr(fl[r:]), _ptr(it),
                               _ptr(it[r:]), _ptr(fl[2 * r:]), _stream()), "byol_ce_topk_fwd")
    return fl[2 * r:], fl[:r]


def ce_bwd(logits, labels, row_lse, grad_out):
    r, c = logits.shape
    d = torch.empty((r, c), dtype=F32, device=logits.device)
    check(lib.byol_ce_bwd(_ptr(logits), _ptr(labels), labels.numel(), _ptr(row_lse), _ptr(grad_out), r, c, logits.stride(0), _ptr(d), c,
                          _stream()), "byol_ce_bwd")
    return d


# ------------------------------------------------------------------------------------------------
# fp32-accurate forward path ("split-bf16"; csrc/split.cu)
# ------------------------------------------------------------------------------------------------
F64 = torch.float64


def split_planes(x2d, T, cpad=None, want_copy=False, copy_out=None):
    """fp32 [M, C] (unit column stride) -> bf16 planes [M, T*cpad] (+ the plain bf16 rounding [M, C])."""
    if x2d.dtype != F32 or not x2d.is_cuda or x2d.stride(-1) != 1:
        raise ValueError("split_planes: need a CUDA fp32 matrix with unit column stride")
    m, c = x2d.shape
    cpad = cpad or c
    planes = torch.empty((m, T * cpad), dtype=BF16, device=x2d.device)
    copy = copy_out if copy_out is not None else \
        (torch.empty((m, c), dtype=BF16, device=x2d.device) if want_copy else None)
    _chk(copy, BF16, "copy_out")
    check(lib.byol_split_planes(_ptr(x2d), _ptr(planes), _ptr(copy), m, c, cpad, x2d.stride(0), T, _stream()),
          "byol_split_planes")
    return planes, copy


def nchw_to_planes(x, T, cpad=8, out=None):
    """fp32 NCHW [N, C<=cpad, H, W] -> bf16 NHWC planes [N, H, W, T*cpad]."""
    _chk(x, F32, "x")
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty((n, h, w, T * cpad), dtype=BF16, device=x.device)
    check(lib.byol_nchw_to_planes(_ptr(x), _ptr(out), n, c, h, w, cpad, T, _stream()), "byol_nchw_to_planes")
    return out


def prep_weight_planes(w, T, cpad, out):
    """fp32 [Cout, Cin, KH, KW] / [out, in] -> bf16 [Cout, taps*T*cpad] with the weight-side plane pattern."""
    _chk(w, F32, "w")
    cout, cin = w.shape[0], w.shape[1]
    taps = w.numel() // (cout * cin)
    check(lib.byol_prep_weight_planes(_ptr(w), _ptr(out), cout, cin, cpad, taps, T, _stream()),
          "byol_prep_weight_planes")
    return out


def stats_f32(y2d, stats64):
    """stats64 (zeroed fp64 [2C]) += [column sums, column sums of squares] of the fp32 matrix y2d."""
    _chk(y2d, F32, "y"); _chk(stats64, F64, "stats")
    m, c = y2d.shape
    check(lib.byol_stats_f32(_ptr(y2d), _ptr(stats64), m, c, _stream()), "byol_stats_f32")
    return stats64


def bn_finalize_lanes_f64(stats64, count, gammas, betas, running_mean, running_var, momentum, eps, coeffs):
    L = len(gammas)
    c = gammas[0].numel()
    g = [_ptr(t) for t in gammas] + [0] * (4 - L)
    b = [_ptr(t) for t in betas] + [0] * (4 - L)
    check(lib.byol_bn_finalize_lanes_f64(_ptr(stats64), float(count), L, g[0], b[0], g[1], b[1], g[2], b[2], g[3],
                                         b[3], _ptr(running_mean), _ptr(running_var), float(momentum), float(eps),
                                         _ptr(coeffs), c, _stream()), "byol_bn_finalize_lanes_f64")
    return coeffs


def bn_apply_f32(y2d, scale, shift, relu, T, resid=None, rscale=None, rshift=None, want_out32=False, want_planes=True,
                 want_copy=False, want_mask=False):
    """act(y*scale + shift (+ residual)) on fp32 [M, C] -> (out32, planes [M, T*C], bf16 copy, mask bits)."""
    _chk(y2d, F32, "y"); _chk(resid, F32, "resid")
    m, c = y2d.shape
    dev = y2d.device
    out32 = torch.empty((m, c), dtype=F32, device=dev) if want_out32 else None
    planes = torch.empty((m, T * c), dtype=BF16, device=dev) if want_planes else None
    copy = torch.empty((m, c), dtype=BF16, device=dev) if want_copy else None
    mask = torch.empty(m * c // 8, dtype=torch.uint8, device=dev) if want_mask else None
    check(lib.byol_bn_apply_f32(_ptr(y2d), _ptr(scale), _ptr(shift), _ptr(resid), _ptr(rscale), _ptr(rshift),
                                _ptr(out32), _ptr(planes), _ptr(copy), _ptr(mask), m, c, int(relu), T, _stream()),
          "byol_bn_apply_f32")
    return out32, planes, copy, mask


def maxpool_f32(x, k=3, s=2, p=1, want_idx=True):
    _chk(x, F32, "x")
    n, h, w, c = x.shape
    ho, wo = conv_out_size(h, k, s, p), conv_out_size(w, k, s, p)
    y = torch.empty((n, ho, wo, c), dtype=F32, device=x.device)
    idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device) if want_idx else None
    check(lib.byol_maxpool_f32(_ptr(x), _ptr(y), _ptr(idx), n, h, w, c, k, s, p, _stream()), "byol_maxpool_f32")
    return y, idx


def avgpool_f32(x):
    _chk(x, F32, "x")
    n, h, w, c = x.shape
    y = torch.empty((n, c), dtype=F32, device=x.device)
    check(lib.byol_avgpool_f32(_ptr(x), _ptr(y), n, h * w, c, _stream()), "byol_avgpool_f32")
    return y
