"""Per-operator CPU references (torch fp32 on the host) used by the GPU parity tests.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function restates, in NCHW fp32 with stock torch
CPU operators, what one byol_b200 kernel computes in NHWC bf16, following the reference call sites:

* conv / linear      : /root/reference/main.py:190-205,237-239 (torchvision ResNet convs, nn.Linear)
* batch norm         : /root/reference/main.py:196,202 + torchvision BatchNorm2d (train mode, momentum 0.1)
* max / avg pooling  : torchvision resnet stem / tail reached from /root/reference/main.py:237
"""
import torch
import torch.nn.functional as F


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def nhwc_to_nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def nchw_to_nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def conv_fprop_ref(x_nhwc, w, stride, pad, bias=None, resid_nhwc=None, relu=False):
    """x_nhwc fp32 [N,H,W,Cin] (already bf16-rounded values), w fp32 [Cout,Cin,KH,KW] (bf16-rounded)."""
    y = F.conv2d(nhwc_to_nchw(x_nhwc), w, bias=bias, stride=stride, padding=pad)
    y = nchw_to_nhwc(y)
    if resid_nhwc is not None:
        y = y + resid_nhwc
    if relu:
        y = torch.relu(y)
    return y


def conv_dgrad_ref(dy_nhwc, w, in_hw, stride, pad):
    h, wd = in_hw
    cin = w.shape[1]
    n = dy_nhwc.shape[0]
    dx = torch.nn.grad.conv2d_input((n, cin, h, wd), w, nhwc_to_nchw(dy_nhwc), stride=stride, padding=pad)
    return nchw_to_nhwc(dx)


def conv_wgrad_ref(x_nhwc, dy_nhwc, w_shape, stride, pad):
    return torch.nn.grad.conv2d_weight(nhwc_to_nchw(x_nhwc), w_shape, nhwc_to_nchw(dy_nhwc), stride=stride,
                                       padding=pad)


def bn_train_ref(x2d, gamma, beta, eps=1e-5):
    """x2d fp32 [M, C]: returns (y, mean, invstd, biased var)."""
    mean = x2d.mean(0)
    var = x2d.var(0, unbiased=False)
    invstd = 1.0 / torch.sqrt(var + eps)
    y = (x2d - mean) * invstd * gamma + beta
    return y, mean, invstd, var


def bn_bwd_ref(dz, x2d, mean, invstd, gamma):
    """Standard batch-norm backward for dz = dL/d(bn output): returns (dx, dgamma, dbeta)."""
    m = x2d.shape[0]
    xhat = (x2d - mean) * invstd
    s1 = dz.sum(0)
    s2 = (dz * xhat).sum(0)
    dx = gamma * invstd * (dz - s1 / m - xhat * s2 / m)
    return dx, s2, s1


def maxpool_ref(x_nhwc, k=3, s=2, p=1):
    y, idx = F.max_pool2d(nhwc_to_nchw(x_nhwc), k, s, p, return_indices=True)
    return nchw_to_nhwc(y), idx


def maxpool_bwd_ref(x_nhwc, dy_nhwc, k=3, s=2, p=1):
    x = nhwc_to_nchw(x_nhwc).clone().requires_grad_(True)
    y = F.max_pool2d(x, k, s, p)
    y.backward(nhwc_to_nchw(dy_nhwc))
    return nchw_to_nhwc(x.grad)
