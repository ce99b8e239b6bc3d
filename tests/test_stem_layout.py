"""CPU restatement of the index algebra behind byol_b200/csrc/conv_stem.cu (no GPU, no extension call): the padded
NHWC4 image + the [7][4][64][8] weight layout + "row m starts 16 bytes (= 8 bf16 = 2 pixels) after row m-1" must
reproduce the 7x7 / stride 2 / pad 3 convolution, and the (input row v) -> (output row a - j, kernel row par + 2j)
pairing of the weight-gradient kernel must reproduce its gradient.  Reference semantics: torchvision ResNet conv1 as
reached from /root/reference/main.py:237."""
import torch
import torch.nn.functional as F

WP = 264   # byol_stem4_row_pixels()


def to_stem4(x):
    """fp32 NCHW [N, C<=4, H, W] -> [N, H+6, WP, 4] zero padded (what nchw_to_stem4_kernel writes)."""
    n, c, h, w = x.shape
    out = torch.zeros(n, h + 6, WP, 4)
    out[:, 3:3 + h, 3:3 + w, :c] = x.permute(0, 2, 3, 1)
    return out


def prep_stem4(wt):
    """fp32 [64, Cin, 7, 7] -> [kh 7][k-chunk 4][cout 64][8]; element e of chunk kc = (kw = 2*kc + e//4, c = e%4)."""
    cout, cin = wt.shape[:2]
    ws = torch.zeros(7, 4, cout, 8)
    for kc in range(4):
        for e in range(8):
            kw, c = 2 * kc + e // 4, e % 4
            if kw < 7 and c < cin:
                ws[:, kc, :, e] = wt[:, c, :, kw].t()
    return ws


def test_stem4_forward_index_algebra():
    g = torch.Generator().manual_seed(5)
    n, h, w = 2, 12, 20
    x = torch.randn(n, 3, h, w, generator=g)
    wt = torch.randn(64, 3, 7, 7, generator=g)
    xs, ws = to_stem4(x), prep_stem4(wt)
    ho, wo = h // 2, w // 2
    y = torch.zeros(n, ho, wo, 64)
    for oh in range(ho):
        for kh in range(7):
            row = xs[:, 2 * oh + kh].reshape(n, -1)                       # [n, WP*4] elements of padded row v
            for kc in range(4):                                           # 16-byte chunk = 8 elements
                # A[m, kc*8 + e] = row[8*m + 8*kc + e]  (LBO = 16 B between chunks, 16 B between rows)
                idx = (8 * torch.arange(wo)[:, None] + 8 * kc + torch.arange(8)[None, :])
                a = row[:, idx]                                           # [n, wo, 8]
                y[:, oh] += torch.einsum("nme,oe->nmo", a, ws[kh, kc])
    ref = F.conv2d(x, wt, stride=2, padding=3).permute(0, 2, 3, 1)
    assert torch.allclose(y, ref, atol=1e-4, rtol=1e-4)


def test_stem4_wgrad_row_pairing():
    g = torch.Generator().manual_seed(6)
    n, h, w = 2, 8, 12
    x = torch.randn(n, 3, h, w, generator=g, requires_grad=False)
    wt = torch.randn(64, 3, 7, 7, generator=g, requires_grad=True)
    dy = torch.randn(n, h // 2, w // 2, 64, generator=g)
    F.conv2d(x, wt, stride=2, padding=3).backward(dy.permute(0, 3, 1, 2))
    xs = to_stem4(x)
    ho, wo, hp = h // 2, w // 2, h + 6
    dw = torch.zeros(64, 3, 7, 7)
    for v in range(hp):                       # one unit = one padded input row
        a, par = v >> 1, v & 1
        row = xs[:, v].reshape(n, -1)
        idx = 8 * torch.arange(wo)[:, None] + torch.arange(32)[None, :]   # B[k = pixel, n = kwslot*4 + c]
        b = row[:, idx]                                                    # [n, wo, 32]
        for j in range(4):                    # chain 0: j = 0, 1 (rows a, a-1); chain 1: j = 2, 3 (rows a-2, a-3)
            oh, kh = a - j, par + 2 * j
            if kh > 6 or oh < 0 or oh >= ho:  # kh = 7 is the ignored accumulator half, rows outside are zero tiles
                continue
            d = torch.einsum("nko,nkq->oq", dy[:, oh], b)                  # [64, 32]
            d = d.view(64, 8, 4)                                            # (kwslot, c)
            dw[:, :, kh, :] += d[:, :7, :3].permute(0, 2, 1)
    assert torch.allclose(dw, wt.grad, atol=1e-3, rtol=1e-4)
