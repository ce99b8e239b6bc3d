"""CPU: the index algebra of maxpool_bwd_k3s2_kernel (csrc/elementwise.cu) restated with numpy.

For k = 3, s = 2, p = 1 and even sizes the 2 x 2 input block (2a .. 2a+1, 2b .. 2b+1) is touched by exactly the four
windows (a, b), (a, b+1), (a+1, b), (a+1, b+1), at the fixed window positions the kernel hard-codes.  The restatement
scatters dy through that table and must equal torch autograd's max_pool2d backward (first-maximum tie-breaking, the
argmax saved as window position kh*3 + kw)."""
import numpy as np
import torch
import torch.nn.functional as F

# (pixel in the block: 0 = (2a, 2b), 1 = (2a, 2b+1), 2 = (2a+1, 2b), 3 = (2a+1, 2b+1)) <- [(window, window position)]
TABLE = {0: [(0, 4)], 1: [(0, 5), (1, 3)], 2: [(0, 7), (2, 1)], 3: [(0, 8), (1, 6), (2, 2), (3, 0)]}


def test_block_table_matches_window_geometry():
    for q, srcs in TABLE.items():
        ih, iw = q >> 1, q & 1                       # pixel (2a + ih, 2b + iw) with a = b = 0 -> use a = b = 1 below
        a = b = 1
        py, px = 2 * a + ih, 2 * b + iw
        found = set()
        for oh in range(0, 4):
            for ow in range(0, 4):
                kh, kw = py - (2 * oh - 1), px - (2 * ow - 1)
                if 0 <= kh < 3 and 0 <= kw < 3:
                    found.add(((oh - a) * 2 + (ow - b), kh * 3 + kw))
        assert found == set(srcs), (q, found, srcs)


def test_block_scatter_equals_autograd():
    g = torch.Generator().manual_seed(0)
    n, c, h = 2, 5, 12
    x = torch.relu(torch.randn(n, c, h, h, generator=g)).requires_grad_(True)      # ties at zero like post-ReLU maps
    y, idx = F.max_pool2d(x, 3, 2, 1, return_indices=True)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    ho = h // 2
    # argmax as window position (what the forward kernel stores)
    ii = idx.numpy()
    oh, ow = np.meshgrid(np.arange(ho), np.arange(ho), indexing="ij")
    pos = (ii // h - (2 * oh - 1)) * 3 + (ii % h - (2 * ow - 1))
    dx = np.zeros((n, c, h, h), dtype=np.float32)
    dyn = dy.numpy()
    for a in range(ho):
        for b in range(ho):
            wins = [(a, b), (a, b + 1), (a + 1, b), (a + 1, b + 1)]
            for q, srcs in TABLE.items():
                for w, p in srcs:
                    wa, wb = wins[w]
                    if wa < ho and wb < ho:
                        hit = pos[:, :, wa, wb] == p
                        dx[:, :, 2 * a + (q >> 1), 2 * b + (q & 1)] += np.where(hit, dyn[:, :, wa, wb], 0.0)
    np.testing.assert_allclose(dx, x.grad.numpy(), rtol=0, atol=0)
