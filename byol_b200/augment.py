"""On-device two-view augmentation — the torchvision branch of the reference's
``build_train_and_test_transforms`` (/root/reference/main.py:386-397) for a batch of decoded images that already live
in HBM: ``RandomResizedCrop(R)``, ``RandomHorizontalFlip(0.5)``, ``RandomApply([ColorJitter(0.8s, 0.8s, 0.8s, 0.2s)], 0.8)``,
``RandomGrayscale(0.2)``, ``GaussianBlur(kernel 0.1 R, p 0.5)``.

    aug = TwoViewAugment(image_size=224, seed=0)
    aug1, aug2 = aug(images)            # images: fp32 CUDA [N, 3, H, W] in [0, 1] -> two fp32 [N, 3, 224, 224]

Every call draws fresh parameters from a counter-based RNG (seed, call counter, sample, view), so a run is
reproducible and the two views of a sample are independent.  ``apply(images, params)`` runs the pipeline on explicit
parameter records (tests/test_gpu_augment.py checks every stage against torchvision on identical parameters).
"""
import torch

from . import ops
from ._lib import lib, check

RECORD = lib.byol_augment_record_floats()


class TwoViewAugment(object):
    def __init__(self, image_size=224, color_jitter_strength=1.0, seed=0, p_flip=0.5, p_jitter=0.8, p_gray=0.2,
                 p_blur=0.5, blur=True):
        self.R = int(image_size)
        self.strength = float(color_jitter_strength)
        self.seed = int(seed)
        self.p = (float(p_flip), float(p_jitter), float(p_gray), float(p_blur))
        k = int(0.1 * self.R) if blur else 0      # main.py:396 kernel_size = int(0.1 * image_size); made odd
        self.ksize = (k | 1) if k > 0 else 0
        self.calls = 0

    def sample_params(self, n, hs, ws, device):
        params = torch.empty((2, n, RECORD), dtype=torch.float32, device=device)
        check(lib.byol_augment_params(params.data_ptr(), n, hs, ws, self.seed, self.calls, self.strength, self.p[0],
                                      self.p[1], self.p[2], self.p[3], ops._stream()), "byol_augment_params")
        self.calls += 1
        return params

    def apply(self, images, params):
        if not images.is_cuda or images.dtype != torch.float32 or images.dim() != 4 or images.shape[1] != 3:
            raise ValueError("TwoViewAugment: expected a CUDA fp32 [N, 3, H, W] batch in [0, 1] (no CPU path)")
        images = images.contiguous()
        n, _, hs, ws = images.shape
        if tuple(params.shape) != (2, n, RECORD) or params.dtype != torch.float32 or not params.is_contiguous():
            raise ValueError("TwoViewAugment: params must be a contiguous fp32 [2, N, %d] tensor" % RECORD)
        out = torch.empty((2, n, 3, self.R, self.R), dtype=torch.float32, device=images.device)
        tmp = torch.empty_like(out) if self.ksize else None
        gray = torch.empty(2 * n, dtype=torch.float64, device=images.device)
        check(lib.byol_augment_apply(images.data_ptr(), params.data_ptr(), out.data_ptr(),
                                     0 if tmp is None else tmp.data_ptr(), gray.data_ptr(), n, hs, ws, self.R, self.ksize,
                                     ops._stream()), "byol_augment_apply", kernels=4 if self.ksize else 2)
        return out[0], out[1]

    def __call__(self, images):
        n, _, hs, ws = images.shape
        return self.apply(images, self.sample_params(n, hs, ws, images.device))
