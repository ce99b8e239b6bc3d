"""Synthetic two-view loader standing in for datasets.loader.get_loader (/root/reference/main.py:417)."""
import torch


class _SyntheticTwoView(object):
    def __init__(self, batch_size, image_size, num_batches, num_classes, seed):
        self.batch_size, self.image_size, self.num_batches = batch_size, image_size, num_batches
        self.num_classes, self.seed = num_classes, seed

    def __len__(self):
        return self.num_batches

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        for _ in range(self.num_batches):
            shape = (self.batch_size, 3, self.image_size, self.image_size)
            yield (torch.rand(shape, generator=g), torch.rand(shape, generator=g),
                   torch.randint(0, self.num_classes, (self.batch_size,), generator=g))


class _Loader(object):
    def __init__(self, batch_size, image_size, steps, num_classes=1000, seed=1234):
        self.input_shape = [3, image_size, image_size]
        self.output_size = num_classes
        self.num_train_samples = batch_size * steps
        self.num_test_samples = batch_size
        self.num_valid_samples = 0
        self.train_loader = _SyntheticTwoView(batch_size, image_size, steps, num_classes, seed)
        self.test_loader = _SyntheticTwoView(batch_size, image_size, 1, num_classes, seed + 1)

    def set_all_epochs(self, epoch):
        pass


def get_loader(**kwargs):
    import os
    return _Loader(batch_size=kwargs.get("batch_size", 32),
                   image_size=kwargs.get("image_size_override", 224),
                   steps=int(os.environ.get("BYOL_SYNTH_STEPS", "2")))
