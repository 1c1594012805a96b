"""Synthetic stand-in for the reference's ``training/data.py`` (which needs webdataset / braceexpand): the same
``ClassificationDataset`` constructor keywords, seeded random ``(pixel_values [B,3,R,R] in [0,1], class_ids [B])`` batches,
loaders with a ``num_batches`` attribute (training/train_maskgit_imagenet.py:277-293,315)."""
import torch


class _Loader:
    def __init__(self, num_batches, batch_size, resolution, num_classes, seed):
        self.num_batches, self.batch_size, self.resolution = num_batches, batch_size, resolution
        self.num_classes, self.seed = num_classes, seed

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        for _ in range(self.num_batches):
            yield (torch.rand(self.batch_size, 3, self.resolution, self.resolution, generator=g),
                   torch.randint(0, self.num_classes, (self.batch_size,), generator=g))

    def __len__(self):
        return self.num_batches


class ClassificationDataset:
    NUM_CLASSES = 10

    def __init__(self, train_shards_path_or_url=None, eval_shards_path_or_url=None, num_train_examples=0,
                 per_gpu_batch_size=1, global_batch_size=1, num_workers=0, resolution=256, center_crop=True,
                 random_flip=False, shuffle_buffer_size=0, pin_memory=False, persistent_workers=False, **kwargs):
        nb = max(1, int(num_train_examples) // max(1, int(global_batch_size)))
        self._train_dataloader = _Loader(nb, per_gpu_batch_size, resolution, self.NUM_CLASSES, 100)
        self._eval_dataloader = _Loader(2, per_gpu_batch_size, resolution, self.NUM_CLASSES, 200)

    @property
    def train_dataloader(self):
        return self._train_dataloader

    @property
    def eval_dataloader(self):
        return self._eval_dataloader


class Text2ImageDataset(ClassificationDataset):
    pass
