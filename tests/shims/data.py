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


class _TextLoader:
    """dict batches of training/data.py's Text2ImageDataset (:455-603): ``image`` [B,3,R,R] in [0,1], tokenised captions
    ``input_ids`` [B,L], and the micro-conditioning fields as the default collate leaves them (``orig_size`` /
    ``crop_coords``: a pair of [B] tensors, ``aesthetic_score``: [B]); pre-encoded shards carry ``image_input_ids`` and
    ``encoder_hidden_states`` instead (:561-573)."""

    def __init__(self, num_batches, batch_size, resolution, seq_len, vocab, seed, pre_encoded=None):
        self.num_batches, self.batch_size, self.resolution = num_batches, batch_size, resolution
        self.seq_len, self.vocab, self.seed, self.pre_encoded = seq_len, vocab, seed, pre_encoded

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        B = self.batch_size
        for _ in range(self.num_batches):
            if self.pre_encoded is not None:
                n_tok, codebook, kv, width = self.pre_encoded
                yield {"image_input_ids": torch.randint(0, codebook, (B, n_tok), generator=g),
                       "encoder_hidden_states": torch.randn(B, kv, width, generator=g)}
                continue
            yield {"image": torch.rand(B, 3, self.resolution, self.resolution, generator=g),
                   "input_ids": torch.randint(1, self.vocab, (B, self.seq_len), generator=g),
                   "orig_size": [torch.full((B,), 256), torch.full((B,), 256)],
                   "crop_coords": [torch.zeros(B, dtype=torch.long), torch.zeros(B, dtype=torch.long)],
                   "aesthetic_score": torch.rand(B, generator=g) * 4 + 4}

    def __len__(self):
        return self.num_batches


class Text2ImageDataset:
    PRE_ENCODED = None  # tests set (n_tokens, codebook_size, kv_len, width) for the is_pre_encoded branch

    def __init__(self, train_shards_path_or_url=None, eval_shards_path_or_url=None, tokenizer=None, max_seq_length=16,
                 num_train_examples=0, per_gpu_batch_size=1, global_batch_size=1, num_workers=0, resolution=256,
                 center_crop=True, random_flip=False, shuffle_buffer_size=0, pin_memory=False, persistent_workers=False,
                 is_pre_encoded=False, **kwargs):
        nb = max(1, int(num_train_examples) // max(1, int(global_batch_size)))
        vocab = getattr(tokenizer, "vocab_size", 32) if tokenizer is not None else 32
        pre = self.PRE_ENCODED if is_pre_encoded else None
        self.train_dataloader = _TextLoader(nb, per_gpu_batch_size, resolution, max_seq_length, vocab, 300, pre)
        self.eval_dataloader = _TextLoader(2, per_gpu_batch_size, resolution, max_seq_length, vocab, 400, pre)
