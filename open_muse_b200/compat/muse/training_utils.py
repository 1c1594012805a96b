"""``muse.training_utils`` of the drop-in package: the logging diagnostics training/train_muse.py computes from the logits
the model returns (train_muse.py:1309-1383 -> muse/training_utils.py:299-455) and the seeding helpers (:27-60).

Host-side torch analytics, not part of the hot path: entropies / cross-entropy / token probabilities grouped by the share
of masked tokens per image, ten buckets ``(k/10, (k+1)/10]``.  Written against the reference's observable behaviour and
compared with it value for value in tests/test_train_script_cpu.py."""
import os
import random

import numpy as np
import torch
import torch.nn.functional as F


def enable_full_determinism(seed: int):
    """reference :27-44"""
    set_seed(seed)
    os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
    os.environ["CUBLAS_WORKSPACE_CONFIG"] = ":16:8"
    torch.use_deterministic_algorithms(True)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def set_seed(seed: int):
    """reference :47-58"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def input_ids_to_masked_buckets(input_ids, mask_id, total_buckets=10):
    """bucket k <=> k/10 < masked share <= (k+1)/10 (an image without masked tokens lands in bucket 0, as upstream)"""
    assert total_buckets == 10
    share = (input_ids == mask_id).sum(-1) / input_ids.shape[-1]
    edges = torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9], dtype=share.dtype, device=share.device)
    return torch.bucketize(share, edges, right=False)


def average_by_buckets(values, masked_buckets, total_buckets):
    """mean of ``values`` per bucket, 0 for empty buckets.  Upstream scatters with the per-image bucket index even when
    ``values`` is longer (the cross-entropy diagnostic passes per-token values): only the first ``len(index)`` values
    take part there, which this keeps."""
    values = values.reshape(-1)[: masked_buckets.numel()]
    total = torch.zeros(total_buckets, device=values.device, dtype=values.dtype).index_add_(0, masked_buckets, values)
    count = torch.bincount(masked_buckets, minlength=total_buckets).clamp(min=1)
    return total / count


def pixel_entropy_per_percent_masked_bucket(logits, input_ids, mask_id):
    """mean per-token predictive entropy over the masked tokens of an image, averaged per bucket"""
    masked = input_ids == mask_id
    logp = F.log_softmax(logits, dim=-1)
    ent = -(logp.exp() * logp).sum(-1) * masked
    per_image = ent.sum(-1) / masked.sum(-1)
    return average_by_buckets(per_image, input_ids_to_masked_buckets(input_ids, mask_id), 10)


def image_entropy_per_percent_masked_bucket(logits, input_ids, mask_id):
    """entropy of the token distribution averaged over an image's masked positions, averaged per bucket"""
    masked = input_ids == mask_id
    p = (F.softmax(logits, dim=-1) * masked[..., None]).sum(-2) / masked.sum(-1, keepdim=True)
    per_image = -(p * p.log()).sum(-1)
    return average_by_buckets(per_image, input_ids_to_masked_buckets(input_ids, mask_id), 10)


def cross_entropy_per_percent_masked_bucket(logits, labels, input_ids, mask_id, output_size, label_smoothing):
    ce = F.cross_entropy(logits.reshape(-1, output_size), labels.reshape(-1), ignore_index=-100,
                         label_smoothing=label_smoothing, reduction="none")
    return average_by_buckets(ce, input_ids_to_masked_buckets(input_ids, mask_id), 10)


def token_probability_distributions_per_percent_masked_bucket(logits, input_ids, mask_id):
    """for each non-empty bucket the predicted distribution of ONE masked token -- upstream indexes the batch with the
    bucket's own number (``masked_buckets[masked_buckets == k][0] == k``), i.e. takes image k, first masked position"""
    import pandas as pd

    probs = F.softmax(logits, dim=-1)
    buckets = input_ids_to_masked_buckets(input_ids, mask_id)
    rows = []
    for k in range(10):
        if not bool((buckets == k).any()):
            continue
        dist = probs[k][input_ids[k] == mask_id][0].cpu().numpy()
        rows.extend({"bucket": k, "masked_pixel_prob": v} for v in dist)
    return pd.DataFrame(rows)
