from open_muse_b200.sampling import *  # noqa: F401,F403
from open_muse_b200.sampling import cosine_schedule, get_mask_chedule, gumbel_sample, mask_by_random_topk, top_k  # noqa: F401
