import enum
import random

import numpy as np
import torch


class DistributedType(str, enum.Enum):
    NO = "NO"
    MULTI_GPU = "MULTI_GPU"
    DEEPSPEED = "DEEPSPEED"


def set_seed(seed, device_specific=False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
