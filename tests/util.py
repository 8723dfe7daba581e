import hashlib

import numpy as np


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def kps_xyr(stage):
    return np.ascontiguousarray(stage, np.float32)
