"""audio_zen/constant.py"""
import math

import numpy as np

PI = math.pi
EPSILON = np.finfo(np.float32).eps  # constant.py:9
MAX_INT16 = np.iinfo(np.int16).max
