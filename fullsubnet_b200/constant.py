"""Numeric constants the hot path shares with the reference (values of audio_zen/constant.py:6-10).

Written as literals so that the library (fsn_common.cuh) and the Python host agree bit for bit."""
PI = 3.141592653589793
# float32 machine epsilon, 2**-23: denominator guard of build_complex_ideal_ratio_mask (mask.py:22) and of the
# cumulative / improved_fullsubnet norms
EPSILON = 1.1920928955078125e-07
# full-scale of 16-bit PCM (int16 scaling of the inference host loop, base_inferencer.py:181-182)
MAX_INT16 = 32767
