"""transcriptclean_b200 - B200-native implementation of TranscriptClean's per-read correction
path (correct_transcript(), /root/reference/TranscriptClean/TranscriptClean.py:406-460).

Host side (this package, Python): SAM -> columnar buffers, reference tables, text rendering,
the `transcriptclean` option surface.  Device side (csrc/, CUDA for sm_100a) behind the C ABI
in include/tc_b200.h, loaded with ctypes.  There is no CPU fallback: without the CUDA library
or without a GPU every entry point raises.
"""
__version__ = "0.1.0"
