"""Kept for tools/make_golden.py: the reader now lives in the product package (avvad/h5min.py)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "audio-visual-vad_b200"))
from avvad.h5min import H5File, lzf_decompress, read_wav_int16  # noqa: F401,E402
