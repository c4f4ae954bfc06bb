"""WSJ0 (CSR-1) path bookkeeping (packages/dataset/csr1_wjs0.py:19-57) -- used only by the legacy SE scripts."""
import os
from glob import glob

_DIRS = {"train": "si_tr_s", "validation": "si_dt_05", "test": "si_et_05"}


def speech_list(input_speech_dir, dataset_type='train'):
    data_dir = os.path.join(input_speech_dir, 'CSR-1-WSJ-0/WAV/wsj0', _DIRS.get(dataset_type, ""))
    files = sorted(glob(data_dir + '/**/*.wav', recursive=True))
    return [os.path.relpath(p, input_speech_dir) for p in files]
