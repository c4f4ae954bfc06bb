"""WSJ0 (CSR-1) path bookkeeping (packages/dataset/csr1_wjs0.py:19-57) -- used only by the legacy SE scripts."""
import os
from glob import glob

_DIRS = {"train": "si_tr_s", "validation": "si_dt_05", "test": "si_et_05"}


def speech_list(input_speech_dir, dataset_type='train'):
    data_dir = os.path.join(input_speech_dir, 'CSR-1-WSJ-0/WAV/wsj0', _DIRS.get(dataset_type, ""))
    files = sorted(glob(data_dir + '/**/*.wav', recursive=True))
    return [os.path.relpath(p, input_speech_dir) for p in files]


def _pickle_path(base_dir, dataset_type, suffix):
    return base_dir + 'CSR-1-WSJ-0/' + _DIRS.get(dataset_type, "") + '_' + suffix + '.p'


def write_dataset(data, output_data_dir, dataset_type, suffix='unlabeled_frames'):
    """Pickle (protocol 4) `data` to <output_data_dir>CSR-1-WSJ-0/<split>_<suffix>.p (csr1_wjs0.py:59-95)."""
    import pickle
    os.makedirs(output_data_dir + 'CSR-1-WSJ-0/', exist_ok=True)
    with open(_pickle_path(output_data_dir, dataset_type, suffix), 'wb') as f:
        pickle.dump(data, f, protocol=4)
    print("data is stored in " + output_data_dir)


def read_dataset(data_dir, dataset_type, suffix='unlabeled_frames'):
    """Inverse of write_dataset (csr1_wjs0.py:98-129)."""
    import pickle
    with open(_pickle_path(data_dir, dataset_type, suffix), 'rb') as f:
        return pickle.load(f)
