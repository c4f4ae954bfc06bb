"""NTCD-TIMIT path bookkeeping with the reference's function names and return conventions
(packages/dataset/ntcd_timit.py:57-96,149-469): sorted recursive globs under the corpus layout, paths returned
relative to the data root."""
import os
import pathlib
from glob import glob

_SPLIT_DIR = {"train": "train/", "validation": "dev/", "test": "test/"}


def proc_video_audio_pair_dict(input_video_dir, dataset_type='train', labels='vad_labels', upsampled=False, dct=False,
                               norm_video=False):
    """Sorted (video .h5 paths, label .h5 paths), relative to input_video_dir (ntcd_timit.py:149-191)."""
    sub = _SPLIT_DIR.get(dataset_type, "")
    video_dir = input_video_dir + 'ntcd_timit/matlab_raw/' + sub
    audio_dir = input_video_dir + 'ntcd_timit/Clean/' + sub
    if upsampled:
        pattern = '**/*_upsampled.h5'
    elif dct:
        pattern = '**/*_dct.h5'
    elif norm_video:
        pattern = '**/*_normvideo.h5'
    else:
        pattern = '**/*[!dct][!upsampled][!normvideo].h5'  # character classes, as in the reference
    videos = sorted(glob(video_dir + pattern, recursive=True))
    labels_ = sorted(glob(audio_dir + '**/*_' + labels + '.h5', recursive=True))
    rel = lambda ps: [os.path.relpath(p, input_video_dir) for p in ps]
    return rel(videos), rel(labels_)


def proc_noisy_clean_pair_dict(input_speech_dir, dataset_type='train', dataset_size='complete', labels='vad_labels',
                               upsampled=False):
    """{noisy wav path: clean label .h5 path}, both relative to input_speech_dir, for every noise type / SNR of the
    configuration (ntcd_timit.py:384-469)."""
    clean_dir = input_speech_dir + 'ntcd_timit/Clean/' + _SPLIT_DIR.get(dataset_type, "")
    suffix = labels + ('_upsampled' if upsampled else '')
    label_files = sorted(glob(clean_dir + '**/*' + suffix + '.h5', recursive=True))
    shortpaths = []
    for f in label_files:
        short = str(pathlib.Path(*pathlib.Path(f).parts[-3:]))
        short = os.path.splitext(short)[0].replace('_' + suffix, '')
        shortpaths.append(short + '.wav')
    label_rel = [os.path.relpath(f, input_speech_dir) for f in label_files]
    noise_types = ['Babble', 'Cafe', 'Car', 'LR', 'Street', 'White']
    snrs = ['-5', '0', '5']
    if dataset_size == 'subset':
        noise_types, snrs = ['Babble'], ['-5']
    pairs = {}
    for noise in noise_types:
        for snr in snrs:
            base = os.path.join('ntcd_timit', 'Noisy', noise, snr)
            pairs.update(zip([os.path.join(base, s) for s in shortpaths], label_rel))
    return pairs


def speech_list(input_speech_dir, dataset_type='train'):
    """(raw clean wav paths, processed output paths) of the volunteers/lipspeakers tree (ntcd_timit.py:98-147)."""
    data_dir = input_speech_dir + 'ntcd_timit/Clean/volunteers/'
    files = sorted(glob(data_dir + '**/*.wav', recursive=True))
    rel = [os.path.relpath(p, input_speech_dir) for p in files]
    out = []
    for p in rel:
        parts = pathlib.Path(p).parts
        out.append(os.path.join('ntcd_timit', 'Clean', _SPLIT_DIR.get(dataset_type, ""), parts[-3], parts[-1]))
    return rel, out


def video_list(input_video_dir, dataset_type='train', upsampled=False):
    sub = _SPLIT_DIR.get(dataset_type, "")
    files = sorted(glob(input_video_dir + 'ntcd_timit/matlab_raw/' + sub + '**/*.mat', recursive=True))
    return [os.path.relpath(p, input_video_dir) for p in files]


# ---- raw-corpus listers used by the data-preparation scripts (ntcd_timit.py:57-96,193-381) ------------------------

_NOISE_TYPES = ['Babble', 'Cafe', 'Car', 'LR', 'Street', 'White']
_SNRS = ['-5', '0', '5']
_NOISY_ROOT = 'ntcd_timit/u/drspeech/data/TCDTIMIT/Noisy_TCDTIMIT'


def kaldi_list(input_video_dir, dataset_type='train', labels='vad_labels', upsampled=False):
    """Sorted (.ark paths, .scp paths) of the fMLLR features, relative to input_video_dir (ntcd_timit.py:57-96)."""
    data_dir = input_video_dir + 'ntcd_timit/kaldi_fMLLR/' + _SPLIT_DIR.get(dataset_type, "")
    rel = lambda ext: [os.path.relpath(p, input_video_dir) for p in sorted(glob(data_dir + '**/*' + ext, recursive=True))]
    return rel('.ark'), rel('.scp')


def _raw_utterances(input_speech_dir, dataset_type):
    """(speaker/straightcam/utt.wav, split/speaker/utt.wav) for every .mat file of the split, in sorted order."""
    data_dir = input_speech_dir + 'ntcd_timit/matlab_raw/' + _SPLIT_DIR.get(dataset_type, "")
    noisy_short, out_short = [], []
    for path in sorted(glob(data_dir + '**/*.mat', recursive=True)):
        stem = os.path.splitext(os.path.basename(path))[0]
        noisy_short.append(path.split('/')[-2] + '/straightcam/' + stem + '.wav')
        out_short.append(os.path.splitext(str(pathlib.Path(*pathlib.Path(path).parts[-3:])))[0] + '.wav')
    return noisy_short, out_short


def _conditions(dataset_size):
    if dataset_size == 'subset':
        return [('Babble', '-5')]
    return [(n, s) for n in _NOISE_TYPES for s in _SNRS]


def noisy_speech_dict(input_speech_dir, dataset_type='train', dataset_size='complete'):
    """{raw noisy wav -> processed noisy wav} over all noise types / SNRs (ntcd_timit.py:193-281)."""
    noisy_short, out_short = _raw_utterances(input_speech_dir, dataset_type)
    pairs = {}
    for noise, snr in _conditions(dataset_size):
        src_dir = os.path.join(_NOISY_ROOT, noise, snr, 'volunteers')
        dst_dir = os.path.join('ntcd_timit', 'Noisy', noise, snr)
        pairs.update({os.path.join(src_dir, a): os.path.join(dst_dir, b) for a, b in zip(noisy_short, out_short)})
    return pairs


def noisy_clean_pair_dict(input_speech_dir, dataset_type='train', dataset_size='complete'):
    """{raw noisy wav -> clean wav of the same utterance} (ntcd_timit.py:285-381)."""
    noisy_short, _ = _raw_utterances(input_speech_dir, dataset_type)
    clean_dir = 'ntcd_timit/Clean/' + _SPLIT_DIR.get(dataset_type, "")
    pairs = {}
    for noise, snr in _conditions(dataset_size):
        src_dir = os.path.join(_NOISY_ROOT, noise, snr, 'volunteers')
        for a in noisy_short:
            noisy = os.path.join(src_dir, a)
            pairs[noisy] = clean_dir + noisy.split('/')[-3] + '/' + os.path.basename(noisy)
    return pairs
