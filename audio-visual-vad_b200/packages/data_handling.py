"""Dataset classes the three train scripts build (same names, constructor keywords and item tuples as
packages/data_handling.py:192-566).

The per-utterance front end (peak-normalise -> STFT -> |.|^2 -> log, data_handling.py:441-457) is a CUDA kernel here.
In the main process ``__getitem__`` returns the computed (513, T) tensor like the reference.  Inside a DataLoader worker
-- the unchanged scripts fork 16 of them after CUDA is initialised (train_AV_net.py:50,141-146), where CUDA cannot be
used -- it returns a packages.processing.deferred.DeferredLogPower instead; collate_many2many_{AV,audio} turn those into
one batched device call that runs in the parent when the batch arrives (see deferred.py).  The legacy HDF5*/VideoFrames
datasets (data_handling.py:19-189, unused by the VAD scripts) are kept as thin host-side classes at the end."""
import math
import os

import numpy as np
import torch
from torch.utils.data import Dataset

from packages._io import load_wav, read_h5
from packages.dataset.ntcd_timit import proc_noisy_clean_pair_dict, proc_video_audio_pair_dict
from packages.processing.deferred import DeferredLogPower, in_worker
from packages.processing.stft import stft_pytorch, stft_frame_count


def _log_power(wave, ds):
    """data_handling.py:441-457: x / max|x| -> STFT -> re^2 + im^2 -> log(. + eps), (513, T); deferred inside a
    DataLoader worker."""
    nfft, hop = int(ds.wlen_sec * ds.fs), int(ds.hop_percent * int(ds.wlen_sec * ds.fs))
    if nfft != 1024 or hop != 256 or ds.center or ds.win != 'hann':
        raise NotImplementedError("libavvad front end supports nfft=1024, hop=256, win='hann', center=False")
    T = stft_frame_count(wave.shape[-1], ds.fs, ds.wlen_sec, ds.hop_percent, ds.pad_at_end)
    item = DeferredLogPower(wave.to(torch.float32).contiguous(), T, ds.eps)
    return item if in_worker() else item.materialise()


class _StftConfig:
    def _set_stft(self, fs, wlen_sec, win, hop_percent, center, pad_mode, pad_at_end, eps):
        self.fs, self.wlen_sec, self.win, self.hop_percent = fs, wlen_sec, win, hop_percent
        self.center, self.pad_mode, self.pad_at_end, self.eps = center, pad_mode, pad_at_end, eps


class WavWholeSequenceSpectrogramLabeledFrames(Dataset):
    """Video-only items (video (67,67,T), label (y_dim,T), length) (data_handling.py:192-229)."""

    def __init__(self, input_video_dir, dataset_type, labels='vad_labels', upsampled=False, dct=False, norm_video=False):
        self.dataset_type, self.input_video_dir, self.labels = dataset_type, input_video_dir, labels
        self.video_file_paths, self.audio_file_paths = proc_video_audio_pair_dict(
            input_video_dir=input_video_dir, dataset_type=dataset_type, labels=labels, upsampled=upsampled, dct=dct,
            norm_video=norm_video)
        self.dataset_len = len(self.video_file_paths)

    def __getitem__(self, i):
        data = read_h5(self.input_video_dir + self.video_file_paths[i], "X")
        label = read_h5(self.input_video_dir + self.audio_file_paths[i], "Y")
        return torch.Tensor(data), torch.Tensor(label), data.shape[-1]

    def __len__(self):
        return self.dataset_len


class _NoisyBase(Dataset, _StftConfig):
    def __init__(self, input_video_dir, dataset_type, dataset_size, labels='vad_labels', upsampled=False, fs=16000,
                 wlen_sec=64e-3, win='hann', hop_percent=0.25, center=True, pad_mode='reflect', pad_at_end=True,
                 eps=1e-8):
        self.input_video_dir, self.dataset_type, self.dataset_size = input_video_dir, dataset_type, dataset_size
        self.labels, self.upsampled = labels, upsampled
        self._set_stft(fs, wlen_sec, win, hop_percent, center, pad_mode, pad_at_end, eps)
        pairs = proc_noisy_clean_pair_dict(input_speech_dir=input_video_dir, dataset_type=dataset_type,
                                           dataset_size=dataset_size, labels=labels, upsampled=upsampled)
        self.noisy_clean_pair_paths = list(pairs.items())
        self.dataset_len = len(self.noisy_clean_pair_paths)

    def __len__(self):
        return self.dataset_len

    def _wave(self, i):
        noisy, clean = self.noisy_clean_pair_paths[i]
        wave, _ = load_wav(self.input_video_dir + noisy)
        return wave[0], clean

    def _video_path(self, clean):
        p = clean.replace('Clean', 'matlab_raw').replace('_' + self.labels, '')
        p = os.path.splitext(p)[0] + ('.h5' if self.upsampled else '_normvideo.h5')
        return self.input_video_dir + p


class NoisyWavWholeSequenceSpectrogramLabeledFrames(_NoisyBase):
    """Audio-only items (log-power (513,T), label (y_dim,T), length) (data_handling.py:231-324)."""

    def __getitem__(self, i):
        wave, clean = self._wave(i)
        data = _log_power(wave, self)
        label = torch.Tensor(read_h5(self.input_video_dir + clean, "Y"))
        n = min(data.shape[-1], label.shape[-1])  # the reference's longer-label branch overwrites data: not reproduced
        return data[..., :n], label[..., :n] if label.shape[-1] > n else label, n


class AudioVisualSequenceLabeledFrames(_NoisyBase):
    """AV items (log-power (513,T), video (67,67,T), label (y_dim,T), length), all trimmed to the common length
    (data_handling.py:387-495)."""

    def __getitem__(self, i):
        wave, clean = self._wave(i)
        spec = _log_power(wave, self)
        video = torch.Tensor(read_h5(self._video_path(clean), "X"))
        label = torch.Tensor(read_h5(self.input_video_dir + clean, "Y"))
        n = min(spec.shape[-1], video.shape[-1], label.shape[-1])
        return spec[..., :n], video[..., :n], label[..., :n], n


class AudioVisualSequenceWavLabeledFrames(_NoisyBase):
    """Waveform variant (data_handling.py:497-566): (peak-normalised wave (N,), video (67,67,T), label (y_dim,T),
    time_length, tf_length) -- untrimmed, video always from '<utt>_upsampled.h5', tf_length = video frames; pair with
    collate_many2many_AV_waveform (which reads lengths = item[-1], time_lengths = item[-2])."""

    def __getitem__(self, i):
        wave, clean = self._wave(i)
        data = wave / torch.max(torch.abs(wave))
        p = clean.replace('Clean', 'matlab_raw').replace('_' + self.labels, '')
        video = torch.Tensor(read_h5(self.input_video_dir + os.path.splitext(p)[0] + '_upsampled.h5', "X"))
        label = torch.Tensor(read_h5(self.input_video_dir + clean, "Y"))
        return data, video, label, data.shape[-1], video.shape[-1]


class NoisyWavWholeSequenceWavLabeledFrames(_NoisyBase):
    """Waveform variant of the audio-only dataset (data_handling.py:326-385): (wave, label, time_length, tf_length)."""

    def __getitem__(self, i):
        wave, clean = self._wave(i)
        data = wave / torch.max(torch.abs(wave))
        label = torch.Tensor(read_h5(self.input_video_dir + clean, "Y"))
        return data, label, data.shape[-1], label.shape[-1]


# ---- legacy datasets (data_handling.py:19-189): frame / sequence views of one big "X_<split>" / "Y_<split>" pair ---------

class _H5Pair(Dataset):
    """Shared part of the three HDF5* datasets: the file is opened lazily in the worker (first __getitem__), the two
    arrays are (features, N) and (labels, N) with the frame index last."""

    def __init__(self, output_h5_dir, dataset_type, rdcc_nbytes, rdcc_nslots):
        self.output_h5_dir, self.dataset_type = output_h5_dir, dataset_type
        self.rdcc_nbytes, self.rdcc_nslots = rdcc_nbytes, rdcc_nslots  # chunk-cache hints of h5py; unused by read_h5
        self.n_frames = read_h5(output_h5_dir, "X_" + dataset_type).shape[-1]

    def open_hdf5(self):
        self.data = read_h5(self.output_h5_dir, "X_" + self.dataset_type)
        self.labels = read_h5(self.output_h5_dir, "Y_" + self.dataset_type)
        self.f = True

    def _ensure_open(self):
        if not hasattr(self, "f"):
            self.open_hdf5()


class HDF5SpectrogramLabeledFrames(_H5Pair):
    """Item i = (spectrogram column i, label column i) (data_handling.py:51-80)."""

    def __init__(self, output_h5_dir, dataset_type, rdcc_nbytes, rdcc_nslots):
        super().__init__(output_h5_dir, dataset_type, rdcc_nbytes, rdcc_nslots)
        self.dataset_len = self.n_frames

    def __getitem__(self, i):
        self._ensure_open()
        return self.data[:, i], self.labels[:, i]

    def __len__(self):
        return self.dataset_len


class HDF5SequenceSpectrogramLabeledFrames(_H5Pair):
    """Item i = (the up-to-seq_length frames ending at i, the label of frame i, length) (data_handling.py:82-138)."""

    def __init__(self, output_h5_dir, dataset_type, rdcc_nbytes, rdcc_nslots, seq_length):
        super().__init__(output_h5_dir, dataset_type, rdcc_nbytes, rdcc_nslots)
        self.seq_length, self.dataset_len = seq_length, self.n_frames

    def __getitem__(self, i):
        self._ensure_open()
        first = 0 if i < self.seq_length else i + 1 - self.seq_length
        data = np.array(self.data[..., first:i + 1])
        labels = np.array(self.labels[..., i:i + 1])
        return torch.Tensor(data), torch.Tensor(labels), data.shape[-1]

    def __len__(self):
        return self.dataset_len


class HDF5WholeSequenceSpectrogramLabeledFrames(_H5Pair):
    """Item i = the i-th non-overlapping block of seq_length frames with all its labels (data_handling.py:140-189)."""

    def __init__(self, output_h5_dir, dataset_type, rdcc_nbytes, rdcc_nslots, seq_length):
        super().__init__(output_h5_dir, dataset_type, rdcc_nbytes, rdcc_nslots)
        self.seq_length, self.dataset_len = seq_length, math.ceil(self.n_frames / seq_length)

    def __getitem__(self, i):
        self._ensure_open()
        lo = i * self.seq_length
        data = np.array(self.data[..., lo:lo + self.seq_length])
        labels = np.array(self.labels[..., lo:lo + self.seq_length])
        return torch.Tensor(data), torch.Tensor(labels), data.shape[-1]

    def __len__(self):
        return self.dataset_len


class VideoFrames(Dataset):
    """Random seq_length-frame window of one speaker's DCT frames, decoded to 3x67x67, with the label of the frame
    after the window (data_handling.py:19-49; paths relative to the working directory as in the reference)."""

    def __init__(self, data, seq_length):
        self.data, self.seq_length = data, seq_length
        self.index = np.arange(len(self.data))

    def __getitem__(self, i):
        from packages.processing.video import _idct_unnormalised
        from avvad.h5min import H5File
        mat_file_path = os.path.join("data/complete/matlab_raw", self.data[i]) + ".mat"
        try:
            import h5py
            with h5py.File(mat_file_path, "r") as f:
                frames = [np.array(v) for v in f.values()][-1]
        except ImportError:
            f = H5File(mat_file_path)
            frames = f[list(f.keys())[-1]]
        peak = frames.flatten().max() + 1e-8
        video = torch.empty(frames.shape[0], 3, 67, 67)
        for k in range(frames.shape[0]):
            img = _idct_unnormalised(_idct_unnormalised(frames[k].reshape(67, 67)).T).T
            video[k] = torch.from_numpy(np.ascontiguousarray(np.rot90(img / peak * 255.0, 3))).float().expand(3, 67, 67)
        start = np.random.randint(frames.shape[0] - self.seq_length)
        labels = np.load("{}{}.npy".format("data/complete/labels/", self.data[i]))[start + self.seq_length]
        return video[start:start + self.seq_length], labels

    def __len__(self):
        return len(self.data)
