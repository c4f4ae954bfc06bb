"""The unchanged training scripts' input pipeline: datasets built over a processed-corpus tree, iterated through
torch.utils.data.DataLoader with FORKED workers (scripts/train_AV_net.py:50,131-146,256: num_workers=16, default
context, collate_fn=collate_many2many_AV, created before and iterated after the model is on the GPU).

The STFT is a CUDA kernel with no CPU implementation, and a forked child of a CUDA-initialised process cannot use CUDA:
workers return deferred spectrograms, the parent runs ONE batched device front end when the batch comes off the queue
(packages/processing/deferred.py).  The CPU variant of the test injects the oracle as that front end (no GPU here); the
GPU variant runs the real kernel with CUDA initialised in the parent before the workers fork."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from avvad import synth
from oracle import frontend as ofe
from packages.processing import deferred

STFT = dict(fs=16000, wlen_sec=64e-3, win='hann', hop_percent=0.25, center=False, pad_mode='reflect', pad_at_end=True,
            eps=1e-8)


def _oracle_materialiser(waves, n_samples, n_frames, t_max, eps):
    out = torch.zeros(len(n_samples), t_max, 513)
    for b, (n, t) in enumerate(zip(n_samples, n_frames)):
        lp = ofe.logpower(ofe.peak_normalise(waves[b, :n].numpy()), eps=eps, dtype=np.float32)   # (513, T)
        out[b, :t] = torch.from_numpy(np.ascontiguousarray(lp[:, :t].T))
    return out


@pytest.fixture()
def corpus(tmp_path):
    root = str(tmp_path) + "/"
    return root, synth.write_synthetic_corpus(root)


def _expected(made, key):
    wav, video, lab = made[key]
    x = wav.astype(np.float32) / 32768.0
    lp = ofe.logpower(ofe.peak_normalise(x), eps=1e-8, dtype=np.float32)
    n = min(lp.shape[1], video.shape[-1], lab.shape[-1])
    return lp[:, :n], video[..., :n], lab[..., :n], n


def _run_av(root, made, workers, tol):
    from packages.data_handling import AudioVisualSequenceLabeledFrames
    from packages.utils import collate_many2many_AV
    ds = AudioVisualSequenceLabeledFrames(input_video_dir=root, dataset_type='train', dataset_size='subset',
                                          labels='vad_labels', upsampled=True, **STFT)
    assert len(ds) == 3
    loader = DataLoader(ds, batch_size=3, shuffle=False, num_workers=workers, pin_memory=torch.cuda.is_available(),
                        drop_last=False, timeout=0, worker_init_fn=None, collate_fn=collate_many2many_AV)
    batches = list(loader)
    assert len(batches) == 1
    lengths, x, v, y = batches[0]
    assert isinstance(x, torch.Tensor) and not x.is_cuda and x.dtype == torch.float32
    keys = [("train", "01M", "sa1"), ("train", "01M", "sa2"), ("train", "02F", "si1")]
    exp = [_expected(made, k) for k in keys]
    T = max(e[3] for e in exp)
    assert lengths.dtype == torch.int64 and lengths.tolist() == [e[3] for e in exp]
    assert x.shape == (3, T, 513) and v.shape == (3, T, 67, 67) and y.shape == (3, T, 1)
    for b, (lp, vid, lab, n) in enumerate(exp):
        assert np.abs(x[b, :n].numpy() - lp.T).max() <= tol
        assert torch.all(x[b, n:] == 0)                       # collate pads with zeros BEFORE standardisation
        assert np.array_equal(v[b, :n].numpy(), np.moveaxis(vid, -1, 0)) and torch.all(v[b, n:] == 0)
        assert np.array_equal(y[b, :n, 0].numpy(), lab[0]) and torch.all(y[b, n:] == 0)
    return lengths, x, v, y


def test_av_dataset_through_forked_workers_cpu(corpus):
    root, made = corpus
    deferred.set_materialiser(_oracle_materialiser)
    try:
        a = _run_av(root, made, 2, 1e-5)
        b = _run_av(root, made, 0, 1e-5)
        for s, t in zip(a, b):
            assert torch.equal(s, t)
    finally:
        deferred.set_materialiser(None)


def test_audio_and_wav_datasets_with_their_collates(corpus):
    """The audio-only dataset and both *Wav* datasets (items end in (time_length, tf_length), which is what
    collate_many2many_{audio,AV}_waveform index) through forked workers."""
    from packages.data_handling import (AudioVisualSequenceWavLabeledFrames, NoisyWavWholeSequenceSpectrogramLabeledFrames,
                                        NoisyWavWholeSequenceWavLabeledFrames)
    from packages.utils import collate_many2many_audio, collate_many2many_audio_waveform, collate_many2many_AV_waveform
    root, made = corpus
    kw = dict(input_video_dir=root, dataset_type='train', dataset_size='subset', labels='vad_labels', **STFT)
    deferred.set_materialiser(_oracle_materialiser)
    try:
        ds = NoisyWavWholeSequenceSpectrogramLabeledFrames(upsampled=True, **kw)
        lengths, x, y = next(iter(DataLoader(ds, batch_size=3, num_workers=2, collate_fn=collate_many2many_audio)))
        assert x.shape == (3, int(lengths.max()), 513) and y.shape[:2] == x.shape[:2]
        item = ds[0]                                            # main process: a real tensor, like the reference
        assert isinstance(item[0], torch.Tensor) and item[0].shape == (513, item[2])
    finally:
        deferred.set_materialiser(None)
    # the waveform datasets never touch the front end; their label files are the non-upsampled naming in the reference
    import os
    import shutil
    for split, spk, utt in made:
        d = os.path.join(root, "ntcd_timit", "Clean", split, spk)
        shutil.copyfile(os.path.join(d, f"{utt}_vad_labels_upsampled.h5"), os.path.join(d, f"{utt}_vad_labels.h5"))
    ds = AudioVisualSequenceWavLabeledFrames(**kw)
    wav0, video0, lab0 = made[("train", "01M", "sa1")]
    data, video, label, time_length, tf_length = ds[0]
    assert time_length == len(wav0) == data.shape[-1] and tf_length == video0.shape[-1] == video.shape[-1]
    assert label.shape[-1] == lab0.shape[-1]                    # untrimmed, as in the reference
    assert float(data.abs().max()) == 1.0
    lengths, a, v, y = next(iter(DataLoader(ds, batch_size=3, num_workers=2, collate_fn=collate_many2many_AV_waveform)))
    assert a.shape == (3, 26500) and v.shape == (3, int(lengths.max()), 67, 67)
    assert lengths.tolist() == [made[k][1].shape[-1] for k in [("train", "01M", "sa1"), ("train", "01M", "sa2"), ("train", "02F", "si1")]]
    ds = NoisyWavWholeSequenceWavLabeledFrames(**kw)
    lengths, a, y = next(iter(DataLoader(ds, batch_size=3, num_workers=2, collate_fn=collate_many2many_audio_waveform)))
    assert a.shape == (3, 26500) and y.shape == (3, int(lengths.max()), 1)


@pytest.mark.gpu
def test_av_dataset_through_forked_workers_after_cuda_init(corpus):
    """Exactly the order of scripts/train_AV_net.py: CUDA initialised in the parent, then fork workers, then iterate."""
    root, made = corpus
    torch.zeros(1, device="cuda")                               # the model is on the GPU before the loaders iterate
    assert torch.cuda.is_initialized()
    a = _run_av(root, made, 2, 2e-3)
    b = _run_av(root, made, 0, 2e-3)
    for s, t in zip(a, b):
        assert torch.equal(s, t)                                # batched in the parent == in place, bit for bit
    # and the front end honours the device of a CUDA input / the current device (packages/processing/stft.py)
    from packages.processing.stft import stft_pytorch
    wav = torch.tensor(made[("train", "01M", "sa1")][0].astype(np.float32) / 32768.0)
    s_cpu = stft_pytorch(wav, fs=16000, wlen_sec=64e-3, hop_percent=0.25, center=False)
    s_dev = stft_pytorch(wav.cuda(), fs=16000, wlen_sec=64e-3, hop_percent=0.25, center=False)
    assert not s_cpu.is_cuda and s_dev.is_cuda and torch.equal(s_cpu, s_dev.cpu())
