"""Collate functions and helpers with the reference's names and return conventions
(packages/utils.py:5-6,42-185): zero-pad every item of a batch to the longest sequence, move time
to axis 1, return ``(lengths LongTensor, padded..., target)``."""
import torch


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def my_collate(batch):
    """Legacy many-to-one collate (packages/utils.py:9-40, imported by scripts/train_video_net.py:19): items
    ``(video (W,H,C,T_i), label, T_i)`` -> ``(lengths, video (B,T,C,H,W) zero-padded and squeezed, target (B,1))``."""
    lengths = [item[2] for item in batch]
    T = max(lengths)
    W, H, C, _ = batch[0][0].shape
    data = torch.zeros((len(batch), T, C, H, W))
    target = torch.zeros((len(batch), 1))
    for i, (item, L) in enumerate(zip(batch, lengths)):
        data[i, :L] = item[0][..., :L].permute(3, 2, 1, 0)
        target[i] = item[1]
    # the reference squeezes singleton axes before its final permute; for C > 1 and B, T > 1 the result is (B,T,C,H,W)
    return torch.LongTensor(lengths), data.contiguous(), target


def _pad_time_last(items, seq_length):
    """Stack tensors (..., T_i) into (B, T, ...) with zero padding on the right."""
    out = items[0].new_zeros((len(items), seq_length) + tuple(items[0].shape[:-1]))
    for i, t in enumerate(items):
        L = t.shape[-1]
        out[i, :L] = t.movedim(-1, 0)
    return out.contiguous()


def _audio_batch(items, seq_length):
    """(B,T,513) from per-utterance (513,T_i) spectrograms.  Items are tensors in the main process and
    DeferredLogPower objects inside DataLoader workers (packages/processing/deferred.py): those become one batched
    device front-end call, executed here (num_workers=0) or by the parent process when it receives the batch."""
    from packages.processing.deferred import DeferredLogPower, DeferredLogPowerBatch, in_worker
    if not isinstance(items[0], DeferredLogPower):
        return _pad_time_last([it.float() for it in items], seq_length)
    batch = DeferredLogPowerBatch(items)
    return batch if in_worker() else batch.materialise(seq_length)


def collate_many2many_video(batch):
    lengths = [item[-1] for item in batch]
    T = max(lengths)
    data = _pad_time_last([item[0].float() for item in batch], T)      # (B,T,H,W)
    target = _pad_time_last([item[1].float() for item in batch], T)    # (B,T,y_dim)
    return torch.LongTensor(lengths), data, target


def collate_many2many_audio(batch):
    lengths = [item[-1] for item in batch]
    T = max(lengths)
    data = _audio_batch([item[0] for item in batch], T)                # (B,T,x_dim)
    target = _pad_time_last([item[1].float() for item in batch], T)
    return torch.LongTensor(lengths), data, target


def collate_many2many_AV(batch):
    lengths = [item[-1] for item in batch]
    T = max(lengths)
    audio = _audio_batch([item[0] for item in batch], T)               # (B,T,x_dim)
    video = _pad_time_last([item[1].float() for item in batch], T)     # (B,T,H,W)
    target = _pad_time_last([item[2].float() for item in batch], T)    # (B,T,y_dim)
    return torch.LongTensor(lengths), audio, video, target


def _pad_wave(waves, n):
    out = waves[0].new_zeros((len(waves), n))
    for i, w in enumerate(waves):
        out[i, :w.shape[-1]] = w
    return out


def collate_many2many_audio_waveform(batch):
    """Items (wave (N,), label (y_dim,T), time_length, tf_length): waveform batches for the on-device
    front end (packages/utils.py:110-146)."""
    lengths = [item[-1] for item in batch]
    time_lengths = [item[-2] for item in batch]
    data = _pad_wave([item[0].float() for item in batch], max(time_lengths))
    target = _pad_time_last([item[1].float() for item in batch], max(lengths))
    return torch.LongTensor(lengths), data, target


def collate_many2many_AV_waveform(batch):
    """Items (wave, video (H,W,T), label, time_length, tf_length) (packages/utils.py:187-227)."""
    lengths = [item[-1] for item in batch]
    time_lengths = [item[-2] for item in batch]
    T = max(lengths)
    audio = _pad_wave([item[0].float() for item in batch], max(time_lengths))
    video = _pad_time_last([item[1].float() for item in batch], T)
    target = _pad_time_last([item[2].float() for item in batch], T)
    return torch.LongTensor(lengths), audio, video, target
