"""Plot helpers the evaluate scripts import at module top (packages/visualization.py:8-331).  Plotting is out of scope
for the hot path; matplotlib / librosa are imported lazily so that importing this module never fails."""
import numpy as np


def _plt():
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    return plt


def display_multiple_signals(signal_list, fs, vmin=-60, vmax=20, wlen_sec=50e-3, hop_percent=0.5, xticks_sec=1.0,
                             fontsize=50, **_):
    """One row per (signal, spectrogram/mask, optional vad) entry; returns the matplotlib figure."""
    plt = _plt()
    fig, axes = plt.subplots(len(signal_list), 1, figsize=(20, 6 * len(signal_list)), squeeze=False)
    for ax, item in zip(axes[:, 0], signal_list):
        x = np.asarray(item[0])
        if x.ndim == 1:
            ax.plot(np.arange(len(x)) / float(fs), x)
        else:
            ax.imshow(x, origin="lower", aspect="auto", vmin=vmin, vmax=vmax)
    return fig


def display_multiple_spectro(signal_list, fs, vmin=-60, vmax=20, wlen_sec=50e-3, hop_percent=0.5, xticks_sec=1.0,
                             fontsize=50, **_):
    return display_multiple_signals(signal_list, fs, vmin, vmax, wlen_sec, hop_percent, xticks_sec, fontsize)


# ---- single-panel helpers (packages/visualization.py:8-199): same names, arguments and return values (the image /
# figure handle), drawn with matplotlib alone -- the reference goes through librosa.display ----------------------

def _time_axis(frames, fs, wlen_sec, hop_percent):
    nfft = int(wlen_sec * fs)
    hop_sec = int(hop_percent * nfft) / fs
    return hop_sec, frames * hop_sec


def display_waveplot(x, fs=16e3, ymax=1., ymin=-1., xticks_sec=1.0, fontsize=50):
    plt = _plt()
    x = np.asarray(x)
    time_sec = len(x) / fs
    plt.rcParams.update({'font.size': fontsize})
    img = plt.plot(np.arange(len(x)) / float(fs), x)
    plt.ylabel('Amplitude', fontsize=fontsize + 10)
    plt.xlabel('Time (s)', fontsize=fontsize + 10)
    plt.xticks(np.arange(0, time_sec, step=xticks_sec), fontsize=fontsize)
    plt.yticks(fontsize=fontsize)
    plt.ylim(ymin, ymax)
    plt.xlim(0, time_sec)
    return img


def _specshow(values, fs, vmin, vmax, wlen_sec, hop_percent, xticks_sec, cmap, fontsize):
    plt = _plt()
    if values.shape[0] == 1:  # a (1, T) VAD track is drawn as a full-height band, as in the reference
        values = np.repeat(values, 513, axis=0)
    hop_sec, time_sec = _time_axis(values.shape[1], fs, wlen_sec, hop_percent)
    plt.rcParams.update({'font.size': fontsize})
    img = plt.imshow(values, origin='lower', aspect='auto', vmin=vmin, vmax=vmax, cmap=cmap,
                     extent=(0.0, time_sec + hop_sec, 0.0, fs / 2e3))
    plt.ylabel('Frequency (kHz)', fontsize=fontsize + 10)
    plt.xlabel('Time (s)', fontsize=fontsize + 10)
    plt.xticks(np.arange(0, time_sec + hop_sec, step=xticks_sec), fontsize=fontsize)
    plt.yticks(fontsize=fontsize)
    return img


def display_spectrogram(complex_spec, convert_to_db=False, fs=16e3, vmin=-60, vmax=10, wlen_sec=50e-3, hop_percent=0.5,
                        xticks_sec=1.0, cmap='magma', fontsize=50):
    amp = np.abs(np.asarray(complex_spec))
    if convert_to_db:  # librosa.amplitude_to_db defaults: 20 log10(max(amp, 1e-5)), floored 80 dB below the peak
        db = 20.0 * np.log10(np.maximum(amp, 1e-5))
        amp = np.maximum(db, db.max() - 80.0)
    return _specshow(amp, fs, vmin, vmax, wlen_sec, hop_percent, xticks_sec, cmap, fontsize)


def display_power_spectro(psd, fs=16e3, vmin=-60, vmax=10, wlen_sec=50e-3, hop_percent=0.5, xticks_sec=1.0,
                          cmap='magma', fontsize=50):
    p = np.abs(np.asarray(psd))  # librosa.power_to_db defaults: 10 log10(max(p, 1e-10)), floored 80 dB below the peak
    db = 10.0 * np.log10(np.maximum(p, 1e-10))
    return _specshow(np.maximum(db, db.max() - 80.0), fs, vmin, vmax, wlen_sec, hop_percent, xticks_sec, cmap, fontsize)


def display_wav_spectro_mask(x, x_tf, x_ibm, fs=16e3, vmin=-60, vmax=10, wlen_sec=50e-3, hop_percent=0.5,
                             xticks_sec=1.0, fontsize=50):
    """Waveform, dB spectrogram and binary mask stacked, each image with its own colour bar; returns the figure."""
    plt = _plt()
    import matplotlib.gridspec as grd
    fig = plt.figure(figsize=(20, 25))
    gs = grd.GridSpec(3, 2, height_ratios=[5, 10, 10], width_ratios=[10, 0.5], wspace=0.1, hspace=0.3, left=0.08)
    plt.subplot(gs[0])
    display_waveplot(x=x, fs=fs, xticks_sec=xticks_sec, fontsize=fontsize)
    plt.subplot(gs[2])
    display_spectrogram(x_tf, True, fs, vmin, vmax, wlen_sec, hop_percent, xticks_sec, 'magma', fontsize)
    plt.colorbar(cax=plt.subplot(gs[3]), format='%+2.0f dB')
    plt.subplot(gs[4])
    display_spectrogram(x_ibm, False, fs, 0, 1, wlen_sec, hop_percent, xticks_sec, 'Greys_r', fontsize)
    plt.colorbar(cax=plt.subplot(gs[5]), format='%0.1f')
    return fig
