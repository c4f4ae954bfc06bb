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
