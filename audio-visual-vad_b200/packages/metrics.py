"""SI-SDR components and summary statistics (packages/metrics.py:5-131): host-side reporting over a few thousand
scalars, numpy only (scipy's t quantile is used when available, a normal quantile otherwise)."""
import numpy as np


def mean_confidence_interval(data, confidence=0.95, round=3):
    a = 1.0 * np.array(data)
    n = len(a)
    m, se = np.mean(a), np.std(a, ddof=1) / np.sqrt(n) if n > 1 else 0.0
    try:
        from scipy import stats
        h = se * stats.t.ppf((1 + confidence) / 2., n - 1) if n > 1 else 0.0
    except ImportError:
        h = se * 1.959963984540054
    return np.round(m, round), np.round(h, round)


def si_sdr_components(s_hat, s, n):
    alpha_s = np.dot(s_hat, s) / np.linalg.norm(s) ** 2
    s_target = alpha_s * s
    alpha_n = np.dot(s_hat, n) / np.linalg.norm(n) ** 2
    e_noise = alpha_n * n
    e_art = s_hat - s_target - e_noise
    return s_target, e_noise, e_art


def energy_ratios(s_hat, s, n):
    s_target, e_noise, e_art = si_sdr_components(s_hat, s, n)
    si_sdr = 10 * np.log10(np.linalg.norm(s_target) ** 2 / np.linalg.norm(e_noise + e_art) ** 2)
    si_sir = 10 * np.log10(np.linalg.norm(s_target) ** 2 / np.linalg.norm(e_noise) ** 2)
    si_sar = 10 * np.log10(np.linalg.norm(s_target) ** 2 / np.linalg.norm(e_art) ** 2)
    return si_sdr, si_sir, si_sar


def compute_stats(metrics_keys, all_metrics, all_snr_db=None, all_noise_types=None, model_data_dir=None,
                  confidence=0.95, all_speaker_ids=None, **_):
    """Prints mean +- confidence interval per metric (overall, then per SNR / noise type / speaker when given)."""
    arr = np.asarray(all_metrics, dtype=float)
    lines = []
    for i, key in enumerate(metrics_keys):
        m, h = mean_confidence_interval(arr[:, i], confidence=confidence)
        lines.append("{}: {} +- {}".format(key, m, h))
    for name, groups in (("SNR", all_snr_db), ("noise", all_noise_types), ("speaker", all_speaker_ids)):
        if groups is None:
            continue
        groups = np.asarray(groups)
        for gval in np.unique(groups):
            sel = arr[groups == gval]
            for i, key in enumerate(metrics_keys):
                m, h = mean_confidence_interval(sel[:, i], confidence=confidence)
                lines.append("{} {} | {}: {} +- {}".format(name, gval, key, m, h))
    text = "\n".join(lines)
    print(text)
    if model_data_dir:
        with open(str(model_data_dir) + "stats.txt", "w") as f:
            f.write(text + "\n")
    return text
