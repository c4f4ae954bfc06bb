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


def compute_stats(metrics_keys, all_metrics, model_data_dir, confidence, all_snr_db=None, all_noise_types=None,
                  all_speakers=None):
    """Prints mean and confidence half-width per metric -- overall, then per input SNR / noise type / speaker when those
    lists are given (packages/metrics.py:62-131: same parameter names and order, same tables; nothing is written to
    disk).  Returns the overall {metric: {'avg', '+/-'}} dictionary for convenience (the reference returns None)."""
    columns = {key: [row[i] for row in all_metrics] for i, key in enumerate(metrics_keys)}
    header = "{:<10} {:<10} {:<10}".format('METRIC', 'AVERAGE', 'CONF. INT.')

    def table(select):
        stats = {}
        print(header)
        for key, col in columns.items():
            m, h = mean_confidence_interval(select(col), confidence=confidence)
            stats[key] = {'avg': m, '+/-': h}
            print("{:<10} {:<10} {:<10}".format(key, m, h))
        print('\n')
        return stats

    overall = table(lambda col: col)
    if all_snr_db is not None:
        snr = np.asarray(all_snr_db)
        for value in np.unique(snr):
            print('Input SNR = {:.2f}'.format(value))
            table(lambda col, v=value: np.asarray(col)[np.where(snr == v)])
    if all_noise_types is not None:
        for value in set(all_noise_types):
            print('Noise type = {}'.format(value))
            table(lambda col, v=value: [x for x, g in zip(col, all_noise_types) if g == v])
    if all_speakers is not None:
        for value in set(all_speakers):
            print('Speaker = {}'.format(value))
            table(lambda col, v=value: [x for x, g in zip(col, all_speakers) if g == v])
    return overall
