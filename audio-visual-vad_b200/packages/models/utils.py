"""Initialisers, packed-sequence helpers, loss and metrics with the reference's names and
signatures (packages/models/utils.py:5-55,108-113,164-203).  The VAE-era helpers of that file
(packages/models/utils.py:57-106,116-162: label enumeration, log-sum-exp, Itakura-Saito / ELBO losses, mask MSEs) are
plain host-level torch expressions; they are provided because scripts/train_video_net.py:18 imports one of them."""
import torch
from torch.nn.utils.rnn import pad_packed_sequence


def weights_init_normal(m, mean=0.0, std=0.005):
    name = m.__class__.__name__
    if any(tag in name for tag in ("Linear", "Conv2d", "ConvTranspose2d")):
        m.weight.data.normal_(mean, std)
        if m.bias is not None:
            m.bias.data.zero_()
    elif "Norm" in name or "lstm" in name:
        m.weight.data.normal_(1.0, 0.02)
        if m.bias is not None:
            m.bias.data.zero_()


def method1(packed):
    """List of the un-padded outputs of a packed sequence (utils.py:28-34)."""
    output, sizes = pad_packed_sequence(packed, batch_first=True)
    return [output[i, :sizes[i]] for i in range(output.size(0))]


def method3(packed, lengths):
    """Last valid item of every sequence in a PackedSequence (utils.py:36-55)."""
    starts = torch.cat((torch.zeros(2, dtype=torch.int64), torch.cumsum(packed.batch_sizes, 0)))
    sorted_lengths = lengths[packed.sorted_indices]
    idx = starts[sorted_lengths] + torch.arange(lengths.size(0))
    return packed.data[idx][packed.unsorted_indices]


def binary_cross_entropy(r, x, eps):
    """-mean(x log(sigmoid(r)+eps) + (1-x) log(1-sigmoid(r)+eps)) over all elements (utils.py:113)."""
    p = torch.sigmoid(r)
    return -torch.mean(x * torch.log(p + eps) + (1 - x) * torch.log(1 - p + eps))


def f1_loss(y_hat_hard: torch.Tensor, y: torch.Tensor, epsilon=1e-8):
    """(accuracy, precision, recall, f1) as 0-dim tensors (utils.py:164-203)."""
    y_pred = y_hat_hard.detach()
    y_true = y.detach()
    assert y_true.ndim == 1
    assert y_pred.ndim == 1 or y_pred.ndim == 2
    if y_pred.ndim == 2:
        y_pred = y_pred.argmax(dim=1)
    tp = (y_true * y_pred).sum().to(torch.float32)
    tn = ((1 - y_true) * (1 - y_pred)).sum().to(torch.float32)
    fp = ((1 - y_true) * y_pred).sum().to(torch.float32)
    fn = (y_true * (1 - y_pred)).sum().to(torch.float32)
    accuracy = (tp + tn) / (tp + tn + fp + fn + epsilon)
    precision = tp / (tp + fp + epsilon)
    recall = tp / (tp + fn + epsilon)
    f1 = 2 * (precision * recall) / (precision + recall + epsilon)
    return accuracy, precision, recall, f1


# ---- helpers the VAD scripts import but never reach the device path (packages/models/utils.py:57-162) ----------

def enumerate_discrete(x, y_dim):
    """(batch*y_dim, y_dim) float one-hot rows: `batch` copies of label 0, then of label 1, ... on x's device."""
    batch = x.size(0)
    labels = torch.arange(y_dim, device=x.device).repeat_interleave(batch)
    return torch.nn.functional.one_hot(labels, y_dim).float()


def onehot(k):
    """Returns encode(label) -> length-k one-hot vector (all zeros when label >= k)."""
    def encode(label):
        y = torch.zeros(k)
        if label < k:
            y[label] = 1
        return y
    return encode


def log_sum_exp(tensor, dim=-1, sum_op=torch.sum):
    """log(sum_op(exp(tensor)) + 1e-8-stabilised) along `dim`, keepdim=True, with the max subtracted first."""
    m, _ = torch.max(tensor, dim=dim, keepdim=True)
    return torch.log(sum_op(torch.exp(tensor - m), dim=dim, keepdim=True) + 1e-8) + m


def binary_cross_entropy_2classes(r1, r2, x, eps):
    return -torch.mean(torch.sum(x * torch.log(r1 + eps) + (1 - x) * torch.log(r2 + eps), dim=-1))


def _is_divergence_terms(x, r, eps):
    return x / r - torch.log(x + eps) + torch.log(r) - 1


def ikatura_saito_divergence(r, x, eps):
    return torch.sum(_is_divergence_terms(x, r, eps), dim=-1)


def _kl_terms(mu, logvar):
    return logvar - mu.pow(2) - logvar.exp()


def elbo(x, r, mu, logvar, eps):
    recon = torch.mean(torch.sum(_is_divergence_terms(x, r, eps), dim=-1))
    kl = -0.5 * torch.mean(torch.sum(_kl_terms(mu, logvar), dim=-1))
    return recon + kl, recon, kl


def L_loss(x, r, mu, logvar, eps):
    recon = torch.sum(_is_divergence_terms(x, r, eps), dim=-1)
    kl = -0.5 * torch.sum(_kl_terms(mu, logvar), dim=-1)
    return recon + kl, recon, kl


def U_loss(x, r, mu, logvar, y_hat_soft, eps):
    L, recon, kl = L_loss(x, r, mu, logvar, eps)
    L = L.view_as(y_hat_soft.t()).t()
    H = -y_hat_soft * torch.log(y_hat_soft + eps) - (1 - y_hat_soft) * torch.log(1 - y_hat_soft + eps)
    L_soft = torch.sum(y_hat_soft * L, dim=-1)
    return torch.mean(L_soft - H[:, 0]), torch.mean(L), torch.mean(recon), torch.mean(kl)


def mean_square_error_signal(x, y, y_hat):
    return torch.mean(torch.sum(torch.square((y - y_hat) * x), dim=-1))


def mean_square_error_mask(y, y_hat):
    return torch.mean(torch.sum(torch.square(y - y_hat), dim=-1))


def magnitude_spectrum_approxiamation_loss(x, s, y_hat):
    d = s - y_hat * x
    return torch.mean(torch.sum(torch.real(d * d.conj()), dim=-1))
