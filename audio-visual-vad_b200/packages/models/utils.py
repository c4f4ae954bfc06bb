"""Initialisers, packed-sequence helpers, loss and metrics with the reference's names and
signatures (packages/models/utils.py:5-55,108-113,164-203).  The VAE-era helpers of that file
(elbo, L_loss, U_loss, IS divergence, ...) are unused by every script and are not provided."""
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
