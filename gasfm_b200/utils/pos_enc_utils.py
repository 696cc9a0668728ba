"""Sin/cos positional embedding of the observation coordinates; mirror of the interface of the
reference's ``code/utils/pos_enc_utils.py`` (``get_embedder(n_freq, in_dim) -> (fn, d_out)``).
Disabled (n_freq = 0) in every shipped configuration."""
import torch


class Embedder:
    """x -> [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(F-1) x), cos(2^(F-1) x)] (log-spaced bands)."""

    def __init__(self, input_dims, num_freqs, include_input=True):
        self.input_dims = input_dims
        self.include_input = include_input
        self.freq_bands = [float(2.0 ** k) for k in range(num_freqs)]
        self.out_dim = input_dims * ((1 if include_input else 0) + 2 * num_freqs)

    def embed(self, inputs):
        parts = [inputs] if self.include_input else []
        for f in self.freq_bands:
            parts.append(torch.sin(inputs * f))
            parts.append(torch.cos(inputs * f))
        return torch.cat(parts, dim=-1)


def get_embedder(pos_emb_n_freq, in_dim):
    emb = Embedder(in_dim, pos_emb_n_freq)
    return emb.embed, emb.out_dim
