"""FreqEncoder -- NeRF positional encoding, drop-in for the reference's src/encoder/freqencoder.py
(config #1 of BASELINE.json uses it as the CPU-runnable encoder).  Elementwise glue; not part of
the fused hash-grid path."""
import torch
import torch.nn as nn


class FreqEncoder(nn.Module):
    def __init__(self, input_dim, max_freq_log2, N_freqs, log_sampling=True, include_input=True,
                 periodic_fns=(torch.sin, torch.cos)):
        super().__init__()
        self.input_dim = input_dim
        self.include_input = include_input
        self.periodic_fns = periodic_fns
        self.output_dim = (input_dim if include_input else 0) + input_dim * N_freqs * len(periodic_fns)
        if log_sampling:
            bands = 2.0 ** torch.linspace(0.0, max_freq_log2, N_freqs)
        else:
            bands = torch.linspace(2.0 ** 0.0, 2.0 ** max_freq_log2, N_freqs)
        self.freq_bands = bands.numpy().tolist()

    def forward(self, input, bound):
        out = [input] if self.include_input else []
        for freq in self.freq_bands:
            scaled = input * freq
            out.extend(fn(scaled) for fn in self.periodic_fns)
        return torch.cat(out, dim=-1)
