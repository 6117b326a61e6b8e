"""DensityNetwork -- drop-in for the reference's src/network/network.py:5-58: same constructor,
attributes (``layers``, ``activations``, ``bound``, ``encoder``, ``in_dim``), state_dict keys
(``encoder.embeddings``, ``layers.{i}.weight/bias``) and ``forward(x [..,3]) -> [.., out_dim]``.

When the configuration is the one every shipped YAML uses (hash grid with L*C == 32, hidden 32,
out_dim 1) the whole forward -- normalise, 16-level gather, MLP, head -- is ONE kernel and the
backward is one kernel + a tiny deterministic reduction.  Other shapes compose the hash-grid
op with ordinary layers.
"""
import torch
import torch.nn as nn

from ..encoder.hashgrid import HashEncoder
from ..fused import DensityFn, NetMeta


class DensityNetwork(nn.Module):
    def __init__(self, encoder, bound=0.2, num_layers=8, hidden_dim=256, skips=[4], out_dim=1, last_activation="sigmoid"):
        super().__init__()
        self.nunm_layers = num_layers  # (sic) attribute name kept for compatibility with the reference
        self.hidden_dim = hidden_dim
        self.skips = skips
        self.encoder = encoder
        self.in_dim = encoder.output_dim
        self.bound = bound
        self.last_activation = last_activation

        widths_in = [self.in_dim]
        for i in range(1, num_layers - 1):
            widths_in.append(hidden_dim + self.in_dim if i in skips else hidden_dim)
        self.layers = nn.ModuleList([nn.Linear(w, hidden_dim) for w in widths_in])
        self.layers.append(nn.Linear(hidden_dim, out_dim))

        self.activations = nn.ModuleList([nn.LeakyReLU() for _ in range(num_layers - 1)])
        heads = {"sigmoid": nn.Sigmoid, "relu": nn.LeakyReLU, "tanh": nn.Tanh, "none": nn.Identity}
        if last_activation not in heads:
            raise NotImplementedError("Unknown last activation")
        self.activations.append(heads[last_activation]())

        # range check of HashEncoder.forward (hashgrid.py:122): costs one host sync per call;
        # render()'s fused path clamps positions into range and skips it.
        self.check_range = True

    # ---- fused path -----------------------------------------------------------------------
    def fused_meta(self):
        """NetMeta if the fused kernels cover this configuration, else None."""
        enc = self.encoder
        if not isinstance(enc, HashEncoder):
            return None
        meta = NetMeta(enc._offsets_np, enc.input_dim, enc.level_dim, enc.base_resolution, self.in_dim, self.hidden_dim,
                       self.layers[-1].out_features, [s for s in self.skips], self.last_activation, self.bound, len(self.layers))
        meta.arith = getattr(self, "arith", None)   # None: fused.DEFAULT_ARITH; or _lib.ARITH_TC / _lib.ARITH_SIMT for this network
        return meta if meta.fused_supported() else None

    def flat_params(self):
        out = []
        for lin in self.layers:
            out += [lin.weight, lin.bias]
        return out

    def forward(self, x):
        meta = self.fused_meta()
        if meta is not None and x.is_cuda and not x.requires_grad and x.dtype == torch.float32:
            prefix = list(x.shape[:-1])
            flat = x.reshape(-1, 3)
            flags = torch.zeros(1, dtype=torch.int32, device=x.device) if self.check_range else None
            sigma = DensityFn.apply(flat, self.encoder.embeddings, meta, flags, *self.flat_params())
            if flags is not None and (int(flags.item()) & 1):
                raise ValueError(f"HashGrid encoder: inputs range [{x.min().item()}, {x.max().item()}] not in [{-self.bound}, {self.bound}]!")
            return sigma.view(prefix + [1])
        return self._forward_layers(x)

    # ---- generic composition (any encoder / width) -----------------------------------------
    def _forward_layers(self, x):
        x = self.encoder(x, self.bound)
        input_pts = x[..., :self.in_dim]
        for i, (linear, activation) in enumerate(zip(self.layers, self.activations)):
            if i in self.skips:
                x = torch.cat([input_pts, x], -1)
            x = activation(linear(x))
        return x
