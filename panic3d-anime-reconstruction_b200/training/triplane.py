"""``OSGDecoder`` with the reference's parameter names and call surface (reference
training/triplane.py:516-548; FullyConnectedLayer: training/networks_stylegan2.py:101-136).

The reference's ``TriPlaneGenerator`` itself stays reference Python: it
instantiates ``ImportanceRenderer`` / ``RaySampler`` by import name, which ``panic3d_b200.dropin``
redirects here.  This stand-alone decoder exists so the package is usable (bench, tests, serving)
without the reference tree; its state_dict keys match (`net.0.weight`, `net.0.bias`, `net.2.*`)."""
import numpy as np
import torch


class FullyConnectedLayer(torch.nn.Module):
    """Parameter container with StyleGAN2's equalised-learning-rate gains (linear activation only)."""

    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        if activation != 'linear':
            raise NotImplementedError('only the linear FullyConnectedLayer of OSGDecoder is provided')
        self.in_features, self.out_features, self.activation = in_features, out_features, activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def extra_repr(self):
        return f'in_features={self.in_features:d}, out_features={self.out_features:d}, activation={self.activation:s}'


class OSGDecoder(torch.nn.Module):
    """32 -> 64 (softplus) -> 1+32.  Evaluated inside the CUDA renderer; ``forward`` on its own goes
    through ``ImportanceRenderer.run_model``-style kernels only via the renderer (no eager path)."""

    def __init__(self, n_features, options):
        super().__init__()
        self.hidden_dim = 64
        self.force_sigmoid = False
        self.net = torch.nn.Sequential(
            FullyConnectedLayer(n_features, self.hidden_dim, lr_multiplier=options['decoder_lr_mul']),
            torch.nn.Softplus(),
            FullyConnectedLayer(self.hidden_dim, 1 + options['decoder_output_dim'], lr_multiplier=options['decoder_lr_mul']),
        )

    def forward(self, sampled_features, ray_directions, force_sigmoid=None):
        raise NotImplementedError('OSGDecoder is evaluated inside panic3d_b200.ImportanceRenderer (fused gather+MLP); '
                                  'use renderer.run_model(planes, decoder, coords, dirs, options) for point queries')

    def set_force_sigmoid(self, state):
        self.force_sigmoid = state
        return self.force_sigmoid
