"""Parameter-holding sub-modules with the reference's names, shapes and initialisers.

The compute of these layers is fused into the encoder schedules (``engine.py``); the classes exist
so that ``state_dict`` keys, construction order (RNG consumption) and ``initialize()`` match the
reference (layers.py:151-160, 178-188, 265-278, 294-316).
"""
import math

import torch
import torch.nn as nn


class Attention(nn.Module):
    """additive attention parameters: affine1 [att, feat] (+bias), affine2 [1, att]"""

    def __init__(self, feature_dim: int, attention_dim: int):
        super().__init__()
        self.affine1 = nn.Linear(feature_dim, attention_dim, bias=True)
        self.affine2 = nn.Linear(attention_dim, 1, bias=False)

    def initialize(self):
        nn.init.xavier_uniform_(self.affine1.weight, gain=nn.init.calculate_gain('tanh'))
        nn.init.zeros_(self.affine1.bias)
        nn.init.xavier_uniform_(self.affine2.weight)


class ScaledDotProduct_CandidateAttention(nn.Module):
    """K [att, feat] (no bias), Q [att, query] (+bias); scores scaled by 1/sqrt(att)"""

    def __init__(self, feature_dim: int, query_dim: int, attention_dim: int):
        super().__init__()
        self.K = nn.Linear(feature_dim, attention_dim, bias=False)
        self.Q = nn.Linear(query_dim, attention_dim, bias=True)
        self.attention_scalar = math.sqrt(float(attention_dim))

    def initialize(self):
        nn.init.xavier_uniform_(self.K.weight)
        nn.init.xavier_uniform_(self.Q.weight)
        nn.init.zeros_(self.Q.bias)


class GCNLayer(nn.Module):
    def __init__(self, in_dim, out_dim, residual=False, layer_norm=False):
        super().__init__()
        self.residual = residual
        self.layer_norm = layer_norm
        if self.residual and in_dim != out_dim:
            raise Exception('To facilitate residual connection, in_dim must equal to out_dim')
        self.W = nn.Linear(in_dim, out_dim, bias=True)
        if self.layer_norm:                                   # layers.py:274-275 (same construction order)
            self.layer_normalization = nn.LayerNorm(normalized_shape=[out_dim])

    def initialize(self):
        nn.init.xavier_uniform_(self.W.weight, gain=nn.init.calculate_gain('relu'))
        nn.init.zeros_(self.W.bias)


class GCN(nn.Module):
    def __init__(self, in_dim, out_dim, hidden_dim=0, num_layers=1, dropout=0.1, residual=False, layer_norm=False):
        super().__init__()
        self.num_layers = num_layers
        self.dropout_rate = dropout
        layers = []
        if num_layers == 1:
            layers.append(GCNLayer(in_dim, out_dim, residual=residual, layer_norm=layer_norm))
        else:
            self.dropout = nn.Dropout(dropout, inplace=True)
            layers.append(GCNLayer(in_dim, hidden_dim, residual=residual, layer_norm=layer_norm))
            for _ in range(1, num_layers - 1):
                layers.append(GCNLayer(hidden_dim, hidden_dim, residual=residual, layer_norm=layer_norm))
            layers.append(GCNLayer(hidden_dim, out_dim, residual=residual, layer_norm=layer_norm))
        self.gcn_layers = nn.ModuleList(layers)

    def initialize(self):
        for layer in self.gcn_layers:
            layer.initialize()
