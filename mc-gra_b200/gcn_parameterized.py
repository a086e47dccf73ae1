"""`gcn_parameterized.PGDAttack` -- the GCN-parameterised adjacency estimate (MC-GRA/gcn_parameterized.py).

Only `get_modified_adj` is runnable in the reference (its `attack()` crashes on the missing `adj_changes`,
SURVEY.md 0.2); that forward is what this module provides: A = zhat zhat^T with zhat the row-normalised output of the
relu-GCN stack applied with the IDENTITY adjacency (gcn_parameterized.py:406-416), gram evaluated by the native kernel."""
from copy import deepcopy

import torch
import torch.nn.functional as F

from . import _native as N
from ._native import call, ptr
from .base_attack import BaseAttack


class PGDAttack(BaseAttack):
    def __init__(self, features=None, model=None, embedding=None, H_A=None, Y_A=None, nnodes=None, loss_type='CE',
                 feature_shape=None, attack_structure=True, attack_features=False, device='cpu'):
        super(PGDAttack, self).__init__(model, nnodes, attack_structure, attack_features, device)
        assert attack_features or attack_structure, 'attack_features or attack_structure cannot be both False'
        self.loss_type = loss_type
        self.features = features.to(device)
        self.embedding = embedding
        self.H_A = H_A
        self.Y_A = Y_A
        self.complementary = None
        self.complementary_after = None
        if attack_structure:
            assert nnodes is not None, 'Please give nnodes='
            self.gc = deepcopy(embedding.gc)

    def get_modified_adj(self, ori_adj=None):
        """zhat zhat^T, zhat = F.normalize(relu-GCN stack(X, I)) (gcn_parameterized.py:406-416)."""
        x = self.features.detach().to(self.device, torch.float32)
        with torch.no_grad():
            for layer in self.gc:                 # adjacency is the identity: every layer is x W + b
                layer = layer.to(self.device)
                x = F.relu(x @ layer.weight + (layer.bias if layer.bias is not None else 0))
        x = x.contiguous()
        n, d = x.shape
        st = N.stream_ptr()
        z = torch.empty_like(x)
        call("mcgra_row_normalize", ptr(x), n, d, 2.0, ptr(z), st)
        out = torch.zeros(n, n, dtype=torch.float32, device=x.device)
        call("mcgra_gram_accumulate", ptr(z), d, n, 3, None, ptr(out), n, 0, n, st)
        return out

    def attack(self, *args, **kwargs):
        raise AttributeError("'PGDAttack' object has no attribute 'adj_changes' -- the reference's "
                             "gcn_parameterized.PGDAttack.attack is dead as shipped (gcn_parameterized.py:120-122,165)")
