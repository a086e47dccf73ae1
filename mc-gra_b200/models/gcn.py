"""Dense GCN victim + embedding model -- same surface as the reference's MC-GRA/models/gcn.py
(`GraphConvolution`, `embedding_GCN`, `GCN` with `.gc` list, `.gc1/.gc2`, `.linear1`, `nclass/nfeat/hidden_sizes`).

Inside the attack loop these modules' forward passes are NOT called: `topology_attack.PGDAttack` reads their
weights and runs the propagation as native kernels over the tiled triangle (engine.py).  The module forward below
serves the one-off work around the loop (victim training `fit`, which SURVEY.md scopes as "stays PyTorch", and the
constants H_A / Y_A): plain torch ops with autograd.
"""
import math
from copy import deepcopy

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim
from torch.nn.modules.module import Module
from torch.nn.parameter import Parameter

from .. import utils


class GraphConvolution(Module):
    """adj @ (input @ W) + b  (models/gcn.py:13-51)."""

    def __init__(self, in_features, out_features, with_bias=True):
        super(GraphConvolution, self).__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = Parameter(torch.empty(in_features, out_features))
        if with_bias:
            self.bias = Parameter(torch.empty(out_features))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1. / math.sqrt(self.weight.size(1))
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.uniform_(-stdv, stdv)

    def forward(self, input, adj):
        support = torch.spmm(input, self.weight) if input.is_sparse else torch.mm(input, self.weight)
        output = torch.spmm(adj, support) if adj.is_sparse else torch.mm(adj, support)
        return output + self.bias if self.bias is not None else output

    def __repr__(self):
        return f"{self.__class__.__name__} ({self.in_features} -> {self.out_features})"


class embedding_GCN(nn.Module):
    """First `nlayer` relu-GC layers of the victim (models/gcn.py:54-85); `.gc` is a plain list that the driver
    replaces with a deepcopy of the victim's (main.py:185-190)."""

    def __init__(self, nfeat, nhid, nlayer=2, with_bias=True, device=None):
        super(embedding_GCN, self).__init__()
        assert device is not None, "Please specify 'device'!"
        self.device = device
        self.nfeat = nfeat
        self.nlayer = nlayer
        self.hidden_sizes = [nhid]
        self.gc1 = GraphConvolution(nfeat, nhid, with_bias=with_bias)
        self.gc = [GraphConvolution(nfeat, nhid, with_bias=with_bias)]
        for _ in range(nlayer - 1):
            self.gc.append(GraphConvolution(nhid, nhid, with_bias=with_bias))
        self.with_bias = with_bias

    def forward(self, x, adj):
        for i in range(self.nlayer):
            layer = self.gc[i].to(self.device)
            x = F.relu(layer(x, adj))
        return x

    def initialize(self):
        self.gc1.reset_parameters()
        for layer in self.gc:
            layer.reset_parameters()

    def set_layers(self, nlayer):
        self.nlayer = nlayer


class GCN(nn.Module):
    """L x relu-GC -> Linear -> log_softmax (models/gcn.py:87-174) with the reference's trainers."""

    def __init__(self, nfeat, nhid, nclass, nlayer=2, dropout=0.5, lr=0.01, weight_decay=5e-4, with_relu=True,
                 with_bias=True, device=None):
        super(GCN, self).__init__()
        assert device is not None, "Please specify 'device'!"
        self.device = device
        self.nfeat = nfeat
        self.hidden_sizes = [nhid]
        self.nclass = nclass
        self.nlayer = nlayer
        self.gc = [GraphConvolution(nfeat, nhid, with_bias=with_bias)]
        for _ in range(nlayer - 1):
            self.gc.append(GraphConvolution(nhid, nhid, with_bias=with_bias))
        self.gc1 = self.gc[0]
        self.gc2 = self.gc[1]
        self.linear1 = nn.Linear(nhid, nclass, bias=with_bias)
        self.dropout = dropout
        self.lr = lr
        self.weight_decay = weight_decay if with_relu else 0
        self.with_relu = with_relu
        self.with_bias = with_bias
        self.output = None
        self.best_model = None
        self.best_output = None
        self.adj_norm = None
        self.features = None

    def forward(self, x, adj):
        for i, layer in enumerate(self.gc):
            layer = layer.to(self.device)
            x = F.relu(layer(x, adj)) if self.with_relu else layer(x, adj)
            if i != len(self.gc) - 1:
                x = F.dropout(x, self.dropout, training=self.training)
        return F.log_softmax(self.linear1(x), dim=1)

    def initialize(self):
        for layer in self.gc:
            layer.reset_parameters()

    def fit(self, features, adj, labels, idx_train, idx_val=None, train_iters=200, initialize=True, verbose=False,
            normalize=True, patience=500, **kwargs):
        """One-off victim training (models/gcn.py:182-241); PyTorch autograd, outside the hot path."""
        self.device = self.gc1.weight.device
        if initialize:
            self.initialize()
        if type(adj) is not torch.Tensor:
            features, adj, labels = utils.to_tensor(features, adj, labels, device=self.device)
        else:
            features, adj, labels = features.to(self.device), adj.to(self.device), labels.to(self.device)
        self.adj_norm = utils.normalize_adj_tensor(adj) if normalize else adj
        self.features = features
        self.labels = labels
        optimizer = optim.Adam(self.parameters(), lr=self.lr, weight_decay=self.weight_decay)
        if idx_val is None:
            self.train()
            for _ in range(train_iters):
                optimizer.zero_grad()
                F.nll_loss(self.forward(self.features, self.adj_norm)[idx_train], labels[idx_train]).backward()
                optimizer.step()
            self.eval()
            self.output = self.forward(self.features, self.adj_norm)
            return
        best_loss_val, best_acc_val, left = 100, 0, patience
        weights = deepcopy(self.state_dict())
        for i in range(train_iters):
            self.train()
            optimizer.zero_grad()
            F.nll_loss(self.forward(self.features, self.adj_norm)[idx_train], labels[idx_train]).backward()
            optimizer.step()
            self.eval()
            with torch.no_grad():
                output = self.forward(self.features, self.adj_norm)
                loss_val = F.nll_loss(output[idx_val], labels[idx_val])
                acc_val = utils.accuracy(output[idx_val], labels[idx_val])
            improved = False
            if best_loss_val > loss_val:
                best_loss_val, improved = loss_val, True
            if patience >= train_iters and acc_val > best_acc_val:
                best_acc_val, improved = acc_val, True
            if improved:
                self.output = output
                weights = deepcopy(self.state_dict())
                left = patience
            elif patience < train_iters:
                left -= 1
                if i > patience and left <= 0:
                    break
        self.load_state_dict(weights)

    def predict(self, features=None, adj=None):
        self.eval()
        if features is None and adj is None:
            return self.forward(self.features, self.adj_norm)
        if type(adj) is not torch.Tensor:
            features, adj = utils.to_tensor(features, adj, device=self.device)
        self.features = features
        self.adj_norm = utils.normalize_adj_tensor(adj)
        return self.forward(self.features, self.adj_norm)

    def test(self, idx_test):
        self.eval()
        output = self.predict()
        acc_test = utils.accuracy(output[idx_test], self.labels[idx_test])
        return acc_test
