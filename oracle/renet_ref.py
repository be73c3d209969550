"""TEST INFRASTRUCTURE ONLY -- the ReNet oracle.

`renet.py` does not exist in the reference tree (reseg.py:5 and attenet2.py:4 have the import
commented out); the only in-tree statement is the contract in
/root/reference/code/lib/archs/modules/README.md:225-256: a bidirectional GRU over every row,
then a bidirectional GRU over every column of the result, (N,C,H,W) -> (N,2*n_units,H/ph,W/pw).
The oracle is therefore torch.nn.GRU (CPU, fp32 or fp64) composed exactly that way; every ReNet
parity claim names this file.  "parity unpinned" by reference tests: none exist for ReNet.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class ReNetRef(nn.Module):
    def __init__(self, n_input, n_units, patch_size=(1, 1)):
        super().__init__()
        self.ph, self.pw = patch_size
        self.rnn_hor = nn.GRU(n_input * self.ph * self.pw, n_units, num_layers=1, batch_first=True, bidirectional=True)
        self.rnn_ver = nn.GRU(n_units * 2, n_units, num_layers=1, batch_first=True, bidirectional=True)

    def tile(self, x):
        ph, pw = self.ph, self.pw
        n_h_pad = (ph - x.size(2) % ph) % ph
        n_w_pad = (pw - x.size(3) % pw) % pw
        if n_h_pad or n_w_pad:
            x = F.pad(x, (n_w_pad // 2, n_w_pad - n_w_pad // 2, n_h_pad // 2, n_h_pad - n_h_pad // 2))
        b, c, h, w = x.size()
        x = x.view(b, c, h // ph, ph, w // pw, pw).permute(0, 2, 4, 1, 3, 5)
        return x.contiguous().view(b, h // ph, w // pw, ph * pw * c).permute(0, 3, 1, 2)

    def forward(self, x):
        if self.ph != 1 or self.pw != 1:
            x = self.tile(x)
        b, c, h, w = x.shape
        rows = x.permute(0, 2, 3, 1).reshape(b * h, w, c)
        rows, _ = self.rnn_hor(rows)
        x = rows.reshape(b, h, w, -1)
        cols = x.permute(0, 2, 1, 3).reshape(b * w, h, x.shape[-1])
        cols, _ = self.rnn_ver(cols)
        x = cols.reshape(b, w, h, -1).permute(0, 3, 2, 1)
        return x.contiguous()
