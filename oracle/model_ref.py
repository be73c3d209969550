"""TEST INFRASTRUCTURE ONLY -- the designed hot path composed from REFERENCE ops on the CPU, used
  * as the parity oracle of isa_b200.archs.EmbeddingPath / ReSeg (same weights, same input), and
  * as the reference arm / cpu_baseline of bench.py (timed on the box's host cores).
Composition (SURVEY.md section 3.1 "designed step"):
  ReNet = nn.GRU rows -> columns          /root/reference/code/lib/archs/modules/README.md:225-256
  attention = MultiHeadAttention            /root/reference/code/lib/archs/modules/utils.py:167-225 (restated: oracle/attention_ref.py)
  loss = DiscriminativeLoss                 /root/reference/code/lib/losses/discriminative.py:162-213 (restated in torch below so that
                                            autograd gives the backward, exactly like the reference's broadcast graph)
  clustering = sklearn KMeans               /root/reference/code/lib/prediction.py:52-85
"""
import numpy as np
import torch
import torch.nn as nn

from .attention_ref import MHARef
from .renet_ref import ReNetRef


class EmbeddingPathRef(nn.Module):
    def __init__(self, n_input=256, n_units=100, n_head=2, d_model=24, d_k=12, d_v=12):
        super().__init__()
        self.renet1 = ReNetRef(n_input, n_units)
        self.renet2 = ReNetRef(2 * n_units, n_units)
        self.attn_in = nn.Linear(2 * n_units, d_model)
        self.attention = MHARef(n_head, d_model, d_k, d_v)

    def forward(self, feats):
        x = self.renet2(self.renet1(feats))
        n, c, h, w = x.shape
        tok = x.permute(0, 2, 3, 1).reshape(n, h * w, c)
        a = self.attn_in(tok)
        a, _ = self.attention(a, a, a)
        return torch.cat([tok, a], dim=2).reshape(n, h, w, -1).permute(0, 3, 1, 2)


def discriminative_loss_torch(input, target, n_objects, max_n_objects, delta_v=0.5, delta_d=1.5, norm=2):
    """The reference's own broadcast-and-reduce graph (discriminative.py:7-62, 65-95 else-branch, 149-160,
    162-188), verbatim in structure so that CPU timing and autograd cost are the reference's."""
    bs, n_filters, height, width = input.size()
    n_instances = target.size(1)
    inp = input.permute(0, 2, 3, 1).contiguous().view(bs, height * width, n_filters)
    tgt = target.permute(0, 2, 3, 1).contiguous().view(bs, height * width, n_instances)
    n_loc = height * width
    pred_repeated = inp.unsqueeze(2).expand(bs, n_loc, n_instances, n_filters)
    gt_expanded = tgt.unsqueeze(3)
    pred_masked = pred_repeated * gt_expanded
    means = []
    for i in range(bs):
        n = int(n_objects[i])
        m = pred_masked[i, :, :n].sum(0) / gt_expanded[i, :, :n].sum(0)
        m = m / torch.norm(m, 2, 1, True)
        if max_n_objects - n != 0:
            m = torch.cat((m, torch.zeros(max_n_objects - n, n_filters)), dim=0)
        means.append(m)
    means = torch.stack(means)
    mexp = means.unsqueeze(1).expand(bs, n_loc, n_instances, n_filters)
    gexp = tgt.unsqueeze(3).expand(bs, n_loc, n_instances, n_filters)
    _var = (torch.clamp(torch.norm((pred_repeated - mexp), norm, 3) - delta_v, min=0.0) ** 2) * gexp[:, :, :, 0]
    var_term = 0.0
    for i in range(bs):
        n = int(n_objects[i])
        var_term = var_term + torch.sum(_var[i, :, :n]) / torch.sum(gexp[i, :, :n, 0])
    var_term = var_term / bs
    t = torch.sum(tgt, dim=2, keepdim=True)
    num = int(torch.sum(t))
    l2 = torch.norm(inp * t, 2, 2)
    reg_term = torch.sum(torch.pow(l2 - 1, 2)) / num
    return 1.0 * var_term + 0.005 * reg_term, means


class _SkipVGG16Ref(nn.Module):
    """Same layers / parameter names as isa_b200.archs.SkipVGG16 (stock convolutions, CPU)."""

    def __init__(self, n_input=3):
        super().__init__()

        def block(cin, cout, n):
            layers = []
            for i in range(n):
                layers += [nn.Conv2d(cin if i == 0 else cout, cout, 3, padding=1), nn.ReLU(inplace=True)]
            return nn.Sequential(*layers)

        self.stage1 = block(n_input, 64, 2)
        self.stage2 = block(64, 128, 2)
        self.stage3 = block(128, 256, 3)
        self.pool = nn.MaxPool2d(2, 2)

    def forward(self, x):
        s1 = self.stage1(x)
        s2 = self.stage2(self.pool(s1))
        return self.stage3(self.pool(s2)), s1, s2


class ReSegRef(nn.Module):
    """isa_b200.archs.ReSeg with the hot path swapped for reference ops; state_dict keys are identical,
    so weights interchange in both directions."""

    def __init__(self, n_classes, n_input=3, n_embedding=24, n_units=100, n_head=2, d_k=12, d_v=12):
        super().__init__()
        self.base = _SkipVGG16Ref(n_input)
        self.path = EmbeddingPathRef(256, n_units, n_head, n_embedding, d_k, d_v)
        c = 2 * n_units + n_embedding
        self.upsampling1 = nn.ConvTranspose2d(c, 100, kernel_size=(2, 2), stride=(2, 2))
        self.relu1 = nn.ReLU()
        self.upsampling2 = nn.ConvTranspose2d(100 + 128, 50, kernel_size=(2, 2), stride=(2, 2))
        self.relu2 = nn.ReLU()
        self.sem_seg_output = nn.Conv2d(50 + 64, n_classes, kernel_size=(1, 1), stride=(1, 1))
        self.ins_seg_output = nn.Conv2d(50 + 64, n_embedding, kernel_size=(1, 1), stride=(1, 1))

    def forward(self, training, x):
        feats, s1, s2 = self.base(x)
        y = self.path(feats)
        y = self.relu1(self.upsampling1(y))
        y = torch.cat((y, s2), dim=1)
        y = self.relu2(self.upsampling2(y))
        y = torch.cat((y, s1), dim=1)
        return self.sem_seg_output(y), self.ins_seg_output(y)
