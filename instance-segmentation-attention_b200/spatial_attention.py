"""Spatial / channel / local attention layers of the instance-embedding block, reference names and
signatures (/root/reference/code/lib/archs/modules/utils.py):

  Decoder(num_layers, d_model, d_inner, n_head, d_k, d_v)            utils.py:49-69    single-query readout
  _ScalePDAttention(d_k, d_v, d_model, dilation_rate, n_head=2)      utils.py:248-303  local 3x3 dilated attention
  _AttenAsppBlock(dilation_rate, d_model, d_k, d_v, d_inner, n_head) utils.py:72-135
  AttentionLayer(channel, reduction=2, multiply=True)                utils.py:402-420  squeeze-excite
  SpatialAttentionLayer(d_model, reduction=2, d_h=None, multiply=True)  utils.py:457-523 masked spatial softmax
  maskBN(num_features, ...)                                          utils.py:529-591
  HardAttentionLayer(d_model, d_k, d_h, random_thred=0.2, reduction=2)  utils.py:613-663 per-instance softmax
  make_position_encoding(xp, batch, length, n_units, f=10000.)       utils.py:332-344
  position_encoding_2d(n_units, h, w)                                reseg.py:132-137 (ReSeg.set_position_encoding)

The HBM-bound parts -- masked softmax over H*W, the per-instance (b, n, H*W) expansion, global pooling and
gating, the readout contraction, the nine-neighbour stencil -- are sm_100a kernels behind the C-ABI
(csrc/spatial_ops.cu, csrc/local_attention.cu) with hand-written backward passes; 1x1 / 3x3 convolutions,
BatchNorm and pooling stay PyTorch (cuDNN, out of scope).  Masks are 0/1 valued (the reference casts
`1 - mask` to uint8, so only 0/1 is meaningful there as well).  No CPU fallback.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.parameter import Parameter

from . import _lib
from .attention import DecoderLayer


def _f32c(t):
    return t.contiguous().float()


# ----------------------------------------------------------------------------- autograd wrappers
class _MaskedSoftmaxHW(torch.autograd.Function):
    """y[b,k,:] = softmax over {p : mask[b,k,p] != 0} of x[b,:], times scale[b,k]; 0 outside the mask."""

    @staticmethod
    def forward(ctx, x, mask, scale, nan_to_zero):
        lib = _lib.load()
        _lib.require_cuda(x, "x")
        B, HW = x.shape
        K = mask.shape[1]
        assert mask.shape[0] == B and mask.shape[2] == HW
        x = _f32c(x)
        if mask.dtype == torch.float32:
            kind = 1
        else:
            kind = 0
            mask = mask.view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8)
        mask = mask.contiguous()
        scale_c = _f32c(scale.reshape(B * K)) if scale is not None else None
        y = torch.empty(B, K, HW, device=x.device, dtype=torch.float32)
        stats = torch.empty(B * K, 2, device=x.device, dtype=torch.float32)
        wsb = lib.isa_masked_softmax_hw_workspace_bytes(B, K, HW)
        ws = torch.empty(wsb, device=x.device, dtype=torch.uint8)
        rc = lib.isa_masked_softmax_hw_fwd(_lib.ptr(x), _lib.ptr(mask), kind, B, K, HW, _lib.ptr(scale_c), int(nan_to_zero),
                                           _lib.ptr(y), _lib.ptr(stats), _lib.ptr(ws), wsb, _lib.stream_ptr(x.device))
        _lib.check(rc, "isa_masked_softmax_hw_fwd")
        ctx.save_for_backward(y, stats, scale_c)
        ctx.dims = (B, K, HW)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        y, stats, scale_c = ctx.saved_tensors
        B, K, HW = ctx.dims
        dy = _f32c(dy)
        dx = torch.empty(B, HW, device=y.device, dtype=torch.float32)
        wsb = lib.isa_masked_softmax_hw_workspace_bytes(B, K, HW)
        ws = torch.empty(wsb, device=y.device, dtype=torch.uint8)
        rc = lib.isa_masked_softmax_hw_bwd(_lib.ptr(y), _lib.ptr(dy), _lib.ptr(stats), _lib.ptr(scale_c), B, K, HW,
                                           _lib.ptr(dx), _lib.ptr(ws), wsb, _lib.stream_ptr(y.device))
        _lib.check(rc, "isa_masked_softmax_hw_bwd")
        return dx, None, None, None


def masked_softmax_hw(x, mask, scale=None, nan_to_zero=False):
    """x (B, HW), mask (B, K, HW) 0/1 (float32 / uint8 / bool), scale (B, K) or None -> (B, K, HW)."""
    return _MaskedSoftmaxHW.apply(x, mask, scale, nan_to_zero)


def _row_dot(a, b, rows, HW, b_rows_div):
    lib = _lib.load()
    out = torch.empty(rows, device=a.device, dtype=torch.float32)
    wsb = lib.isa_row_dot_workspace_bytes(rows, HW)
    ws = torch.empty(wsb, device=a.device, dtype=torch.uint8)
    rc = lib.isa_row_dot(_lib.ptr(a), _lib.ptr(b), rows, HW, b_rows_div, _lib.ptr(out), _lib.ptr(ws), wsb, _lib.stream_ptr(a.device))
    _lib.check(rc, "isa_row_dot")
    return out


def _row_affine(x, g, c, rows, HW):
    lib = _lib.load()
    y = torch.empty_like(x)
    rc = lib.isa_row_affine(_lib.ptr(x), _lib.ptr(g), _lib.ptr(c), rows, HW, _lib.ptr(y), _lib.stream_ptr(x.device))
    _lib.check(rc, "isa_row_affine")
    return y


class _SqueezeExcite(torch.autograd.Function):
    """y = x * sigmoid(W2 relu(W1 mean_hw(x) + b1) + b2): two streaming passes over x forward (pool, gate), two
    backward (dgate = <dy, x>, dx = dy * gate + dmean / HW); the (b, c) MLP and its gradient are a few tiny ops."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, multiply):
        _lib.require_cuda(x, "x")
        x = _f32c(x)
        b, c, h, w = x.shape
        HW = h * w
        mean = _row_dot(x, None, b * c, HW, 1).view(b, c) / HW
        hid = torch.relu(F.linear(mean, w1, b1))
        gate = torch.sigmoid(F.linear(hid, w2, b2))
        ctx.save_for_backward(x, w1, w2, mean, hid, gate)
        ctx.multiply = multiply
        if not multiply:
            return gate.view(b, c, 1, 1)
        return _row_affine(x, gate.reshape(-1).contiguous(), None, b * c, HW)

    @staticmethod
    def backward(ctx, dy):
        x, w1, w2, mean, hid, gate = ctx.saved_tensors
        b, c, h, w = x.shape
        HW = h * w
        if ctx.multiply:
            dy = _f32c(dy)
            dgate = _row_dot(dy, x, b * c, HW, 1).view(b, c)
        else:
            dgate = dy.reshape(b, c).float()
        dz2 = dgate * gate * (1 - gate)
        dw2 = dz2.t() @ hid
        db2 = dz2.sum(0)
        dz1 = (dz2 @ w2) * (hid > 0).float()
        dw1 = dz1.t() @ mean
        db1 = dz1.sum(0)
        dmean = (dz1 @ w1) / HW
        if ctx.multiply:
            dx = _row_affine(dy, gate.reshape(-1).contiguous(), dmean.reshape(-1).contiguous(), b * c, HW)
        else:
            dx = dmean.view(b, c, 1, 1).expand(b, c, h, w).contiguous()
        return dx, dw1, db1, dw2, db2, None


class _Readout(torch.autograd.Function):
    """sigmoid(bmm(q (b,1,C), enc (b,C,HW))) -> (b, HW)"""

    @staticmethod
    def forward(ctx, q, enc):
        lib = _lib.load()
        _lib.require_cuda(enc, "enc_output")
        q, enc = _f32c(q), _f32c(enc)
        B, C, HW = enc.shape
        out = torch.empty(B, HW, device=enc.device, dtype=torch.float32)
        rc = lib.isa_readout_fwd(_lib.ptr(q), _lib.ptr(enc), B, C, HW, _lib.ptr(out), _lib.stream_ptr(enc.device))
        _lib.check(rc, "isa_readout_fwd")
        ctx.save_for_backward(q, enc, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        q, enc, out = ctx.saved_tensors
        B, C, HW = enc.shape
        dout = _f32c(dout)
        dz = torch.empty_like(out)
        denc = torch.empty_like(enc) if ctx.needs_input_grad[1] else None
        rc = lib.isa_readout_bwd(_lib.ptr(q), _lib.ptr(out), _lib.ptr(dout), B, C, HW, _lib.ptr(dz), _lib.ptr(denc),
                                 _lib.stream_ptr(enc.device))
        _lib.check(rc, "isa_readout_bwd")
        dq = _row_dot(enc, dz, B * C, HW, C).view(B, C) if ctx.needs_input_grad[0] else None
        return dq, denc


class _LocalAttention(torch.autograd.Function):

    @staticmethod
    def forward(ctx, Q, K, V, nomask, dil, scale):
        lib = _lib.load()
        _lib.require_cuda(Q, "Q")
        Q, K, V = _f32c(Q), _f32c(K), _f32c(V)
        Bh, dk, h, w = Q.shape
        dv = V.shape[1]
        nm = None
        mb = 1
        if nomask is not None:
            nm = _f32c(nomask).reshape(-1, h, w)
            mb = nm.shape[0]
        need = any(ctx.needs_input_grad[:3])
        out = torch.empty(Bh, dv, h, w, device=Q.device, dtype=torch.float32)
        P = torch.empty(Bh, 9, h, w, device=Q.device, dtype=torch.float32) if need else None
        rc = lib.isa_local_attention_fwd(_lib.ptr(Q), _lib.ptr(K), _lib.ptr(V), _lib.ptr(nm), mb, Bh, dk, dv, h, w, int(dil),
                                         float(scale), _lib.ptr(out), _lib.ptr(P), _lib.stream_ptr(Q.device))
        _lib.check(rc, "isa_local_attention_fwd")
        if need:
            ctx.save_for_backward(Q, K, V, P)
            ctx.cfg = (int(dil), float(scale))
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        Q, K, V, P = ctx.saved_tensors
        dil, scale = ctx.cfg
        Bh, dk, h, w = Q.shape
        dv = V.shape[1]
        dout = _f32c(dout)
        dQ, dK, dV = torch.empty_like(Q), torch.empty_like(K), torch.empty_like(V)
        dS = torch.empty_like(P)
        rc = lib.isa_local_attention_bwd(_lib.ptr(Q), _lib.ptr(K), _lib.ptr(V), _lib.ptr(P), _lib.ptr(dout), Bh, dk, dv, h, w, dil,
                                         scale, _lib.ptr(dQ), _lib.ptr(dK), _lib.ptr(dV), _lib.ptr(dS), _lib.stream_ptr(Q.device))
        _lib.check(rc, "isa_local_attention_bwd")
        return dQ, dK, dV, None, None, None


def local_attention(Q, K, V, nomask, dilation, scale):
    """Q, K (Bh, d_k, h, w), V (Bh, d_v, h, w); nomask (mb, 1, h, w) non-zero = excluded -> (Bh, d_v, h, w)."""
    return _LocalAttention.apply(Q, K, V, nomask, dilation, scale)


# ----------------------------------------------------------------------------- reference-named modules
class Decoder(nn.Module):
    """utils.py:49-69.  `layer_stack` is built (and therefore part of the state_dict) but, as in the reference,
    forward only evaluates the single-query readout."""

    def __init__(self, num_layers, d_model, d_inner, n_head, d_k, d_v):
        super(Decoder, self).__init__()
        self.num_layers = num_layers
        self.layer_stack = nn.ModuleList([
            DecoderLayer(d_model, d_inner, n_head, d_k, d_v, last=(i == num_layers - 1))
            for i in range(num_layers)])
        self.linear = None

    def forward(self, input, enc_output, mask=None):
        b, c, h, w = enc_output.shape
        return _Readout.apply(input, enc_output.reshape(b, c, h * w))


class _ScalePDAttention(nn.Module):
    """utils.py:248-303"""

    def __init__(self, d_k, d_v, d_model, dilation_rate, n_head=2):
        super(_ScalePDAttention, self).__init__()
        self.qk_w = nn.Conv2d(in_channels=d_model // n_head, out_channels=2 * d_k, kernel_size=(1, 1), stride=(1, 1))
        self.v_w = nn.Conv2d(in_channels=d_model // n_head, out_channels=d_v, kernel_size=(1, 1), stride=(1, 1))
        self.fc = nn.Conv2d(in_channels=n_head * d_v, out_channels=d_model, kernel_size=(1, 1), stride=(1, 1))
        self.layer_norm = nn.InstanceNorm2d(d_model)
        self.d = dilation_rate
        self.d_k = d_k
        self.n_head = n_head
        self.d_v = d_v
        self.d_model = d_model
        nn.init.normal_(self.qk_w.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.v_w.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_v)))
        nn.init.xavier_normal_(self.fc.weight)

    def forward(self, qk, v, nomask=None):
        '''qk: b, c, h, w'''
        residual = qk
        b0, c0, h, w = qk.shape
        qk = qk.reshape(b0 * self.n_head, c0 // self.n_head, h, w)
        v = v.reshape(b0 * self.n_head, c0 // self.n_head, h, w)
        c = c0 // self.n_head
        QK = self.qk_w(qk)
        V = self.v_w(v)
        Q, K = torch.split(QK, [self.d_k, self.d_k], dim=1)
        # image i of the folded batch reads nomask[i % b0]: the reference tiles the mask with .repeat (utils.py:272)
        atten = local_attention(Q, K, V, nomask, self.d, c ** -0.5)
        atten = atten.reshape(b0, self.d_v * self.n_head, h, w)
        out = self.fc(atten)
        out = self.layer_norm(out + residual)
        return out


class _AttenAsppBlock(nn.Module):
    """utils.py:72-135"""

    def __init__(self, dilation_rate, d_model, d_k, d_v, d_inner, n_head, bn_start=True):
        super(_AttenAsppBlock, self).__init__()
        self.d = dilation_rate
        self.d_v = d_v
        self.n_head = n_head
        self.attention = _ScalePDAttention(d_k, d_v, d_model, dilation_rate)
        self.w1 = nn.Conv2d(in_channels=d_model, out_channels=d_inner, kernel_size=(1, 1), stride=(1, 1))
        self.act = nn.LeakyReLU(0.01)
        self.w2 = nn.Conv2d(in_channels=d_inner, out_channels=d_model, kernel_size=(1, 1), stride=(1, 1))
        self.layer_norm = nn.InstanceNorm2d(d_model)

    def forward(self, _input, mask=None):
        '''mask: b, 1, h, w (1 = inside)'''
        nomask = 1 - mask
        atten = self.attention(_input, _input, nomask)
        feedF = self.w2(self.act(self.w1(atten)))
        return self.layer_norm(feedF + atten)


class AttentionLayer(nn.Module):
    """utils.py:402-420 (squeeze-excite channel gate)"""

    def __init__(self, channel, reduction=2, multiply=True):
        super(AttentionLayer, self).__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(
            nn.Linear(channel, channel // reduction),
            nn.ReLU(inplace=True),
            nn.Linear(channel // reduction, channel),
            nn.Sigmoid())
        self.multiply = multiply

    def forward(self, x):
        return _SqueezeExcite.apply(x, self.fc[0].weight, self.fc[0].bias, self.fc[2].weight, self.fc[2].bias,
                                    self.multiply == True)  # noqa: E712 (the reference compares with == True)


class SpatialAttentionLayer(nn.Module):
    """utils.py:457-523.  Base (b,c,h,w), y_ (b,1,h,w) 0/1 mask (or an int: no mask), h_t (b,c) or None."""

    def __init__(self, d_model, reduction=2, d_h=None, multiply=True):
        super(SpatialAttentionLayer, self).__init__()
        self.l_v = nn.Conv2d(d_model, d_model // reduction, 1, 1)
        self.l_h = nn.Linear(d_model, d_model // reduction, bias=False)
        self.spatial_fc = nn.Sequential(
            nn.Tanh(),
            nn.Conv2d(d_model // reduction, 1, 1, 1))
        self.bn = nn.BatchNorm2d(d_model)
        self.multiply = multiply

    def forward(self, Base, y_, h_t=None, use_sigmoid=False, decoder=False):
        b, c, h, w = Base.size()
        base = self.l_v(Base * y_)
        if h_t is None:
            Base_mask = Base * y_
            h_t = torch.mean(Base_mask.view(b, c, -1), dim=2)
        h_t = self.l_h(h_t)
        h_t = h_t.view(b, h_t.shape[1], 1, 1)
        base = base + h_t
        beta = self.spatial_fc(base)   # b, 1, h, w
        if use_sigmoid:
            beta = torch.sigmoid(beta).view(b, 1, h, w)
        else:
            if type(y_) != int and not decoder:
                mask = y_.reshape(b, 1, h * w)
                y_sum = torch.sum(y_, dim=[1, 2, 3]).reshape(b, 1).float()
            else:
                mask = torch.ones(b, 1, h * w, device=Base.device, dtype=torch.uint8)
                y_sum = torch.full((b, 1), float(h * w), device=Base.device)
            beta = masked_softmax_hw(beta.reshape(b, h * w), mask, y_sum, False).view(b, 1, h, w)
        if self.multiply:
            paste = self.bn(Base * beta)
            paste = paste * y_
            return Base + paste
        return beta


class _MaskBN(torch.autograd.Function):
    """Training-mode maskBN (utils.py:568-591) on the device: masked statistics, normalisation and the gradient through
    mean and var, as deterministic two-stage reductions (csrc/spatial_ops.cu mbn_*).  Returns (y, mean, var)."""

    @staticmethod
    def forward(ctx, x, mask, weight, bias, eps):
        lib = _lib.load()
        _lib.require_cuda(x, "input")
        x = _f32c(x)
        b, c, h, w = x.shape
        mask = _f32c(mask.to(torch.float32).reshape(b, -1, h * w))
        mc = mask.shape[1]
        y = torch.empty_like(x)
        stats = torch.empty(2 * c + b * mc, device=x.device, dtype=torch.float32)
        wsb = lib.isa_mask_bn_workspace_bytes(b, c, h * w)
        ws = torch.empty(wsb, device=x.device, dtype=torch.uint8)
        wt = _f32c(weight) if weight is not None else None
        bs = _f32c(bias) if bias is not None else None
        rc = lib.isa_mask_bn_fwd(_lib.ptr(x), _lib.ptr(mask), mc, b, c, h * w, _lib.ptr(wt), _lib.ptr(bs), float(eps),
                                 _lib.ptr(y), _lib.ptr(stats), _lib.ptr(ws), wsb, _lib.stream_ptr(x.device))
        _lib.check(rc, "isa_mask_bn_fwd")
        ctx.save_for_backward(x, mask, stats, wt)
        ctx.meta = (b, c, h * w, mc, float(eps), weight is not None, bias is not None)
        mean, var = stats[:c].clone(), stats[c:2 * c].clone()
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, dy, _dmean, _dvar):
        lib = _lib.load()
        x, mask, stats, wt = ctx.saved_tensors
        b, c, hw, mc, eps, has_w, has_b = ctx.meta
        dy = _f32c(dy)
        dx = torch.empty_like(x)
        dw = torch.empty(c, device=x.device, dtype=torch.float32) if has_w else None
        db = torch.empty(c, device=x.device, dtype=torch.float32) if has_b else None
        wsb = lib.isa_mask_bn_workspace_bytes(b, c, hw)
        ws = torch.empty(wsb, device=x.device, dtype=torch.uint8)
        rc = lib.isa_mask_bn_bwd(_lib.ptr(x), _lib.ptr(mask), mc, _lib.ptr(dy), _lib.ptr(stats), _lib.ptr(wt), eps, b, c, hw,
                                 _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ws), wsb, _lib.stream_ptr(x.device))
        _lib.check(rc, "isa_mask_bn_bwd")
        return dx, None, dw, db, None


class maskBN(nn.Module):
    """utils.py:529-591: batch statistics weighted by a (b,c,h,w) mask, per-image normaliser sum(mask) + 1.
    The running-average update keeps the reference's (reversed) momentum convention (utils.py:583-584)."""
    _version = 2

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super(maskBN, self).__init__()
        self.num_features = num_features
        self.eps = eps
        self.momentum = momentum
        self.affine = affine
        self.track_running_stats = track_running_stats
        if self.affine:
            self.weight = Parameter(torch.Tensor(num_features))
            self.bias = Parameter(torch.Tensor(num_features))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)
        if self.track_running_stats:
            self.register_buffer('running_mean', torch.zeros(num_features))
            self.register_buffer('running_var', torch.ones(num_features))
            self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))
        else:
            self.register_parameter('running_mean', None)
            self.register_parameter('running_var', None)
            self.register_parameter('num_batches_tracked', None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.track_running_stats:
            self.running_mean.zero_()
            self.running_var.fill_(1)
            self.num_batches_tracked.zero_()
        if self.affine:
            self.weight.data.uniform_()
            self.bias.data.zero_()

    def forward(self, input, mask):
        b, c, h, w = input.shape
        factor = 0.0
        if self.training and self.track_running_stats:
            self.num_batches_tracked += 1
            factor = 1.0 / self.num_batches_tracked.item() if self.momentum is None else self.momentum
        if self.training or not self.track_running_stats:
            assert mask.shape[1] in (1, c), "maskBN: the mask has one channel or one per feature"
            out, mean, var = _MaskBN.apply(input, mask, self.weight if self.affine else None,
                                           self.bias if self.affine else None, self.eps)
            if self.track_running_stats:
                with torch.no_grad():      # the reference's convention (utils.py:580-581): running * factor + (1 - factor) * batch
                    self.running_mean.copy_(self.running_mean * factor + (1 - factor) * mean)
                    self.running_var.copy_(self.running_var * factor + (1 - factor) * var)
            return out
        w_ = self.weight.view(1, c, 1, 1) if self.affine else 1.0
        b_ = self.bias.view(1, c, 1, 1) if self.affine else 0.0
        return (input - self.running_mean.view(1, c, 1, 1)) / torch.pow(self.running_var.view(1, c, 1, 1) + self.eps, 0.5) * w_ + b_

    def extra_repr(self):
        return '{num_features}, eps={eps}, momentum={momentum}, affine={affine}, ' \
               'track_running_stats={track_running_stats}'.format(**self.__dict__)


class HardAttentionLayer(nn.Module):
    """utils.py:613-663.  S (b,c,h,w), sem_seg (b,1,h,w) 0/1, ins_seg (b,n,h,w) 0/1 instance masks
    -> (e_t_split (b,n,h,w): softmax of the attention logits within each instance, e_t_org (b,1,h,w))."""

    def __init__(self, d_model, d_k, d_h, random_thred=0.2, reduction=2):
        super(HardAttentionLayer, self).__init__()
        self.l1 = nn.Conv2d(d_model, d_k, 1, 1)
        self.l2 = nn.Linear(d_model, d_k, bias=False)
        self.attend_fc = nn.Sequential(
            nn.Tanh(),
            nn.Conv2d(d_k, 1, 3, 1, 1))
        self.bn = maskBN(1)

    def forward(self, S, sem_seg, ins_seg, h_t=None):
        b, n, h, w = ins_seg.size()
        S = F.avg_pool2d(S, kernel_size=3, stride=1, padding=1)
        e_t_org = self.l1(S)
        e_t_org = self.attend_fc(e_t_org)
        e_t_org = self.bn(e_t_org, sem_seg)
        e_t_org = F.avg_pool2d(e_t_org, kernel_size=3, stride=1, padding=1) * sem_seg
        mask = ins_seg.reshape(b, n, h * w)
        if mask.dtype not in (torch.float32, torch.uint8, torch.bool):
            mask = (mask != 0).view(torch.uint8)
        e_t_split = masked_softmax_hw(e_t_org.reshape(b, h * w), mask, None, True).view(b, n, h, w)
        return e_t_split, e_t_org


def make_position_encoding(xp, batch, length, n_units, f=10000.):
    """utils.py:332-344 (sinusoidal code, float32)"""
    assert (n_units % 2 == 0)
    position_block = xp.broadcast_to(xp.arange(length)[None, None, :], (batch, n_units // 2, length)).astype('f')
    unit_block = xp.broadcast_to(xp.arange(n_units // 2)[None, :, None], (batch, n_units // 2, length)).astype('f')
    rad_block = position_block / (f * 1.) ** (unit_block / (n_units // 2))
    sin_block = xp.sin(rad_block)
    cos_block = xp.cos(rad_block)
    return xp.concatenate([sin_block, cos_block], axis=1)


def position_encoding_2d(n_units, h, w, device=None):
    """reseg.py:132-137 ReSeg.set_position_encoding: (1, n_units, h, w), first half coded along h, second along w."""
    h_vec = np.tile(make_position_encoding(np, 1, h, n_units // 2, f=10000.)[:, :, :, np.newaxis], (1, 1, 1, w))
    w_vec = np.tile(make_position_encoding(np, 1, w, n_units // 2, f=10000.)[:, :, np.newaxis, :], (1, 1, h, 1))
    vec = torch.from_numpy(np.concatenate([h_vec, w_vec], axis=1))
    return vec.to(device) if device is not None else vec
