"""Attention family of the instance-embedding block, reference names and signatures
(/root/reference/code/lib/archs/modules/utils.py):

  ScaledDotProductAttention(temperature, attn_dropout=0.1)      utils.py:305-329
  MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=0.1)    utils.py:167-225
  PositionwiseFeedForward(d_in, d_hid, dropout=0.1)             utils.py:229-246
  DecoderLayer(d_model, d_inner, n_head, d_k, d_v, ...)         utils.py:138-164

The dense QK^T / softmax / PV contraction runs in the tcgen05/TMEM kernel of csrc/attention.cu
through the C-ABI; projections, LayerNorm and the feed-forward are plain PyTorch (library GEMMs).

Differences from the reference, all deliberate:
  * `attn` (the L_q x L_k probabilities the reference returns) is materialised only when the module
    is built / called with `return_attn=True`; otherwise None is returned in its place.  At L = 4096
    the tensor is 268 MB per head-batch of 4.
  * attention-probability dropout inside the fused kernel is not implemented: in training mode
    `attn_dropout` must be 0 (eval mode ignores dropout as PyTorch does).
  * masks are bool/uint8 with 1 = masked, exactly as masked_fill(mask, -inf) reads them; a mask of
    shape (rows, 1, L_k) is treated as a key mask and never expanded.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .pointwise import add_layer_norm, thin_linear


def _prep_mask(mask, BH, Lq, Lk, device):
    """-> (key_mask (rows, Lk) u8 or None, full_mask (rows, Lq, Lk) u8 or None, rows)."""
    if mask is None:
        return None, None, 1
    if mask.dim() != 3:
        raise ValueError("mask must be (rows, L_q or 1, L_k)")
    m = mask.to(device=device)
    if m.dtype != torch.uint8:
        m = m.to(torch.uint8) if m.dtype != torch.bool else m.view(torch.uint8)
    rows = m.shape[0]
    if BH % rows != 0:
        raise ValueError("mask rows (%d) must divide n_head*batch (%d)" % (rows, BH))
    if m.shape[2] != Lk:
        raise ValueError("mask last dim %d != L_k %d" % (m.shape[2], Lk))
    if m.shape[1] == 1:
        return m.reshape(rows, Lk).contiguous(), None, rows
    if m.shape[1] != Lq:
        raise ValueError("mask middle dim must be 1 or L_q")
    return None, m.contiguous(), rows


class _SdpaFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, q, k, v, key_mask, full_mask, n_mask_rows, temperature):
        lib = _lib.load()
        for t, nme in ((q, "q"), (k, "k"), (v, "v")):
            _lib.require_cuda(t, nme)
            if t.dtype != torch.float32:
                raise TypeError("attention computes in fp32 (bf16 hi/lo tensor-core products); got %s" % t.dtype)
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        BH, Lq, d = q.shape
        Lk, dv = k.shape[1], v.shape[2]
        if k.shape[0] != BH or v.shape[0] != BH or k.shape[2] != d or v.shape[1] != Lk:
            raise ValueError("inconsistent q/k/v shapes")
        out = torch.empty(BH, Lq, dv, device=q.device, dtype=torch.float32)
        lse2 = torch.empty(BH, Lq, device=q.device, dtype=torch.float32)
        wsb = lib.isa_attention_workspace_bytes(BH, Lq, Lk)
        ws = torch.empty(wsb, device=q.device, dtype=torch.uint8)
        rc = lib.isa_attention_fwd(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), BH, Lq, Lk, d, dv, float(temperature),
                                   _lib.ptr(key_mask), _lib.ptr(full_mask), int(n_mask_rows),
                                   _lib.ptr(out), _lib.ptr(lse2), _lib.ptr(ws), wsb, _lib.stream_ptr(q.device))
        _lib.check(rc, "isa_attention_fwd")
        ctx.save_for_backward(q, k, v, out, lse2, key_mask, full_mask)
        ctx.cfg = (BH, Lq, Lk, d, dv, float(temperature), int(n_mask_rows))
        ctx.mark_non_differentiable(lse2)
        return out, lse2

    @staticmethod
    def backward(ctx, dout, _dlse):
        lib = _lib.load()
        q, k, v, out, lse2, key_mask, full_mask = ctx.saved_tensors
        BH, Lq, Lk, d, dv, temperature, n_mask_rows = ctx.cfg
        dout = dout.contiguous().float()
        dq = torch.empty_like(q)
        dk = torch.empty_like(k)
        dvv = torch.empty_like(v)
        wsb = lib.isa_attention_workspace_bytes(BH, Lq, Lk)
        ws = torch.empty(wsb, device=q.device, dtype=torch.uint8)
        rc = lib.isa_attention_bwd(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(out), _lib.ptr(dout), _lib.ptr(lse2),
                                   BH, Lq, Lk, d, dv, temperature, _lib.ptr(key_mask), _lib.ptr(full_mask), n_mask_rows,
                                   _lib.ptr(dq), _lib.ptr(dk), _lib.ptr(dvv), _lib.ptr(ws), wsb, _lib.stream_ptr(q.device))
        _lib.check(rc, "isa_attention_bwd")
        return dq, dk, dvv, None, None, None, None


def attention_probs(q, k, lse2, temperature, key_mask=None, full_mask=None, n_mask_rows=1):
    """The (BH, L_q, L_k) softmax probabilities, recomputed from the saved log-sum-exp."""
    lib = _lib.load()
    BH, Lq, d = q.shape
    Lk = k.shape[1]
    attn = torch.empty(BH, Lq, Lk, device=q.device, dtype=torch.float32)
    rc = lib.isa_attention_probs(_lib.ptr(q.contiguous()), _lib.ptr(k.contiguous()), _lib.ptr(lse2), BH, Lq, Lk, d,
                                 float(temperature), _lib.ptr(key_mask), _lib.ptr(full_mask), int(n_mask_rows),
                                 _lib.ptr(attn), _lib.stream_ptr(q.device))
    _lib.check(rc, "isa_attention_probs")
    return attn


def scaled_dot_product_attention(q, k, v, temperature, mask=None, return_attn=False):
    """q (BH,Lq,d), k (BH,Lk,d), v (BH,Lk,dv) fp32 CUDA; mask as in the module docstring."""
    BH, Lq, _ = q.shape
    Lk = k.shape[1]
    key_mask, full_mask, rows = _prep_mask(mask, BH, Lq, Lk, q.device)
    out, lse2 = _SdpaFn.apply(q, k, v, key_mask, full_mask, rows, temperature)
    attn = None
    if return_attn:
        attn = attention_probs(q.detach(), k.detach(), lse2, temperature, key_mask, full_mask, rows)
    return out, attn


class ScaledDotProductAttention(nn.Module):
    ''' Scaled Dot-Product Attention (utils.py:305-329) '''

    def __init__(self, temperature, attn_dropout=0.1, return_attn=False):
        super().__init__()
        self.temperature = float(temperature)
        self.attn_dropout = float(attn_dropout)
        self.dropout = nn.Dropout(attn_dropout)
        self.return_attn = return_attn

    def forward(self, q, k, v, mask=None, last=False):
        if last:
            # utils.py:315-318: raw correlation, no scaling, no softmax
            return torch.bmm(q, k.transpose(1, 2))
        if self.training and self.attn_dropout > 0:
            raise NotImplementedError("attention-probability dropout is not fused into the sm_100a kernel: "
                                      "build the module with attn_dropout=0 for training (eval mode is unaffected)")
        return scaled_dot_product_attention(q, k, v, self.temperature, mask, self.return_attn)


class MultiHeadAttention(nn.Module):
    ''' Multi-Head Attention module (utils.py:167-225) '''

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1, attn_dropout=None, return_attn=False):
        super().__init__()
        self.n_head = n_head
        self.d_k = d_k
        self.d_v = d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k)
        self.w_ks = nn.Linear(d_model, n_head * d_k)
        self.w_vs = nn.Linear(d_model, n_head * d_v)
        nn.init.normal_(self.w_qs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_ks.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_vs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_v)))
        # the reference's inner module always has attn_dropout=0.1 (utils.py:184 + :308)
        self.attention = ScaledDotProductAttention(temperature=np.power(d_k, 0.5),
                                                   attn_dropout=0.1 if attn_dropout is None else attn_dropout,
                                                   return_attn=return_attn)
        self.layer_norm = nn.LayerNorm(d_model)
        self.fc = nn.Linear(n_head * d_v, d_model)
        nn.init.xavier_normal_(self.fc.weight)
        self.dropout = nn.Dropout(dropout)

    def forward(self, q, k, v, mask=None, last=False):
        d_k, d_v, n_head = self.d_k, self.d_v, self.n_head
        sz_b, len_q, _ = q.size()
        sz_b, len_k, _ = k.size()
        sz_b, len_v, _ = v.size()
        residual = q
        q = thin_linear(q, self.w_qs).view(sz_b, len_q, n_head, d_k)
        k = thin_linear(k, self.w_ks).view(sz_b, len_k, n_head, d_k)
        v = thin_linear(v, self.w_vs).view(sz_b, len_v, n_head, d_v)
        q = q.permute(2, 0, 1, 3).contiguous().view(-1, len_q, d_k)  # (n*b) x lq x dk
        k = k.permute(2, 0, 1, 3).contiguous().view(-1, len_k, d_k)
        v = v.permute(2, 0, 1, 3).contiguous().view(-1, len_v, d_v)
        # the reference repeats the mask n_head times (utils.py:210-211); the kernel indexes it
        # modulo the batch instead, so no copy is made.
        if not last:
            output, attn = self.attention(q, k, v, mask=mask)
            output = output.view(n_head, sz_b, len_q, d_v)
            output = output.permute(1, 2, 0, 3).contiguous().view(sz_b, len_q, -1)  # b x lq x (n*dv)
            output = self.dropout(thin_linear(output, self.fc))
            output = add_layer_norm(output, residual, self.layer_norm)
            return output, attn
        correlation = self.attention(q, k, v, mask=None, last=True)
        correlation = torch.sigmoid(correlation)
        correlation = correlation.squeeze(1)
        return correlation, None


class PositionwiseFeedForward(nn.Module):
    ''' A two-feed-forward-layer module (utils.py:229-246) '''

    def __init__(self, d_in, d_hid, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, 1)
        self.w_2 = nn.Conv1d(d_hid, d_in, 1)
        self.layer_norm = nn.LayerNorm(d_in)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        residual = x
        output = x.transpose(1, 2)
        output = self.w_2(F.relu(self.w_1(output)))
        output = output.transpose(1, 2)
        output = self.dropout(output)
        output = self.layer_norm(output + residual)
        return output


class DecoderLayer(nn.Module):
    ''' Compose with three layers (utils.py:138-164) '''

    def __init__(self, d_model, d_inner, n_head, d_k, d_v, dropout=0.1, last=False, attn_dropout=None):
        super(DecoderLayer, self).__init__()
        self.last = last
        if last:
            n_head = 1
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout, attn_dropout=attn_dropout)
        self.enc_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout, attn_dropout=attn_dropout)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner, dropout=dropout)

    def forward(self, dec_input, enc_output, mask):
        mask = mask.unsqueeze(1).byte()
        slf_attn_mask = 1 - mask    # 1 = masked = outside the foreground
        dec_output, dec_slf_attn = self.slf_attn(dec_input, dec_input, dec_input, mask=None)
        dec_output, dec_enc_attn = self.enc_attn(dec_output, enc_output, enc_output, mask=slf_attn_mask, last=self.last)
        if not self.last:
            dec_output = self.pos_ffn(dec_output)
        return dec_output, dec_slf_attn, dec_enc_attn
