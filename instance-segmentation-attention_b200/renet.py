"""ReNet layer: a bidirectional GRU swept over every row, then over every column.

Reference contract: /root/reference/code/lib/archs/modules/README.md:225-256
    ReNet(n_input, n_units, patch_size=(1, 1), use_coordinates=False, usegpu=True)
    (N, C_in, H, W) -> (N, 2 * n_units, H / ph, W / pw)
(`renet.py` is absent from the reference tree -- reseg.py:5 has the import commented out -- so
the arithmetic is PyTorch's nn.GRU, which is also the oracle: oracle/renet_ref.py.)

The parameters live in two real `nn.GRU` modules named like upstream (`rnn_hor`, `rnn_ver`), so
state_dicts interchange (`rnn_hor.weight_ih_l0`, `..._reverse`, ...); they are never *called*:
the input projection is one GEMM over all tokens and the recurrence runs in the persistent
sm_100a scan kernel (csrc/gru_scan.cu) through the C-ABI.  Internally everything is token-major
(channels-last), so the returned tensor is a channels_last view -- no transposes between the
row sweep, the column sweep and the following convolution.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


def _geometry(B, H, W, mode):
    """(n_seq, T, inner, outer_tok_stride, inner_tok_stride, t_tok_stride) for tokens (b*H+h)*W+w."""
    if mode == "hor":   # sequences = rows, steps along w
        return B * H, W, B * H, 0, W, 1
    return B * W, H, W, H * W, 1, W  # sequences = columns, steps along h


def _split3(src2d, dst, col0, order, shift=0, pos=(1, 1, 0)):
    """fp32 (rows, cols) view with unit column stride -> the three bf16 parts at columns [col0, col0+cols) of
    dst (rows, 3, ctot) (csrc/split_bf16.cu); `shift`/`pos` select a row-shifted, zero-filled read."""
    lib = _lib.load()
    rows, cols = src2d.shape
    assert src2d.stride(1) == 1 and dst.dtype == torch.bfloat16 and dst.is_contiguous()
    ctot = dst.shape[2]
    rc = lib.isa_split_bf16x3(src2d.data_ptr(), rows, cols, src2d.stride(0), dst.data_ptr() + 2 * col0, 3 * ctot, ctot,
                              order, shift, pos[0], pos[1], pos[2], _lib.stream_ptr(src2d.device))
    _lib.check(rc, "isa_split_bf16x3")


def _tensor_core_ok(*dims):
    """The bf16x3 GEMM path needs 4-column granularity (8 B bf16 stores, 16 B fp32 loads)."""
    return all(d % 4 == 0 for d in dims)


class _GruSweepFn(torch.autograd.Function):
    """out[tokens, 2n] = biGRU over the sequences described by `mode` of x[tokens, Cin].

    Projections: x W_ih^T, dG W_ih and dG^T [x | h_prev | 1] run as bf16 tensor-core GEMMs over hi/lo-split
    operands (fp32-level accuracy, see csrc/split_bf16.cu); shapes whose widths are not multiples of 4 take
    the plain fp32 library GEMM."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, B, H, W, mode):
        lib = _lib.load()
        _lib.require_cuda(x, "x")
        n = w_hh.shape[2]
        tokens = B * H * W
        x2 = x.reshape(tokens, -1)
        cin = x2.shape[1]
        w2 = w_ih.reshape(6 * n, cin)
        fast = _tensor_core_ok(cin, n)
        if fast:
            # input projection of both directions in one GEMM: [tokens][2][3n]; b_ih is added by the scan kernel
            xs = torch.empty(tokens, 3, cin, device=x.device, dtype=torch.bfloat16)
            ws = torch.empty(6 * n, 3, cin, device=x.device, dtype=torch.bfloat16)
            _split3(x2, xs, 0, 0)
            _split3(w2, ws, 0, 1)
            gx = torch.mm(xs.view(tokens, 3 * cin), ws.view(6 * n, 3 * cin).t(), out_dtype=torch.float32)
            bias_in = b_ih.contiguous()
            del xs
        else:
            gx = torch.addmm(b_ih.reshape(-1), x2, w2.t())
            bias_in = None
        out = torch.empty(tokens, 2 * n, device=x.device, dtype=torch.float32)
        need_grad = any(ctx.needs_input_grad)
        stash = torch.empty(tokens, 2, 4 * n, device=x.device, dtype=torch.float32) if need_grad else None
        n_seq, T, inner, outer, inner_s, t_s = _geometry(B, H, W, mode)
        w_hh_c = w_hh.contiguous()
        rc = lib.isa_gru_scan_fwd(_lib.ptr(gx), _lib.ptr(w_hh_c), _lib.ptr(b_hh.contiguous()), _lib.ptr(bias_in), n_seq, T, n,
                                  inner, outer, inner_s, t_s, _lib.ptr(out), _lib.ptr(stash), _lib.stream_ptr(x.device))
        _lib.check(rc, "isa_gru_scan_fwd")
        if need_grad:
            ctx.save_for_backward(x2, w_ih, w_hh_c, out, stash)
            ctx.geom = (B, H, W, mode, n)
            ctx.x_shape = tuple(x.shape)
            ctx.fast = fast
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        x2, w_ih, w_hh, out, stash = ctx.saved_tensors
        B, H, W, mode, n = ctx.geom
        tokens = B * H * W
        cin = x2.shape[1]
        dout = dout.contiguous()
        dgx = torch.empty(tokens, 2, 3 * n, device=dout.device, dtype=torch.float32)
        dghn = torch.empty(tokens, 2, n, device=dout.device, dtype=torch.float32)
        n_seq, T, inner, outer, inner_s, t_s = _geometry(B, H, W, mode)
        rc = lib.isa_gru_scan_bwd(_lib.ptr(dout), _lib.ptr(out), _lib.ptr(stash), _lib.ptr(w_hh), n_seq, T, n,
                                  inner, outer, inner_s, t_s, _lib.ptr(dgx), _lib.ptr(dghn), _lib.stream_ptr(dout.device))
        _lib.check(rc, "isa_gru_scan_bwd")
        dg2 = dgx.view(tokens, 6 * n)
        w2 = w_ih.reshape(6 * n, cin)
        if ctx.fast:
            return _GruSweepFn._backward_tensor_core(ctx, x2, w2, w_ih, w_hh, out, dg2, dghn.view(tokens, 2 * n))
        dx = dg2 @ w2 if ctx.needs_input_grad[0] else None
        dw_ih = (dg2.t() @ x2).view_as(w_ih)
        db_ih = dg2.sum(0).view(2, 3 * n)
        # h_{t-1} per direction: the forward output shifted by one step along the sweep axis
        o5 = out.view(B, H, W, 2, n)
        hprev = torch.zeros_like(o5)
        if mode == "hor":
            hprev[:, :, 1:, 0] = o5[:, :, :-1, 0]
            hprev[:, :, :-1, 1] = o5[:, :, 1:, 1]
        else:
            hprev[:, 1:, :, 0] = o5[:, :-1, :, 0]
            hprev[:, :-1, :, 1] = o5[:, 1:, :, 1]
        hprev = hprev.view(tokens, 2, n)
        dw_hh = torch.empty_like(w_hh)
        db_hh = torch.empty(2, 3 * n, device=dout.device, dtype=torch.float32)
        for d in range(2):
            hp = hprev[:, d]
            g_rz = dgx[:, d, :2 * n]
            g_n = dghn[:, d]
            dw_hh[d, :2 * n] = g_rz.t() @ hp
            dw_hh[d, 2 * n:] = g_n.t() @ hp
            db_hh[d, :2 * n] = g_rz.sum(0)
            db_hh[d, 2 * n:] = g_n.sum(0)
        return dx.view(ctx.x_shape) if dx is not None else None, dw_ih, dw_hh, db_ih, db_hh, None, None, None, None

    @staticmethod
    def _backward_tensor_core(ctx, x2, w2, w_ih, w_hh, out, dg2, dghn2):
        """Two bf16x3 GEMM groups over ONE split of the gate gradients:
             left  L = [dgx (6n) | dghn (2n)]                      (tokens, 3, 8n)   parts (hi, hi, lo)
             right R = [x (cin) | h_prev dir0 (n) | h_prev dir1 (n) | 1 (8)]  (tokens, 3, cin+2n+8)  parts (hi, lo, hi)
           dx    = L_kconcat (tokens, 24n) @ W_pad (24n, cin)      (rows of W_pad under the dghn columns are zero)
           C     = sum_parts L_p^T R_p  (8n, cin+2n+8): dW_ih, dW_hh blocks and, from the ones column, every bias gradient."""
        B, H, W, mode, n = ctx.geom
        tokens = B * H * W
        cin = x2.shape[1]
        dev = x2.device
        nl = 8 * n
        nr = cin + 2 * n + 8
        L = torch.empty(tokens, 3, nl, device=dev, dtype=torch.bfloat16)
        _split3(dg2, L, 0, 0)
        _split3(dghn2, L, 6 * n, 0)
        dx = None
        if ctx.needs_input_grad[0]:
            wp = torch.zeros(3, nl, cin, device=dev, dtype=torch.bfloat16)
            # part q of W (order 1) under part q of L, rows [0, 6n) of each 8n block
            wtmp = torch.empty(6 * n, 3, cin, device=dev, dtype=torch.bfloat16)
            _split3(w2, wtmp, 0, 1)
            wp[:, :6 * n] = wtmp.permute(1, 0, 2)
            dx = torch.mm(L.view(tokens, 3 * nl), wp.view(3 * nl, cin), out_dtype=torch.float32).view(ctx.x_shape)
        R = torch.empty(tokens, 3, nr, device=dev, dtype=torch.bfloat16)
        _split3(x2, R, 0, 1)
        o2 = out.view(tokens, 2 * n)
        if mode == "hor":   # step axis = w (token stride 1), position = token % W
            step, pos_div, pos_mod = 1, 1, W
        else:               # step axis = h (token stride W), position = (token / W) % H
            step, pos_div, pos_mod = W, W, H
        # direction 0 walks forward: h_{t-1} is the previous position; direction 1 walks backward: the next one
        _split3(o2[:, :n], R, cin, 1, shift=-step, pos=(pos_div, pos_mod, -1))
        _split3(o2[:, n:], R, cin + n, 1, shift=step, pos=(pos_div, pos_mod, 1))
        ones = R[:, :, cin + 2 * n:]
        ones.zero_()
        ones[:, 0, 0] = 1.0
        ones[:, 2, 0] = 1.0
        C = (torch.mm(L[:, 0].t(), R[:, 0], out_dtype=torch.float32)
             + torch.mm(L[:, 1].t(), R[:, 1], out_dtype=torch.float32)
             + torch.mm(L[:, 2].t(), R[:, 2], out_dtype=torch.float32))
        kb = cin + 2 * n   # the ones column
        dw_ih = C[:6 * n, :cin].reshape(w_ih.shape)
        db_ih = C[:6 * n, kb].reshape(2, 3 * n)
        dw_hh = torch.empty_like(w_hh)
        db_hh = torch.empty(2, 3 * n, device=dev, dtype=torch.float32)
        for d in range(2):
            cols = slice(cin + d * n, cin + (d + 1) * n)
            dw_hh[d, :2 * n] = C[3 * n * d:3 * n * d + 2 * n, cols]
            dw_hh[d, 2 * n:] = C[6 * n + d * n:6 * n + (d + 1) * n, cols]
            db_hh[d, :2 * n] = C[3 * n * d:3 * n * d + 2 * n, kb]
            db_hh[d, 2 * n:] = C[6 * n + d * n:6 * n + (d + 1) * n, kb]
        return dx, dw_ih, dw_hh, db_ih, db_hh, None, None, None, None


def _stack_gru(gru):
    w_ih = torch.stack([gru.weight_ih_l0, gru.weight_ih_l0_reverse])
    w_hh = torch.stack([gru.weight_hh_l0, gru.weight_hh_l0_reverse])
    b_ih = torch.stack([gru.bias_ih_l0, gru.bias_ih_l0_reverse])
    b_hh = torch.stack([gru.bias_hh_l0, gru.bias_hh_l0_reverse])
    return w_ih, w_hh, b_ih, b_hh


def bigru_sweep(x_nhwc, gru, mode):
    """x_nhwc (B,H,W,Cin) contiguous float32 CUDA -> (B,H,W,2n)."""
    B, H, W, _ = x_nhwc.shape
    w_ih, w_hh, b_ih, b_hh = _stack_gru(gru)
    out = _GruSweepFn.apply(x_nhwc, w_ih, w_hh, b_ih, b_hh, B, H, W, mode)
    return out.view(B, H, W, -1)


class ReNet(nn.Module):
    """See module docstring.  `use_coordinates` is accepted for signature parity; it is False in
    every shipped setting (settings/CVPPP/model_settings.py:18) and not implemented here."""

    def __init__(self, n_input, n_units, patch_size=(1, 1), use_coordinates=False, usegpu=True):
        super(ReNet, self).__init__()
        if use_coordinates:
            raise NotImplementedError("ReNet(use_coordinates=True) is outside the hot path (off in every shipped setting)")
        self.patch_size_height = int(patch_size[0])
        self.patch_size_width = int(patch_size[1])
        assert self.patch_size_height >= 1 and self.patch_size_width >= 1
        self.tiling = not (self.patch_size_height == 1 and self.patch_size_width == 1)
        self.n_units = n_units
        self.usegpu = usegpu
        rnn_hor_n_inputs = n_input * self.patch_size_height * self.patch_size_width
        self.rnn_hor = nn.GRU(rnn_hor_n_inputs, n_units, num_layers=1, batch_first=True, bidirectional=True)
        self.rnn_ver = nn.GRU(n_units * 2, n_units, num_layers=1, batch_first=True, bidirectional=True)

    def tile(self, x):
        """(N,C,H,W) -> (N, C*ph*pw, H/ph, W/pw): zero-pad to a multiple of the patch, fold patches into channels."""
        ph, pw = self.patch_size_height, self.patch_size_width
        n_h_pad = (ph - x.size(2) % ph) % ph
        n_w_pad = (pw - x.size(3) % pw) % pw
        if n_h_pad or n_w_pad:
            x = F.pad(x, (n_w_pad // 2, n_w_pad - n_w_pad // 2, n_h_pad // 2, n_h_pad - n_h_pad // 2))
        b, c, h, w = x.size()
        x = x.view(b, c, h // ph, ph, w // pw, pw).permute(0, 2, 4, 1, 3, 5)
        return x.contiguous().view(b, h // ph, w // pw, ph * pw * c)  # NHWC

    def forward(self, x):
        if self.tiling:
            x = self.tile(x)
        else:
            x = x.permute(0, 2, 3, 1).contiguous()  # no copy if x is channels_last
        x = x.float()
        x = bigru_sweep(x, self.rnn_hor, "hor")
        x = bigru_sweep(x, self.rnn_ver, "ver")
        return x.permute(0, 3, 1, 2)  # (N, 2n, H, W), channels_last memory
