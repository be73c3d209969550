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


def _tensor_core_ok(*dims):
    """The fused-split tcgen05 GEMMs read 16-byte aligned fp32 runs: widths must be multiples of 4."""
    return all(d % 4 == 0 for d in dims)


# useful FLOPs of the three projection GEMMs of one sweep (bench.py rooflines): 2 * M * N * K
def bench_gemm_flops(tokens, n=100):
    avg_cin = (256 + 200 + 200 + 200) / 4.0      # row sweep of layer 1 reads the 256-channel map, the others 2n = 200
    return {"isa_renet_proj_fwd": 2.0 * tokens * 6 * n * avg_cin,
            "isa_renet_proj_dx": 2.0 * tokens * 6 * n * avg_cin,
            "isa_renet_proj_wgrad": 2.0 * tokens * (6 * n * avg_cin + 6 * n * n)}


class _GruSweepFn(torch.autograd.Function):
    """out[tokens, 2n] = biGRU over the sequences described by `mode` of x[tokens, Cin].

    Projections: x W_ih^T, dG W_ih and [dgx | dghn]^T [x | h_prev | 1] run on tcgen05 (csrc/proj_gemm.cu): the fp32
    operands are split into bf16 hi/lo parts inside the kernel's producer warps (fp32-level accuracy), h_prev is read
    from the forward output with a row shift, nothing is transposed, copied or concatenated in global memory.  Shapes
    whose widths are not multiples of 4 take the plain fp32 library GEMM."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, B, H, W, mode):
        lib = _lib.load()
        _lib.require_cuda(x, "x")
        n = w_hh.shape[2]
        tokens = B * H * W
        x2 = x.reshape(tokens, -1)
        cin = x2.shape[1]
        w2 = w_ih.reshape(6 * n, cin).contiguous()
        fast = _tensor_core_ok(cin, n)
        if fast:
            # input projection of both directions in one GEMM: [tokens][2][3n]; b_ih is added by the scan kernel
            x2 = x2.contiguous()
            gx = torch.empty(tokens, 6 * n, device=x.device, dtype=torch.float32)
            wsb = lib.isa_renet_proj_workspace_bytes(cin, 6 * n)
            wpk = torch.empty(wsb, device=x.device, dtype=torch.uint8)
            rc = lib.isa_renet_proj_fwd(_lib.ptr(x2), _lib.ptr(w2), tokens, cin, 6 * n, _lib.ptr(gx), _lib.ptr(wpk), wsb,
                                        _lib.stream_ptr(x.device))
            _lib.check(rc, "isa_renet_proj_fwd")
            bias_in = b_ih.contiguous()
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
            return _GruSweepFn._backward_tensor_core(ctx, x2, w2.contiguous(), w_ih, w_hh, out, dgx, dghn)
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
    def _backward_tensor_core(ctx, x2, w2, w_ih, w_hh, out, dgx, dghn):
        """dx = dgx W_ih (one GEMM) and every weight / bias gradient from ONE product over all tokens,
        [dgx | dghn]^T [x | h_prev(dir 0) | h_prev(dir 1) | 1], split over the tokens and folded in a fixed order."""
        lib = _lib.load()
        B, H, W, mode, n = ctx.geom
        tokens = B * H * W
        cin = x2.shape[1]
        dev = x2.device
        st = _lib.stream_ptr(dev)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(tokens, cin, device=dev, dtype=torch.float32)
            wsb = lib.isa_renet_proj_workspace_bytes(cin, 6 * n)
            wpk = torch.empty(wsb, device=dev, dtype=torch.uint8)
            rc = lib.isa_renet_proj_dx(_lib.ptr(dgx), _lib.ptr(w2), tokens, 6 * n, cin, _lib.ptr(dx), _lib.ptr(wpk), wsb, st)
            _lib.check(rc, "isa_renet_proj_dx")
            dx = dx.view(ctx.x_shape)
        if mode == "hor":   # step axis = w (token stride 1), position = token % W
            step, pos_div, pos_mod = 1, 1, W
        else:               # step axis = h (token stride W), position = (token / W) % H
            step, pos_div, pos_mod = W, W, H
        dw_ih = torch.empty_like(w_ih)
        dw_hh = torch.empty_like(w_hh)
        db_ih = torch.empty(2, 3 * n, device=dev, dtype=torch.float32)
        db_hh = torch.empty(2, 3 * n, device=dev, dtype=torch.float32)
        wsb = lib.isa_renet_proj_wgrad_workspace_bytes(tokens, cin, n)
        ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
        rc = lib.isa_renet_proj_wgrad(_lib.ptr(dgx), _lib.ptr(dghn), _lib.ptr(x2), _lib.ptr(out), tokens, cin, n, step, pos_div, pos_mod,
                                      _lib.ptr(dw_ih), _lib.ptr(dw_hh), _lib.ptr(db_ih), _lib.ptr(db_hh), _lib.ptr(ws), wsb, st)
        _lib.check(rc, "isa_renet_proj_wgrad")
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
