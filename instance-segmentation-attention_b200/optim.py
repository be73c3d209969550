"""Flat-buffer optimizer for the training step: gradient-norm clipping + Adadelta in two kernel launches
(csrc/pointwise.cu: isa_adadelta_step) instead of PyTorch's multi-tensor chains.

The reference builds `optim.Adadelta(parameters, lr, weight_decay)` (/root/reference/code/lib/model.py:141-160) and
calls `clip_grad_norm_` then `optimizer.step()` every mini-batch (model.py:271-281).  `FusedAdadelta` keeps that
interface (a `torch.optim.Optimizer` with one param group, so `ReduceLROnPlateau` drives `param_groups[0]['lr']`) and
the same arithmetic; parameters, gradients and both state tensors live in flat fp32 buffers that the module's
parameters / `.grad`s are views of (the same buffer is what the data-parallel all-reduce sends)."""
import torch

from . import _lib


class FusedAdadelta(torch.optim.Optimizer):

    def __init__(self, params, lr=1.0, rho=0.9, eps=1e-6, weight_decay=0.0):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FusedAdadelta: no trainable parameters")
        super(FusedAdadelta, self).__init__(params, dict(lr=lr, rho=rho, eps=eps, weight_decay=weight_decay))
        dev = params[0].device
        for p in params:
            _lib.require_cuda(p, "parameter")
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("FusedAdadelta: fp32 parameters on one device only")
        # every parameter starts on a 16-byte boundary of the flat buffers (vectorised kernels, TMA-friendly views)
        self._offsets = []
        off = 0
        for p in params:
            self._offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.flat_param = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(off, device=dev, dtype=torch.float32)
        self.square_avg = torch.zeros(off, device=dev, dtype=torch.float32)
        self.acc_delta = torch.zeros(off, device=dev, dtype=torch.float32)
        self.grad_norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self._params = params
        for p, o in zip(params, self._offsets):
            v = self.flat_param[o:o + p.numel()].view_as(p)
            v.copy_(p.data)
            p.data = v
            g = self.flat_grad[o:o + p.numel()].view_as(p)
            if p.grad is not None:
                g.copy_(p.grad)
            p.grad = g
        lib = _lib.load()
        self._wsb = lib.isa_adadelta_workspace_bytes()
        self._ws = torch.empty(self._wsb, device=dev, dtype=torch.uint8)

    def zero_grad(self, set_to_none=False):
        """One memset; the .grad views stay attached (set_to_none is ignored on purpose)."""
        self.flat_grad.zero_()
        for p, o in zip(self._params, self._offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)

    @torch.no_grad()
    def step(self, closure=None, clip_grad_norm=0.0):
        """clip_grad_norm > 0: scale the gradient by min(1, clip / (||g|| + 1e-6)) first (clip_grad_norm_ semantics).
        The pre-clipping norm is left in `self.grad_norm` (device tensor, no synchronisation)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        g = self.param_groups[0]
        rc = lib.isa_adadelta_step(self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.square_avg.data_ptr(),
                                   self.acc_delta.data_ptr(), self.numel, float(g['lr']), float(g['rho']), float(g['eps']),
                                   float(g['weight_decay']), float(clip_grad_norm or 0.0), self.grad_norm.data_ptr(),
                                   self._ws.data_ptr(), self._wsb, _lib.stream_ptr(self.flat_param.device))
        _lib.check(rc, "isa_adadelta_step")
        return loss

    def state_dict(self):
        return {"param_groups": [{k: v for k, v in self.param_groups[0].items() if k != 'params'}],
                "square_avg": self.square_avg.clone(), "acc_delta": self.acc_delta.clone()}

    def load_state_dict(self, sd):
        for k, v in sd["param_groups"][0].items():
            self.param_groups[0][k] = v
        self.square_avg.copy_(sd["square_avg"])
        self.acc_delta.copy_(sd["acc_delta"])
