"""Drop-in `DiscriminativeLoss` backed by the sm_100a segmented-reduction kernels.

Same constructor, call signature and return value as the reference class
(/root/reference/code/lib/losses/discriminative.py:191-213):

    crit = DiscriminativeLoss(delta_var, delta_dist, norm, usegpu=True)
    loss, cluster_means = crit(input, target, n_objects, max_n_objects)

input  (bs, C, H, W) float32 CUDA
target (bs, K, H, W) float / int64 / uint8 one-hot (or soft) masks as the reference's
       collate emits them (dataset.py:354-376), or -- extension -- a uint8 label map
       (bs, H, W) with 255 = background (1 byte per pixel instead of 4K..8K)
n_objects (bs,) ints;  max_n_objects == K.
`q_denominator` (optional 1-element float CUDA tensor) overrides int(sum(target)) in the q-regulariser
(discriminative.py:153-159) -- data parallel passes global_foreground / world_size.

The shipped composite is  1.0 * variance + 0.005 * q_regulariser  with L2-normalised
means (discriminative.py:168-186).  `terms=` switches the distance / regulariser terms
on (they are defined in the reference, :98-147, but not summed).
"""
import torch
from torch.nn.modules.loss import _Loss

from . import _lib

TGT_LABEL_U8, TGT_DENSE_F32, TGT_DENSE_I64, TGT_DENSE_U8 = 0, 1, 2, 3

# weights of (variance, distance, regulariser, q_regulariser)
SHIPPED_TERMS = (1.0, 0.0, 0.0, 0.005)


def _target_kind(target, input):
    if target.dim() == 3:
        if target.dtype != torch.uint8:
            raise TypeError("a (bs,H,W) label-map target must be uint8 (255 = background)")
        return TGT_LABEL_U8
    if target.dim() != 4:
        raise ValueError("target must be (bs,K,H,W) masks or a (bs,H,W) uint8 label map")
    if target.dtype == torch.float32:
        return TGT_DENSE_F32
    if target.dtype == torch.int64:
        return TGT_DENSE_I64
    if target.dtype in (torch.uint8, torch.bool):
        return TGT_DENSE_U8
    raise TypeError("unsupported target dtype %s" % target.dtype)


class _DiscLossFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, input, target, n_objects, K, delta_v, delta_d, norm, normalize, terms, q_den=None):
        lib = _lib.load()
        _lib.require_cuda(input, "input")
        _lib.require_cuda(target, "target")
        if input.dtype != torch.float32:
            raise TypeError("DiscriminativeLoss computes in fp32; got %s" % input.dtype)
        bs, C, H, W = input.shape
        x = input.contiguous()
        kind = _target_kind(target, input)
        tgt = target.contiguous()
        if tgt.dtype == torch.bool:
            tgt = tgt.view(torch.uint8)
        nobj = torch.as_tensor(n_objects).reshape(-1).to(device=x.device, dtype=torch.int32).contiguous()
        if nobj.numel() != bs:
            raise ValueError("n_objects must have one entry per image")
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        terms_out = torch.empty(4, device=x.device, dtype=torch.float32)  # var, dist, reg, qreg
        means = torch.empty(bs, K, C, device=x.device, dtype=torch.float32)
        ws_bytes = lib.isa_disc_loss_workspace_bytes(bs, C, K, H, W)
        ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
        rc = lib.isa_disc_loss_fwd(
            _lib.ptr(x), _lib.ptr(tgt), kind, _lib.ptr(nobj), bs, C, H, W, K,
            delta_v, delta_d, norm, int(normalize), terms[0], terms[1], terms[2], terms[3],
            _lib.ptr(q_den), loss.data_ptr(), terms_out.data_ptr(), _lib.ptr(means),
            _lib.ptr(ws), ws_bytes, _lib.stream_ptr(x.device))
        _lib.check(rc, "isa_disc_loss_fwd")
        ctx.save_for_backward(x, tgt, nobj, means, ws, q_den)
        ctx.cfg = (kind, bs, C, H, W, K, delta_v, delta_d, norm, int(normalize), terms)
        ctx.mark_non_differentiable(terms_out)
        return loss, means, terms_out

    @staticmethod
    def backward(ctx, grad_loss, grad_means, _grad_terms):
        lib = _lib.load()
        x, tgt, nobj, means, ws, q_den = ctx.saved_tensors
        kind, bs, C, H, W, K, delta_v, delta_d, norm, normalize, terms = ctx.cfg
        if grad_loss is None:
            grad_loss = torch.zeros((), device=x.device, dtype=torch.float32)
        gl = grad_loss.reshape(1).to(torch.float32).contiguous()
        gm = None
        if grad_means is not None:
            gm = grad_means.to(torch.float32).contiguous()
        grad = torch.empty_like(x)
        rc = lib.isa_disc_loss_bwd(
            _lib.ptr(x), _lib.ptr(tgt), kind, _lib.ptr(nobj), bs, C, H, W, K,
            delta_v, delta_d, norm, normalize, terms[0], terms[1], terms[2], terms[3],
            _lib.ptr(q_den), _lib.ptr(means), _lib.ptr(gl), _lib.ptr(gm), _lib.ptr(grad),
            _lib.ptr(ws), ws.numel(), _lib.stream_ptr(x.device))
        _lib.check(rc, "isa_disc_loss_bwd")
        return grad, None, None, None, None, None, None, None, None, None


class DiscriminativeLoss(_Loss):
    """See module docstring; reference: lib/losses/discriminative.py:191-213."""

    def __init__(self, delta_var, delta_dist, norm, size_average=True, reduce=True, usegpu=True,
                 terms=SHIPPED_TERMS, normalize_means=True):
        super(DiscriminativeLoss, self).__init__()
        self.reduce = reduce
        self.delta_var = float(delta_var)
        self.delta_dist = float(delta_dist)
        self.norm = int(norm)
        self.usegpu = usegpu
        self.terms = tuple(float(t) for t in terms)
        self.normalize_means = bool(normalize_means)
        self.last_terms = None
        assert self.norm in [1, 2]
        assert len(self.terms) == 4
        if not usegpu:
            raise _lib.IsaError("DiscriminativeLoss(usegpu=False): this build has no CPU path "
                                "(the CPU restatement lives in oracle/ and is test-only)")

    def forward(self, input, target, n_objects, max_n_objects, q_denominator=None):
        K = int(max_n_objects)
        if target.dim() == 4 and target.size(1) != K:
            raise ValueError("target has %d instance channels but max_n_objects=%d "
                             "(the reference broadcast fails on this too)" % (target.size(1), K))
        loss, means, terms = _DiscLossFn.apply(input, target, n_objects, K, self.delta_var, self.delta_dist,
                                               self.norm, self.normalize_means, self.terms, q_denominator)
        self.last_terms = terms  # (var, dist, reg, qreg), unweighted, detached
        return loss, means


def onehot_to_labels(target, with_count=False):
    """(bs,K,H,W) one-hot masks -> ((bs,H,W) uint8 label map, device int flag `not_onehot`[, device int64 fg count])."""
    lib = _lib.load()
    _lib.require_cuda(target, "target")
    kind = _target_kind(target, None)
    if kind == TGT_LABEL_U8:
        raise ValueError("target is already a label map")
    tgt = target.contiguous()
    if tgt.dtype == torch.bool:
        tgt = tgt.view(torch.uint8)
    bs, K, H, W = tgt.shape
    labels = torch.empty(bs, H, W, device=tgt.device, dtype=torch.uint8)
    flag = torch.empty(1, device=tgt.device, dtype=torch.int32)
    count = torch.empty(1, device=tgt.device, dtype=torch.int64) if with_count else None
    rc = lib.isa_onehot_to_labels(_lib.ptr(tgt), kind, bs, K, H, W, _lib.ptr(labels), _lib.ptr(flag), _lib.ptr(count),
                                  _lib.stream_ptr(tgt.device))
    _lib.check(rc, "isa_onehot_to_labels")
    return (labels, flag, count) if with_count else (labels, flag)


def label_fg_count(labels, K):
    """Foreground pixels (label < K) of a uint8 label map -> 1-element device int64 tensor."""
    lib = _lib.load()
    _lib.require_cuda(labels, "labels")
    lab = labels.contiguous()
    count = torch.empty(1, device=lab.device, dtype=torch.int64)
    rc = lib.isa_label_fg_count(_lib.ptr(lab), lab.numel(), int(K), _lib.ptr(count), _lib.stream_ptr(lab.device))
    _lib.check(rc, "isa_label_fg_count")
    return count
