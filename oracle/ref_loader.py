"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Loads the *unmodified* reference source files from /root/reference (read-only,
present only in the build container, NOT on the GPU box) so that
  * tests/golden/make_golden.py can generate golden vectors from the reference itself,
  * `-m "not gpu"` tests can pin the oracle restatements in this directory against it.

Nothing is copied: modules are executed from where they lie, behind the shim set
of SURVEY.md Appendix B (the reference targets Python 3.6 / torch 0.4):
  * sys.path entries mirroring code/train.py:3-6, so `config`, `MobileNetDenseASPP`
    resolve as top-level names;
  * `Tensor.cuda()` / `Module.cuda()` identity when no GPU is visible (the module-level
    `.cuda()` at code/lib/archs/modules/utils.py:11-12 would otherwise fail at import);
  * `Tensor.masked_fill` accepts uint8 masks again (utils.py:294,323,507,648).
"""
import importlib.util
import os
import sys

REF_ROOT = os.environ.get("ISA_REFERENCE_ROOT", "/root/reference")
REF_CODE = os.path.join(REF_ROOT, "code")

_loaded = {}
_shimmed = False


def available():
    return os.path.isfile(os.path.join(REF_CODE, "lib", "losses", "discriminative.py"))


def _install_shims():
    global _shimmed
    if _shimmed:
        return
    import torch
    for sub in ("lib/archs/modules", "lib/archs", "lib/losses", "lib", "settings/CVPPP", ""):
        p = os.path.join(REF_CODE, sub)
        if p not in sys.path:
            sys.path.append(p)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    _orig_masked_fill = torch.Tensor.masked_fill

    def masked_fill(self, mask, value):
        if mask.dtype == torch.uint8:
            mask = mask.bool()
        return _orig_masked_fill(self, mask, value)

    torch.Tensor.masked_fill = masked_fill
    import collections
    import collections.abc
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable
    _shimmed = True


def load_file(relpath, name=None):
    """Executes one reference source file (path relative to /root/reference/code) as a module."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_shims()
    name = name or ("_isa_ref_" + relpath.replace("/", "_").replace(".py", ""))
    if name in _loaded:
        return _loaded[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_CODE, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _loaded[name] = mod
    return mod


def discriminative():
    """lib/losses/discriminative.py (runs unmodified with usegpu=False)."""
    return load_file("lib/losses/discriminative.py")


def attention_utils():
    """lib/archs/modules/utils.py: ScaledDotProductAttention, MultiHeadAttention, DecoderLayer,
    SpatialAttentionLayer, HardAttentionLayer, AttentionLayer, make_position_encoding ..."""
    return load_file("lib/archs/modules/utils.py")
