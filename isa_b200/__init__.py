"""Importable name of the package that lives in `instance-segmentation-attention_b200/`
(a directory name Python cannot import directly because of the hyphens).

    import isa_b200
    from isa_b200.losses import DiscriminativeLoss
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "instance-segmentation-attention_b200")
__path__.append(_PKG_DIR)

from . import _lib  # noqa: E402,F401

__all__ = ["_lib"]
