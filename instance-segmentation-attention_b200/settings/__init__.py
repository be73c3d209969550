"""Settings classes with the reference's names and fields
(/root/reference/code/settings/__init__.py:1-3, settings/CVPPP/*.py) plus a Cityscapes-shaped
variant that the reference lacks (train.py:39 asserts CVPPP only)."""
from .cvppp import DataSettings as CVPPPDataSettings  # noqa: F401
from .cvppp import ModelSettings as CVPPPModelSettings  # noqa: F401
from .cvppp import TrainingSettings as CVPPPTrainingSettings  # noqa: F401
from .cityscapes import ModelSettings as CityscapesModelSettings  # noqa: F401
from .cityscapes import TrainingSettings as CityscapesTrainingSettings  # noqa: F401
