"""nnr_b200 -- B200-native CNE+SUE hot path behind the reference's module API.

Importing the package loads ``_lib/libnnr_b200.so`` (built by ``csrc/build.sh``); it raises if the
library is missing -- there is no CPU or PyTorch fallback.
"""
from . import _lib, ops, engine          # noqa: F401  (fail loudly if the CUDA library is absent)
from .newsEncoders import CNE, NewsEncoder  # noqa: F401
from .userEncoders import SUE, UserEncoder  # noqa: F401
from .variantEncoders import CNE_Content, CNE_Title, CNE_wo_CA, CNE_wo_CS, SUE_wo_GCN, SUE_wo_HCA  # noqa: F401
from .model import Model                 # noqa: F401
from . import corpus, metrics, scoring, trainer  # noqa: F401
