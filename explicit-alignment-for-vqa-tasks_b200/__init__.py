"""B200-native CLIP-prefix language-model step (drop-in for ``src/models/clipcap.py``).

Host side: Python/PyTorch mirror of the reference's model interface
(``ClipCaptionPrefixB200``) over a C-ABI CUDA library (``csrc/``, ``include/eavqa_b200.h``).
Import as ``eavqa_b200`` (see ``eavqa_b200.py`` at the repo root).
"""
from . import synthetic  # noqa: F401
from .model import ClipCaptionModelB200, ClipCaptionPrefixB200  # noqa: F401

__all__ = ["synthetic", "ClipCaptionModelB200", "ClipCaptionPrefixB200"]
