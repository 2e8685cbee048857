"""Two-symbol stand-in for `omegaconf`, only so the reference tree imports in the build container (oracle use only).

The reference touches omegaconf at ldm/modules/utils.py:13 and openaimodel.py:479 (type check on ListConfig).
"""


class OmegaConf:  # pragma: no cover - never instantiated by the hot path
    pass
