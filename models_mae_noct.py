"""`import models_mae_noct` from the repo root -> B200-native implementation (reference: models_mae_noct.py,
consumed by FSC_pretrain.py:202)."""
from countr_b200.models_mae_noct import *  # noqa: F401,F403
from countr_b200.models_mae_noct import MaskedAutoencoderViTNoCT  # noqa: F401
