"""`import models_crossvit` from the repo root -> B200-native implementation (reference: models_crossvit.py)."""
from countr_b200.models_crossvit import *  # noqa: F401,F403
from countr_b200.models_crossvit import (Attention, CrossAttention, CrossAttentionBlock, DropPath, Mlp, drop_path,  # noqa: F401
                                         to_2tuple)
