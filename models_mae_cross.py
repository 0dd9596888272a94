"""`import models_mae_cross` from the repo root resolves to the B200-native implementation, exactly
as the reference scripts expect (FSC_finetune_cross.py:27, demo.py:18)."""
from countr_b200.models_mae_cross import *  # noqa: F401,F403
from countr_b200.models_mae_cross import SupervisedMAE  # noqa: F401
