"""countr_b200 — B200-native (sm_100a) implementation of CounTR's forward/backward hot path.

The package mirrors the reference's Python surface (models_mae_cross / models_crossvit /
models_mae_noct); all device work goes through libcountr_sm100.so (include/countr_b200.h).
"""
__version__ = "0.1.0"
