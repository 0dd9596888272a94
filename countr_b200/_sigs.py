"""argtypes / restype of the plain-argument entry points (include/countr_b200.h)."""
from ctypes import c_double, c_float, c_int32, c_int64, c_uint64, c_void_p

P, I, L, F = c_void_p, c_int32, c_int64, c_float

SIGS = {
    "countr_memset_zero": [P, c_int64, P],
    "countr_layernorm_fwd": [P, P, P, P, P, P, P, I, I, F, I, P],
    "countr_layernorm_bwd": [P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, P],
    "countr_layernorm_bwd_blocks": [I],
    "countr_attention_fwd": [P, P, P, I, I, I, I, F, I, P],
    "countr_attention_bwd": [P, P, P, P, P, P, I, I, I, I, F, I, P],
    "countr_cross_attn_core": [P, P, P, P, P, I, I, I, I, I, F, I, I, P],
    "countr_cast_f32_to_16": [P, P, L, F, I, P],
    "countr_cast_transpose_f32_to_16": [P, P, I, I, I, P],
    "countr_patchify": [P, I, L, L, L, L, P, I, I, I, I, I, I, P],
    "countr_conv_weight_pack": [P, P, I, I, I, I, P],
    "countr_gn_relu_upsample2x": [P, P, P, P, P, I, I, I, I, I, F, I, P],
    "countr_gn_relu_conv1x1": [P, P, P, P, P, P, P, I, I, I, I, F, I, P],
    "countr_upsample2x_f32": [P, P, I, I, I, I, P],
    "countr_exemplar_conv1": [P, I, L, L, L, L, L, P, P, P, I, I, I, I, I, P],
    "countr_inorm_relu_pool": [P, P, P, P, P, P, I, I, I, I, F, I, I, P],
    "countr_upsample2x_bwd": [P, I, P, I, I, I, P],
    "countr_gn_relu_bwd_reduce": [P, P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, F, I, P],
    "countr_gn_bwd_apply": [P, P, P, P, P, P, P, I, I, I, I, F, I, P],
    "countr_colsum": [P, I, P, L, I, L, P],
    "countr_softmax_bwd_rows": [P, P, P, L, I, F, I, P],
    "countr_cross_attn_core_bwd": [P, P, P, P, P, P, P, P, I, I, I, I, I, F, I, I, P],
    "countr_inorm_relu_pool_bwd": [P, P, P, P, P, P, P, P, I, I, I, I, I, I, P],
    "countr_exemplar_conv1_dw": [P, I, L, L, L, L, L, P, P, I, I, I, I, P],
    "countr_conv_dw_unpack": [P, P, I, I, P],
    "countr_gather_rows": [P, P, P, I, I, I, I, P],
    "countr_mae_unshuffle": [P, P, P, P, P, I, I, I, I, P],
    "countr_mae_loss": [P, P, I, L, L, L, L, P, P, I, I, I, I, I, I, P],
    "countr_cast_scaled_f32_to_16": [P, P, P, L, I, P],
    "countr_finetune_loss": [P, I, P, I, P, L, c_uint64, F, P, F, P, P, P, P, P, I, I, I, P],
    "countr_grad_stats": [P, L, P, P, P],
    "countr_adamw_update": [P, I, P, I, P, P, P, P, P, P, F, F, F, F, F, I, P],
    "countr_window_blend": [P, I, P, I, I, I, I, P, P],
    "countr_weight_refresh": [P, P, I, I, I, P],
    "countr_crop_resize_boxes": [P, L, L, L, L, P, P, I, I, I, I, I, I, P],
    "countr_crop_resize": [P, L, L, L, L, P, P, I, I, I, I, I, I, I, P],
    "countr_rect_mass": [P, I, I, P, I, F, P, P],
    "countr_aug_noise_clamp": [P, P, L, F, c_uint64, P],
    "countr_aug_color_jitter": [P, P, P, P, I, I, I, P],
    "countr_aug_gaussian_blur": [P, P, P, P, I, I, I, I, I, P],
    "countr_aug_hflip": [P, P, P, I, I, I, I, P],
    "countr_aug_mosaic": [P, I, I, I, P, P],
    "countr_aug_mosaic_dots": [P, I, I, P, P, P],
    "countr_density_filter": [P, P, P, I, I, I, P, I, F, P],
    "countr_aug_affine": [P, P, I, I, I, P, P],
    "countr_aug_affine_dots": [P, I, c_double, c_double, I, I, P, P, P],
    "countr_grouped_dw": [P, I, I, P],
    "countr_grouped_colsum": [P, I, P],
    "countr_density_from_dots": [P, P, I, I, c_double, c_double, I, I, I, I, I, I, P, I, F, P, P, P],
}


SIZE_QUERIES = {
    "countr_finetune_loss_scratch_bytes": [I],
    "countr_grad_stats_scratch_bytes": [],
    "countr_attention_bwd_workspace_bytes": [I, I, I, I],
}


def declare(lib):
    for name, args in SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int32
    for name, args in SIZE_QUERIES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int64
