"""ctypes signatures of every symbol include/b200u.h declares (kept in the same order).

tests/test_abi.py checks this table against the header and against the built library, so a
symbol added to one and not the others fails the CPU suite.
"""
import ctypes as C

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_ll = C.c_longlong
_ull = C.c_ulonglong
_sz = C.c_size_t
_u32 = C.c_uint32

SIGNATURES = {
    "b200u_last_error_string": (C.c_char_p, []),
    "b200u_version": (_i, []),
    "b200u_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "b200u_launch_count": (_ll, []),
    "b200u_set_pdl": (_i, [_i]),
    "b200u_set_sm_limit": (_i, [_i]),
    "b200u_set_bwd_streams": (_i, [_i]),
    "b200u_set_fused_layernorm": (_i, [_i]),
    "b200u_set_attention_impl": (_i, [_i]),
    "b200u_prof_enable": (_i, [_i]),
    "b200u_prof_collect": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i)]),
    "b200u_gemm": (_i, [_p, _p]),
    "b200u_gemm_debug_stamps": (_i, [_p]),
    "b200u_layernorm_fwd": (_i, [_p, _i, _p, _p, _p, _i, _p, _p, _i, _i, _f, _p, _p]),
    "b200u_layernorm_bwd": (_i, [_p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _i, _p]),
    "b200u_colsum_accum": (_i, [_p, _i, _p, _i, _i, _p]),
    "b200u_dgelu_mul": (_i, [_p, _p, _p, _sz, _p]),
    "b200u_cast_f32_to_bf16": (_i, [_p, _p, _sz, _p]),
    "b200u_slice_sum_bf16": (_i, [_p, _p, _i, _sz, _p]),
    "b200u_ce_finish": (_i, [_p, _p, _p, _p, _i, _i, _p]),
    "b200u_gather_rows": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "b200u_gather_rows_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "b200u_txt_embed_fwd": (_i, [_p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p, _p]),
    "b200u_img_embed_fwd": (_i, [_p] * 16 + [_i, _i, _i, _f, _p, _p]),
    "b200u_embedding_scatter_add": (_i, [_p, _p, _i, _i, _ll, _p, _i, _i, _ll, _ll, _p]),
    "b200u_embedding_segment_add": (_i, [_p, _p, _p, _p, _i, _i, _ll, _ll, _p, _i, _p]),
    "b200u_input_errors": (_i, [_p, _i]),
    "b200u_pos_linear_wgrad": (_i, [_p, _p, _p, _i, _i, _p]),
    "b200u_attention_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "b200u_attention_bwd_scratch_bytes": (_sz, [_i, _i, _i]),
    "b200u_attention_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "b200u_bert_layer_fwd": (_i, [_p, _p, _p, _p, _p]),
    "b200u_bert_layer_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p]),
    "b200u_pooler_fwd": (_i, [_p, _ll, _p, _p, _p, _i, _i, _p]),
    "b200u_pooler_bwd": (_i, [_p, _p, _p, _ll, _p, _p, _p, _p, _ll, _i, _i, _p]),
    "b200u_linear_small_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "b200u_linear_small_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "b200u_bce_logits": (_i, [_p, _p, _f, _f, _p, _p, _p, _i, _p]),
    "b200u_cosine_cost": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "b200u_ipot": (_i, [_p, _p, _p, _p, _i, _i, _i, _f, _i, _i, _p]),
    "b200u_ot_distance": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "b200u_cosine_cost_bwd": (_i, [_p] * 10 + [_i, _i, _i, _i, _p]),
    "b200u_counter_add": (_i, [_p, _ull, _p]),
    "b200u_grad_sumsq": (_i, [_p, _sz, _p, _p, _sz, _sz, _p]),
    "b200u_clip_coef": (_i, [_p, _f, _f, _p, _p, _p]),
    "b200u_adam_step": (_i, [_p, _p, _p, _p, _p, _sz, _p, _p, _p, _i, _p, _p, _p, _f, _f, _f, _i, _p, _sz, _sz, _p]),
}
