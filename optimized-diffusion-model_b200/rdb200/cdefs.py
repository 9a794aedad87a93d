"""ctypes mirrors of the structs in include/rdb200.h (field order and types must match exactly;
tests/test_abi.py cross-checks sizeof/offsetof against a tiny C program compiled from the header)."""
import ctypes as C

RD_PREC_BF16, RD_PREC_F32X3 = 0, 1
RD_OP_CONV, RD_OP_ATTN_CORE, RD_OP_TEMB, RD_OP_IN_CONV, RD_OP_OUT_HEAD, RD_OP_ATTN_BLOCK = 1, 2, 3, 4, 5, 6

i32 = C.c_int32
vp = C.c_void_p


class GtoHaloCodec(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "data_mean", "data_std", "shooting_time_min", "shooting_time_span", "coast_time_min", "coast_time_span",
        "halo_energy_min", "halo_energy_span", "fuel_mass_min", "fuel_mass_span", "manifold_length_min",
        "manifold_length_span", "thrust")] + [("n_triplets", i32)]


class ConvSrc(C.Structure):
    _fields_ = [("ptr", vp), ("C", i32), ("Hs", i32), ("Ws", i32)]


class OpConv(C.Structure):
    _fields_ = [
        ("src", ConvSrc * 2), ("nsrc", i32), ("H_in", i32), ("W_in", i32), ("pad", i32), ("stride", i32),
        ("H_out", i32), ("W_out", i32), ("ntaps", i32), ("C_out", i32), ("gn_groups", i32), ("gn_silu", i32),
        ("gn_eps", C.c_float), ("gn_gamma", vp), ("gn_beta", vp), ("w", vp), ("bias", vp), ("tproj", vp),
        ("tproj_stride", i32), ("tproj_off", i32), ("tproj_wrap", i32), ("residual", vp), ("out_scale", C.c_float), ("out", vp),
        ("B2", i32), ("samples_per_cta", i32), ("precision", i32), ("out_stride", i32),
        ("sc_src", ConvSrc * 2), ("sc_nsrc", i32), ("_pad0", i32)]


class OpAttn(C.Structure):
    _fields_ = [("qkv", vp), ("out", vp), ("B2", i32), ("T", i32), ("C", i32), ("precision", i32)]


class OpAttnBlock(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("wqkv_t", vp), ("wproj_t", vp), ("bqkv", vp), ("bproj", vp), ("gamma", vp),
                ("beta", vp), ("B2", i32), ("T", i32), ("C", i32), ("groups", i32), ("eps", C.c_float),
                ("out_scale", C.c_float)]


class OpTemb(C.Structure):
    _fields_ = [("time_table", vp), ("label_w", vp), ("labels", vp), ("dense_w", vp), ("dense_b", vp), ("out", vp),
                ("step_ctr", vp), ("row_idx", vp), ("B2", i32), ("temb_dim", i32), ("num_classes", i32),
                ("n_out_total", i32)]


class OpInConv(C.Structure):
    _fields_ = [("x", vp), ("w", vp), ("bias", vp), ("out", vp), ("B", i32), ("B2", i32), ("C_in", i32),
                ("C_out", i32), ("H", i32), ("W", i32), ("precision", i32)]


class OpOutHead(C.Structure):
    _fields_ = [("h", vp), ("gamma", vp), ("beta", vp), ("w", vp), ("bias", vp), ("cfg_w", vp),
                ("cfg_w_scalar", C.c_float), ("score", vp), ("B", i32), ("B2", i32), ("C", i32), ("C_img", i32),
                ("H", i32), ("W", i32), ("groups", i32), ("cfg", i32), ("eps", C.c_float), ("precision", i32),
                ("sigma_table", vp), ("step_ctr", vp)]


class _OpUnion(C.Union):
    _fields_ = [("conv", OpConv), ("attn", OpAttn), ("attn_block", OpAttnBlock), ("temb", OpTemb), ("inconv", OpInConv), ("outhead", OpOutHead)]


class Op(C.Structure):
    _fields_ = [("kind", i32), ("_pad", i32), ("u", _OpUnion)]


class SamplerDesc(C.Structure):
    _fields_ = [("forward", vp), ("x", vp), ("score", vp), ("partial", vp), ("g_table", vp), ("step_ctr", vp),
                ("noise_tape", vp), ("seed", C.c_uint64), ("snr", C.c_float), ("dt", C.c_float),
                ("sqrt_dt", C.c_float), ("B", i32), ("D", i32), ("n_corrector_steps", i32)]
