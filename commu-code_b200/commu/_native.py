"""ctypes binding of libcommu_b200.so (the C-ABI declared in include/commu_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_PKG_ROOT, "lib", "libcommu_b200.so")

_lib = None
_lock = threading.Lock()

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_int64 = ctypes.c_int64
c_float = ctypes.c_float


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("a", c_void_p), ("lda", c_int64), ("a_mn_major", c_int),
        ("b", c_void_p), ("ldb", c_int64), ("b_mn_major", c_int),
        ("m", c_int), ("n", c_int), ("k", c_int),
        ("split_k", c_int), ("alpha", c_float),
        ("bias", c_void_p), ("relu", c_int),
        ("relu_mask", c_void_p), ("ld_mask", c_int64),
        ("add_f32", c_void_p), ("ld_add", c_int64),
        ("out_bf16", c_void_p), ("ld_out_bf16", c_int64),
        ("out_f32", c_void_p), ("ld_out_f32", c_int64),
        ("f32_atomic", c_int), ("impl", c_int),
        ("drop_p", c_float), ("drop_seed", ctypes.c_uint64),
    ]


class DecLinearArgs(ctypes.Structure):
    """CommuDecLinear of include/commu_b200.h."""
    _fields_ = [
        ("prologue", c_int), ("epilogue", c_int),
        ("B", c_int), ("K", c_int), ("N", c_int), ("d_true", c_int),
        ("split_k", c_int), ("pdl", c_int),
        ("tokens", c_void_p), ("emb", c_void_p), ("emb_scale", c_float),
        ("z", c_void_p), ("ldz", c_int64), ("gamma", c_void_p), ("beta", c_void_p), ("eps", c_float),
        ("a_bf16", c_void_p), ("lda", c_int64),
        ("x_out", c_void_p), ("ldx", c_int64),
        ("w", c_void_p), ("ldw", c_int64), ("bias", c_void_p),
        ("res", c_void_p), ("ldr", c_int64),
        ("out_f32", c_void_p), ("ldo", c_int64),
        ("out_bf16", c_void_p), ("ldob", c_int64),
        ("q_out", c_void_p), ("k_cache", c_void_p), ("v_cache", c_void_p),
        ("H", c_int), ("C", c_int), ("slot", c_int),
        ("dev_state", c_void_p),
    ]


PRO_EMBED, PRO_LN, PRO_BF16 = 0, 1, 2
EPI_QKV, EPI_RES, EPI_RELU, EPI_LOGITS = 0, 1, 2, 3


def dec_linear_args(**kw):
    """Builds a DecLinearArgs; tensors become device pointers (a prepared struct can be reused every step)."""
    a = DecLinearArgs()
    for k, v in kw.items():
        if hasattr(v, "data_ptr"):
            v = v.data_ptr()
        setattr(a, k, v)
    return a


def dec_linear(args):
    check(lib().commu_decode_fused_linear(ctypes.byref(args), stream_ptr()))


def build_if_needed():
    """(Re)build the shared library in-tree when nvcc is available and sources are newer."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("commu_b200_build", os.path.join(_PKG_ROOT, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def lib():
    """Loads the shared library (once)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "commu_b200: %s is missing. Build it with `python commu-code_b200/build.py` "
                "(needs nvcc). There is no CPU or PyTorch fallback for the hot path." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        L.commu_last_error.restype = ctypes.c_char_p
        L.commu_launch_count.restype = ctypes.c_longlong
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().commu_last_error()
        raise RuntimeError("commu_b200 native call failed (%d): %s" % (rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(a, b, *, m, n, k, lda=None, ldb=None, a_mn=False, b_mn=False, split_k=1, alpha=1.0,
         bias=None, relu=False, relu_mask=None, ld_mask=0, add_f32=None, ld_add=0,
         out_bf16=None, ld_out_bf16=0, out_f32=None, ld_out_f32=0, f32_atomic=False, impl=0,
         drop_p=0.0, drop_seed=0):
    """C[m,n] = alpha * sum_k A[m,k] B[n,k] with the fused epilogue of commu_gemm_bf16."""
    args = GemmArgs()
    args.a = a.data_ptr(); args.lda = lda if lda is not None else a.stride(0); args.a_mn_major = int(a_mn)
    args.b = b.data_ptr(); args.ldb = ldb if ldb is not None else b.stride(0); args.b_mn_major = int(b_mn)
    args.m, args.n, args.k = m, n, k
    args.split_k = split_k
    args.alpha = alpha
    args.bias = bias.data_ptr() if bias is not None else None
    args.relu = int(relu)
    args.relu_mask = relu_mask.data_ptr() if relu_mask is not None else None
    args.ld_mask = ld_mask or (relu_mask.stride(0) if relu_mask is not None else 0)
    args.add_f32 = add_f32.data_ptr() if add_f32 is not None else None
    args.ld_add = ld_add or (add_f32.stride(0) if add_f32 is not None else 0)
    args.out_bf16 = out_bf16.data_ptr() if out_bf16 is not None else None
    args.ld_out_bf16 = ld_out_bf16 or (out_bf16.stride(0) if out_bf16 is not None else 0)
    args.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
    args.ld_out_f32 = ld_out_f32 or (out_f32.stride(0) if out_f32 is not None else 0)
    args.f32_atomic = int(f32_atomic)
    args.impl = impl
    args.drop_p = drop_p
    args.drop_seed = drop_seed & 0xFFFFFFFFFFFFFFFF
    check(lib().commu_gemm_bf16(ctypes.byref(args), stream_ptr()))


def attn_sizes(T, M, B, H):
    """Byte sizes of the buffers of the materialised attention backward for one call of shape (T, M, B, H):
    (p_save, mt_save, workspace, leading part of the workspace that must be zero on first use)."""
    out = [c_int64(0) for _ in range(4)]
    fn = lib().commu_relattn_bwd_sizes
    fn.argtypes = [c_int, c_int, c_int, c_int] + [ctypes.POINTER(c_int64)] * 4
    fn.restype = c_int
    check(fn(T, M, B, H, *[ctypes.byref(o) for o in out]))
    return tuple(int(o.value) for o in out)


_attn_ws = {}


def attn_bwd_workspace(T, M, B, H, device):
    """The backward workspace of a shape, allocated (and zeroed) once and then shared by every layer and step of
    that shape - the kernels keep its zero regions valid (include/commu_b200.h, commu_relattn_bwd).  At most two
    shapes stay resident (a training run alternates between at most the warm-up and the steady-state memory length)."""
    import torch
    key = (T, M, B, H, str(device))
    ws = _attn_ws.get(key)
    if ws is None:
        _, _, ws_bytes, _ = attn_sizes(T, M, B, H)
        while len(_attn_ws) >= 2:
            _attn_ws.pop(next(iter(_attn_ws)))
        ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=device)
        _attn_ws[key] = ws
    return ws


# ------------------------------------------------------------------------------------------------
# Typed signatures of the remaining entry points (kept in one table so a CPU test can verify that
# every symbol declared in include/commu_b200.h is exported and bound).
# ------------------------------------------------------------------------------------------------
P, I, L, F, U = c_void_p, c_int, c_int64, c_float, ctypes.c_uint
SIGNATURES = {
    "commu_abi_version": [],
    "commu_device_info": [P, P, P],
    "commu_prof_arm": [U],
    "commu_prof_read": [I, P, P],
    "commu_relattn_bwd_set_impl": [I, I, I],
    "commu_relattn_set_dropout": [F, ctypes.c_uint64],
    "commu_dropout": [P, I, L, P, L, L, I, F, ctypes.c_uint64, P, L, P, L, P],
    "commu_relattn_bwd_dr_tc": [P, P, L, P, P, L, P, L, I, P, I, I, I, I, I, I, F, P, P, L, P, P, P, P, P],
    "commu_relattn_bwd_dq_tc": [P, P, L, P, P, L, P, L, I, P, I, I, I, I, I, I, F, P, P, L, P, P, L, P, P, P],
    "commu_relattn_bwd_dkv_tc": [P, P, L, P, P, L, P, L, I, P, I, I, I, I, I, I, F, P, P, L, P, P, P, L, P],
    "commu_decode_linear": [P, L, P, L, I, P, I, P, L, P, L, I, I, I, P],
    "commu_decode_linear_tiled": [P, L, P, L, I, P, I, P, L, P, L, I, I, I, I, P, P, P],
    "commu_decode_qkv_tiled": [P, L, P, L, I, I, I, I, P, P, P, I, I, P, P, P, P],
    "commu_pad_heads": [P, L, I, I, I, I, P, I, L, L, L, P, P],
    "commu_decode_advance": [P, I, I, I, P],
    "commu_decode_attn": [P, P, P, P, I, P, P, I, I, I, I, I, F, P, L, P, P, I, P, P, P],
    "commu_decode_fused_linear": [P, P],
    "commu_decode_attn_split": [P, P, P, P, P, P, I, I, I, I, I, F, I, P, P, P, P, L, P, I, I, P],
    "commu_sample": [P, L, I, I, F, I, F, P, ctypes.c_uint64, ctypes.c_uint64, P, P, L, P, P],
    "commu_comm_unique_id": [P, P],
    "commu_comm_init": [P, P, I, I],
    "commu_allreduce_sum_f32": [P, L, P],
    "commu_comm_destroy": [],
    "commu_gemm_bf16": [P, P],
    "commu_embed_fwd": [P, P, I, I, F, L, P, L, P, L, P],
    "commu_embed_bwd": [P, P, L, I, F, L, P, P],
    "commu_pos_table": [P, I, I, I, I, P, P, P],
    "commu_layernorm_fwd": [P, L, P, P, I, I, F, L, P, L, P, L, P, P, P],
    "commu_layernorm_bwd": [P, L, P, L, P, P, P, I, I, L, P, L, P, L, P, P, F, ctypes.c_uint64, P],
    "commu_nll_fwd": [P, L, I, P, L, P, P, P],
    "commu_nll_bwd": [P, L, I, I, P, P, P, L, P, L, P],
    "commu_colsum_bf16": [P, L, I, L, P, P],
    "commu_cast_pad": [P, L, I, I, I, I, I, I, P, L, I, P],
    "commu_unpad_accum": [P, L, I, I, I, I, I, I, P, L, F, P],
    "commu_sumsq": [P, L, P, P],
    "commu_clip_adam": [P, P, P, P, L, F, F, F, F, I, P, F, F, F, P, P, P],
    "commu_relattn_fwd": [P, L, P, P, L, P, L, I, P, P, P, I, I, I, I, I, I, F, P, L, P, P, P, P],
    "commu_relattn_fwd_tc": [P, L, P, P, L, P, L, I, P, P, P, I, I, I, I, I, I, F, P, L, P, P, P, P, P, P],
    "commu_relattn_bwd": [P, P, L, P, P, L, P, L, I, P, I, I, I, I, I, I, F, P, L, P, P, L, P, P, L,
                          P, P, L, P, P, P, P, P, P, L, P],
    "commu_relattn_bwd_sizes": [I, I, I, I, P, P, P, P],
    "commu_relattn_bwd_mat": [P, P, L, P, P, L, P, L, I, P, I, I, I, I, I, I, F, P, P, L, P, P, P, P, L, P, L,
                              P, P, L, P, P, P, P],
}
_bound = set()


def _as_arg(x):
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return c_void_p(x.data_ptr())
    return x


def call(name, *args):
    """Calls an entry point with torch tensors converted to device pointers; the last argument
    (the stream) is appended automatically for every function that takes one."""
    L_ = lib()
    fn = getattr(L_, name)
    sig = SIGNATURES[name]
    if name not in _bound:
        fn.argtypes = sig
        fn.restype = c_int
        _bound.add(name)
    conv = [_as_arg(a) for a in args]
    if len(conv) == len(sig) - 1:
        conv.append(stream_ptr())
    if len(conv) != len(sig):
        raise TypeError("%s expects %d arguments, got %d" % (name, len(sig), len(conv)))
    check(fn(*conv))
