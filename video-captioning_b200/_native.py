"""ctypes binding of the C ABI in ``include/vc_b200.h`` (``libvc_b200.so``).

There is deliberately no CPU or PyTorch fallback: if the library cannot be built/loaded, or a call
returns a non-zero status, a ``RuntimeError`` is raised.  PyTorch is used here only for device memory
(``torch.empty`` workspaces / outputs) and the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import hashlib
import logging
import os
import shutil
import subprocess
import threading
from typing import Dict, Optional

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO_ROOT = os.path.dirname(_PKG_DIR)
_CSRC = os.path.join(_PKG_DIR, "csrc")
_INCLUDE = os.path.join(_REPO_ROOT, "include")
LIB_PATH = os.path.join(_PKG_DIR, "libvc_b200.so")

ATTN_BAHDANAU, ATTN_LUONG_DOT, ATTN_LUONG_GENERAL, ATTN_LUONG_CONCAT, ATTN_MULTIHEAD = range(5)
PREC_FP32, PREC_BF16 = 0, 1
METHOD_GREEDY, METHOD_BEAM = 0, 1

ATTENTION_IDS = {
    "bahdanau": ATTN_BAHDANAU, "luong_dot": ATTN_LUONG_DOT, "luong_general": ATTN_LUONG_GENERAL,
    "luong_concat": ATTN_LUONG_CONCAT, "multihead": ATTN_MULTIHEAD,
}
PRECISION_IDS = {"fp32": PREC_FP32, "bf16": PREC_BF16}

# every symbol include/vc_b200.h declares (tests check the .so exports all of them)
EXPORTED_SYMBOLS = (
    "vc_last_error", "vc_version", "vc_launch_count", "vc_profile_begin", "vc_profile_end", "vc_model_create", "vc_model_set_weight", "vc_model_finalize",
    "vc_model_destroy", "vc_workspace_bytes", "vc_encoder_forward", "vc_attn_precompute",
    "vc_decode_greedy", "vc_decode_beam", "vc_beam_nbest", "vc_generate", "vc_generate_ex", "vc_host_pack_bf16", "vc_host_stage_rows", "vc_convert_bf16",
    "vc_forward_teacher", "vc_linear",
    "vc_attention_step", "vc_beam_select",
)


class ModelDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "feature_dim", "hidden_dim", "embed_dim", "attn_dim", "vocab_size", "enc_layers", "dec_layers",
        "attention", "num_heads", "precision")]


class DecodeParams(ctypes.Structure):
    _fields_ = [("method", ctypes.c_int32), ("beam_size", ctypes.c_int32), ("max_length", ctypes.c_int32),
                ("start_token_id", ctypes.c_int32), ("end_token_id", ctypes.c_int32),
                ("length_penalty", ctypes.c_float), ("temperature", ctypes.c_float),
                ("diverse_beams", ctypes.c_int32)]


def nvcc_command(out_path: str = LIB_PATH):
    return ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
            "-shared", "-Xcompiler", "-fPIC", "-I", _INCLUDE, "-o", out_path, os.path.join(_CSRC, "capi.cu"),
            os.path.join(_CSRC, "host_pack.cpp")]


def _source_files():
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC)) + [os.path.join(_INCLUDE, "vc_b200.h")]


def sources_hash() -> str:
    """Hash of everything the library is compiled from (mtimes do not survive a copy of the tree)."""
    h = hashlib.sha256()
    for f in _source_files():
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


_HASH_PATH = LIB_PATH + ".srchash"


def library_is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(_HASH_PATH)):
        return False
    with open(_HASH_PATH) as fh:
        return fh.read().strip() == sources_hash()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/capi.cu for sm_100a into libvc_b200.so (in-tree).  Cross-compiles without a GPU.  The hash of the
    sources is stored beside the library, so a library left over from older sources is never loaded silently."""
    if not force and library_is_current():
        return LIB_PATH
    # One builder at a time across processes (the ranks of a torchrun launch all get here when the library is stale): an
    # exclusive lock on a side file, the state re-checked under it, a per-process temporary name, an atomic rename.
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and library_is_current():
                return LIB_PATH
            tmp = f"{LIB_PATH}.tmp{os.getpid()}"
            proc = subprocess.run(nvcc_command(tmp), capture_output=True, text=True)
            if proc.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libvc_b200.so:\n" + proc.stdout + proc.stderr)
            if os.path.exists(_HASH_PATH):
                os.remove(_HASH_PATH)          # never a new library beside the old hash
            os.replace(tmp, LIB_PATH)
            with open(_HASH_PATH, "w") as fh:
                fh.write(sources_hash())
            if verbose:
                print(proc.stdout + proc.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_lib = None
_lib_lock = threading.Lock()


def load_library() -> ctypes.CDLL:
    """Load (building first if needed) the native library.  Raises if that is impossible."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not library_is_current():
            # missing, or compiled from other sources than the ones in the tree (an edited csrc/ or header): rebuild
            if shutil.which("nvcc") is None:
                raise RuntimeError(f"{LIB_PATH} is missing or stale (csrc/ changed since it was built) and nvcc is not on "
                                   "PATH to rebuild it; there is no fallback path")
            build_library()
        lib = ctypes.CDLL(LIB_PATH)
        vp, i32, i64, f32p, i32p = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p
        sz = ctypes.c_size_t
        lib.vc_last_error.restype = ctypes.c_char_p
        lib.vc_last_error.argtypes = []
        lib.vc_version.restype = ctypes.c_int
        lib.vc_launch_count.restype = ctypes.c_longlong
        lib.vc_launch_count.argtypes = []
        lib.vc_profile_begin.argtypes = []
        lib.vc_profile_end.argtypes = [vp, vp]
        lib.vc_model_create.argtypes = [ctypes.POINTER(ModelDesc), ctypes.POINTER(vp)]
        lib.vc_model_set_weight.argtypes = [vp, ctypes.c_char_p, f32p, i64, vp]
        lib.vc_model_finalize.argtypes = [vp, vp]
        lib.vc_model_destroy.argtypes = [vp]
        lib.vc_model_destroy.restype = None
        lib.vc_workspace_bytes.argtypes = [vp, i32, i32, i32, i32]
        lib.vc_workspace_bytes.restype = sz
        lib.vc_encoder_forward.argtypes = [vp, f32p, i32, i32, i32p, f32p, f32p, vp, sz, vp]
        lib.vc_attn_precompute.argtypes = [vp, i32, i32, vp, sz, vp]
        lib.vc_decode_greedy.argtypes = [vp, i32, i32, f32p, ctypes.POINTER(DecodeParams), i32p, f32p, vp, sz, vp]
        lib.vc_decode_beam.argtypes = [vp, i32, i32, f32p, ctypes.POINTER(DecodeParams), i32p, i32p, f32p, vp, sz, vp]
        lib.vc_beam_nbest.argtypes = [vp, i32, i32, ctypes.POINTER(DecodeParams), i32, i32p, i32p, f32p, vp, sz, vp]
        lib.vc_generate.argtypes = [vp, f32p, i32, i32, i32p, f32p, ctypes.POINTER(DecodeParams), i32p, i32p, f32p,
                                    f32p, vp, sz, vp]
        lib.vc_generate_ex.argtypes = [vp, vp, i32, i32, i32, i32p, f32p, ctypes.POINTER(DecodeParams), i32p, i32p, f32p,
                                       f32p, vp, sz, vp]
        lib.vc_host_pack_bf16.argtypes = [vp, vp, sz, i32]
        lib.vc_host_stage_rows.argtypes = [vp, i64, i64, i32, vp, i32, i32]
        lib.vc_convert_bf16.argtypes = [f32p, vp, i64, vp]
        lib.vc_forward_teacher.argtypes = [vp, f32p, i32, i32, i32p, f32p, i32p, i32, f32p, f32p, f32p, vp, sz, vp]
        lib.vc_linear.argtypes = [i32, f32p, f32p, f32p, f32p, i32, i32, i32, i32, vp, sz, vp]
        lib.vc_attention_step.argtypes = [vp, f32p, f32p, f32p, i32, i32, i32, f32p, f32p, vp, sz, vp]
        lib.vc_beam_select.argtypes = [f32p, f32p, i32, i32, i32, i32p, i32p, f32p, vp, sz, vp]
        for name in EXPORTED_SYMBOLS:
            fn = getattr(lib, name)
            if name not in ("vc_last_error", "vc_model_destroy", "vc_workspace_bytes", "vc_launch_count"):
                fn.restype = ctypes.c_int
        _lib = lib
        return lib


KERNEL_CLASSES = ("convert", "enc_feature_proj", "enc_input_proj", "enc_recurrent", "enc_output_proj",
                  "attn_precompute", "attn_query_proj", "attn_step", "attn_output_proj", "dec_lstm",
                  "dec_context_proj", "dec_vocab", "select", "reorder_embed", "misc")


_replayed_launches = 0     # kernel launches executed through CUDA-graph replays (the library counts at enqueue time)
_profiling = False         # per-class event timing needs the plain launch path


def launch_count() -> int:
    """Kernel launches of this package executed so far (plain launches + launches inside replayed CUDA graphs)."""
    return int(load_library().vc_launch_count()) + _replayed_launches


def profile_begin() -> None:
    global _profiling
    _profiling = True
    check(load_library().vc_profile_begin(), "vc_profile_begin")


def profile_end() -> Dict[str, Dict[str, float]]:
    """-> {class: {"ms": summed device milliseconds, "scopes": event-bracketed launch scopes}}"""
    global _profiling
    _profiling = False
    n = len(KERNEL_CLASSES)
    ms = (ctypes.c_float * n)()
    cnt = (ctypes.c_int32 * n)()
    check(load_library().vc_profile_end(ms, cnt), "vc_profile_end")
    return {KERNEL_CLASSES[i]: {"ms": float(ms[i]), "scopes": int(cnt[i])} for i in range(n)}


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load_library().vc_last_error().decode(errors="replace")
        if status == 1:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device: this package has no CPU path "
                           f"(got device {t.device})")


class NativeModel:
    """Owns one ``vc_model_t`` handle built from a reference-layout state_dict."""

    def __init__(self, desc: Dict[str, int], state_dict: Dict[str, torch.Tensor], device: torch.device):
        self.lib = load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("NativeModel needs a CUDA device; there is no CPU fallback")
        self.desc = ModelDesc(**desc)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.vc_model_create(ctypes.byref(self.desc), ctypes.byref(self._h)), "vc_model_create")
            st = _stream(self.device)
            keep = []
            for key, val in state_dict.items():
                t = val.detach().to(dtype=torch.float32).contiguous()
                keep.append(t)
                check(self.lib.vc_model_set_weight(self._h, key.encode(), _ptr(t), t.numel(), st),
                      f"vc_model_set_weight({key})")
            check(self.lib.vc_model_finalize(self._h, st), "vc_model_finalize")
            torch.cuda.current_stream(self.device).synchronize()
        self._ws: Optional[torch.Tensor] = None
        self._graphs_on = os.environ.get("VC_CUDA_GRAPHS", "1") != "0"
        self._graphs: Dict[tuple, tuple] = {}
        self._graph_seen: Dict[tuple, int] = {}

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self.lib.vc_model_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    # -- helpers
    @property
    def V(self):
        return self.desc.vocab_size

    @property
    def H(self):
        return self.desc.hidden_dim

    def workspace_bytes(self, B, T, K, S) -> int:
        return int(self.lib.vc_workspace_bytes(self._h, B, T, K, S))

    def _workspace(self, B, T, K, S) -> torch.Tensor:
        need = self.workspace_bytes(B, T, K, S)
        if self._ws is None or self._ws.numel() < need:
            self._graphs.clear()                  # captured graphs point into the old workspace
            self._graph_seen.clear()
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _params(self, method, K, S, start, end, length_penalty=1.0, temperature=1.0, diverse=False):
        return DecodeParams(METHOD_BEAM if method == "beam" else METHOD_GREEDY, int(K), int(S), int(start), int(end),
                            float(length_penalty), float(temperature), int(bool(diverse)))

    def _prep_feats(self, feats, allow_bf16=False):
        require_cuda(feats, "video_features")
        if allow_bf16 and feats.dtype == torch.bfloat16:
            f = feats.detach().contiguous()      # host-packed ingest: already rounded (vc_generate_ex, VC_DTYPE_BF16)
        else:
            f = feats.detach().to(dtype=torch.float32).contiguous()
        if f.dim() != 3 or f.shape[2] != self.desc.feature_dim:
            raise ValueError(f"video_features must be [B,T,{self.desc.feature_dim}], got {tuple(f.shape)}")
        return f

    @staticmethod
    def _prep_mask(mask, B, T, device):
        if mask is None:
            return None, None
        m = mask.detach().to(device=device, dtype=torch.float32).contiguous()
        if tuple(m.shape) != (B, T):
            raise ValueError(f"video_mask must be [{B},{T}], got {tuple(m.shape)}")
        lengths = m.sum(dim=1).to(torch.int32).contiguous()   # encoder.py:75
        return m, lengths

    # -- entry points
    def encoder_forward(self, feats, mask=None):
        f = self._prep_feats(feats)
        B, T, _ = f.shape
        m, lengths = self._prep_mask(mask, B, T, self.device)
        enc_out = torch.empty(B, T, self.H, dtype=torch.float32, device=self.device)
        final = torch.empty(B, self.H, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            ws = self._workspace(B, T, 1, 1)
            check(self.lib.vc_encoder_forward(self._h, _ptr(f), B, T, _ptr(lengths), _ptr(enc_out), _ptr(final),
                                              _ptr(ws), ws.numel(), _stream(self.device)), "vc_encoder_forward")
        return enc_out, final

    def generate(self, feats, start, end, max_length, mask=None, method="greedy", beam_size=5, length_penalty=1.0,
                 temperature=1.0, diverse=False, want_attention=True, nbest=0):
        """-> (tokens, lengths, scores, attention[, nbest_tokens [B,N,S+1], nbest_lengths [B,N], nbest_scores [B,N]]);
        the n-best triple (vc_beam_nbest) is appended when ``nbest`` > 0 (beam only)."""
        f = self._prep_feats(feats, allow_bf16=self.desc.precision == PREC_BF16)
        dtype_id = 1 if f.dtype == torch.bfloat16 else 0
        B, T, _ = f.shape
        m, lengths = self._prep_mask(mask, B, T, self.device)
        S = int(max_length)
        beam = method == "beam"
        K = int(beam_size) if beam else 1
        p = self._params(method, K, S, start, end, length_penalty, temperature, diverse)
        N = int(nbest) if beam else 0

        def alloc():
            base = (torch.empty(B, S + 1 if beam else S, dtype=torch.int32, device=self.device),
                    torch.empty(B, dtype=torch.int32, device=self.device) if beam else None,
                    torch.empty(B, dtype=torch.float32, device=self.device) if beam else None,
                    torch.empty(B, S, T, dtype=torch.float32, device=self.device) if (want_attention and not beam) else None)
            if N > 0:
                base += (torch.empty(B, N, S + 1, dtype=torch.int32, device=self.device),
                         torch.empty(B, N, dtype=torch.int32, device=self.device),
                         torch.empty(B, N, dtype=torch.float32, device=self.device))
            return base

        def launch(outs, ws):
            tokens, lens, scores, attn = outs[:4]
            check(self.lib.vc_generate_ex(self._h, _ptr(f), dtype_id, B, T, _ptr(lengths), _ptr(m), ctypes.byref(p),
                                          _ptr(tokens), _ptr(lens), _ptr(scores), _ptr(attn), _ptr(ws), ws.numel(),
                                          _stream(self.device)), "vc_generate")
            if N > 0:
                check(self.lib.vc_beam_nbest(self._h, B, T, ctypes.byref(p), N, _ptr(outs[4]), _ptr(outs[5]), _ptr(outs[6]),
                                             _ptr(ws), ws.numel(), _stream(self.device)), "vc_beam_nbest")

        with torch.cuda.device(self.device):
            ws = self._workspace(B, T, K, S)
            # The whole call is ~850 dependent launches: on small batches the host cannot enqueue them as fast as the
            # GPU runs them.  A call whose arguments (shapes AND buffer addresses) repeat is captured once as a CUDA graph
            # and replayed; the replay writes into the graph's own output buffers, which are cloned for the caller.
            key = None
            if self._graphs_on and not _profiling and m is None and f.data_ptr() == feats.data_ptr():
                key = (f.data_ptr(), dtype_id, B, T, K, S, method, int(start), int(end), float(length_penalty),
                       float(temperature), bool(diverse), bool(want_attention), N, ws.data_ptr(), ws.numel())
            ent = self._graphs.get(key) if key is not None else None
            global _replayed_launches
            if ent is not None:
                ent[0].replay()
                _replayed_launches += ent[4]
                return tuple(None if t is None else t.clone() for t in ent[1])
            if key is not None:
                if len(self._graph_seen) > 256:
                    self._graph_seen.clear()
                seen = self._graph_seen.get(key, 0) + 1
                self._graph_seen[key] = seen
                if seen >= 2:                      # second identical call: worth capturing
                    outs = alloc()
                    try:
                        g = torch.cuda.CUDAGraph()
                        n0 = int(self.lib.vc_launch_count())
                        with torch.cuda.graph(g, capture_error_mode="thread_local"):
                            launch(outs, ws)
                        n_launch = int(self.lib.vc_launch_count()) - n0      # counted once, at capture, for this call
                        if len(self._graphs) >= 16:
                            self._graphs.pop(next(iter(self._graphs)))
                        self._graphs[key] = (g, outs, f, ws, n_launch)     # keeps the captured buffers alive
                        g.replay()
                        return tuple(None if t is None else t.clone() for t in outs)
                    except RuntimeError as e:
                        # stream capture failed (torch raises RuntimeError / its subclass AcceleratorError for a launch the
                        # driver cannot record, and for a non-zero status of the library inside the capture): say so once and
                        # keep this handle on plain launches.  Anything else (ValueError of a bad argument, ...) propagates.
                        logging.getLogger(__name__).warning("CUDA-graph capture of generate() failed, using plain launches "
                                                            "for this handle: %s", e)
                        self._graphs_on = False
                        self._graphs.clear()
                        torch.cuda.synchronize(self.device)
            outs = alloc()
            launch(outs, ws)
        return outs

    def forward_teacher(self, feats, input_tokens, mask=None, want_attention=True):
        f = self._prep_feats(feats)
        B, T, _ = f.shape
        m, lengths = self._prep_mask(mask, B, T, self.device)
        tok = input_tokens.detach().to(device=self.device, dtype=torch.int32).contiguous()
        L = tok.shape[1]
        if tok.numel() and (int(tok.min()) < 0 or int(tok.max()) >= self.V):
            raise IndexError(f"input_tokens outside the vocabulary [0, {self.V})")     # nn.Embedding raises too (decoder.py:130)
        logits = torch.empty(B, L, self.V, dtype=torch.float32, device=self.device)
        attn = torch.empty(B, L, T, dtype=torch.float32, device=self.device) if want_attention else None
        enc_out = torch.empty(B, T, self.H, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            ws = self._workspace(B, T, 1, L)
            check(self.lib.vc_forward_teacher(self._h, _ptr(f), B, T, _ptr(lengths), _ptr(m), _ptr(tok), L,
                                              _ptr(logits), _ptr(attn), _ptr(enc_out), _ptr(ws), ws.numel(),
                                              _stream(self.device)), "vc_forward_teacher")
        return logits, attn, enc_out

    def attention_step(self, enc_out, hidden, mask, K, want_weights=True):
        """One attention step.  ``want_weights=False`` returns (context, None) through the context-only kernels the
        decode loop uses (the weights are an extra output of explain_prediction / teacher forcing)."""
        require_cuda(enc_out, "enc_out")
        e = enc_out.detach().float().contiguous()
        h = hidden.detach().float().contiguous()
        B, T, H = e.shape
        R = B * K
        m = None if mask is None else mask.detach().float().contiguous()
        ctx = torch.empty(R, H, dtype=torch.float32, device=self.device)
        w = torch.empty(R, T, dtype=torch.float32, device=self.device) if want_weights else None
        with torch.cuda.device(self.device):
            ws = self._workspace(B, T, K, 1)
            check(self.lib.vc_attention_step(self._h, _ptr(e), _ptr(h), _ptr(m), B, T, K, _ptr(ctx), _ptr(w), _ptr(ws),
                                             ws.numel(), _stream(self.device)), "vc_attention_step")
        return ctx, w


def host_pack_bf16(src: torch.Tensor, dst: torch.Tensor, threads: int) -> None:
    """dst (host, bf16) = src (host, fp32) rounded to nearest even, on ``threads`` host threads (GIL released)."""
    if src.device.type != "cpu" or dst.device.type != "cpu" or src.dtype != torch.float32 or dst.dtype != torch.bfloat16:
        raise ValueError("host_pack_bf16: host fp32 source and host bf16 destination expected")
    if not (src.is_contiguous() and dst.is_contiguous()) or src.numel() != dst.numel():
        raise ValueError("host_pack_bf16: contiguous buffers of equal length expected")
    check(load_library().vc_host_pack_bf16(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()), src.numel(),
                                           int(threads)), "vc_host_pack_bf16")


_STAGE_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def host_stage_rows(src_rows, n_rows: int, F: int, src_dtype: torch.dtype, dst: torch.Tensor, threads: int) -> None:
    """dst (host [n_rows, F], contiguous) row r = the F-element source frame at host address src_rows[r] (numpy uint64
    array; 0 = zero row), converted src_dtype -> dst.dtype (vc_host_stage_rows)."""
    if dst.device.type != "cpu" or not dst.is_contiguous() or dst.numel() != n_rows * F:
        raise ValueError("host_stage_rows: contiguous host destination of n_rows*F elements expected")
    if src_rows.dtype.name != "uint64" or src_rows.size != n_rows or not src_rows.flags["C_CONTIGUOUS"]:
        raise ValueError("host_stage_rows: src_rows must be a contiguous uint64 array of n_rows addresses")
    check(load_library().vc_host_stage_rows(ctypes.c_void_p(src_rows.ctypes.data), int(n_rows), int(F),
                                            _STAGE_DTYPES[src_dtype], ctypes.c_void_p(dst.data_ptr()),
                                            _STAGE_DTYPES[dst.dtype], int(threads)), "vc_host_stage_rows")


def convert_bf16(src: torch.Tensor, dst: torch.Tensor) -> None:
    """dst (device, bf16) = src (device, fp32), round to nearest even, on the current stream."""
    require_cuda(src, "src")
    require_cuda(dst, "dst")
    if src.dtype != torch.float32 or dst.dtype != torch.bfloat16 or src.numel() != dst.numel():
        raise ValueError("convert_bf16: fp32 source and bf16 destination of equal length expected")
    if not (src.is_contiguous() and dst.is_contiguous()):
        raise ValueError("convert_bf16: contiguous buffers expected")
    with torch.cuda.device(src.device):
        check(load_library().vc_convert_bf16(_ptr(src), _ptr(dst), src.numel(), _stream(src.device)), "vc_convert_bf16")


def linear(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, precision: str = "fp32",
           apply_tanh: bool = False) -> torch.Tensor:
    """``A @ W.T + bias`` through the native GEMM kernel of the given precision (parity-test entry)."""
    require_cuda(A, "A")
    lib = load_library()
    A = A.detach().float().contiguous()
    W = W.detach().float().contiguous()
    b = None if bias is None else bias.detach().float().contiguous()
    M, K = A.shape
    N = W.shape[0]
    C = torch.empty(M, N, dtype=torch.float32, device=A.device)
    ws = torch.empty(2 * (M * K + N * K) + 1024, dtype=torch.uint8, device=A.device)
    with torch.cuda.device(A.device):
        check(lib.vc_linear(PRECISION_IDS[precision], _ptr(A), _ptr(W), _ptr(b), _ptr(C), M, N, K, int(apply_tanh),
                            _ptr(ws), ws.numel(), _stream(A.device)), "vc_linear")
    return C


def beam_select(logits: torch.Tensor, scores: torch.Tensor, B: int, K: int):
    """One reference beam selection step (video_captioning_model.py:209-220) on the device."""
    require_cuda(logits, "logits")
    lib = load_library()
    lg = logits.detach().float().contiguous()
    sc = scores.detach().float().contiguous()
    V = lg.shape[1]
    R = B * K
    parent = torch.empty(R, dtype=torch.int32, device=lg.device)
    token = torch.empty(R, dtype=torch.int32, device=lg.device)
    new_scores = torch.empty(R, dtype=torch.float32, device=lg.device)
    ws = torch.empty(R * K * 8 + R * 32 + B * 32 + 8192, dtype=torch.uint8, device=lg.device)
    with torch.cuda.device(lg.device):
        check(lib.vc_beam_select(_ptr(lg), _ptr(sc), B, K, V, _ptr(parent), _ptr(token), _ptr(new_scores), _ptr(ws),
                                 ws.numel(), _stream(lg.device)), "vc_beam_select")
    return parent, token, new_scores
