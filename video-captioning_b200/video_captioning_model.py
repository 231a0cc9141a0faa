"""VideoCaptioningModel: the reference's model API over the native B200 path.

Same constructor, ``generate(...)`` signature, return dict and ``state_dict`` layout as the reference's
``src/models/video_captioning_model.py`` (ctor :13-33, forward :35-77, generate :79-125), so a reference
checkpoint loads with ``load_state_dict`` and ``predict.py``-style callers work unchanged.  All
arithmetic runs in ``libvc_b200.so``; there is no PyTorch fallback.

Semantics kept on purpose (SURVEY.md sections 3.3, 8a):
  * greedy: tokens [B, L<=S] without START; stops only when every row emits END in the same step
    (decoder.py:275); ``attention_weights`` [B, L, T].
  * beam: the reference initialises all K beam scores to 0 (:194) so beams tie and the result equals
    greedy with START prepended, truncated after the first END.  Its batched (B>1) bookkeeping is
    broken (:277-282, RuntimeError on staggered END), and the Predictor only ever calls it with B=1
    (predictor.py:102).  Batched contract here: row i == the reference's B=1 call on video i, rows
    right-padded with START (:288-300).  ``diverse_beams=True`` opts into standard beam search.
"""
from __future__ import annotations

import collections
import os
import threading
import time
import weakref
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _native
from .decoder import CaptionDecoder
from .encoder import VideoEncoder


def _attention_desc(att) -> Dict[str, int]:
    return {"attention": _native.ATTENTION_IDS[att.attention_kind], "num_heads": int(getattr(att, "num_heads", 1))}


def standalone_attention_handle(att) -> "_native.NativeModel":
    """Native handle for calling an attention module on its own (tests / explain): a minimal model
    whose non-attention weights are zeros."""
    p = next(att.parameters(), None)
    device = p.device if p is not None else torch.device("cuda")
    key = tuple((n, t._version, t.data_ptr()) for n, t in att.state_dict().items()) + (str(device),)
    cached = getattr(att, "_standalone", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    H, A = att.encoder_dim, att.attention_dim
    F = E = 64
    V = 8
    desc = dict(feature_dim=F, hidden_dim=H, embed_dim=E, attn_dim=A, vocab_size=V, enc_layers=1, dec_layers=1,
                precision=_native.PREC_FP32, **_attention_desc(att))
    z = lambda *s: torch.zeros(*s, device=device)
    sd = {"encoder.feature_projection.weight": z(H, F), "encoder.feature_projection.bias": z(H),
          "encoder.output_projection.weight": z(H, 2 * H), "encoder.output_projection.bias": z(H),
          "decoder.embedding.weight": z(V, E), "decoder.lstm.weight_ih_l0": z(4 * H, E + H),
          "decoder.lstm.weight_hh_l0": z(4 * H, H), "decoder.lstm.bias_ih_l0": z(4 * H), "decoder.lstm.bias_hh_l0": z(4 * H),
          "decoder.context_projection.weight": z(H, 2 * H + E), "decoder.context_projection.bias": z(H),
          "decoder.output_projection.weight": z(V, H), "decoder.output_projection.bias": z(V)}
    for sfx in ("", "_reverse"):
        sd[f"encoder.lstm.weight_ih_l0{sfx}"] = z(4 * H, H)
        sd[f"encoder.lstm.weight_hh_l0{sfx}"] = z(4 * H, H)
        sd[f"encoder.lstm.bias_ih_l0{sfx}"] = z(4 * H)
        sd[f"encoder.lstm.bias_hh_l0{sfx}"] = z(4 * H)
    for n, t in att.state_dict().items():
        sd["decoder.attention." + n] = t
    h = _native.NativeModel(desc, sd, device)
    att._standalone = (key, h)
    return h


class VideoCaptioningModel(nn.Module):
    """Encoder-decoder captioning model; parameters in the reference layout, compute in CUDA."""

    def __init__(self, config, vocabulary_size: int, attention_type: str = "bahdanau", num_heads: int = 8,
                 precision: str = "fp32", chunk_size: int = 2048):
        super().__init__()
        if precision not in _native.PRECISION_IDS:
            raise ValueError(f"precision must be one of {list(_native.PRECISION_IDS)}")
        self.config = config
        self.vocabulary_size = vocabulary_size
        self._check_dims(precision)
        self.encoder = VideoEncoder(config)
        self.decoder = CaptionDecoder(config, vocabulary_size, attention_type, num_heads)
        self.feature_extractor = None     # CNN extractors are out of scope (precomputed features)
        self.precision = precision
        self.chunk_size = int(chunk_size)  # videos per native call (bounds the workspace)
        self.host_chunk_size = 256         # videos per H2D chunk when the features live in host memory
        # bf16 mode: part of a host batch is rounded to bf16 on the host cores and crosses PCIe at half the size
        # (_generate_from_host_packed); VC_HOST_PACK=0 or host_pack=False keeps the plain fp32 transfer
        # "auto" (default): on when this rank has the node's host memory system mostly to itself.  Packing reads 1.31 MB and
        # writes 0.66 MB per video next to the DMA reads; measured: 1 rank on a 16-core box 41-42k vs 35k captions/s with
        # plain copies (wins); 4 ranks on a 32-core box 112k vs 137k for the node (loses); 2 ranks 72k (36k per GPU).
        _hp = os.environ.get("VC_HOST_PACK", "auto")
        _lw = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        self.host_pack = (_lw <= 2 and (os.cpu_count() or 1) // _lw >= 8) if _hp == "auto" else (_hp != "0")
        self.host_piece_size = int(os.environ.get("VC_HOST_PIECE", "64"))   # videos per transfer piece of the packed ingest
        self.host_inflight = 3             # piece copies queued ahead on the copy stream
        self.host_window_size = 1024       # videos staged on the device at a time (0.67 GB of bf16 at the MSVD shape)
        self.host_chunk_fractions = tuple(float(x) for x in os.environ.get("VC_HOST_CHUNKS", "0.5,0.8,1.0").split(","))
        self.host_pack_threads = int(os.environ.get("VC_HOST_PACK_THREADS", "0")) or max(
            1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
        self._packed = None
        self._packed_key = None
        self._copy_stream = None
        self._staging = None
        self._staging_key = None
        self._native_key = None
        self._native_handle: Optional[_native.NativeModel] = None
        self.encoder._owner = weakref.ref(self)

    # ------------------------------------------------------------------ native handle management
    def _check_dims(self, precision: str) -> None:
        """Size limits of the native kernels, raised where the model is built rather than at the first generate():
        F, H, E, A multiples of 8 (vectorised rows), 64 in bf16 mode (tensor-core k-blocks); the vocabulary may have
        any size (len(vocabulary) is arbitrary in reference checkpoints; it is padded internally)."""
        m = self.config.model
        if m.encoder_hidden_dim != m.decoder_hidden_dim:
            raise ValueError("encoder_hidden_dim != decoder_hidden_dim is not supported: the reference draws a fresh random "
                             "Linear on every call in that case (decoder.py:97-99), so there is nothing to reproduce")
        dims = dict(cnn_feature_dim=m.cnn_feature_dim, encoder_hidden_dim=m.encoder_hidden_dim,
                    embedding_dim=m.embedding_dim, attention_dim=m.attention_dim)
        mult = 8
        bad = {k: v for k, v in dims.items() if v <= 0 or v % mult}
        if precision == "bf16":
            bad.update({k: v for k, v in dims.items() if k != "attention_dim" and v % 64})
        if bad:
            raise ValueError(f"unsupported model dimensions for precision={precision}: {bad} (need multiples of 8; "
                             "cnn_feature_dim / hidden / embedding multiples of 64 in bf16 mode)")
        if self.vocabulary_size < 4:
            raise ValueError("vocabulary_size must be >= 4 (the four special tokens)")

    def set_precision(self, precision: str) -> "VideoCaptioningModel":
        if precision not in _native.PRECISION_IDS:
            raise ValueError(f"precision must be one of {list(_native.PRECISION_IDS)}")
        self._check_dims(precision)
        self.precision = precision
        return self

    def refresh_native(self) -> "VideoCaptioningModel":
        """Rebuild the native weight copies.  Needed only after parameters were changed behind autograd's back
        (``p.data.copy_()`` / ``p.data = ...`` do not bump the tensor version the handle cache watches);
        ``load_state_dict``, ``.to()`` and in-place ops on the parameters are picked up automatically."""
        self._native_handle = None
        self._native_key = None
        return self

    def _desc(self) -> Dict[str, int]:
        m = self.config.model
        d = dict(feature_dim=m.cnn_feature_dim, hidden_dim=m.encoder_hidden_dim, embed_dim=m.embedding_dim,
                 attn_dim=m.attention_dim, vocab_size=self.vocabulary_size, enc_layers=m.encoder_num_layers,
                 dec_layers=m.decoder_num_layers, precision=_native.PRECISION_IDS[self.precision])
        d.update(_attention_desc(self.decoder.attention))
        return d

    def _handle(self) -> _native.NativeModel:
        """(Re)build the native handle when parameters, device, precision or attention module changed."""
        sd = self.state_dict()
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError("VideoCaptioningModel must be moved to a CUDA device (model.to('cuda')); "
                               "this package has no CPU path")
        key = (self.precision, str(p.device), self.decoder.attention.attention_kind,
               tuple((n, t._version, t.data_ptr()) for n, t in sd.items()))
        if self._native_handle is None or key != self._native_key:
            self._native_handle = None
            self._native_handle = _native.NativeModel(self._desc(), sd, p.device)
            self._native_key = key
        return self._native_handle

    # ------------------------------------------------------------------ reference API
    def generate(self, video_features: torch.Tensor, start_token_id: int, end_token_id: int, max_length: int = 20,
                 video_mask: Optional[torch.Tensor] = None, method: str = "greedy", **kwargs) -> Dict[str, torch.Tensor]:
        """video_captioning_model.py:79-125.  kwargs: greedy ``temperature``; beam ``beam_size``,
        ``length_penalty``; opt-in, beyond the reference: ``diverse_beams=True`` (real beam search: only beam 0 is
        live at step 0 instead of K tied copies, :194) and ``num_return_sequences=N`` which adds the n-best lists
        ``nbest_tokens`` [B,N,L], ``nbest_lengths`` [B,N] (0 = none), ``nbest_scores`` [B,N] (what
        inference/predictor.py:353 asks the beam to return)."""
        if method not in ("greedy", "beam"):
            raise ValueError(f"Unsupported generation method: {method}")
        if method == "greedy":
            allowed = {"temperature"}
        else:
            allowed = {"beam_size", "length_penalty", "diverse_beams", "num_return_sequences"}
        bad = set(kwargs) - allowed
        if bad:
            raise TypeError(f"generate(method='{method}') got unexpected keyword arguments {sorted(bad)}")
        h = self._handle()
        B = video_features.shape[0]
        nbest = int(kwargs.get("num_return_sequences", 0) or 0)
        if B == 0:     # empty batch (e.g. an empty shard): well-formed empty results, no device work
            dev = h.device
            T = video_features.shape[1] if video_features.dim() == 3 else 0
            if method == "greedy":
                return {"generated_tokens": torch.zeros(0, 0, dtype=torch.int64, device=dev),
                        "attention_weights": torch.zeros(0, 0, T, device=dev)}
            res = {"generated_tokens": torch.zeros(0, 1, dtype=torch.int64, device=dev),
                   "lengths": torch.zeros(0, dtype=torch.int64, device=dev), "scores": torch.zeros(0, device=dev)}
            if nbest > 0:
                res.update(nbest_tokens=torch.zeros(0, nbest, 1, dtype=torch.int64, device=dev),
                           nbest_lengths=torch.zeros(0, nbest, dtype=torch.int64, device=dev),
                           nbest_scores=torch.zeros(0, nbest, device=dev))
            return res
        if nbest < 0 or nbest > 2 * int(kwargs.get("beam_size", 5)):
            raise ValueError("num_return_sequences must be in [0, 2*beam_size]")
        gen = lambda x, mk: h.generate(x, start_token_id, end_token_id, max_length, mk, method,
                                       beam_size=kwargs.get("beam_size", 5), length_penalty=kwargs.get("length_penalty", 1.0),
                                       temperature=kwargs.get("temperature", 1.0), diverse=kwargs.get("diverse_beams", False),
                                       nbest=nbest)
        if hasattr(video_features, "pack_piece"):     # a host piece source (VideoCaptionPredictor), bf16 mode
            if self.precision != "bf16":
                raise ValueError("host piece sources need precision='bf16'")
            outs = self._generate_from_host_packed(h, None, video_mask, gen, pack=True, source=video_features)
        elif video_features.device.type == "cpu":
            if video_features.dim() != 3 or video_features.dtype not in (torch.float32, torch.float16, torch.bfloat16):
                raise ValueError("host video_features must be a float32 / float16 / bfloat16 [B,T,F] tensor")
            if video_features.dtype == torch.bfloat16 and self.precision != "bf16":
                raise ValueError("bfloat16 host features need precision='bf16'")
            can_pack = self.precision == "bf16" and video_features.dtype == torch.float32
            if can_pack:
                # bf16 mode always goes through the piece pipeline (every feature rounded once to bf16, so the result does
                # not depend on the route); host_pack only decides whether host cores take part
                outs = self._generate_from_host_packed(h, video_features.contiguous(), video_mask, gen, pack=bool(self.host_pack))
            else:
                outs = self._generate_from_host(h, video_features, video_mask, gen)
        else:
            outs = []
            for lo in range(0, B, self.chunk_size):
                hi = min(B, lo + self.chunk_size)
                outs.append(gen(video_features[lo:hi], None if video_mask is None else video_mask[lo:hi]))
        tokens = torch.cat([o[0] for o in outs], dim=0)
        if method == "greedy":
            attn = torch.cat([o[3] for o in outs], dim=0)
            # decoder.py:275: the loop stops after the first step at which every row emitted END
            all_end = (tokens == end_token_id).all(dim=0)
            idx = torch.nonzero(all_end)
            L = int(idx[0, 0]) + 1 if idx.numel() else tokens.shape[1]
            return {"generated_tokens": tokens[:, :L].to(torch.int64), "attention_weights": attn[:, :L]}
        lens = torch.cat([o[1] for o in outs], dim=0)
        scores = torch.cat([o[2] for o in outs], dim=0)
        L = int(lens.max())
        res = {"generated_tokens": tokens[:, :L].to(torch.int64), "lengths": lens.to(torch.int64), "scores": scores}
        if nbest > 0:
            nl = torch.cat([o[5] for o in outs], dim=0)
            Ln = max(1, int(nl.max()))
            res.update(nbest_tokens=torch.cat([o[4] for o in outs], dim=0)[:, :, :Ln].to(torch.int64),
                       nbest_lengths=nl.to(torch.int64), nbest_scores=torch.cat([o[6] for o in outs], dim=0))
        return res

    def _generate_from_host(self, h, feats: torch.Tensor, mask, gen):
        """Feature ingest for HOST tensors (predictor.py:101-107 does one blocking H2D per video): the batch is
        streamed to the device in chunks on a copy stream, double-buffered, so the PCIe transfer of chunk
        i+1 overlaps the compute of chunk i.  Pass pinned memory (``tensor.pin_memory()``) for full speed.
        The arithmetic still runs only on the GPU."""
        # fp32 features cross the link as they are; features stored as halves (float16 .npy files load that way, the
        # reference's np.load -> tensor path is dtype-agnostic, predictor.py:101-102) or already rounded to bf16
        # (VideoCaptionPredictor's staging in bf16 mode) cross at half the bytes and are widened / rounded on the device
        dev = h.device
        want = torch.bfloat16 if self.precision == "bf16" and feats.dtype != torch.float32 else torch.float32
        B, T, F = feats.shape
        chunk = max(1, min(self.host_chunk_size, self.chunk_size, B))
        compute = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        key = (chunk, T, F, str(dev), feats.dtype)
        if self._staging_key != key:
            self._staging = [torch.empty(chunk, T, F, dtype=feats.dtype, device=dev) for _ in range(2)]
            self._staging_key = key
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        spans = self._host_spans(B, chunk)

        def enqueue_copy(i):
            lo, hi = spans[i]
            with torch.cuda.stream(self._copy_stream):
                if i >= 2:
                    self._copy_stream.wait_event(free[i % 2])
                else:
                    self._copy_stream.wait_stream(compute)      # earlier users of the staging buffers
                self._staging[i % 2][: hi - lo].copy_(feats[lo:hi], non_blocking=True)
                ready[i % 2].record(self._copy_stream)

        outs = []
        enqueue_copy(0)
        for i, (lo, hi) in enumerate(spans):
            if i + 1 < len(spans):
                enqueue_copy(i + 1)
            compute.wait_event(ready[i % 2])
            mk = None if mask is None else mask[lo:hi].to(dev, non_blocking=True)
            x = self._staging[i % 2][: hi - lo]
            if x.dtype != want:
                x = x.to(want)          # fp16 -> fp32 (exact) or fp16 -> bf16 (the rounding bf16 mode applies anyway)
            outs.append(gen(x, mk))
            free[i % 2].record(compute)
        return outs

    def _generate_from_host_packed(self, h, feats, mask, gen, pack: bool = True, source=None):
        """bf16-mode ingest of HOST fp32 features.  The PCIe link (54 GB/s, 1.3 MB per video) is the end-to-end
        bottleneck and bf16 mode rounds the features to bf16 anyway, so the batch is cut into pieces of
        ``host_piece_size`` videos that reach the device by one of two routes, whichever is free:
          * raw    : fp32 piece -> H2D -> vc_convert_bf16 on the copy stream (the link's share)
          * packed : vc_host_pack_bf16 on the host cores (a worker thread, GIL released) -> pinned bf16 -> H2D at half
                     the bytes (the cores' share)
        Within a chunk of ``host_chunk_size`` videos the copy loop takes raw pieces from the front and the packer takes
        pieces from the back until they meet; every piece lands in one device bf16 buffer and the chunk is decoded as
        soon as its pieces are in (vc_generate_ex, VC_DTYPE_BF16), while the next chunk is on its way.  All features are rounded exactly once, to nearest even, on
        either route, so results do not depend on the split.  Rows are returned in input order.

        ``source`` (instead of ``feats``): an object with ``shape`` = (B,T,F) and ``pack_piece(lo, hi, dst_bf16)`` that
        writes videos [lo,hi) rounded to bf16 into a pinned buffer -- the Predictor's list of per-video arrays, resized and
        rounded by vc_host_stage_rows piece by piece.  Every piece then takes the packed route (there is no contiguous
        fp32 batch to send raw), still overlapped with the copies and the decode of the previous chunk."""
        dev = h.device
        B, T, F = source.shape if source is not None else feats.shape
        piece = max(1, min(self.host_piece_size, self.chunk_size, B))
        window = min(B, max(piece, min(self.chunk_size, self.host_window_size) // piece * piece))
        compute = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        copy = self._copy_stream
        key = (window, piece, T, F, str(dev))
        if self._packed_key != key:
            self._packed = dict(
                dev16=torch.empty(window, T, F, dtype=torch.bfloat16, device=dev),
                raw=[torch.empty(piece, T, F, dtype=torch.float32, device=dev) for _ in range(2)],
                host16=[torch.empty(piece, T, F, dtype=torch.bfloat16).pin_memory() for _ in range(4)])
            self._packed_key = key
        st = self._packed
        pinned = feats.is_pinned() if source is None else False
        outs = {}
        chunk_free = {}                      # chunk slot of the window -> event: its last decode has read dev16
        raw_free = [None, None]
        for w0 in range(0, B, window):
            w1 = min(B, w0 + window)
            # the device buffer is reused by every window (and call): its previous readers -- all decodes enqueued so far --
            # must be done before the first piece of this window lands (chunk boundaries differ between windows)
            copy.wait_stream(compute)
            chunk_free.clear()
            chunks = self._packed_chunks(w0, w1, piece)
            pieces, span, owner = [], [], []
            for ci, (clo, chi) in enumerate(chunks):
                first = len(pieces)
                for lo in range(clo, chi, piece):
                    pieces.append((lo, min(chi, lo + piece)))
                    owner.append(ci)
                span.append([first, len(pieces)])
            left = [b - a for a, b in span]   # pieces of each chunk not yet enqueued
            lock = threading.Lock()
            # per chunk: [next raw piece (from its front), one past the last unclaimed piece (the packer takes from its
            # back)]; both sides work on the earliest chunk that still has unclaimed pieces, so chunks complete -- and
            # are decoded -- one after the other while the transfer of the later ones is still running

            def claim(from_back):              # call with the lock held
                for fb in span:
                    if fb[0] < fb[1]:
                        if from_back:
                            fb[1] -= 1
                            return fb[1]
                        fb[0] += 1
                        return fb[0] - 1
                return None

            ready = collections.deque()
            have = threading.Semaphore(0)
            bufs = collections.deque((b, None) for b in st["host16"])     # (pinned bf16 buffer, event of its last H2D)
            buf_sem = threading.Semaphore(len(st["host16"]))
            errors = []

            def packer():
                try:
                    while True:
                        buf_sem.acquire()          # a free pinned buffer first: a claimed piece is a commitment
                        with lock:
                            idx = claim(True)
                        if idx is None:
                            buf_sem.release()
                            return
                        with lock:
                            buf, ev = bufs.popleft()
                        if ev is not None:
                            ev.synchronize()
                        lo, hi = pieces[idx]
                        t0 = time.perf_counter()
                        if source is not None:
                            source.pack_piece(lo, hi, buf[: hi - lo])
                        else:
                            _native.host_pack_bf16(feats[lo:hi], buf[: hi - lo], self.host_pack_threads)
                        stats["pack_s"] += time.perf_counter() - t0
                        stats["packed"] += 1
                        with lock:
                            ready.append((idx, buf))
                        have.release()
                except Exception as e:      # surfaced by the copy loop
                    errors.append(e)
                    have.release()

            stats = dict(pack_s=0.0, packed=0, raw=0, sync_s=0.0, idle_s=0.0, h2d_bytes=0, t0=time.perf_counter())
            if w0 > 0:                        # later windows of the same call accumulate
                stats["h2d_bytes"] = self.host_stats.get("h2d_bytes", 0)
            self.host_stats = stats
            worker = threading.Thread(target=packer if pack else (lambda: None), daemon=True)
            worker.start()
            inflight = collections.deque()     # events of enqueued piece copies, oldest first
            n_raw = 0
            done = 0
            while done < len(pieces):
                if errors:
                    raise errors[0]
                item = None
                with lock:
                    if ready:
                        item = ("packed",) + ready.popleft()
                    elif source is None:
                        idx = claim(False)
                        if idx is not None:
                            item = ("raw", idx, None)
                if item is None:               # the packer holds the last pieces
                    t0 = time.perf_counter()
                    have.acquire()
                    stats["idle_s"] += time.perf_counter() - t0
                    continue
                if item[0] == "packed":
                    have.acquire(blocking=False)
                kind, idx, buf = item
                lo, hi = pieces[idx]
                ci = owner[idx]
                dst = st["dev16"][lo - w0: hi - w0]
                with torch.cuda.stream(copy):
                    if ci in chunk_free:
                        copy.wait_event(chunk_free.pop(ci))
                    stats["h2d_bytes"] += (hi - lo) * T * F * (2 if kind == "packed" else 4)
                    if kind == "packed":
                        dst.copy_(buf[: hi - lo], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy)
                        with lock:
                            bufs.append((buf, ev))
                        buf_sem.release()
                    else:
                        slot = n_raw % 2
                        n_raw += 1
                        if raw_free[slot] is not None:
                            copy.wait_event(raw_free[slot])
                        src = feats[lo:hi]
                        st["raw"][slot][: hi - lo].copy_(src, non_blocking=pinned)
                        _native.convert_bf16(st["raw"][slot][: hi - lo], dst)
                        raw_free[slot] = torch.cuda.Event()
                        raw_free[slot].record(copy)
                        ev = raw_free[slot]
                    inflight.append(ev)
                    left[ci] -= 1
                    if left[ci] == 0:          # chunk complete: decode it
                        rdy = torch.cuda.Event()
                        rdy.record(copy)
                        compute.wait_event(rdy)
                        clo, chi = chunks[ci]
                        with torch.cuda.stream(compute):
                            # the mask is transferred on the COMPUTE stream: its consumers (lengths, encoder, every
                            # attention step) run there, and its memory then belongs to that stream's allocator pool
                            mk = None if mask is None else mask[clo:chi].to(dev, non_blocking=True)
                            outs[clo] = gen(st["dev16"][clo - w0: chi - w0], mk)
                            fr = torch.cuda.Event()
                            fr.record(compute)
                        chunk_free[ci] = fr
                done += 1
                # pace the loop to the link: at most two piece copies queued, so that a piece packed meanwhile is
                # sent instead of the next raw one
                t0 = time.perf_counter()
                while len(inflight) > self.host_inflight:
                    inflight.popleft().synchronize()
                stats["sync_s"] += time.perf_counter() - t0
            worker.join()
            stats["raw"] = n_raw
            stats["loop_s"] = time.perf_counter() - stats.pop("t0")
            if errors:
                raise errors[0]
        return [outs[k] for k in sorted(outs)]

    def _packed_chunks(self, w0: int, w1: int, piece: int):
        """Decode chunks of one window of the packed ingest.  A chunk's decode has a latency floor (160 dependent encoder
        steps, 20 decode steps: ~4 ms however few videos it holds), so few large chunks beat many small ones; the last
        chunk is the smaller one because its decode is the part of the step no transfer overlaps.  ``host_chunk_fractions``
        are cumulative fractions of the window, rounded to whole pieces."""
        n = w1 - w0
        cuts = sorted({min(n, max(piece, int(round(f * n / piece)) * piece)) for f in self.host_chunk_fractions} | {n})
        lo, out = 0, []
        for c in cuts:
            if c > lo:
                out.append((w0 + lo, w0 + c))
                lo = c
        return out

    @staticmethod
    def _host_spans(B: int, chunk: int):
        """Chunk schedule of the host-feature pipeline: uniform chunks.  The transfer is the bottleneck (1.3 MB per
        video over PCIe), so the wall time is ~ total copy time + the compute of the last chunk.  Tapering the last
        chunks (128/64/64 videos) was measured and is WORSE (27.9k vs 35k captions/s): a chunk's compute has a
        latency floor (160 dependent encoder steps, 20 decode steps) that small chunks do not amortise."""
        return [(lo, min(B, lo + chunk)) for lo in range(0, B, chunk)]

    def forward(self, video_features, input_tokens, target_tokens, video_mask=None) -> Dict[str, torch.Tensor]:
        """Teacher-forced forward, video_captioning_model.py:35-77 (inference only: no autograd graph)."""
        h = self._handle()
        logits, attn, enc_out = h.forward_teacher(video_features, input_tokens, video_mask)
        return {"logits": logits, "encoder_outputs": enc_out, "attention_weights": attn, "target_tokens": target_tokens}

    def get_trainable_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def freeze_encoder(self) -> None:
        for p in self.encoder.parameters():
            p.requires_grad = False

    def unfreeze_encoder(self) -> None:
        for p in self.encoder.parameters():
            p.requires_grad = True
