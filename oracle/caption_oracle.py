"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference caption-generation path.

Closed-form torch-CPU arithmetic (plain ``@`` + sigmoid/tanh; no ``nn.LSTM``/``nn.Linear`` modules)
of angadbawa/Video-Captioning's batched caption-generation path.  Each function cites the reference
file:line it follows (paths relative to /root/reference/src).  It is the checker for the CUDA
path and the "port" CPU baseline of bench.py; the product never imports it.

Parity pinning: the reference has no tests/golden vectors.  This restatement
is pinned against outputs of the *unmodified reference modules* run in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle_vs_reference.py`` also
compares live wherever /root/reference exists).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch


def attention_type_of(sd) -> str:
    """Infer the attention variant from state_dict keys (SURVEY.md section 8b key table)."""
    if "decoder.attention.encoder_projection.weight" in sd:
        return "bahdanau"
    if "decoder.attention.linear_in.weight" in sd:
        return "luong_general"
    if "decoder.attention.linear_query.weight" in sd:
        return "luong_concat"
    if "decoder.attention.query_linear.weight" in sd:
        return "multihead"
    return "luong_dot"


class CaptionOracle:
    """Holds a reference-layout state_dict and evaluates the hot path on the CPU."""

    def __init__(self, state_dict, num_encoder_layers=2, num_decoder_layers=2, num_heads=8,
                 dtype=torch.float32):
        self.dtype = dtype
        self.p = {k: torch.as_tensor(np.asarray(v)).to(dtype) if not torch.is_tensor(v) else v.to(dtype)
                  for k, v in state_dict.items()}
        self.Le = num_encoder_layers
        self.Ld = num_decoder_layers
        self.num_heads = num_heads
        self.attention = attention_type_of(self.p)
        self.H = self.p["encoder.output_projection.weight"].shape[0]
        self.V = self.p["decoder.output_projection.weight"].shape[0]

    # ------------------------------------------------------------------ encoder (a1, a1')
    @staticmethod
    def _lstm_cell(gates, c):
        """PyTorch gate order i,f,g,o (nn.LSTM as used at encoder.py:35-42 / decoder.py:44-50)."""
        H = c.shape[-1]
        i = torch.sigmoid(gates[..., 0 * H:1 * H])
        f = torch.sigmoid(gates[..., 1 * H:2 * H])
        g = torch.tanh(gates[..., 2 * H:3 * H])
        o = torch.sigmoid(gates[..., 3 * H:4 * H])
        c2 = f * c + i * g
        return o * torch.tanh(c2), c2

    def _run_direction(self, x, w_ih, w_hh, b_ih, b_hh, reverse, lengths):
        """One direction of one bi-LSTM layer over x [B,T,in]; zero initial state (encoder.py:84).

        With ``lengths`` it follows pack_padded_sequence semantics (encoder.py:74-82): each row only
        runs over its first ``len`` frames (the reverse direction starts at frame len-1), outputs
        beyond ``len`` are zero and the returned final state is the state after the row's last
        valid step.
        """
        B, T, _ = x.shape
        H = w_hh.shape[1]
        xp = x @ w_ih.t() + b_ih  # all-timestep input projection
        h = x.new_zeros(B, H)
        c = x.new_zeros(B, H)
        out = x.new_zeros(B, T, H)
        steps = range(T - 1, -1, -1) if reverse else range(T)
        for t in steps:
            gates = xp[:, t] + (h @ w_hh.t() + b_hh)
            h2, c2 = self._lstm_cell(gates, c)
            if lengths is not None:
                valid = (t < lengths).to(x.dtype).unsqueeze(1)
                h = valid * h2 + (1 - valid) * h
                c = valid * c2 + (1 - valid) * c
                out[:, t] = valid * h2
            else:
                h, c = h2, c2
                out[:, t] = h2
        return out, h

    def encode(self, feats: torch.Tensor, mask: Optional[torch.Tensor] = None
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        """VideoEncoder.forward, encoder.py:52-98 (dropout = identity in eval mode)."""
        p = self.p
        x = feats.to(self.dtype)
        lengths = None
        if mask is not None:
            lengths = mask.sum(dim=1).to(torch.long)                         # encoder.py:75
        proj = x @ p["encoder.feature_projection.weight"].t() + p["encoder.feature_projection.bias"]  # :70
        layer_in = proj
        h_last = None
        for layer in range(self.Le):
            outs, finals = [], []
            for sfx, rev in (("", False), ("_reverse", True)):
                o, hN = self._run_direction(
                    layer_in, p[f"encoder.lstm.weight_ih_l{layer}{sfx}"], p[f"encoder.lstm.weight_hh_l{layer}{sfx}"],
                    p[f"encoder.lstm.bias_ih_l{layer}{sfx}"], p[f"encoder.lstm.bias_hh_l{layer}{sfx}"], rev, lengths)
                outs.append(o)
                finals.append(hN)
            layer_in = torch.cat(outs, dim=2)                                 # [fwd ; bwd]
            h_last = finals
        lstm_out = layer_in
        if lengths is not None:
            lstm_out = lstm_out[:, : int(lengths.max())]                      # pad_packed pads to batch max, :80-82
        w_o, b_o = p["encoder.output_projection.weight"], p["encoder.output_projection.bias"]
        enc_out = lstm_out @ w_o.t() + b_o                                    # :87
        final = torch.cat(h_last, dim=1) @ w_o.t() + b_o                      # :92-96 (hidden[-2:], same W_o)
        return enc_out, final

    # ------------------------------------------------------------------ attention (a3, a4, a5)
    def precompute_keys(self, enc_out):
        """Loop-invariant projections (attention.py:52 / :140 / :241-242), hoisted out of the step."""
        p, a = self.p, self.attention
        if a == "bahdanau":
            return (enc_out @ p["decoder.attention.encoder_projection.weight"].t()
                    + p["decoder.attention.encoder_projection.bias"],)
        if a == "luong_concat":
            return (enc_out @ p["decoder.attention.linear_context.weight"].t()
                    + p["decoder.attention.linear_context.bias"],)
        if a == "multihead":
            return (enc_out @ p["decoder.attention.key_linear.weight"].t() + p["decoder.attention.key_linear.bias"],
                    enc_out @ p["decoder.attention.value_linear.weight"].t() + p["decoder.attention.value_linear.bias"])
        return ()

    def attend(self, enc_out, h, mask, pre=None):
        """(context, weights) for one query per row.  attention.py:32-73, 103-187, 220-275."""
        p, a = self.p, self.attention
        if pre is None:
            pre = self.precompute_keys(enc_out)
        if a == "bahdanau":
            q = h @ p["decoder.attention.decoder_projection.weight"].t() + p["decoder.attention.decoder_projection.bias"]
            comb = torch.tanh(pre[0] + q.unsqueeze(1))                        # :56
            s = (comb @ p["decoder.attention.attention_linear.weight"].t()).squeeze(-1) \
                + p["decoder.attention.attention_linear.bias"]                # :57
        elif a == "luong_dot":
            s = torch.einsum("rh,rth->rt", h, enc_out)                        # :118-125
        elif a == "luong_general":
            q = h @ p["decoder.attention.linear_in.weight"].t()               # :128 (no bias)
            s = torch.einsum("rh,rth->rt", q, enc_out)                        # :129-132
        elif a == "luong_concat":
            q = h @ p["decoder.attention.linear_query.weight"].t() + p["decoder.attention.linear_query.bias"]
            comb = torch.tanh(q.unsqueeze(1) + pre[0])                        # :145
            s = (comb @ p["decoder.attention.linear_v.weight"].t()).squeeze(-1)  # :146 (no bias)
        elif a == "multihead":
            R, T, H = enc_out.shape
            n, d = self.num_heads, H // self.num_heads
            q = h @ p["decoder.attention.query_linear.weight"].t() + p["decoder.attention.query_linear.bias"]
            Q = q.view(R, n, 1, d)
            K = pre[0].view(R, T, n, d).transpose(1, 2)
            Vv = pre[1].view(R, T, n, d).transpose(1, 2)
            s = (Q @ K.transpose(-2, -1)) / (d ** 0.5)                        # :250  [R,n,1,T]
            if mask is not None:
                s = s.masked_fill(mask[:, None, None, :] == 0, -1e9)          # :253-255
            w = torch.softmax(s, dim=-1)                                      # :258
            ctx = (w @ Vv).transpose(1, 2).reshape(R, H)                      # :262-267
            ctx = ctx @ p["decoder.attention.output_linear.weight"].t() + p["decoder.attention.output_linear.bias"]
            return ctx, w.mean(dim=1).squeeze(1)                              # :270-273
        else:
            raise ValueError(a)
        if mask is not None:
            s = s.masked_fill(mask == 0, -1e9)                                # :61 / :174-175
        w = torch.softmax(s, dim=1)                                           # :64 / :178
        ctx = torch.einsum("rt,rth->rh", w, enc_out)                          # :68-71 / :182-185
        return ctx, w

    # ------------------------------------------------------------------ decoder step (a2, a6)
    def init_hidden(self, final):
        """CaptionDecoder.init_hidden_state, decoder.py:81-106: h0[l]=final, c0=0."""
        h = final.unsqueeze(0).repeat(self.Ld, 1, 1)
        return h, torch.zeros_like(h)

    def forward_step(self, tok, state, enc_out, mask, pre=None):
        """CaptionDecoder.forward_step, decoder.py:108-171.  tok: [R] int64."""
        p = self.p
        h, c = state
        emb = p["decoder.embedding.weight"][tok]                              # :130
        ctx, w = self.attend(enc_out, h[-1], mask, pre)                       # :135-138 (top layer, prev step)
        x = torch.cat([emb, ctx], dim=1)                                      # :143-146
        hs, cs = [], []
        for layer in range(self.Ld):                                          # :152
            gates = (x @ p[f"decoder.lstm.weight_ih_l{layer}"].t() + p[f"decoder.lstm.bias_ih_l{layer}"]) \
                + (h[layer] @ p[f"decoder.lstm.weight_hh_l{layer}"].t() + p[f"decoder.lstm.bias_hh_l{layer}"])
            h2, c2 = self._lstm_cell(gates, c[layer])
            hs.append(h2)
            cs.append(c2)
            x = h2
        pin = torch.cat([hs[-1], ctx, emb], dim=1)                            # :157-161
        o = torch.tanh(pin @ p["decoder.context_projection.weight"].t() + p["decoder.context_projection.bias"])  # :164-165
        logits = o @ p["decoder.output_projection.weight"].t() + p["decoder.output_projection.bias"]            # :169
        return logits, (torch.stack(hs), torch.stack(cs)), w

    # ------------------------------------------------------------------ greedy (a7)
    def greedy(self, feats, start_id, end_id, max_length=20, mask=None, temperature=1.0,
               return_logits=False) -> Dict[str, torch.Tensor]:
        """generate(method='greedy'): video_captioning_model.py:104-140 + decoder.py:223-289."""
        feats = torch.as_tensor(feats)
        enc_out, final = self.encode(feats, mask)
        B = feats.shape[0]
        if mask is None:
            mask = torch.ones(B, feats.shape[1], dtype=self.dtype)            # :108-112
        pre = self.precompute_keys(enc_out)
        state = self.init_hidden(final)
        tok = torch.full((B,), start_id, dtype=torch.long)
        toks, ws, lg = [], [], []
        for _ in range(max_length):
            logits, state, w = self.forward_step(tok, state, enc_out, mask, pre)
            if temperature != 1.0:
                logits = logits / temperature                                 # decoder.py:265
            nxt = torch.argmax(logits, dim=1)                                 # :269
            toks.append(nxt)
            ws.append(w)
            lg.append(logits)
            if bool((nxt == end_id).all()):                                   # :275
                break
            tok = nxt
        out = {"generated_tokens": torch.stack(toks, dim=1), "attention_weights": torch.stack(ws, dim=1)}
        if return_logits:
            out["logits"] = torch.stack(lg, dim=1)
        return out

    # ------------------------------------------------------------------ beam (a9), B=1 semantics
    def _beam_one(self, enc_out, final, mask, start_id, end_id, max_length, K, length_penalty, diverse=False):
        """_beam_search_generate for ONE video, video_captioning_model.py:148-302.

        Restates the reference literally for batch_size == 1 (the only well-defined case,
        SURVEY.md section 3.3): scores start at ZERO for all K beams (:194) so all beams tie.

        ``diverse=True`` is the repair the reference itself asks for (inference/predictor.py:353 "modify beam
        search to return multiple hypotheses"): the SAME loop with ``scores[1:] = -inf`` at step 0, so the K
        rows stop being copies of each other; completed hypotheses go to a per-video list with the
        length-normalised score of :237-242, which now matters.  Nothing else changes.

        Returns a dict: ``best`` (tokens incl. leading START), ``best_score`` (normalised score of the best
        completed hypothesis, else the raw score of live beam 0, :274-286), ``step_scores`` (per-step top-K
        candidate scores), ``nbest`` = [(tokens, score)]: the K best completed hypotheses by normalised score
        (descending, first-completed first among equals -- what ``max`` at :277-281 keeps), followed by the
        beams still live after ``max_length`` steps (score / generated_length ** length_penalty).
        """
        V = self.V
        enc = enc_out.expand(K, -1, -1).contiguous()                          # :179-181
        msk = mask.expand(K, -1).contiguous()                                 # :187-189
        pre = tuple(t.expand(K, -1, -1).contiguous() for t in self.precompute_keys(enc_out))
        seqs = torch.full((K, 1), start_id, dtype=torch.long)                 # :191-193
        scores = torch.zeros(K, dtype=self.dtype)                             # :194
        if diverse:
            scores[1:] = float("-inf")
        state = self.init_hidden(final.expand(K, -1).contiguous())            # :196
        completed: List[Tuple[torch.Tensor, float]] = []
        step_scores = []
        live = True
        for _ in range(max_length):                                           # :202
            R = seqs.shape[0]
            logits, state, _ = self.forward_step(seqs[:, -1], state, enc[:R], msk[:R], tuple(t[:R] for t in pre))
            logp = torch.log_softmax(logits, dim=-1)                          # :209
            cand = (scores.unsqueeze(1) + logp).view(1, -1)                   # :211-213
            top_s, top_i = torch.topk(cand, K, dim=1)                         # :215
            beam_i = top_i // V                                               # :219
            tok_i = top_i % V                                                 # :220
            step_scores.append(top_s[0].clone())
            new_seqs, new_scores, keep = [], [], []
            for k in range(K):                                                # :226-249
                ob = int(beam_i[0, k])
                t = int(tok_i[0, k])
                ns = torch.cat([seqs[ob], torch.tensor([t])])
                if t == end_id:
                    completed.append((ns, float(top_s[0, k]) / ((len(ns) - 1) ** length_penalty)))  # :237-242
                else:
                    new_seqs.append(ns)
                    new_scores.append(top_s[0, k])
                    keep.append(ob)
            if not new_seqs:                                                  # :251
                live = False
                break
            seqs = torch.stack(new_seqs)                                      # :254-266 (equal lengths for B=1)
            scores = torch.stack(new_scores)                                  # :267
            idx = torch.tensor(keep)
            state = (state[0][:, idx].clone(), state[1][:, idx].clone())      # :269-272
        if completed:                                                         # :274-282
            best, best_score = max(completed, key=lambda x: x[1])
        else:
            best, best_score = seqs[0], float(scores[0])                      # :286
        nbest = sorted(completed, key=lambda x: -x[1])[:K]                    # stable: first-completed first among equals
        if live:
            gen = seqs.shape[1] - 1
            nbest += [(seqs[k], float(scores[k]) / (gen ** length_penalty)) for k in range(seqs.shape[0])]
        return {"best": best, "best_score": best_score, "step_scores": step_scores, "nbest": nbest}

    def beam(self, feats, start_id, end_id, max_length=20, mask=None, beam_size=5, length_penalty=1.0,
             return_scores=False, diverse=False, num_return=0) -> Dict[str, torch.Tensor]:
        """Batched contract: row i == the reference's B=1 call on video i
        (predictor.py:102 always calls with B=1); rows right-padded with START (:288-300).
        ``num_return`` > 0 adds the n-best lists (see _beam_one): ``nbest_tokens`` [B,N,L] START-padded,
        ``nbest_lengths`` [B,N] (0 = no such hypothesis), ``nbest_scores`` [B,N] (-inf there)."""
        feats = torch.as_tensor(feats)
        enc_out, final = self.encode(feats, mask)
        B = feats.shape[0]
        if mask is None:
            mask = torch.ones(B, feats.shape[1], dtype=self.dtype)
        res1 = [self._beam_one(enc_out[b:b + 1], final[b:b + 1], mask[b:b + 1], start_id, end_id,
                               max_length, beam_size, length_penalty, diverse) for b in range(B)]
        rows = [r["best"] for r in res1]
        L = max(len(r) for r in rows)
        out = torch.full((B, L), start_id, dtype=torch.long)
        lens = torch.zeros(B, dtype=torch.long)
        for b, r in enumerate(rows):
            out[b, : len(r)] = r
            lens[b] = len(r)
        res = {"generated_tokens": out, "lengths": lens,
               "scores": torch.tensor([r["best_score"] for r in res1], dtype=torch.float64)}
        if return_scores:
            res["step_scores"] = [r["step_scores"] for r in res1]
        if num_return > 0:
            N = num_return
            nt = torch.full((B, N, max_length + 1), start_id, dtype=torch.long)
            nl = torch.zeros(B, N, dtype=torch.long)
            nsc = torch.full((B, N), float("-inf"), dtype=torch.float64)
            for b, r in enumerate(res1):
                for j, (t, sc) in enumerate(r["nbest"][:N]):
                    nt[b, j, : len(t)] = t
                    nl[b, j] = len(t)
                    nsc[b, j] = sc
            res.update(nbest_tokens=nt, nbest_lengths=nl, nbest_scores=nsc)
        return res

    def sequence_logprob(self, feats, tokens, lengths, mask=None) -> torch.Tensor:
        """Sum of log-softmax probabilities of given START-prefixed rows (teacher-forced through
        forward_step): the un-normalised beam score of video_captioning_model.py:209-213 for a FIXED token
        sequence.  tokens [B,L] (index 0 = START), lengths [B] (incl. START) -> [B] float64."""
        tokens = torch.as_tensor(tokens).long()
        lengths = torch.as_tensor(lengths).long()
        n = int(lengths.max()) - 1
        lg = self.forward_teacher(feats, tokens[:, :n], mask)["logits"]
        lp = torch.log_softmax(lg.double(), dim=-1)
        tgt = tokens[:, 1:n + 1]
        g = lp.gather(2, tgt.unsqueeze(-1)).squeeze(-1)
        valid = torch.arange(n)[None, :] < (lengths - 1)[:, None]
        return (g * valid).sum(dim=1)

    # ------------------------------------------------------------------ teacher forcing (f1)
    def forward_teacher(self, feats, input_tokens, mask=None) -> Dict[str, torch.Tensor]:
        """VideoCaptioningModel.forward / CaptionDecoder.forward, video_captioning_model.py:35-77,
        decoder.py:173-221: logits [B,L,V], attention [B,L,T]."""
        feats = torch.as_tensor(feats)
        input_tokens = torch.as_tensor(input_tokens)
        enc_out, final = self.encode(feats, mask)
        B = feats.shape[0]
        if mask is None:
            mask = torch.ones(B, feats.shape[1], dtype=self.dtype)
        pre = self.precompute_keys(enc_out)
        state = self.init_hidden(final)
        lg, ws = [], []
        for t in range(input_tokens.shape[1]):
            logits, state, w = self.forward_step(input_tokens[:, t], state, enc_out, mask, pre)
            lg.append(logits)
            ws.append(w)
        return {"logits": torch.stack(lg, dim=1), "attention_weights": torch.stack(ws, dim=1),
                "encoder_outputs": enc_out, "encoder_final": final}


# ---------------------------------------------------------------------- host-side helpers (a10, a11)
def resize_features(feats: np.ndarray, target_length: int) -> np.ndarray:
    """predictor.py:292-315 for one [T',F] array: uniform subsample at floor(linspace) or zero-pad."""
    Tp, F = feats.shape
    if Tp == target_length:
        return feats
    if Tp > target_length:
        idx = torch.linspace(0, Tp - 1, target_length, dtype=torch.long).numpy()   # :310
        return feats[idx]
    pad = np.zeros((target_length - Tp, F), dtype=feats.dtype)                     # :314-315
    return np.concatenate([feats, pad], axis=0)


def decode_caption(tokens, idx2word, pad_token="<PAD>", start_token="<START>", end_token="<END>",
                   remove_special_tokens=True) -> str:
    """Vocabulary.decode_caption, vocabulary.py:161-194 (compares the word strings, :183-189)."""
    words = []
    for t in tokens:
        t = int(t)
        if t not in idx2word:                                                       # :179 unknown ids dropped
            continue
        w = idx2word[t]
        if remove_special_tokens and w in (pad_token, start_token, end_token):      # :183-186 skip, no stop
            continue
        if w == end_token:                                                          # :189 only reachable w/o removal
            break
        words.append(w)
    return " ".join(words)
