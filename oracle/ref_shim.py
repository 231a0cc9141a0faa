"""TEST INFRASTRUCTURE ONLY -- import shim that loads the UNMODIFIED reference model files.

The reference package does not import as shipped (broken dataclass defaults
on py>=3.11, missing modules in ``src/config/__init__.py``, path hacks).  The four hot-path model
files only use ``Config`` as a type annotation, so they load untouched once ``src.config.config``
is stubbed.  Nothing is copied: the files are executed from where they lie under /root/reference.

Only usable where /root/reference exists (the build container).  ``available()`` says so.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from types import SimpleNamespace

# Where the unmodified reference sources are read from: $VC_REFERENCE_ROOT, else /root/reference (the build container),
# else oracle/_ref -- a git-ignored, byte-identical copy of the five files this shim executes, made by
# ``vendor_reference()`` (called from __graft_entry__.build() where /root/reference exists) so that bench.py's
# ``--impl reference`` arm can time the reference itself on the GPU box, where /root/reference does not exist.
_VENDOR_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_VENDORED = ("src/models/attention.py", "src/models/encoder.py", "src/models/decoder.py",
             "src/models/video_captioning_model.py", "src/data/vocabulary.py")


def _pick_root() -> str:
    cands = [os.environ.get("VC_REFERENCE_ROOT"), "/root/reference", _VENDOR_ROOT]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "src", "models", "video_captioning_model.py")):
            return c
    return cands[1]


REF_ROOT = _pick_root()
_REF_SRC = os.path.join(REF_ROOT, "src")
_loaded = {}


def available() -> bool:
    return os.path.isfile(os.path.join(_REF_SRC, "models", "video_captioning_model.py"))


def is_vendored_copy() -> bool:
    return os.path.abspath(REF_ROOT) == os.path.abspath(_VENDOR_ROOT)


def vendor_reference(src_root: str = "/root/reference") -> bool:
    """Copy the five reference files this shim executes, unmodified, into oracle/_ref (git-ignored; it travels to the
    GPU box with the working tree like the built .so).  Returns False where the reference is not mounted."""
    import shutil
    if not os.path.isfile(os.path.join(src_root, "src", "models", "video_captioning_model.py")):
        return False
    for rel in _VENDORED:
        dst = os.path.join(_VENDOR_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, rel), dst)
    return True


def _pkg(name: str) -> None:
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = []  # namespace-like, bypasses the broken __init__.py files
        sys.modules[name] = m


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load_reference():
    """Return a namespace with the reference's attention / encoder / decoder / model modules."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError(f"reference sources not found under {REF_ROOT}")
    for p in ("src", "src.config", "src.models", "src.data", "src.utils", "src.inference"):
        _pkg(p)
    cfgmod = types.ModuleType("src.config.config")

    class Config:  # the reference only uses it as an annotation
        pass

    cfgmod.Config = Config
    sys.modules["src.config.config"] = cfgmod
    att = _load("src.models.attention", f"{_REF_SRC}/models/attention.py")
    enc = _load("src.models.encoder", f"{_REF_SRC}/models/encoder.py")
    dec = _load("src.models.decoder", f"{_REF_SRC}/models/decoder.py")
    vcm = _load("src.models.video_captioning_model", f"{_REF_SRC}/models/video_captioning_model.py")
    ns = SimpleNamespace(attention=att, encoder=enc, decoder=dec, model=vcm)
    _loaded["ns"] = ns
    return ns


def load_reference_vocabulary():
    load_reference()
    if "vocab" not in _loaded:
        _loaded["vocab"] = _load("src.data.vocabulary", f"{_REF_SRC}/data/vocabulary.py")
    return _loaded["vocab"]


def build_reference_model(cfg, vocab_size: int, attention: str = "bahdanau", num_heads: int = 8,
                          state_dict=None):
    """Construct the reference VideoCaptioningModel (eval mode) and optionally load a state_dict.

    ``attention``: 'bahdanau' (what decoder.py:38 hard-codes), or 'luong_general' / 'luong_dot' /
    'luong_concat' / 'multihead', swapped in after construction exactly as SURVEY.md section 0 item 3
    describes (same call signature, attention.py:76,190).
    """
    import torch

    ns = load_reference()
    model = ns.model.VideoCaptioningModel(cfg, vocab_size)
    if attention == "bahdanau":
        pass
    elif attention.startswith("luong_"):
        model.decoder.attention = ns.attention.LuongAttention(cfg, attention.split("_", 1)[1])
    elif attention == "multihead":
        model.decoder.attention = ns.attention.MultiHeadAttention(cfg, num_heads)
    else:
        raise ValueError(attention)
    if state_dict is not None:
        sd = {k: torch.as_tensor(v) for k, v in state_dict.items()}
        model.load_state_dict(sd)
    return model.eval()


class _DiverseTorch:
    """Proxy for the ``torch`` module global of the reference's video_captioning_model.py: identical except
    that the beam-score initialisation ``torch.zeros(batch_size * beam_size, device=...)`` (:194, the only
    1-D ``zeros`` call with an int size in that file) yields [0, -inf, ..., -inf] per video.  This is the
    one-line repair the reference asks for (inference/predictor.py:353); every other line of the reference's
    beam loop runs unmodified."""

    def __init__(self, real, beam_size):
        self._real, self._k = real, beam_size

    def __getattr__(self, name):
        return getattr(self._real, name)

    def zeros(self, *size, **kw):
        z = self._real.zeros(*size, **kw)
        if len(size) == 1 and isinstance(size[0], int) and z.dim() == 1 and size[0] % self._k == 0 and not kw.get("dtype"):
            z = z.view(-1, self._k)
            z[:, 1:] = float("-inf")
            z = z.reshape(-1)
        return z


def reference_diverse_beam(model, feats_1, start_id, end_id, max_length, beam_size, length_penalty=1.0):
    """Run the UNMODIFIED reference ``_beam_search_generate`` for one video (B=1) with two minimal repairs
    applied from outside (no reference source is edited):
      1. the score initialisation is [0, -inf, ...] (see _DiverseTorch);
      2. ``decoder.forward_step`` receives ``encoder_outputs[:rows]`` / ``mask[:rows]``: once a hypothesis has
         completed the reference re-stacks only the live rows (:254-272) but keeps passing the K-row encoder
         tensors (:205-207), which raises a shape error (SURVEY.md section 3.3 iv).  For B=1 all K encoder rows are
         copies of the same video, so slicing is exact.
    Returns the reference's token row (its best completed hypothesis, :274-286)."""
    import torch

    ns = load_reference()
    mod = ns.model
    real = mod.torch
    orig_step = model.decoder.forward_step

    def step(input_token, hidden_state, encoder_outputs, encoder_mask=None):
        n = input_token.shape[0]
        return orig_step(input_token, hidden_state, encoder_outputs[:n],
                         None if encoder_mask is None else encoder_mask[:n])

    mod.torch = _DiverseTorch(real, beam_size)
    model.decoder.forward_step = step
    try:
        with torch.no_grad():
            return model.generate(feats_1, start_id, end_id, max_length=max_length, method="beam",
                                  beam_size=beam_size, length_penalty=length_penalty)["generated_tokens"][0]
    finally:
        mod.torch = real
        del model.decoder.forward_step
