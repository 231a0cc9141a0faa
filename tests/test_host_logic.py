"""CPU: host-side logic of the drop-in boundary (state_dict layout, feature resize, vocabulary,
sharding + the world_size-2 gloo gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import video_captioning_b200 as vc
from oracle import synth
from oracle.caption_oracle import decode_caption as oracle_decode
from oracle.caption_oracle import resize_features as oracle_resize
from video_captioning_b200.predictor import resize_features
from video_captioning_b200.sharding import gather_captions, shard_bounds


@pytest.mark.parametrize("att", synth.ATTENTION_TYPES)
def test_state_dict_layout_matches_reference_keys(att):
    cfg = synth.make_config("tiny")
    sd = synth.make_state_dict(cfg, 1000, att)
    m = vc.VideoCaptioningModel(cfg, 1000, attention_type=att)
    own = m.state_dict()
    assert sorted(own) == sorted(sd)
    for k in sd:
        assert tuple(own[k].shape) == sd[k].shape, k


def test_swapping_attention_module_like_the_oracle_does():
    cfg = synth.make_config("tiny")
    m = vc.VideoCaptioningModel(cfg, 1000)
    m.decoder.attention = vc.LuongAttention(cfg, "dot")
    assert m._desc()["attention"] == 1
    m.decoder.attention = vc.MultiHeadAttention(cfg, 4)
    assert m._desc()["attention"] == 4 and m._desc()["num_heads"] == 4
    with pytest.raises(ValueError):
        vc.create_attention_mechanism(cfg, "nope")


def test_mismatched_hidden_dims_rejected():
    cfg = synth.make_config("tiny")
    cfg.model.decoder_hidden_dim = 64
    with pytest.raises(ValueError, match="not supported"):
        vc.VideoCaptioningModel(cfg, 1000)


@pytest.mark.parametrize("tp", [5, 16, 17, 80, 200, 1000])
def test_resize_features_matches_oracle(tp):
    x = np.random.default_rng(tp).standard_normal((tp, 7)).astype(np.float32)
    assert np.array_equal(resize_features(x, 16), oracle_resize(x, 16))
    assert resize_features(x, 16).shape == (16, 7)


def test_vocabulary_decode_matches_oracle():
    v = vc.Vocabulary.from_words(["a", "man", "is", "running"])
    assert len(v) == 8
    for toks in ([1, 4, 5, 2, 6, 7, 0, 99], [4, 5, 6], [], [2, 4], [1, 1, 2]):
        for rm in (True, False):
            assert v.decode_caption(toks, rm) == oracle_decode(toks, v.idx2word, remove_special_tokens=rm)
    v2 = vc.Vocabulary.from_package(v.to_package())
    assert v2.word2idx == v.word2idx and v2.end_idx == 2


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 65536, 1000):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 3 if rank == 0 else 2
    L = 4 + rank
    toks = torch.full((n, L), 10 + rank, dtype=torch.int64)
    lens = torch.full((n,), L, dtype=torch.int64)
    t, l = gather_captions(toks, lens, pad_id=1)
    q.put((rank, t.tolist(), l.tolist()))
    dist.destroy_process_group()


def test_gather_captions_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    exp_t = [[10, 10, 10, 10, 1]] * 3 + [[11] * 5] * 2
    for _, t, l in res:
        assert t == exp_t and l == [4, 4, 4, 5, 5]


def test_host_chunk_schedule_covers_batch():
    """The host-feature pipeline's chunk schedule (video_captioning_model._host_spans): contiguous cover of
    [0, B) with no chunk above the staging size."""
    import video_captioning_b200 as vc
    spans_of = vc.VideoCaptioningModel._host_spans
    for B in (1, 7, 31, 128, 129, 255, 256, 257, 1000, 1024, 4096):
        for chunk in (1, 8, 64, 256):
            sp = spans_of(B, chunk)
            assert sp[0][0] == 0 and sp[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(sp, sp[1:]))
            assert all(0 < hi - lo <= chunk for lo, hi in sp)


def test_host_pack_bf16_matches_round_to_nearest_even():
    """vc_host_pack_bf16 (host cores, AVX-512 or scalar) == torch's fp32 -> bf16 rounding, incl. NaN / Inf / denormals,
    odd lengths, unaligned starts and every thread count."""
    import torch
    from video_captioning_b200 import _native
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3 * 80 * 4096 + 37, generator=g) * 3
    x[:6] = torch.tensor([float("nan"), float("inf"), -float("inf"), 1e-40, -0.0, 65504.0])
    x[100:200] = torch.arange(100, dtype=torch.float32) * 0.0078125 + 1.0        # exact ties at the bf16 rounding point
    ref = x.to(torch.bfloat16)
    for threads in (1, 3, 16):
        for off in (0, 1, 5):
            src = x[off:].contiguous()
            dst = torch.zeros(src.numel(), dtype=torch.bfloat16)
            _native.host_pack_bf16(src, dst, threads)
            a, b = dst.view(torch.int16), ref[off:].view(torch.int16)
            nan = torch.isnan(src)
            assert torch.equal(a[~nan], b[~nan]) and torch.isnan(dst[nan]).all()


def test_packed_ingest_chunk_schedule():
    """Decode chunks of the packed host ingest: contiguous cover of the window, piece-aligned cuts, last chunk smallest of
    the defaults (its decode is the part of a step no transfer overlaps)."""
    import video_captioning_b200 as vc
    from oracle import synth
    m = vc.VideoCaptioningModel(synth.make_config("tiny"), 1000, precision="bf16")
    for (w0, w1, piece) in ((0, 1024, 64), (0, 23, 2), (14, 23, 2), (0, 5, 8), (100, 1124, 64), (0, 1000, 64)):
        ch = m._packed_chunks(w0, w1, piece)
        assert ch[0][0] == w0 and ch[-1][1] == w1
        for (a, b), (c, d) in zip(ch, ch[1:]):
            assert b == c and a < b
        for a, b in ch[:-1]:
            assert (b - w0) % piece == 0
    ch = m._packed_chunks(0, 1024, 64)
    assert [b - a for a, b in ch] == [512, 320, 192]
    m.host_chunk_fractions = (1.0,)
    assert m._packed_chunks(0, 1024, 64) == [(0, 1024)]


def test_host_pack_auto_depends_on_ranks_per_node(monkeypatch):
    """Host-side packing shares the node's cores and memory bandwidth: on for at most two ranks per node (measured),
    always overridable."""
    import video_captioning_b200 as vc
    from oracle import synth
    cfg = synth.make_config("tiny")
    monkeypatch.delenv("VC_HOST_PACK", raising=False)
    monkeypatch.setattr("os.cpu_count", lambda: 32)
    for lw, want in (("1", True), ("2", True), ("4", False), ("8", False)):
        monkeypatch.setenv("LOCAL_WORLD_SIZE", lw)
        assert vc.VideoCaptioningModel(cfg, 1000, precision="bf16").host_pack is want
    monkeypatch.setattr("os.cpu_count", lambda: 4)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    assert vc.VideoCaptioningModel(cfg, 1000, precision="bf16").host_pack is False      # too few cores to be worth it
    monkeypatch.setenv("VC_HOST_PACK", "1")
    assert vc.VideoCaptioningModel(cfg, 1000, precision="bf16").host_pack is True
    monkeypatch.setenv("VC_HOST_PACK", "0")
    assert vc.VideoCaptioningModel(cfg, 1000, precision="bf16").host_pack is False


class _FakeModel:
    """Stands in for VideoCaptioningModel in the sharding test (no GPU here): row i's tokens encode its video id."""

    def generate(self, video_features, start_token_id, end_token_id, max_length=20, video_mask=None, method="greedy", **kw):
        n = video_features.shape[0]
        if n == 0:
            return {"generated_tokens": torch.zeros(0, 1, dtype=torch.int64), "lengths": torch.zeros(0, dtype=torch.int64),
                    "scores": torch.zeros(0)}
        ids = video_features[:, 0, 0].to(torch.int64)
        toks = torch.stack([torch.full((n,), start_token_id), ids + 10, torch.full((n,), end_token_id)], dim=1)
        return {"generated_tokens": toks, "lengths": torch.full((n,), 3, dtype=torch.int64), "scores": torch.zeros(n)}


def _empty_shard_worker(rank, world, port, q):
    from video_captioning_b200.sharding import ShardedCaptioner
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = torch.arange(1, dtype=torch.float32).view(1, 1, 1).expand(1, 2, 4).clone()      # ONE video, two ranks
    out = ShardedCaptioner(_FakeModel()).generate(x, 1, 2, max_length=5, method="beam", already_sharded=False, beam_size=3)
    q.put((rank, out["generated_tokens"].tolist(), out["lengths"].tolist()))
    dist.destroy_process_group()


def test_sharded_captioner_empty_shard_gloo_world2():
    """Fewer videos than ranks: the rank with the empty shard must still take part in the gather."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_empty_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    for _, t, l in res:
        assert t == [[1, 10, 2]] and l == [3]


def test_model_dimension_limits_raise_at_construction():
    """Arbitrary vocabulary sizes are accepted (len(vocabulary) of a reference checkpoint is arbitrary); unsupported
    F/H/E/A raise a ValueError when the model is built, not at the first generate()."""
    cfg = synth.make_config("tiny")
    vc.VideoCaptioningModel(cfg, 1003)                      # odd vocabulary: fine
    vc.VideoCaptioningModel(cfg, 1003, precision="bf16")
    bad = synth.make_config("tiny", E=100)
    with pytest.raises(ValueError):
        vc.VideoCaptioningModel(bad, 1000)
    bad = synth.make_config("tiny", H=136)                  # multiple of 8, not of 64
    vc.VideoCaptioningModel(bad, 1000)
    with pytest.raises(ValueError):
        vc.VideoCaptioningModel(bad, 1000, precision="bf16")
    with pytest.raises(ValueError):
        vc.VideoCaptioningModel(bad, 1000).set_precision("bf16")
    with pytest.raises(ValueError):
        vc.VideoCaptioningModel(cfg, 3)


def test_library_staleness_is_detected_by_source_hash(tmp_path, monkeypatch):
    """A libvc_b200.so built from other sources than the tree's must not be loaded silently."""
    from video_captioning_b200 import _native
    assert _native.library_is_current()                     # conftest / build() left a current library
    monkeypatch.setattr(_native, "_HASH_PATH", str(tmp_path / "h"))
    assert not _native.library_is_current()                 # no hash file: treated as stale
    (tmp_path / "h").write_text("0" * 64)
    assert not _native.library_is_current()
    (tmp_path / "h").write_text(_native.sources_hash())
    assert _native.library_is_current()


def test_host_stage_rows_equals_reference_resize():
    """vc_host_stage_rows (the Predictor's batched resize/pad + optional bf16 rounding, one native pass) against the
    oracle's restatement of predictor.py:292-315 per video: longer videos subsampled at floor(linspace), shorter ones
    zero-padded, equal ones copied; fp32 -> fp32 exact, fp32 -> bf16 == torch's round-to-nearest-even, fp16 sources."""
    from video_captioning_b200 import _native
    from video_captioning_b200.predictor import resize_indices
    rng = np.random.default_rng(0)
    T, F = 16, 72
    vids = [rng.standard_normal((n, F)).astype(np.float32) for n in (16, 40, 7, 1, 33, 16, 200)]
    exp = np.stack([oracle_resize(v, T) for v in vids])
    for src_np, src_t in ((np.float32, torch.float32), (np.float16, torch.float16)):
        arrs = [np.ascontiguousarray(v.astype(src_np)) for v in vids]
        rows = np.empty((len(arrs), T), dtype=np.uint64)
        for b, a in enumerate(arrs):
            idx = resize_indices(a.shape[0], T)
            rows[b] = np.where(idx >= 0, a.ctypes.data + idx * (F * a.itemsize), 0).astype(np.uint64)
        want = torch.from_numpy(np.stack([oracle_resize(a, T) for a in arrs]))
        for dst_t in ((torch.float32, torch.bfloat16) if src_t == torch.float32 else (torch.float16, torch.bfloat16, torch.float32)):
            for threads in (1, 5):
                dst = torch.full((len(arrs), T, F), 7.0).to(dst_t)
                _native.host_stage_rows(rows.reshape(-1), len(arrs) * T, F, src_t, dst, threads)
                assert torch.equal(dst, want.to(torch.float32).to(dst_t) if dst_t != src_t else want), (src_t, dst_t, threads)
    assert np.array_equal(exp[2][7:], np.zeros((9, F), np.float32))


def test_result_packaging_matches_predict_py(tmp_path):
    """save_batch_results / save_single_result / save_multiple_captions write what src/predict.py:55-71, :105-137, :174-189
    write: the same keys, indent 2, and a captions file with an empty line for each failed video."""
    import json
    results = [{"video_path": "a.mp4", "caption": "a man is running", "tokens": [1, 5, 6, 2], "method": "beam"},
               {"video_path": "b.mp4", "caption": "", "error": "Feature file not found: b.npy"},
               {"video_path": "c.mp4", "caption": "a dog", "tokens": [7, 2], "method": "greedy",
                "attention_weights": torch.ones(2, 4)}]
    vc.save_batch_results(results, tmp_path / "o.json", tmp_path / "c.txt", method="beam", max_length=20, beam_size=5,
                          length_penalty=1.0, temperature=1.0)
    d = json.load(open(tmp_path / "o.json"))
    assert list(d) == ["parameters", "results"]
    assert d["parameters"] == {"method": "beam", "max_length": 20, "beam_size": 5, "length_penalty": 1.0, "temperature": 1.0}
    assert d["results"][1]["error"].startswith("Feature file") and d["results"][2]["attention_weights"] == [[1.0] * 4] * 2
    assert open(tmp_path / "c.txt").read() == "a man is running\n\na dog\n"
    assert open(tmp_path / "o.json").read().startswith('{\n  "parameters": {\n    "method"')
    vc.save_single_result(results[0], "a.mp4", tmp_path / "s.json", method="beam")
    d = json.load(open(tmp_path / "s.json"))
    assert list(d) == ["video_path", "caption", "method", "tokens", "parameters"] and d["tokens"] == [1, 5, 6, 2]
    caps = [{"caption": "x", "score": 1.0 / np.float64(0.7), "tokens": [3], "temperature": np.float64(0.7)}]
    vc.save_multiple_captions(caps, "a.mp4", tmp_path / "m.json", num_captions=1, method="greedy")
    d = json.load(open(tmp_path / "m.json"))
    assert list(d) == ["video_path", "captions", "parameters"] and abs(d["captions"][0]["temperature"] - 0.7) < 1e-12
