"""The reference's zstd-pickled prompt-embedding cache (lakonlab/datasets/image_prompts.py:86-91,
:281-309, :357-391): committed fixtures (tools/make_golden_zstd.py), two independent zstd implementations (libzstd via
ctypes, pyarrow's bundled copy), legacy keys, `encoder_hidden_states_scale`, `pad_seq_len`, latents, test-mode noise."""
import gzip
import json
import os
import pickle

import pytest
import torch

from lakonlab.datasets import ImagePrompts, build_dataloader, build_dataset, parse_prompt_embeds
from lakonlab.datasets import zstd_cache as Z

GOLD = os.path.join(os.path.dirname(__file__), "golden", "zstd_cache")


def _tensors(seed, seq, dim=32, pooled=8):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(seq, dim, generator=g), torch.randn(pooled, generator=g), (torch.arange(seq) < seq - 1).to(torch.int64))


def test_libzstd_is_the_codec_and_agrees_with_pyarrow():
    assert Z._libzstd() is not None, "libzstd.so.1 is part of the image"
    import pyarrow as pa
    blob = os.urandom(1000) + b"arcflow" * 5000
    ours = Z.zstd_compress(blob)
    assert ours[:4] == b"\x28\xb5\x2f\xfd"                       # zstd frame magic
    assert pa.Codec("zstd").decompress(ours, decompressed_size=len(blob)).to_pybytes() == blob
    theirs = pa.Codec("zstd").compress(blob, asbytes=True)
    assert Z.zstd_decompress(theirs) == blob and Z._pyarrow_decompress(ours) == blob
    assert Z.zstd_decompress(ours + theirs) == blob + blob      # concatenated frames
    assert Z.zstd_decompress(Z.zstd_compress(b"")) == b""
    with pytest.raises(ValueError):
        Z.zstd_decompress(ours[:-5])                             # truncated frame
    with pytest.raises(ValueError):
        Z.zstd_decompress(b"not a zstd frame at all")


def test_fixture_records_decode_with_both_implementations():
    for name in ("p000", "p001", "p002"):
        with open(os.path.join(GOLD, name + ".zst"), "rb") as f:
            payload = f.read()
        a = pickle.loads(Z.zstd_decompress(payload))
        b = pickle.loads(Z._pyarrow_decompress(payload))
        assert a.keys() == b.keys() and a["prompt"] == b["prompt"]
    # p002 is a streamed frame: no content size in its header (the path `zstandard`'s stream_writer produces)
    with open(os.path.join(GOLD, "p002.zst"), "rb") as f:
        payload = f.read()
    assert Z._libzstd().ZSTD_getFrameContentSize(payload, len(payload)) == Z._CONTENTSIZE_UNKNOWN
    assert "prompt_embed_kwargs" not in Z.read_record(GOLD, "p001")   # legacy flat keys on disk


def test_dataset_semantics_on_the_fixture():
    ds = build_dataset(dict(type="ImagePrompts", data_root=os.path.dirname(GOLD), cache_dir="zstd_cache",
                            cache_datalist_path=os.path.join(GOLD, "datalist.jsonl.gz"), pad_seq_len=6,
                            latent_size=(16, 128, 128), bucketize=True))
    assert len(ds) == 3 and ds.bucket_ids == [0, 1, 0]
    s0, s1, s2 = ds[0], ds[1], ds[2]
    e0, p0, m0 = _tensors(0, 5)
    pe = s0["prompt_embed_kwargs"]
    assert s0["name"] == "a photo of a cat" and "encoder_hidden_states_scale" not in pe
    want = (e0 / 0.5).bfloat16().float() * 0.5                   # stored payload x stored scale, fp32
    assert pe["encoder_hidden_states"].dtype == torch.float32 and pe["encoder_hidden_states"].shape == (6, 32)
    assert torch.equal(pe["encoder_hidden_states"][:5], want) and pe["encoder_hidden_states"][5].abs().sum() == 0
    assert torch.equal(pe["encoder_hidden_states_mask"], torch.tensor([1, 1, 1, 1, 0, 0]))
    assert pe["pooled_projections"].dtype == torch.float32 and torch.equal(pe["pooled_projections"], p0.half().float())
    assert s0["latents"].shape == (16, 8, 12)                    # the record's latent_size wins over the default
    e1, p1, m1 = _tensors(1, 9)                                  # legacy keys, truncated to pad_seq_len
    pe = s1["prompt_embed_kwargs"]
    assert torch.equal(pe["encoder_hidden_states"], ((e1 / 0.25).bfloat16().float() * 0.25)[:6])
    assert torch.equal(pe["encoder_hidden_states_mask"], m1[:6]) and torch.equal(pe["pooled_projections"], p1)
    assert s1["latents"].shape == (16, 128, 128)
    g = torch.Generator().manual_seed(7)                         # cached latents x latents_scale
    assert torch.equal(s2["latents"], torch.randn(16, 4, 4, generator=g).half().float() * 2.0)
    # test mode: per-index seeded noise of the latent shape instead of latents
    t = ImagePrompts(cache_dir=GOLD, test_mode=True)
    assert "latents" not in t[2] and torch.equal(t[2]["noise"], torch.randn(16, 4, 4, generator=torch.Generator().manual_seed(2)))
    assert ImagePrompts(cache_dir=GOLD, ignore_cached_latents=True, latent_size=(16, 2, 2))[2]["latents"].shape == (16, 2, 2)
    # start / end / repeat index mapping (image_prompts.py:170-181, :346-350)
    r = ImagePrompts(cache_dir=GOLD, start_ind=1, end_ind=3, repeat=2)
    assert len(r) == 4 and [r[i]["name"] for i in range(4)] == ["a dog", "a dog", "a bird", "a bird"]


def test_write_read_round_trip_and_datalists(tmp_path):
    e, p, m = _tensors(5, 4)
    d = str(tmp_path / "cache")
    for i, legacy in enumerate((False, True)):
        Z.write_record(d, f"s{i}", f"prompt {i}", dict(encoder_hidden_states=e.bfloat16(), pooled_projections=p,
                                                         encoder_hidden_states_mask=m), legacy_keys=legacy)
    a, b = Z.read_record(d, "s0"), Z.read_record(d, "s1")
    assert set(b) == {"prompt", "prompt_embeds", "pooled_prompt_embeds", "prompt_embeds_mask"}
    pa_, pb_ = parse_prompt_embeds(a), parse_prompt_embeds(b)
    assert all(torch.equal(pa_[k], pb_[k]) for k in pa_) and set(pa_) == set(pb_)
    # no datalist: the directory is listed and the listing is saved where the datalist was expected
    dl = str(tmp_path / "list.jsonl")
    names, buckets = Z.parse_datalist(d, dl)
    assert names == ["s0", "s1"] and buckets is None and os.path.isfile(dl)
    assert [json.loads(x)["filename"] for x in open(dl).read().splitlines()] == ["s0", "s1"]
    assert Z.parse_datalist(d, dl)[0] == ["s0", "s1"]
    js = str(tmp_path / "list.json")
    json.dump(["/x/y/s1.zst", "s0.zst"], open(js, "w"))
    assert Z.parse_datalist(d, js)[0] == ["s1", "s0"]
    with gzip.open(str(tmp_path / "h.jsonl.gz"), "wt") as f:
        f.write(json.dumps({"image_hash": "s1", "bucket_id": 3}))
    assert Z.parse_datalist(d, str(tmp_path / "h.jsonl.gz"), bucketize=True) == (["s1"], [3])
    with pytest.raises(ValueError):
        Z.parse_datalist(d, None, bucketize=True)
    batch = next(iter(build_dataloader(ImagePrompts(cache_dir=d, pad_seq_len=8, latent_size=(16, 4, 4)), samples_per_gpu=2,
                                       shuffle=False)))
    assert batch["prompt_embed_kwargs"]["encoder_hidden_states"].shape == (2, 8, 32)
    assert batch["prompt_embed_kwargs"]["encoder_hidden_states_mask"].shape == (2, 8)
