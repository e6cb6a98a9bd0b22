"""Writes tests/golden/zstd_cache/: a small prompt-embedding cache in the reference's on-disk format (one zstd-compressed
pickle per sample + a .jsonl.gz datalist), following the reader at lakonlab/datasets/image_prompts.py:357-391 and the key
conventions at :86-91, :281-309. Records:
  p000  new-style `prompt_embed_kwargs` with a bf16 payload + `encoder_hidden_states_scale`, a mask, `latent_size`
  p001  legacy flat keys (prompt_embeds / prompt_embeds_scale / pooled_prompt_embeds / prompt_embeds_mask)
  p002  cached latents + `latents_scale`; written as a STREAMED frame (no content size in the frame header, what
        `zstandard`'s stream_writer emits) through pyarrow's bundled zstd — a second implementation of the format
Values are seeded, so the test recomputes what the reader must return.
"""
import gzip
import json
import os
import pickle
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lakonlab.datasets.zstd_cache import write_record  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "zstd_cache")


def tensors(seed, seq, dim=32, pooled=8):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(seq, dim, generator=g), torch.randn(pooled, generator=g),
            (torch.arange(seq) < seq - 1).to(torch.int64))


def main():
    os.makedirs(OUT, exist_ok=True)
    e, p, m = tensors(0, 5)
    write_record(OUT, "p000", "a photo of a cat", dict(encoder_hidden_states=(e / 0.5).bfloat16(), encoder_hidden_states_scale=0.5,
                                                       pooled_projections=p.half(), encoder_hidden_states_mask=m),
                 latent_size=(16, 8, 12))
    e, p, m = tensors(1, 9)
    write_record(OUT, "p001", "a dog", dict(encoder_hidden_states=(e / 0.25).bfloat16(), encoder_hidden_states_scale=0.25,
                                            pooled_projections=p, encoder_hidden_states_mask=m), legacy_keys=True)
    e, p, m = tensors(2, 3)
    g = torch.Generator().manual_seed(7)
    rec = dict(prompt="a bird", prompt_embed_kwargs=dict(encoder_hidden_states=e, pooled_projections=p),
               latents=torch.randn(16, 4, 4, generator=g).half(), latents_scale=2.0)
    import pyarrow as pa
    sink = pa.BufferOutputStream()
    with pa.CompressedOutputStream(sink, "zstd") as f:      # streaming writer: the frame header carries no content size
        f.write(pickle.dumps(rec, protocol=pickle.HIGHEST_PROTOCOL))
    with open(os.path.join(OUT, "p002.zst"), "wb") as f:
        f.write(sink.getvalue().to_pybytes())
    with gzip.open(os.path.join(OUT, "datalist.jsonl.gz"), "wt", encoding="utf-8") as f:
        f.write("\n".join(json.dumps({"filename": n, "size_idx": i % 2}) for i, n in enumerate(["p000", "p001", "p002"])))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
