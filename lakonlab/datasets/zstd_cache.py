"""The reference's on-disk prompt-embedding cache: one zstd-compressed pickle per sample
(lakonlab/datasets/image_prompts.py:357-391 reads `<cache_dir>/<name>.zst` with
`zstd.ZstdDecompressor().stream_reader` + `pickle.load`; the records are written by the text-encoder caching tools with
`zstd.ZstdCompressor().compress(pickle.dumps(record))`).

Record (a plain dict):  prompt: str;  prompt_embed_kwargs: {encoder_hidden_states [S_t, C] (+ encoder_hidden_states_scale),
pooled_projections [P], encoder_hidden_states_mask [S_t]}  — or the legacy flat keys prompt_embeds / prompt_embeds_scale /
pooled_prompt_embeds / prompt_embeds_mask (image_prompts.py:86-91);  optional latents (+ latents_scale) or latent_size.

The `zstandard` Python package is not in this image, but the zstd FORMAT is what matters: frames are coded here through
`libzstd.so.1` (ctypes; the same C library `zstandard` wraps), with `pyarrow.Codec('zstd')` (pyarrow's bundled copy) as
the second implementation the tests cross-check against. Frames written by `zstandard` carry their content size (one-shot
`compress`) or not (`stream_writer`); both decode here (the streaming API handles unknown sizes and multi-frame files).
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import gzip
import io
import json
import os
import pickle
from typing import Dict, List, Optional, Sequence, Tuple

_ZSTD = None
_CONTENTSIZE_UNKNOWN = (1 << 64) - 1
_CONTENTSIZE_ERROR = (1 << 64) - 2


class _Buf(C.Structure):      # ZSTD_inBuffer / ZSTD_outBuffer share this layout
    _fields_ = [("ptr", C.c_void_p), ("size", C.c_size_t), ("pos", C.c_size_t)]


def _libzstd():
    global _ZSTD
    if _ZSTD is not None:
        return _ZSTD or None
    for name in (ctypes.util.find_library("zstd"), "libzstd.so.1", "libzstd.so"):
        if not name:
            continue
        try:
            lib = C.CDLL(name)
        except OSError:
            continue
        lib.ZSTD_compressBound.restype = C.c_size_t
        lib.ZSTD_compressBound.argtypes = [C.c_size_t]
        lib.ZSTD_compress.restype = C.c_size_t
        lib.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        lib.ZSTD_isError.restype = C.c_uint
        lib.ZSTD_isError.argtypes = [C.c_size_t]
        lib.ZSTD_getErrorName.restype = C.c_char_p
        lib.ZSTD_getErrorName.argtypes = [C.c_size_t]
        lib.ZSTD_getFrameContentSize.restype = C.c_ulonglong
        lib.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
        lib.ZSTD_createDStream.restype = C.c_void_p
        lib.ZSTD_freeDStream.argtypes = [C.c_void_p]
        lib.ZSTD_initDStream.restype = C.c_size_t
        lib.ZSTD_initDStream.argtypes = [C.c_void_p]
        lib.ZSTD_decompressStream.restype = C.c_size_t
        lib.ZSTD_decompressStream.argtypes = [C.c_void_p, C.POINTER(_Buf), C.POINTER(_Buf)]
        lib.ZSTD_DStreamOutSize.restype = C.c_size_t
        _ZSTD = lib
        return lib
    _ZSTD = False
    return None


def _check(lib, code: int, what: str) -> int:
    if lib.ZSTD_isError(code):
        raise ValueError(f"zstd {what} failed: {lib.ZSTD_getErrorName(code).decode()}")
    return code


def zstd_compress(data: bytes, level: int = 3) -> bytes:
    """One zstd frame with the content size in its header (what `zstandard.ZstdCompressor().compress` emits)."""
    lib = _libzstd()
    if lib is None:
        import pyarrow as pa
        return pa.Codec("zstd", compression_level=level).compress(data, asbytes=True)
    bound = lib.ZSTD_compressBound(len(data))
    dst = C.create_string_buffer(bound)
    n = _check(lib, lib.ZSTD_compress(dst, bound, data, len(data), level), "compress")
    return dst.raw[:n]


def zstd_decompress(data: bytes) -> bytes:
    """Decodes a whole `.zst` payload: any number of concatenated frames, with or without a content-size header."""
    lib = _libzstd()
    if lib is None:
        return _pyarrow_decompress(data)
    ds = lib.ZSTD_createDStream()
    if not ds:
        raise MemoryError("ZSTD_createDStream")
    try:
        _check(lib, lib.ZSTD_initDStream(ds), "initDStream")
        chunk = int(lib.ZSTD_DStreamOutSize())
        hint = lib.ZSTD_getFrameContentSize(data, len(data))
        if hint not in (_CONTENTSIZE_UNKNOWN, _CONTENTSIZE_ERROR) and 0 < hint < (1 << 34):
            chunk = max(chunk, int(hint))
        src = C.create_string_buffer(data, len(data))
        inb = _Buf(C.cast(src, C.c_void_p), len(data), 0)
        out = io.BytesIO()
        dst = C.create_string_buffer(chunk)
        ret = 0
        while inb.pos < inb.size:
            outb = _Buf(C.cast(dst, C.c_void_p), chunk, 0)
            ret = _check(lib, lib.ZSTD_decompressStream(ds, C.byref(outb), C.byref(inb)), "decompress")
            out.write(dst.raw[:outb.pos])
            while outb.pos == chunk:      # the output buffer was filled: drain what the decoder still holds
                outb = _Buf(C.cast(dst, C.c_void_p), chunk, 0)
                ret = _check(lib, lib.ZSTD_decompressStream(ds, C.byref(outb), C.byref(inb)), "decompress")
                out.write(dst.raw[:outb.pos])
        if ret != 0:
            raise ValueError("zstd decompress failed: truncated frame")
        return out.getvalue()
    finally:
        lib.ZSTD_freeDStream(ds)


def _pyarrow_decompress(data: bytes) -> bytes:
    import pyarrow as pa
    return pa.CompressedInputStream(pa.BufferReader(data), "zstd").read()


def dumps_record(record: Dict, level: int = 3) -> bytes:
    return zstd_compress(pickle.dumps(record, protocol=pickle.HIGHEST_PROTOCOL), level)


def loads_record(payload: bytes) -> Dict:
    """`pickle.load` of the decompressed stream, as the reference does (image_prompts.py:361-362) — like there, cache
    shards are trusted input (they are produced by the user's own caching run)."""
    return pickle.loads(zstd_decompress(payload))


def write_record(cache_dir: str, name: str, prompt: str, prompt_embed_kwargs: Dict, latents=None, latents_scale=None,
                 latent_size: Optional[Sequence[int]] = None, legacy_keys: bool = False, level: int = 3) -> str:
    """Writes `<cache_dir>/<name>.zst` in the layout image_prompts.py:357-391 reads. legacy_keys=True stores the flat
    pre-`prompt_embed_kwargs` names (prompt_embeds, prompt_embeds_scale, pooled_prompt_embeds, prompt_embeds_mask)."""
    from . import PROMPT_KEY_MAPS
    rec: Dict = dict(prompt=prompt)
    if legacy_keys:
        new_to_old = {v: k for k, v in PROMPT_KEY_MAPS.items()}
        for k, v in prompt_embed_kwargs.items():
            rec[new_to_old.get(k, k)] = v
    else:
        rec["prompt_embed_kwargs"] = dict(prompt_embed_kwargs)
    if latents is not None:
        rec["latents"] = latents
        if latents_scale is not None:
            rec["latents_scale"] = latents_scale
    elif latent_size is not None:
        rec["latent_size"] = tuple(latent_size)
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, name + ".zst")
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(dumps_record(rec, level))
    os.replace(tmp, path)
    return path


def read_record(cache_dir: str, name: str) -> Dict:
    with open(os.path.join(cache_dir, name + ".zst"), "rb") as f:
        return loads_record(f.read())


def parse_datalist(dir_path: str, datalist_path: Optional[str] = None, bucketize: bool = False,
                   datalist_must_exist: bool = False) -> Tuple[List[str], Optional[List[int]]]:
    """Sample names (+ bucket ids) of a cache directory: a `.jsonl` / `.jsonl.gz` datalist with `filename` (or
    `image_hash`) and, when bucketizing, `size_idx` / `bucket_id` per line; a `.json` list of paths; or, without a
    datalist, the sorted directory listing — which is then saved to `datalist_path` (image_prompts.py:201-270)."""
    if datalist_path is not None and os.path.isfile(datalist_path):
        names, buckets = [], []
        if datalist_path.endswith((".jsonl", ".jsonl.gz")):
            opener = gzip.open if datalist_path.endswith(".gz") else open
            with opener(datalist_path, "rt", encoding="utf-8") as f:
                lines = [ln for ln in f.read().splitlines() if ln.strip()]
            for ln in lines:
                item = json.loads(ln)
                if "filename" in item:
                    names.append(item["filename"])
                elif "image_hash" in item:
                    names.append(item["image_hash"])
                else:
                    raise ValueError("No valid key to identify data item.")
                if bucketize:
                    if "size_idx" in item:
                        buckets.append(item["size_idx"])
                    elif "bucket_id" in item:
                        buckets.append(item["bucket_id"])
                    else:
                        raise ValueError("Either `size_idx` or `bucket_id` must be present in datalist for bucketize.")
        elif datalist_path.endswith(".json"):
            if bucketize:
                raise ValueError("Bucketize not supported for json datalist.")
            with open(datalist_path, "rb") as f:
                names = [os.path.splitext(os.path.basename(p))[0] for p in json.load(f)]
        else:
            raise ValueError("Datalist file must be .jsonl, .jsonl.gz or .json")
        return names, (buckets if bucketize else None)
    if datalist_must_exist:
        raise FileNotFoundError(f"Datalist file {datalist_path} does not exist.")
    if bucketize:
        raise ValueError("Bucketize not supported when datalist is not provided.")
    names = sorted(os.path.splitext(p)[0] for p in os.listdir(dir_path) if not p.endswith(".tmp"))
    if datalist_path is not None:
        if datalist_path.endswith((".jsonl", ".jsonl.gz")):
            text = "\n".join(json.dumps({"filename": n}) for n in names)
            opener = gzip.open if datalist_path.endswith(".gz") else open
            with opener(datalist_path, "wt", encoding="utf-8") as f:
                f.write(text)
        elif datalist_path.endswith(".json"):
            with open(datalist_path, "w", encoding="utf-8") as f:
                json.dump(names, f)
    return names, None
