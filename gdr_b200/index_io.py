"""On-disk form of the cluster-contiguous CSR index.

The reference keeps its index as three pickles — `doc_embedding.pkl` (list of [768] fp32 tensors,
GDR_model/main_models.py:182-187, 806-814), `indexmap*.pkl` (`id_mapping`, :874-889) and the pickled tree (:225-226,
727-728) — which is infeasible at the 100M-document scale (153 GB of Python objects).  This module defines one
mmap-able file with the arrays in exactly the layout `ClusterStore` keeps in HBM, a converter from the reference's
pickles, and a chunked uploader (pinned staging buffer, so a shard larger than host RAM streams to the device).

File layout (little endian), every section 4096-byte aligned:
    header   64 bytes: magic "GDRCSR01", u32 dtype (0 = fp32, 1 = bf16), u32 dim, u64 n_rows, u64 n_clusters,
             u64 off_offsets, u64 off_docid, u64 off_emb, u64 off_keys (0 = no keys)
    offsets  int64 [n_clusters + 1]
    docid    int64 [n_rows]
    emb      dtype [n_rows, dim], row-major, cluster c = rows offsets[c] .. offsets[c+1]-1
    keys     u64 length + UTF-8 JSON list of the cluster-id strings ("3-17-22", ...), in cluster order
"""
from __future__ import annotations

import json
import os
import pickle
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

MAGIC = b"GDRCSR01"
_HDR = struct.Struct("<8sIIQQQQQQ")      # 64 bytes
_ALIGN = 4096


def _align(x: int) -> int:
    return (x + _ALIGN - 1) // _ALIGN * _ALIGN


def write_index(path: str, emb: torch.Tensor, offsets, docid, keys: Optional[Sequence[str]] = None) -> None:
    """emb [N, D] fp32 or bf16 (CPU), offsets [C+1], docid [N], keys optional list of C cluster-id strings."""
    emb = emb.detach().to("cpu").contiguous()
    if emb.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("emb must be float32 or bfloat16")
    offsets = np.ascontiguousarray(np.asarray(offsets), dtype=np.int64)
    docid = np.ascontiguousarray(np.asarray(docid), dtype=np.int64)
    n_rows, dim = emb.shape
    C = offsets.size - 1
    if offsets[0] != 0 or offsets[-1] != n_rows or np.any(np.diff(offsets) < 0) or docid.size != n_rows:
        raise ValueError("offsets / docid do not describe emb")
    if keys is not None and len(keys) != C:
        raise ValueError("one key per cluster expected")
    off_offsets = _align(_HDR.size)
    off_docid = _align(off_offsets + offsets.nbytes)
    off_emb = _align(off_docid + docid.nbytes)
    emb_bytes = n_rows * dim * emb.element_size()
    off_keys = _align(off_emb + emb_bytes) if keys is not None else 0
    with open(path, "wb") as f:
        f.write(_HDR.pack(MAGIC, 1 if emb.dtype == torch.bfloat16 else 0, dim, n_rows, C, off_offsets, off_docid, off_emb, off_keys))
        f.seek(off_offsets); f.write(offsets.tobytes())
        f.seek(off_docid); f.write(docid.tobytes())
        f.seek(off_emb)
        raw = emb.view(torch.int16) if emb.dtype == torch.bfloat16 else emb
        step = max(1, (64 << 20) // max(1, dim * emb.element_size()))
        for i in range(0, n_rows, step):
            f.write(raw[i:i + step].numpy().tobytes())
        if keys is not None:
            blob = json.dumps(list(keys)).encode("utf-8")
            f.seek(off_keys); f.write(struct.pack("<Q", len(blob))); f.write(blob)


def read_header(path: str) -> Dict[str, int]:
    with open(path, "rb") as f:
        magic, dtype, dim, n_rows, C, o_off, o_doc, o_emb, o_keys = _HDR.unpack(f.read(_HDR.size))
    if magic != MAGIC:
        raise ValueError(f"{path}: not a GDRCSR01 index file")
    return dict(dtype=dtype, dim=dim, n_rows=n_rows, n_clusters=C, off_offsets=o_off, off_docid=o_doc, off_emb=o_emb, off_keys=o_keys)


def map_index(path: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray, Optional[List[str]], torch.dtype]:
    """Memory-map the file: (emb as uint16/float32 memmap [N, D], offsets, docid, keys, torch dtype).  Nothing is read
    until touched, so a 150 GB index opens instantly."""
    h = read_header(path)
    dt = torch.bfloat16 if h["dtype"] == 1 else torch.float32
    np_dt = np.uint16 if h["dtype"] == 1 else np.float32
    offsets = np.memmap(path, dtype=np.int64, mode="r", offset=h["off_offsets"], shape=(h["n_clusters"] + 1,))
    docid = np.memmap(path, dtype=np.int64, mode="r", offset=h["off_docid"], shape=(h["n_rows"],))
    emb = np.memmap(path, dtype=np_dt, mode="r", offset=h["off_emb"], shape=(h["n_rows"], h["dim"]))
    keys = None
    if h["off_keys"]:
        with open(path, "rb") as f:
            f.seek(h["off_keys"])
            (n,) = struct.unpack("<Q", f.read(8))
            keys = json.loads(f.read(n).decode("utf-8"))
    return emb, offsets, docid, keys, dt


def load_store(path: str, device="cuda", clusters: Optional[np.ndarray] = None, chunk_rows: int = 1 << 18):
    """Open the file and build a ClusterStore in HBM.  `clusters` (ascending global cluster ids) selects this rank's
    shard of a cluster-sharded corpus; the rows are streamed through a pinned staging buffer `chunk_rows` at a time."""
    from .store import ClusterStore
    emb, offsets, docid, keys, dt = map_index(path)
    sizes = np.diff(offsets)
    if clusters is None:
        ranges = [(0, emb.shape[0])]
        loc_off, loc_doc, loc_keys = np.asarray(offsets), np.asarray(docid), keys
    else:
        clusters = np.asarray(clusters, dtype=np.int64)
        ranges = [(int(offsets[c]), int(offsets[c + 1])) for c in clusters]
        loc_off = np.zeros(clusters.size + 1, dtype=np.int64)
        loc_off[1:] = np.cumsum(sizes[clusters])
        loc_doc = np.concatenate([docid[a:b] for a, b in ranges]) if ranges else np.zeros(0, np.int64)
        loc_keys = [keys[c] for c in clusters] if keys is not None else None
    n_loc = int(loc_off[-1])
    dev_emb = torch.empty((n_loc, emb.shape[1]), dtype=dt, device=device)
    stage = torch.empty((chunk_rows, emb.shape[1]), dtype=torch.int16 if dt == torch.bfloat16 else torch.float32).pin_memory()
    stage_np = stage.numpy()
    pos = 0
    for a, b in ranges:
        for i in range(a, b, chunk_rows):
            n = min(chunk_rows, b - i)
            stage_np[:n] = emb[i:i + n].view(stage_np.dtype) if dt == torch.bfloat16 else emb[i:i + n]
            src = stage[:n].view(torch.bfloat16) if dt == torch.bfloat16 else stage[:n]
            dev_emb[pos:pos + n].copy_(src, non_blocking=False)
            pos += n
    return ClusterStore(dev_emb, torch.from_numpy(np.array(loc_off, dtype=np.int64)), torch.from_numpy(np.array(loc_doc, dtype=np.int64)), loc_keys)


def convert_pickles(doc_embedding_pkl: str, indexmap_pkl: str, out_path: str, dtype=torch.bfloat16) -> Dict[str, int]:
    """Reference pickles -> index file.  `doc_embedding.pkl` unpickles to an int-indexable container of [D] tensors
    (main_models.py:182-187), `indexmap*.pkl` to `Dict[str, List[int]]` (:874-889)."""
    from .store import csr_from_reference
    with open(doc_embedding_pkl, "rb") as f:
        doc_embed = pickle.load(f)
    with open(indexmap_pkl, "rb") as f:
        id_mapping = pickle.load(f)
    emb, offsets, docid, keys = csr_from_reference(doc_embed, id_mapping)
    write_index(out_path, emb.to(dtype), offsets.numpy(), docid.numpy(), keys)
    return read_header(out_path)


def main(argv=None) -> int:
    """`python -m gdr_b200.index_io convert doc_embedding.pkl indexmap.pkl corpus.gdr [--dtype bf16|f32]` turns the reference's
    two pickles into the mmap-able index file; `python -m gdr_b200.index_io info corpus.gdr` prints its header.  Host only."""
    import argparse
    import json
    ap = argparse.ArgumentParser(prog="python -m gdr_b200.index_io", description=main.__doc__)
    sub = ap.add_subparsers(dest="cmd", required=True)
    c = sub.add_parser("convert", help="reference pickles -> index file")
    c.add_argument("doc_embedding_pkl")
    c.add_argument("indexmap_pkl")
    c.add_argument("out")
    c.add_argument("--dtype", choices=["bf16", "f32"], default="bf16")
    i = sub.add_parser("info", help="print the header of an index file")
    i.add_argument("path")
    args = ap.parse_args(argv)
    if args.cmd == "convert":
        hdr = convert_pickles(args.doc_embedding_pkl, args.indexmap_pkl, args.out, torch.bfloat16 if args.dtype == "bf16" else torch.float32)
    else:
        hdr = read_header(args.path)
    print(json.dumps({k: int(v) for k, v in hdr.items()}))
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(main())
