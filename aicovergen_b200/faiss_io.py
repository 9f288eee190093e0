"""Reader / writer for faiss' binary `IndexIVFFlat` files (fourcc "IwFl", flat L2/IP coarse quantiser "IxF2"/"IxFI"),
the format of the `added_IVF{n}_Flat_nprobe_1_*.index` files RVC voice models ship with (reference call site
`faiss.read_index(file_index)`, vc_infer_pipeline.py:505; faiss-cpu 1.7.3 pinned in requirements.txt:3 is absent here).

Layout restated from the published faiss serialisation (faiss/impl/index_write.cpp, 1.7.x), little-endian:
    u32 "IwFl"
    index header : i32 d | i64 ntotal | i64 dummy | i64 dummy | u8 is_trained | i32 metric_type [| f32 metric_arg if metric > 1]
    u64 nlist | u64 nprobe
    quantiser    : u32 "IxF2"/"IxFI"/"IxFl" | index header | u64 n_floats | f32[n_floats]      (nlist x d centroids)
    direct map   : u8 type | u64 n | i64[n]  (| u64 n_pairs | (i64,i64)[n_pairs] when type == 2, hashtable)
    (NO code_size field here: "IwFl" is fourcc + ivf header + inverted lists, the reader derives code_size = 4 d;
     only the quantising IVF types - IwSq, IwPQ ... - and the legacy "IvFl" layout carry one)
    inverted lists: u32 "ilar" | u64 nlist | u64 code_size | u32 "full" | u64 nlist | u64 sizes[nlist]
                                                          (or u32 "sprs" | u64 2k | (u64 list, u64 size)[k])
                    then per non-empty list: codes u8[size*code_size] | ids i64[size]
`parity unpinned`: no faiss build and no index file exist in this environment; tests/test_formats_cpu.py parses a
byte string assembled field by field from the layout above (independently of the writer below) besides the round trip.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import BinaryIO

import numpy as np

METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1


@dataclass
class IvfFlatData:
    d: int
    nlist: int
    nprobe: int
    metric: int
    centroids: np.ndarray      # [nlist, d] float32
    vectors: np.ndarray        # [ntotal, d] float32 in id order (what reconstruct_n(0, ntotal) returns)
    list_of: np.ndarray        # [ntotal] int64: inverted list holding each id
    ids_sequential: bool       # ids are exactly 0..ntotal-1 (true for indexes built with index.add)


class FaissFormatError(ValueError):
    pass


def _rd(f: BinaryIO, fmt: str):
    n = struct.calcsize(fmt)
    b = f.read(n)
    if len(b) != n:
        raise FaissFormatError("truncated faiss index file")
    v = struct.unpack("<" + fmt, b)
    return v[0] if len(v) == 1 else v


def _fourcc(f: BinaryIO) -> str:
    b = f.read(4)
    if len(b) != 4:
        raise FaissFormatError("truncated faiss index file")
    return b.decode("latin1")


def _read_header(f: BinaryIO):
    d = _rd(f, "i")
    ntotal = _rd(f, "q")
    _rd(f, "q"), _rd(f, "q")
    is_trained = _rd(f, "B")
    metric = _rd(f, "i")
    if metric > 1:
        _rd(f, "f")
    return d, ntotal, bool(is_trained), metric


def _read_array(f: BinaryIO, dtype, count: int) -> np.ndarray:
    nbytes = int(count) * np.dtype(dtype).itemsize
    b = f.read(nbytes)
    if len(b) != nbytes:
        raise FaissFormatError("truncated faiss index file")
    return np.frombuffer(b, dtype=dtype, count=int(count))


def read_ivfflat(path: str) -> IvfFlatData:
    with open(path, "rb") as f:
        cc = _fourcc(f)
        if cc != "IwFl":
            raise FaissFormatError(f"{path}: index type {cc!r} is not IndexIVFFlat ('IwFl'); RVC ships IVF-Flat indexes")
        d, ntotal, _, metric = _read_header(f)
        nlist, nprobe = _rd(f, "Q"), _rd(f, "Q")
        qcc = _fourcc(f)
        if qcc not in ("IxF2", "IxFI", "IxFl"):
            raise FaissFormatError(f"{path}: coarse quantiser {qcc!r} is not a flat index")
        qd, qn, _, _ = _read_header(f)
        nfl = _rd(f, "Q")
        if qd != d or qn != nlist or nfl != nlist * d:
            raise FaissFormatError(f"{path}: quantiser shape mismatch (d {qd} vs {d}, n {qn} vs nlist {nlist}, floats {nfl})")
        centroids = _read_array(f, "<f4", nfl).reshape(nlist, d).copy()
        dm_type = _rd(f, "B")
        dm_n = _rd(f, "Q")
        _read_array(f, "<i8", dm_n)
        if dm_type == 2:
            npairs = _rd(f, "Q")
            _read_array(f, "<i8", 2 * npairs)
        code_size = 4 * d                      # IndexIVFFlat: not stored, ivfl->code_size = d * sizeof(float)
        il = _fourcc(f)
        if il != "ilar":
            raise FaissFormatError(f"{path}: inverted lists {il!r} are not ArrayInvertedLists ('ilar')")
        il_nlist, il_cs = _rd(f, "Q"), _rd(f, "Q")
        if il_nlist != nlist or il_cs != code_size:
            raise FaissFormatError(f"{path}: inverted-list header mismatch (nlist {il_nlist} vs {nlist}, code_size {il_cs} vs 4*d = {code_size})")
        kind = _fourcc(f)
        sizes = np.zeros(nlist, dtype=np.int64)
        if kind == "full":
            n = _rd(f, "Q")
            if n != nlist:
                raise FaissFormatError(f"{path}: {n} list sizes for {nlist} lists")
            sizes[:] = _read_array(f, "<u8", n)
        elif kind == "sprs":
            n = _rd(f, "Q")
            pairs = _read_array(f, "<u8", n).reshape(-1, 2)
            sizes[pairs[:, 0].astype(np.int64)] = pairs[:, 1]
        else:
            raise FaissFormatError(f"{path}: unknown list-size encoding {kind!r}")
        if int(sizes.sum()) != ntotal:
            raise FaissFormatError(f"{path}: list sizes sum to {int(sizes.sum())}, ntotal is {ntotal}")
        codes = np.empty((ntotal, d), dtype=np.float32)
        ids = np.empty(ntotal, dtype=np.int64)
        list_of_pos = np.empty(ntotal, dtype=np.int64)
        pos = 0
        for li in range(nlist):
            n = int(sizes[li])
            if n == 0:
                continue
            codes[pos:pos + n] = _read_array(f, "<f4", n * d).reshape(n, d)
            ids[pos:pos + n] = _read_array(f, "<i8", n)
            list_of_pos[pos:pos + n] = li
            pos += n
    seq = bool(ntotal == 0 or (ids.min() == 0 and ids.max() == ntotal - 1 and np.unique(ids).size == ntotal))
    if seq:
        vectors = np.empty_like(codes)
        vectors[ids] = codes
        list_of = np.empty(ntotal, dtype=np.int64)
        list_of[ids] = list_of_pos
    else:   # custom ids (add_with_ids): keep storage order; reconstruct_n is then not id-addressable, same as faiss without a direct map
        vectors, list_of = codes, list_of_pos
    return IvfFlatData(d, nlist, int(nprobe), int(metric), centroids, vectors, list_of, seq)


def _write_header(f: BinaryIO, d: int, ntotal: int, metric: int):
    f.write(struct.pack("<iqqqBi", d, ntotal, 1 << 20, 1 << 20, 1, metric))


def write_ivfflat(path: str, centroids: np.ndarray, vectors: np.ndarray, list_of: np.ndarray, nprobe: int = 1,
                  metric: int = METRIC_L2) -> None:
    """Writes the same byte layout faiss.write_index produces for an IndexIVFFlat whose vectors were added in id order."""
    centroids = np.ascontiguousarray(centroids, dtype="<f4")
    vectors = np.ascontiguousarray(vectors, dtype="<f4")
    list_of = np.asarray(list_of, dtype=np.int64)
    nlist, d = centroids.shape
    ntotal = vectors.shape[0]
    assert vectors.shape[1] == d and list_of.shape == (ntotal,) and (ntotal == 0 or (0 <= list_of.min() and list_of.max() < nlist))
    order = np.argsort(list_of, kind="stable")
    sizes = np.bincount(list_of, minlength=nlist).astype("<u8")
    with open(path, "wb") as f:
        f.write(b"IwFl")
        _write_header(f, d, ntotal, metric)
        f.write(struct.pack("<QQ", nlist, nprobe))
        f.write(b"IxF2" if metric == METRIC_L2 else b"IxFI")
        _write_header(f, d, nlist, metric)
        f.write(struct.pack("<Q", nlist * d))
        f.write(centroids.tobytes())
        f.write(struct.pack("<BQ", 0, 0))                       # no direct map
        f.write(b"ilar")
        f.write(struct.pack("<QQ", nlist, 4 * d))
        if int((sizes > 0).sum()) > nlist // 2:
            f.write(b"full")
            f.write(struct.pack("<Q", nlist))
            f.write(sizes.tobytes())
        else:
            nz = np.nonzero(sizes)[0]
            f.write(b"sprs")
            f.write(struct.pack("<Q", 2 * nz.size))
            f.write(np.stack([nz.astype("<u8"), sizes[nz]], 1).astype("<u8").tobytes())
        pos = 0
        for li in range(nlist):
            n = int(sizes[li])
            if n == 0:
                continue
            sel = order[pos:pos + n]
            f.write(vectors[sel].tobytes())
            f.write(sel.astype("<i8").tobytes())
            pos += n
