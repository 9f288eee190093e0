"""Device-side IVF-Flat (nprobe=1) feature index — drop-in for the faiss index object used at
vc_infer_pipeline.py:505-507 (`read_index`, `reconstruct_n`, `ntotal`) and :421 (`search(npy, k=8)`).

Coarse quantiser = exact-fp32 SIMT tap-GEMM (|c|^2 - 2 q.c) + row argmin; the list scan, top-8, the
`(1/d)^2` weighting, reconstruction and `index_rate` blend are one kernel (b200vc_ivf_scan_blend), so
features never leave HBM (the reference does D2H -> faiss-cpu -> H2D, :414-431).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _ffi, ops
from . import tapgemm as tg
from .tapgemm import Epi


class IvfIndexB200:
    def __init__(self, centroids: np.ndarray, vectors: np.ndarray, device: str = "cuda:0", list_of: Optional[np.ndarray] = None):
        """centroids [nlist, d], vectors [ntotal, d] in insertion (id) order.  `list_of` [ntotal]: the inverted list of
        every vector as stored in an index file; when absent vectors go to their nearest centroid (what faiss' add does)."""
        self.device = torch.device(device)
        with torch.cuda.device(self.device):      # the list assignment below launches kernels (see _ffi.on_device)
            self._build(centroids, vectors, list_of)

    def _build(self, centroids, vectors, list_of):
        cent = torch.from_numpy(np.ascontiguousarray(centroids, dtype=np.float32))
        vecs = torch.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32))
        self.ntotal, self.d = int(vecs.shape[0]), int(vecs.shape[1])
        self.nlist = int(cent.shape[0])
        self._host_vectors = vecs.numpy()
        self.cent_m2 = (cent * -2.0).contiguous().to(self.device)          # GEMM gives -2 q.c
        self.cent_n2 = (cent.double() ** 2).sum(1).float().to(self.device)  # + |c|^2 as per-column bias
        # ---- inverted lists: assign every database vector to its nearest centroid (on the device, same kernels)
        vdev = vecs.to(self.device)
        assign = self._coarse(vdev) if list_of is None else torch.from_numpy(np.ascontiguousarray(list_of, dtype=np.int64)).to(self.device)
        order = torch.argsort(assign.long(), stable=True)                   # list order == insertion order
        self.ids = order.contiguous()                                        # int64: sorted position -> original id
        self.vecs_sorted = vdev[order].contiguous()
        counts = torch.bincount(assign.long(), minlength=self.nlist)
        offs = torch.zeros(self.nlist + 1, dtype=torch.int64, device=self.device)
        offs[1:] = torch.cumsum(counts, 0)
        self.offsets = offs.to(torch.int32)
        del vdev

    # ---- faiss-compatible surface -------------------------------------------------
    def reconstruct_n(self, i0: int, n: int) -> np.ndarray:
        return self._host_vectors[i0:i0 + n].copy()

    @_ffi.on_device
    def search(self, x: np.ndarray, k: int = 8) -> Tuple[np.ndarray, np.ndarray]:
        """(squared-L2 [T,k], ids [T,k]); only k=8 is on the RVC path."""
        if k != 8:
            raise ValueError("IvfIndexB200.search supports k=8 (vc_infer_pipeline.py:421)")
        q = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(self.device)
        out = torch.empty_like(q)
        D = torch.empty(q.shape[0], 8, device=self.device)
        I = torch.empty(q.shape[0], 8, device=self.device, dtype=torch.int64)
        self._scan(q, out, 0.0, D, I)
        return D.cpu().numpy(), I.cpu().numpy()

    # ---- device path used by VC.vc ---------------------------------------------------
    def _coarse(self, q: torch.Tensor) -> torch.Tensor:
        T = q.shape[0]
        assign = torch.empty(T, dtype=torch.int32, device=self.device)
        step = 32768
        for s in range(0, T, step):
            qb = q[s:s + step]
            sc = torch.empty(qb.shape[0], self.nlist, device=self.device)
            tg.linear(qb, self.cent_m2, sc, Epi(bias=self.cent_n2), backend=tg.BACKEND_SIMT, name="ivf.coarse")()
            ops.argmin_rows(sc, assign[s:s + step])
        return assign

    def _scan(self, q, out, rate, D=None, I=None):
        assign = self._coarse(q)
        ops.ivf_scan_blend(q, assign, self.offsets, self.ids, self.vecs_sorted, out, rate, D, I)

    @_ffi.on_device
    @torch.no_grad()
    def search_blend(self, feats: torch.Tensor, index_rate: float) -> torch.Tensor:
        """feats [T, d] (device, fp32) -> index_rate * weighted-NN reconstruction + (1-index_rate) * feats."""
        q = feats.contiguous()
        out = torch.empty_like(q)
        self._scan(q, out, float(index_rate))
        return out


def write_index_npz(path: str, centroids: np.ndarray, vectors: np.ndarray):
    """Portable container for an IVF-Flat index (centroids + vectors in id order)."""
    np.savez(path, centroids=np.asarray(centroids, dtype=np.float32), vectors=np.asarray(vectors, dtype=np.float32))


def read_index(path: str, device: str = "cuda:0") -> IvfIndexB200:
    """Counterpart of faiss.read_index at vc_infer_pipeline.py:505: faiss' binary IndexIVFFlat files ("IwFl", what RVC
    voice models ship) through aicovergen_b200.faiss_io, or the `.npz` form of write_index_npz.  Unsupported index types
    raise (the reference's own behaviour on a failed read is to continue without an index, :508-510)."""
    if str(path).endswith(".npz"):
        z = np.load(path)
        return IvfIndexB200(z["centroids"], z["vectors"], device)
    from .faiss_io import METRIC_L2, read_ivfflat
    data = read_ivfflat(str(path))
    if data.metric != METRIC_L2:
        raise NotImplementedError(f"{path}: metric {data.metric}; the RVC path uses L2 IVF-Flat indexes")
    if not data.ids_sequential:
        raise NotImplementedError(f"{path}: custom ids (add_with_ids) are not addressable by reconstruct_n(0, ntotal)")
    return IvfIndexB200(data.centroids, data.vectors, device, list_of=data.list_of)
