"""On-disk formats: faiss IndexIVFFlat binary files."""
import struct

import numpy as np
import pytest

from aicovergen_b200 import faiss_io


def _toy(nlist=7, d=12, n=200, seed=0):
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((nlist, d)).astype(np.float32)
    vec = rng.standard_normal((n, d)).astype(np.float32)
    d2 = ((vec[:, None, :] - cent[None]) ** 2).sum(-1)
    return cent, vec, d2.argmin(1)


@pytest.mark.parametrize("nlist,n", [(7, 200), (40, 25), (3, 0)])
def test_ivfflat_round_trip(tmp_path, nlist, n):
    cent, vec, lof = _toy(nlist=nlist, n=n)
    p = str(tmp_path / "added_IVF7_Flat_nprobe_1_toy.index")
    faiss_io.write_ivfflat(p, cent, vec, lof, nprobe=1)
    data = faiss_io.read_ivfflat(p)
    assert (data.d, data.nlist, data.nprobe, data.metric) == (12, nlist, 1, faiss_io.METRIC_L2)
    assert data.ids_sequential
    np.testing.assert_array_equal(data.centroids, cent)
    np.testing.assert_array_equal(data.vectors, vec)          # reconstruct_n(0, ntotal) order
    np.testing.assert_array_equal(data.list_of, lof)


def test_ivfflat_byte_layout_matches_faiss_headers(tmp_path):
    """Field widths of the published faiss 1.7 serialisation: 4-byte fourcc, 33-byte index header, u64 nlist/nprobe."""
    cent, vec, lof = _toy()
    p = str(tmp_path / "t.index")
    faiss_io.write_ivfflat(p, cent, vec, lof)
    b = open(p, "rb").read()
    assert b[:4] == b"IwFl"
    d, ntotal, dummy1, dummy2, trained, metric = struct.unpack_from("<iqqqBi", b, 4)
    assert (d, ntotal, dummy1, dummy2, trained, metric) == (12, 200, 1 << 20, 1 << 20, 1, 1)
    nlist, nprobe = struct.unpack_from("<QQ", b, 4 + 33)
    assert (nlist, nprobe) == (7, 1)
    assert b[4 + 33 + 16: 4 + 33 + 20] == b"IxF2"
    assert b.find(b"ilar") > 0 and b.find(b"full") > b.find(b"ilar")


def test_ivfflat_hand_assembled_faiss_bytes(tmp_path):
    """A file assembled field by field from faiss 1.7.x's serialisation (index_write.cpp: fourcc "IwFl" + write_ivf_header
    + write_InvertedLists; NO code_size between the direct map and "ilar" for IndexIVFFlat), without the repo's writer:
    d=2, nlist=3 (list 1 empty -> exercises skipping), ntotal=4 added in id order, maintained direct map (type 1: Array)."""
    d, nlist = 2, 3
    cent = np.array([[0, 0], [10, 10], [5, 5]], dtype="<f4")
    vec = np.array([[0.1, 0], [5, 5.2], [0, 0.3], [4.9, 5]], dtype="<f4")      # ids 0..3 -> lists 0, 2, 0, 2
    hdr = lambda n: struct.pack("<i", d) + struct.pack("<q", n) + struct.pack("<qq", 1 << 20, 1 << 20) + b"\x01" + struct.pack("<i", 1)
    b = b"IwFl" + hdr(4) + struct.pack("<QQ", nlist, 1)
    b += b"IxF2" + hdr(nlist) + struct.pack("<Q", nlist * d) + cent.tobytes()
    # direct map, type Array: entry id -> (list_no << 32 | offset)
    dm = np.array([(0 << 32) | 0, (2 << 32) | 0, (0 << 32) | 1, (2 << 32) | 1], dtype="<i8")
    b += b"\x01" + struct.pack("<Q", 4) + dm.tobytes()
    b += b"ilar" + struct.pack("<QQ", nlist, 4 * d) + b"full" + struct.pack("<Q", nlist) + np.array([2, 0, 2], dtype="<u8").tobytes()
    b += vec[[0, 2]].tobytes() + np.array([0, 2], dtype="<i8").tobytes()
    b += vec[[1, 3]].tobytes() + np.array([1, 3], dtype="<i8").tobytes()
    assert len(b) == 4 + 33 + 16 + 4 + 33 + 8 + 24 + 1 + 8 + 32 + 4 + 16 + 4 + 8 + 24 + 2 * (16 + 16)
    p = tmp_path / "added_IVF3_Flat_nprobe_1_hand.index"
    p.write_bytes(b)
    data = faiss_io.read_ivfflat(str(p))
    assert (data.d, data.nlist, data.nprobe, data.metric, data.ids_sequential) == (2, 3, 1, faiss_io.METRIC_L2, True)
    np.testing.assert_array_equal(data.centroids, cent)
    np.testing.assert_array_equal(data.vectors, vec)
    np.testing.assert_array_equal(data.list_of, [0, 2, 0, 2])
    # the legacy layout with a u64 code_size before "ilar" must be rejected, not misparsed
    bad = b.replace(b"ilar", struct.pack("<Q", 4 * d) + b"ilar")
    p.write_bytes(bad)
    with pytest.raises(faiss_io.FaissFormatError):
        faiss_io.read_ivfflat(str(p))
    # and the writer emits exactly these bytes for the same index without a direct map
    q = tmp_path / "w.index"
    faiss_io.write_ivfflat(str(q), cent, vec, np.array([0, 2, 0, 2]))
    want = b.replace(b"\x01" + struct.pack("<Q", 4) + dm.tobytes(), b"\x00" + struct.pack("<Q", 0))
    assert q.read_bytes() == want


def test_rejects_other_index_types(tmp_path):
    p = tmp_path / "x.index"
    p.write_bytes(b"IxF2" + b"\0" * 64)
    with pytest.raises(faiss_io.FaissFormatError):
        faiss_io.read_ivfflat(str(p))
    p.write_bytes(b"IwFl" + b"\0" * 10)
    with pytest.raises(faiss_io.FaissFormatError):
        faiss_io.read_ivfflat(str(p))


@pytest.mark.parametrize("fold_bn", [True, False])
def test_onnx_convtdfnet_round_trip(tmp_path, fold_bn):
    """state dict -> ONNX bytes (the node sequence an eval-mode export emits) -> hand-decoded protobuf -> state dict: the
    recovered network computes the same function (oracle net on CPU) and the hyper-parameters are re-derived from shapes."""
    import torch

    from aicovergen_b200 import onnx_io
    from aicovergen_b200.synthetic import make_mdx_state_dict
    from oracle import mdx as om

    sd = make_mdx_state_dict(dim_f=64, dim_t=16, g=8, l=2, n=2, bn=4)
    p = str(tmp_path / "UVR_MDXNET_toy.onnx")
    onnx_io.write_convtdfnet_onnx(p, sd, fold_bn=fold_bn)
    inits, nodes = onnx_io.read_onnx(p)
    assert sum(nd.op_type == "ConvTranspose" for nd in nodes) == 2
    assert sum(nd.op_type == "BatchNormalization" for nd in nodes) == (2 * 5 if fold_bn else 2 * 5 + 1 + 2 * 5 + 2 + 2)
    sd2 = onnx_io.convtdfnet_state_dict(p, dim_t=16)
    assert [int(v) for v in sd2["_meta"]] == [int(v) for v in sd["_meta"]]
    x = torch.randn(1, 4, 64, 16, generator=torch.Generator().manual_seed(0))
    y1, y2 = om.convtdfnet(sd, x), om.convtdfnet(sd2, x)
    assert float((y1 - y2).abs().max()) <= 2e-5 * float(y1.abs().max())
    if not fold_bn:
        for k_ in sd:
            if k_ != "_meta":
                assert torch.equal(sd[k_].float(), sd2[k_].float()), k_


def test_onnx_hand_assembled_protobuf_bytes(tmp_path):
    """A ModelProto assembled BY HAND from the protobuf wire format and onnx.proto's field numbers — not through
    write_convtdfnet_onnx — with what exporter-written files carry and this package's writer does not: ir_version /
    producer / opset_import / doc_string fields to skip, graph inputs / outputs / value_info, unpacked (proto2-style) and packed
    `dims`, `raw_data` and packed `float_data` initialisers, an int64 initialiser, float / int / ints / string / tensor
    attributes, a negative int attribute (10-byte varint)."""
    import struct

    from aicovergen_b200.onnx_io import read_onnx

    def vi(v):                       # varint
        v &= (1 << 64) - 1
        out = bytearray()
        while True:
            b = v & 0x7F
            v >>= 7
            out.append(b | (0x80 if v else 0))
            if not v:
                return bytes(out)

    key = lambda fno, wt: vi((fno << 3) | wt)
    ld = lambda fno, payload: key(fno, 2) + vi(len(payload)) + payload
    w = np.arange(24, dtype=np.float32).reshape(2, 3, 4) - 7.5
    # TensorProto: dims (1) unpacked varints, data_type (2) = 1 FLOAT, name (8), raw_data (9), doc_string (12, skipped)
    t_raw = b"".join(key(1, 0) + vi(d) for d in w.shape) + key(2, 0) + vi(1) + ld(8, b"conv.weight") + ld(9, w.tobytes()) + ld(12, b"doc")
    # TensorProto with packed dims and packed float_data (4)
    b = np.array([0.5, -1.25, 3.0], dtype=np.float32)
    t_flt = ld(1, vi(3)) + key(2, 0) + vi(1) + ld(4, struct.pack("<3f", *b)) + ld(8, b"conv.bias")
    # int64 initialiser (a Reshape target), data_type 7, raw_data
    shp = np.array([1, -1, 4], dtype=np.int64)
    t_i64 = ld(1, vi(3)) + key(2, 0) + vi(7) + ld(8, b"shape") + ld(9, shp.tobytes())
    # AttributeProto: name (1), f (2, fixed32), i (3), s (4), t (5), ints (8), type (20)
    a_ints_packed = ld(1, b"kernel_shape") + ld(8, vi(3) + vi(3)) + key(20, 0) + vi(7)
    a_ints_unpacked = ld(1, b"pads") + b"".join(key(8, 0) + vi(p) for p in (1, 1, 1, 1)) + key(20, 0) + vi(7)
    a_int = ld(1, b"axis") + key(3, 0) + vi(-1) + key(20, 0) + vi(2)
    a_flt = ld(1, b"epsilon") + key(2, 5) + struct.pack("<f", 1e-5) + key(20, 0) + vi(1)
    a_str = ld(1, b"auto_pad") + ld(4, b"NOTSET") + key(20, 0) + vi(3)
    a_tensor = ld(1, b"value") + ld(5, t_i64) + key(20, 0) + vi(4)
    n_conv = ld(1, b"x") + ld(1, b"conv.weight") + ld(1, b"conv.bias") + ld(2, b"y") + ld(3, b"Conv_0") + ld(4, b"Conv") + \
        ld(5, a_ints_packed) + ld(5, a_ints_unpacked) + ld(5, a_str) + ld(7, b"")          # domain (7) = ""
    n_bn = ld(1, b"y") + ld(2, b"z") + ld(4, b"BatchNormalization") + ld(5, a_flt)
    n_cat = ld(1, b"z") + ld(1, b"z") + ld(2, b"c") + ld(4, b"Concat") + ld(5, a_int)
    n_const = ld(2, b"k") + ld(4, b"Constant") + ld(5, a_tensor)
    value_info = ld(1, b"x") + ld(2, ld(1, key(1, 0) + vi(1)))                             # ValueInfoProto{name, type{tensor_type{elem_type}}}
    graph = ld(1, n_conv) + ld(1, n_bn) + ld(1, n_cat) + ld(1, n_const) + ld(2, b"torch_jit") + ld(5, t_raw) + ld(5, t_flt) + \
        ld(5, t_i64) + ld(10, b"graph doc") + ld(11, value_info) + ld(12, value_info) + ld(13, value_info)
    opset = ld(1, b"") + key(2, 0) + vi(11)
    model = key(1, 0) + vi(6) + ld(2, b"pytorch") + ld(3, b"1.10") + ld(7, graph) + ld(8, opset) + key(5, 0) + vi(0)
    path = tmp_path / "hand.onnx"
    path.write_bytes(model)

    inits, nodes = read_onnx(str(path))
    assert set(inits) == {"conv.weight", "conv.bias", "shape"}
    assert inits["conv.weight"].dtype == np.float32 and np.array_equal(inits["conv.weight"], w)
    assert np.array_equal(inits["conv.bias"], b) and inits["conv.bias"].shape == (3,)
    assert inits["shape"].dtype == np.int64 and inits["shape"].tolist() == [1, -1, 4]
    assert [n.op_type for n in nodes] == ["Conv", "BatchNormalization", "Concat", "Constant"]
    conv = nodes[0]
    assert conv.inputs == ["x", "conv.weight", "conv.bias"] and conv.outputs == ["y"] and conv.name == "Conv_0"
    assert conv.attrs["kernel_shape"] == [3, 3] and conv.attrs["pads"] == [1, 1, 1, 1]
    assert abs(nodes[1].attrs["epsilon"] - 1e-5) < 1e-12 and nodes[2].attrs["axis"] == -1
