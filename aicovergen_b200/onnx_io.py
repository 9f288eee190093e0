"""Minimal ONNX reader (hand-decoded protobuf wire format — no `onnx`/`onnxruntime` package here) and the mapping of a
UVR MDX-Net ("ConvTDFNet") graph to the state dict of `aicovergen_b200.mdx.ConvTDFNetB200`.

Reference call site: `ort.InferenceSession(model_path)` (mdx.py:74); the `.onnx` files come from the TRvlvr model repo
(download_models.py:23-26) and are absent here, so the graph→architecture mapping is restated from the public KUIELab
ConvTDFNet export shape and is `parity unpinned`; it is exercised by a round trip through `write_convtdfnet_onnx`, which
emits the node sequence a `torch.onnx.export` of that module produces in eval mode (Conv+BatchNorm folded or explicit,
Linear as MatMul with a [in, out] initialiser).

Wire format used (protobuf): ModelProto.graph = field 7; GraphProto.node = 1, .initializer = 5; NodeProto.input = 1,
.output = 2, .name = 3, .op_type = 4, .attribute = 5; AttributeProto.name = 1, .f = 2, .i = 3, .ints = 8;
TensorProto.dims = 1, .data_type = 2, .float_data = 4, .name = 8, .raw_data = 9.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Tuple

import numpy as np

BN_EPS = 1e-5


# ----------------------------------------------------------------------------------------------- wire format
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Yields (field number, wire type, value) — value is an int for varint/fixed, a memoryview-able bytes slice for type 2."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_or_single_ints(wt: int, v) -> List[int]:
    if wt == 0:
        return [_signed(v)]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_signed(x))
    return out


_DTYPES = {1: "<f4", 6: "<i4", 7: "<i8", 10: "<f2", 11: "<f8"}


def _tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype, name, raw, floats = 1, "", None, []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_or_single_ints(wt, v)
        elif fno == 2:
            dtype = v
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 4:
            floats += list(struct.unpack(f"<{len(v) // 4}f", v)) if wt == 2 else [struct.unpack("<f", v)[0]]
    if dtype not in _DTYPES:
        raise ValueError(f"initializer {name!r}: unsupported ONNX data_type {dtype}")
    arr = np.frombuffer(raw, dtype=_DTYPES[dtype]).copy() if raw is not None else np.asarray(floats, dtype=np.float32)
    return name, arr.reshape(dims) if dims else arr


@dataclass
class Node:
    op_type: str
    inputs: List[str]
    outputs: List[str]
    name: str = ""
    attrs: Dict[str, object] = field(default_factory=dict)


def _node(buf: bytes) -> Node:
    nd = Node("", [], [])
    for fno, wt, v in _fields(buf):
        if fno == 1:
            nd.inputs.append(bytes(v).decode())
        elif fno == 2:
            nd.outputs.append(bytes(v).decode())
        elif fno == 3:
            nd.name = bytes(v).decode()
        elif fno == 4:
            nd.op_type = bytes(v).decode()
        elif fno == 5:
            an, val = "", None
            ints: List[int] = []
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    an = bytes(v2).decode()
                elif f2 == 2:
                    val = struct.unpack("<f", v2)[0]
                elif f2 == 3:
                    val = _signed(v2)
                elif f2 == 8:
                    ints += _packed_or_single_ints(w2, v2)
            nd.attrs[an] = ints if ints else val
    return nd


def read_onnx(path: str) -> Tuple[Dict[str, np.ndarray], List[Node]]:
    """(initializers by name, nodes in graph order) of an ONNX file."""
    buf = open(path, "rb").read()
    graph = None
    for fno, wt, v in _fields(buf):
        if fno == 7 and wt == 2:
            graph = v
    if graph is None:
        raise ValueError(f"{path}: no GraphProto (field 7) in the model")
    inits: Dict[str, np.ndarray] = {}
    nodes: List[Node] = []
    for fno, wt, v in _fields(graph):
        if fno == 5 and wt == 2:
            name, arr = _tensor(v)
            inits[name] = arr
        elif fno == 1 and wt == 2:
            nodes.append(_node(v))
    return inits, nodes


# ----------------------------------------------------------------------------------------------- ConvTDFNet mapping
@dataclass
class _Op:
    kind: str                 # conv | convT | linear
    weight: np.ndarray        # torch layout: conv [Cout,Cin,kh,kw], convT [Cin,Cout,kh,kw], linear [out,in]
    bias: np.ndarray | None
    bn: Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray] | None = None   # (weight, bias, mean, var)


def _compute_ops(inits: Dict[str, np.ndarray], nodes: List[Node]) -> List[_Op]:
    ops: List[_Op] = []
    for nd in nodes:
        t = nd.op_type
        if t in ("Conv", "ConvTranspose"):
            w = inits[nd.inputs[1]].astype(np.float32)
            b = inits[nd.inputs[2]].astype(np.float32) if len(nd.inputs) > 2 and nd.inputs[2] in inits else None
            ops.append(_Op("conv" if t == "Conv" else "convT", w, b))
        elif t == "MatMul":
            wn = nd.inputs[1] if nd.inputs[1] in inits else (nd.inputs[0] if nd.inputs[0] in inits else None)
            if wn is None:
                continue                                   # activation x activation product: not a layer
            ops.append(_Op("linear", np.ascontiguousarray(inits[wn].astype(np.float32).T), None))   # [in,out] -> [out,in]
        elif t == "Gemm":
            w = inits[nd.inputs[1]].astype(np.float32)
            if not nd.attrs.get("transB", 0):
                w = np.ascontiguousarray(w.T)
            b = inits[nd.inputs[2]].astype(np.float32) if len(nd.inputs) > 2 and nd.inputs[2] in inits else None
            ops.append(_Op("linear", w, b))
        elif t == "BatchNormalization":
            if not ops or ops[-1].bn is not None:
                raise ValueError("BatchNormalization without a preceding Conv/MatMul")
            ops[-1].bn = tuple(inits[n].astype(np.float32) for n in nd.inputs[1:5])
    return ops


def convtdfnet_state_dict(path: str, dim_t: int) -> Dict[str, "object"]:
    """State dict (torch tensors, keys of ConvTDFNetB200) from a ConvTDFNet `.onnx` file.  `dim_t` comes from the
    model_data.json entry (mdx.py:81-90): the graph itself is shape-agnostic along time."""
    import torch

    inits, nodes = read_onnx(path)
    ops = _compute_ops(inits, nodes)
    if len(ops) < 4 or ops[0].kind != "conv" or ops[-1].kind != "conv":
        raise ValueError(f"{path}: not a ConvTDFNet graph (first/last layer must be 1x1 convolutions)")
    n = sum(1 for o in ops if o.kind == "convT")
    g, dim_c = int(ops[0].weight.shape[0]), int(ops[0].weight.shape[1])
    l = 0
    while 1 + l < len(ops) and ops[1 + l].kind == "conv" and ops[1 + l].weight.shape[-1] > 1:
        l += 1
    k = int(ops[1].weight.shape[-1])
    lin0 = ops[1 + l]
    if lin0.kind != "linear":
        raise ValueError(f"{path}: expected the TDF Linear after {l} TFC convolutions, found {lin0.kind}")
    dim_f, bnf = int(lin0.weight.shape[1]), int(lin0.weight.shape[1] // lin0.weight.shape[0])
    sd: Dict[str, torch.Tensor] = {}
    it = iter(ops)

    def put_bn(name, c, bn):
        if bn is None:      # folded into the preceding layer by the exporter: identity
            bn = (np.ones(c, np.float32), np.zeros(c, np.float32), np.zeros(c, np.float32), np.full(c, 1.0 - BN_EPS, np.float32))
        for key, arr in zip(("weight", "bias", "running_mean", "running_var"), bn):
            sd[f"{name}.{key}"] = torch.from_numpy(np.ascontiguousarray(arr))

    def take(kind, conv_name, bn_name, out_ch_axis=0):
        op = next(it)
        if op.kind != kind:
            raise ValueError(f"{path}: expected {kind} for {conv_name}, found {op.kind}")
        sd[conv_name + ".weight"] = torch.from_numpy(op.weight)
        c = op.weight.shape[out_ch_axis]
        if kind != "linear":
            sd[conv_name + ".bias"] = torch.from_numpy(op.bias if op.bias is not None else np.zeros(c, np.float32))
        return op, c

    def tfc_tdf(prefix, c):
        for j in range(l):
            op, _ = take("conv", f"{prefix}.tfc.H.{j}.0", None)
            put_bn(f"{prefix}.tfc.H.{j}.1", c, op.bn)
        op, _ = take("linear", f"{prefix}.tdf.0", None)
        put_bn(f"{prefix}.tdf.1", c, op.bn)
        op, _ = take("linear", f"{prefix}.tdf.3", None)
        put_bn(f"{prefix}.tdf.4", c, op.bn)

    op, c = take("conv", "first_conv.0", None)
    put_bn("first_conv.1", c, op.bn)
    for i in range(n):
        tfc_tdf(f"encoding_blocks.{i}", c)
        op, c = take("conv", f"ds.{i}.0", None)
        put_bn(f"ds.{i}.1", c, op.bn)
    tfc_tdf("bottleneck_block", c)
    for i in range(n):
        op, c = take("convT", f"us.{i}.0", None, out_ch_axis=1)
        put_bn(f"us.{i}.1", c, op.bn)
        tfc_tdf(f"decoding_blocks.{i}", c)
    take("conv", "final_conv.0", None)
    if next(it, None) is not None:
        raise ValueError(f"{path}: trailing layers after final_conv")
    sd["_meta"] = torch.tensor([dim_f, dim_t, g, l, n, bnf, k, dim_c])
    return sd


# ----------------------------------------------------------------------------------------------- writer (tests / export)
def _enc_varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(fno: int, payload: bytes) -> bytes:
    return _enc_varint((fno << 3) | 2) + _enc_varint(len(payload)) + payload


def _tensor_proto(name: str, arr: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(arr, dtype="<f4")
    out = b"".join(_enc_varint((1 << 3) | 0) + _enc_varint(int(d)) for d in arr.shape)
    out += _enc_varint((2 << 3) | 0) + _enc_varint(1)
    out += _ld(8, name.encode()) + _ld(9, arr.tobytes())
    return out


def _node_proto(op_type: str, inputs: List[str], outputs: List[str], ints: Dict[str, List[int]] | None = None) -> bytes:
    out = b"".join(_ld(1, s.encode()) for s in inputs) + b"".join(_ld(2, s.encode()) for s in outputs)
    out += _ld(4, op_type.encode())
    for an, vals in (ints or {}).items():
        a = _ld(1, an.encode()) + b"".join(_enc_varint((8 << 3) | 0) + _enc_varint(v) for v in vals)
        a += _enc_varint((20 << 3) | 0) + _enc_varint(7)          # AttributeProto.type = INTS
        out += _ld(5, a)
    return out


def write_convtdfnet_onnx(path: str, sd: Dict[str, "object"], fold_bn: bool = True) -> None:
    """Serialises a ConvTDFNet state dict as the ONNX graph an eval-mode export produces: Conv(+folded BatchNorm when
    `fold_bn`) → Relu, Transpose, TFC convs, TDF MatMul → BatchNormalization → Relu, …  Only what the reader needs is
    exact (node order, op types, initialiser layouts); value_info and opset records are omitted."""
    meta = [int(v) for v in sd["_meta"]]
    dim_f, dim_t, g, l, n, bnf, k, dim_c = meta
    nodes: List[bytes] = []
    inits: List[bytes] = []
    cnt = [0]

    def A(name):
        return np.asarray(sd[name].detach().cpu().numpy() if hasattr(sd[name], "detach") else sd[name], dtype=np.float32)

    def tname():
        cnt[0] += 1
        return f"t{cnt[0]}"

    def bn_params(p):
        return A(p + ".weight"), A(p + ".bias"), A(p + ".running_mean"), A(p + ".running_var")

    def emit_bn(x, p):
        y = tname()
        names = []
        for key, arr in zip(("weight", "bias", "running_mean", "running_var"), bn_params(p)):
            nm = f"{p}.{key}"
            inits.append(_tensor_proto(nm, arr))
            names.append(nm)
        nodes.append(_node_proto("BatchNormalization", [x] + names, [y]))
        return y

    def conv(x, cp, bp, op="Conv", ch_axis=0):
        w, b = A(cp + ".weight"), A(cp + ".bias")
        if fold_bn:
            gam, bet, mu, var = bn_params(bp)
            s = gam / np.sqrt(var + BN_EPS)
            shape = [1] * w.ndim
            shape[ch_axis] = -1
            w, b = w * s.reshape(shape), (b - mu) * s + bet
        wn, bnm, y = f"onnx::{op}_{cnt[0]}w", f"onnx::{op}_{cnt[0]}b", tname()
        inits.extend([_tensor_proto(wn, w), _tensor_proto(bnm, b)])
        nodes.append(_node_proto(op, [x, wn, bnm], [y], {"kernel_shape": list(w.shape[2:]), "strides": [1, 1]}))
        if not fold_bn:
            y = emit_bn(y, bp)
        r = tname()
        nodes.append(_node_proto("Relu", [y], [r]))
        return r

    def linear(x, wp, bp):
        wn, y = f"onnx::MatMul_{cnt[0]}", tname()
        inits.append(_tensor_proto(wn, A(wp + ".weight").T))          # exporter stores Linear weights as [in, out]
        nodes.append(_node_proto("MatMul", [x, wn], [y]))
        y = emit_bn(y, bp)
        r = tname()
        nodes.append(_node_proto("Relu", [y], [r]))
        return r

    def tfc_tdf(x, p):
        for j in range(l):
            x = conv(x, f"{p}.tfc.H.{j}.0", f"{p}.tfc.H.{j}.1")
        t = linear(x, f"{p}.tdf.0", f"{p}.tdf.1")
        t = linear(t, f"{p}.tdf.3", f"{p}.tdf.4")
        y = tname()
        nodes.append(_node_proto("Add", [x, t], [y]))
        return y

    x = conv("input", "first_conv.0", "first_conv.1")
    t = tname()
    nodes.append(_node_proto("Transpose", [x], [t], {"perm": [0, 1, 3, 2]}))
    x = t
    skips = []
    for i in range(n):
        x = tfc_tdf(x, f"encoding_blocks.{i}")
        skips.append(x)
        x = conv(x, f"ds.{i}.0", f"ds.{i}.1")
    x = tfc_tdf(x, "bottleneck_block")
    for i in range(n):
        x = conv(x, f"us.{i}.0", f"us.{i}.1", op="ConvTranspose", ch_axis=1)
        m = tname()
        nodes.append(_node_proto("Mul", [x, skips[-1 - i]], [m]))
        x = tfc_tdf(m, f"decoding_blocks.{i}")
    t = tname()
    nodes.append(_node_proto("Transpose", [x], [t], {"perm": [0, 1, 3, 2]}))
    wn, bnm = "final_conv.0.weight", "final_conv.0.bias"
    inits.extend([_tensor_proto(wn, A(wn)), _tensor_proto(bnm, A(bnm))])
    nodes.append(_node_proto("Conv", [t, wn, bnm], ["output"], {"kernel_shape": [1, 1]}))
    graph = b"".join(_ld(1, nd) for nd in nodes) + _ld(2, b"ConvTDFNet") + b"".join(_ld(5, t_) for t_ in inits)
    model = _enc_varint((1 << 3) | 0) + _enc_varint(7) + _ld(7, graph)
    with open(path, "wb") as f:
        f.write(model)
