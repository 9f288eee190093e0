"""Pins oracle/hubert.py (the restatement of fairseq 0.12.2 `HubertModel.extract_features`, which cannot be installed
here) against an INDEPENDENT implementation of the same published architecture: `transformers.HubertModel` with the
default `HubertConfig()` (= hubert-base: conv 512x7 kernels (10,3,3,3,3,2,2) strides (5,2,2,2,2,2,2), GroupNorm front-end,
12 x 768/3072/12 heads, pos-conv 128/16, post-LN), loaded with the same fairseq-named synthetic state dict through
`oracle.hubert.to_hf_state_dict`.  `hidden_states[k]` is fairseq's `output_layer=k`."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("seconds", [1.37, 3.0])
def test_hubert_oracle_matches_transformers(seconds):
    transformers = pytest.importorskip("transformers")
    from aicovergen_b200.synthetic import make_hubert_state_dict
    from oracle import hubert as ohub

    sd = make_hubert_state_dict()
    cfg = transformers.HubertConfig()
    assert (cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.intermediate_size) == (768, 12, 12, 3072)
    assert tuple(cfg.conv_kernel) == (10, 3, 3, 3, 3, 2, 2) and tuple(cfg.conv_stride) == (5, 2, 2, 2, 2, 2, 2)
    assert cfg.feat_extract_norm == "group" and not cfg.do_stable_layer_norm
    assert (cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups) == (128, 16)
    hf = transformers.HubertModel(cfg).eval()
    res = hf.load_state_dict(ohub.to_hf_state_dict(sd), strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k == "masked_spec_embed" for k in res.missing_keys), res.missing_keys
    rng = np.random.default_rng(0)
    n = int(16000 * seconds)
    t = np.arange(n) / 16000.0
    x = (0.1 * rng.standard_normal(n) + 0.4 * np.sin(2 * np.pi * (100 * t + 45 * t * t))).astype(np.float32)
    src = torch.from_numpy(x)[None]
    with torch.no_grad():
        hs = hf(src, output_hidden_states=True).hidden_states
    for layer in (9, 12):         # v1 voice models read layer 9 (+ final_proj), v2 layer 12 (vc_infer_pipeline.py:401-406)
        got = ohub.extract_features(sd, src, layer)
        ref = hs[layer]
        assert got.shape == ref.shape == (1, (n - 400) // 320 + 1, 768)
        err = float((got - ref).abs().max())
        print(f"[hubert oracle vs transformers] {seconds}s layer {layer}: max abs diff {err:.2e} (ref max {float(ref.abs().max()):.2f})")
        assert err < 5e-5
