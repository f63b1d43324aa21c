"""Pins for the oracle (it restates an un-vendored dependency, so it is checked against the third-party
implementations that ARE installed here: transformers' LlamaModel and logits processors, torch.stft/istft) and for
the host-side packing rules the CUDA kernels rely on (implicit-GEMM conv, transposed-conv phases, fragment order)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cbx_b200.config import ModelConfig, T3Config, HiFTConfig
from cbx_b200.pack import conv_pack, convT_pack, frag_order, llama3_inv_freq
from cbx_b200.weights import random_state_dict, schema, synthetic_conditionals
from oracle import flow as OF, hift as OH, t3 as OT


def test_schema_counts_match_survey():
    """16 779 264 parameters per Llama layer, 66.08 M MACs per frame for the CFM estimator."""
    S = schema(ModelConfig())
    per_layer = sum(math.prod(sh) for n, (sh, _) in S.items() if n.startswith("t3.tfmr.layers.0."))
    assert per_layer == 16779264
    est = sum(math.prod(sh) for n, (sh, _) in S.items() if n.startswith("flow.decoder.estimator.") and n.endswith("weight") and len(sh) >= 2
              and "time_mlp" not in n and ".mlp.1." not in n)
    assert abs(est - 66.08e6) / 66.08e6 < 0.01


def test_llama_trunk_matches_transformers():
    from transformers import LlamaConfig, LlamaModel
    c = T3Config(n_layers=2)
    sd = {k: v for k, v in random_state_dict(ModelConfig(t3=c), 3, parts=("t3",)).items()}
    hf = LlamaModel(LlamaConfig(vocab_size=8, hidden_size=c.dim, intermediate_size=c.ffn, num_hidden_layers=c.n_layers, num_attention_heads=c.n_heads,
                                num_key_value_heads=c.n_heads, head_dim=c.head_dim, rms_norm_eps=c.rms_eps, rope_theta=c.rope_theta, hidden_act="silu",
                                attention_bias=False, mlp_bias=False, max_position_embeddings=131072,
                                rope_scaling=dict(factor=8.0, high_freq_factor=4.0, low_freq_factor=1.0, original_max_position_embeddings=8192, rope_type="llama3")))
    hsd = hf.state_dict()
    for k in hsd:
        if k == "embed_tokens.weight":
            continue
        hsd[k].copy_(sd["t3.tfmr." + k])
    hf.eval()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 9, c.dim, generator=g) * 0.05
    with torch.no_grad():
        ref = hf(inputs_embeds=x, use_cache=True)
        mine, kv = OT.llama_forward(sd, c, x)
        assert torch.allclose(mine, ref.last_hidden_state, atol=2e-5, rtol=1e-4)
        x1 = torch.randn(2, 1, c.dim, generator=g) * 0.05
        ref1 = hf(inputs_embeds=x1, past_key_values=ref.past_key_values, use_cache=True)
        mine1, _ = OT.llama_forward(sd, c, x1, kv)
        assert torch.allclose(mine1, ref1.last_hidden_state, atol=2e-5, rtol=1e-4)
    assert torch.allclose(llama3_inv_freq(c), OT.llama3_inv_freq(c))


def test_sampler_matches_transformers_processors():
    from transformers import MinPLogitsWarper, RepetitionPenaltyLogitsProcessor, TopPLogitsWarper
    g = torch.Generator().manual_seed(1)
    for trial in range(5):
        lg = torch.randn(2, 8194, generator=g) * 2.0
        hist = torch.randint(0, 8194, (1, 40), generator=g)
        w, temp, rp, mp, tp = 0.5, 0.8, 1.2, 0.05, 0.9
        mine = OT.process_logits(lg, hist[0].tolist(), w, temp, rp, mp, tp)
        x = (lg[0] + w * (lg[0] - lg[1]))[None]
        x = RepetitionPenaltyLogitsProcessor(penalty=rp)(hist, x.clone())
        x = x / temp
        x = MinPLogitsWarper(min_p=mp)(None, x)
        x = TopPLogitsWarper(top_p=tp)(None, x)
        assert torch.equal(torch.isinf(mine), torch.isinf(x[0]))
        keep = ~torch.isinf(mine)
        assert torch.allclose(mine[keep], x[0][keep])
    # multinomial restated as argmax(p / Exp(1))
    fl = OT.process_logits(lg, [6561], 0.5, 0.8, 1.2, 0.05, 0.95)
    q = torch.empty(8194).exponential_(generator=g)
    tok = OT.sample_from(fl, q)
    assert not math.isinf(fl[tok]) and 0 <= tok < 8194


def test_rel_shift_and_relpos_table():
    T, D = 7, 16
    pe = OF.espnet_rel_pos_emb(T, D)
    assert pe.shape == (2 * T - 1, D)
    div = torch.exp(torch.arange(0, D, 2).float() * -(math.log(10000.0) / D))
    for j in (0, 3, 6, 12):
        rel = T - 1 - j
        assert torch.allclose(pe[j, 0::2], torch.sin(rel * div), atol=1e-6) and torch.allclose(pe[j, 1::2], torch.cos(rel * div), atol=1e-6)
    x = torch.arange(T * (2 * T - 1)).float().view(1, 1, T, 2 * T - 1)
    y = OF._rel_shift(x)
    for i in range(T):
        for j in range(T):
            assert y[0, 0, i, j] == x[0, 0, i, j + T - 1 - i]      # the index the CUDA attention kernel adds as bias


def test_implicit_gemm_conv_packing():
    """conv1d (dilated, causal or 'same') == rows of a zero-haloed channels-last buffer times conv_pack(W)^T."""
    g = torch.Generator().manual_seed(2)
    for (cin, cout, k, d) in [(16, 8, 3, 1), (8, 8, 7, 3), (8, 16, 11, 5)]:
        T = 40
        x = torch.randn(1, cin, T, generator=g)
        w = torch.randn(cout, cin, k, generator=g)
        pad = d * (k - 1) // 2
        ref = F.conv1d(x, w, dilation=d, padding=pad)[0].t()
        halo = 32
        buf = torch.zeros(halo + T + halo, cin)
        buf[halo:halo + T] = x[0].t()
        W = conv_pack(w)
        rows = torch.stack([torch.cat([buf[halo - pad + t + j * d] for j in range(k)]) for t in range(T)])
        assert torch.allclose(rows @ W.t(), ref, atol=1e-4)


def test_transposed_conv_phase_packing():
    """ConvTranspose1d == one GEMM over all u output phases with the row/phase scatter used by hift.cu."""
    g = torch.Generator().manual_seed(3)
    for (u, k) in [(8, 16), (5, 11), (3, 7)]:
        cin, cout, T = 6, 4, 13
        p = (k - u) // 2
        x = torch.randn(1, cin, T, generator=g)
        w = torch.randn(cin, cout, k, generator=g)
        b = torch.randn(cout, generator=g)
        ref = F.conv_transpose1d(x, w, b, stride=u, padding=p)[0].t()       # [T*u][cout]
        W, bias = convT_pack(w, b, u)
        taps = (k + u - 1) // u
        buf = torch.zeros(taps - 1 + T + 1, cin)
        buf[taps - 1: taps - 1 + T] = x[0].t()
        out = torch.zeros(T * u, cout)
        for q in range(T + 1):
            row = buf[q: q + taps].reshape(-1) @ W.t() + bias
            for r in range(u):
                t = q * u + r - p
                if 0 <= t < T * u:
                    out[t] = row[r * cout:(r + 1) * cout]
        assert torch.allclose(out, ref, atol=1e-4)


def test_fragment_order_is_a_permutation():
    w = torch.arange(32 * 48, dtype=torch.float32).view(32, 48)
    f = frag_order(w)
    assert f.shape == (2, 3, 32, 8) and sorted(f.flatten().tolist()) == w.flatten().tolist()
    lane = 13  # g = 3, tg = 1 -> rows 3 / 11, cols 2,3 and 10,11 of tile (1, 2)
    tile = w[16:32, 32:48]
    assert f[1, 2, lane].tolist() == [tile[3, 2], tile[3, 3], tile[11, 2], tile[11, 3], tile[3, 10], tile[3, 11], tile[11, 10], tile[11, 11]]


def test_hift_source_and_istft_properties():
    hc = HiFTConfig()
    sd = random_state_dict(ModelConfig(), 0, parts=("hift",))
    T = 6
    f0 = torch.tensor([[0.0, 5.0, 120.0, 180.0, 90.0, 9.9]])
    g = torch.Generator().manual_seed(5)
    ph = torch.rand(9, generator=g)
    nz = torch.randn(9, T * 480, generator=g)
    s = OH.sine_source(sd, hc, f0, ph, nz)
    assert s.shape == (1, 1, T * 480) and s.abs().max() <= 1.0
    # decode output length and iSTFT/STFT consistency of the 16-point transform the kernels hand-code
    x = torch.randn(1, 480 * 4, generator=g)
    win = OH.hann(16)
    X = torch.stft(x, 16, 4, 16, window=win, return_complex=True)
    assert X.shape[-1] == 120 * 4 + 1
    y = torch.istft(X, 16, 4, 16, window=win)
    assert y.shape[-1] == 480 * 4 and torch.allclose(y, x[:, : y.shape[-1]], atol=1e-5)
    tf = OH.trim_fade(24000)
    assert tf.shape == (960,) and (tf[:480] == 0).all() and abs(float(tf[-1]) - 1.0) < 1e-6


def test_oracle_end_to_end_tiny_is_deterministic():
    cfg = ModelConfig.tiny()
    sd = random_state_dict(cfg, 0)
    conds = synthetic_conditionals(cfg, prompt_tokens=10)
    text = torch.tensor([[255, 5, 6, 7, 0]] * 2)
    g = torch.Generator().manual_seed(0)
    noise = [torch.empty(8194).exponential_(generator=g) for _ in range(3)]
    with torch.no_grad():
        a = list(OT.inference_stream(sd, cfg.t3, conds["t3"], text, 3, noise_fn=lambda i: noise[i]))
        b = list(OT.inference_stream(sd, cfg.t3, conds["t3"], text, 3, noise_fn=lambda i: noise[i]))
        assert a == b and len(a) == 3
        mel = OF.flow_inference(sd, cfg.flow, torch.tensor([1, 2, 3]), conds["gen"])
        assert mel.shape == (1, 80, 6) and torch.isfinite(mel).all()
