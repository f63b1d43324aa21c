"""ORACLE (test infrastructure, not product): CPU/PyTorch fp32 restatement of the
S3Gen token->mel path (UpsampleConformerEncoder + CausalConditionalCFM with the
ConditionalDecoder estimator).  PARITY UNPINNED: restates the published algorithm of
the un-vendored dependency chatterbox (reference requirements.txt:9) — upstream files
models/s3gen/flow.py, transformer/upsample_encoder.py, transformer/attention.py,
transformer/embedding.py, flow_matching.py, decoder.py, matcha/transformer.py,
matcha/decoder.py — anchored on the reference call sites src/tts_streaming.py:316-320,
:586-590 (s3gen.inference arguments; finalize defaults to True) and :655-677
(token accumulation, EOS append, <6561 filter, pad to >=3).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.
"""
import math

import torch
import torch.nn.functional as F

from cbx_b200.config import FlowConfig


def _ln(x, sd, p, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


# ----------------------------------------------------------------------------- encoder (K7)
def espnet_rel_pos_emb(T: int, D: int, device=None) -> torch.Tensor:
    """EspnetRelPositionalEncoding.position_encoding: (2T-1, D); row j is relative position T-1-j."""
    pos = torch.arange(T - 1, -T, -1, dtype=torch.float32, device=device)[:, None]
    div = torch.exp(torch.arange(0, D, 2, dtype=torch.float32, device=device) * -(math.log(10000.0) / D))
    pe = torch.zeros(2 * T - 1, D, device=device)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def _rel_shift(x):
    """(B,H,T,2T-1) -> (B,H,T,T): out[i,j] = x[i, j + T-1-i]."""
    B, H, T, N = x.shape
    xp = torch.cat([x.new_zeros(B, H, T, 1), x], dim=-1).view(B, H, N + 1, T)
    return xp[:, :, 1:].reshape(B, H, T, N)[..., : N // 2 + 1]


def _rel_mha(sd, p, x, pos_emb, H):
    B, T, D = x.shape
    dk = D // H
    q = _lin(x, sd, p + "linear_q").view(B, T, H, dk)
    k = _lin(x, sd, p + "linear_k").view(B, T, H, dk).transpose(1, 2)
    v = _lin(x, sd, p + "linear_v").view(B, T, H, dk).transpose(1, 2)
    pp = F.linear(pos_emb, sd[p + "linear_pos.weight"]).view(1, -1, H, dk).transpose(1, 2)
    qu = (q + sd[p + "pos_bias_u"]).transpose(1, 2)
    qv = (q + sd[p + "pos_bias_v"]).transpose(1, 2)
    ac = qu @ k.transpose(-1, -2)
    bd = _rel_shift(qv @ pp.transpose(-1, -2))
    att = ((ac + bd) / math.sqrt(dk)).softmax(-1)
    o = (att @ v).transpose(1, 2).reshape(B, T, D)
    return _lin(o, sd, p + "linear_out")


def _conformer_layer(sd, p, x, pos_emb, H):
    x = x + _rel_mha(sd, p + "self_attn.", _ln(x, sd, p + "norm_mha", 1e-12), pos_emb, H)
    h = _ln(x, sd, p + "norm_ff", 1e-12)
    h = _lin(F.silu(_lin(h, sd, p + "feed_forward.w_1")), sd, p + "feed_forward.w_2")
    return x + h


def _embed(sd, p, x):
    """LinearNoSubsampling + EspnetRelPositionalEncoding (x * sqrt(D), pos_emb)."""
    x = _ln(_lin(x, sd, p + "out.0"), sd, p + "out.1", 1e-5)
    D = x.shape[-1]
    return x * math.sqrt(D), espnet_rel_pos_emb(x.shape[1], D, x.device)[None]


def encoder_forward(sd, c: FlowConfig, tok_emb: torch.Tensor) -> torch.Tensor:
    """(1, Ttok, 512) -> (1, 2*Ttok, 512)."""
    e = "flow.encoder."
    x, pe = _embed(sd, e + "embed.", tok_emb)
    # PreLookaheadLayer
    y = F.pad(x.transpose(1, 2), (0, c.pre_lookahead))
    y = F.leaky_relu(F.conv1d(y, sd[e + "pre_lookahead_layer.conv1.weight"], sd[e + "pre_lookahead_layer.conv1.bias"]))
    y = F.conv1d(F.pad(y, (2, 0)), sd[e + "pre_lookahead_layer.conv2.weight"], sd[e + "pre_lookahead_layer.conv2.bias"])
    x = x + y.transpose(1, 2)
    for i in range(c.enc_blocks):
        x = _conformer_layer(sd, e + f"encoders.{i}.", x, pe, c.enc_heads)
    # Upsample1D: nearest x2, left pad 4, conv k5
    y = F.interpolate(x.transpose(1, 2), scale_factor=2.0, mode="nearest")
    y = F.conv1d(F.pad(y, (4, 0)), sd[e + "up_layer.conv.weight"], sd[e + "up_layer.conv.bias"])
    x, pe = _embed(sd, e + "up_embed.", y.transpose(1, 2))
    for i in range(c.up_blocks):
        x = _conformer_layer(sd, e + f"up_encoders.{i}.", x, pe, c.enc_heads)
    return _ln(x, sd, e + "after_norm", 1e-5)


# ----------------------------------------------------------------------------- estimator (K9)
def _causal_conv(x, w, b):
    return F.conv1d(F.pad(x, (w.shape[-1] - 1, 0)), w, b)


def _causal_block(sd, p, x):
    """CausalBlock1D: CausalConv1d(k3) -> LayerNorm over channels -> Mish.  x: (B,C,T)."""
    h = _causal_conv(x, sd[p + "block.0.weight"], sd[p + "block.0.bias"])
    h = _ln(h.transpose(1, 2), sd, p + "block.2").transpose(1, 2)
    return F.mish(h)


def _resnet(sd, p, x, temb):
    h = _causal_block(sd, p + "block1.", x)
    h = h + _lin(F.mish(temb), sd, p + "mlp.1")[:, :, None]
    h = _causal_block(sd, p + "block2.", h)
    return h + F.conv1d(x, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"])


def _tfm_block(sd, p, x, heads, hd):
    """matcha BasicTransformerBlock (self-attention only, GELU feed-forward).  x: (B,T,C)."""
    B, T, C = x.shape
    n = _ln(x, sd, p + "norm1")
    sp = lambda t: t.view(B, T, heads, hd).transpose(1, 2)
    q, k, v = (sp(F.linear(n, sd[p + f"attn1.{m}.weight"])) for m in ("to_q", "to_k", "to_v"))
    a = ((q @ k.transpose(-1, -2)) * (hd ** -0.5)).softmax(-1)
    o = (a @ v).transpose(1, 2).reshape(B, T, heads * hd)
    x = x + _lin(o, sd, p + "attn1.to_out.0")
    n = _ln(x, sd, p + "norm3")
    return x + _lin(F.gelu(_lin(n, sd, p + "ff.net.0.proj")), sd, p + "ff.net.2")


def time_embedding(sd, c: FlowConfig, t: torch.Tensor) -> torch.Tensor:
    """SinusoidalPosEmb(in_ch, scale 1000) -> TimestepEmbedding(SiLU).  t: (B,) -> (B, 4*ch)."""
    half = c.in_ch // 2
    e = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * -(math.log(10000.0) / (half - 1)))
    e = 1000.0 * t[:, None] * e[None]
    e = torch.cat([e.sin(), e.cos()], dim=-1)
    d = "flow.decoder.estimator.time_mlp."
    return _lin(F.silu(_lin(e, sd, d + "linear_1")), sd, d + "linear_2")


def estimator_forward(sd, c: FlowConfig, x, mu, t, spks, cond):
    """ConditionalDecoder.forward with an all-ones mask.  x,mu,cond: (B,80,T); spks: (B,80); t: (B,)."""
    d = "flow.decoder.estimator."
    temb = time_embedding(sd, c, t)
    T = x.shape[-1]
    h = torch.cat([x, mu, spks[:, :, None].expand(-1, -1, T), cond], dim=1)

    def stage(p, h):
        h = _resnet(sd, p + "0.", h, temb).transpose(1, 2)
        for j in range(c.n_blocks):
            h = _tfm_block(sd, p + f"1.{j}.", h, c.heads, c.head_dim)
        return h.transpose(1, 2)

    h = stage(d + "down_blocks.0.", h)
    skip = h
    h = _causal_conv(h, sd[d + "down_blocks.0.2.weight"], sd[d + "down_blocks.0.2.bias"])
    for i in range(c.n_mid):
        h = stage(d + f"mid_blocks.{i}.", h)
    h = stage(d + "up_blocks.0.", torch.cat([h, skip], dim=1))
    h = _causal_conv(h, sd[d + "up_blocks.0.2.weight"], sd[d + "up_blocks.0.2.bias"])
    h = _causal_block(sd, d + "final_block.", h)
    return F.conv1d(h, sd[d + "final_proj.weight"], sd[d + "final_proj.bias"])


# ----------------------------------------------------------------------------- CFM (K8) + flow.inference (K7)
def cfm_t_span(n: int) -> torch.Tensor:
    return 1 - torch.cos(torch.linspace(0, 1, n + 1) * 0.5 * math.pi)


def solve_euler(sd, c: FlowConfig, mu, spks, cond, trace=None):
    """CausalConditionalCFM.forward/solve_euler: fixed noise buffer, cosine schedule,
    CFG-batched estimator (row 1 = zeros for mu/spks/cond), v = (1+r) v_c - r v_u."""
    T = mu.shape[-1]
    x = sd["flow.decoder.rand_noise"][:, :, :T].to(mu)
    ts = cfm_t_span(c.n_timesteps).to(mu)
    t, dt = ts[0], ts[1] - ts[0]
    z = torch.zeros_like
    for step in range(1, len(ts)):
        v = estimator_forward(sd, c, torch.cat([x, x]), torch.cat([mu, z(mu)]), t.expand(2),
                              torch.cat([spks, z(spks)]), torch.cat([cond, z(cond)]))
        v = (1.0 + c.cfg_rate) * v[:1] - c.cfg_rate * v[1:]
        x = x + dt * v
        t = t + dt
        if trace is not None:
            trace.append(x.clone())
        if step < len(ts) - 1:
            dt = ts[step + 1] - t
    return x


def flow_inference(sd, c: FlowConfig, speech_tokens, ref, trace=None):
    """CausalMaskedDiffWithXvec.inference(finalize=True) -> mel (1, 80, 2*n)."""
    emb = F.normalize(ref["embedding"].float(), dim=1)
    spks = _lin(emb, sd, "flow.spk_embed_affine_layer")
    tok = torch.cat([ref["prompt_token"].long().view(1, -1), speech_tokens.long().view(1, -1)], dim=1)
    te = sd["flow.input_embedding.weight"][tok.clamp(min=0)]
    h = encoder_forward(sd, c, te)
    mu = _lin(h, sd, "flow.encoder_proj").transpose(1, 2)
    pf = ref["prompt_feat"].float()
    L1 = pf.shape[1]
    cond = torch.zeros_like(mu)
    cond[:, :, :L1] = pf.transpose(1, 2)
    if trace is not None:
        trace.append(mu.clone())
    feat = solve_euler(sd, c, mu, spks, cond, trace)
    return feat[:, :, L1:]
