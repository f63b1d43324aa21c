"""Checkpoint -> device layout.  Pure data movement: fuses / reorders the upstream-named tensors
(see weights.schema) into the packed tensors libcbx_b200 publishes through cbx_tensor_info:
  * linear weights stay [N][K] bf16; conv weights become [C_out][k][C_in] (implicit-GEMM K order);
  * T3 decode weights are additionally stored in mma.m16n8k16 A-fragment order (t3_kernels.cu);
  * transposed convs become one [u*C_out][taps][C_in] matrix covering all u output phases (hift.cu).
"""
import math

import torch

from .config import ModelConfig


def frag_order(w: torch.Tensor) -> torch.Tensor:
    """[N][K] -> [N/16][K/16][32 lanes][8] in the register order of an m16n8k16 A fragment."""
    N, K = w.shape
    assert N % 16 == 0 and K % 16 == 0
    t = w.view(N // 16, 16, K // 16, 16).permute(0, 2, 1, 3)  # [s][kt][r][c]
    lane = torch.arange(32)
    g, tg = lane // 4, lane % 4
    rows = torch.stack([g, g, g + 8, g + 8, g, g, g + 8, g + 8], dim=1)
    cols = torch.stack([2 * tg, 2 * tg + 1, 2 * tg, 2 * tg + 1, 2 * tg + 8, 2 * tg + 9, 2 * tg + 8, 2 * tg + 9], dim=1)
    return t[:, :, rows, cols].contiguous()


def conv_pack(w: torch.Tensor, cin_pad: int = None) -> torch.Tensor:
    """Conv1d weight [C_out][C_in][k] -> [C_out][k*C_in(_pad)]."""
    co, ci, k = w.shape
    t = w.permute(0, 2, 1)
    if cin_pad and cin_pad > ci:
        t = torch.nn.functional.pad(t, (0, cin_pad - ci))
    return t.reshape(co, -1).contiguous()


def convT_pack(w: torch.Tensor, b: torch.Tensor, u: int):
    """ConvTranspose1d weight [C_in][C_out][k], stride u -> ([u*C_out][taps*C_in], bias tiled u times).
    Output phase r of input-rate row q reads x[q-(taps-1)+m'] with kernel tap j = r + (taps-1-m')*u."""
    ci, co, k = w.shape
    taps = (k + u - 1) // u
    out = torch.zeros(u, co, taps, ci)
    for r in range(u):
        for m in range(taps):
            j = r + (taps - 1 - m) * u
            if j < k:
                out[r, :, m, :] = w[:, :, j].t()
    return out.reshape(u * co, taps * ci).contiguous(), b.repeat(u).contiguous()


def llama3_inv_freq(c) -> torch.Tensor:
    inv = 1.0 / (c.rope_theta ** (torch.arange(0, c.head_dim, 2, dtype=torch.float32) / c.head_dim))
    low_wl = c.rope_orig_max_pos / c.rope_low_freq_factor
    high_wl = c.rope_orig_max_pos / c.rope_high_freq_factor
    wl = 2 * math.pi / inv
    scaled = torch.where(wl > low_wl, inv / c.rope_factor, inv)
    smooth = (c.rope_orig_max_pos / wl - c.rope_low_freq_factor) / (c.rope_high_freq_factor - c.rope_low_freq_factor)
    smoothed = (1 - smooth) * scaled / c.rope_factor + smooth * scaled
    mid = ~(wl < high_wl) * ~(wl > low_wl)
    return torch.where(mid, smoothed, scaled)


def pack_state_dict(sd, cfg: ModelConfig):
    P = {}
    bf = lambda t: t.to(torch.bfloat16).contiguous()
    f32 = lambda t: t.float().contiguous()

    def lin(dst, src, bias=True):
        P[dst + ".w"] = bf(sd[src + ".weight"])
        if bias:
            P[dst + ".b"] = f32(sd[src + ".bias"])

    def ln(dst, src):
        P[dst + ".g"] = f32(sd[src + ".weight"])
        P[dst + ".b"] = f32(sd[src + ".bias"])

    def conv(dst, src, cin_pad=None):
        P[dst + ".w"] = bf(conv_pack(sd[src + ".weight"], cin_pad))
        P[dst + ".b"] = f32(sd[src + ".bias"])

    # ------------------------------------------------------------------ T3
    t3 = cfg.t3
    P["t3.text_emb"] = f32(sd["t3.text_emb.weight"])
    P["t3.speech_emb"] = f32(sd["t3.speech_emb.weight"])
    P["t3.text_pos"] = f32(sd["t3.text_pos_emb.emb.weight"])
    P["t3.speech_pos"] = f32(sd["t3.speech_pos_emb.emb.weight"])
    P["t3.final_norm"] = f32(sd["t3.tfmr.norm.weight"])
    P["t3.inv_freq"] = llama3_inv_freq(t3)
    head = sd["t3.speech_head.weight"]
    vpad = (head.shape[0] + 15) // 16 * 16
    P["t3.head_f"] = frag_order(bf(torch.nn.functional.pad(head, (0, 0, 0, vpad - head.shape[0]))))
    lin("t3.spkr", "t3.cond_enc.spkr_enc")
    P["t3.emo_w"] = f32(sd["t3.cond_enc.emotion_adv_fc.weight"][:, 0])
    P["t3.perc_query"] = f32(sd["t3.cond_enc.perceiver.pre_attention_query"][0])
    pa = "t3.cond_enc.perceiver.attn."
    ln("t3.pnorm", pa + "norm")
    lin("t3.pq", pa + "to_q")
    P["t3.pkv.w"] = bf(torch.cat([sd[pa + "to_k.weight"], sd[pa + "to_v.weight"]]))
    P["t3.pkv.b"] = f32(torch.cat([sd[pa + "to_k.bias"], sd[pa + "to_v.bias"]]))
    lin("t3.po", pa + "proj_out")
    for i in range(t3.n_layers):
        s, d = f"t3.tfmr.layers.{i}.", f"t3.l{i}."
        P[d + "ln1"] = f32(sd[s + "input_layernorm.weight"])
        P[d + "ln2"] = f32(sd[s + "post_attention_layernorm.weight"])
        wqkv = bf(torch.cat([sd[s + f"self_attn.{n}_proj.weight"] for n in "qkv"]))
        wo = bf(sd[s + "self_attn.o_proj.weight"])
        g, u = bf(sd[s + "mlp.gate_proj.weight"]), bf(sd[s + "mlp.up_proj.weight"])
        wd = bf(sd[s + "mlp.down_proj.weight"])
        P[d + "wqkv"], P[d + "wo"], P[d + "wd"] = wqkv, wo, wd
        P[d + "wgu"] = torch.stack([g, u], dim=1).reshape(2 * g.shape[0], -1).contiguous()           # rows (gate_j, up_j)
        P[d + "wqkv_f"], P[d + "wo_f"], P[d + "wd_f"] = frag_order(wqkv), frag_order(wo), frag_order(wd)
        gu16 = torch.stack([g.view(-1, 16, g.shape[1]), u.view(-1, 16, u.shape[1])], dim=1)            # strips (gate_j, up_j)
        P[d + "wgu_f"] = frag_order(gu16.reshape(2 * g.shape[0], -1))
    # ------------------------------------------------------------------ flow encoder
    fc = cfg.flow
    P["flow.tok_emb"] = f32(sd["flow.input_embedding.weight"])
    P["flow.spk.w"] = f32(sd["flow.spk_embed_affine_layer.weight"])
    P["flow.spk.b"] = f32(sd["flow.spk_embed_affine_layer.bias"])
    lin("flow.enc_proj", "flow.encoder_proj")
    e = "flow.encoder."
    for dst, src in (("embed", "embed"), ("up_embed", "up_embed")):
        lin("flow." + dst, e + src + ".out.0")
        ln("flow." + dst + "_ln", e + src + ".out.1")
    conv("flow.pl1", e + "pre_lookahead_layer.conv1")
    conv("flow.pl2", e + "pre_lookahead_layer.conv2")
    conv("flow.upconv", e + "up_layer.conv")
    ln("flow.after_norm", e + "after_norm")
    for grp, dstp, n in (("encoders", "flow.enc", fc.enc_blocks), ("up_encoders", "flow.up", fc.up_blocks)):
        for i in range(n):
            s, d = e + f"{grp}.{i}.", f"{dstp}{i}."
            ln(d + "nm", s + "norm_mha")
            ln(d + "nf", s + "norm_ff")
            a = s + "self_attn."
            wq, bq = sd[a + "linear_q.weight"], sd[a + "linear_q.bias"]
            P[d + "qkv4.w"] = bf(torch.cat([wq, wq, sd[a + "linear_k.weight"], sd[a + "linear_v.weight"]]))
            P[d + "qkv4.b"] = f32(torch.cat([bq + sd[a + "pos_bias_u"].flatten(), bq + sd[a + "pos_bias_v"].flatten(),
                                            sd[a + "linear_k.bias"], sd[a + "linear_v.bias"]]))
            lin(d + "pos", a + "linear_pos", bias=False)
            lin(d + "out", a + "linear_out")
            lin(d + "w1", s + "feed_forward.w_1")
            lin(d + "w2", s + "feed_forward.w_2")
    # ------------------------------------------------------------------ CFM estimator
    c = "flow.decoder.estimator."
    lin("cfm.t1", c + "time_mlp.linear_1")
    lin("cfm.t2", c + "time_mlp.linear_2")
    stages = [c + "down_blocks.0."] + [c + f"mid_blocks.{i}." for i in range(fc.n_mid)] + [c + "up_blocks.0."]
    for r, sp in enumerate(stages):
        d = f"cfm.r{r}."
        lin(d + "tmlp", sp + "0.mlp.1")
        conv(d + "c1", sp + "0.block1.block.0")
        ln(d + "n1", sp + "0.block1.block.2")
        conv(d + "c2", sp + "0.block2.block.0")
        ln(d + "n2", sp + "0.block2.block.2")
        conv(d + "res", sp + "0.res_conv")
        for j in range(fc.n_blocks):
            s, t = sp + f"1.{j}.", d + f"t{j}."
            ln(t + "n1", s + "norm1")
            P[t + "qkv.w"] = bf(torch.cat([sd[s + f"attn1.{n}.weight"] for n in ("to_q", "to_k", "to_v")]))
            lin(t + "out", s + "attn1.to_out.0")
            ln(t + "n3", s + "norm3")
            lin(t + "ff0", s + "ff.net.0.proj")
            lin(t + "ff2", s + "ff.net.2")
    conv("cfm.down_conv", c + "down_blocks.0.2")
    conv("cfm.up_conv", c + "up_blocks.0.2")
    conv("cfm.final_conv", c + "final_block.block.0")
    ln("cfm.final_ln", c + "final_block.block.2")
    conv("cfm.final_proj", c + "final_proj")
    P["cfm.noise"] = f32(sd["flow.decoder.rand_noise"][0].t())
    # ------------------------------------------------------------------ HiFT
    hc = cfg.hift
    m = "mel2wav."
    conv("hift.conv_pre", m + "conv_pre", cin_pad=128)       # mel channels 80 -> 128: whole K tiles for the tcgen05 conv
    conv("hift.conv_post", m + "conv_post")
    nk = len(hc.resblock_kernels)

    def resblock(dst, src):
        for j in range(3):
            conv(dst + f"c1_{j}", src + f"convs1.{j}")
            conv(dst + f"c2_{j}", src + f"convs2.{j}")
            P[dst + f"a1_{j}"] = f32(sd[src + f"activations1.{j}.alpha"])
            P[dst + f"a2_{j}"] = f32(sd[src + f"activations2.{j}.alpha"])

    for i, u in enumerate(hc.upsample_rates):
        w, b = convT_pack(sd[m + f"ups.{i}.weight"], sd[m + f"ups.{i}.bias"], u)
        P[f"hift.ups{i}.w"], P[f"hift.ups{i}.b"] = bf(w), f32(b)
        conv(f"hift.sdown{i}", m + f"source_downs.{i}", cin_pad=64)
        resblock(f"hift.sres{i}.", m + f"source_resblocks.{i}.")
        for k in range(nk):
            resblock(f"hift.res{i * nk + k}.", m + f"resblocks.{i * nk + k}.")
    for l in range(hc.f0_layers):
        conv(f"hift.f0c{l}", m + f"f0_predictor.condnet.{2 * l}", cin_pad=128 if l == 0 else None)
    P["hift.f0w"] = f32(sd[m + "f0_predictor.classifier.weight"][0])
    P["hift.f0b"] = f32(sd[m + "f0_predictor.classifier.bias"])
    P["hift.lw"] = f32(sd[m + "m_source.l_linear.weight"][0])
    P["hift.lb"] = f32(sd[m + "m_source.l_linear.bias"])
    n = hc.sr // 50
    fade = torch.zeros(2 * n)
    fade[n:] = (torch.cos(torch.linspace(math.pi, 0, n)) + 1) / 2
    P["hift.fade"] = fade
    return P
