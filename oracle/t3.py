"""ORACLE (test infrastructure, not product): CPU/PyTorch fp32 restatement of the
T3 speech-token decoder.  PARITY UNPINNED: the reference's arithmetic lives in
the un-vendored, unpinned dependency `git+https://github.com/akashdeep000/chatterbox.git`
(reference requirements.txt:9; a fork of resemble-ai/chatterbox, PyPI `chatterbox-tts`);
this file restates that package's published algorithm (models/t3/t3.py,
modules/cond_enc.py, modules/perceiver.py, modules/learned_pos_emb.py,
llama_configs.py `Llama_520M`, inference/t3_hf_backend.py) and anchors on the
reference's own call sites: src/tts_streaming.py:282-292 (warm-up call),
:420-435 (inference_stream arguments), :473-478 (CFG row duplication, SOT/EOT pad).
The trunk is cross-checked against transformers' LlamaModel and the sampler
against transformers' logits warpers in tests/test_oracle_pins.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.
"""
import math

import torch
import torch.nn.functional as F

from cbx_b200.config import T3Config


# ----------------------------------------------------------------------------- RoPE
def llama3_inv_freq(c: T3Config) -> torch.Tensor:
    """transformers `_compute_llama3_parameters` (rope_type="llama3")."""
    inv = 1.0 / (c.rope_theta ** (torch.arange(0, c.head_dim, 2, dtype=torch.float32) / c.head_dim))
    low_wl = c.rope_orig_max_pos / c.rope_low_freq_factor
    high_wl = c.rope_orig_max_pos / c.rope_high_freq_factor
    wl = 2 * math.pi / inv
    scaled = torch.where(wl > low_wl, inv / c.rope_factor, inv)
    smooth = (c.rope_orig_max_pos / wl - c.rope_low_freq_factor) / (c.rope_high_freq_factor - c.rope_low_freq_factor)
    smoothed = (1 - smooth) * scaled / c.rope_factor + smooth * scaled
    mid = ~(wl < high_wl) * ~(wl > low_wl)
    return torch.where(mid, smoothed, scaled)


def rope_cos_sin(c: T3Config, positions: torch.Tensor):
    f = positions.float()[:, None] * llama3_inv_freq(c).to(positions.device)[None, :]
    emb = torch.cat([f, f], dim=-1)
    return emb.cos(), emb.sin()


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def rms_norm(x, w, eps):
    v = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x.float() * torch.rsqrt(v + eps)).to(x.dtype)


# ----------------------------------------------------------------------------- trunk
def llama_forward(sd, c: T3Config, x, kv=None, prefix="t3.tfmr.", collect=None, attn_probe=None):
    """x: (B, S, D) input embeds; kv: list of (k, v) per layer each (B, H, S0, hd) or None.
    Returns (final-normed hidden (B,S,D), new kv).  attn_probe = (layer, list): the attention probabilities of
    batch row 0 at that layer, averaged over the heads (S, past + S), are appended to the list (the forward hook
    of AlignmentStreamAnalyzer)."""
    B, S, D = x.shape
    past = 0 if kv is None else kv[0][0].shape[2]
    pos = torch.arange(past, past + S, device=x.device)
    cos, sin = rope_cos_sin(c, pos)
    new_kv = []
    H, hd = c.n_heads, c.head_dim
    for i in range(c.n_layers):
        p = f"{prefix}layers.{i}."
        h = rms_norm(x, sd[p + "input_layernorm.weight"], c.rms_eps)
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"]).view(B, S, H, hd).transpose(1, 2)
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"]).view(B, S, H, hd).transpose(1, 2)
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"]).view(B, S, H, hd).transpose(1, 2)
        q = q * cos + _rot_half(q) * sin
        k = k * cos + _rot_half(k) * sin
        if kv is not None:
            k = torch.cat([kv[i][0], k], dim=2)
            v = torch.cat([kv[i][1], v], dim=2)
        new_kv.append((k, v))
        att = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        if S > 1:
            mask = torch.ones(S, past + S, dtype=torch.bool, device=x.device).tril(past)
            att = att.masked_fill(~mask, float("-inf"))
        att = att.softmax(-1)
        if attn_probe is not None and attn_probe[0] == i:
            attn_probe[1].append(att[0].mean(0))
        o = (att @ v).transpose(1, 2).reshape(B, S, D)
        x = x + F.linear(o, sd[p + "self_attn.o_proj.weight"])
        h = rms_norm(x, sd[p + "post_attention_layernorm.weight"], c.rms_eps)
        g = F.linear(h, sd[p + "mlp.gate_proj.weight"])
        u = F.linear(h, sd[p + "mlp.up_proj.weight"])
        x = x + F.linear(F.silu(g) * u, sd[p + "mlp.down_proj.weight"])
        if collect is not None:
            collect.append(x)
    return rms_norm(x, sd[prefix + "norm.weight"], c.rms_eps), new_kv


# ----------------------------------------------------------------------------- conditioning prefix (K0)
def _attn_block2(sd, p, x1, x2, heads):
    """perceiver.AttentionBlock2: x1 + proj_out(attn(q(norm x1), k(norm x2), v(norm x2)))."""
    D = x1.shape[-1]
    n1 = F.layer_norm(x1, (D,), sd[p + "norm.weight"], sd[p + "norm.bias"])
    n2 = F.layer_norm(x2, (D,), sd[p + "norm.weight"], sd[p + "norm.bias"])
    q = F.linear(n1, sd[p + "to_q.weight"], sd[p + "to_q.bias"])
    k = F.linear(n2, sd[p + "to_k.weight"], sd[p + "to_k.bias"])
    v = F.linear(n2, sd[p + "to_v.weight"], sd[p + "to_v.bias"])
    B, T1, _ = q.shape
    sp = lambda t: t.view(B, t.shape[1], heads, D // heads).transpose(1, 2)
    a = (sp(q) @ sp(k).transpose(-1, -2)) / math.sqrt(D // heads)
    h = (a.softmax(-1) @ sp(v)).transpose(1, 2).reshape(B, T1, D)
    return x1 + F.linear(h, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])


def cond_prefix(sd, c: T3Config, t3_cond) -> torch.Tensor:
    """T3CondEnc.forward + T3.prepare_conditioning -> (1, 34, D)."""
    spk = F.linear(t3_cond["speaker_emb"].view(-1, c.speaker_embed_size).float(),
                   sd["t3.cond_enc.spkr_enc.weight"], sd["t3.cond_enc.spkr_enc.bias"])[:, None]
    toks = t3_cond["cond_prompt_speech_tokens"].long()
    pe = sd["t3.speech_pos_emb.emb.weight"][: toks.shape[1]]
    prompt = sd["t3.speech_emb.weight"][toks] + pe[None]
    q0 = sd["t3.cond_enc.perceiver.pre_attention_query"].expand(prompt.shape[0], -1, -1)
    pa = "t3.cond_enc.perceiver.attn."
    pre = _attn_block2(sd, pa, q0, prompt, c.perceiver_heads)
    per = _attn_block2(sd, pa, pre, pre, c.perceiver_heads)
    emo = F.linear(t3_cond["emotion_adv"].view(-1, 1, 1).float(), sd["t3.cond_enc.emotion_adv_fc.weight"])
    return torch.cat([spk, per, emo], dim=1)


def prepare_input_embeds(sd, c: T3Config, t3_cond, text_tokens: torch.Tensor, cfg_weight: float):
    """T3.prepare_input_embeds + the BOS handling of T3.inference: returns (B, L, D).
    Row 1 (unconditional) keeps position embeddings but has its token embedding zeroed.
    The speech BOS appears twice when cfg_weight > 0 (upstream quirk)."""
    text_tokens = torch.atleast_2d(text_tokens).long()
    B, L = text_tokens.shape
    cond = cond_prefix(sd, c, t3_cond).expand(B, -1, -1)
    te = sd["t3.text_emb.weight"][text_tokens].clone()
    if cfg_weight > 0.0:
        te[1].zero_()
    te = te + sd["t3.text_pos_emb.emb.weight"][:L][None]
    bos = sd["t3.speech_emb.weight"][c.start_speech_token] + sd["t3.speech_pos_emb.emb.weight"][0]
    bos = bos[None, None].expand(B, 1, -1)
    parts = [cond, te, bos]
    if cfg_weight > 0.0:
        parts.append(bos)
    return torch.cat(parts, dim=1)


def speech_logits(sd, hidden_last):
    return F.linear(hidden_last, sd["t3.speech_head.weight"])


# ----------------------------------------------------------------------------- sampling (K6)
def process_logits(logits2, generated_ids, cfg_weight, temperature, rep_penalty, min_p, top_p):
    """(B,V) fp32 logits -> (V,) filtered logits.  Order fixed by BASELINE.json north_star:
    CFG mix -> repetition penalty -> temperature -> min-p -> top-p (each stage follows the
    transformers processor of the same name)."""
    lg = logits2.float()
    if cfg_weight > 0.0:
        lg = lg[0] + cfg_weight * (lg[0] - lg[1])
    else:
        lg = lg[0]
    lg = lg.clone()
    if rep_penalty != 1.0 and len(generated_ids):
        ids = torch.as_tensor(generated_ids, dtype=torch.long, device=lg.device).unique()
        s = lg[ids]
        lg[ids] = torch.where(s < 0, s * rep_penalty, s / rep_penalty)
    if temperature != 1.0:
        lg = lg / temperature
    if min_p > 0.0:
        probs = lg.softmax(-1)
        remove = probs < min_p * probs.max()
        order = torch.argsort(lg, descending=True)
        keep_first = torch.zeros_like(remove)
        keep_first[order[:1]] = True
        lg = lg.masked_fill(remove & ~keep_first, float("-inf"))
    if top_p < 1.0:
        sl, si = torch.sort(lg, descending=False, stable=True)
        cum = sl.softmax(-1).cumsum(-1)
        rem = cum <= (1.0 - top_p)
        rem[-1] = False
        lg = lg.masked_fill(torch.zeros_like(rem).scatter(0, si, rem), float("-inf"))
    return lg


# ----------------------------------------------------------------------------- alignment-based EOS control (8f.3)
class AlignmentAnalyzer:
    """Restatement of AlignmentStreamAnalyzer.step of the upstream chatterbox package ([U]: the fork the reference
    imports is not on this machine and unpinned, so is the version of this heuristic -- parity unpinned; the engine
    keeps it off by default).  Per frame it appends the head-averaged attention of the newest query over the text
    span (columns above the frame counter zeroed), tracks the text position (argmax of the newest row when it moved
    by -3..+6), and decides: bit 0 = suppress EOS (position short of the last three text tokens), bit 1 = force EOS
    (after completion, a final-column sum >= 10 -- long tail -- or row maxima over the earlier columns summing
    to > 5 -- repetition).  The first frame holds every prefilled query from the first BOS on (two rows with CFG)."""

    def __init__(self, S: int):
        self.S = S
        self.alignment = torch.zeros(0, S)
        self.curr_frame_pos = 0
        self.text_position = 0
        self.started = False
        self.started_at = None
        self.complete = False
        self.completed_at = None

    def step(self, A_chunk: torch.Tensor) -> int:
        A_chunk = A_chunk.detach().float().cpu().clone()
        A_chunk[:, self.curr_frame_pos + 1:] = 0
        self.alignment = torch.cat((self.alignment, A_chunk), dim=0)
        A = self.alignment
        T, S = A.shape
        cur_text_posn = int(A_chunk[-1].argmax())
        discontinuity = not (-4 < cur_text_posn - self.text_position < 7)
        if not discontinuity:
            self.text_position = cur_text_posn
        false_start = (not self.started) and (float(A[-2:, -2:].max()) > 0.1 or float(A[:, :4].max()) < 0.5)
        self.started = not false_start
        if self.started and self.started_at is None:
            self.started_at = T
        self.complete = self.complete or self.text_position >= S - 3
        if self.complete and self.completed_at is None:
            self.completed_at = T
        long_tail = self.complete and float(A[self.completed_at:, -3:].sum(dim=0).max()) >= 10
        rest = A[self.completed_at:, :-5] if self.complete else None
        repetition = self.complete and rest.numel() > 0 and float(rest.max(dim=1).values.sum()) > 5
        ctl = (2 if (long_tail or repetition) else 0) | (1 if cur_text_posn < S - 3 else 0)
        self.curr_frame_pos += 1
        self.cur_text_posn = cur_text_posn
        return ctl

    @staticmethod
    def apply(logits, ctl: int, eos_idx: int):
        """The analyzer's edit of the step logits (any leading shape): force first, then suppress."""
        if ctl & 2:
            logits = -(2 ** 15) * torch.ones_like(logits)
            logits[..., eos_idx] = 2 ** 15
        if ctl & 1:
            logits = logits.clone()
            logits[..., eos_idx] = -(2 ** 15)
        return logits


def sample_from(filtered_logits, exp_noise):
    """torch.multinomial(softmax(l), 1) restated with explicit Exp(1) noise q:
    argmax(p / q) (ties -> lowest index), the formulation torch itself uses."""
    p = filtered_logits.softmax(-1)
    return int(torch.argmax(p / exp_noise))


def inference_stream(sd, c: T3Config, t3_cond, text_tokens, max_new_tokens, temperature=0.8, cfg_weight=0.5,
                     rep_penalty=1.2, min_p=0.05, top_p=0.95, noise_fn=None, return_logits=False, alignment_layer=None,
                     trace=None):
    """Generator of speech token ids (T3.inference_stream of the fork, driven by
    src/tts_streaming.py:483-501).  noise_fn(step) -> (V,) Exp(1) noise.  alignment_layer: run the
    AlignmentAnalyzer on that trunk layer's attention (upstream: 9) and edit the step logits with its decision;
    trace (list) receives (alignment rows of the frame, ctl) per step."""
    x = prepare_input_embeds(sd, c, t3_cond, text_tokens, cfg_weight)
    probe = None
    if alignment_layer is not None:
        L = torch.atleast_2d(text_tokens).shape[1]
        i0 = x.shape[1] - L - (2 if cfg_weight > 0.0 else 1)
        analyzer = AlignmentAnalyzer(L)
        probe = (min(alignment_layer, c.n_layers - 1), [])
    h, kv = llama_forward(sd, c, x, attn_probe=probe)
    generated = [c.start_speech_token]
    for i in range(max_new_tokens):
        logits = speech_logits(sd, h[:, -1])
        if probe is not None:
            att = probe[1].pop()
            rows = att[i0 + L:, i0:i0 + L] if i == 0 else att[:, i0:i0 + L]
            ctl = analyzer.step(rows)
            if trace is not None:
                trace.append((rows.detach().float().cpu(), ctl))
        used = logits if probe is None else AlignmentAnalyzer.apply(logits, ctl, c.stop_speech_token)
        fl =process_logits(used, generated, cfg_weight, temperature, rep_penalty, min_p, top_p)
        tok = sample_from(fl, noise_fn(i))
        generated.append(tok)
        if return_logits:
            yield tok, logits
        else:
            yield tok
        if tok == c.stop_speech_token:
            return
        e = sd["t3.speech_emb.weight"][tok] + sd["t3.speech_pos_emb.emb.weight"][i + 1]
        e = e[None, None].expand(x.shape[0], 1, -1)
        h, kv = llama_forward(sd, c, e, kv, attn_probe=probe)
