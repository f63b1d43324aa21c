"""Checkpoint loading: what `ChatterboxTTS.from_local(MODEL_PATH, device)` reads (reference src/tts_streaming.py:252-258) ->
the flat, weight-norm-free state dict of `weights.schema` that `pack.pack_state_dict` uploads.

`from_local` [upstream chatterbox/tts.py] loads `ve.safetensors`, `t3_cfg.safetensors`, `s3gen.safetensors`,
`tokenizer.json` and `conds.pt` from one directory.  This module converts the two model files that feed the hot path:

  * `t3_cfg.safetensors` -- T3's own state dict (keys `tfmr.*`, `text_emb.*`, `speech_emb.*`, `cond_enc.*`, `speech_head.*`,
    `text_pos_emb.*`, `speech_pos_emb.*`; some releases wrap it under a leading "model." / store `{"model": [sd]}`): gets the
    `t3.` prefix.  `text_head.*` (never evaluated at inference) is dropped.
  * `s3gen.safetensors` -- S3Token2Wav's state dict: `flow.*`, `mel2wav.*` are kept (weight-norm folded: both the
    `parametrizations.weight.original0/1` and the legacy `weight_g/weight_v` spellings), `tokenizer.*` / `speaker_encoder.*`
    (conditioning encoders) are returned separately for the voice-conditioning path.  `flow.decoder.rand_noise` is a
    buffer created at module init under `set_all_random_seed(0)` and is not in the file; it is re-drawn the same way.

A directory with neither a merged `cbx_b200.safetensors` nor the upstream files is an ERROR (the reference's from_local
fails too); seeded random weights are used only when the caller passes a state dict or sets CBX_ALLOW_RANDOM_WEIGHTS=1
(BASELINE.json configs: "random-init Chatterbox weights").
"""
import os
import re
import warnings
from collections import OrderedDict

import torch

from .config import ModelConfig
from .weights import schema, random_state_dict

MERGED = "cbx_b200.safetensors"
UPSTREAM_T3, UPSTREAM_S3GEN, UPSTREAM_CONDS, UPSTREAM_VE = "t3_cfg.safetensors", "s3gen.safetensors", "conds.pt", "ve.safetensors"
COND_PREFIXES = ("tokenizer.", "speaker_encoder.", "ve.")


class CheckpointError(RuntimeError):
    pass


def fold_weight_norm(sd):
    """weight = g * v / ||v|| (norm over every dim but 0, torch.nn.utils.weight_norm's default dim=0) for both spellings."""
    out = OrderedDict()
    pairs = {}
    for k, v in sd.items():
        m = re.match(r"(.*)\.parametrizations\.weight\.original([01])$", k)
        if m:
            pairs.setdefault(m.group(1), {})["g" if m.group(2) == "0" else "v"] = v
            continue
        m = re.match(r"(.*)\.weight_([gv])$", k)
        if m:
            pairs.setdefault(m.group(1), {})[m.group(2)] = v
            continue
        out[k] = v
    for base, gv in pairs.items():
        if "g" not in gv or "v" not in gv:
            raise CheckpointError(f"weight-norm pair incomplete for {base}")
        g, v = gv["g"].float(), gv["v"].float()
        norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape([-1] + [1] * (v.dim() - 1))
        out[base + ".weight"] = g * v / norm
    return out


def upstream_rand_noise(cfg: ModelConfig) -> torch.Tensor:
    """CausalConditionalCFM.__init__ [upstream flow_matching.py]: set_all_random_seed(0); torch.randn([1, 80, 50 * 300])."""
    g = torch.Generator().manual_seed(0)
    return torch.randn([1, cfg.flow.mel, cfg.flow.noise_len], generator=g)


def _unwrap(sd):
    if "model" in sd and not torch.is_tensor(sd["model"]):      # {"model": [state_dict]} releases
        m = sd["model"]
        sd = m[0] if isinstance(m, (list, tuple)) else m
    if sd and all(k.startswith("model.") for k in sd):
        sd = {k[len("model."):]: v for k, v in sd.items()}
    return sd


def convert_upstream(t3_sd, s3gen_sd, cfg: ModelConfig = None, strict: bool = True):
    """(t3_cfg state dict, s3gen state dict) -> (hot-path state dict keyed as weights.schema, conditioning-encoder tensors)."""
    cfg = cfg or ModelConfig()
    want = schema(cfg)
    out, extra = OrderedDict(), OrderedDict()
    for k, v in _unwrap(dict(t3_sd)).items():
        if k.startswith("text_head."):
            continue
        out["t3." + k] = v
    for k, v in fold_weight_norm(_unwrap(dict(s3gen_sd))).items():
        if k.startswith(("tokenizer.", "speaker_encoder.")):
            extra[k] = v
        else:
            out[k] = v
    if "flow.decoder.rand_noise" not in out:
        out["flow.decoder.rand_noise"] = upstream_rand_noise(cfg)
    # buffers that ride in some releases' files but are recomputed here (RoPE tables, window functions, masks)
    for k in [k for k in out if k not in want]:
        extra[k] = out.pop(k)
    missing = [k for k in want if k not in out]
    bad = [k for k, (shape, _) in want.items() if k in out and tuple(out[k].shape) != tuple(shape)]
    if strict and (missing or bad):
        raise CheckpointError(f"upstream checkpoint does not cover the hot path: {len(missing)} missing (e.g. {missing[:3]}), "
                              f"{len(bad)} with a wrong shape (e.g. {[(k, tuple(out[k].shape), want[k][0]) for k in bad[:3]]})")
    return OrderedDict((k, out[k].float()) for k in want if k in out), extra


def load_conds(path, device="cpu"):
    """conds.pt [upstream tts.py Conditionals.load]: {"t3": T3Cond fields, "gen": ref_dict} -> the dict voice_put takes."""
    kw = torch.load(path, map_location=device, weights_only=True)
    t3, gen = kw["t3"], kw["gen"]
    t3 = t3 if isinstance(t3, dict) else t3.__dict__
    return {"t3": {k: t3[k] for k in ("speaker_emb", "cond_prompt_speech_tokens", "emotion_adv")}, "gen": dict(gen)}


def load_checkpoint(model_path: str, cfg: ModelConfig = None, seed: int = 0):
    """-> (state dict, conditioning-encoder tensors or None, source: "merged" | "upstream" | "random")."""
    cfg = cfg or ModelConfig()
    from safetensors.torch import load_file
    merged = os.path.join(model_path or "", MERGED)
    if os.path.exists(merged):
        sd = load_file(merged)
        enc = OrderedDict((k, v) for k, v in sd.items() if k.startswith(COND_PREFIXES))
        return OrderedDict((k, v) for k, v in sd.items() if not k.startswith(COND_PREFIXES)), (enc or None), "merged"
    t3p, s3p = os.path.join(model_path or "", UPSTREAM_T3), os.path.join(model_path or "", UPSTREAM_S3GEN)
    if os.path.exists(t3p) and os.path.exists(s3p):
        sd, extra = convert_upstream(load_file(t3p), load_file(s3p), cfg)
        vep = os.path.join(model_path or "", UPSTREAM_VE)
        if os.path.exists(vep):          # VoiceEncoder state dict (lstm.*, proj.*): merged under "ve."
            for k, v in load_file(vep).items():
                extra["ve." + k] = v
        return sd, extra, "upstream"
    if os.environ.get("CBX_ALLOW_RANDOM_WEIGHTS", "0") == "1":
        warnings.warn(f"no checkpoint under {model_path!r}: using seeded RANDOM weights (CBX_ALLOW_RANDOM_WEIGHTS=1) -- the audio is noise")
        return random_state_dict(cfg, seed), random_state_dict(cfg, seed, parts=("cond",)), "random"
    raise CheckpointError(
        f"no checkpoint under MODEL_PATH={model_path!r}: expected {MERGED}, or the upstream {UPSTREAM_T3} + {UPSTREAM_S3GEN} "
        f"(reference: ChatterboxTTS.from_local, src/tts_streaming.py:252-258).  Benchmarks / tests pass a state dict explicitly; "
        f"set CBX_ALLOW_RANDOM_WEIGHTS=1 to run on seeded random weights.")
