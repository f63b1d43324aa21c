"""Upstream checkpoint converter (cbx_b200/checkpoint.py): key mapping, weight-norm folding in both spellings, the
re-drawn CFM noise buffer, key coverage against the engine's schema, and the hard failure on a missing checkpoint
(reference: ChatterboxTTS.from_local, src/tts_streaming.py:252-258)."""
import os

import pytest
import torch

from cbx_b200 import checkpoint as CK
from cbx_b200.weights import random_state_dict, schema


def _to_upstream(sd, style):
    """The hot-path state dict re-spelled the way the upstream files spell it."""
    t3, s3 = {}, {}
    g = torch.Generator().manual_seed(0)
    for k, v in sd.items():
        if k.startswith("t3."):
            t3[k[3:]] = v.clone()
        elif k == "flow.decoder.rand_noise":
            continue                                  # a buffer made at module init, not in the file
        elif k.startswith("mel2wav.") and k.endswith(".weight") and v.dim() == 3 and "source_downs" not in k:
            base = k[: -len(".weight")]
            norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
            scale = torch.rand(v.shape[0], 1, 1, generator=g) + 0.5     # v is any positive multiple of the direction
            if style == "parametrizations":
                s3[base + ".parametrizations.weight.original0"] = norm.clone()
                s3[base + ".parametrizations.weight.original1"] = v * scale
            else:
                s3[base + ".weight_g"] = norm.clone()
                s3[base + ".weight_v"] = v * scale
        else:
            s3[k] = v.clone()
    t3["text_head.weight"] = torch.zeros(704, 1024)
    s3["tokenizer.encoder.conv1.weight"] = torch.zeros(4, 4, 3)
    s3["speaker_encoder.head.weight"] = torch.zeros(4, 4)
    return t3, s3


@pytest.mark.parametrize("style", ["parametrizations", "legacy"])
def test_convert_upstream_round_trip(tiny_cfg, style):
    sd = random_state_dict(tiny_cfg, 0)
    t3, s3 = _to_upstream(sd, style)
    out, extra = CK.convert_upstream(t3, s3, tiny_cfg)
    want = schema(tiny_cfg)
    assert list(out.keys()) == list(want.keys()), "converter must cover exactly the tensors the engine reads"
    for k, v in sd.items():
        if k == "flow.decoder.rand_noise":
            assert out[k].shape == v.shape and torch.equal(out[k], CK.upstream_rand_noise(tiny_cfg))
        else:
            assert torch.allclose(out[k], v, rtol=1e-5, atol=1e-6), k
    assert "tokenizer.encoder.conv1.weight" in extra and "speaker_encoder.head.weight" in extra
    assert not any(k.startswith("t3.text_head") for k in out)


def test_convert_accepts_wrapped_t3(tiny_cfg):
    sd = random_state_dict(tiny_cfg, 0)
    t3, s3 = _to_upstream(sd, "parametrizations")
    out, _ = CK.convert_upstream({"model." + k: v for k, v in t3.items()}, s3, tiny_cfg)
    assert torch.equal(out["t3.speech_emb.weight"], sd["t3.speech_emb.weight"])


def test_convert_reports_missing_and_misshaped(tiny_cfg):
    sd = random_state_dict(tiny_cfg, 0)
    t3, s3 = _to_upstream(sd, "legacy")
    del t3["speech_head.weight"]
    s3["flow.encoder_proj.weight"] = torch.zeros(3, 3)
    with pytest.raises(CK.CheckpointError) as ei:
        CK.convert_upstream(t3, s3, tiny_cfg)
    assert "t3.speech_head.weight" in str(ei.value) and "flow.encoder_proj.weight" in str(ei.value)


def test_packer_accepts_converted_checkpoint(tiny_cfg):
    """converted dict -> packer -> every tensor the engine's manifest asks for (no device needed)."""
    from cbx_b200.pack import pack_state_dict
    t3, s3 = _to_upstream(random_state_dict(tiny_cfg, 0), "parametrizations")
    out, _ = CK.convert_upstream(t3, s3, tiny_cfg)
    packed = pack_state_dict(out, tiny_cfg)
    assert "t3.l0.wqkv_f" in packed and "hift.ups0.w" in packed and packed["cfm.noise"].shape == (15000, 80)


def test_missing_checkpoint_is_an_error(tmp_path, tiny_cfg, monkeypatch):
    monkeypatch.delenv("CBX_ALLOW_RANDOM_WEIGHTS", raising=False)
    with pytest.raises(CK.CheckpointError):
        CK.load_checkpoint(str(tmp_path), tiny_cfg)
    monkeypatch.setenv("CBX_ALLOW_RANDOM_WEIGHTS", "1")
    with pytest.warns(UserWarning):
        sd, extra, src = CK.load_checkpoint(str(tmp_path), tiny_cfg)
    assert src == "random" and "t3.speech_head.weight" in sd


def test_upstream_files_are_found_and_converted(tmp_path, tiny_cfg):
    from safetensors.torch import save_file
    sd = random_state_dict(tiny_cfg, 1)
    t3, s3 = _to_upstream(sd, "parametrizations")
    save_file({k: v.contiguous() for k, v in t3.items()}, os.path.join(tmp_path, CK.UPSTREAM_T3))
    save_file({k: v.contiguous() for k, v in s3.items()}, os.path.join(tmp_path, CK.UPSTREAM_S3GEN))
    out, extra, src = CK.load_checkpoint(str(tmp_path), tiny_cfg)
    assert src == "upstream" and torch.allclose(out["mel2wav.conv_pre.weight"], sd["mel2wav.conv_pre.weight"], atol=1e-6)
    assert extra and any(k.startswith("tokenizer.") for k in extra)
