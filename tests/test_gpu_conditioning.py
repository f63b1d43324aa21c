"""GPU parity of the voice-conditioning encoders (cbx_b200/conditioning.py over csrc/cond.cu, through the C-ABI) against the
fp32 oracle (oracle/cond.py) on the same seeded weights and the same synthetic clip.  Tolerances: signal front ends and
embeddings <= 2e-4 relative L2 (fp32 both sides, different summation orders); S3Tokenizer ids: every id equal, except where
the pre-rounding FSQ value lies within 1e-3 of a rounding boundary (reported; none expected at these sizes)."""
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _wave(n, sr, seed=0):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / sr
    x = 0.3 * torch.sin(2 * math.pi * 180 * t) + 0.2 * torch.sin(2 * math.pi * 440 * t + 1.0) + 0.1 * torch.sin(2 * math.pi * 2300 * t) + 0.05 * torch.randn(n, generator=g)
    env = torch.clamp(torch.sin(2 * math.pi * 1.5 * t), min=0) ** 0.5
    return (x * env).float()


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-12))


@pytest.fixture(scope="module", params=["tiny", "full"])
def enc(request):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cbx_b200.config import ModelConfig
    from cbx_b200.conditioning import ConditioningEncoders
    from cbx_b200.weights import random_state_dict
    cfg = ModelConfig.tiny() if request.param == "tiny" else ModelConfig()
    sd = random_state_dict(cfg, 3, parts=("cond",))
    return ConditioningEncoders(sd, cfg.cond, device=0), sd, request.param


def test_front_ends_match_the_oracle(enc):
    from oracle import cond as O
    e, sd, _ = enc
    w24 = _wave(24000 * 2 + 311, 24000, seed=1)
    d24 = e._dev_wave(w24)
    d16 = e.resample(d24, 24000, 16000)
    w16 = O.resample(w24, 24000, 16000)
    assert d16.shape[0] == w16.shape[0] and _rel(d16, w16) < 1e-5
    assert _rel(e.resample(d24, 22050, 24000), O.resample(w24, 22050, 24000)) < 1e-5
    assert _rel(e.mel_24k(d24), O.mel_24k(w24)) < 2e-4
    assert _rel(e.log_mel_16k(d16), O.log_mel_16k(w16).T) < 2e-4
    mel, n = e.ve_mel(d16)
    assert _rel(mel[:n], O.ve_mel(w16)) < 2e-4
    f = O.kaldi_fbank(w16)
    assert _rel(e.kaldi_fbank(d16), f - f.mean(dim=0, keepdim=True)) < 2e-4
    x = torch.cat([torch.zeros(6000), w16, torch.zeros(9000)])
    assert e.trim_silence(e._dev_wave(x)).shape[0] == O.trim_silence(x).shape[0]


def test_networks_match_the_oracle(enc):
    from oracle import cond as O
    e, sd, size = enc
    w24 = _wave(24000 * 3 + 97, 24000, seed=2)
    w16 = O.resample(w24, 24000, 16000)
    d16 = e._dev_wave(w16)
    with torch.no_grad():
        ref_tok = O.s3_tokens_from_wav(sd, w16)
        ref_xv = O.xvector_from_wav(sd, w16)
        ref_spk = O.voice_embed(sd, w16)
    tok = e.s3_tokens_from_wav(d16).long().cpu()
    assert tok.shape == ref_tok.shape
    agree = float((tok == ref_tok).float().mean())
    xv, spk = e.xvector_from_wav(d16), e.voice_embed(d16)
    rep = {"size": size, "tokens": int(tok.shape[0]), "token_agreement": agree, "xvector_rel": _rel(xv, ref_xv), "speaker_emb_rel": _rel(spk, ref_spk)}
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/parity_conditioning_{size}.json", "w") as f:
        json.dump(rep, f, indent=1)
    assert agree == 1.0, rep
    assert rep["xvector_rel"] < 2e-4 and rep["speaker_emb_rel"] < 2e-4, rep


def test_prepare_conditionals_end_to_end(enc):
    """reference :357-384 from the decoded clip on: shapes, ranges and values of everything voice_put receives."""
    from oracle import cond as O
    e, sd, size = enc
    w = _wave(24000 * 4, 24000, seed=4)
    got = e.prepare_conditionals(w.numpy(), 24000)
    with torch.no_grad():
        ref = O.prepare_conditionals(sd, w)
    assert torch.equal(got["gen"]["prompt_token"], ref["gen"]["prompt_token"])
    assert torch.equal(got["t3"]["cond_prompt_speech_tokens"], ref["t3"]["cond_prompt_speech_tokens"])
    assert got["gen"]["prompt_feat"].shape == ref["gen"]["prompt_feat"].shape == (1, 200, 80)
    assert _rel(got["gen"]["prompt_feat"], ref["gen"]["prompt_feat"]) < 2e-4
    assert _rel(got["gen"]["embedding"], ref["gen"]["embedding"]) < 2e-4
    assert _rel(got["t3"]["speaker_emb"], ref["t3"]["speaker_emb"]) < 2e-4
    assert int(got["gen"]["prompt_token_len"][0]) == 100


def test_engine_prepare_conditionals_from_a_wav_file(tmp_path):
    """The engine entry point: a WAV file -> cached voice -> a request that speaks with it (reference :357-384, :386-406)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import asyncio
    from scipy.io import wavfile
    from cbx_b200.config import ModelConfig
    from cbx_b200.engine import TextToSpeechEngine, SamplingDefaults
    from cbx_b200.weights import random_state_dict
    cfg = ModelConfig.tiny()
    path = str(tmp_path / "speaker.wav")
    wavfile.write(path, 22050, (_wave(22050 * 3, 22050, seed=7).numpy() * 32767).astype(np.int16))
    eng = TextToSpeechEngine("cuda:0", cfg=cfg, state_dict=random_state_dict(cfg, 0), encoder_state_dict=random_state_dict(cfg, 0, parts=("cond",)),
                             sampling=SamplingDefaults(tokens_per_word=10), native_kwargs=dict(max_s3_tokens=400, n_lanes=2))
    asyncio.run(eng.ainit())
    try:
        eng.prepare_conditionals(path)
        assert "speaker.wav" in eng.voice_cache
    finally:
        eng.shutdown()
