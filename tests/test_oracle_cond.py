"""Pins of the conditioning oracle's signal front ends against torchaudio / torch (CPU), and shape / invariant checks of the
network bodies (oracle/cond.py; reference call sites src/tts_streaming.py:357-384)."""
import math

import pytest
import torch

from cbx_b200.config import ModelConfig
from cbx_b200.weights import random_state_dict
from oracle import cond as O

torchaudio = pytest.importorskip("torchaudio")


def _wave(n, sr, seed=0):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / sr
    x = 0.3 * torch.sin(2 * math.pi * 180 * t) + 0.2 * torch.sin(2 * math.pi * 440 * t + 1.0) + 0.05 * torch.randn(n, generator=g)
    env = torch.clamp(torch.sin(2 * math.pi * 1.5 * t), min=0) ** 0.5
    return (x * env).float()


def test_resampler_matches_torchaudio():
    x = _wave(24000 * 2 + 137, 24000)
    ref = torchaudio.functional.resample(x, 24000, 16000)
    got = O.resample(x, 24000, 16000)
    assert got.shape == ref.shape and torch.allclose(got, ref, atol=1e-6)
    ref = torchaudio.functional.resample(x, 22050, 24000)
    got = O.resample(x, 22050, 24000)
    assert got.shape == ref.shape and torch.allclose(got, ref, atol=1e-5)


def test_mel_filters_match_torchaudio_slaney():
    for sr, n_fft, n_mels, fmax in ((24000, 1920, 80, 8000.0), (16000, 400, 128, 8000.0), (16000, 400, 40, 8000.0)):
        ref = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, fmax, n_mels, sr, norm="slaney", mel_scale="slaney").T
        got = O.mel_filters(sr, n_fft, n_mels, 0.0, fmax)
        assert torch.allclose(got, ref, atol=1e-6), (sr, n_mels)


def test_stfts_match_torch_stft():
    x = _wave(16000, 16000)
    ref = torch.stft(x, 400, 160, window=torch.hann_window(400), return_complex=True).abs()
    assert torch.allclose(O.stft_mag(x, 400, 160, O.hann(400), True, 0), ref, atol=2e-4)
    x = _wave(24000, 24000)
    xp = torch.nn.functional.pad(x[None, None], (720, 720), mode="reflect")[0, 0]
    ref = torch.stft(xp, 1920, 480, win_length=1920, window=torch.hann_window(1920), center=False, return_complex=True).abs()
    assert torch.allclose(O.stft_mag(x, 1920, 480, O.hann(1920), False, 720), ref, atol=2e-3)
    assert O.mel_24k(x).shape == (24000 // 480, 80)          # two mel frames per 25 Hz token
    assert O.log_mel_16k(_wave(16000, 16000)).shape == (128, 100)


def test_kaldi_fbank_matches_torchaudio():
    x = _wave(16000 * 2, 16000, seed=3)
    ref = torchaudio.compliance.kaldi.fbank(x[None], num_mel_bins=80, dither=0.0, sample_frequency=16000)
    got = O.kaldi_fbank(x)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=2e-3, rtol=1e-4)


def test_lstm_matches_torch_lstm():
    g = torch.Generator().manual_seed(1)
    lstm = torch.nn.LSTM(40, 256, num_layers=3, batch_first=True)
    sd = {"ve.lstm." + k: v.detach() for k, v in lstm.state_dict().items()}
    x = torch.randn(3, 50, 40, generator=g)
    with torch.no_grad():
        _, (h, _) = lstm(x)
    assert torch.allclose(O.lstm_forward(sd, x), h[-1], atol=1e-5)


def test_partial_windows_follow_upstream():
    assert O.ve_partials(160) == (77, 1, 160)
    assert O.ve_partials(1000)[0] == 77
    step, n, target = O.ve_partials(1000)
    assert target >= 1000 - step and (target - 160) % step == 0 and n == (target - 160) // step + 1


def test_trim_silence_cuts_quiet_edges():
    x = torch.cat([torch.zeros(8000), _wave(16000, 16000, seed=2), torch.zeros(8000)])
    y = O.trim_silence(x)
    assert 12000 < y.shape[0] < 22000


def test_prepare_conditionals_shapes_tiny():
    cfg = ModelConfig.tiny()
    sd = random_state_dict(cfg, 0, parts=("cond",))
    wav = _wave(24000 * 3, 24000, seed=5)
    c = O.prepare_conditionals(sd, wav)
    n = c["gen"]["prompt_token"].shape[1]
    assert n == 75 and c["gen"]["prompt_feat"].shape == (1, 150, 80) and c["gen"]["embedding"].shape == (1, 192)
    assert c["t3"]["cond_prompt_speech_tokens"].shape == (1, 75) and c["t3"]["speaker_emb"].shape == (1, 256)
    assert int(c["gen"]["prompt_token"].min()) >= 0 and int(c["gen"]["prompt_token"].max()) < 6561
    assert abs(float(c["t3"]["speaker_emb"].norm()) - 1.0) < 1e-5
    assert torch.isfinite(c["gen"]["embedding"]).all() and torch.isfinite(c["gen"]["prompt_feat"]).all()
