"""Host-side tables of the conditioning path (cbx_b200/conditioning.py) against torchaudio, on CPU: what the GPU kernels are
fed must be the reference's filter banks / resampling kernel / window bookkeeping."""
import numpy as np
import pytest
import torch

from cbx_b200 import conditioning as Cn

torchaudio = pytest.importorskip("torchaudio")


def test_slaney_banks_match_torchaudio():
    for sr, n_fft, n_mels, fmax in ((24000, 1920, 80, 8000.0), (16000, 400, 128, 8000.0), (16000, 400, 40, 8000.0)):
        ref = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, fmax, n_mels, sr, norm="slaney", mel_scale="slaney").T.numpy()
        assert np.allclose(Cn.slaney_mel_bank(sr, n_fft, n_mels, 0.0, fmax), ref, atol=1e-6)


def test_kaldi_bank_matches_torchaudio():
    ref, _ = torchaudio.compliance.kaldi.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    got = Cn.kaldi_mel_bank()
    assert got.shape == (80, 257) and np.allclose(got[:, :256], ref.numpy(), atol=5e-5) and not got[:, 256].any()


def test_sinc_kernel_matches_torchaudio():
    for orig, new in ((24000, 16000), (22050, 24000), (48000, 24000)):
        k, width, o, n = Cn.sinc_kernel(orig, new)
        import math
        g = math.gcd(orig, new)
        ref, rw = torchaudio.functional.functional._get_sinc_resample_kernel(orig, new, g)
        assert width == rw and (o, n) == (orig // g, new // g)
        assert np.allclose(k, ref[:, 0].numpy(), atol=1e-5)      # torchaudio builds it in float32, this table in float64


def test_partials_and_wav_loading(tmp_path):
    from oracle import cond as O
    for n in (1, 100, 160, 161, 237, 500, 1000, 1601):
        assert Cn.ve_partials(n) == O.ve_partials(n)
    from scipy.io import wavfile
    x = (np.sin(np.arange(8000) * 0.05) * 20000).astype(np.int16)
    p = str(tmp_path / "a.wav")
    wavfile.write(p, 16000, np.stack([x, x], axis=1))
    w, sr = Cn.load_wav(p)
    assert sr == 16000 and w.shape == (8000,) and w.dtype == np.float32 and abs(float(np.abs(w).max()) - 20000 / 32768) < 1e-4


def test_encoders_refuse_to_run_without_weights_or_gpu():
    if torch.cuda.is_available():
        with pytest.raises(KeyError):
            Cn.ConditioningEncoders({}, device=0)
    else:
        with pytest.raises(RuntimeError):
            Cn.ConditioningEncoders({}, device=0)
