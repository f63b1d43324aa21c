"""ORACLE (test infrastructure, not product): CPU/PyTorch fp32 restatement of the
HiFT vocoder (ConvRNNF0Predictor, SourceModuleHnNSF/SineGen, HiFTGenerator.decode,
iSTFT) and of S3Token2Wav.inference's trim-fade.  PARITY UNPINNED: restates the
published algorithm of the un-vendored dependency chatterbox (reference
requirements.txt:9; upstream models/s3gen/hifigan.py, f0_predictor.py, s3gen.py),
anchored on the reference call sites src/tts_streaming.py:586-590 (arguments),
:694-699 (cache_source threading, `wav[previous_length:]`), and the fixed 960-samples-
per-token output length the engine relies on.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.
"""
import math

import torch
import torch.nn.functional as F

from cbx_b200.config import HiFTConfig
from cbx_b200.weights import hift_source_down_specs


def hann(n, device=None):
    """scipy.signal.get_window('hann', n, fftbins=True)."""
    return 0.5 - 0.5 * torch.cos(2 * math.pi * torch.arange(n, dtype=torch.float32, device=device) / n)


def f0_predict(sd, c: HiFTConfig, mel):
    """(1,80,T) -> (1,T) non-negative f0 in Hz."""
    p = "mel2wav.f0_predictor."
    x = mel
    for l in range(c.f0_layers):
        x = F.elu(F.conv1d(x, sd[p + f"condnet.{2 * l}.weight"], sd[p + f"condnet.{2 * l}.bias"], padding=1))
    return torch.abs(F.linear(x.transpose(1, 2), sd[p + "classifier.weight"], sd[p + "classifier.bias"]).squeeze(-1))


def sine_source(sd, c: HiFTConfig, f0, phase, noise):
    """SineGen + SourceModuleHnNSF with explicit randomness.
    f0: (1,T) Hz at mel rate; phase: (H+1,) initial phases (phase[0] is forced to 0 upstream);
    noise: (H+1, T*up) standard normal.  Returns source (1,1,T*up)."""
    up = c.upsample_total
    f = f0.repeat_interleave(up, dim=1)  # nn.Upsample(scale_factor=up), nearest
    H = c.nb_harmonics + 1
    mult = torch.arange(1, H + 1, dtype=torch.float32, device=f.device)[:, None]
    fmat = f * mult / c.sr                                   # (H, L)
    theta = 2 * math.pi * (torch.cumsum(fmat.double(), dim=-1) % 1).float()
    ph = phase.clone().view(H, 1)
    ph[0] = 0
    sines = c.nsf_alpha * torch.sin(theta + ph)
    uv = (f > c.voiced_threshold).float()
    namp = uv * c.nsf_sigma + (1 - uv) * c.nsf_alpha / 3
    sines = sines * uv + namp * noise
    m = "mel2wav.m_source.l_linear."
    return torch.tanh(F.linear(sines.transpose(0, 1), sd[m + "weight"], sd[m + "bias"])).view(1, 1, -1)


def snake(x, alpha):
    a = alpha[None, :, None]
    return x + (1.0 / (a + 1e-9)) * torch.sin(x * a) ** 2


def _resblock(sd, p, x, k, dils):
    for j, d in enumerate(dils):
        xt = snake(x, sd[p + f"activations1.{j}.alpha"])
        xt = F.conv1d(xt, sd[p + f"convs1.{j}.weight"], sd[p + f"convs1.{j}.bias"], dilation=d, padding=(k * d - d) // 2)
        xt = snake(xt, sd[p + f"activations2.{j}.alpha"])
        xt = F.conv1d(xt, sd[p + f"convs2.{j}.weight"], sd[p + f"convs2.{j}.bias"], padding=(k - 1) // 2)
        x = xt + x
    return x


def decode(sd, c: HiFTConfig, mel, s, trace=None):
    """HiFTGenerator.decode: mel (1,80,T), source s (1,1,T*up) -> wav (1, T*up)."""
    m = "mel2wav."
    win = hann(c.n_fft, mel.device)
    spec = torch.stft(s.squeeze(1), c.n_fft, c.hop, c.n_fft, window=win, return_complex=True)
    s_stft = torch.cat([spec.real, spec.imag], dim=1)
    x = F.conv1d(mel, sd[m + "conv_pre.weight"], sd[m + "conv_pre.bias"], padding=3)
    nk = len(c.resblock_kernels)
    for i, (u, k) in enumerate(zip(c.upsample_rates, c.upsample_kernels)):
        x = F.leaky_relu(x, c.lrelu_slope)
        x = F.conv_transpose1d(x, sd[m + f"ups.{i}.weight"], sd[m + f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        if i == len(c.upsample_rates) - 1:
            x = F.pad(x, (1, 0), mode="reflect")
        st, ks, pd = hift_source_down_specs(c)[i]
        si = F.conv1d(s_stft, sd[m + f"source_downs.{i}.weight"], sd[m + f"source_downs.{i}.bias"], stride=st, padding=pd)
        si = _resblock(sd, m + f"source_resblocks.{i}.", si, c.source_resblock_kernels[i], c.resblock_dilations)
        x = x + si
        xs = None
        for j, k2 in enumerate(c.resblock_kernels):
            r = _resblock(sd, m + f"resblocks.{i * nk + j}.", x, k2, c.resblock_dilations)
            xs = r if xs is None else xs + r
        x = xs / nk
        if trace is not None:
            trace.append(x.clone())
    x = F.leaky_relu(x)
    x = F.conv1d(x, sd[m + "conv_post.weight"], sd[m + "conv_post.bias"], padding=3)
    nb = c.n_fft // 2 + 1
    mag = torch.exp(x[:, :nb]).clip(max=1e2)
    ph = torch.sin(x[:, nb:])
    wav = torch.istft(torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)), c.n_fft, c.hop, c.n_fft, window=win)
    return wav.clamp(-c.audio_limit, c.audio_limit)


def trim_fade(sr=24000, device=None):
    n = sr // 50
    f = torch.zeros(2 * n, device=device)
    f[n:] = (torch.cos(torch.linspace(math.pi, 0, n, device=device)) + 1) / 2
    return f


def hift_inference(sd, c: HiFTConfig, mel, cache_source, phase, noise):
    """HiFTGenerator.inference + S3Token2Wav.inference trim-fade -> (wav (1,L), source (1,1,L))."""
    f0 = f0_predict(sd, c, mel)
    s = sine_source(sd, c, f0, phase, noise)
    if cache_source is not None and cache_source.shape[2] != 0:
        s[:, :, : cache_source.shape[2]] = cache_source
    wav = decode(sd, c, mel, s)
    tf = trim_fade(c.sr, wav.device)
    wav[:, : len(tf)] *= tf
    return wav, s
