"""ORACLE (test infrastructure, not product): CPU/PyTorch fp32 restatement of the voice-conditioning encoders that
`prepare_conditionals` drives (reference src/tts_streaming.py:357-384): 24 kHz prompt mel (`s3gen.embed_ref` :366),
S3Tokenizer-v2 speech tokens (:369-372, and inside embed_ref), CAMPPlus x-vector (inside embed_ref), VoiceEncoder speaker
embedding (:374-375).

PARITY: the arithmetic lives in the un-vendored dependency chatterbox (reference requirements.txt:9; upstream
models/s3gen/{s3gen.py,utils/mel.py,xvector.py}, models/s3tokenizer/s3tokenizer.py (+ the `s3tokenizer` package's
model_v2.py), models/voice_encoder/{voice_encoder.py,melspec.py}) and is restated from its published algorithm.  The signal
front ends ARE pinned, against torchaudio which is in this image (tests/test_oracle_cond.py): the 24 k -> 16 k sinc resampler
against torchaudio.functional.resample, the Kaldi fbank against torchaudio.compliance.kaldi.fbank, the slaney mel filter
bank against torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney") (= librosa.filters.mel), the STFTs
against torch.stft, the LSTM against torch.nn.LSTM.  The network bodies (S3Tokenizer-v2 encoder, CAMPPlus) are UNPINNED
restatements.

Deliberate deviation, stated: the reference loads and resamples the file with librosa (soxr_hq), which is not in this
image and is not bit-reproducible with any torch resampler; file -> 24 kHz here uses the same windowed-sinc
resampler as embed_ref's 24 k -> 16 k step.  `librosa.effects.trim(top_db=20)` inside embeds_from_wavs is restated below.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
import math

import torch
import torch.nn.functional as F

S3GEN_SR, S3_SR = 24000, 16000


# ------------------------------------------------------------------------------------------------ signal front ends
def sinc_resample_kernel(orig, new, lowpass_filter_width=6, rolloff=0.99):
    """torchaudio.functional._get_sinc_resample_kernel (sinc_interp_hann): (new', 1, 2*width + orig') kernel, width; the
    ratio reduced by its gcd."""
    g = math.gcd(int(orig), int(new))
    orig, new = int(orig) // g, int(new) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float64)[:, None, None] / new + idx
    t = (t * base).clamp(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base / orig
    k = torch.where(t == 0, torch.ones_like(t), t.sin() / t) * window * scale
    return k.float(), width, orig, new


def resample(wav, orig, new):
    """wav (L,) -> (ceil(new * L / orig),)  (torchaudio.functional.resample defaults)."""
    if orig == new:
        return wav
    k, width, o, n = sinc_resample_kernel(orig, new)
    L = wav.shape[-1]
    x = F.pad(wav[None, None], (width, width + o))
    y = F.conv1d(x, k, stride=o)                       # (1, n, L / o + 1)
    y = y.transpose(1, 2).reshape(-1)
    return y[: math.ceil(n * L / o)]


def _hz_to_mel_slaney(f):
    f = torch.as_tensor(f, dtype=torch.float64)
    lin = f / (200.0 / 3)
    logstep = math.log(6.4) / 27.0
    return torch.where(f >= 1000.0, 15.0 + torch.log(f.clamp(min=1e-9) / 1000.0) / logstep, lin)


def _mel_to_hz_slaney(m):
    logstep = math.log(6.4) / 27.0
    return torch.where(m >= 15.0, 1000.0 * torch.exp(logstep * (m - 15.0)), m * (200.0 / 3))


def mel_filters(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) (slaney scale, slaney norm): (n_mels, n_fft // 2 + 1)."""
    fmax = sr / 2 if fmax is None else fmax
    freqs = torch.linspace(0, sr / 2, n_fft // 2 + 1, dtype=torch.float64)
    pts = _mel_to_hz_slaney(torch.linspace(float(_hz_to_mel_slaney(fmin)), float(_hz_to_mel_slaney(fmax)), n_mels + 2, dtype=torch.float64))
    fdiff = pts[1:] - pts[:-1]
    ramps = pts[:, None] - freqs[None]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = torch.clamp(torch.minimum(lower, upper), min=0)
    w = w * (2.0 / (pts[2:] - pts[:-2]))[:, None]
    return w.float()


def hann(n):
    """torch.hann_window(n) (periodic)."""
    return (0.5 - 0.5 * torch.cos(2 * math.pi * torch.arange(n, dtype=torch.float64) / n)).float()


def stft_mag(x, n_fft, hop, win, center, pad):
    """|STFT| (bins, frames) of a reflect-padded signal: `pad` samples both sides when center is False (matcha mel), n_fft // 2
    when center is True (torch.stft / librosa default)."""
    p = n_fft // 2 if center else pad
    x = F.pad(x[None, None], (p, p), mode="reflect")[0, 0]
    fr = x.unfold(0, n_fft, hop) * win                       # (frames, n_fft)
    return torch.fft.rfft(fr.double(), dim=-1).abs().float().T


def mel_24k(wav24):
    """upstream s3gen/utils/mel.py mel_spectrogram(n_fft 1920, 80 mels, 24 kHz, hop 480, win 1920, fmin 0, fmax 8000,
    center False): (T, 80) log-mel, the prompt_feat of embed_ref."""
    mag = stft_mag(wav24, 1920, 480, hann(1920), False, (1920 - 480) // 2)
    spec = torch.sqrt(mag.double() ** 2 + 1e-9).float()
    mel = mel_filters(24000, 1920, 80, 0, 8000) @ spec
    return torch.log(torch.clamp(mel, min=1e-5)).T.contiguous()


def log_mel_16k(wav16):
    """upstream S3Tokenizer.log_mel_spectrogram (whisper front end, 128 mels): (128, T) with T = len // 160."""
    mag = stft_mag(wav16, 400, 160, hann(400), True, 0)
    power = mag[:, :-1] ** 2
    mel = mel_filters(16000, 400, 128) @ power
    ls = torch.clamp(mel, min=1e-10).log10()
    ls = torch.maximum(ls, ls.max() - 8.0)
    return (ls + 4.0) / 4.0


def ve_mel(wav16):
    """upstream voice_encoder/melspec.py melspectrogram (40 mels, n_fft 400, hop 160, power 2, "amp", not normalised): (T, 40)."""
    mag = stft_mag(wav16, 400, 160, hann(400), True, 0)
    return (mel_filters(16000, 400, 40, 0, 8000) @ (mag ** 2)).T.contiguous()


def trim_silence(wav, top_db=20.0, frame_length=2048, hop_length=512):
    """librosa.effects.trim: frames whose RMS is within top_db of the loudest frame are non-silent; cut to their span."""
    x = F.pad(wav[None, None], (frame_length // 2, frame_length // 2), mode="constant")[0, 0]     # librosa.feature.rms pads with zeros (center=True, pad_mode="constant")
    fr = x.unfold(0, frame_length, hop_length)
    mse = (fr.double() ** 2).mean(dim=1)
    db = 10.0 * torch.log10(torch.clamp(mse, min=1e-10))      # power_to_db(mse, ref=max, top_db=None): amin 1e-10
    db = db - db.max()
    nz = torch.nonzero(db > -top_db).flatten()
    if nz.numel() == 0:
        return wav[:0]
    start, end = int(nz[0]) * hop_length, min(wav.shape[0], (int(nz[-1]) + 1) * hop_length)
    return wav[start:end]


def kaldi_fbank(wav16, n_mels=80):
    """torchaudio.compliance.kaldi.fbank(num_mel_bins=80, sample 16 kHz, dither 0) defaults: 25 ms povey window, 10 ms shift,
    snip_edges, DC removal, pre-emphasis 0.97, 512-point power spectrum, HTK-mel triangles from 20 Hz to Nyquist, log: (T, 80)."""
    wl, ws, nfft = 400, 160, 512
    n = 1 + (wav16.shape[0] - wl) // ws
    fr = wav16.unfold(0, wl, ws)[:n].clone()
    fr = fr - fr.mean(dim=1, keepdim=True)
    prev = torch.cat([fr[:, :1], fr[:, :-1]], dim=1)          # replicate the first sample
    fr = fr - 0.97 * prev
    win = (0.5 - 0.5 * torch.cos(2 * math.pi * torch.arange(wl, dtype=torch.float64) / (wl - 1))).pow(0.85).float()
    fr = F.pad(fr * win, (0, nfft - wl))
    power = torch.fft.rfft(fr.double(), dim=-1).abs().pow(2).float()          # (n, 257)
    mel = lambda f: 1127.0 * math.log(1.0 + f / 700.0)
    lo, hi = mel(20.0), mel(8000.0)
    delta = (hi - lo) / (n_mels + 1)
    fm = 1127.0 * torch.log(1.0 + (16000.0 / nfft) * torch.arange(nfft // 2, dtype=torch.float64) / 700.0)[None]
    b = torch.arange(n_mels, dtype=torch.float64)[:, None]
    left, center, right = lo + b * delta, lo + (b + 1) * delta, lo + (b + 2) * delta
    bins = torch.clamp(torch.minimum((fm - left) / (center - left), (right - fm) / (right - center)), min=0).float()   # (80, 256)
    bins = F.pad(bins, (0, 1))
    return torch.log(torch.clamp(power @ bins.T, min=torch.finfo(torch.float32).eps))


# ------------------------------------------------------------------------------------------------ S3Tokenizer v2
def _rotary_tables(T, dim=64, theta=10000.0):
    """precompute_freqs_cis of s3tokenizer model_v2.py: angles repeated over both halves of the head dimension."""
    fr = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    ang = torch.outer(torch.arange(T).float(), fr)
    ang = torch.cat([ang, ang], dim=-1)
    return ang.cos(), ang.sin()


def _apply_rotary(x, cos, sin):
    """x (B, T, H, D): x * cos + rotate_half(x) * sin."""
    D = x.shape[-1]
    xr = torch.cat([-x[..., D // 2:], x[..., : D // 2]], dim=-1)
    return x * cos[None, :, None, :] + xr * sin[None, :, None, :]


def s3_tokenize(sd, mel, n_head=20, p="tokenizer."):
    """S3TokenizerV2.quantize for one utterance: mel (128, T) -> (T // 4,) int64 codes in [0, 6561).
    AudioEncoderV2: conv1 (k3 s2 p1) GELU, conv2 (k3 s2 p1) GELU, 6 x [x + FSMN-attention(LN x); x + MLP(LN x)], FSQ codebook
    (project_down to 8 dims, tanh, * 0.999..., round, + 1, base-3 digits)."""
    x = mel[None]
    x = F.gelu(F.conv1d(x, sd[p + "encoder.conv1.weight"], sd[p + "encoder.conv1.bias"], stride=2, padding=1))
    x = F.gelu(F.conv1d(x, sd[p + "encoder.conv2.weight"], sd[p + "encoder.conv2.bias"], stride=2, padding=1))
    x = x.transpose(1, 2)                                   # (1, T', 1280)
    B, T, Dm = x.shape
    hd = Dm // n_head
    cos, sin = _rotary_tables(T, hd)
    i = 0
    while p + f"encoder.blocks.{i}.attn_ln.weight" in sd:
        b = p + f"encoder.blocks.{i}."
        h = F.layer_norm(x, (Dm,), sd[b + "attn_ln.weight"], sd[b + "attn_ln.bias"], 1e-5)
        q = F.linear(h, sd[b + "attn.query.weight"], sd[b + "attn.query.bias"]).view(B, T, n_head, hd)
        k = F.linear(h, sd[b + "attn.key.weight"]).view(B, T, n_head, hd)
        v = F.linear(h, sd[b + "attn.value.weight"], sd[b + "attn.value.bias"]).view(B, T, n_head, hd)
        q, k = _apply_rotary(q, cos, sin), _apply_rotary(k, cos, sin)
        vf = v.reshape(B, T, Dm).transpose(1, 2)
        fsm = F.conv1d(F.pad(vf, (15, 15)), sd[b + "attn.fsmn_block.weight"], groups=Dm).transpose(1, 2) + v.reshape(B, T, Dm)
        sc = hd ** -0.25
        w = torch.softmax((q.permute(0, 2, 1, 3) * sc) @ (k.permute(0, 2, 3, 1) * sc), dim=-1)
        a = (w @ v.permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(B, T, Dm)
        x = x + F.linear(a, sd[b + "attn.out.weight"], sd[b + "attn.out.bias"]) + fsm
        h = F.layer_norm(x, (Dm,), sd[b + "mlp_ln.weight"], sd[b + "mlp_ln.bias"], 1e-5)
        x = x + F.linear(F.gelu(F.linear(h, sd[b + "mlp.0.weight"], sd[b + "mlp.0.bias"])), sd[b + "mlp.2.weight"], sd[b + "mlp.2.bias"])
        i += 1
    h = torch.tanh(F.linear(x, sd[p + "quantizer._codebook.project_down.weight"], sd[p + "quantizer._codebook.project_down.bias"]))
    h = (h * 0.9990000128746033).round() + 1
    powers = 3 ** torch.arange(8)
    return (h[0].long() * powers).sum(dim=-1)


def s3_tokens_from_wav(sd, wav16, max_len=None):
    """S3Tokenizer.forward for one clip: log-mel (cut to 4 frames per requested token), quantize."""
    mel = log_mel_16k(wav16)
    if max_len is not None:
        mel = mel[:, : max_len * 4]
    return s3_tokenize(sd, mel)


# ------------------------------------------------------------------------------------------------ CAMPPlus x-vector
def _bn(sd, p, x, affine=True):
    shape = [1, -1] + [1] * (x.dim() - 2)
    y = (x - sd[p + "running_mean"].view(shape)) / torch.sqrt(sd[p + "running_var"].view(shape) + 1e-5)
    return y * sd[p + "weight"].view(shape) + sd[p + "bias"].view(shape) if affine else y


def _res2d(sd, p, x, stride):
    y = F.relu(_bn(sd, p + "bn1.", F.conv2d(x, sd[p + "conv1.weight"], stride=(stride, 1), padding=1)))
    y = _bn(sd, p + "bn2.", F.conv2d(y, sd[p + "conv2.weight"], padding=1))
    if p + "shortcut.0.weight" in sd:
        x = _bn(sd, p + "shortcut.1.", F.conv2d(x, sd[p + "shortcut.0.weight"], stride=(stride, 1)))
    return F.relu(y + x)


def _seg_pool(x, seg=100):
    s = F.avg_pool1d(x, kernel_size=seg, stride=seg, ceil_mode=True)
    return s.unsqueeze(-1).expand(*s.shape, seg).reshape(*s.shape[:-1], -1)[..., : x.shape[-1]]


CAMPPLUS_BLOCKS = ((12, 3, 1), (24, 3, 2), (16, 3, 2))


def campplus(sd, feat, p="speaker_encoder.", blocks=None):
    """CAMPPlus(feat_dim 80, embedding 192, growth 32, bn_size 4, init 128): feat (T, 80) mean-normalised fbank -> (192,).
    The layer count of each dense block is read off the checkpoint (reduced test configurations), kernel / dilation are
    upstream's (3, 1), (3, 2), (3, 2)."""
    if blocks is None:
        blocks = []
        for bi, (_, k, dil) in enumerate(CAMPPLUS_BLOCKS):
            n = 0
            while p + f"xvector.block{bi + 1}.tdnnd{n + 1}.linear1.weight" in sd:
                n += 1
            blocks.append((n, k, dil))
    x = feat.T[None, None]                                   # (1, 1, 80, T)
    h = p + "head."
    x = F.relu(_bn(sd, h + "bn1.", F.conv2d(x, sd[h + "conv1.weight"], padding=1)))
    for layer in ("layer1", "layer2"):
        for j in range(2):
            x = _res2d(sd, h + f"{layer}.{j}.", x, 2 if j == 0 else 1)
    x = F.relu(_bn(sd, h + "bn2.", F.conv2d(x, sd[h + "conv2.weight"], stride=(2, 1), padding=1)))
    x = x.reshape(1, -1, x.shape[-1])                        # (1, 32 * 10, T)
    xv = p + "xvector."
    x = F.relu(_bn(sd, xv + "tdnn.nonlinear.batchnorm.", F.conv1d(x, sd[xv + "tdnn.linear.weight"], stride=2, padding=2)))
    for bi, (n_layers, k, dil) in enumerate(blocks):
        for li in range(n_layers):
            l = xv + f"block{bi + 1}.tdnnd{li + 1}."
            y = F.conv1d(F.relu(_bn(sd, l + "nonlinear1.batchnorm.", x)), sd[l + "linear1.weight"])
            y = F.relu(_bn(sd, l + "nonlinear2.batchnorm.", y))
            c = l + "cam_layer."
            loc = F.conv1d(y, sd[c + "linear_local.weight"], padding=(k - 1) // 2 * dil, dilation=dil)
            ctx = y.mean(-1, keepdim=True) + _seg_pool(y)
            ctx = F.relu(F.conv1d(ctx, sd[c + "linear1.weight"], sd[c + "linear1.bias"]))
            m = torch.sigmoid(F.conv1d(ctx, sd[c + "linear2.weight"], sd[c + "linear2.bias"]))
            x = torch.cat([x, loc * m], dim=1)
        t = xv + f"transit{bi + 1}."
        x = F.conv1d(F.relu(_bn(sd, t + "nonlinear.batchnorm.", x)), sd[t + "linear.weight"])
    x = F.relu(_bn(sd, xv + "out_nonlinear.batchnorm.", x))
    stats = torch.cat([x.mean(-1), x.std(-1, unbiased=True)], dim=-1)            # (1, 1024)
    e = F.conv1d(stats[..., None], sd[xv + "dense.linear.weight"])
    return _bn(sd, xv + "dense.nonlinear.batchnorm.", e, affine=False)[0, :, 0]


def xvector_from_wav(sd, wav16):
    f = kaldi_fbank(wav16)
    return campplus(sd, f - f.mean(dim=0, keepdim=True))


# ------------------------------------------------------------------------------------------------ VoiceEncoder
def ve_partials(n_frames, rate=1.3, overlap=0.5, win=160, min_coverage=0.8):
    """voice_encoder.get_num_wins / get_frame_step: (frame_step, n_wins, target_len)."""
    step = int(round((S3_SR / rate) / win)) if rate is not None else int(round(win * (1 - overlap)))
    n_wins, rem = divmod(max(n_frames - win + step, 0), step)
    if n_wins == 0 or (rem + (win - step)) / win >= min_coverage:
        n_wins += 1
    return step, n_wins, win + step * (n_wins - 1)


def lstm_forward(sd, x, p="ve.lstm.", layers=3):
    """torch.nn.LSTM(batch_first) restated: x (B, T, I) -> last layer's final hidden state (B, H)."""
    h_last = None
    for l in range(layers):
        wi, wh = sd[p + f"weight_ih_l{l}"], sd[p + f"weight_hh_l{l}"]
        b = sd[p + f"bias_ih_l{l}"] + sd[p + f"bias_hh_l{l}"]
        H = wh.shape[1]
        h = torch.zeros(x.shape[0], H)
        c = torch.zeros(x.shape[0], H)
        xp = x @ wi.T + b
        out = []
        for t in range(x.shape[1]):
            g = xp[:, t] + h @ wh.T
            i, f, gg, o = g.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            out.append(h)
        x = torch.stack(out, dim=1)
        h_last = h
    return h_last


def voice_embed(sd, wav16, trim=True):
    """VoiceEncoder.embeds_from_wavs([wav], 16 kHz)[0]: trim (top_db 20), 40-mel, 160-frame partials at rate 1.3, 3 x LSTM-256,
    proj + ReLU, L2 norm per partial, mean, L2 norm: (256,)."""
    if trim:
        wav16 = trim_silence(wav16)
    mel = ve_mel(wav16)                                      # (T, 40)
    step, n_wins, target = ve_partials(mel.shape[0])
    if target > mel.shape[0]:
        mel = F.pad(mel, (0, 0, 0, target - mel.shape[0]))
    parts = torch.stack([mel[i * step: i * step + 160] for i in range(n_wins)])
    h = lstm_forward(sd, parts)
    e = F.relu(F.linear(h, sd["ve.proj.weight"], sd["ve.proj.bias"]))
    e = e / e.norm(dim=1, keepdim=True)
    m = e.mean(dim=0)
    return m / m.norm()


# ------------------------------------------------------------------------------------------------ prepare_conditionals
def embed_ref(sd, wav24):
    """S3Token2Mel.embed_ref for a 24 kHz clip (<= 10 s): prompt_feat (T, 80), prompt_token (T // 2,), embedding (192,)."""
    feat = mel_24k(wav24)
    wav16 = resample(wav24, S3GEN_SR, S3_SR)
    emb = xvector_from_wav(sd, wav16)
    tok = s3_tokens_from_wav(sd, wav16)
    if feat.shape[0] != 2 * tok.shape[0]:                    # upstream: "mel_len = 2 * stoken_len" is enforced by cutting the tokens
        tok = tok[: feat.shape[0] // 2]
        feat = feat[: 2 * tok.shape[0]]
    return {"prompt_feat": feat, "prompt_token": tok, "embedding": emb}


def prepare_conditionals(sd, wav24, speech_cond_prompt_len=150, exaggeration=0.5, dec_cond_len=10 * S3GEN_SR, enc_cond_len=6 * S3_SR):
    """reference src/tts_streaming.py:357-384 from the 24 kHz waveform on."""
    wav16 = resample(wav24, S3GEN_SR, S3_SR)
    gen = embed_ref(sd, wav24[:dec_cond_len])
    t3_tok = s3_tokens_from_wav(sd, wav16[:enc_cond_len], max_len=speech_cond_prompt_len)
    spk = voice_embed(sd, wav16)
    return {"t3": {"speaker_emb": spk[None], "cond_prompt_speech_tokens": t3_tok[None], "emotion_adv": exaggeration * torch.ones(1, 1, 1)},
            "gen": {"prompt_token": gen["prompt_token"][None], "prompt_token_len": torch.tensor([gen["prompt_token"].shape[0]]),
                    "prompt_feat": gen["prompt_feat"][None], "prompt_feat_len": None, "embedding": gen["embedding"][None]}}
