"""Voice-conditioning encoders on the GPU (reference src/tts_streaming.py:357-384 `prepare_conditionals`: `s3gen.embed_ref`
:366, `s3gen.tokenizer.forward` :370-372, `ve.embeds_from_wavs` :374-375; upstream S3Token2Mel.embed_ref, S3TokenizerV2,
CAMPPlus, VoiceEncoder).

Host logic only: every computation is a kernel of libcbx_b200.so (csrc/cond.cu, C-ABI `cbx_cond_*`); torch provides device
buffers and the stream.  Constant tables (windows, mel filter banks, the resampler's sinc kernel) are built once on the host
in float64 and uploaded.  There is no CPU fallback: without a CUDA device the constructor raises.

Deviation from the reference, stated: the reference reads and resamples the file with librosa (soxr_hq), which is neither
in this image nor reproducible with any torch resampler; file -> 24 kHz here uses the same windowed-sinc
polyphase resampler as embed_ref's own 24 k -> 16 k step.
"""
import ctypes as C
import math
import threading

import numpy as np
import torch

from . import lib as L
from .config import CondConfig, S3_SR, S3GEN_SR

DEC_COND_LEN = 10 * S3GEN_SR      # reference :167
ENC_COND_LEN = 6 * S3_SR          # reference :166


# ---------------------------------------------------------------------------------------------- constant tables (host, float64)
def _mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-9) / 1000.0) / (math.log(6.4) / 27.0), f / (200.0 / 3))


def _hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((math.log(6.4) / 27.0) * (m - 15.0)), m * (200.0 / 3))


def slaney_mel_bank(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    """librosa.filters.mel defaults (slaney scale and area normalisation): [n_mels][n_fft/2 + 1]."""
    fmax = sr / 2 if fmax is None else fmax
    freqs = np.linspace(0, sr / 2, n_fft // 2 + 1)
    pts = _hz_slaney(np.linspace(_mel_slaney(fmin), _mel_slaney(fmax), n_mels + 2))
    lower = (freqs[None] - pts[:-2, None]) / (pts[1:-1] - pts[:-2])[:, None]
    upper = (pts[2:, None] - freqs[None]) / (pts[2:] - pts[1:-1])[:, None]
    w = np.maximum(0.0, np.minimum(lower, upper)) * (2.0 / (pts[2:] - pts[:-2]))[:, None]
    return w.astype(np.float32)


def kaldi_mel_bank(n_mels=80, nfft=512, sr=16000, low=20.0):
    """Kaldi's HTK-mel triangles (torchaudio.compliance.kaldi.get_mel_banks), padded with a zero Nyquist column: [n_mels][nfft/2 + 1]."""
    mel = lambda f: 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)
    lo, hi = mel(low), mel(sr / 2)
    delta = (hi - lo) / (n_mels + 1)
    fm = mel((sr / nfft) * np.arange(nfft // 2))[None]
    b = np.arange(n_mels)[:, None]
    left, center, right = lo + b * delta, lo + (b + 1) * delta, lo + (b + 2) * delta
    bins = np.maximum(0.0, np.minimum((fm - left) / (center - left), (right - fm) / (right - center)))
    return np.pad(bins, ((0, 0), (0, 1))).astype(np.float32)


def sinc_kernel(orig, new, width_taps=6, rolloff=0.99):
    """torchaudio's sinc_interp_hann resampling kernel for the gcd-reduced ratio: (kernel [new'][klen], width, orig', new')."""
    g = math.gcd(int(orig), int(new))
    orig, new = int(orig) // g, int(new) // g
    base = min(orig, new) * rolloff
    width = math.ceil(width_taps * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None] / orig
    t = (np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx) * base
    t = np.clip(t, -width_taps, width_taps)
    window = np.cos(t * math.pi / width_taps / 2) ** 2
    t = t * math.pi
    k = np.where(t == 0, 1.0, np.sin(t) / np.where(t == 0, 1.0, t)) * window * (base / orig)
    return k.astype(np.float32), width, orig, new


def ve_partials(n_frames, rate=1.3, win=160, min_coverage=0.8):
    """VoiceEncoder partial-utterance windows (upstream get_frame_step / get_num_wins): (step, n_wins, target_len)."""
    step = int(round((S3_SR / rate) / win))
    n_wins, rem = divmod(max(n_frames - win + step, 0), step)
    if n_wins == 0 or (rem + (win - step)) / win >= min_coverage:
        n_wins += 1
    return step, n_wins, win + step * (n_wins - 1)


class ConditioningEncoders:
    """Device-resident weights of the three conditioning networks + the op sequences that run them."""

    def __init__(self, state_dict, cfg: CondConfig = None, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError("ConditioningEncoders needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg or CondConfig()
        self.lib = L.load()
        self.dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        need = [k for k in ("tokenizer.encoder.conv1.weight", "speaker_encoder.head.conv1.weight", "ve.lstm.weight_ih_l0") if k not in state_dict]
        if need:
            raise KeyError(f"checkpoint has no conditioning-encoder weights ({need[0]} ...): voices cannot be computed from audio")
        up = lambda t: torch.as_tensor(t, dtype=torch.float32).contiguous().to(self.dev)
        sd = state_dict
        self.w = {}
        for k, v in sd.items():
            if not (k.startswith("tokenizer.") or k.startswith("speaker_encoder.") or k.startswith("ve.")):
                continue
            v = v.float()
            if v.dim() == 3 and "fsmn_block" not in k:
                v = v.permute(0, 2, 1)                    # conv1d [C_out][C_in][k] -> [C_out][k][C_in]: one GEMM row per output channel
            self.w[k] = up(v.reshape(v.shape[0], -1) if v.dim() == 3 else v)
        # BatchNorm (inference) folded to per-channel scale / shift
        self.bn = {}
        for k in list(sd.keys()):
            if k.endswith("running_var") and (k.startswith("speaker_encoder.")):
                p = k[: -len("running_var")]
                inv = 1.0 / torch.sqrt(sd[k].double() + 1e-5)
                g = sd[p + "weight"].double() if p + "weight" in sd else torch.ones_like(inv)
                b = sd[p + "bias"].double() if p + "bias" in sd else torch.zeros_like(inv)
                self.bn[p] = (up((g * inv).float()), up((b - sd[p + "running_mean"].double() * g * inv).float()))
        # LSTM: transposed recurrent weights, summed biases
        self.lstm = []
        for l in range(self.cfg.ve_layers):
            self.lstm.append((up(sd[f"ve.lstm.weight_ih_l{l}"]), up(sd[f"ve.lstm.weight_hh_l{l}"].t().contiguous()),
                              up(sd[f"ve.lstm.bias_ih_l{l}"] + sd[f"ve.lstm.bias_hh_l{l}"])))
        hann = lambda n: 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n, dtype=np.float64) / n)
        self.win1920, self.win400 = up(hann(1920)), up(hann(400))
        self.povey = up((0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400, dtype=np.float64) / 399)) ** 0.85)
        self.mel80_24k = up(slaney_mel_bank(24000, 1920, 80, 0.0, 8000.0))
        self.mel128 = up(slaney_mel_bank(16000, 400, 128))
        self.mel40 = up(slaney_mel_bank(16000, 400, 40, 0.0, 8000.0))
        self.kaldi80 = up(kaldi_mel_bank())
        self.eye80 = up(np.eye(80, dtype=np.float32))
        self._kern = {}
        # one long-lived stream for every conditioning run: a request thread's fresh stream misses the caching allocator's
        # per-stream pools, and each of the ~200 temporaries of a run would fall through to cudaMalloc
        self.stream = torch.cuda.Stream(device=self.dev)
        self.lock = threading.Lock()
        self.n_blocks = 0
        while f"tokenizer.encoder.blocks.{self.n_blocks}.attn_ln.weight" in self.w:
            self.n_blocks += 1
        self.xv_blocks = []
        for bi, (_n, k, dil) in enumerate(((12, 3, 1), (24, 3, 2), (16, 3, 2))):     # kernel / dilation are upstream's; depth from the checkpoint
            n = 0
            while f"speaker_encoder.xvector.block{bi + 1}.tdnnd{n + 1}.linear1.weight" in self.w:
                n += 1
            self.xv_blocks.append((n, k, dil))

    # ------------------------------------------------------------------------------------------ plumbing
    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def _new(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, device=self.dev, dtype=dtype)

    def _gemm(self, A, W, out, M, N, K, lda=None, ldw=None, ldc=None, **kw):
        a = L.SgemmArgs()
        a.A, a.W, a.C = A.data_ptr() + 4 * kw.pop("a_off", 0), W.data_ptr() + 4 * kw.pop("w_off", 0), out.data_ptr() + 4 * kw.pop("c_off", 0)
        a.M, a.N, a.K, a.batch = M, N, K, kw.pop("batch", 1)
        a.kc = kw.pop("kc", K)
        a.lda = lda if lda is not None else a.kc
        a.ldw = ldw if ldw is not None else K
        a.ldc = ldc if ldc is not None else N
        a.a_stride, a.a_dil, a.a_pad, a.a_rows = kw.pop("a_stride", 1), kw.pop("a_dil", 1), kw.pop("a_pad", 0), kw.pop("a_rows", 0)
        a.a_bs, a.w_bs, a.c_bs = kw.pop("a_bs", 0), kw.pop("w_bs", 0), kw.pop("c_bs", 0)
        a.w_trans, a.act, a.alpha, a.a_relu = kw.pop("w_trans", 0), kw.pop("act", 0), kw.pop("alpha", 1.0), kw.pop("a_relu", 0)
        for name in ("a_scale", "a_shift", "bias", "o_scale", "o_shift", "mul", "res", "res2"):
            t = kw.pop(name, None)
            setattr(a, name, t.data_ptr() if t is not None else None)
        a.ldm, a.mul_bs, a.mul_div = kw.pop("ldm", 0), kw.pop("mul_bs", 0), kw.pop("mul_div", 1)
        a.ldr, a.res_bs = kw.pop("ldr", a.ldc), kw.pop("res_bs", 0)
        assert not kw, f"unknown sgemm arguments {list(kw)}"
        L.check(self.lib.cbx_cond_sgemm(C.byref(a), self._st()))
        return out

    def _dev_wave(self, wav):
        t = torch.as_tensor(np.asarray(wav.detach().cpu().numpy() if torch.is_tensor(wav) else wav, dtype=np.float32).reshape(-1))
        return t.to(self.dev)

    # ------------------------------------------------------------------------------------------ signal front ends
    def resample(self, x, orig, new):
        """x device (L,) -> device (ceil(new L / orig),) (torchaudio.functional.resample)."""
        if int(orig) == int(new):
            return x
        key = (int(orig), int(new))
        if key not in self._kern:
            k, width, o, n = sinc_kernel(orig, new)
            self._kern[key] = (torch.from_numpy(k).to(self.dev), width, o, n)
        k, width, o, n = self._kern[key]
        n_out = math.ceil(n * x.shape[0] / o)
        y = self._new(n_out)
        L.check(self.lib.cbx_cond_resample(x.data_ptr(), x.shape[0], y.data_ptr(), n_out, o, n, k.data_ptr(), k.shape[1], width, self._st()))
        return y

    def _spec(self, x, n_fft, hop, pad, frame_len, window, n_frames, mode, remove_dc=0, preemph=0.0):
        out = self._new(n_frames, n_fft // 2 + 1)
        L.check(self.lib.cbx_cond_frames_dft(x.data_ptr(), x.shape[0], n_fft, hop, pad, frame_len, window.data_ptr(), remove_dc, preemph, mode,
                                             n_frames, out.data_ptr(), n_fft // 2 + 1, self._st()))
        return out

    def mel_24k(self, wav24):
        """(T, 80) log-mel of embed_ref (upstream s3gen/utils/mel.py: n_fft 1920, hop 480, center False, reflect pad 720)."""
        n = (wav24.shape[0] + 2 * 720 - 1920) // 480 + 1
        spec = self._spec(wav24, 1920, 480, 720, 1920, self.win1920, n, 2)
        mel = self._gemm(spec, self.mel80_24k, self._new(n, 80), n, 80, 961)
        L.check(self.lib.cbx_cond_mel_log(mel.data_ptr(), mel.numel(), 0, 1e-5, None, self._st()))
        return mel

    def log_mel_16k(self, wav16):
        """(T, 128) whisper log-mel of S3Tokenizer (time-major; the last STFT frame is dropped)."""
        n = wav16.shape[0] // 160
        power = self._spec(wav16, 400, 160, 200, 400, self.win400, n, 1)
        mel = self._gemm(power, self.mel128, self._new(n, 128), n, 128, 201)
        L.check(self.lib.cbx_cond_mel_log(mel.data_ptr(), mel.numel(), 1, 0.0, self._new(1).data_ptr(), self._st()))
        return mel

    def ve_mel(self, wav16, pad_to=0):
        n = wav16.shape[0] // 160 + 1
        power = self._spec(wav16, 400, 160, 200, 400, self.win400, n, 1)
        out = torch.zeros(max(n, pad_to), 40, device=self.dev)
        self._gemm(power, self.mel40, out, n, 40, 201)
        return out, n

    def kaldi_fbank(self, wav16):
        """(T, 80) Kaldi log-fbank with the mean over time removed (CAMPPlus.extract_feature)."""
        n = 1 + (wav16.shape[0] - 400) // 160
        power = self._spec(wav16, 512, 160, 0, 400, self.povey, n, 1, remove_dc=1, preemph=0.97)
        f = self._gemm(power, self.kaldi80, self._new(n, 80), n, 80, 257)
        L.check(self.lib.cbx_cond_mel_log(f.data_ptr(), f.numel(), 0, float(np.finfo(np.float32).eps), None, self._st()))
        L.check(self.lib.cbx_cond_col_stats(f.data_ptr(), 80, n, 80, 0, None, None, None, 0, None, self._st()))
        return f

    def trim_silence(self, wav16, top_db=20.0):
        """librosa.effects.trim(top_db=20): the frame energies come from the GPU (2048-sample frames, hop 512, zero-padded
        edges), the cut points are a host decision like the reference's."""
        n = wav16.shape[0]
        nf = 1 + n // 512
        xp = torch.zeros(n + 2048, device=self.dev)
        xp[1024: 1024 + n].copy_(wav16)
        ones = torch.ones(1, 2048, device=self.dev)
        sq = self._new(n + 2048)
        # x^2 through the GEMM epilogue (x * x via mul): one row per sample, then frame sums as a strided GEMM against ones
        self._gemm(xp, torch.ones(1, 1, device=self.dev), sq, n + 2048, 1, 1, lda=1, ldw=1, ldc=1, mul=xp, ldm=1)
        en = self._gemm(sq, ones, self._new(nf, 1), nf, 1, 2048, lda=512, ldw=2048, ldc=1, kc=2048, alpha=1.0 / 2048)
        mse = en.reshape(-1).double().cpu().numpy()
        db = 10.0 * np.log10(np.maximum(mse, 1e-10))
        nz = np.nonzero(db - db.max() > -top_db)[0]
        if nz.size == 0:
            return wav16[:0]
        return wav16[int(nz[0]) * 512: min(n, (int(nz[-1]) + 1) * 512)]

    # ------------------------------------------------------------------------------------------ S3Tokenizer v2
    def s3_tokenize(self, mel):
        """mel device (T, 128) -> int32 device (T // 4 ...) codes (S3TokenizerV2.quantize)."""
        w, c, lib, st = self.w, self.cfg, self.lib, self._st()
        D, H = c.tok_dim, c.tok_heads
        hd = D // H
        T0 = mel.shape[0]
        T1 = (T0 + 2 - 3) // 2 + 1
        T = (T1 + 2 - 3) // 2 + 1
        p = "tokenizer.encoder."
        x1 = self._gemm(mel, w[p + "conv1.weight"], self._new(T1, D), T1, D, 3 * c.tok_mels, kc=c.tok_mels, a_stride=2, a_pad=1, a_rows=T0, bias=w[p + "conv1.bias"], act=2)
        x = self._gemm(x1, w[p + "conv2.weight"], self._new(T, D), T, D, 3 * D, kc=D, a_stride=2, a_pad=1, a_rows=T1, bias=w[p + "conv2.bias"], act=2)
        h, q, k, v, fsm, a = (self._new(T, D) for _ in range(6))
        S = self._new(H, T, T)
        ff = self._new(T, 4 * D)
        for i in range(self.n_blocks):
            b = p + f"blocks.{i}."
            L.check(lib.cbx_cond_layernorm(x.data_ptr(), D, h.data_ptr(), D, T, D, w[b + "attn_ln.weight"].data_ptr(), w[b + "attn_ln.bias"].data_ptr(), 1e-5, st))
            self._gemm(h, w[b + "attn.query.weight"], q, T, D, D, bias=w[b + "attn.query.bias"])
            self._gemm(h, w[b + "attn.key.weight"], k, T, D, D)
            self._gemm(h, w[b + "attn.value.weight"], v, T, D, D, bias=w[b + "attn.value.bias"])
            L.check(lib.cbx_cond_rotary(q.data_ptr(), D, T, H, hd, hd ** -0.25, st))
            L.check(lib.cbx_cond_rotary(k.data_ptr(), D, T, H, hd, hd ** -0.25, st))
            L.check(lib.cbx_cond_dwconv_add(v.data_ptr(), D, w[b + "attn.fsmn_block.weight"].data_ptr(), c.tok_fsmn_kernel, fsm.data_ptr(), D, T, D, st))
            self._gemm(q, k, S, T, T, hd, lda=D, ldw=D, ldc=T, batch=H, a_bs=hd, w_bs=hd, c_bs=T * T)
            L.check(lib.cbx_cond_softmax(S.data_ptr(), T, T * T, T, T, H, st))
            self._gemm(S, v, a, T, hd, T, lda=T, ldw=D, ldc=D, batch=H, a_bs=T * T, w_bs=hd, c_bs=hd, w_trans=1)
            self._gemm(a, w[b + "attn.out.weight"], x, T, D, D, bias=w[b + "attn.out.bias"], res=x, res2=fsm)
            L.check(lib.cbx_cond_layernorm(x.data_ptr(), D, h.data_ptr(), D, T, D, w[b + "mlp_ln.weight"].data_ptr(), w[b + "mlp_ln.bias"].data_ptr(), 1e-5, st))
            self._gemm(h, w[b + "mlp.0.weight"], ff, T, 4 * D, D, bias=w[b + "mlp.0.bias"], act=2)
            self._gemm(ff, w[b + "mlp.2.weight"], x, T, D, 4 * D, bias=w[b + "mlp.2.bias"], res=x)
        pd = self._gemm(x, w["tokenizer.quantizer._codebook.project_down.weight"], self._new(T, 8), T, 8, D, bias=w["tokenizer.quantizer._codebook.project_down.bias"])
        codes = self._new(T, dtype=torch.int32)
        L.check(lib.cbx_cond_fsq(pd.data_ptr(), codes.data_ptr(), T, st))
        return codes

    def s3_tokens_from_wav(self, wav16, max_len=None):
        mel = self.log_mel_16k(wav16)
        if max_len is not None:
            mel = mel[: max_len * 4]
        return self.s3_tokenize(mel.contiguous())

    # ------------------------------------------------------------------------------------------ CAMPPlus
    def _conv2d(self, x, wkey, bnkey, Cin, Cout, F, T, ks, sf, relu, res=None, t_major=False):
        Fo = (F + 2 * (ks // 2) - ks) // sf + 1
        y = self._new(T, Cout * Fo) if t_major else self._new(Cout, Fo, T)
        sc, sh = self.bn[bnkey]
        L.check(self.lib.cbx_cond_conv2d(x.data_ptr(), self.w[wkey].data_ptr(), sc.data_ptr(), sh.data_ptr(), res.data_ptr() if res is not None else None,
                                         y.data_ptr(), Cin, Cout, F, T, ks, sf, 1 if relu else 0, 1 if t_major else 0, self._st()))
        return y, Fo

    def campplus(self, feat):
        """feat device (T, 80) mean-normalised fbank -> device (192,) x-vector."""
        w, bn, c, lib, st = self.w, self.bn, self.cfg, self.lib, self._st()
        T = feat.shape[0]
        x = self._gemm(self.eye80, feat, self._new(80, T), 80, T, 80)            # [T][80] -> [1][80][T]
        h = "speaker_encoder.head."
        x, F = self._conv2d(x, h + "conv1.weight", h + "bn1.", 1, 32, 80, T, 3, 1, True)
        for layer in ("layer1", "layer2"):
            for j in range(2):
                r = h + f"{layer}.{j}."
                sf = 2 if j == 0 else 1
                sc = x
                if r + "shortcut.0.weight" in w:
                    sc, _ = self._conv2d(x, r + "shortcut.0.weight", r + "shortcut.1.", 32, 32, F, T, 1, sf, False)
                y, Fo = self._conv2d(x, r + "conv1.weight", r + "bn1.", 32, 32, F, T, 3, sf, True)
                x, F = self._conv2d(y, r + "conv2.weight", r + "bn2.", 32, 32, Fo, T, 3, 1, True, res=sc)
        x, F = self._conv2d(x, h + "conv2.weight", h + "bn2.", 32, 32, F, T, 3, 2, True, t_major=True)     # [T][320]
        xv = "speaker_encoder.xvector."
        ch = c.xv_init
        T1 = (T + 4 - 5) // 2 + 1
        sc, sh = bn[xv + "tdnn.nonlinear.batchnorm."]
        cmax = ch + self.xv_blocks[0][0] * c.xv_growth
        X = self._new(T1, cmax)
        self._gemm(x, w[xv + "tdnn.linear.weight"], X, T1, ch, 5 * 32 * F, kc=32 * F, a_stride=2, a_pad=2, a_rows=T, o_scale=sc, o_shift=sh, act=1, ldc=cmax)
        bnc = 4 * c.xv_growth
        nseg = -(-T1 // 100)
        y2, ctx, c1, m, gmean = self._new(T1, bnc), self._new(nseg, bnc), self._new(nseg, bnc // 2), self._new(nseg, c.xv_growth), self._new(bnc)
        for bi, (n_layers, k, dil) in enumerate(self.xv_blocks):
            for li in range(n_layers):
                l = xv + f"block{bi + 1}.tdnnd{li + 1}."
                cin = ch + li * c.xv_growth
                s1, h1 = bn[l + "nonlinear1.batchnorm."]
                s2, h2 = bn[l + "nonlinear2.batchnorm."]
                self._gemm(X, w[l + "linear1.weight"], y2, T1, bnc, cin, lda=cmax, a_scale=s1, a_shift=h1, a_relu=1, o_scale=s2, o_shift=h2, act=1)
                L.check(lib.cbx_cond_col_stats(y2.data_ptr(), bnc, T1, bnc, 2, None, None, gmean.data_ptr(), 100, ctx.data_ptr(), st))
                self._gemm(ctx, w[l + "cam_layer.linear1.weight"], c1, nseg, bnc // 2, bnc, bias=w[l + "cam_layer.linear1.bias"], act=1)
                self._gemm(c1, w[l + "cam_layer.linear2.weight"], m, nseg, c.xv_growth, bnc // 2, bias=w[l + "cam_layer.linear2.bias"], act=3)
                self._gemm(y2, w[l + "cam_layer.linear_local.weight"], X, T1, c.xv_growth, k * bnc, kc=bnc, a_dil=dil, a_pad=(k - 1) // 2 * dil, a_rows=T1,
                           mul=m, ldm=c.xv_growth, mul_div=100, ldc=cmax, c_off=cin)
            ch += n_layers * c.xv_growth
            t = xv + f"transit{bi + 1}."
            st_, sh_ = bn[t + "nonlinear.batchnorm."]
            nxt = ch // 2 + (self.xv_blocks[bi + 1][0] * c.xv_growth if bi + 1 < len(self.xv_blocks) else 0)
            Xn = self._new(T1, nxt)
            self._gemm(X, w[t + "linear.weight"], Xn, T1, ch // 2, ch, lda=cmax, a_scale=st_, a_shift=sh_, a_relu=1, ldc=nxt)
            X, cmax, ch = Xn, nxt, ch // 2
        so, ho = bn[xv + "out_nonlinear.batchnorm."]
        stats = self._new(1, 2 * ch)
        L.check(lib.cbx_cond_col_stats(X.data_ptr(), cmax, T1, ch, 1, so.data_ptr(), ho.data_ptr(), stats.data_ptr(), 0, None, st))
        sd_, hd_ = bn[xv + "dense.nonlinear.batchnorm."]
        return self._gemm(stats, w[xv + "dense.linear.weight"], self._new(1, c.xv_dim), 1, c.xv_dim, 2 * ch, o_scale=sd_, o_shift=hd_).reshape(-1)

    def xvector_from_wav(self, wav16):
        return self.campplus(self.kaldi_fbank(wav16))

    # ------------------------------------------------------------------------------------------ VoiceEncoder
    def voice_embed(self, wav16, trim=True):
        """device (256,) utterance embedding of VoiceEncoder.embeds_from_wavs([wav], 16 kHz)."""
        c, lib, st = self.cfg, self.lib, self._st()
        if trim:
            wav16 = self.trim_silence(wav16).contiguous()
        n = wav16.shape[0] // 160 + 1
        step, B, target = ve_partials(n)
        mel, _ = self.ve_mel(wav16, pad_to=target)
        H, Tw = c.ve_hidden, 160
        xp, hs, hl = self._new(B, Tw, 4 * H), self._new(B, Tw, H), self._new(B, H)
        for l, (w_ih, w_hh_t, bias) in enumerate(self.lstm):
            if l == 0:      # partial b = mel rows [b * step, b * step + 160): overlapping windows straight from the mel buffer
                self._gemm(mel, w_ih, xp, Tw, 4 * H, c.ve_mels, batch=B, a_bs=step * c.ve_mels, c_bs=Tw * 4 * H, bias=bias)
            else:
                self._gemm(hs, w_ih, xp, B * Tw, 4 * H, H, bias=bias)
            L.check(lib.cbx_cond_lstm_layer(xp.data_ptr(), w_hh_t.data_ptr(), hs.data_ptr(), hl.data_ptr(), B, Tw, H, st))
        e = self._gemm(hl, self.w["ve.proj.weight"], self._new(B, c.ve_embed), B, c.ve_embed, H, bias=self.w["ve.proj.bias"])
        L.check(lib.cbx_cond_l2norm_rows(e.data_ptr(), B, c.ve_embed, 1, st))
        m = self._new(1, c.ve_embed)
        L.check(lib.cbx_cond_mean_rows(e.data_ptr(), B, c.ve_embed, m.data_ptr(), st))
        L.check(lib.cbx_cond_l2norm_rows(m.data_ptr(), 1, c.ve_embed, 0, st))
        return m.reshape(-1)

    # ------------------------------------------------------------------------------------------ the reference's call sites
    def embed_ref(self, wav24):
        """S3Token2Mel.embed_ref on a 24 kHz device clip: dict of device tensors (prompt_feat (T, 80), prompt_token (T/2,), embedding (192,))."""
        feat = self.mel_24k(wav24)
        wav16 = self.resample(wav24, S3GEN_SR, S3_SR)
        emb = self.xvector_from_wav(wav16)
        tok = self.s3_tokens_from_wav(wav16)
        if feat.shape[0] != 2 * tok.shape[0]:
            tok = tok[: feat.shape[0] // 2]
            feat = feat[: 2 * tok.shape[0]]
        return {"prompt_feat": feat, "prompt_token": tok, "embedding": emb}

    def prepare_conditionals(self, wav, sr, speech_cond_prompt_len=150, exaggeration=0.5):
        """reference :357-384 from the decoded waveform on; returns host tensors in the layout voice_put / the reference's
        Conditionals use."""
        with self.lock, torch.cuda.stream(self.stream):
            x = self._dev_wave(wav)
            wav24 = self.resample(x, sr, S3GEN_SR)
            wav16 = self.resample(wav24, S3GEN_SR, S3_SR)
            gen = self.embed_ref(wav24[:DEC_COND_LEN].contiguous())
            t3_tok = self.s3_tokens_from_wav(wav16[:ENC_COND_LEN].contiguous(), max_len=speech_cond_prompt_len)
            spk = self.voice_embed(wav16)
            self.stream.synchronize()
        n = gen["prompt_token"].shape[0]
        return {"t3": {"speaker_emb": spk.cpu()[None], "cond_prompt_speech_tokens": t3_tok.cpu().long()[None], "emotion_adv": exaggeration * torch.ones(1, 1, 1)},
                "gen": {"prompt_token": gen["prompt_token"].cpu().long()[None], "prompt_token_len": torch.tensor([n]),
                        "prompt_feat": gen["prompt_feat"].cpu()[None], "prompt_feat_len": None, "embedding": gen["embedding"].cpu()[None]}}


def load_wav(path):
    """(float32 mono waveform in [-1, 1], sample rate) of a PCM / float WAV file (reference: librosa.load, :362)."""
    from scipy.io import wavfile
    sr, x = wavfile.read(path)
    if x.dtype == np.int16:
        x = x.astype(np.float32) / 32768.0
    elif x.dtype == np.int32:
        x = x.astype(np.float32) / 2147483648.0
    elif x.dtype == np.uint8:
        x = (x.astype(np.float32) - 128.0) / 128.0
    else:
        x = x.astype(np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1)
    return x, int(sr)
