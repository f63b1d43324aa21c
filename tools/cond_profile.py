#!/usr/bin/env python
"""Stage timings of the GPU voice-conditioning path on a 10 s clip (CUDA events, median of 5)."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from cbx_b200.config import ModelConfig
from cbx_b200.conditioning import ConditioningEncoders
from cbx_b200.weights import random_state_dict

cfg = ModelConfig()
enc = ConditioningEncoders(random_state_dict(cfg, 0, parts=("cond",)), cfg.cond, device=0)
t = np.arange(24000 * 10) / 24000.0
wav = (0.3 * np.sin(2 * np.pi * 170 * t) * np.clip(np.sin(2 * np.pi * 1.3 * t), 0, None) + 0.03 * np.random.default_rng(0).standard_normal(t.shape[0])).astype(np.float32)
w24 = enc._dev_wave(wav)
w16 = enc.resample(w24, 24000, 16000).contiguous()


def timed(name, fn):
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{name:34s} {statistics.median(ts[1:]):8.2f} ms")


timed("resample 24k -> 16k", lambda: enc.resample(w24, 24000, 16000))
timed("mel 24 kHz (80)", lambda: enc.mel_24k(w24))
timed("log-mel 16 kHz (128)", lambda: enc.log_mel_16k(w16))
mel = enc.log_mel_16k(w16).contiguous()
timed("S3Tokenizer encoder + FSQ (250 tok)", lambda: enc.s3_tokenize(mel))
timed("kaldi fbank", lambda: enc.kaldi_fbank(w16))
fb = enc.kaldi_fbank(w16)
timed("CAMPPlus", lambda: enc.campplus(fb))
timed("VoiceEncoder (trim + mel + LSTM)", lambda: enc.voice_embed(w16))
timed("prepare_conditionals (device part)", lambda: enc.prepare_conditionals(wav, 24000))
# VoiceEncoder sub-stages
from cbx_b200.conditioning import ve_partials
import ctypes as C
from cbx_b200 import lib as L
timed("  VE: trim_silence", lambda: enc.trim_silence(w16))
w16t = enc.trim_silence(w16).contiguous()
n = w16t.shape[0] // 160 + 1
step, B, target = ve_partials(n)
timed("  VE: mel (40)", lambda: enc.ve_mel(w16t, pad_to=target))
mel40, _ = enc.ve_mel(w16t, pad_to=target)
H = 256
xp, hs, hl = enc._new(B, 160, 4 * H), enc._new(B, 160, H), enc._new(B, H)
w_ih, w_hh_t, bias = enc.lstm[0]
timed(f"  VE: layer-0 input GEMM (B={B})", lambda: enc._gemm(mel40, w_ih, xp, 160, 4 * H, 40, batch=B, a_bs=step * 40, c_bs=160 * 4 * H, bias=bias))
timed("  VE: one LSTM layer recurrence", lambda: L.check(enc.lib.cbx_cond_lstm_layer(xp.data_ptr(), w_hh_t.data_ptr(), hs.data_ptr(), hl.data_ptr(), B, 160, H, enc._st())))
w_ih1, _, bias1 = enc.lstm[1]
timed("  VE: layer-1 input GEMM", lambda: enc._gemm(hs, w_ih1, xp, B * 160, 4 * H, H, bias=bias1))
