#!/usr/bin/env python
"""Per-kernel-class device time of one batched S3Gen call (in-engine per-launch profiler)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")):
    sys.path.insert(0, p)
import torch
from cbx_b200.config import ModelConfig
from cbx_b200.native import NativeEngine
from cbx_b200.weights import random_state_dict, synthetic_conditionals

cfg = ModelConfig()
eng = NativeEngine(cfg, max_streams=8, n_lanes=1)
eng.load_state_dict(random_state_dict(cfg, 0))
conds = synthetic_conditionals(cfg)
v = eng.voice_put("default", conds["t3"], conds["gen"])
names = ["gemm", "attn", "gemv", "decode_attn", "sampler", "norm", "elementwise", "hift_misc"]
out = {}
for (B, n) in ((1, 35), (8, 35), (8, 140)):
    calls = [(v, [(i * 37 + 11 * b) % 6561 for i in range(n)], None, 1 + b) for b in range(B)]
    eng.s3gen_infer_batch(calls)
    torch.cuda.synchronize()
    eng.lib.cbx_profile_begin()
    eng.s3gen_infer_batch(calls)
    torch.cuda.synchronize()
    cnt, ms, work = (C.c_int64 * 8)(), (C.c_double * 8)(), (C.c_double * 8)()
    eng.lib.cbx_profile_end(cnt, ms, work, 8)
    out[f"B{B}_n{n}"] = {names[i]: {"launches": int(cnt[i]), "ms": round(ms[i], 2), "rate": round(work[i] / max(ms[i], 1e-9) / (1e9 if i < 2 else 1e6), 1)} for i in range(8) if cnt[i]}
print(json.dumps(out, indent=1))
eng.close()
