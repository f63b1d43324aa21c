#!/usr/bin/env python
"""One eager S3Gen call (and a few T3 steps) between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")):
    sys.path.insert(0, p)
import torch
from cbx_b200.config import ModelConfig
from cbx_b200.native import NativeEngine
from cbx_b200.weights import random_state_dict, synthetic_conditionals

what = sys.argv[1] if len(sys.argv) > 1 else "s3gen"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 35
cfg = ModelConfig()
eng = NativeEngine(cfg, max_streams=8, n_lanes=1)
eng.load_state_dict(random_state_dict(cfg, 0))
conds = synthetic_conditionals(cfg)
v = eng.voice_put("default", conds["t3"], conds["gen"])
toks = [(i * 37) % 6561 for i in range(n)]
text = [255] + [(7 * i) % 700 + 1 for i in range(145)] + [0]
if what == "s3batch":        # one batched call of B x n tokens (argv: s3batch n B)
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    calls = [(v, [(i * 37 + 11 * b) % 6561 for i in range(n)], None, 1 + b) for b in range(B)]
    eng.s3gen_infer_batch(calls)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.s3gen_infer_batch(calls)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
elif what == "s3gen":
    mel = eng.flow_infer(v, toks)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    mel = eng.flow_infer(v, toks)
    wav, src = eng.hift_infer(mel)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    nstreams = n
    slots = [eng.t3_open(v, text, seed=i, max_new=400) for i in range(nstreams)]
    noise = torch.empty(2, nstreams, 8194, device="cuda").exponential_()
    eng.t3_step(slots, 2, noise=noise)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.t3_step(slots, 2, noise=noise)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
