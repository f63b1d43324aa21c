#!/usr/bin/env python
"""Short, fixed program for `ncu`: T3 prefill + 3 decode steps with the GEMV kernels + 2 with the persistent kernel,
then ONE batched S3Gen call (4 calls of 35 tokens, T = 458 CFM frames each, un-graphed).  Run it plain first (must exit 0),
then under ncu (profiles/README.md has the exact commands)."""
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
text = [255] + [(7 * i) % 700 + 1 for i in range(145)] + [0]
slot = eng.t3_open(v, text, seed=1, max_new=100)
eng.t3_step([slot], 3)
eng.t3_set_persistent(True)
eng.t3_step([slot], 2)
eng.t3_set_persistent(False)
n, done = eng.t3_poll(slot)
eng.t3_close(slot)
calls = [(v, [(i * 37 + 11 * b) % 6561 for i in range(35)], None, 1 + b) for b in range(4)]
outs = eng.s3gen_infer_batch(calls)
torch.cuda.synchronize()
assert n == 5 and all(torch.isfinite(w).all() for w, _ in outs)
print("ok", n, [tuple(w.shape) for w, _ in outs])
eng.close()
