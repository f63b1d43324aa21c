#!/usr/bin/env python
"""Fixed program for `ncu --profile-from-start off`: N streams (default 8 = 16 rows) decoded to ~290 cached positions, then TWO
decode steps with the GEMV-path kernels inside a cudaProfilerStart/Stop window.  Prints the event-timed step first."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")):
    sys.path.insert(0, p)
import torch
from cbx_b200.config import ModelConfig
from cbx_b200.native import NativeEngine
from cbx_b200.weights import random_state_dict, synthetic_conditionals

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = ModelConfig()
eng = NativeEngine(cfg, max_streams=max(8, n), n_lanes=1)
eng.load_state_dict(random_state_dict(cfg, 0))
conds = synthetic_conditionals(cfg)
v = eng.voice_put("default", conds["t3"], conds["gen"])
text = [255] + [(7 * i) % 700 + 1 for i in range(145)] + [0]
slots = [eng.t3_open(v, text, seed=1 + i, max_new=400) for i in range(n)]
eng.t3_step(slots, 100)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.t3_step(slots, 20); b.record(); torch.cuda.synchronize()
print("rows", 2 * n, "step_ms", a.elapsed_time(b) / 20)
torch.cuda.profiler.start()
eng.t3_step(slots, 2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
for s in slots:
    eng.t3_close(s)
print("ok")
eng.close()
