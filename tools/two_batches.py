#!/usr/bin/env python
"""Does running two S3Gen batches concurrently (two workspaces, two host threads) beat one batch twice the size?"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")):
    sys.path.insert(0, p)
import torch
from cbx_b200.config import ModelConfig
from cbx_b200.native import NativeEngine
from cbx_b200.weights import random_state_dict, synthetic_conditionals
cfg = ModelConfig(); eng = NativeEngine(cfg, max_streams=8, n_lanes=1); eng.load_state_dict(random_state_dict(cfg, 0))
conds = synthetic_conditionals(cfg); v = eng.voice_put("default", conds["t3"], conds["gen"])
def calls(B, n, off=0): return [(v, [(i * 37 + 11 * (b + off)) % 6561 for i in range(n)], None, 1 + b) for b in range(B)]
def run_threads(k, B, n, reps=3):
    def work(i):
        torch.cuda.set_device(0)
        with torch.cuda.stream(torch.cuda.Stream()):
            for _ in range(reps):
                eng.s3gen_infer_batch(calls(B, n, 16 * i))
            torch.cuda.current_stream().synchronize()
    th = [threading.Thread(target=work, args=(i,)) for i in range(k)]
    torch.cuda.synchronize(); t0 = time.time()
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    return (time.time() - t0) * 1e3 / reps
for n in (35, 140):
    for (k, B) in ((1, 8), (2, 4), (1, 4), (2, 2), (2, 8), (1, 16)):
        run_threads(k, B, n, reps=1)
        ms = run_threads(k, B, n)
        print(f"n={n} threads={k} batch={B}: {ms:.1f} ms per round, {k * B * n * 0.04 / ms * 1e3:.0f} audio-s/s", flush=True)
eng.close()
