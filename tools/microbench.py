#!/usr/bin/env python
"""Stage-level timings on one B200 (CUDA events): T3 decode steps at 1..8 streams, S3Gen call latency by token count."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")):
    sys.path.insert(0, p)
import torch
from cbx_b200.config import ModelConfig
from cbx_b200.native import NativeEngine
from cbx_b200.weights import random_state_dict, synthetic_conditionals


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    cfg = ModelConfig()
    stream_counts = tuple(int(x) for x in os.environ.get("MICRO_STREAMS", "1,2,4,8").split(","))
    eng = NativeEngine(cfg, max_streams=max(8, max(stream_counts)), n_lanes=1)
    eng.load_state_dict(random_state_dict(cfg, 0))
    conds = synthetic_conditionals(cfg)
    v = eng.voice_put("default", conds["t3"], conds["gen"])
    out = {}
    text = [255] + [(7 * i) % 700 + 1 for i in range(145)] + [0]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    if only == "align":   # decode step with and without alignment-based EOS control (two extra kernels at the probe layer)
        for ns in stream_counts:
            for on in (False, True):
                eng.t3_set_alignment_eos(on, 9)
                slots = [eng.t3_open(v, text, seed=i, max_new=1000) for i in range(ns)]
                eng.t3_step(slots, 20)
                torch.cuda.synchronize()
                a, b = ev(), ev()
                a.record(); eng.t3_step(slots, 200); b.record()
                torch.cuda.synchronize()
                out[f"t3_streams{ns}_alignment_{'on' if on else 'off'}"] = {"ms_per_step": a.elapsed_time(b) / 200}
                if on:
                    out[f"t3_streams{ns}_alignment_on"]["analyzer"] = eng.t3_alignment_peek(slots[0], rows=False)[0]
                for s in slots:
                    eng.t3_close(s)
        print(json.dumps(out, indent=1))
        return
    for ns in (() if only == "s3b" else stream_counts):
        t0 = time.time()
        slots = [eng.t3_open(v, text, seed=i, max_new=1000) for i in range(ns)]
        torch.cuda.synchronize()
        prefill_ms = (time.time() - t0) * 1e3 / ns
        eng.t3_step(slots, 20)
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        eng.t3_step(slots, 200)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 200
        pos = 34 + len(text) + 1 + 120
        wbytes = 30 * 16779264 * 2 + 8208 * 1024 * 2
        kvbytes = 122880 * pos * 2 * ns
        import ctypes as C
        tr = (C.c_ulonglong * 32)()
        eng.lib.cbx_t3_mega_trace(tr)
        names = ["P1 wait z + norm", "P1 qkv items", "P2 attention", "P3 o-proj", "P4 wait y", "P4 norm + gate/up", "P5 down"]
        trace = {f"{i:02d} {names[i]}": int(tr[i + 1]) - int(tr[i]) for i in range(7)}
        trace["detail_rel_to_P1_start"] = {k: int(tr[i]) - int(tr[0]) for k, i in (("z added", 16), ("norm done", 1), ("qkv done", 2), ("att polled", 8), ("att sync1", 9), ("att rope", 10), ("att loop", 11),
                                                                                   ("att reduce", 12), ("att sync3", 13), ("att done", 3), ("oproj staged", 14), ("oproj done", 4), ("y added", 5),
                                                                                   ("gu done", 6), ("act staged", 15), ("down done", 7))}
        pr = (C.c_longlong * 1024)()
        eng.lib.cbx_t3_mega_prof(pr)
        prof = [(int(pr[4 * i]), int(pr[4 * i + 1]), int(pr[4 * i + 2])) for i in range(148)]
        out[f"t3_streams{ns}"] = {"prof_cycles_mbar_cnt_total": prof, "layer1_phase_ns": trace, "ms_per_step": ms, "tok_s": ns * 1e3 / ms, "audio_s_per_s": ns * 1e3 / ms / 25, "prefill_ms": prefill_ms,
                                  "hbm_gbs": (wbytes + kvbytes) / (ms * 1e-3) / 1e9}
        for s in slots:
            eng.t3_close(s)
    for n in (() if only in ("t3", "s3b") else (35, 70, 140, 245)):
        toks = [(i * 37) % 6561 for i in range(n)]
        for _ in range(2):
            eng.s3gen_infer(v, toks, seed=1)
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(3):
            eng.s3gen_infer(v, toks, seed=1)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        a.record()
        for _ in range(3):
            eng.flow_infer(v, toks)
        b.record()
        torch.cuda.synchronize()
        fms = a.elapsed_time(b) / 3
        T = 2 * (194 + n)
        flops = 2 * (66.08e6 + 57344 * T) * T * 20
        out[f"s3gen_n{n}"] = {"ms": ms, "flow_eager_ms": fms, "T": T, "cfm_tflops_vs_graph_total": flops / (ms * 1e-3) / 1e12, "audio_s": n * 0.04}
    if only in ("", "s3b"):
        for n in (35, 140):
            toks = [[(i * 37 + 11 * b) % 6561 for i in range(n + (b % 3))] for b in range(8)]
            for B in (1, 2, 4, 8):
                calls = [(v, toks[b], None, 1 + b) for b in range(B)]
                eng.s3gen_infer_batch(calls)
                torch.cuda.synchronize()
                t0 = time.time()
                a, b_ = ev(), ev()
                a.record()
                for _ in range(2):
                    eng.s3gen_infer_batch(calls)
                b_.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b_) / 2
                out[f"s3gen_batch_n{n}_B{B}"] = {"ms": ms, "wall_ms": (time.time() - t0) * 500, "ms_per_call": ms / B, "audio_s_per_s": B * n * 0.04 / (ms * 1e-3)}
    out["gemm_tc_launches"] = int(eng.lib.cbx_gemm_tc_launches())
    print(json.dumps(out, indent=1))
    eng.close()


if __name__ == "__main__":
    main()
