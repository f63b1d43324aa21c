#!/usr/bin/env python
"""Benchmark of the Chatterbox hot path (T3 -> S3Gen CFM -> HiFT) on B200.

Contract: `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line (rank 0).
A step = one pass of the hot path over BASELINE.json configs[1]: a single stream synthesising a
200-word synthetic paragraph in streamed chunks (cfg 0.5, temp 0.8, slice 35, overlap "full",
crossfade 30 ms, 10 speech tokens per word, random-init weights seed 0, fixed sampling seed).
  value : audio-seconds per second with the PCM left in HBM (engine pipeline, no D2H of audio)
  e2e   : the same through the public `TextToSpeechEngine.stream()` call: host text in, host PCM bytes out
  roofline / kernels : one extra instrumented pass (per-launch CUDA events, T3 un-graphed)
  cpu_baseline : the oracle (CPU port of the reference path) on a bounded sample, rank 0, N=1
`--impl reference` times the oracle alone on the host cores (the reference has no CUDA code of its own;
its dependency `chatterbox` is not installable here).
"""
import argparse
import asyncio
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORDS, TOK_PER_WORD = 200, 10
REQ = dict(cfg_guidance_weight=0.5, synthesis_temperature=0.8, text_processing_chunk_size=150, audio_tokens_per_slice=35,
           remove_trailing_milliseconds=0, remove_leading_milliseconds=0, chunk_overlap_strategy="full",
           crossfade_duration_milliseconds=30)


def synthetic_text(n_words: int, seed: int = 1234) -> str:
    """seeded 5-letter lowercase words, a period every 12 words."""
    r = random.Random(seed)
    out = []
    for i in range(n_words):
        w = "".join(r.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(5))
        out.append(w + ("." if (i + 1) % 12 == 0 or i == n_words - 1 else ""))
    return " ".join(out)


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.p = gpu, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "which": "measured (MEASURED_PEAKS.json; sustained bf16)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tflops": 1400.0, "which": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------ oracle legs
def oracle_sample(threads=None):
    """Bounded CPU sample of the same workload: first text chunk, T3 prefill + 42 decode steps (first slice +
    look-ahead), then one S3Gen call on the first 35-token slice (prompt 194 tokens / 388 frames).  fp32, eager."""
    import torch
    from oracle import t3 as OT, flow as OF, hift as OH
    from cbx_b200.config import ModelConfig
    from cbx_b200.weights import random_state_dict, synthetic_conditionals
    from cbx_b200.text_processing import split_text_into_chunks, SyntheticTokenizer
    if threads:
        torch.set_num_threads(threads)
    cfg = ModelConfig()
    if not hasattr(oracle_sample, "sd"):
        oracle_sample.sd = random_state_dict(cfg, 0)
        oracle_sample.conds = synthetic_conditionals(cfg)
    sd, conds = oracle_sample.sd, oracle_sample.conds
    chunk = split_text_into_chunks(synthetic_text(WORDS), 150)[0]
    ids = [255] + SyntheticTokenizer().text_to_tokens(chunk)[0].tolist() + [0]
    text = torch.tensor([ids, ids])
    g = torch.Generator().manual_seed(1234)
    t0 = time.time()
    with torch.no_grad():
        toks = []
        for tok in OT.inference_stream(sd, cfg.t3, conds["t3"], text, 42, noise_fn=lambda i: torch.empty(8194).exponential_(generator=g)):
            toks.append(tok)
        sl = [t for t in toks[:35] if t < 6561]
        while len(sl) < 3:
            sl.append(0)
        mel = OF.flow_inference(sd, cfg.flow, torch.tensor(sl), conds["gen"])
        T = mel.shape[-1]
        wav, _ = OH.hift_inference(sd, cfg.hift, mel, None, torch.zeros(9), torch.randn(9, T * 480, generator=g))
    dt = time.time() - t0
    return wav.shape[-1] / 24000.0, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    for _ in range(args.warmup):
        oracle_sample()
    secs, times = 0.0, 0.0
    for _ in range(args.steps):
        a, dt, th = oracle_sample()
        secs += a
        times += dt
    v = secs / times
    line = {"metric": "audio_sec_per_sec", "value": v, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": times / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "configs[1] single stream, 200-word paragraph (bounded sample per step: first chunk, 42 T3 steps + first 35-token S3Gen slice)"},
            "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": th, "kind": "port",
                             "sample": "oracle fp32 eager: T3 prefill + 42 decode steps, one S3Gen call on 35 tokens (T=458 CFM frames), 1.4 s audio"},
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def aggregate_ranks(ms, e2e_ms, audio_s, e2e_audio_s, world, device):
    """Replicas only: every rank runs the same work on its own GPU.  The job's time is the MAX over ranks
    (device-timed), the job's output is the SUM of what the ranks produced; no data-path collective exists."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms, e2e_ms], device=device, dtype=torch.float64)
    a = torch.tensor([audio_s, e2e_audio_s], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    return {"ms_max": ms_max, "e2e_ms_max": e2e_ms_max, "value": float(a[0]) / (ms_max / 1e3), "e2e_value": float(a[1]) / (e2e_ms_max / 1e3)}


def t3_step_leg(eng, pk):
    import torch
    nat = eng.native
    out = {}
    text = [255] + [(7 * i) % 700 + 1 for i in range(145)] + [0]
    for name, persistent in (("gemv_kernels", False), ("persistent_kernel", True)):
        try:
            nat.t3_set_persistent(persistent)
            slot = nat.t3_open(eng.voice_cache["default"], text, seed=1, max_new=1000)
            nat.t3_step([slot], 20)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nat.t3_step([slot], 200)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 200
            nat.t3_close(slot)
            pos = 34 + len(text) + 1 + 120
            nbytes = 30 * 16779264 * 2 + 8208 * 1024 * 2 + 122880 * pos * 2
            gbs = nbytes / (ms * 1e-3) / 1e9
            out[name] = {"ms_per_step": ms, "bytes_per_step": nbytes, "achieved": gbs, "unit": "GB/s", "peak": pk["hbm_gbs"], "frac": gbs / pk["hbm_gbs"],
                         "tokens_per_s": 1e3 / ms}
        except Exception as ex:   # report, never hide
            out[name] = {"error": str(ex)}
    nat.t3_set_persistent(False)
    return out


# ------------------------------------------------------------------------------------------------ B200 leg
def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from cbx_b200.engine import TextToSpeechEngine, SamplingDefaults
    from cbx_b200 import lib as L

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    text = synthetic_text(WORDS)
    sampling = SamplingDefaults(tokens_per_word=TOK_PER_WORD)
    eng = TextToSpeechEngine(f"cuda:{local}", concurrent_requests=int(os.environ.get("BENCH_CONCURRENT", "8")), sampling=sampling, seed=0)
    lib = L.load()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def value_step():
        n = [0]
        eng.device_sink = True
        eng._seq = 0
        eng._run_request(text, None, REQ["cfg_guidance_weight"], REQ["synthesis_temperature"], REQ["text_processing_chunk_size"],
                         REQ["audio_tokens_per_slice"], 0, 0, "full", 30, "bench", None, lambda k: n.__setitem__(0, n[0] + k), time.time())
        torch.cuda.synchronize()
        eng.device_sink = False
        return n[0]

    async def e2e_step():
        nbytes, t0, first = 0, time.time(), None
        eng._seq = 0
        async for chunk in eng.stream(text=text, output_format="raw_pcm", voice_id=None, request_id="bench", cancellation_token=None, **REQ):
            if first is None and len(chunk):
                first = (time.time() - t0) * 1e3
            nbytes += len(chunk)
        return nbytes, first

    async def concurrent_leg(n_streams=8, words=100):
        """BASELINE.json configs[2]: 8 concurrent streams batched on one B200, 100-word prompts, cached voice conditioning."""
        texts = [synthetic_text(words, seed=1234 + i) for i in range(n_streams)]

        async def one(i):
            nb, t0, first = 0, time.time(), None
            async for chunk in eng.stream(text=texts[i], output_format="raw_pcm", voice_id=None, request_id=f"c{i}", cancellation_token=None, **REQ):
                if first is None and len(chunk):
                    first = (time.time() - t0) * 1e3
                nb += len(chunk)
            return nb, first
        await asyncio.gather(*[one(i) for i in range(n_streams)])      # warm-up: captures the S3Gen graphs of every lane
        torch.cuda.synchronize()
        t0 = time.time()
        res = await asyncio.gather(*[one(i) for i in range(n_streams)])
        torch.cuda.synchronize()
        dt = time.time() - t0
        audio = sum(r[0] for r in res) / 2 / 24000.0
        firsts = sorted(r[1] for r in res if r[1] is not None)
        return {"workload": "configs[2]: 8 concurrent streams on one B200, 100-word prompts, cached voice conditioning", "streams": n_streams,
                "value": audio / dt, "unit": "audio-s/s", "per_stream_x_realtime": audio / dt / n_streams,
                "first_chunk_ms_p50": firsts[len(firsts) // 2] if firsts else None, "seconds": dt}

    async def main():
        await eng.ainit()
        for _ in range(max(args.warmup, 3)):
            value_step()
        launches0 = eng.native.gpu_launches()
        clocks = ClockSampler(local)
        barrier()
        clocks.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        samples = 0
        for _ in range(args.steps):
            samples += value_step()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        clk = clocks.stop()
        launches = eng.native.gpu_launches() - launches0
        # e2e through stream(): host text in, host PCM out
        for _ in range(1):
            await e2e_step()
        barrier()
        t0 = time.time()
        e2e_bytes, firsts = 0, []
        for _ in range(args.steps):
            nb, f = await e2e_step()
            e2e_bytes += nb
            firsts.append(f)
        torch.cuda.synchronize()
        barrier()
        e2e_s = time.time() - t0
        audio_s = samples / 24000.0
        e2e_audio = e2e_bytes / 2 / 24000.0
        agg = aggregate_ranks(ms, e2e_s * 1e3, audio_s, e2e_audio, world, "cuda")
        ms_max, e2e_ms_max, value, e2e_val = agg["ms_max"], agg["e2e_ms_max"], agg["value"], agg["e2e_value"]
        if rank != 0:
            return
        # instrumented pass for the roofline (rank 0)
        lib.cbx_profile_begin()
        value_step()
        n = 8
        counts, pms, work = (C.c_int64 * n)(), (C.c_double * n)(), (C.c_double * n)()
        lib.cbx_profile_end(counts, pms, work, n)
        names = ["gemm_mma_kernel (CFM/HiFT/encoder/T3-prefill GEMM + implicit conv)", "attn_kernel (CFM/encoder/prefill attention)",
                 "t3_decode_step (GEMV projections + decode attention + sampler of one step, graph replay)", "decode_attn_kernel", "sampler_kernel", "norm_kernel", "elementwise", "hift misc"]
        kern = []
        for i in range(n):
            if counts[i]:
                kern.append({"kernel": names[i], "launches": int(counts[i]), "ms": pms[i], "work": work[i],
                             "rate": work[i] / (pms[i] * 1e-3) / (1e12 if i < 2 else 1e9), "rate_unit": "TFLOP/s" if i < 2 else "GB/s"})
        tot = sum(k["ms"] for k in kern) or 1.0
        for k in kern:
            k["share"] = k["ms"] / tot
        pk = peaks()
        dom = max(range(n), key=lambda i: pms[i])
        if dom < 2:
            ach = work[dom] / (pms[dom] * 1e-3) / 1e12
            roof = {"kernel": names[dom], "bound": "tensor", "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"], "traffic": None,
                    "peak_source": pk["which"], "avg_launch_us": pms[dom] * 1e3 / counts[dom]}
        else:
            ach = work[dom] / (pms[dom] * 1e-3) / 1e9
            roof = {"kernel": names[dom], "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None,
                    "peak_source": pk["which"], "avg_launch_us": pms[dom] * 1e3 / counts[dom]}
        gv = next((k for k in kern if k["kernel"].startswith("t3_decode_step")), None)
        if gv:
            roof["t3_decode_step_in_workload"] = {"bound": "hbm", "achieved": gv["rate"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gv["rate"] / pk["hbm_gbs"]}
        # dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), when there is one
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")) as fh:
                tr = json.load(fh)
            key = "gemm_tc_kernel" if dom == 0 else ("attn_tc_kernel" if dom == 1 else ("t3_decode_step" if dom == 2 else None))
            if key in tr:
                roof["traffic"] = tr[key]["dram_bytes_per_launch"]
                roof["traffic_note"] = tr[key].get("note", "")
        except Exception:
            pass
        # T3 decode step alone, both implementations (200 steps of one stream, CUDA events): bytes = weights + KV read
        roof["t3_step"] = t3_step_leg(eng, pk)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            a, dt, th = oracle_sample()
            cpu = {"value": a / dt, "unit": "audio-s/s", "cores": th, "kind": "port",
                   "sample": "oracle fp32 eager on host: T3 prefill + 42 decode steps and one S3Gen call on the first 35-token slice (1.4 s audio); %.1f s" % dt}
        firsts = [f for f in firsts if f is not None]
        line = {"metric": "audio_sec_per_sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "configs[1]: single stream per GPU, 200-word paragraph streamed in chunks, cfg 0.5, temp 0.8, slice 35, overlap full, crossfade 30 ms, 10 speech tokens/word, fixed seed",
                           "chunk_parallelism": eng.chunk_parallelism,
                           "audio_s_per_step": audio_s / args.steps, "weights": "random-init seed 0", "l2": "no flush needed: every T3 step streams 1.02 GB of weights (> 126 MB L2)"},
                "clocks": clk,
                "e2e": {"value": e2e_val, "unit": "audio-s/s", "h2d_bytes_per_step": int(4 * (WORDS * 6 + WORDS * TOK_PER_WORD * 4)), "d2h_bytes_per_step": int(e2e_bytes / args.steps + 4 * WORDS * TOK_PER_WORD),
                        "first_chunk_ms_p50": statistics.median(firsts) if firsts else None, "rtf": (e2e_ms_max / 1e3) / max(e2e_audio, 1e-9)},
                "gpu_launches": int(launches), "roofline": roof, "kernels": kern}
        if cpu:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_concurrent:
            line["concurrent8"] = await concurrent_leg()
        line["s3gen_batch_sizes"] = {str(k): int(v) for k, v in sorted(eng.s3gen.batches.items())}
        line["s3gen_padding"] = {"tokens_requested": int(eng.s3gen.pad_stats[0]), "tokens_after_padding": int(eng.s3gen.pad_stats[1])}
        print(json.dumps(line), flush=True)

    asyncio.run(main())
    eng.shutdown()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-concurrent", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
