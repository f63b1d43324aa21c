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
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "which": "measured (MEASURED_PEAKS.json; sustained bf16)", "which_hbm": "measured (MEASURED_PEAKS.json; copy bandwidth, read + write bytes)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tflops": 1400.0, "which": "fallback (B200_PROFILING.md)", "which_hbm": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------ oracle legs
WORKLOAD = ("configs[1]: single stream per GPU, 200-word paragraph streamed in chunks, cfg 0.5, temp 0.8, slice 35, overlap full, "
            "crossfade 30 ms, 10 speech tokens/word, fixed seed")


def host_threads():
    """All host cores, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)
    return n


def slice_schedule(n_tokens, slice_len=35, look_ahead=7):
    """Token counts of the slices the reference emits for one text chunk (src/tts_streaming.py:499-501, :535-565)."""
    out, consumed = [], 0
    while n_tokens - consumed >= slice_len + look_ahead:
        out.append(slice_len)
        consumed += slice_len
    if n_tokens - consumed > 0:
        out.append(n_tokens - consumed)
    return out


class OracleChunk:
    """The reference's CPU path for the FIRST text chunk of the benchmark paragraph, on the oracle port: T3 prefill + decode
    (CFG rows, sampling), then per 35-token slice one S3Gen call over [prompt | all tokens so far] ("full" overlap: token
    accumulation, EOS(0) append, <6561 filter, source cache chained; src/tts_streaming.py:655-699) and only the new tail
    counted as audio.  `n_slices` bounds the sample: the T3 loop stops at the tokens those slices need (35 k + 7 look-ahead)."""

    def __init__(self):
        import torch
        from cbx_b200.config import ModelConfig
        from cbx_b200.weights import random_state_dict, synthetic_conditionals
        from cbx_b200.text_processing import split_text_into_chunks, SyntheticTokenizer
        self.cfg = ModelConfig()
        self.sd = random_state_dict(self.cfg, 0)
        self.conds = synthetic_conditionals(self.cfg)
        chunk = split_text_into_chunks(synthetic_text(WORDS), 150)[0]
        self.words = len(chunk.split())
        ids = [255] + SyntheticTokenizer().text_to_tokens(chunk)[0].tolist() + [0]
        self.text = torch.tensor([ids, ids])
        self.n_tokens = TOK_PER_WORD * self.words
        self.schedule = slice_schedule(self.n_tokens)

    def run(self, n_slices=None):
        """-> (audio seconds emitted, seconds, {"t3_s", "s3gen_s": [per slice]})"""
        import torch
        from oracle import t3 as OT, flow as OF, hift as OH
        sched = self.schedule if n_slices is None else self.schedule[:max(1, n_slices)]
        whole = len(sched) == len(self.schedule)
        need = self.n_tokens if whole else sum(sched) + 7
        g = torch.Generator().manual_seed(1234)
        t0 = time.time()
        with torch.no_grad():
            toks = []
            for tok in OT.inference_stream(self.sd, self.cfg.t3, self.conds["t3"], self.text, need,
                                           noise_fn=lambda i: torch.empty(8194).exponential_(generator=g)):
                toks.append(tok)
            t_t3 = time.time() - t0
            acc, cache, prev_len, emitted, per = [], None, 0, 0, []
            for k, n in enumerate(sched):
                t1 = time.time()
                acc = acc + toks[sum(sched[:k]): sum(sched[:k]) + n]
                cur = list(acc) + ([0] if (whole and k == len(sched) - 1) else [])
                cur = [t for t in cur if t < 6561]
                while len(cur) < 3:
                    cur.append(0)
                mel = OF.flow_inference(self.sd, self.cfg.flow, torch.tensor(cur), self.conds["gen"])
                T = mel.shape[-1]
                wav, cache = OH.hift_inference(self.sd, self.cfg.hift, mel, cache, torch.zeros(9), torch.randn(9, T * 480, generator=g))
                emitted += wav.shape[-1] - prev_len
                prev_len = wav.shape[-1]
                per.append(time.time() - t1)
        return emitted / 24000.0, time.time() - t0, {"t3_s": t_t3, "s3gen_s": per}

    def slices_for_budget(self, probe, seconds):
        """Largest slice count whose predicted cost fits `seconds`, from a one-slice probe (T3 cost per token; S3Gen cost taken
        as proportional to the sequence length 388 + 70 k frames)."""
        t3_tok = probe["t3_s"] / 42.0
        s1 = probe["s3gen_s"][0] / 458.0
        best = 1
        for k in range(1, len(self.schedule) + 1):
            whole = k == len(self.schedule)
            ntok = self.n_tokens if whole else sum(self.schedule[:k]) + 7
            cost = t3_tok * ntok + sum(s1 * (388 + 2 * sum(self.schedule[:j + 1])) for j in range(k))
            if cost <= seconds:
                best = k
        return best

    def describe(self, k, th):
        sched = self.schedule[:k]
        return (f"oracle fp32 eager on {th} host threads: first text chunk of the paragraph ({self.words} words, {self.n_tokens} tokens), "
                f"T3 prefill + {self.n_tokens if k == len(self.schedule) else sum(sched) + 7} decode steps, {k} of its {len(self.schedule)} "
                f"full-overlap S3Gen slices {sched} (prompt 194 tokens; every slice re-synthesises all accumulated tokens)")


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port; the reference's own model package is not installable),
    all host cores, same config / metric / unit as the B200 arm; each step is a bounded sample of that workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    th = host_threads()
    oc = OracleChunk()
    _, _, probe = oc.run(1)                                   # untimed warm-up, also the cost probe
    budget = float(os.environ.get("BENCH_REF_BUDGET_S", "240"))
    k = oc.slices_for_budget(probe, budget / max(1, args.steps))
    for _ in range(max(0, args.warmup - 1)):
        oc.run(1)                                             # warm-up steps need not be full samples: eager CPU has no graphs to capture
    secs, times = 0.0, 0.0
    for _ in range(args.steps):
        a, dt, _ = oc.run(k)
        secs += a
        times += dt
    v = secs / times
    line = {"metric": "audio_sec_per_sec", "value": v, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": times / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference", "config": bench_config(),
            "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": th, "kind": "port", "sample": oc.describe(k, th),
                             "audio_s_per_step": secs / args.steps},
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bench_config():
    """`config` of BOTH arms (the driver compares them): the workload, not how much of it a step samples."""
    return {"workload": WORKLOAD, "words": WORDS, "tokens_per_word": TOK_PER_WORD, "weights": "random-init seed 0",
            "voice": "cached conditioning of trump.wav's shapes (194 prompt tokens, 388 mel frames)",
            "l2": "no flush needed: every T3 step streams 1.02 GB of weights (> 126 MB L2)"}


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
def conditioning_leg(eng, with_cpu):
    """configs[4]'s per-request extra: voice-cloning conditioning computed from a clip (reference prepare_conditionals :357-384) --
    10 s of synthetic 24 kHz audio through the GPU encoders (host waveform in, host conditioning tensors out), median of 3; the
    oracle port on the host cores beside it when the CPU baseline is on."""
    import numpy as np
    import torch
    t = np.arange(24000 * 10) / 24000.0
    rng = np.random.default_rng(0)
    wav = (0.3 * np.sin(2 * np.pi * 170 * t) * np.clip(np.sin(2 * np.pi * 1.3 * t), 0, None) + 0.03 * rng.standard_normal(t.shape[0])).astype(np.float32)
    enc = eng.conditioning_encoders()
    ts = []
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.time()
        c = enc.prepare_conditionals(wav, 24000)
        ts.append((time.time() - t0) * 1e3)
    out = {"workload": "prepare_conditionals on a 10 s clip (24 kHz): mel, S3Tokenizer-v2, CAMPPlus, VoiceEncoder", "clip_s": 10.0,
           "gpu_ms": statistics.median(ts[1:]), "prompt_tokens": int(c["gen"]["prompt_token"].shape[1]), "dtype": "f32"}
    if with_cpu:
        from oracle import cond as OC
        from cbx_b200.config import ModelConfig
        from cbx_b200.weights import random_state_dict
        sd = random_state_dict(ModelConfig(), 0, parts=("cond",))
        with torch.no_grad():
            t0 = time.time()
            OC.prepare_conditionals(sd, torch.from_numpy(wav))
            out["cpu_port_ms"] = (time.time() - t0) * 1e3
            out["cpu_cores"] = host_threads()
    return out


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
    from cbx_b200.config import ModelConfig
    from cbx_b200.weights import random_state_dict
    text = synthetic_text(WORDS)
    sampling = SamplingDefaults(tokens_per_word=TOK_PER_WORD)
    # BASELINE.json configs: random-init Chatterbox weights (seed 0), passed explicitly -- the engine refuses to invent weights
    eng = TextToSpeechEngine(f"cuda:{local}", state_dict=random_state_dict(ModelConfig(), 0), concurrent_requests=int(os.environ.get("BENCH_CONCURRENT", "8")),
                             sampling=sampling, seed=0, encoder_state_dict=random_state_dict(ModelConfig(), 0, parts=("cond",)))
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

    async def run_streams(texts, tag):
        """All texts as concurrent requests through stream(): -> (audio seconds, wall seconds, first-PCM-chunk ms per request)."""
        async def one(i):
            nb, t0, first = 0, time.time(), None
            async for chunk in eng.stream(text=texts[i], output_format="raw_pcm", voice_id=None, request_id=f"{tag}{i}", cancellation_token=None, **REQ):
                if first is None and len(chunk):
                    first = (time.time() - t0) * 1e3
                nb += len(chunk)
            return nb, first
        torch.cuda.synchronize()
        t0 = time.time()
        res = await asyncio.gather(*[one(i) for i in range(len(texts))])
        torch.cuda.synchronize()
        dt = time.time() - t0
        return sum(r[0] for r in res) / 2 / 24000.0, dt, [r[1] for r in res if r[1] is not None]

    def pct(xs, q):
        xs = sorted(xs)
        return xs[min(len(xs) - 1, int(round(q * (len(xs) - 1))))] if xs else None

    async def concurrent_leg(n_streams=8, words=100, repeats=5):
        """BASELINE.json configs[2]: 8 concurrent streams batched on one B200, 100-word prompts, cached voice conditioning.
        One untimed pass, then `repeats` timed passes (all requests arrive together); first-chunk percentiles over all
        repeats x streams."""
        texts = [synthetic_text(words, seed=1234 + i) for i in range(n_streams)]
        await run_streams(texts, "w")
        vals, firsts = [], []
        for r in range(repeats):
            a, dt, f = await run_streams(texts, f"c{r}_")
            vals.append(a / dt)
            firsts += f
        v = statistics.median(vals)
        return {"workload": "configs[2]: 8 concurrent streams on one B200, 100-word prompts, cached voice conditioning", "streams": n_streams,
                "repeats": repeats, "value": v, "unit": "audio-s/s", "values": vals, "spread": (max(vals) - min(vals)) / v, "per_stream_x_realtime": v / n_streams,
                "first_chunk_ms_p50": pct(firsts, 0.5), "first_chunk_ms_p95": pct(firsts, 0.95), "first_chunk_ms_min": min(firsts) if firsts else None,
                "target": {"audio_s_per_s": 400.0, "first_chunk_ms_p50": 150.0}}

    async def mixed_leg(per_gpu=8, repeats=2):
        """BASELINE.json configs[3]: 8 x N concurrent streams of mixed 20-400-word prompts (word counts randint(20, 400), seed
        1234), dealt round-robin to the ranks as the reference's ZeroMQ PUSH socket deals requests to workers (src/master.py:56-98,
        SURVEY 8e: not load-aware).  Every rank runs its share concurrently; job time = max over ranks, audio = sum over ranks."""
        r = random.Random(1234)
        counts = [r.randint(20, 400) for _ in range(per_gpu * world)]
        mine = [(i, c) for i, c in enumerate(counts) if i % world == rank]
        texts = [synthetic_text(c, seed=5000 + i) for i, c in mine]
        await run_streams(texts[:2], "mw")
        out = []
        for k in range(repeats):
            barrier()
            a, dt, f = await run_streams(texts, f"m{k}_")
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            tmin = t.clone()
            au = torch.tensor([a], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
                dist.all_reduce(au, op=dist.ReduceOp.SUM)
            out.append((float(au[0]) / float(t[0]), float(t[0]), float(tmin[0]), f))
        best = max(out, key=lambda x: x[0])
        return {"workload": f"configs[3]: {per_gpu * world} concurrent streams of mixed 20-400-word prompts over {world} GPU(s), round-robin dispatch",
                "streams": per_gpu * world, "words_total": sum(counts), "value": statistics.median(o[0] for o in out), "unit": "audio-s/s", "values": [o[0] for o in out],
                "slowest_rank_s": best[1], "fastest_rank_s": best[2], "imbalance": best[1] / max(best[2], 1e-9),
                "first_chunk_ms_p50_rank0": pct(best[3], 0.5), "first_chunk_ms_p95_rank0": pct(best[3], 0.95)}

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
        mixed = None
        if not args.no_concurrent and (world > 1 or os.environ.get("BENCH_MIXED", "0") == "1"):
            mixed = await mixed_leg()
        if rank != 0:
            return
        # instrumented pass for the roofline (rank 0)
        lib.cbx_profile_begin()
        value_step()
        n = 8
        counts, pms, work = (C.c_int64 * n)(), (C.c_double * n)(), (C.c_double * n)()
        lib.cbx_profile_end(counts, pms, work, n)
        names = ["gemm_tc_kernel class (tcgen05 GEMM + implicit conv: CFM / HiFT / encoder / T3 prefill)", "attn_fa_kernel class (tcgen05 flash attention, P and O in tensor memory: CFM / encoder / T3 prefill)",
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
                    "peak_source": pk["which_hbm"], "avg_launch_us": pms[dom] * 1e3 / counts[dom]}
            # the tensor-bound runner-up next to it: both classes matter in this workload
            alt = max(range(2), key=lambda i: pms[i])
            a2 = work[alt] / (pms[alt] * 1e-3) / 1e12
            roof["tensor_class"] = {"kernel": names[alt], "bound": "tensor", "achieved": a2, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": a2 / pk["tflops"],
                                    "share": pms[alt] / sum(pms[:n])}
        gv = next((k for k in kern if k["kernel"].startswith("t3_decode_step")), None)
        if gv:
            roof["t3_decode_step_in_workload"] = {"bound": "hbm", "achieved": gv["rate"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gv["rate"] / pk["hbm_gbs"]}
        # dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), when there is one
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")) as fh:
                tr = json.load(fh)
            key = "gemm_tc_kernel" if dom == 0 else ("attn_fa_kernel" if dom == 1 else None)
            if key in tr:
                roof["traffic"] = tr[key]["dram_bytes_per_launch"]
                roof["traffic_note"] = ("ncu --set full --clock-control none (tools/ncu_capture_r2.sh: one batched S3Gen call of 8 x 140 tokens; cold-cache, "
                                        "serialised); dram read + write bytes averaged over the captured launches of " + key + "; profiles/ncu_traffic_r2.json")
            elif dom == 2:      # one decode step = 30 x (QKV + O + gate/up + down) captured launches + the head: weights once (the KV cache is extra)
                g = tr.get("gemv_kernel", {}).get("dram_bytes_per_launch")
                if g:
                    roof["traffic"] = g * 4 * 30 + 16.8e6
                    roof["traffic_note"] = "30 x 4 captured gemv_kernel launches (2 rows) + the 16.8 MB head; decode attention's KV reads are not in this figure"
        except Exception:
            pass
        # T3 decode step alone, both implementations (200 steps of one stream, CUDA events): bytes = weights + KV read
        roof["t3_step"] = t3_step_leg(eng, pk)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            th = host_threads()
            oc = OracleChunk()
            _, _, probe = oc.run(1)
            k = oc.slices_for_budget(probe, float(os.environ.get("BENCH_CPU_BUDGET_S", "25")))
            a, dt, _ = oc.run(k)
            cpu = {"value": a / dt, "unit": "audio-s/s", "cores": th, "kind": "port", "sample": oc.describe(k, th) + "; %.1f s for %.2f s of audio" % (dt, a)}
        firsts = [f for f in firsts if f is not None]
        line = {"metric": "audio_sec_per_sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": bench_config(), "engine": {"chunk_parallelism": eng.chunk_parallelism, "chunk_parallelism_short_requests": eng.chunk_parallelism_short, "audio_s_per_step": audio_s / args.steps / world},
                "clocks": clk,
                "e2e": {"value": e2e_val, "unit": "audio-s/s", "h2d_bytes_per_step": int(4 * (WORDS * 6 + WORDS * TOK_PER_WORD * 4)), "d2h_bytes_per_step": int(e2e_bytes / args.steps + 4 * WORDS * TOK_PER_WORD),
                        "first_chunk_ms_p50": statistics.median(firsts) if firsts else None, "rtf": (e2e_ms_max / 1e3) / max(e2e_audio, 1e-9)},
                "gpu_launches": int(launches), "roofline": roof, "kernels": kern}
        if cpu:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_concurrent:
            line["concurrent8"] = await concurrent_leg()
        if world == 1:
            line["conditioning"] = conditioning_leg(eng, cpu is not None)
        if mixed is not None:
            line["mixed_streams"] = mixed
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
        run_reference(a)      # rank 0 only; the other ranks return at once
    else:
        run_b200(a)
