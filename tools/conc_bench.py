#!/usr/bin/env python
"""Aggregate throughput of N concurrent streams on one GPU as a function of the number of S3Gen lanes."""
import asyncio
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")):
    sys.path.insert(0, p)
import torch
import bench
from cbx_b200.engine import TextToSpeechEngine, SamplingDefaults
from cbx_b200.weights import random_state_dict
from cbx_b200.config import ModelConfig

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lanes_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8]
words = int(sys.argv[3]) if len(sys.argv) > 3 else 100
sd = random_state_dict(ModelConfig(), 0)
for lanes in lanes_list:
    eng = TextToSpeechEngine("cuda:0", concurrent_requests=streams, sampling=SamplingDefaults(tokens_per_word=10), seed=0, state_dict=sd,
                             native_kwargs=dict(n_lanes=lanes, max_streams=max(16, streams)))

    async def main():
        await eng.ainit()
        texts = [bench.synthetic_text(words, seed=1234 + i) for i in range(streams)]

        async def one(i):
            nb, t0, first = 0, time.time(), None
            async for chunk in eng.stream(text=texts[i], output_format="raw_pcm", voice_id=None, request_id=f"c{i}", cancellation_token=None, **bench.REQ):
                if first is None and len(chunk):
                    first = (time.time() - t0) * 1e3
                nb += len(chunk)
            return nb, first
        await asyncio.gather(*[one(i) for i in range(streams)])
        torch.cuda.synchronize()
        for w in range(int(os.environ.get("CONC_EXTRA_WAVES", "0"))):     # more waves back to back: first chunks and batch sizes per wave
            nb0 = dict(eng.s3gen.batches)
            r = await asyncio.gather(*[one(i) for i in range(streams)])
            torch.cuda.synchronize()
            print(json.dumps({"wave": w, "first_chunk_ms": sorted(round(x[1], 1) for x in r),
                              "s3gen_batches": {k: v - nb0.get(k, 0) for k, v in sorted(eng.s3gen.batches.items()) if v - nb0.get(k, 0)}}), flush=True)
        b0, r0, n0 = eng.s3gen.busy_s, eng.scheduler.busy_s, dict(eng.s3gen.batches)
        eng.scheduler.row_hist.clear()
        t0 = time.time()
        res = await asyncio.gather(*[one(i) for i in range(streams)])
        torch.cuda.synchronize()
        dt = time.time() - t0
        audio = sum(r[0] for r in res) / 2 / 24000.0
        firsts = sorted(r[1] for r in res)
        print(json.dumps({"streams": streams, "lanes": lanes, "audio_s_per_s": audio / dt, "seconds": dt, "first_chunk_ms_p50": firsts[len(firsts) // 2], "first_chunk_ms_all": [round(f, 1) for f in firsts], "s3gen_batches": dict(sorted(eng.s3gen.batches.items())),
                          "t3_rounds": eng.scheduler.rounds,
                          "s3gen_busy_s": eng.s3gen.busy_s - b0, "t3_busy_s": eng.scheduler.busy_s - r0, "t3_rows_hist": dict(sorted(eng.scheduler.row_hist.items())),
                          "s3gen_batches_this_wave": {k: v - n0.get(k, 0) for k, v in sorted(eng.s3gen.batches.items()) if v - n0.get(k, 0)}}), flush=True)
    asyncio.run(main())
    eng.shutdown()
    del eng
    torch.cuda.empty_cache()
