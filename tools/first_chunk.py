#!/usr/bin/env python
"""Where the time to the first PCM chunk goes (single request, engine trace)."""
import asyncio, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")):
    sys.path.insert(0, p)
import torch
import bench
from cbx_b200.engine import TextToSpeechEngine, SamplingDefaults

async def main():
    if os.environ.get('FC_GC') == '0':
        import gc; gc.collect(); gc.freeze(); gc.disable()
    from cbx_b200.config import ModelConfig
    from cbx_b200.weights import random_state_dict
    eng = TextToSpeechEngine("cuda:0", concurrent_requests=8, sampling=SamplingDefaults(tokens_per_word=bench.TOK_PER_WORD), seed=0,
                             state_dict=random_state_dict(ModelConfig(), 0))
    await eng.ainit()
    text = bench.synthetic_text(bench.WORDS)
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
        eng._seq = 0
        eng.stats["trace"].clear() if "trace" in eng.stats else None
        t0 = time.time(); first = None
        async for chunk in eng.stream(text=text, output_format="raw_pcm", voice_id=None, request_id="bench", cancellation_token=None, **bench.REQ):
            if first is None and len(chunk):
                first = (time.time() - t0) * 1e3
        if os.environ.get("FC_SYNC", "1") == "1":
            torch.cuda.synchronize()
        tr = list(eng.stats["trace"])
        print(it, "first chunk %.1f ms" % first, "total %.0f ms" % ((time.time() - t0) * 1e3), tr if first > 100 or it == 1 else tr[3:9], flush=True)
    eng.shutdown()
asyncio.run(main())
