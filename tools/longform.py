#!/usr/bin/env python
"""BASELINE.json configs[4] on one GPU: a 2000-word document with VOICE-CLONING CONDITIONING COMPUTED FOR THE REQUEST (a 10 s clip in
VOICES_DIR, not cached: prepare_conditionals runs the GPU encoders inside the timed request) streamed through the engine (about 80
text chunks, 20 000 speech tokens, ~13 minutes of audio): also checks slot / KV-page recycling and ordering at length."""
import asyncio, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench
from cbx_b200.engine import TextToSpeechEngine, SamplingDefaults

async def main():
    words = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    import tempfile
    from scipy.io import wavfile
    from cbx_b200.config import ModelConfig
    from cbx_b200.weights import random_state_dict
    vdir = tempfile.mkdtemp()
    os.environ["VOICES_DIR"] = vdir
    t = np.arange(24000 * 10) / 24000.0
    clip = 0.3 * np.sin(2 * np.pi * 170 * t) * np.clip(np.sin(2 * np.pi * 1.3 * t), 0, None) + 0.03 * np.random.default_rng(0).standard_normal(t.shape[0])
    wavfile.write(os.path.join(vdir, "speaker.wav"), 24000, (clip * 32767).astype(np.int16))
    cfg = ModelConfig()
    eng = TextToSpeechEngine("cuda:0", concurrent_requests=8, sampling=SamplingDefaults(tokens_per_word=bench.TOK_PER_WORD), seed=0,
                             state_dict=random_state_dict(cfg, 0), encoder_state_dict=random_state_dict(cfg, 0, parts=("cond",)))
    await eng.ainit()
    eng.conditioning_encoders()          # weights resident (the reference loads its encoders in from_local as well)
    text = bench.synthetic_text(words)
    # steady-state serving: one short request first (allocator pools, first-use sizes), as bench.py's warm-up steps do
    async for _ in eng.stream(text=bench.synthetic_text(100, seed=7), output_format="raw_pcm", voice_id=None, request_id="warm", cancellation_token=None, **bench.REQ):
        pass
    torch.cuda.synchronize()
    t0 = time.time(); first = None; nbytes = 0; peak = 0
    async for chunk in eng.stream(text=text, output_format="raw_pcm", voice_id="speaker.wav", request_id="long", cancellation_token=None, **bench.REQ):
        if first is None and len(chunk):
            first = (time.time() - t0) * 1e3
        if len(chunk):
            peak = max(peak, int(np.abs(np.frombuffer(chunk, dtype=np.int16)).max()))
        nbytes += len(chunk)
    torch.cuda.synchronize()
    dt = time.time() - t0
    audio = nbytes / 2 / 24000.0
    print({"words": words, "audio_s": round(audio, 1), "seconds": round(dt, 2), "audio_s_per_s": round(audio / dt, 1), "first_chunk_ms": round(first, 1),
           "peak_int16": peak, "batches": dict(sorted(eng.s3gen.batches.items())), "mem_GB": round(torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9, 1)})
    eng.shutdown()
asyncio.run(main())
