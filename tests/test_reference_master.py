"""Drop-in at the DISPATCHER boundary: the UNMODIFIED reference master (/root/reference/src/master.py: spawn_workers,
result_listener, shutdown_workers) spawns TWO unmodified worker processes (`python -m src.worker <id> <device>`,
src/master.py:56-75), waits for their ready reports (:80-86, the barrier), broadcasts the voice list for cache warming (:88-96)
and routes the pickled result stream back to per-request queues (:29-53) -- with this repository's TextToSpeechEngine behind
every worker (the one-import change of INTEGRATION.md section A, installed in the worker processes by a sitecustomize module so
that no reference file is touched).  The model behind the engines is the deterministic FakeNative (host logic, no GPU): every
request's PCM must be the golden PCM of the unmodified reference engine, the ZeroMQ PUSH socket must have dealt the requests to
both workers, and both must have warmed the broadcast voice.  Needs /root/reference: skipped where it does not exist."""
import asyncio
import os
import pickle
import sys
import zlib

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference checkout not present on this machine")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = np.load(os.path.join(HERE, "golden", "engine_pcm.npz"))

SITE = '''
import os, sys, types
if os.environ.get("CBX_TEST_WORKER_SHIM") == "1" and len(sys.argv) >= 1:
    def _install():
        import cbx_b200.engine as E
        from fake_backend import FakeNative
        logdir = os.environ["CBX_TEST_LOG_DIR"]

        def _log(line):
            with open(os.path.join(logdir, f"{os.getpid()}.log"), "a") as f:
                f.write(line + "\\n")

        class Engine(E.TextToSpeechEngine):          # the worker constructs TextToSpeechEngine(device=...)
            def __init__(self, device):
                super().__init__(device, backend=FakeNative(), concurrent_requests=4)
                _log("engine " + device)

            def stream(self, *a, **kw):
                _log("request " + kw["request_id"])
                return super().stream(*a, **kw)

            def prepare_conditionals(self, path):
                _log("voice " + os.path.basename(path))
                return super().prepare_conditionals(path)

        shim = types.ModuleType("src.tts_streaming")
        shim.TextToSpeechEngine = Engine
        shim.CancellationToken = E.CancellationToken
        sys.modules["src.tts_streaming"] = shim
    _install()
'''


def test_unmodified_reference_master_dispatches_to_two_workers(tmp_path, monkeypatch):
    import zmq
    import zmq.asyncio
    import scipy.io.wavfile as wavfile
    from fake_backend import SCENARIOS, scenario_text
    site, logs, voices = tmp_path / "site", tmp_path / "logs", tmp_path / "voices"
    for d in (site, logs, voices):
        d.mkdir()
    (site / "sitecustomize.py").write_text(SITE)
    wavfile.write(str(voices / "alice.wav"), 24000, (np.sin(np.arange(24000 * 2) * 0.05) * 8000).astype(np.int16))
    monkeypatch.setenv("PYTHONPATH", os.pathsep.join([str(site), os.path.join(ROOT, "chatterbox-tts_b200"), ROOT, HERE, REF]))
    monkeypatch.setenv("CBX_TEST_WORKER_SHIM", "1")
    monkeypatch.setenv("CBX_TEST_LOG_DIR", str(logs))
    monkeypatch.setenv("VOICES_DIR", str(voices))
    monkeypatch.setenv("PRELOADED_VOICES_DIR", str(tmp_path / "none"))
    monkeypatch.setenv("API_KEY", "test-key")
    monkeypatch.setenv("OMP_NUM_THREADS", "2")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import src.ipc as ipc
    import src.master as M                      # the unmodified dispatcher
    monkeypatch.setattr(M.settings, "WORKERS_PER_DEVICE", 2, raising=False)
    monkeypatch.setattr(M.settings, "VOICES_DIR", str(voices), raising=False)
    monkeypatch.setattr(M.settings, "PRELOADED_VOICES_DIR", str(tmp_path / "none"), raising=False)
    monkeypatch.setattr(M.torch.cuda, "is_available", lambda: False)     # "cpu" device list, as on a machine without GPUs (:60-64)
    M.worker_processes.clear(); M.ready_workers.clear(); M.active_requests.clear()
    scs = [s for s in SCENARIOS if s["name"] in ("full_fade30", "long_eos")]
    jobs = [(f"{sc['name']}#{i}", sc) for i in range(2) for sc in scs]          # four requests

    async def run():
        ctx = zmq.asyncio.Context()
        job_push, result_pull, bcast_pub = ipc.setup_master_sockets(ctx)
        listener = asyncio.create_task(M.result_listener(result_pull))
        try:
            M.spawn_workers(bcast_pub)                                           # spawns, then waits for ready and broadcasts the voices
            assert len(M.worker_processes) == 2
            for _ in range(1800):
                if len(M.ready_workers) == 2:
                    break
                assert all(p.poll() is None for p in M.worker_processes), "a worker process died during start-up"
                await asyncio.sleep(0.1)
            assert M.ready_workers == {0, 1}, "ready barrier: both workers must report"

            def log_lines():
                out = {}
                for f in os.listdir(logs):
                    out[f] = open(os.path.join(logs, f)).read().split("\n")
                return out
            for _ in range(100):                                                 # warm-up broadcast reaches BOTH engines
                if sum(any(l == "voice alice.wav" for l in ls) for ls in log_lines().values()) == 2:
                    break
                await asyncio.sleep(0.1)
            assert sum(any(l == "voice alice.wav" for l in ls) for ls in log_lines().values()) == 2
            # requests through the master's own bookkeeping: a queue per request id, the listener fills it
            for rid, sc in jobs:
                M.active_requests[rid] = asyncio.Queue()
                await job_push.send(pickle.dumps(ipc.TTSRequest(
                    request_id=rid, text=scenario_text(sc["words"]), output_format="raw_pcm", voice_id=None, cfg_guidance_weight=0.5,
                    synthesis_temperature=0.8, text_processing_chunk_size=sc["chunk"], audio_tokens_per_slice=sc["slice"],
                    remove_trailing_milliseconds=sc["trail"], remove_leading_milliseconds=sc["lead"],
                    chunk_overlap_strategy=sc["overlap"], crossfade_duration_milliseconds=sc["fade"])))
            for rid, sc in jobs:
                pcm = b""
                while True:
                    msg = await asyncio.wait_for(M.active_requests[rid].get(), 60)
                    assert msg.request_id == rid
                    if msg.is_final:
                        break
                    pcm += msg.chunk
                pcm = np.frombuffer(pcm, dtype=np.int16)
                k = sc["name"]
                assert pcm.shape[0] == int(GOLD[k + "_len"][0]) and zlib.crc32(pcm.tobytes()) == int(GOLD[k + "_crc"][0]), rid
            served = [sum(l.startswith("request ") for l in ls) for ls in log_lines().values()]
            assert sorted(served) == [2, 2], f"the PUSH socket deals requests round-robin to the two workers: {served}"
        finally:
            M.shutdown_workers()
            listener.cancel()
            await asyncio.gather(listener, return_exceptions=True)
            for s in (job_push, result_pull, bcast_pub):
                s.close(linger=0)
            ctx.term()
            M.worker_processes.clear(); M.ready_workers.clear(); M.active_requests.clear()

    asyncio.run(run())
