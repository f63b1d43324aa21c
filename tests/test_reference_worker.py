"""Drop-in at the worker boundary: the UNMODIFIED reference worker (/root/reference/src/worker.py: handle_request,
listen_for_jobs, listen_for_broadcasts) runs over real ZeroMQ sockets against a fake master, with this repository's
TextToSpeechEngine swapped in by the one-import change INTEGRATION.md section A describes (here: a `src.tts_streaming`
module object that re-exports cbx_b200.engine's classes).  Pickled TTSRequest in -> TTSStreamChunk stream out -> final
marker; cancel_request / clear_voice_cache / warm_up_voices broadcasts honoured (reference src/worker.py:22-58, 104-136,
src/ipc.py:25-59).  The model behind the engine is the deterministic FakeNative (host logic only, no GPU); the PCM must be
the golden PCM of the unmodified reference engine.  Needs /root/reference: skipped where it does not exist."""
import asyncio
import os
import pickle
import sys
import types
import zlib

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference checkout not present on this machine")

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "engine_pcm.npz"))


def _import_reference_worker():
    from fake_backend import FakeNative
    import cbx_b200.engine as E
    os.environ.setdefault("API_KEY", "test-key")
    if REF not in sys.path:
        sys.path.insert(0, REF)

    class Engine(E.TextToSpeechEngine):          # the worker constructs TextToSpeechEngine(device=...)
        def __init__(self, device):
            super().__init__(device, backend=FakeNative(), concurrent_requests=4)

    shim = types.ModuleType("src.tts_streaming")  # the one-import change of INTEGRATION.md, done without touching the file
    shim.TextToSpeechEngine = Engine
    shim.CancellationToken = E.CancellationToken
    import src  # noqa: F401  (reference package)
    sys.modules["src.tts_streaming"] = shim
    sys.modules.pop("src.worker", None)
    import src.worker as W
    import src.ipc as ipc
    return W, ipc, Engine


def _request(ipc, rid, sc, text):
    return ipc.TTSRequest(request_id=rid, text=text, output_format="raw_pcm", voice_id=None, cfg_guidance_weight=0.5,
                          synthesis_temperature=0.8, text_processing_chunk_size=sc["chunk"], audio_tokens_per_slice=sc["slice"],
                          remove_trailing_milliseconds=sc["trail"], remove_leading_milliseconds=sc["lead"],
                          chunk_overlap_strategy=sc["overlap"], crossfade_duration_milliseconds=sc["fade"])


def test_unmodified_reference_worker_streams_our_engine_over_zeromq(tmp_path):
    import zmq
    import zmq.asyncio
    from fake_backend import SCENARIOS, scenario_text
    W, ipc, Engine = _import_reference_worker()
    sc = next(s for s in SCENARIOS if s["name"] == "full_fade30")
    long_sc = next(s for s in SCENARIOS if s["name"] == "long_eos")

    async def run():
        ctx = zmq.asyncio.Context()
        job_push, result_pull, bcast_pub = ipc.setup_master_sockets(ctx)            # the fake master binds :5555/:5556/:5557
        engine = Engine(device="cpu")
        await engine.ainit()
        job_socket, result_socket, broadcast_socket = ipc.setup_worker_sockets(ctx)   # exactly what worker.main() does
        await result_socket.send(pickle.dumps(ipc.WorkerStatus(worker_id=0, status="ready")))
        tasks = [asyncio.create_task(W.listen_for_jobs(engine, job_socket, result_socket)),
                 asyncio.create_task(W.listen_for_broadcasts(engine, broadcast_socket))]
        try:
            st = pickle.loads(await asyncio.wait_for(result_pull.recv(), 10))
            assert isinstance(st, ipc.WorkerStatus) and st.status == "ready"
            await asyncio.sleep(0.3)                                                 # let the SUB socket finish subscribing

            async def collect(rid, stop_after=None):
                chunks, final = [], False
                while not final:
                    msg = pickle.loads(await asyncio.wait_for(result_pull.recv(), 30))
                    assert isinstance(msg, ipc.TTSStreamChunk) and msg.request_id == rid
                    final = msg.is_final
                    if not final:
                        chunks.append(msg.chunk)
                        if stop_after is not None and len(chunks) == stop_after:
                            await bcast_pub.send(pickle.dumps(ipc.BroadcastCommand("cancel_request", {"request_id": rid})))
                return b"".join(chunks)

            # 1. job in -> chunks out -> final marker; PCM equals the unmodified reference engine's
            await job_push.send(pickle.dumps(_request(ipc, sc["name"], sc, scenario_text(sc["words"]))))
            pcm = np.frombuffer(await collect(sc["name"]), dtype=np.int16)
            k = sc["name"]
            assert pcm.shape[0] == int(GOLD[k + "_len"][0]) and zlib.crc32(pcm.tobytes()) == int(GOLD[k + "_crc"][0])
            assert np.array_equal(pcm[:4000], GOLD[k + "_head"])
            # 2. cancel_request broadcast stops a long stream early, the final marker still arrives
            await job_push.send(pickle.dumps(_request(ipc, "to-cancel", long_sc, scenario_text(long_sc["words"]))))
            part = np.frombuffer(await collect("to-cancel", stop_after=1), dtype=np.int16)
            assert 0 < part.shape[0] < int(GOLD["long_eos_len"][0])
            # 3. warm_up_voices / clear_voice_cache broadcasts reach the engine's voice cache
            wav = tmp_path / "alice.wav"
            import scipy.io.wavfile as wavfile
            wavfile.write(str(wav), 24000, (np.sin(np.arange(24000 * 2) * 0.05) * 8000).astype(np.int16))
            engine.voice_manager.voices_dir = str(tmp_path)
            await bcast_pub.send(pickle.dumps(ipc.BroadcastCommand("warm_up_voices", {"voice_ids": ["alice.wav"]})))
            for _ in range(100):
                if "alice.wav" in engine.voice_cache:
                    break
                await asyncio.sleep(0.05)
            assert "alice.wav" in engine.voice_cache
            await bcast_pub.send(pickle.dumps(ipc.BroadcastCommand("clear_voice_cache", {"voice_id": "alice.wav"})))
            for _ in range(100):
                if "alice.wav" not in engine.voice_cache:
                    break
                await asyncio.sleep(0.05)
            assert "alice.wav" not in engine.voice_cache
        finally:
            for t in tasks:
                t.cancel()
            await asyncio.gather(*tasks, return_exceptions=True)
            engine.shutdown()
            for s in (job_push, result_pull, bcast_pub, job_socket, result_socket, broadcast_socket):
                s.close(linger=0)
            ctx.term()

    asyncio.run(run())
