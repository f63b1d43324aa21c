"""Thin Python object over the C-ABI engine.  torch is used only for device buffers and streams;
every computation happens inside libcbx_b200.so."""
import ctypes as C
import os
import threading

import numpy as np
import torch

from . import lib as L
from .config import ModelConfig
from .pack import pack_state_dict

SPEECH_V = 8194


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _f32(a):
    if torch.is_tensor(a):
        a = a.detach().cpu().float().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


_FADE_CACHE = {}
_FADE_LOCK = threading.Lock()


def fade_curves(fade_len: int, device):
    """(fade_in, fade_out) exactly as the reference builds them per request (src/tts_streaming.py:867-871).  The values
    depend only on (length, device), so they are built once and kept: building them per request allocates on the request
    thread's fresh CUDA stream, which misses the caching allocator's per-stream pools and falls through to cudaMalloc --
    measured as 40-90 ms before the request's first T3 launch in one request out of three (tools/first_chunk.py)."""
    key = (int(fade_len), str(device))
    with _FADE_LOCK:
        c = _FADE_CACHE.get(key)
        if c is None:
            t = torch.linspace(0, 1, fade_len, device=device)
            c = (torch.sin(t * 0.5 * torch.pi), torch.cos(t * 0.5 * torch.pi))
            torch.cuda.synchronize(device)      # other streams read these
            if len(_FADE_CACHE) > 64:
                _FADE_CACHE.clear()
            _FADE_CACHE[key] = c
    return c


class NativeEngine:
    """One engine per GPU (reference: one worker process per device, src/master.py:56-77)."""

    def __init__(self, cfg: ModelConfig = None, device: int = 0, max_streams=8, max_seq=1536, max_text=512,
                 max_s3_tokens=1056, max_prompt_tokens=250, n_voices=None, n_lanes=2):
        if not torch.cuda.is_available():
            raise RuntimeError("NativeEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg or ModelConfig()
        if n_voices is None:
            n_voices = int(os.environ.get("CBX_VOICE_SLOTS", "32"))
        self.lib = L.load()
        self.device = device
        torch.cuda.set_device(device)
        f = self.cfg.flow
        self.ccfg = L.CbxConfig(self.cfg.t3.n_layers, f.enc_blocks, f.up_blocks, f.n_blocks, f.n_mid, f.n_timesteps, f.cfg_rate,
                                max_streams, max_seq, max_text, max_s3_tokens, max_prompt_tokens, n_voices, n_lanes)
        h = C.c_void_p()
        L.check(self.lib.cbx_engine_create(C.byref(self.ccfg), device, C.byref(h)))
        self.h = h
        self._voices = {}          # key -> {"slot", "refs", "used"}
        self._retired = []         # records of keys that were re-put / dropped while requests still read the old slot
        self._tick = 0
        self._voice_lock = threading.Lock()
        self.n_voices = n_voices
        self.max_s3_tokens = max_s3_tokens

    # ------------------------------------------------------------------ checkpoint
    def tensor_manifest(self):
        out = []
        name = C.create_string_buffer(256)
        numel, dt = C.c_int64(), C.c_int()
        for i in range(self.lib.cbx_tensor_count(self.h)):
            L.check(self.lib.cbx_tensor_info(self.h, i, name, 256, C.byref(numel), C.byref(dt)))
            out.append((name.value.decode(), numel.value, dt.value))
        return out

    def load_state_dict(self, sd):
        packed = pack_state_dict(sd, self.cfg)
        for name, numel, dt in self.tensor_manifest():
            if name not in packed:
                raise KeyError(f"checkpoint packer produced no tensor named {name}")
            t = packed[name]
            want = torch.float32 if dt == 0 else torch.bfloat16
            if t.dtype != want or t.numel() != numel:
                raise ValueError(f"{name}: packer gave {t.dtype} x{t.numel()}, engine wants {want} x{numel}")
            t = t.contiguous()
            L.check(self.lib.cbx_tensor_upload(self.h, name.encode(), C.c_void_p(t.data_ptr()), t.numel() * t.element_size()))
        L.check(self.lib.cbx_finalize(self.h))

    # ------------------------------------------------------------------ voices
    # The reference's voice_cache is an unbounded dict (src/tts_streaming.py:178); the device cache has `n_voices` slots, so
    # it evicts least-recently-used voices -- but never one that an open request still reads (refcount > 0), and re-putting
    # a key that is in use goes to a NEW slot (the old one is recycled when its last reader releases it).
    def _free_slot_locked(self):
        used = {r["slot"] for r in self._voices.values()} | {r["slot"] for r in self._retired}
        for i in range(self.n_voices):
            if i not in used:
                return i
        idle = [k for k, r in self._voices.items() if r["refs"] == 0 and k != "default"]
        if not idle:
            raise RuntimeError(f"voice cache is full: all {self.n_voices} slots are pinned by requests in flight (raise CBX_VOICE_SLOTS)")
        victim = min(idle, key=lambda k: self._voices[k]["used"])
        slot = self._voices.pop(victim)["slot"]
        L.check(self.lib.cbx_voice_drop(self.h, slot))
        return slot

    def voice_put(self, key, t3_cond: dict, gen: dict) -> int:
        """Caches one voice's conditioning on the device; returns its slot (reference voice_cache, tts_streaming.py:178)."""
        with self._voice_lock:
            rec = self._voices.get(key)
            if rec is not None and rec["refs"] == 0:
                slot = rec["slot"]                       # nobody reads it: rewrite in place
            else:
                if rec is not None:                      # in use: the readers keep the old slot until they release it
                    self._retired.append(self._voices.pop(key))
                slot = self._free_slot_locked()
            spk = _f32(t3_cond["speaker_emb"]).reshape(-1)
            ct = _i32(torch.as_tensor(t3_cond["cond_prompt_speech_tokens"]).cpu().numpy().reshape(-1))
            emo = float(torch.as_tensor(t3_cond["emotion_adv"]).reshape(-1)[0])
            pt = _i32(torch.as_tensor(gen["prompt_token"]).cpu().numpy().reshape(-1))
            pf = _f32(gen["prompt_feat"]).reshape(-1, 80)
            xv = _f32(gen["embedding"]).reshape(-1)
            L.check(self.lib.cbx_voice_put(self.h, slot, spk.ctypes.data, ct.ctypes.data, len(ct), emo, pt.ctypes.data, len(pt),
                                           pf.ctypes.data, pf.shape[0], xv.ctypes.data, _stream_ptr()))
            self._tick += 1
            self._voices[key] = {"slot": slot, "refs": 0, "used": self._tick}
            return slot

    def voice_slot(self, key):
        with self._voice_lock:
            rec = self._voices.get(key)
            return None if rec is None else rec["slot"]

    def voice_acquire(self, key) -> int:
        """Pins the voice for a request (T3 prefills + S3Gen jobs read its device buffers); pair with voice_release(slot)."""
        with self._voice_lock:
            rec = self._voices[key]
            rec["refs"] += 1
            self._tick += 1
            rec["used"] = self._tick
            return rec["slot"]

    def voice_release(self, slot: int):
        with self._voice_lock:
            for rec in self._voices.values():
                if rec["slot"] == slot:
                    rec["refs"] = max(0, rec["refs"] - 1)
                    return
            for rec in self._retired:
                if rec["slot"] == slot:
                    rec["refs"] -= 1
                    if rec["refs"] <= 0:
                        self._retired.remove(rec)
                        L.check(self.lib.cbx_voice_drop(self.h, slot))
                    return

    def voice_drop(self, key):
        with self._voice_lock:
            rec = self._voices.pop(key, None)
            if rec is None:
                return
            if rec["refs"] > 0:
                self._retired.append(rec)                # still read by a request: recycled on its release
            else:
                L.check(self.lib.cbx_voice_drop(self.h, rec["slot"]))

    # ------------------------------------------------------------------ T3
    def t3_open(self, voice, text_ids, cfg_weight=0.5, temperature=0.8, rep_penalty=1.2, min_p=0.05, top_p=0.95, seed=0, max_new=1000):
        ids = _i32(text_ids).reshape(-1)
        slot = C.c_int()
        L.check(self.lib.cbx_t3_open(self.h, voice, ids.ctypes.data, len(ids), cfg_weight, temperature, rep_penalty, min_p, top_p,
                                     seed, max_new, C.byref(slot), _stream_ptr()))
        return slot.value

    def t3_open_batch(self, reqs):
        """reqs: list of (voice, text_ids, cfg_weight, temperature, rep_penalty, min_p, top_p, seed, max_new) -> slots; ONE prefill
        pass for all of them (cbx_t3_open_batch, at most 8), each stream identical to its own t3_open."""
        arr = (L.T3OpenReq * len(reqs))()
        keep = []
        for i, (voice, text_ids, cfg_w, temp, rep, min_p, top_p, seed, max_new) in enumerate(reqs):
            ids = _i32(text_ids).reshape(-1)
            keep.append(ids)
            arr[i] = L.T3OpenReq(voice, ids.ctypes.data, len(ids), cfg_w, temp, rep, min_p, top_p, seed, max_new)
        slots = (C.c_int * len(reqs))()
        L.check(self.lib.cbx_t3_open_batch(self.h, arr, len(reqs), slots, _stream_ptr()))
        return [int(x) for x in slots]

    def t3_step(self, slots, n_steps=1, noise=None):
        s = _i32(slots)
        nptr = C.c_void_p(noise.data_ptr()) if noise is not None else None
        if noise is not None:
            assert noise.is_cuda and noise.dtype == torch.float32 and noise.numel() == n_steps * len(s) * SPEECH_V
        L.check(self.lib.cbx_t3_step(self.h, s.ctypes.data, len(s), n_steps, nptr, _stream_ptr()))

    def t3_set_persistent(self, on: bool):
        L.check(self.lib.cbx_t3_set_persistent(self.h, 1 if on else 0))

    def t3_set_priority(self, high: bool):
        """T3 work on the engine's high- or low-priority stream (include/cbx_b200.h::cbx_t3_set_priority)."""
        L.check(self.lib.cbx_t3_set_priority(self.h, 1 if high else 0))

    def t3_set_alignment_eos(self, on: bool, layer: int = 9):
        """Alignment-based EOS control for the generators opened after this call (include/cbx_b200.h)."""
        L.check(self.lib.cbx_t3_set_alignment_eos(self.h, 1 if on else 0, int(layer)))

    ALIGN_FIELDS = ("on", "i0", "S", "frame_pos", "text_pos", "rows", "started", "started_at", "complete", "completed_at", "has_pre", "ctl", "cur_posn")

    def t3_alignment_peek(self, slot, rows=True):
        """(state dict, newest alignment row, prefilled BOS row) of a stream's analyzer."""
        iv, fv = np.zeros(13, np.int32), np.zeros(6, np.float32)
        L.check(self.lib.cbx_t3_alignment_peek(self.h, slot, iv.ctypes.data, fv.ctypes.data, None, None, _stream_ptr()))
        st = dict(zip(self.ALIGN_FIELDS, (int(v) for v in iv)))
        st.update(first4_max=float(fv[0]), prev_last2=float(fv[1]), tail3=[float(v) for v in fv[2:5]], rep_sum=float(fv[5]))
        if not rows or st["S"] <= 0:
            return st, None, None
        cur, pre = np.zeros(st["S"], np.float32), np.zeros(st["S"], np.float32)
        L.check(self.lib.cbx_t3_alignment_peek(self.h, slot, iv.ctypes.data, fv.ctypes.data, cur.ctypes.data, pre.ctypes.data, _stream_ptr()))
        return st, cur, pre

    def t3_alignment_poke(self, slot, st):
        """Test hook: writes back a state dict obtained from t3_alignment_peek."""
        iv = np.asarray([st[k] for k in self.ALIGN_FIELDS], np.int32)
        fv = np.asarray([st["first4_max"], st["prev_last2"], *st["tail3"], st["rep_sum"]], np.float32)
        L.check(self.lib.cbx_t3_alignment_poke(self.h, slot, iv.ctypes.data, fv.ctypes.data, _stream_ptr()))

    def t3_poll(self, slot):
        n, d = C.c_int(), C.c_int()
        L.check(self.lib.cbx_t3_poll(self.h, slot, C.byref(n), C.byref(d), _stream_ptr()))
        return n.value, bool(d.value)

    def t3_tokens(self, slot, start, count):
        out = np.empty(count, dtype=np.int32)
        L.check(self.lib.cbx_t3_tokens(self.h, slot, start, count, out.ctypes.data, _stream_ptr()))
        return out

    def t3_logits(self, slot):
        out = np.empty((2, SPEECH_V), dtype=np.float32)
        L.check(self.lib.cbx_t3_logits(self.h, slot, out.ctypes.data, _stream_ptr()))
        return out

    def t3_close(self, slot):
        L.check(self.lib.cbx_t3_close(self.h, slot))

    def healthy(self):
        """False once the CUDA context has taken a sticky device fault (include/cbx_b200.h::cbx_engine_health)."""
        return self.lib.cbx_engine_health(self.h) == 0

    def t3_stats(self):
        """(free KV pages, open stream slots)"""
        f, o = C.c_int(), C.c_int()
        L.check(self.lib.cbx_t3_stats(self.h, C.byref(f), C.byref(o)))
        return f.value, o.value

    # ------------------------------------------------------------------ S3Gen
    def s3gen_infer(self, voice, tokens, cache_source=None, seed=0, phase=None, noise=None, return_mel=False):
        tok = _i32(tokens).reshape(-1)
        n = len(tok)
        dev = torch.device("cuda", self.device)
        wav = torch.empty(1, 960 * n, device=dev, dtype=torch.float32)
        src = torch.empty(1, 1, 960 * n, device=dev, dtype=torch.float32)
        mel = torch.empty(2 * n, 80, device=dev, dtype=torch.float32) if return_mel else None
        m = 0 if cache_source is None else cache_source.shape[-1]
        cptr = C.c_void_p(cache_source.contiguous().data_ptr()) if m else None
        ph = _f32(phase) if phase is not None else None
        L.check(self.lib.cbx_s3gen_infer(self.h, voice, tok.ctypes.data, n, cptr, m, C.c_void_p(wav.data_ptr()), C.c_void_p(src.data_ptr()),
                                         C.c_void_p(mel.data_ptr()) if return_mel else None,
                                         ph.ctypes.data if ph is not None else None,
                                         C.c_void_p(noise.data_ptr()) if noise is not None else None, seed, _stream_ptr()))
        return (wav, src, mel) if return_mel else (wav, src)

    def s3gen_infer_batch(self, calls, return_mel=False):
        """calls: list of (voice, tokens, cache_source, seed[, emit_from]) -> list of (wav, source[, mel]); one batched token->mel pass
        for all calls (cbx_s3gen_infer_batch), results equal s3gen_infer call by call.  cache_source is a tensor, None, or
        the index of an EARLIER call of this batch whose source output is the cache (the vocoder runs call by call in
        order, so consecutive slices of one text chunk can share a batch).  emit_from > 0: only wav[..., emit_from:] is valid
        (the vocoder decodes a window; the samples before it are uninitialised)."""
        dev = torch.device("cuda", self.device)
        arr = (L.S3GenCall * len(calls))()
        keep, outs = [], []
        for i, call in enumerate(calls):
            voice, tokens, cache_source, seed = call[:4]
            emit_from = int(call[4]) if len(call) > 4 else 0
            tok = _i32(tokens).reshape(-1)
            n = len(tok)
            wav = torch.empty(1, 960 * n, device=dev, dtype=torch.float32)
            src = torch.empty(1, 1, 960 * n, device=dev, dtype=torch.float32)
            mel = torch.empty(2 * n, 80, device=dev, dtype=torch.float32) if return_mel else None
            if isinstance(cache_source, int):
                assert 0 <= cache_source < i, "a chained cache_source must refer to an earlier call of the batch"
                cs = outs[cache_source][1]
                m = cs.shape[-1]
            else:
                m = 0 if cache_source is None else cache_source.shape[-1]
                cs = cache_source.contiguous() if m else None
            keep.append((tok, cs))
            arr[i] = L.S3GenCall(voice, tok.ctypes.data, n, cs.data_ptr() if m else None, m, wav.data_ptr(), src.data_ptr(),
                                 mel.data_ptr() if return_mel else None, seed, emit_from)
            outs.append((wav, src, mel) if return_mel else (wav, src))
        L.check(self.lib.cbx_s3gen_infer_batch(self.h, arr, len(calls), _stream_ptr()))
        return outs

    def flow_infer(self, voice, tokens):
        tok = _i32(tokens).reshape(-1)
        mel = torch.empty(2 * len(tok), 80, device=torch.device("cuda", self.device), dtype=torch.float32)
        L.check(self.lib.cbx_flow_infer(self.h, voice, tok.ctypes.data, len(tok), C.c_void_p(mel.data_ptr()), _stream_ptr()))
        return mel

    def hift_infer(self, mel, cache_source=None, seed=0, phase=None, noise=None):
        mel = mel.contiguous().float()
        T = mel.shape[0]
        dev = mel.device
        wav = torch.empty(1, 480 * T, device=dev, dtype=torch.float32)
        src = torch.empty(1, 1, 480 * T, device=dev, dtype=torch.float32)
        m = 0 if cache_source is None else cache_source.shape[-1]
        cptr = C.c_void_p(cache_source.contiguous().data_ptr()) if m else None
        ph = _f32(phase) if phase is not None else None
        L.check(self.lib.cbx_hift_infer(self.h, C.c_void_p(mel.data_ptr()), T, cptr, m, C.c_void_p(wav.data_ptr()), C.c_void_p(src.data_ptr()),
                                        ph.ctypes.data if ph is not None else None,
                                        C.c_void_p(noise.data_ptr()) if noise is not None else None, seed, _stream_ptr()))
        return wav, src

    def hift_infer_window(self, mel, w0, cache_source=None, seed=0):
        """hift_infer with the convolution stack over mel frames [w0, T) only: samples before (w0 + 20) * 480 are not valid."""
        mel = mel.contiguous().float()
        T = mel.shape[0]
        wav = torch.zeros(1, 480 * T, device=mel.device, dtype=torch.float32)
        src = torch.empty(1, 1, 480 * T, device=mel.device, dtype=torch.float32)
        m = 0 if cache_source is None else cache_source.shape[-1]
        cptr = C.c_void_p(cache_source.contiguous().data_ptr()) if m else None
        L.check(self.lib.cbx_hift_infer_window(self.h, C.c_void_p(mel.data_ptr()), T, cptr, m, C.c_void_p(wav.data_ptr()), C.c_void_p(src.data_ptr()),
                                               seed, int(w0), _stream_ptr()))
        return wav, src

    def hift_f0(self, mel):
        mel = mel.contiguous().float()
        f0 = torch.empty(mel.shape[0], device=mel.device, dtype=torch.float32)
        L.check(self.lib.cbx_hift_f0(self.h, C.c_void_p(mel.data_ptr()), mel.shape[0], C.c_void_p(f0.data_ptr()), _stream_ptr()))
        return f0

    def hift_source(self, f0, phase=None, noise=None, seed=0):
        f0 = f0.contiguous().float()
        T = f0.shape[0]
        src = torch.empty(1, 1, 480 * T, device=f0.device, dtype=torch.float32)
        ph = _f32(phase) if phase is not None else None
        L.check(self.lib.cbx_hift_source(self.h, C.c_void_p(f0.data_ptr()), T, ph.ctypes.data if ph is not None else None,
                                         C.c_void_p(noise.data_ptr()) if noise is not None else None, seed, C.c_void_p(src.data_ptr()), _stream_ptr()))
        return src

    def crossfade_pcm(self, cur, n_out, prev_tail=None, fade_len=0, out=None, fade_in=None, fade_out=None):
        # `out`: caller-owned int16 device buffer (the engine keeps one per request: an allocation on a request's fresh
        # stream misses the caching allocator's per-stream pools and falls through to cudaMalloc, 4-50 ms with the GPU busy)
        # fade_in / fade_out: the request's curves (reference src/tts_streaming.py:867-871); built here the same way when absent
        out = torch.empty(n_out, device=cur.device, dtype=torch.int16) if out is None else out[:n_out]
        if prev_tail is not None and fade_len > 0 and fade_in is None:
            fade_in, fade_out = fade_curves(fade_len, cur.device)
        L.check(self.lib.cbx_crossfade_pcm(self.h, C.c_void_p(cur.data_ptr()), n_out,
                                           C.c_void_p(prev_tail.data_ptr()) if prev_tail is not None else None, fade_len,
                                           C.c_void_p(fade_in.data_ptr()) if fade_in is not None else None,
                                           C.c_void_p(fade_out.data_ptr()) if fade_out is not None else None,
                                           C.c_void_p(out.data_ptr()), _stream_ptr()))
        return out

    def gpu_launches(self):
        return int(self.lib.cbx_gpu_launches(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.cbx_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
