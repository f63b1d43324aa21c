"""Shared state of the shim: one NativeEngine per process and the voice-slot bookkeeping.

The native voice cache holds T3 and S3Gen conditioning together (cbx_voice_put); the reference hands them to the model
separately (t3_cond to T3.inference_stream, ref_dict to S3Gen.inference), so each object gets its own slot: a T3Cond's slot
carries a minimal S3Gen part and vice versa.  Slots are keyed by a token stored on the object and recycled LRU."""
import collections
import itertools
import os
import threading
import zlib

import numpy as np
import torch

_factory = None
_lock = threading.Lock()
_ids = itertools.count(1)


def set_backend_factory(fn):
    """Tests inject a factory returning an object with NativeEngine's interface; the product path builds a NativeEngine."""
    global _factory
    _factory = fn


class Backend:
    def __init__(self, ckpt_dir: str, device: str):
        from cbx_b200.config import ModelConfig
        self.cfg = ModelConfig()
        self.device = device
        self.lock = threading.Lock()
        self.slots = collections.OrderedDict()      # key -> native voice slot name (LRU order)
        self.weights_source, self.encoder_sd = "injected", None
        if _factory is not None:
            self.native = _factory(ckpt_dir, device)
            self.encoder_sd = getattr(self.native, "encoder_sd", None)      # tests may hang conditioning-encoder weights on the injected engine
        else:
            if "cuda" not in str(device):
                raise RuntimeError("the B200-native chatterbox shim needs a cuda:N device; there is no CPU fallback")
            from cbx_b200.native import NativeEngine
            from cbx_b200.checkpoint import load_checkpoint
            gpu = int(str(device).split(":")[-1]) if ":" in str(device) else 0
            n = int(os.environ.get("CONCURRENT_REQUESTS_PER_WORKER", "1"))
            # merged / upstream checkpoint files; a missing checkpoint raises like from_local does, unless
            # CBX_ALLOW_RANDOM_WEIGHTS=1 (BASELINE.json configs: random-init weights)
            sd, self.encoder_sd, self.weights_source = load_checkpoint(ckpt_dir or "", self.cfg, 0)
            self.native = NativeEngine(self.cfg, device=gpu, max_streams=max(8, n), n_lanes=max(2, min(n, 4)))
            self.native.load_state_dict(sd)
            if os.environ.get("CBX_ALIGNMENT_EOS", "0") == "1":    # the package's AlignmentStreamAnalyzer, opt-in
                self.native.t3_set_alignment_eos(True, int(os.environ.get("CBX_ALIGNMENT_LAYER", "9")))
        self.max_slots = max(2, getattr(self.native, "n_voices", 16) - 1)

    def _dummy_gen(self):
        fc = self.cfg.flow
        return {"prompt_token": torch.zeros(1, 1, dtype=torch.long), "prompt_token_len": torch.tensor([1]),
                "prompt_feat": torch.zeros(1, 2, fc.mel), "prompt_feat_len": None, "embedding": torch.zeros(1, fc.spk_dim)}

    def _dummy_t3(self):
        t = self.cfg.t3
        spk = torch.zeros(1, t.speaker_embed_size)
        spk[0, 0] = 1.0
        return {"speaker_emb": spk, "cond_prompt_speech_tokens": torch.zeros(1, 1, dtype=torch.long), "emotion_adv": torch.zeros(1, 1, 1)}

    def slot_for(self, key: str, t3: dict = None, gen: dict = None) -> int:
        with self.lock:
            if key in self.slots:
                self.slots.move_to_end(key)
                return self.native.voice_slot(key) if hasattr(self.native, "voice_slot") else 0
            while len(self.slots) >= self.max_slots:
                old, _ = self.slots.popitem(last=False)
                self.native.voice_drop(old)
            slot = self.native.voice_put(key, t3 if t3 is not None else self._dummy_t3(), gen if gen is not None else self._dummy_gen())
            self.slots[key] = slot
            return slot


    def encoders(self, what: str):
        """The GPU conditioning encoders (cbx_b200/conditioning.py), built on first use from the checkpoint's tokenizer.* /
        speaker_encoder.* / ve.* tensors.  A checkpoint without them is an error: a voice is never invented."""
        with self.lock:
            if getattr(self, "_enc", None) is None:
                if _factory is not None and not self.encoder_sd:
                    raise RuntimeError(f"{what}: the injected test backend has no conditioning-encoder weights")
                if not self.encoder_sd:
                    raise RuntimeError(f"{what}: the checkpoint holds no conditioning-encoder weights (tokenizer.*, speaker_encoder.*, ve.*)")
                from cbx_b200.conditioning import ConditioningEncoders
                gpu = int(str(self.device).split(":")[-1]) if ":" in str(self.device) else 0
                self._enc = ConditioningEncoders(self.encoder_sd, self.cfg.cond, device=gpu)
            return self._enc


def new_key(prefix: str) -> str:
    return f"{prefix}-{next(_ids)}"


def seed_of(*parts) -> int:
    h = 0
    for p in parts:
        h = zlib.crc32(np.asarray(p).tobytes(), h)
    return h & 0x7FFFFFFF
