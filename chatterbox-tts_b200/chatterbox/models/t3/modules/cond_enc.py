"""chatterbox.models.t3.modules.cond_enc.T3Cond (reference call site src/tts_streaming.py:377-381)."""
import torch


class T3Cond:
    def __init__(self, speaker_emb=None, cond_prompt_speech_tokens=None, emotion_adv=0.5, clap_emb=None, cond_prompt_speech_emb=None):
        self.speaker_emb = speaker_emb
        self.cond_prompt_speech_tokens = cond_prompt_speech_tokens
        self.emotion_adv = emotion_adv
        self.clap_emb = clap_emb
        self.cond_prompt_speech_emb = cond_prompt_speech_emb

    def to(self, *, device=None, dtype=None):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                is_fp = v.is_floating_point()
                setattr(self, k, v.to(device=device, dtype=dtype if is_fp else None))
        return self
