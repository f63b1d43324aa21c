"""chatterbox.models.t3.T3: the speech-token decoder behind the generator protocol of T3.inference_stream
(reference call sites src/tts_streaming.py:283-292, :420-435; consumer :499-519, :539, :555)."""
import types

import torch

from ..._backend import new_key, seed_of


class T3(torch.nn.Module):
    """An nn.Module without parameters, so the reference's torch.compile(self.tts.t3, ...) wrapper (:264-265) accepts it;
    attribute access on the compiled wrapper forwards here."""

    def __init__(self, backend):
        super().__init__()
        self._b = backend
        t = backend.cfg.t3
        self.hp = types.SimpleNamespace(start_text_token=t.start_text_token, stop_text_token=t.stop_text_token,
                                        speech_cond_prompt_len=t.speech_cond_prompt_len,
                                        start_speech_token=t.start_speech_token, stop_speech_token=t.stop_speech_token)
        self.steps_per_pull = 7     # decode steps enqueued per device round trip (the consumer pulls 42 tokens per hop, :499-501)

    def inference_stream(self, *, t3_cond, text_tokens, max_new_tokens=1000, temperature=0.8, cfg_weight=0.5,
                         repetition_penalty=1.2, min_p=0.05, top_p=0.95):
        """Synchronous generator of (1, 1) token tensors on the caller's device; ends at the stop-speech token or after
        max_new_tokens.  Closing / dropping the generator (cancel path) releases the stream's KV pages."""
        b = self._b
        nat = b.native
        if not hasattr(t3_cond, "_cbx_key"):
            t3_cond._cbx_key = new_key("t3")
        voice = b.slot_for(t3_cond._cbx_key, t3={"speaker_emb": t3_cond.speaker_emb, "cond_prompt_speech_tokens": t3_cond.cond_prompt_speech_tokens,
                                                 "emotion_adv": t3_cond.emotion_adv})
        ids = torch.as_tensor(text_tokens).reshape(-1, torch.as_tensor(text_tokens).shape[-1])[0].to("cpu").tolist()   # CFG rows are duplicates (:475-478)
        dev = text_tokens.device if torch.is_tensor(text_tokens) else torch.device("cpu")
        seed = seed_of(ids, [int(max_new_tokens)])
        slot = nat.t3_open(voice, ids, float(cfg_weight), float(temperature), float(repetition_penalty), float(min_p), float(top_p), seed, int(max_new_tokens))
        sent = 0
        try:
            while True:
                nat.t3_step([slot], self.steps_per_pull)
                n, done = nat.t3_poll(slot)
                if n > sent:
                    for tok in nat.t3_tokens(slot, sent, n - sent).tolist():
                        sent += 1
                        yield torch.tensor([[int(tok)]], dtype=torch.long, device=dev)
                if done or sent >= max_new_tokens:
                    return
        finally:
            nat.t3_close(slot)
