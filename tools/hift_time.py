import sys,os,time
ROOT="/root/repo"
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")): sys.path.insert(0,p)
import torch
from cbx_b200.config import ModelConfig
from cbx_b200.native import NativeEngine
from cbx_b200.weights import random_state_dict, synthetic_conditionals
cfg=ModelConfig(); eng=NativeEngine(cfg,max_streams=8,n_lanes=1); eng.load_state_dict(random_state_dict(cfg,0))
conds=synthetic_conditionals(cfg); v=eng.voice_put("default",conds["t3"],conds["gen"])
for n in (35,140,245):
    mel=torch.randn(2*n,80,device="cuda")*1.5-4
    for _ in range(2): eng.hift_infer(mel,seed=1)
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(5): eng.hift_infer(mel,seed=1)
    b.record(); torch.cuda.synchronize(); print("hift n=%d: %.2f ms"%(n,a.elapsed_time(b)/5))
eng.close()
