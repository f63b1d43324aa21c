#!/usr/bin/env python
"""Per-launch device time of single ops replayed from a CUDA graph (no host launch cost)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chatterbox-tts_b200")):
    sys.path.insert(0, p)
import torch
from cbx_b200 import lib as L

lib = L.load()
dev = torch.device("cuda")
REP = 40


def bench(fn, reps=REP):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            g.replay()
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (5 * reps)


def gemm(M, N, K):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = torch.randn(N, K, device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev)
    def fn():
        L.check(lib.cbx_op_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return fn


def attn(T, H, B):
    qkv = torch.randn(B, T, 3 * H * 64, device=dev).to(torch.bfloat16)
    out = torch.empty(B, T, H * 64, device=dev, dtype=torch.bfloat16)
    def fn():
        L.check(lib.cbx_op_attention(qkv.data_ptr(), out.data_ptr(), T, H, B, 0, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return fn


print("tc enabled:", os.environ.get("CBX_DISABLE_TC", "0") != "1")
for (M, N, K) in [(916, 256, 256), (916, 256, 768), (916, 256, 1024), (916, 1536, 256), (916, 1024, 256), (1756, 256, 768), (1756, 1024, 256),
                  (7328, 256, 768), (7328, 1024, 256), (16800, 64, 704), (8192, 8192, 1024)]:
    us = bench(gemm(M, N, K), reps=10 if M * N * K > 1e10 else REP)
    tr = (C.c_ulonglong * 16)()
    lib.cbx_gemm_tc_trace(tr)
    d = [int(tr[i]) - int(tr[0]) for i in range(10)]
    print(f"gemm M={M:6d} N={N:5d} K={K:5d}: {us:8.2f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s  trace_ns(setup,tma0,mma_issued,acc_ready,epi_done,end | tile_written,after_bar,iter0_done)={d[1:7]} | {d[7:10]}")
for (T, H, B) in [(458, 8, 2), (878, 8, 2), (3664, 8, 2)]:
    us = bench(attn(T, H, B))
    print(f"attn T={T:5d} H={H} B={B}: {us:8.2f} us  {4.0 * T * T * 64 * H * B / us / 1e6:8.1f} TFLOP/s")
x = torch.randn(1 << 16, device=dev)
print("empty-ish torch kernel:", bench(lambda: x.add_(1.0)), "us")
