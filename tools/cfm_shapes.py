#!/usr/bin/env python
"""Where do the CFM estimator's GEMM shapes stand?  Device time per launch (graph-replayed, no host cost) of this library's
GEMM against cuBLAS (torch.matmul, bf16) on the same shapes -- cuBLAS is the practical ceiling for a plain GEMM of that shape
on this chip, the tensor peak is not.  M = 2 x calls x T rows of a batched S3Gen call."""
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


def bench(fn, reps=20):
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


rows = [] if sys.argv[1:2] == ["attn"] else [int(x) for x in (sys.argv[1:] or ["916", "7328", "10688", "21376"])]
print(f"{'M':>6} {'N':>5} {'K':>5} | {'ours us':>8} {'TF/s':>7} | {'cublas us':>9} {'TF/s':>7} | layer")
for M in rows:
    for (N, K, name) in [(1536, 256, "qkv"), (256, 512, "attn out"), (1024, 256, "ff0"), (256, 1024, "ff2"), (256, 768, "resnet conv k3"), (256, 960, "resnet c1 (320 x 3)")]:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = torch.randn(N, K, device=dev).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev)
        outb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        wt = w.t().contiguous()

        res = torch.randn(M, N, device=dev)
        resid = name in ("attn out", "ff2")

        def ours():   # the epilogue the model uses for this layer: bf16 out (qkv), GELU + bf16 (ff0), fp32 residual in / out (out, ff2)
            L.check(lib.cbx_op_gemm_ex(a.data_ptr(), w.data_ptr(), bias.data_ptr(), res.data_ptr() if resid else None, out.data_ptr() if resid else None,
                                       None if resid else outb.data_ptr(), M, N, K, 1 if name == "ff0" else 0, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

        def cublas():
            torch.matmul(a, wt, out=outb)
        u1, u2 = bench(ours), bench(cublas)
        fl = 2.0 * M * N * K
        print(f"{M:6d} {N:5d} {K:5d} | {u1:8.2f} {fl / u1 / 1e6:7.1f} | {u2:9.2f} {fl / u2 / 1e6:7.1f} | {name}")
print("fused block tail (cfm_tail_kernel): out-proj + LN3 + GELU-FF + LN1 + QKV, per launch")
for M in rows:
    bf = lambda *s_: torch.randn(*s_, device=dev).mul_(0.05).to(torch.bfloat16)
    o, h, qkv = bf(M, 512), torch.randn(M, 256, device=dev), torch.empty(M, 1536, device=dev, dtype=torch.bfloat16)
    wout, w0, w2, wq = bf(256, 512), bf(1024, 256), bf(256, 1024), bf(1536, 256)
    v256, v1024 = torch.randn(256, device=dev) * 0.1, torch.randn(1024, device=dev) * 0.1
    for mode, name in ((7, "OUT|FF|QKV"), (3, "OUT|FF"), (4, "QKV"), (11, "OUT|FF id"), (27, "id, S/2")):
        fl = 2.0 * M * ((256 * 512 if mode & 1 else 0) + (2 * 256 * 1024 if mode & 2 else 0) + (256 * 1536 if mode & 4 else 0))
        us = bench(lambda: L.check(lib.cbx_op_cfm_tail(mode, M, o.data_ptr(), h.data_ptr(), wout.data_ptr(), v256.data_ptr(), v256.data_ptr(), v256.data_ptr(),
                                                       w0.data_ptr(), v1024.data_ptr(), w2.data_ptr(), v256.data_ptr(), v256.data_ptr(), v256.data_ptr(),
                                                       wq.data_ptr(), qkv.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))))
        tr = (C.c_ulonglong * 16)()
        lib.cbx_cfm_tail_trace(tr)
        d = [int(tr[i]) - int(tr[0]) for i in range(9)]
        print(f"  M={M:6d} {name:10s}: {us:8.2f} us  {fl / us / 1e6:7.1f} TFLOP/s  ({-(-M // 128)} tiles)  ns since start: deps {d[1]}, out-acc {d[2]}, ln3 {d[3]}, gelu {d[4]}, ff-acc {d[5]}, pass2 {d[6]}, qkv {d[7]}, end {d[8]}")
print("attention (tcgen05), 8 heads x 64:")
for (T, B) in [(458, 2), (458, 16), (668, 2), (668, 16), (878, 16), (668, 32), (2048, 16)]:
    qkv = torch.randn(B, T, 3 * 512, device=dev).to(torch.bfloat16)
    o = torch.empty(B, T, 512, device=dev, dtype=torch.bfloat16)
    us = bench(lambda: L.check(lib.cbx_op_attention(qkv.data_ptr(), o.data_ptr(), T, 8, B, 0, C.c_void_p(torch.cuda.current_stream().cuda_stream))))
    print(f"  T={T:4d} B={B:2d}: {us:8.2f} us  {4.0 * T * T * 64 * 8 * B / us / 1e6:7.1f} TFLOP/s")
    if os.environ.get("CBX_ATTN_FA_DBG", "0") == "4":
        tr = (C.c_ulonglong * 256)()
        lib.cbx_attn_fa_trace(tr)
        t0 = int(tr[128 + 1])
        for j in range(min(8, -(-T // 128))):
            for g in range(2):
                i, k = 64 + (j * 2 + g) * 3, 128 + (j * 2 + g) * 4
                print(f"      j={j} g={g}: issuer wait {int(tr[i]) - t0:6d} -> {int(tr[i + 1]) - t0:6d}, issued {int(tr[i + 2]) - t0:6d} | softmax wait {int(tr[k]) - t0:6d} -> {int(tr[k + 1]) - t0:6d}, S in regs {int(tr[k + 2]) - t0:6d}, max {int(tr[192 + (j * 2 + g) * 4]) - t0:6d}, O ok {int(tr[193 + (j * 2 + g) * 4]) - t0:6d}, turn {int(tr[194 + (j * 2 + g) * 4]) - t0:6d}, exp done {int(tr[195 + (j * 2 + g) * 4]) - t0:6d}, P out {int(tr[k + 3]) - t0:6d}")
x = torch.randn(1 << 16, device=dev)
print("empty-ish torch kernel:", bench(lambda: x.add_(1.0)), "us")
