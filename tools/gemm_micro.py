"""Micro-benchmark of tcar_gemm_tf32: GPU time per launch (CUDA-graph replay, no host overhead) vs K, for the operand
layouts / modes the session path uses."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.load_package()
from tcar_b200 import _native as nv

def tf32_rn(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)

def graph_time(fn, reps=20, iters=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / (iters * reps) * 1000

dev = "cuda"
for (M, N, precise, bmn) in [(512, 128, False, 0), (512, 128, False, 1), (512, 128, True, 1), (10240, 250, False, 0), (10240, 250, True, 1)]:
    row = []
    for K in (32, 64, 128, 256, 512, 1024):
        lda = K
        ldw = (N + 255) // 256 * 256
        A = torch.randn(M, lda, device=dev)
        if bmn:
            W = torch.zeros(K, ldw, device=dev); W[:, :N] = torch.randn(K, N, device=dev) * 0.05; ldb = ldw
        else:
            W = torch.zeros(N, lda, device=dev); W[:, :K] = torch.randn(N, K, device=dev) * 0.05; ldb = lda
        hi = tf32_rn(W); lo = tf32_rn(W - hi)
        out = torch.zeros(M, ldw, device=dev)
        f = lambda: nv.gemm([(A, lda, 0, hi, lo if precise else None, ldb, bmn, K)], M, N, out, ldw, precise=precise)
        row.append("%5.1f" % graph_time(f))
    print(f"M={M:5d} N={N:3d} precise={int(precise)} b_mn={bmn}:  us/launch for K=32..1024: " + " ".join(row), flush=True)
