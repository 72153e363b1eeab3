"""Micro-benchmark of tcar_gemm_tf32 GPU time per launch via CUDA graphs (no host overhead):
back-to-back identical launches vs alternating with a tiny small-smem kernel (carveout reconfiguration)."""
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
small = torch.zeros(1024, device=dev)
print("tiny kernel alone: %.2f us" % graph_time(lambda: small.add_(1)))
for (M, K, N, precise, bmn, label) in [(512, 128, 250, True, 1, "h1"), (512, 500, 500, True, 1, "a_ic"), (10240, 564, 250, True, 1, "U1"),
                                       (10240, 820, 250, True, 1, "U2"), (512, 250, 128, False, 0, "dCT"), (10240, 250, 64, False, 0, "dD"),
                                       (10240, 250, 250, False, 0, "dXi"), (512, 32, 64, False, 0, "1kb")]:
    lda = (K + 3) // 4 * 4
    ldw = (N + 255) // 256 * 256
    A = torch.randn(M, lda, device=dev)
    if bmn:
        W = torch.zeros(K, ldw, device=dev); W[:, :N] = torch.randn(K, N, device=dev) * 0.05
        ldb = ldw
    else:
        W = torch.zeros(N, lda, device=dev); W[:, :K] = torch.randn(N, K, device=dev) * 0.05
        ldb = lda
    hi = tf32_rn(W); lo = tf32_rn(W - hi)
    out = torch.zeros(M, ldw, device=dev)
    f = lambda: nv.gemm([(A, lda, 0, hi, lo if precise else None, ldb, bmn, K)], M, N, out, ldw, precise=precise)
    t1 = graph_time(f)
    t2 = graph_time(lambda: (f(), small.add_(1)))
    print(f"{label:5s} M={M} K={K} N={N} precise={precise}: back-to-back {t1:.1f} us/launch; alternating with tiny kernel {t2:.1f} us/pair", flush=True)
