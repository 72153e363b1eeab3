"""tcar_gemm_tf32 (tcgen05 kind::tf32) against torch fp64 matmul through the C ABI: the three operand-major
combinations the session path uses, multi-segment accumulation, split reduction, fused bias/activation, 3xTF32.

Tolerances: precise (3xTF32) -> 2e-6 of the result scale (fp32 class); fast (TF32) -> 2e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tf32_rn(x):
    """cvt.rna.tf32.f32: round the magnitude to 10 mantissa bits, ties away from zero."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)


def _split(w):
    hi = _tf32_rn(w)
    return hi.contiguous(), _tf32_rn(w - hi).contiguous()


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g) * scale


@pytest.mark.parametrize("M,K,N", [(10240, 500, 250), (512, 128, 250), (512, 500, 500), (77, 320, 320), (1, 64, 250),
                                   (300, 250, 64)])
@pytest.mark.parametrize("precise", [True, False])
def test_forward_kmajor_a_mnmajor_b(native, M, K, N, precise):
    """C = act(A[M,K] . W[K,N] + b): A row-major, W row-major padded to a pitch of 256/512."""
    ldw = (N + 255) // 256 * 256
    lda = (K + 3) // 4 * 4
    A = torch.zeros(M, lda, device="cuda")
    A[:, :K] = _rand(M, K, seed=1)
    W = torch.zeros(K, ldw, device="cuda")
    W[:, :N] = _rand(K, N, seed=2, scale=0.05)
    bias = _rand(N, seed=3)
    hi, lo = _split(W)
    out = torch.full((M, N), float("nan"), device="cuda")
    native.gemm([(A, lda, 0, hi if precise else W, lo if precise else None, ldw, 1, K)], M, N, out, N, bias=bias, act=2,
                precise=precise)
    torch.cuda.synchronize()
    ref = torch.tanh(A[:, :K].double() @ W[:, :N].double() + bias.double())
    err = (out.double() - ref).abs().max().item()
    # fp32-class: a few ulp of the largest sum of |a||w| (what an fp32 FMA chain would give); TF32: 2^-10 of it
    scale = float((A[:, :K].abs() @ W[:, :N].abs()).max())
    assert err < (2e-6 if precise else 1e-3) * scale, f"err {err} scale {scale}"


def test_two_segments_accumulate_like_count_alpha_m(native):
    """U1 = X.W_in + Xc.W_c + D.W_i (modules.py:126-131) as ONE launch.  Xc = X[:, 250:] starts 1000 bytes into a
    row, which TMA cannot address (16-byte rule), so W_c is folded into the bottom half of W_in (X.W_in + Xc.W_c =
    X.(W_in + [0; W_c])) -- the same fold params.py applies when it pre-splits the weights."""
    M = 2000
    X, D = _rand(M, 500, seed=4), _rand(M, 64, seed=5)
    W_in, W_c, W_i = _rand(500, 250, seed=6, scale=0.05), _rand(250, 250, seed=7, scale=0.05), _rand(64, 250, seed=8, scale=0.05)
    Wm = torch.zeros(500, 256, device="cuda")
    Wm[:, :250] = W_in
    Wm[250:, :250] += W_c
    Wi = torch.zeros(64, 256, device="cuda")
    Wi[:, :250] = W_i
    (mh, ml), (ih, il) = _split(Wm), _split(Wi)
    out = torch.full((M, 256), float("nan"), device="cuda")
    native.gemm([(X, 500, 0, mh, ml, 256, 1, 500), (D, 64, 0, ih, il, 256, 1, 64)], M, 250, out, 256, precise=True)
    torch.cuda.synchronize()
    ref = X.double() @ W_in.double() + X[:, 250:].double() @ W_c.double() + D.double() @ W_i.double()
    bound = 4e-6 * float((X.abs() @ Wm[:, :250].abs()).max())
    err = (out[:, :250].double() - ref).abs().max().item()
    assert err < bound, f"err {err} bound {bound}"


@pytest.mark.parametrize("Mred,Kw,N,splits", [(10240, 500, 250, 16), (10240, 64, 250, 8), (512, 500, 500, 1),
                                              (512, 320, 320, 4), (70, 128, 250, 1)])
def test_weight_gradient_both_mnmajor_split(native, Mred, Kw, N, splits):
    """g_W[Kw,N] = X[Mred,Kw]^T . dU[Mred,N]: both operands MN-major, reduction split over CTAs."""
    ldu = (N + 3) // 4 * 4 if N % 4 else N
    ldu = 256 if N == 250 else ldu
    X = _rand(Mred, Kw, seed=9)
    dU = torch.zeros(Mred, ldu, device="cuda")
    dU[:, :N] = _rand(Mred, N, seed=10)
    out = torch.full((Kw, N), float("nan"), device="cuda")
    splits = native.lib().tcar_gemm_tf32_splits(Kw, N, Mred, splits)
    part = torch.zeros(native.lib().tcar_gemm_tf32_part_elems(Kw, N, splits), device="cuda") if splits > 1 else None
    native.gemm([(X, Kw, 1, dU, None, ldu, 1, Mred)], Kw, N, out, N, splits=splits, part=part)
    torch.cuda.synchronize()
    ref = X.double().t() @ dU[:, :N].double()
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 5e-3, f"rel err {err}"


@pytest.mark.parametrize("M,Ko,Ni", [(10240, 250, 250), (10240, 250, 64), (512, 500, 500), (300, 250, 128)])
def test_data_gradient_kmajor_both_accumulate(native, M, Ko, Ni):
    """dX[M,Ni] += dU[M,Ko] . W[Ni,Ko]^T: A K-major, B K-major (W row-major [Ni, pitch])."""
    ldu = (Ko + 255) // 256 * 256
    dU = torch.zeros(M, ldu, device="cuda")
    dU[:, :Ko] = _rand(M, Ko, seed=11)
    W = torch.zeros(Ni, ldu, device="cuda")
    W[:, :Ko] = _rand(Ni, Ko, seed=12, scale=0.05)
    base = _rand(M, Ni, seed=13)
    out = base.clone()
    native.gemm([(dU, ldu, 0, W, None, ldu, 0, Ko)], M, Ni, out, Ni, accumulate=True)
    torch.cuda.synchronize()
    ref = base.double() + dU[:, :Ko].double() @ W[:, :Ko].double().t()
    err = (out.double() - ref).abs().max().item()
    assert err < 1e-2, f"err {err}"


def test_prep_weights_split_is_exact(native):
    theta = _rand(1000 + 64 * 250, seed=14)
    table = torch.tensor([[0, 10, 100, 0, 128, -1, 0], [1000, 64, 250, 10 * 128, 256, -1, 0]], dtype=torch.int32,
                         device="cuda")
    n = 10 * 128 + 64 * 256
    hi, lo = torch.full((n,), float("nan"), device="cuda"), torch.full((n,), float("nan"), device="cuda")
    native.call("tcar_prep_weights", native.ptr(theta), native.ptr(table), 2, native.ptr(hi), native.ptr(lo))
    torch.cuda.synchronize()
    w0 = theta[:1000].view(10, 100)
    got = (hi + lo)[: 10 * 128].view(10, 128)
    assert (got[:, :100] - w0).abs().max() <= w0.abs().max() * 2 ** -21 and (got[:, 100:] == 0).all()
    assert torch.equal(hi[: 10 * 128].view(10, 128)[:, :100], _tf32_rn(w0))
    w1 = theta[1000:].view(64, 250)
    got1 = (hi + lo)[10 * 128:].view(64, 256)
    assert (got1[:, :250] - w1).abs().max() <= w1.abs().max() * 2 ** -21 and (got1[:, 250:] == 0).all()
    assert ((hi.view(torch.int32) & 8191) == 0).all()
