"""tcgen05 TF32 GEMM building block vs a plain PyTorch fp32 reference of the same op."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(mode, M, N, K, seed=0, single_cta=False, narrow=False, narrow64=False):
    from jsrl_corl_b200 import _lib

    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(seed)
    pad = 8  # exercise leading dimensions larger than the logical width
    pad_k = pad + (-K) % 4  # leading dimensions must stay multiples of 4 floats (TMA row alignment)
    pad_m = pad + (-M) % 4
    pad_n = pad + (-N) % 4
    if mode == 0:
        A = torch.randn(M, K + pad_k, device="cuda", generator=g)[:, :K]
        B = torch.randn(N, K + pad_k, device="cuda", generator=g)[:, :K]
        ref = A.double() @ B.double().T
    elif mode == 1:
        A = torch.randn(M, K + pad_k, device="cuda", generator=g)[:, :K]
        B = torch.randn(K, N + pad_n, device="cuda", generator=g)[:, :N]
        ref = A.double() @ B.double()
    else:
        A = torch.randn(K, M + pad_m, device="cuda", generator=g)[:, :M]
        B = torch.randn(K, N + pad_n, device="cuda", generator=g)[:, :N]
        ref = A.double().T @ B.double()
    Cout = torch.full((M, N + 5), float("nan"), device="cuda")
    scratch = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        rc = L.iql_selftest_umma_gemm(mode | (0x100 if single_cta else 0) | (0x200 if narrow else 0) | (0x400 if narrow64 else 0), M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                                      Cout.data_ptr(), Cout.stride(0), scratch.data_ptr(), scratch.numel(), st.cuda_stream)
    _lib.check(rc, None, "iql_selftest_umma_gemm")
    st.synchronize()
    assert torch.isnan(Cout[:, N:]).all(), "the kernel wrote outside its N columns"
    out = Cout[:, :N].double()
    err = (out - ref).norm() / ref.norm()
    return float(err), out, ref


@pytest.mark.parametrize("single_cta", [False, True], ids=["cta_pair", "single_cta"])
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("shape", [(256, 256, 256), (512, 256, 64), (256, 768, 128), (1024, 512, 96)])
def test_umma_gemm_matches_fp32_matmul(mode, shape, single_cta):
    """N a multiple of 256 runs on CTA pairs (tcgen05 cta_group::2) unless mode | 0x100 asks for the single-CTA kernel;
    the two must agree bit for bit (same products, same accumulation order inside the tensor core)."""
    M, N, K = shape
    err, out, ref = _run(mode, M, N, K, single_cta=single_cta)
    # TF32 operands (10-bit mantissa), fp32 accumulation: ~5e-4 relative per product, averaged over K
    assert err < 1.5e-3, (mode, shape, err)
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("mode,shape", [(0, (256, 256, 23)), (0, (256, 6, 256)), (0, (256, 1, 256)), (0, (256, 256, 69)),
                                        (2, (256, 23, 256)), (2, (256, 37, 256)), (2, (256, 69, 256)), (1, (256, 64, 40))])
def test_umma_gemm_skinny_shapes_and_tails(mode, shape):
    """First-layer (K = obs dims) and output-layer (N = 1 / act_dim) shapes: TMA zero-fill of K tails and of N < 32."""
    M, N, K = shape
    err, out, ref = _run(mode, M, N, K)
    assert err < 1.5e-3, (mode, shape, err)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_umma_cta_pair_equals_single_cta(mode):
    _, pair, _ = _run(mode, 512, 512, 160, seed=3)
    _, single, _ = _run(mode, 512, 512, 160, seed=3, single_cta=True)
    assert torch.equal(pair, single)


@pytest.mark.parametrize("single_cta", [False, True], ids=["cta_pair", "single_cta"])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_umma_narrow_tiles_equal_full_tiles(mode, single_cta):
    """N = 128 tiles (mode | 0x200; on CTA pairs each CTA stages 64 columns of B): the hidden-layer dgrad of launches with
    few problems runs this way.  Same products, same accumulation order per output element: bit-identical to N = 256 tiles."""
    _, narrow, ref = _run(mode, 512, 512, 256, seed=5, single_cta=single_cta, narrow=True)
    _, full, _ = _run(mode, 512, 512, 256, seed=5, single_cta=single_cta)
    assert torch.equal(narrow, full)
    assert float((narrow - ref).norm() / ref.norm()) < 1.5e-3
    _, n64, _ = _run(mode, 512, 512, 256, seed=5, single_cta=single_cta, narrow64=True)
    assert torch.equal(n64, full)


def test_umma_gemm_rejects_bad_shapes():
    from jsrl_corl_b200 import _lib

    L = _lib.lib()
    x = torch.zeros(256, 256, device="cuda")
    s = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    rc = L.iql_selftest_umma_gemm(0, 100, 256, 256, x.data_ptr(), 256, x.data_ptr(), 256, x.data_ptr(), 256,
                                  s.data_ptr(), 4096, None)
    assert rc == _lib.IQL_ERR_INVALID
