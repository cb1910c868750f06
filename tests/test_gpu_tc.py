"""tcgen05 / TMA building blocks: every shared-memory operand flavour used by the tensor-core kernels,
checked against torch on exactly representable bf16 inputs (fp32 accumulation -> tight tolerance)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(a_mode, b_mode, N, K):
    from clip_finegrained_alignment_b200 import _lib
    g = torch.Generator().manual_seed(100 * a_mode + 10 * b_mode + N + K)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16)          # logical A[m, k]
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)            # logical B[n, k]
    ref = A.float() @ B.float().t()
    a_src = (A.t().contiguous() if a_mode == 2 else A).cuda()
    b_src = (B.t().contiguous() if b_mode in (1, 3) else B).cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    _lib.call("cfa_tc_selftest", a_mode, b_mode, N, K, a_src.data_ptr(), b_src.data_ptr(), D.data_ptr(),
              _lib.stream_ptr())
    torch.cuda.synchronize()
    return D.cpu(), ref


@pytest.mark.parametrize("a_mode,b_mode,N,K", [
    (0, 0, 208, 128),    # TMA SW128 K-major x TMA SW128 K-major        (S = l . v^T)
    (0, 0, 80, 64),
    (1, 0, 208, 128),    # interleaved K-major A x TMA K-major B          (L += G_kb . l_kb^T, dW += dG_kb . v_kb^T)
    (1, 1, 64, 208),     # interleaved K-major A x TMA tile read MN-major (G_kb = W . v_kb, X = dL . l_kb)
    (1, 1, 64, 80),
    (2, 1, 64, 80),      # interleaved read MN-major A x TMA MN-major B   (dv_kb = dS^T . l_kb)
    (2, 3, 64, 80),      # both interleaved, both MN-major                (dv_kb += W^T . dG_kb)
    (1, 2, 96, 64),
    (1, 3, 64, 208),     # interleaved K-major A x interleaved MN-major B (dl_kb += dL^T . G_kb uses a_mode 2)
])
def test_tcgen05_operand_modes(a_mode, b_mode, N, K):
    D, ref = _run(a_mode, b_mode, N, K)
    err = (D - ref).abs().max().item()
    assert torch.isfinite(D).all(), (a_mode, b_mode)
    assert err <= 1e-3 * max(1.0, ref.abs().max().item()), (a_mode, b_mode, N, K, err)
