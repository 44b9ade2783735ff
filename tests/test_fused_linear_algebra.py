"""CPU check of the algebra behind ALPRO_FUSE_TFC (alpro_b200/engine.py): temporal_attn.proj -> DropPath -> temporal_fc
(vit.py:157-161) as one Linear with the composed weight, and the un-composition of its weight gradient.

    y = x_res + fc(s * proj(o)) = x_res + s * (o Wc^T + W_fc b_p) + b_fc,      Wc = W_fc W_p
    dWc = (s * dy)^T o,  v = colsum(s * dy)
    dW_fc = dWc W_p^T + v (x) b_p,   dW_p = W_fc^T dWc,   db_p = W_fc^T v,   db_fc = colsum(dy),   do = (s * dy) Wc

The GPU tests compare the fused kernels with the two-GEMM path; this one pins the formulas themselves against autograd in
float64 so that a future edit of the engine can be checked without a GPU."""
import torch


def test_composed_linear_matches_two_linears_and_their_gradients():
    g = torch.Generator().manual_seed(3)
    M, d = 37, 24
    o = torch.randn(M, d, generator=g, dtype=torch.float64, requires_grad=True)
    x_res = torch.randn(M, d, generator=g, dtype=torch.float64)
    Wp = torch.randn(d, d, generator=g, dtype=torch.float64, requires_grad=True)
    bp = torch.randn(d, generator=g, dtype=torch.float64, requires_grad=True)
    Wf = torch.randn(d, d, generator=g, dtype=torch.float64, requires_grad=True)
    bf = torch.randn(d, generator=g, dtype=torch.float64, requires_grad=True)
    keep = 0.7
    s = (torch.rand(M, generator=g) < keep).double() / keep      # DropPath factor per row (mask / keep_prob)
    dy = torch.randn(M, d, generator=g, dtype=torch.float64)

    # reference: the two Linear maps with DropPath in between, gradients by autograd
    y_ref = x_res + (s[:, None] * (o @ Wp.t() + bp)) @ Wf.t() + bf
    y_ref.backward(dy)

    # composed forward (what the single GEMM + bias / bias2 epilogue computes)
    with torch.no_grad():
        Wc = Wf @ Wp
        c1 = Wf @ bp
        y = x_res + s[:, None] * (o @ Wc.t()) + s[:, None] * c1 + bf
        assert torch.allclose(y, y_ref, rtol=1e-12, atol=1e-12)

        # un-composed backward
        sdy = s[:, None] * dy
        dWc = sdy.t() @ o
        v = sdy.sum(0)
        assert torch.allclose(sdy @ Wc, o.grad, rtol=1e-10, atol=1e-10)                       # d o
        assert torch.allclose(dWc @ Wp.t() + torch.outer(v, bp), Wf.grad, rtol=1e-10, atol=1e-10)
        assert torch.allclose(Wf.t() @ dWc, Wp.grad, rtol=1e-10, atol=1e-10)
        assert torch.allclose(Wf.t() @ v, bp.grad, rtol=1e-10, atol=1e-10)
        assert torch.allclose(dy.sum(0), bf.grad, rtol=1e-10, atol=1e-10)                    # unscaled column sums
