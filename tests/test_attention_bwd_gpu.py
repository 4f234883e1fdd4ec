"""hupr_attention_bwd (fused flash-style backward of layers.py:126-133, head dim 64) against float64 autograd of the same attention.
Tolerance 1e-4 of max |gradient| (fp32-equivalent hi/lo arithmetic; P = exp(S - lse) with ex2.approx, fp32 accumulation in TMEM,
atomic accumulation of dQ across key blocks)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,s", [(2, 256), (1, 1024)])
def test_attention_bwd_matches_autograd(b, s):
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(51)
    c, ld = 64, 256
    qk = torch.randn(b, s, ld, device="cuda") * 0.5                    # q at channel offset 64, k at 128 of one projection tensor
    v = torch.randn(b, s, c, device="cuda")
    do_wide = torch.randn(b, s, 192, device="cuda")                    # dO at channel offset 64 of a wider gradient tensor
    QK, V, DO = SplitTensor.from_float(qk), SplitTensor.from_float(v), SplitTensor.from_float(do_wide)
    q = QK.float()[:, :, 64:128].double().requires_grad_()
    k = QK.float()[:, :, 128:192].double().requires_grad_()
    vv = V.float().double().requires_grad_()
    do = DO.float()[:, :, 64:128].double()
    logits = q @ k.transpose(1, 2)
    p = torch.softmax(logits, dim=2)
    out = p @ vv
    out.backward(do)
    lse = torch.logsumexp(logits.detach(), dim=2).float().contiguous()
    rowdot = (do * out.detach()).sum(dim=2).float().contiguous()
    dq = torch.zeros(b, s, c, device="cuda")
    dk = torch.zeros(b, s, 128, device="cuda")                         # written at channel offset 64
    dv = torch.full((b, s, c), 0.5, device="cuda")                     # accumulates on top of earlier contents
    ops.attention_bwd(QK, 64, QK, 128, V, 0, DO, 64, lse, rowdot, dq, 0, dk, 64, dv, 0)
    torch.cuda.synchronize()

    def rel(got, ref):
        return float((got.double() - ref).abs().max() / ref.abs().max())
    e_q, e_k, e_v = rel(dq, q.grad), rel(dk[:, :, 64:], k.grad), rel(dv - 0.5, vv.grad)
    print("attention_bwd rel err dq %.3g dk %.3g dv %.3g" % (e_q, e_k, e_v))
    assert float(dk[:, :, :64].abs().max()) == 0.0
    assert e_v < 1e-4 and e_k < 1e-4 and e_q < 1e-4, (e_q, e_k, e_v)
