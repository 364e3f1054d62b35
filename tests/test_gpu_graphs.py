"""GraphedStep (rma_net_b200/graphs.py): a captured loss step replays to the same loss and gradients as the eager one,
at the reference's training shape (train_sample.py:15-30: B=4, N=2048, T=7 stages)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _stages(T, B, N, seed):
    g = torch.Generator().manual_seed(seed)
    dev = torch.device("cuda:0")
    st = [(torch.rand(B, 3, N, generator=g) * 2 - 1).to(dev) for _ in range(T)]
    tgt = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
    return st, tgt


def test_graphed_stage_loss_matches_eager():
    from rma_net_b200.graphs import GraphedStep
    from rma_net_b200.model.loss import loss_cd_stages_fused
    T, B, N = 7, 4, 2048
    st, tgt = _stages(T, B, N, 0)
    step = GraphedStep(lambda *s: loss_cd_stages_fused(list(s[:-1]), s[-1], gamma=0.8), st + [tgt],
                       grads_for=list(range(T)))
    for seed in (1, 2):          # new values through the same graph
        st2, tgt2 = _stages(T, B, N, seed)
        loss_g, grads_g = step(*st2, tgt2)
        loss_g = loss_g.clone()
        grads_g = [g.clone() for g in grads_g]
        leaves = [s.clone().requires_grad_(True) for s in st2]
        loss_e = loss_cd_stages_fused(leaves, tgt2, gamma=0.8)
        loss_e.backward()
        assert abs(float(loss_g) - float(loss_e)) <= 1e-6 * abs(float(loss_e))
        for a, b in zip(grads_g, leaves):
            # the scattered half of the gradient is accumulated with float atomics: order-dependent in the last bits
            torch.testing.assert_close(a, b.grad, rtol=1e-5, atol=1e-8)


def test_graphed_step_rejects_shape_change():
    from rma_net_b200.graphs import GraphedStep
    from rma_net_b200.model.loss import loss_on_cd
    st, tgt = _stages(1, 2, 256, 3)
    x = st[0].transpose(1, 2).contiguous()
    step = GraphedStep(lambda a, b: loss_on_cd(a, b), [x, tgt])
    with pytest.raises(RuntimeError):
        step(x[:, :128], tgt)
