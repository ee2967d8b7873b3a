"""GPU parity of K5 (libdmlb dmlb_adam_step_f32 behind dmlcloud_b200.optim.FlatAdam) — the `optimizer.step()` of the
reference's optimise step (stage.py:287-288, examples/mnist.py:39) — against the numpy oracle (oracle/adam_oracle.py,
pinned to torch.optim.Adam / AdamW by tests/test_oracle_pins.py) and against torch.optim.Adam itself on the same device.

Tolerances (fp32 arithmetic, different but equally valid operation orders): 1e-6 * max|x| against the fp64 oracle per
quantity after 6 steps; torch's own fp32 result differs from that oracle by the same order.  Adam's normalised update
m / (sqrt(v) + eps) is ill-conditioned where the effective gradient g + wd * p cancels to ~eps in the first step: with L2
decay a few elements per million land there (the fp32 numpy oracle shows the same elements at the same magnitude, e.g.
5 of 1,000,003 with a largest deviation of 9.5e-6 at lr = 1e-2), so for the parameters up to 2e-5 * n elements may exceed
the 1e-6 bound, by no more than one learning rate.
"""
import numpy as np
import pytest
import torch

from oracle import adam_oracle

pytestmark = pytest.mark.gpu

CONFIGS = [
    dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False, maximize=False),
    dict(lr=1e-2, betas=(0.8, 0.99), eps=1e-6, weight_decay=0.05, decoupled=False, maximize=False),
    dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1, decoupled=True, maximize=True),
    dict(lr=1e-3, betas=(0.3, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False, maximize=False),  # lerp weight > 0.5
]


def _rel(got, want, floor):
    """max |got - want| relative to the largest magnitude in play (`floor`: the scale the quantity moves on, so that a
    handful of near-zero values cannot turn rounding noise into a large relative error)."""
    return float(np.abs(got.astype(np.float64) - want).max() / max(np.abs(want).max(), floor))


@pytest.mark.parametrize('n', [1, 3, 7, 4099, 10_330, 1_000_003])
@pytest.mark.parametrize('cfg', range(len(CONFIGS)))
def test_adam_kernel_vs_oracle(n, cfg):
    """C ABI on raw buffers: 6 steps, vector path (16-byte aligned) and scalar path (views shifted by one element),
    with and without the fused clip coefficient; the device-resident step count advances only when asked to."""
    from dmlcloud_b200 import _native as N

    c = CONFIGS[cfg]
    lib, st = N.cuda_lib(0), N.stream_ptr()
    rng = np.random.RandomState(17 * n + cfg)
    for shift, clip in ((0, None), (1, None), (0, 0.5)):
        P = rng.randn(n).astype(np.float32)
        M, V = np.zeros(n), np.zeros(n)
        Pd = P.astype(np.float64)
        dev = [torch.zeros(n + shift, dtype=torch.float32, device='cuda') for _ in range(4)]
        p, g, m, v = (t[shift:] for t in dev)
        p.copy_(torch.from_numpy(P))
        state = torch.zeros(2, dtype=torch.int64, device='cuda')
        sumsq = torch.zeros(1, dtype=torch.float64, device='cuda')
        for t in range(1, 7):
            G = (rng.randn(n) * (0.05 if t % 2 else 4.0)).astype(np.float32)
            g.copy_(torch.from_numpy(G))
            coef = 1.0
            if clip is not None:
                sumsq.fill_(float((G.astype(np.float64) ** 2).sum()))
                coef = float(adam_oracle.clip_coef(sumsq.item(), clip))
            if t == 3:  # a launch that must NOT advance the step: run it on scratch copies
                scratch = [x.clone() for x in (p, m, v)]
                N.check(lib.dmlb_adam_step_f32(scratch[0].data_ptr(), g.data_ptr(), scratch[1].data_ptr(),
                                               scratch[2].data_ptr(), n, c['lr'], c['betas'][0], c['betas'][1], c['eps'],
                                               c['weight_decay'], int(c['decoupled']), int(c['maximize']), None, 0.0,
                                               state.data_ptr(), 0, None, st), 'adam(no advance)')
                assert int(state[0].item()) == t - 1
            N.check(lib.dmlb_adam_step_f32(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, c['lr'],
                                           c['betas'][0], c['betas'][1], c['eps'], c['weight_decay'], int(c['decoupled']),
                                           int(c['maximize']), sumsq.data_ptr() if clip is not None else None,
                                           clip or 0.0, state.data_ptr(), 1, None, st), 'adam')
            Pd, M, V = adam_oracle.adam_step(Pd, G, M, V, t, lr=c['lr'], betas=c['betas'], eps=c['eps'],
                                             weight_decay=c['weight_decay'], decoupled=c['decoupled'],
                                             maximize=c['maximize'], coef=coef)
        assert int(state[0].item()) == 6 and int(state[1].item()) == 0
        err = np.abs(p.cpu().numpy().astype(np.float64) - Pd)
        bound = 1e-6 * max(np.abs(Pd).max(), 1.0)
        assert int((err > bound).sum()) <= int(2e-5 * n) and err.max() <= c['lr'], (shift, clip, err.max())
        assert _rel(m.cpu().numpy(), M, 0.1) <= 1e-6 and _rel(v.cpu().numpy(), V, 0.01) <= 1e-6, (shift, clip)
        if shift:  # nothing written in front of the shifted views
            assert all(float(t[0]) == 0.0 for t in dev)


def _cnn_params(seed):
    from torch import nn

    torch.manual_seed(seed)
    model = nn.Sequential(nn.Conv2d(1, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Conv2d(16, 16, 3, padding=1),
                          nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(), nn.Linear(784, 10)).cuda()
    return model


def _grads_for(model, step):
    g = torch.Generator(device='cuda').manual_seed(1000 + step)
    return [torch.randn(p.shape, device='cuda', generator=g) * (0.01 if step % 2 else 1.0) for p in model.parameters()]


@pytest.mark.parametrize('decoupled', [False, True])
@pytest.mark.parametrize('flat_grads', [False, True])
def test_flat_adam_matches_torch_adam_on_mnist_cnn(decoupled, flat_grads):
    """The reference's optimizer object (torch.optim.Adam / AdamW) and FlatAdam fed identical gradients for 8 steps:
    per-parameter launches (gradients are separate tensors) and the one-launch path (gradients are views of one flat
    bucket, graphstep.FlatGradBucket)."""
    from dmlcloud_b200 import _native as N
    from dmlcloud_b200.graphstep import FlatGradBucket
    from dmlcloud_b200.optim import FlatAdam

    a, b = _cnn_params(0), _cnn_params(0)
    wd = 0.02
    ref = (torch.optim.AdamW if decoupled else torch.optim.Adam)(a.parameters(), lr=2e-3, weight_decay=wd)
    opt = FlatAdam(b.parameters(), lr=2e-3, weight_decay=wd, decoupled_weight_decay=decoupled)
    assert all(torch.equal(x, y) for x, y in zip(a.parameters(), b.parameters()))  # flattening kept the values
    bucket = FlatGradBucket(list(b.parameters()), torch.device('cuda', 0)) if flat_grads else None
    for step in range(8):
        grads = _grads_for(a, step)
        for p, q, g in zip(a.parameters(), b.parameters(), grads):
            p.grad = g.clone()
            if flat_grads:
                q.grad.copy_(g)
            else:
                q.grad = g.clone()
        before = N.launch_count()
        ref.step()
        opt.step()
        assert N.launch_count() - before == (1 if flat_grads else 6)
    assert opt.steps_taken() == 8
    if flat_grads:
        assert bucket.attached()
    for p, q in zip(a.parameters(), b.parameters()):
        torch.testing.assert_close(q, p, rtol=1e-5, atol=1e-6)
    # the model still computes with the flattened parameters
    x = torch.randn(4, 1, 28, 28, device='cuda')
    torch.testing.assert_close(b(x), a(x), rtol=1e-4, atol=1e-4)


def test_flat_adam_state_dict_is_interchangeable_with_torch_adam():
    from dmlcloud_b200.optim import FlatAdam

    a, b = _cnn_params(1), _cnn_params(1)
    ref = torch.optim.Adam(a.parameters(), lr=1e-3)
    opt = FlatAdam(b.parameters(), lr=1e-3)
    assert opt.state_dict()['state'] == {}  # nothing stepped yet
    for step in range(3):  # torch optimizer runs alone ...
        for p, g in zip(a.parameters(), _grads_for(a, step)):
            p.grad = g
        ref.step()
    with torch.no_grad():
        for p, q in zip(a.parameters(), b.parameters()):
            q.copy_(p)
    opt.load_state_dict(ref.state_dict())  # ... FlatAdam picks up its state
    assert opt.steps_taken() == 3
    for step in range(3, 6):
        for p, q, g in zip(a.parameters(), b.parameters(), _grads_for(a, step)):
            p.grad, q.grad = g.clone(), g.clone()
        ref.step()
        opt.step()
    for p, q in zip(a.parameters(), b.parameters()):
        torch.testing.assert_close(q, p, rtol=1e-5, atol=1e-6)
    # and back: a fresh torch optimizer continues from FlatAdam's checkpoint
    c = _cnn_params(1)
    with torch.no_grad():
        for q, r in zip(b.parameters(), c.parameters()):
            r.copy_(q)
    ref2 = torch.optim.Adam(c.parameters(), lr=1e-3)
    saved = opt.state_dict()
    for group in saved['param_groups']:
        group['capturable'] = False  # FlatAdam is always capturable; the plain torch optimizer keeps `step` on the host
    ref2.load_state_dict(saved)
    for step in range(6, 8):
        for q, r, g in zip(b.parameters(), c.parameters(), _grads_for(a, step)):
            q.grad, r.grad = g.clone(), g.clone()
        opt.step()
        ref2.step()
    for q, r in zip(b.parameters(), c.parameters()):
        torch.testing.assert_close(q, r, rtol=1e-5, atol=1e-6)


def test_flat_adam_in_a_cuda_graph_advances_its_device_step():
    """Captured once, replayed: every replay is one more Adam step (bias corrections follow the device-resident count).
    Bit-identical to stepping eagerly, because it is the same kernel on the same data."""
    from dmlcloud_b200.graphstep import FlatGradBucket
    from dmlcloud_b200.optim import FlatAdam

    a, b = _cnn_params(2), _cnn_params(2)
    eager, graphed = FlatAdam(a.parameters(), lr=1e-2), FlatAdam(b.parameters(), lr=1e-2)
    ga = FlatGradBucket(list(a.parameters()), torch.device('cuda', 0))
    gb = FlatGradBucket(list(b.parameters()), torch.device('cuda', 0))
    gen = torch.Generator(device='cuda').manual_seed(9)
    ga.flat.copy_(torch.randn(ga.total, device='cuda', generator=gen))
    gb.flat.copy_(ga.flat)
    torch.cuda.synchronize()  # the side stream does not order itself after the default stream
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        graphed.step()  # warm-up step 1
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            graphed.step()
        for _ in range(4):
            graph.replay()
    torch.cuda.synchronize()
    for _ in range(5):
        eager.step()
    assert graphed.steps_taken() == 5 and eager.steps_taken() == 5
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.equal(p, q)
