"""ORACLE — TEST INFRASTRUCTURE ONLY.

numpy restatement of the optimizer step the reference runs at stage.py:287-288 (`optimizer.step()` for every registered
optimizer; the examples register `torch.optim.Adam(lr=1e-3)`, examples/mnist.py:39).  The arithmetic lives in torch
(third-party, unpinned in requirements.txt:1; this image: 2.11.0+cu128): torch/optim/adam.py `_single_tensor_adam`
(lines 347ff: weight decay 69-82, moments 110-129, bias corrections 184-189, update 198-200), and
torch/nn/utils/clip_grad.py for the optional clip coefficient (stage.py:276-285).

    t += 1
    g  = coef * grad                     (coef = min(1, max_norm / (||grad||_2 + 1e-6)) when clipping, else 1; -coef: maximize)
    L2 decay:  g += wd * p               decoupled (AdamW):  p *= 1 - lr * wd
    m  = lerp(m, g, 1 - beta1)           v = beta2 * v + (1 - beta2) * g * g
    p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)

Pinned by tests/test_oracle_pins.py against torch.optim.Adam / AdamW themselves (fp64: 1e-12; fp32: a few ulp).
"""
import numpy as np


def clip_coef(sumsq, max_norm):
    """torch.nn.utils.clip_grad_norm_: fp32 arithmetic on the total norm."""
    total = np.float32(np.sqrt(np.float64(sumsq)))
    c = np.float32(max_norm) / (total + np.float32(1e-6))
    return np.float32(min(c, np.float32(1.0)))


def adam_step(p, g, m, v, t, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False, maximize=False,
              coef=1.0, dtype=np.float64):
    """One step.  `t` is the step count AFTER this step (1 for the first).  Returns new (p, m, v) in `dtype`.
    dtype=np.float32 rounds after every operation like an fp32 kernel without fused multiply-adds."""
    f = dtype
    p, g, m, v = (np.asarray(x, dtype=f).copy() for x in (p, g, m, v))
    b1, b2 = f(betas[0]), f(betas[1])
    g = g * f(-coef if maximize else coef)
    if weight_decay != 0:
        if decoupled:
            p = p * f(1.0 - lr * weight_decay)
        else:
            g = g + f(weight_decay) * p
    w = f(1.0 - betas[0])  # python-float expressions first, one rounding to the working dtype (as torch passes scalars)
    m = m + w * (g - m) if w < 0.5 else g - (g - m) * (f(1) - w)
    v = b2 * v + f(1.0 - betas[1]) * g * g
    step_size = f(np.float64(lr) / (1.0 - np.float64(betas[0]) ** t))
    bc2_sqrt = f(np.sqrt(1.0 - np.float64(betas[1]) ** t))
    p = p - step_size * m / (np.sqrt(v) / bc2_sqrt + f(eps))
    return p, m, v


class OracleAdamLib:
    """Stands in for libdmlb behind `dmlb_adam_step_f32`'s C signature on HOST memory (tests inject it through
    FlatAdam(_lib=...) to run the optimizer's layout / checkpoint logic on a CPU-only box).  fp32 arithmetic."""

    def __init__(self):
        self.launches = 0

    @staticmethod
    def _view(ptr, n, ctype):
        import ctypes

        return np.ctypeslib.as_array((ctype * n).from_address(ptr))

    def dmlb_adam_step_f32(self, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, decoupled,
                           maximize, sumsq, max_norm, state, advance, lr_dev, zero_grad, stream):
        import ctypes

        self.launches += 1
        p, g, m, v = (self._view(x, n, ctypes.c_float) for x in (param, grad, exp_avg, exp_avg_sq))
        st = self._view(state, 2, ctypes.c_int64)
        if lr_dev:
            lr = float(self._view(lr_dev, 1, ctypes.c_double)[0])
        coef = 1.0
        if sumsq:
            coef = float(clip_coef(self._view(sumsq, 1, ctypes.c_double)[0], max_norm))
        p2, m2, v2 = adam_step(p, g, m, v, int(st[0]) + 1, lr=lr, betas=(beta1, beta2), eps=eps, weight_decay=weight_decay,
                               decoupled=bool(decoupled), maximize=bool(maximize), coef=coef, dtype=np.float32)
        p[:], m[:], v[:] = p2, m2, v2
        if zero_grad:
            g[:] = 0
        if advance:
            st[0] += 1
        return 0


def sgd_step(p, g, buf, first, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, maximize=False, coef=1.0,
             dtype=np.float64):
    """torch/optim/sgd.py _single_tensor_sgd (what the reference's `optimizer.step()`, stage.py:287-288, runs for the
    ResNet-18 configuration's SGD), with clip_grad_norm_'s coefficient folded in like csrc/optim_kernels.cu K6.
    `first`: no momentum buffer exists yet (torch clones the gradient into it).  Returns (param, momentum_buffer)."""
    f = dtype
    p, g = p.astype(f), g.astype(f) * f(-coef if maximize else coef)
    if weight_decay != 0:
        g = g + f(weight_decay) * p
    if momentum != 0:
        buf = g.copy() if first else f(momentum) * buf.astype(f) + f(1.0 - dampening) * g
        g = g + f(momentum) * buf if nesterov else buf
    return p - f(lr) * g, buf
