"""FlatAdam: Adam / AdamW whose whole step is ONE libdmlb launch over flat fp32 buffers.

The reference calls `optimizer.step()` on whatever torch optimizer the user registered (stage.py:287-288; the examples
use `torch.optim.Adam(lr=1e-3)`, examples/mnist.py:39).  `FlatAdam` is a drop-in for torch.optim.Adam / AdamW on fp32
CUDA parameters:

  * at construction every parameter of a group is moved into one flat fp32 buffer (`p.data` become views, 16-byte
    aligned slots); `exp_avg` / `exp_avg_sq` are flat buffers of the same layout;
  * `step()` is `dmlb_adam_step_f32` (csrc/optim_kernels.cu, 28 B/elem): one launch per group when the gradients are
    views of one flat buffer with the same layout (what `graphstep.FlatGradBucket` sets up), else one launch per
    parameter; the step count lives in device memory, so the optimizer is CUDA-graph capturable by construction;
  * `step(clip=(sumsq, max_norm))` fuses `clip_grad_norm_` (stage.py:276-285): the coefficient is derived on the device
    from a sum of squares that the gradient all-reduce already produced;
  * `state_dict()` / `load_state_dict()` use torch.optim.Adam's format (per-parameter `step`, `exp_avg`, `exp_avg_sq`),
    so checkpoints are interchangeable with the torch optimizer.

Differences from torch.optim.Adam: no amsgrad, no sparse gradients, fp32 CUDA parameters only; the step count is kept
per group, not per parameter (identical unless some parameters receive no gradient in some steps).
"""
import torch

from . import _native as N

SLOT = 4  # elements: every parameter starts on a 16-byte boundary of the flat buffers


class _FlatBase(torch.optim.Optimizer):
    """Layout + device-resident learning rate shared by FlatAdam and FlatSGD."""

    device_lr = True   # the kernels read lr from device memory: a captured graph follows a scheduler (graphstep.py)
    fused_clip = True  # step(clip=(sumsq, max_norm)) applies clip_grad_norm_'s coefficient inside the update launch
    zero_grad_in_step = False  # True (set by the captured step): the update launch also zeroes the gradients it consumed
    EXTRA_STATE = ()   # names of the per-element state buffers besides `param`

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        if hasattr(self, '_flat'):  # groups added after construction get their own flat buffers
            self._flat.append(self._flatten(self.param_groups[-1]))

    def _lib(self, device):
        return self._lib_override if self._lib_override is not None else N.cuda_lib(device.index)

    def _flatten(self, group):
        params = group['params']
        if not params:
            raise ValueError(f'{type(self).__name__}: empty parameter group')
        device = params[0].device
        if device.type != 'cuda' and self._lib_override is None:
            raise RuntimeError(f'{type(self).__name__} runs on libdmlb CUDA kernels: parameters must live on a CUDA '
                               'device (dmlcloud_b200 has no CPU fallback)')
        self._lib(device)
        offsets, total = [], 0
        for p in params:
            if p.dtype != torch.float32 or p.device != device or p.is_sparse:
                raise RuntimeError(f'{type(self).__name__} expects dense fp32 parameters on one CUDA device')
            offsets.append(total)
            total += -(-p.numel() // SLOT) * SLOT
        flat = torch.zeros(total, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p, off in zip(params, offsets):
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        out = {
            'device': device, 'offsets': offsets, 'total': total, 'param': flat,
            'state': torch.zeros(2, dtype=torch.int64, device=device),  # dmlb_adam_state {step, done|pad}
            'lr': torch.full((1,), float(group['lr']), dtype=torch.float64, device=device), 'lr_host': float(group['lr']),
        }
        for name in self.EXTRA_STATE:
            out[name] = torch.zeros_like(flat)
        return out

    def sync_device_lr(self):
        """Copy group['lr'] (what schedulers edit) into the device-resident value the kernels read — one tiny fill per
        group, only when the number changed.  step() calls it; graphstep calls it before a replay."""
        for group, flat in zip(self.param_groups, self._flat):
            lr = float(group['lr'])
            if lr != flat['lr_host']:
                flat['lr'].fill_(lr)
                flat['lr_host'] = lr

    def _attached(self, group, flat):
        """True while every parameter still is the view of the flat buffer it was given at construction."""
        base = flat['param'].data_ptr()
        return all(p.data_ptr() == base + 4 * off for p, off in zip(group['params'], flat['offsets']))

    @staticmethod
    def _flat_grad_base(group, flat):
        """Address of a flat gradient buffer with this group's layout, or None."""
        g0 = group['params'][0].grad
        if g0 is None:
            return None
        base = g0.data_ptr()
        for p, off in zip(group['params'], flat['offsets']):
            g = p.grad
            if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() != base + 4 * off:
                return None
        return base

    def _launch(self, lib, flat, group, offset, grad_ptr, n, clip, advance, st):
        raise NotImplementedError()

    @torch.no_grad()
    def step(self, closure=None, clip=None):
        """clip: None or (sumsq, max_norm) with `sumsq` a 1-element fp64 CUDA tensor holding sum(grad^2) over exactly the
        gradients this optimizer owns — the gradients are scaled by min(1, max_norm / (sqrt(sumsq) + 1e-6)) on the fly."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._lib_override is None and not torch.cuda.is_current_stream_capturing():
            self.sync_device_lr()
        for group, flat in zip(self.param_groups, self._flat):
            if not self._attached(group, flat):
                raise RuntimeError(f'{type(self).__name__}: a parameter no longer aliases the flat buffer (its .data was '
                                   'replaced)')
            lib = self._lib(flat['device'])
            flat['stepped'] = True  # host-side "a step has been issued": state_dict() then has state without a device read
            clip_args = (clip[0].data_ptr(), float(clip[1])) if clip is not None else (None, 0.0)
            st = N.stream_ptr() if self._lib_override is None else None
            base = self._flat_grad_base(group, flat)
            if base is not None:  # one launch for the whole group
                N.check(self._launch(lib, flat, group, 0, base, flat['total'], clip_args, 1, st), 'optimizer_step')
                continue
            live = [(p, off) for p, off in zip(group['params'], flat['offsets']) if p.grad is not None]
            for i, (p, off) in enumerate(live):
                g = p.grad
                if g.is_sparse or g.dtype != torch.float32:
                    raise RuntimeError(f'{type(self).__name__} expects dense fp32 gradients')
                g = g.contiguous()
                N.check(self._launch(lib, flat, group, off, g.data_ptr(), p.numel(), clip_args, int(i == len(live) - 1),
                                     st), 'optimizer_step')
        return loss

    def steps_taken(self, group=0):
        """Host copy of the device-resident step count (synchronises)."""
        return int(self._flat[group]['state'][0].item())


class FlatAdam(_FlatBase):
    EXTRA_STATE = ('exp_avg', 'exp_avg_sq')

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled_weight_decay=False,
                 maximize=False, _lib=None):
        # _lib: test hook (tests/test_host_logic.py injects oracle/adam_oracle.py behind the C signature to exercise the
        # layout / checkpoint logic on a CPU-only box); the product always runs libdmlb on a CUDA device.
        self._lib_override = _lib
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0:
            raise ValueError('lr, eps and weight_decay must be non-negative')
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError(f'Invalid betas: {betas}')
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                        decoupled_weight_decay=decoupled_weight_decay, maximize=maximize,
                        capturable=True,  # the step count is device-resident: always safe to capture
                        # torch.optim.Adam's remaining group keys, so that a state_dict loads into the torch optimizer
                        amsgrad=False, foreach=None, fused=None, differentiable=False)
        super().__init__(params, defaults)
        self._flat = [self._flatten(group) for group in self.param_groups]

    def _launch(self, lib, flat, group, offset, grad_ptr, n, clip, advance, st):
        beta1, beta2 = group['betas']
        sumsq_ptr, max_norm = clip
        o = 4 * offset
        return lib.dmlb_adam_step_f32(flat['param'].data_ptr() + o, grad_ptr, flat['exp_avg'].data_ptr() + o,
                                      flat['exp_avg_sq'].data_ptr() + o, n, float(group['lr']), float(beta1), float(beta2),
                                      float(group['eps']), float(group['weight_decay']),
                                      int(group['decoupled_weight_decay']), int(group['maximize']), sumsq_ptr, max_norm,
                                      flat['state'].data_ptr(), advance,
                                      flat['lr'].data_ptr() if self._lib_override is None else None,
                                      int(self.zero_grad_in_step), st)

    # -- checkpoint: torch.optim.Adam's format -----------------------------------------------------------------------
    def state_dict(self):
        """torch.optim.Adam's format.  No host synchronisation: `step` stays a device tensor (torch's capturable Adam keeps
        it on the device too), so an asynchronous snapshot can stage everything in one go."""
        state, groups, index = {}, [], 0
        for group, flat in zip(self.param_groups, self._flat):
            step = flat['state'][0].to(torch.float32)
            ids = []
            for p, off in zip(group['params'], flat['offsets']):
                n = p.numel()
                if flat.get('stepped'):  # (torch's Adam has no state before its first step either)
                    state[index] = {'step': step.clone(),
                                    'exp_avg': flat['exp_avg'][off:off + n].view(p.shape).clone(),
                                    'exp_avg_sq': flat['exp_avg_sq'][off:off + n].view(p.shape).clone()}
                ids.append(index)
                index += 1
            packed = {k: v for k, v in group.items() if k != 'params'}
            packed['params'] = ids
            groups.append(packed)
        return {'state': state, 'param_groups': groups}

    def load_state_dict(self, state_dict):
        groups = state_dict['param_groups']
        if len(groups) != len(self.param_groups) or any(len(a['params']) != len(b['params'])
                                                        for a, b in zip(groups, self.param_groups)):
            raise ValueError('loaded state dict has a different number of parameter groups / parameters')
        for saved, group, flat in zip(groups, self.param_groups, self._flat):
            for key, value in saved.items():
                if key not in ('params', 'capturable', 'fused', 'foreach', 'differentiable', 'amsgrad'):
                    group[key] = value
            flat['exp_avg'].zero_()
            flat['exp_avg_sq'].zero_()
            steps = set()
            for idx, p, off in zip(saved['params'], group['params'], flat['offsets']):
                entry = state_dict['state'].get(idx)
                if entry is None:
                    continue
                n = p.numel()
                flat['exp_avg'][off:off + n].copy_(entry['exp_avg'].reshape(-1))
                flat['exp_avg_sq'][off:off + n].copy_(entry['exp_avg_sq'].reshape(-1))
                steps.add(int(entry['step']))
            if len(steps) > 1:
                raise ValueError(f'FlatAdam keeps one step count per group; the loaded state has {sorted(steps)}')
            flat['state'].zero_()
            flat['state'][0] = steps.pop() if steps else 0
            flat['stepped'] = bool(state_dict['state'])


class FlatSGD(_FlatBase):
    """torch.optim.SGD (momentum / dampening / nesterov / weight decay) on flat fp32 buffers: `dmlb_sgd_step_f32`
    (csrc/optim_kernels.cu K6), one launch per group on a flat gradient bucket.  Drop-in for the optimizer the ResNet-18
    configuration registers (reference stage.py:287-288 calls `optimizer.step()` on whatever the user registered);
    `state_dict()` uses torch.optim.SGD's format (`momentum_buffer` per parameter)."""

    EXTRA_STATE = ('momentum_buffer',)

    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, maximize=False,
                 _lib=None):
        self._lib_override = _lib
        if lr < 0.0 or momentum < 0.0 or weight_decay < 0.0:
            raise ValueError('lr, momentum and weight_decay must be non-negative')
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError('Nesterov momentum requires a momentum and zero dampening')
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov,
                        maximize=maximize, capturable=True, foreach=None, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self._flat = [self._flatten(group) for group in self.param_groups]

    def _launch(self, lib, flat, group, offset, grad_ptr, n, clip, advance, st):
        sumsq_ptr, max_norm = clip
        o = 4 * offset
        buf = flat['momentum_buffer'].data_ptr() + o if group['momentum'] != 0 else None
        return lib.dmlb_sgd_step_f32(flat['param'].data_ptr() + o, grad_ptr, buf, n, float(group['lr']),
                                     float(group['momentum']), float(group['dampening']), float(group['weight_decay']),
                                     int(group['nesterov']), int(group['maximize']), sumsq_ptr, max_norm,
                                     flat['state'].data_ptr(), advance,
                                     flat['lr'].data_ptr() if self._lib_override is None else None,
                                     int(self.zero_grad_in_step), st)

    def state_dict(self):
        state, groups, index = {}, [], 0
        for group, flat in zip(self.param_groups, self._flat):
            stepped = bool(flat.get('stepped'))
            ids = []
            for p, off in zip(group['params'], flat['offsets']):
                if stepped and group['momentum'] != 0:
                    state[index] = {'momentum_buffer': flat['momentum_buffer'][off:off + p.numel()].view(p.shape).clone()}
                ids.append(index)
                index += 1
            packed = {k: v for k, v in group.items() if k != 'params'}
            packed['params'] = ids
            packed['_flat_steps'] = flat['state'][0].clone()  # (torch's SGD keeps no step count; K6 needs "first step?")
            groups.append(packed)
        return {'state': state, 'param_groups': groups}

    def load_state_dict(self, state_dict):
        groups = state_dict['param_groups']
        if len(groups) != len(self.param_groups) or any(len(a['params']) != len(b['params'])
                                                        for a, b in zip(groups, self.param_groups)):
            raise ValueError('loaded state dict has a different number of parameter groups / parameters')
        for saved, group, flat in zip(groups, self.param_groups, self._flat):
            for key, value in saved.items():
                if key not in ('params', 'capturable', 'fused', 'foreach', 'differentiable', '_flat_steps'):
                    group[key] = value
            flat['momentum_buffer'].zero_()
            seen = False
            for idx, p, off in zip(saved['params'], group['params'], flat['offsets']):
                entry = state_dict['state'].get(idx)
                if entry is None or entry.get('momentum_buffer') is None:
                    continue
                flat['momentum_buffer'][off:off + p.numel()].copy_(entry['momentum_buffer'].reshape(-1))
                seen = True
            flat['state'].zero_()
            flat['state'][0] = int(saved.get('_flat_steps', 1 if seen else 0))
            flat['stepped'] = seen or int(flat['state'][0]) > 0
