"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the reference's TRAINING step on the CPU -- the bar for row f1.

``engine.py:28-35`` (train_one_epoch): ``model.train()``; ``outputs, loss_list = model(inputs, labels)``;
``loss = loss_list[0]``; ``optimizer.zero_grad(); loss.backward(); optimizer.step()``.  This module restates the
forward in train mode (``oracle/backbone.py`` with ``TrainState``: batch-statistics BatchNorm; ``oracle/head.py``
unchanged -- the head has no mode-dependent layer) and takes the gradients from torch autograd, exactly as the
reference does; the arithmetic library (torch) is not vendored under /root/reference.

Nothing in ``scouter_b200`` implements this yet (``SlotModel.forward`` raises in ``.train()``); the golden vectors of
``oracle/make_golden_train.py`` (dumped from the unmodified reference) and this restatement are the parity bar the
CUDA backward will be held to.
"""
from __future__ import annotations

import torch

from .backbone import TrainState, slot_model_forward


def is_parameter(key: str) -> bool:
    """state_dict keys that are nn.Parameters (everything except BatchNorm buffers)."""
    return not key.endswith(("running_mean", "running_var", "num_batches_tracked"))


def train_step(model: str, sd: dict, x: torch.Tensor, target: torch.Tensor, *, num_classes: int, slots_per_class: int,
               loss_status: int = 1, power: int = 1, lambda_value: float = 1.0, dtype=torch.float32):
    """One forward + backward of engine.py:28-33.  Returns a dict: ``log_probs``, ``loss``/``nll``/``attn_loss``,
    ``grads`` {parameter key: gradient, None for parameters the graph does not reach (``slot.to_q.*``, like the
    reference under ``find_unused_parameters=True``, train.py:140)}, ``bn_updates`` {running_mean / running_var key:
    value after the step}."""
    st = TrainState()
    params = {}
    for k, v in sd.items():
        if is_parameter(k) and v.is_floating_point():
            params[k] = v.detach().to(dtype).clone().requires_grad_(True)
            st[k] = params[k]
        else:
            st[k] = v
    out = slot_model_forward(model, st, x, num_classes=num_classes, slots_per_class=slots_per_class,
                             loss_status=loss_status, power=power, lambda_value=lambda_value, target=target, dtype=dtype)
    keys = list(params)
    grads = torch.autograd.grad(out["loss"], [params[k] for k in keys], allow_unused=True)
    return {"log_probs": out["log_probs"].detach(), "loss": out["loss"].detach(), "nll": out["nll"].detach(),
            "attn_loss": out["attn_loss"].detach(), "grads": dict(zip(keys, grads)), "bn_updates": dict(st.updates)}
