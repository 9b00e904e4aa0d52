"""The reference's training step with all three losses (``trainer.py:179-231``) on the sm_100a modules.

``ExpModule.training_step`` runs ONE forward and up to three losses -- classification every step, the
self-supervised pair (x 0.1) on SSL epochs (``cur_epoch % ssl_epoch_step == 0``), the 2C2P contrastive
loss (x ``cm_weight``) from ``cm_init_epoch`` on -- each followed by its own ``manual_backward``, and
three ``AdamW`` optimisers that ALL hold ALL model parameters (``main.py:158-160``; only their learning
rates differ).  What that sequence computes:

* every ``opt_x.zero_grad()`` clears the one set of ``.grad`` tensors the three optimisers share
  (``set_to_none``), so when the ``step()`` calls finally run (``trainer.py:225-229``) they all see the
  gradient of the LAST loss that ran backward -- the 2C2P loss when it is active, else the SSL loss,
  else the classification loss -- and the gradients of the earlier losses are never applied;
* parameters the last loss does not reach have ``grad is None`` and are skipped by every optimiser
  (no decay, moments and per-parameter step counts untouched);
* ``opt.step()`` always runs, ``opt_ssl.step()`` / ``opt_cm.step()`` only when their loss was computed,
  each with its own moments, step counts and learning rate, one after the other on the same gradient.

:class:`TrainerStep` reproduces exactly that parameter trajectory (``tests/test_trainer_step_gpu.py``
steps the reference's own modules with three ``torch.optim.AdamW`` beside it).  Because a wiped
gradient is unobservable, the backward passes whose result the reference discards are not run
(``run_wiped_backward=True`` runs them anyway, e.g. to time what the reference pays); the FORWARD of
every active loss always runs -- it moves BatchNorm running statistics (SimSiam projectors, the shared
ProteinCNN under the masked sequence, ``Mean2Embed``).

The lazily created SimSiam projectors (``self_supervised_learning.py:126-141``) do
not exist when ``main.py`` builds the optimisers, so the reference never updates them; here they are
outside the flat parameter store for the same reason (build the TrainerStep before the first SSL call).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import functions as Fn
from . import kernels as K
from .modules import CMTargets, CrossModality, binary_cross_entropy
from .params import FlatAdamW
from .train import StaticBatch


class TrainerStep:
    def __init__(self, model, lr=5e-5, ssl_lr=5e-5, cm_lr=5e-5, weight_decay=1e-2, use_ssl=True, use_cm=True,
                 cm_weight=1.0, run_wiped_backward=False):
        self.model = model
        self.flat = model._flat or model.flatten_parameters()
        mk = lambda rate: FlatAdamW(self.flat, lr=rate, weight_decay=weight_decay, per_param_steps=True)  # noqa: E731
        self.opt = mk(lr)
        self.opt_ssl = mk(ssl_lr) if use_ssl else None
        self.opt_cm = mk(cm_lr) if use_cm else None
        self.cm_weight = cm_weight
        self.run_wiped_backward = run_wiped_backward
        self.losses = {}

    def _backward(self, loss, last: bool, retain: bool) -> None:
        """zero_grad + backward of one loss (``opt_x.zero_grad(); manual_backward(loss)``); a backward
        whose gradients the next zero_grad wipes is skipped unless run_wiped_backward."""
        if not last and not self.run_wiped_backward:
            return
        self.flat.zero_grad()
        with Fn.deferred_weight_grads():
            loss.backward(retain_graph=retain)

    def step(self, sb: StaticBatch, meta=None, compute_ssl: bool = False, compute_cm: bool = False,
             calibrate_cm_weight: bool = False, mlm_mask=None) -> dict:
        """One ``training_step``.  meta: the batch's list of ``{'Prot_ID', 'Drug_ID', 'Y'}`` dicts (or a
        prepared :class:`CMTargets`) when compute_cm; calibrate_cm_weight: the first-2C2P-epoch rescaling
        of ``cm_weight`` by powers of ten (``trainer.py:214-219``; two host reads, like the reference);
        mlm_mask: a pre-sampled ``ssl.sample_mlm_mask`` result (tests)."""
        m = self.model
        compute_ssl = compute_ssl and self.opt_ssl is not None
        compute_cm = compute_cm and self.opt_cm is not None
        K.set_dropout_step(self.opt.step_count)
        try:
            _, _, ssl_input, cm_input, score = m(*sb.model_inputs())
            _, cls_loss = binary_cross_entropy(score, sb.y)
            self._backward(cls_loss, last=not (compute_ssl or compute_cm), retain=compute_ssl or compute_cm)
            self.losses = {"train_loss": cls_loss.detach()}
            total = cls_loss.detach().float()
            if compute_ssl:
                m.ssl_model._mask_override = mlm_mask
                try:
                    d = m.ssl_model(**ssl_input)
                finally:
                    m.ssl_model._mask_override = None
                ssl_loss = (d["prot_ssl"] + d["drug_ssl"]) * 0.1
                self._backward(ssl_loss, last=not compute_cm, retain=compute_cm)
                self.losses["ssl_loss"] = ssl_loss.detach()
                total = total + ssl_loss.detach().float()
            if compute_cm:
                if cm_input is None:
                    raise ValueError("compute_cm needs a model that returns the 2C2P inputs (DrugLAMP2C2P)")
                targets = meta if isinstance(meta, CMTargets) else CrossModality.prepare(meta)
                cm_loss = m.cm_model(**cm_input, meta=targets)
                if calibrate_cm_weight and cm_loss.item() > 0:
                    c, cl = cm_loss.item(), cls_loss.item()
                    while c * self.cm_weight / 10 > cl:
                        self.cm_weight /= 10
                    while c * self.cm_weight * 10 < cl:
                        self.cm_weight *= 10
                cm_loss = cm_loss * self.cm_weight
                self._backward(cm_loss, last=True, retain=False)
                self.losses["cm_loss"] = cm_loss.detach()
                total = total + cm_loss.detach().float()
        finally:
            K.set_dropout_step(None)
        self.losses["all_loss"] = total
        # trainer.py:225-229: the three optimisers step, in this order, on the gradient that is left
        self.opt.step()
        if compute_ssl:
            self.opt_ssl.step()
        if compute_cm:
            self.opt_cm.step()
        return self.losses

    def epoch_flags(self, cur_epoch: int, ssl_epoch_step: int, cm_init_epoch: int):
        """(compute_ssl, compute_cm, calibrate) for a 1-based epoch number (``trainer.py:180,190-191,214``)."""
        return (cur_epoch % ssl_epoch_step == 0 and self.opt_ssl is not None,
                cur_epoch >= cm_init_epoch and self.opt_cm is not None,
                cur_epoch == cm_init_epoch)
