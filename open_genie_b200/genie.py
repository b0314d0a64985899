"""Genie — the composed training module (genie/genie.py:18-181), RE-SPECIFIED per SURVEY.md §8 a15.

The reference constructor cannot run at HEAD (it reads undefined attributes such as self.enc_desc / TEST_DESC,
genie.py:37-58) and compute_loss forwards the (quant, idxs) tuple of tokenize() to the dynamics model
(genie.py:109). The intended data flow is kept, with explicit constructor arguments:

    with no_grad: _, tokens = tokenizer.tokenize(video)          # frozen tokenizer
    act_id, act_loss, (rec, q) = latent_action(video)
    dyn_loss = dynamics_model.compute_loss(tokens, act_id[, mask])
    loss = act_loss + dyn_loss                                    # logged as train/act_loss, train/dyn_loss, ...
"""
from __future__ import annotations

from typing import Callable, Iterable

import torch
from torch import Tensor
from torch.optim import Optimizer

from .action import LatentAction
from .dynamics import DynamicsModel
from .lightning_compat import LightningModule
from .optim import FusedAdamW
from .tokenizer import VideoTokenizer
from .utils import Blueprint

OptimizerCallable = Callable[[Iterable], Optimizer]


class Genie(LightningModule):
    def __init__(self, tokenizer: VideoTokenizer, latent_action: LatentAction | dict, dynamics_model: DynamicsModel | dict,
                 optimizer: OptimizerCallable = FusedAdamW, img_prompt: Tensor | None = None):
        super().__init__()
        self.tokenizer = tokenizer.requires_grad_(False)          # pre-trained, frozen (find_unused_parameters=False)
        self.latent_action = latent_action if isinstance(latent_action, LatentAction) else LatentAction(**latent_action)
        self.dynamics_model = (dynamics_model if isinstance(dynamics_model, DynamicsModel)
                               else DynamicsModel(**dynamics_model))
        self.optimizer = optimizer
        self.img_prompt = img_prompt
        self.save_hyperparameters(ignore=['tokenizer', 'latent_action', 'dynamics_model'])

    def compute_loss(self, video: Tensor, mask: Tensor | None = None):
        with torch.no_grad():
            _, tokens = self.tokenizer.tokenize(video)
        act_id, act_loss, (act_rec_loss, act_q_loss) = self.latent_action(video)
        dyn_loss = self.dynamics_model.compute_loss(tokens, act_id, mask=mask)
        loss = act_loss + dyn_loss
        return loss, (('act_loss', act_loss), ('dyn_loss', dyn_loss), ('act_rec_loss', act_rec_loss),
                      ('act_q_loss', act_q_loss))

    def training_step(self, batch: Tensor, batch_idx: int) -> Tensor:
        loss, aux_losses = self.compute_loss(batch)
        self.log_dict({**{'train_loss': loss}, **{f'train/{k}': v for k, v in aux_losses}}, logger=True, on_step=True,
                      sync_dist=True)
        return loss

    def validation_step(self, batch: Tensor, batch_idx: int) -> Tensor:
        loss, aux_losses = self.compute_loss(batch)
        self.log_dict({**{'val_loss': loss}, **{f'val/{k}': v for k, v in aux_losses}}, logger=True, on_step=True,
                      sync_dist=True)
        return loss

    def configure_optimizers(self) -> Optimizer:
        return self.optimizer([p for p in self.parameters() if p.requires_grad])
