"""Drop-in for the reference's `discretization.discretization.Discretization` (discretization/discretization.py:9-81).

Same constructor, attributes (`vocabulary.weight`, `size`, `dim`), methods and return values; on CUDA tensors the
nearest-codeword search runs in libschemahead (`sh_dev_discretize`): the [n*bs, size] distance matrix that
`torch.cdist(...).argmin(dim=1)` materialises is never formed, and the codeword gather is fused behind it.
There is no CPU fallback: CPU tensors raise.
"""
import logging
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from schemanet_b200 import native


class Discretization(nn.Module):
    def __init__(
        self,
        size: int,
        dim: int,
        detach_input_seq: bool = True,
        uniform_range: Tuple[float, float] = [-1, 1]
    ):
        super().__init__()
        self.logger = logging.getLogger("discretization")
        self.size = size
        self.dim = dim
        self.detach_input_seq = detach_input_seq
        self.logger.info("Creating discretization with size: %d, dimension: %d", size, dim)
        self.vocabulary = nn.Embedding(size, dim)
        self._reset_parameters(uniform_range)
        self.kernel_mode = native.DISC_AUTO
        self.activate()

    def _reset_parameters(self, uniform_range: Tuple[float, float]):
        self.logger.info("Initializing with Uniform[%.2f, %.2f]", uniform_range[0], uniform_range[1])
        nn.init.uniform_(self.vocabulary.weight, uniform_range[0], uniform_range[1])

    def initial_vocabulary(self, vocabulary_fp: str):
        self.logger.info("Loading from external vocabulary...")
        vocabulary: torch.Tensor = torch.load(vocabulary_fp, map_location="cpu")
        if vocabulary.shape[0] > self.size:
            self.logger.warning("Too much external vocabulary, using random picked vocabulary...")
            vocabulary = vocabulary[torch.randperm(vocabulary.shape[0])][:self.size]
        with torch.no_grad():
            self.vocabulary.weight.copy_(vocabulary)

    def deactivate(self):
        self.logger.debug("Deactivated discretization!")
        self._activate = False

    def activate(self):
        self.logger.debug("Activated discretization!")
        self._activate = True

    def encode(self, seq: torch.Tensor) -> Tuple[torch.Tensor, torch.LongTensor]:
        if self.detach_input_seq:
            seq = seq.detach()
        n, bs = seq.shape[:2]
        flat = seq.reshape(n * bs, self.dim)
        weight = self.vocabulary.weight
        train_vocab = torch.is_grad_enabled() and weight.requires_grad
        gathered = None
        if self._activate and not train_vocab:
            gathered = torch.empty(n * bs, self.dim, dtype=torch.float32, device=flat.device)
        ingredients = native.discretize(flat.detach(), weight.detach(), out_seq=gathered, mode=self.kernel_mode)
        if self._activate:
            # with a trainable vocabulary the gather stays an autograd op, as in the reference (:66-67)
            flat = self.vocabulary(ingredients) if train_vocab else gathered
        return flat.reshape(n, bs, self.dim), ingredients.reshape(n, bs)

    def forward(self, seq: torch.Tensor) -> Tuple[torch.Tensor, torch.LongTensor]:
        """seq [n, bs, dim] -> (encoded sequence [n, bs, dim], matched ingredient of every token [n, bs])."""
        t = seq.shape[2]
        assert int(t) == self.dim, f"dimension {seq.shape[2]} not match to {self.dim}"
        return self.encode(seq)
