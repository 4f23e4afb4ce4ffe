"""Drop-in for the reference's `discretization.discretization.Discretization` (discretization/discretization.py:9-81).

Same constructor, attributes (`vocabulary.weight`, `size`, `dim`), methods and return values; on CUDA tensors the
nearest-codeword search runs in libschemahead (`sh_dev_discretize`): the [n*bs, size] distance matrix that
`torch.cdist(...).argmin(dim=1)` materialises is never formed, and the codeword gather is fused behind it.
There is no CPU fallback: CPU tensors raise.
"""
import logging
from typing import Sequence, Tuple

import torch
from torch import nn

from schemanet_b200 import native

_LOG = logging.getLogger("discretization")


class Discretization(nn.Module):
    """Visual vocabulary (codebook) of `size` words of dimension `dim`; maps every token to its nearest word.

    detach_input_seq: the tokens are detached before matching (always true for the frozen backbone of the head).
    uniform_range: initialisation interval of the codebook before `initial_vocabulary` loads the k-means result.
    """

    def __init__(self, size: int, dim: int, detach_input_seq: bool = True,
                 uniform_range: Sequence[float] = (-1, 1)):
        super().__init__()
        self.logger = _LOG
        self.size, self.dim = size, dim
        self.detach_input_seq = detach_input_seq
        self.vocabulary = nn.Embedding(size, dim)
        lo, hi = uniform_range
        nn.init.uniform_(self.vocabulary.weight, lo, hi)
        _LOG.info("codebook of %d words x %d dims, U[%.2f, %.2f] init", size, dim, lo, hi)
        self.kernel_mode = native.DISC_AUTO      # SH_DISC_* selector of the CUDA path
        self._activate = True                    # True: the returned sequence is replaced by the matched codewords

    # -- state ----------------------------------------------------------------------------------------------------
    def activate(self):
        self._activate = True

    def deactivate(self):
        self._activate = False

    def initial_vocabulary(self, vocabulary_fp: str):
        """Loads a [n_words, dim] tensor saved by the k-means extraction; a random subset if it has too many rows."""
        words: torch.Tensor = torch.load(vocabulary_fp, map_location="cpu")
        surplus = words.shape[0] - self.size
        if surplus > 0:
            _LOG.warning("external vocabulary has %d words too many: keeping a random subset", surplus)
            words = words[torch.randperm(words.shape[0])[:self.size]]
        with torch.no_grad():
            self.vocabulary.weight.copy_(words)

    # -- matching -------------------------------------------------------------------------------------------------
    def encode(self, seq: torch.Tensor) -> Tuple[torch.Tensor, torch.LongTensor]:
        tokens = seq.detach() if self.detach_input_seq else seq
        n, bs = tokens.shape[0], tokens.shape[1]
        flat = tokens.reshape(n * bs, self.dim)
        codebook = self.vocabulary.weight
        # a trainable codebook keeps the gather as an autograd op (reference :66-67); otherwise it is fused in CUDA
        trainable = torch.is_grad_enabled() and codebook.requires_grad
        fused_out = None
        if self._activate and not trainable:
            fused_out = torch.empty(n * bs, self.dim, dtype=torch.float32, device=flat.device)
        ingredients = native.discretize(flat.detach(), codebook.detach(), out_seq=fused_out, mode=self.kernel_mode)
        if self._activate:
            flat = self.vocabulary(ingredients) if trainable else fused_out
        return flat.reshape(n, bs, self.dim), ingredients.reshape(n, bs)

    def forward(self, seq: torch.Tensor) -> Tuple[torch.Tensor, torch.LongTensor]:
        """seq [n, bs, dim] -> (sequence [n, bs, dim] (codewords if activated), ingredient id of every token [n, bs])."""
        got = int(seq.shape[2])
        assert got == self.dim, f"dimension {got} not match to {self.dim}"
        return self.encode(seq)
