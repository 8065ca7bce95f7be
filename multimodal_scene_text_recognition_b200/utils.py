"""Label codec of the attention decoder (the caller-side step right after the hot path).

``AttnLabelConverter`` keeps the reference's interface (utils.py:102-149): token table
``['[GO]', '[s]', '[PAD]'] + CHARS`` -> ids 0, 1, 2, 3...; ``encode`` pads with [PAD] (id 2) after
[s]; ``decode`` joins the characters of *all* positions and leaves the cut at '[s]' to the caller.
``decode_ids`` is the vectorised version used by this repo's own callers: one D2H copy of the
int32 ids, cut at the first [s] like training_functions.py:188-189 does.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import config


class AttnLabelConverter(object):
    def __init__(self, character: str):
        self.character = ['[GO]', '[s]', '[PAD]'] + list(character)
        self.dict = {ch: i for i, ch in enumerate(self.character)}

    def encode(self, text: Sequence[str], batch_max_length: int = 25):
        """-> (ids [B, batch_max_length + 2] int64 with [GO] first and [PAD] after [s], lengths [B] int32)."""
        lengths = [len(s) + 1 for s in text]
        width = batch_max_length + 2
        batch = torch.full((len(text), width), 2, dtype=torch.long)
        for row, word in enumerate(text):
            ids = [self.dict[ch] for ch in word] + [self.dict['[s]']]
            batch[row, 1:1 + len(ids)] = torch.tensor(ids, dtype=torch.long)
            batch[row, 0] = 0
        dev = config.PRIMARY_DEVICE if torch.cuda.is_available() else 'cpu'
        return batch.to(dev), torch.tensor(lengths, dtype=torch.int32).to(dev)

    def decode(self, text_index, length) -> List[str]:
        rows = text_index.detach().to('cpu').tolist()   # one D2H copy instead of per-element indexing
        return [''.join(self.character[i] for i in rows[k]) for k, _ in enumerate(length)]

    def decode_ids(self, ids) -> List[str]:
        """ids [B, steps] (any int dtype, any device) -> strings cut the way the reference's callers
        cut them: ``pred[:pred.find('[s]')]`` (a missing '[s]' drops the last character)."""
        rows = ids.detach().to('cpu').tolist()
        out = []
        for row in rows:
            s = ''.join(self.character[i] for i in row)
            out.append(s[:s.find('[s]')])
        return out


class Averager(object):
    """Running mean over tensor elements (interface of utils.py:151-172: add / reset / val); the reference's
    training loop imports it next to the converter, so the name has to resolve for training_functions.py."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.n_count, self.sum = 0, 0

    def add(self, v):
        self.n_count += v.data.numel()
        self.sum += v.data.sum()

    def val(self):
        return self.sum / float(self.n_count) if self.n_count else 0
