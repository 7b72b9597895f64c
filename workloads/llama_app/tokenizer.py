"""The demo's tokenizer (reference llm/llama/tokenizer.py:5-66): a llama2.c-style vocabulary with merge scores.

Same results as the reference, including its quirks — characters missing from the vocabulary are dropped, ties between equally
scored merges go to the leftmost pair, duplicate vocabulary strings resolve to their FIRST id, and ``decode`` strips the
CHARACTER SETS ``"<s>"`` and ``"</s>"`` from both ends (so a text ending in "s" loses it). The lookup is a dict built once
instead of ``list.index`` per probe (the reference scans the 32000-entry list for every candidate pair: ~0.2 s per prompt)."""
import json
from typing import List


class Tokenizer:

    def __init__(self, model_path: str):
        with open(model_path, "r", encoding="utf-8") as f:
            model = json.load(f)
        self.vocab = model["tokens"]
        self.scores = model["scores"]
        self.bos_id, self.eos_id = 1, 2
        self._id = {}
        for i, tok in enumerate(self.vocab):
            self._id.setdefault(tok, i)  # first occurrence wins, like list.index

    def str_lookup(self, token: str) -> int:
        return self._id.get(token, -1)

    def encode(self, text: str, add_bos: bool = True, add_eos: bool = False) -> List[int]:
        ids = [i for i in (self._id.get(ch, -1) for ch in text) if i >= 0]
        while True:
            best = None  # (score, position, merged id)
            for pos in range(len(ids) - 1):
                merged = self._id.get(self.vocab[ids[pos]] + self.vocab[ids[pos + 1]], -1)
                if merged != -1 and self.scores[merged] > (best[0] if best else -1e10):
                    best = (self.scores[merged], pos, merged)
            if best is None:
                break
            ids[best[1]:best[1] + 2] = [best[2]]
        if add_bos:
            ids.insert(0, self.bos_id)
        if add_eos:
            ids.append(self.eos_id)
        return ids

    def decode(self, ids: List[int]) -> str:
        return "".join(self.vocab[i] for i in ids).strip("<s>").strip("</s>")
