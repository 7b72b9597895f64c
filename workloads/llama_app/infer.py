"""``python -m workloads.llama_app.infer --prompt "There was a boy" [--cuda]`` — the reference's llm/llama/infer.py:11-64 on
pydynet_b200: load the stories15M checkpoint + tokenizer, greedy-decode with the KV cache, print tokens as they are produced and the
reference's throughput line (total length / elapsed, prompt included — the definition bench.py measures at batch 1024)."""
import argparse
import os
import sys
import time

import numpy as np

import pydynet_b200 as pdn
from workloads.llama import Llama

from .io import load_finetuned_parameters, load_model
from .tokenizer import Tokenizer

CONFIG = dict(vocab_size=32000, dim=288, n_heads=6, ffn_dim=768, max_seq_len=1024, n_layers=6)  # infer.py:19-26


def build(data_dir: str, max_batch_size: int = 1, dtype=np.float32, finetuned: str = None):
    c = CONFIG
    model = Llama(c["vocab_size"], c["dim"], c["n_heads"], c["ffn_dim"], c["max_seq_len"], max_batch_size, c["n_layers"], dtype=dtype)
    load_model(model, os.path.join(data_dir, "stories15M.model.npz"))
    if finetuned is not None:
        load_finetuned_parameters(model, finetuned)
    return Tokenizer(os.path.join(data_dir, "tokenizer.model.np")), model


def generate_text(model, tokenizer, prompt: str, max_new_tokens: int, out=sys.stdout):
    """Streams the decoded continuation to ``out``; returns (total token count, elapsed seconds) like infer.py:49-63."""
    input_ids = np.array([tokenizer.encode(prompt)])
    _, L = input_ids.shape
    start = time.time()
    with pdn.no_grad():
        for step_ids in model.generate(input_ids, max_new_tokens):
            L += 1
            output_id = step_ids[0].numpy().tolist()
            if output_id[-1] in (tokenizer.eos_id, tokenizer.bos_id):
                break
            print(tokenizer.decode(output_id), end="", file=out)
            out.flush()
    return L, time.time() - start


def main(argv=None):
    ap = argparse.ArgumentParser(description="Prompt input, e.g. There was a boy")
    ap.add_argument("--prompt", type=str, default="There was a boy")
    ap.add_argument("--cuda", action="store_true")
    ap.add_argument("--finetuned", type=str, default=None, help="Optional finetuned parameter file (.npz)")
    ap.add_argument("--data-dir", type=str, default="llm/llama/data", help="directory holding stories15M.model.npz and tokenizer.model.np")
    ap.add_argument("--max-new-tokens", type=int, default=1024)
    args = ap.parse_args(argv)
    tokenizer, model = build(args.data_dir, finetuned=args.finetuned)
    if args.cuda and pdn.cuda.is_available():
        model = model.to("cuda:0")
    model.eval()
    print(f"\n{args.prompt}", end="")
    L, elapsed = generate_text(model, tokenizer, args.prompt, args.max_new_tokens)
    print(f"\n\nToken count: {L}, elapsed: {elapsed:.2f}s, {round(L / elapsed)} tokens/s")


if __name__ == "__main__":
    main()
