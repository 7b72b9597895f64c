"""``python -m workloads.llama_app.finetune --text "..." [--cuda]`` — the reference's llm/llama/finetune.py:14-79 on pydynet_b200:
teacher-forced next-token cross entropy on one text, Adam over the parameters whose names start with the given prefixes (default
``lm_head``), then the grad-requiring parameters are saved under their ``_parameters`` names."""
import argparse
import time

import numpy as np

import pydynet_b200 as pdn
import pydynet_b200.optim as optim

from .infer import build
from .io import save_finetuned_parameters


def build_causal_training_pair(tokenizer, text: str, max_seq_len: int):
    """[x0..xN-1] -> [x1..xN] over bos + text + eos, truncated to max_seq_len + 1 tokens (finetune.py:14-27)."""
    ids = tokenizer.encode(text, add_bos=True, add_eos=True)
    if len(ids) < 2:
        raise ValueError("Training text is too short after tokenization.")
    ids = ids[:max_seq_len + 1]
    if len(ids) < 2:
        raise ValueError("Token sequence must contain at least 2 tokens.")
    return np.array([ids[:-1]], dtype=np.int64), np.array([ids[1:]], dtype=np.int64)


def finetune(model, tokenizer, text: str, steps: int, lr: float, trainable=("lm_head", ), log=print):
    n_train, n_frozen = model.set_trainable_parameters(tuple(trainable))
    log(f"Trainable params: {n_train}, Frozen params: {n_frozen}")
    optimizer = optim.Adam(model.parameters(), lr=lr)
    input_ids, target_ids = build_causal_training_pair(tokenizer, text, model.max_seq_len)
    losses = []
    for step in range(1, steps + 1):
        losses.append(model.finetune_step(input_ids, target_ids, optimizer))
        if step == 1 or step % 5 == 0 or step == steps:
            log(f"step={step:04d}, loss={losses[-1]:.6f}")
    return losses


def main(argv=None):
    ap = argparse.ArgumentParser(description="Fine-tune Llama parameters")
    ap.add_argument("--text", type=str, required=True)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--cuda", action="store_true")
    ap.add_argument("--trainable", type=str, default="lm_head", help="Comma-separated parameter name prefixes to train")
    ap.add_argument("--save", type=str, default="llm/llama/data/finetuned_params.npz")
    ap.add_argument("--data-dir", type=str, default="llm/llama/data")
    args = ap.parse_args(argv)
    tokenizer, model = build(args.data_dir)
    if args.cuda and pdn.cuda.is_available():
        model = model.to("cuda:0")
    start = time.time()
    finetune(model, tokenizer, args.text, args.steps, args.lr, tuple(p.strip() for p in args.trainable.split(",") if p.strip()))
    elapsed = time.time() - start
    save_finetuned_parameters(model, args.save)
    print(f"Saved finetuned params to {args.save}")
    print(f"Elapsed: {elapsed:.2f}s")


if __name__ == "__main__":
    main()
