"""Host side of the reference's Llama demo (llm/llama: io.py, tokenizer.py, infer.py, finetune.py) on pydynet_b200 — SURVEY.md
§8(f) row f2: checkpoint / fine-tune parameter IO, the tokenizer and the two command-line drivers, so that the README demo of the
reference runs on the B200 backend with the same files and the same results."""
from .io import load_model, load_finetuned_parameters, save_finetuned_parameters  # noqa: F401
from .tokenizer import Tokenizer  # noqa: F401
