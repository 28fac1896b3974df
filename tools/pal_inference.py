"""PAL (program-aided language model) evaluation: the model writes a Python ``solution()`` for every math word problem, the
program is executed in a sandboxed runtime with a time limit, and its return value is compared with the gold answer.

Counterpart of the reference's ``tools/pal_inference.py`` (GSM8K + PAL prompting).  There is no network here, so the
dataset is a local JSON-lines file with ``question`` / ``answer`` fields in the GSM8K format (the gold number follows
``####``); results are written as JSON lines and can be resumed with ``--append``.

    python tools/pal_inference.py /path/to/hf_model results.jsonl --dataset gsm8k_test.jsonl --max_length 1024
"""
from __future__ import annotations

import argparse
import copy
import json
import multiprocessing as mp
import os
import re
import sys
from typing import Any, Dict, List, Optional

PROMPT_HEAD = '''Let's use python to solve math problems. Here are some examples of how to do it,

Q: Olivia has $23. She bought five bagels for $3 each. How much money does she have left?
```python
def solution():
    """Olivia has $23. She bought five bagels for $3 each. How much money does she have left?"""
    money_initial = 23
    bagels = 5
    bagel_cost = 3
    money_spent = bagels * bagel_cost
    money_left = money_initial - money_spent
    result = money_left
    return result
```

Q: There were nine computers in the server room. Five more computers were installed each day, from monday to thursday. How many computers are now in the server room?
```python
def solution():
    """There were nine computers in the server room. Five more computers were installed each day, from monday to thursday. How many computers are now in the server room?"""
    computers_initial = 9
    computers_per_day = 5
    num_days = 4
    computers_added = computers_per_day * num_days
    computers_total = computers_initial + computers_added
    result = computers_total
    return result
```

How about this question?
Q: {question}'''


class GenericRuntime:
    """Executes generated code in its own namespace; ``answer_expr`` is evaluated afterwards."""

    GLOBAL_DICT: Dict[str, Any] = {}
    HEADERS: List[str] = ["import math"]

    def __init__(self):
        self._global_vars = copy.copy(self.GLOBAL_DICT)
        for h in self.HEADERS:
            self.exec_code(h)

    def exec_code(self, code_piece: str) -> None:
        exec(code_piece, self._global_vars)  # noqa: S102 - this IS the point of PAL; run under `run_with_timeout`

    def eval_code(self, expr: str) -> Any:
        return eval(expr, self._global_vars)  # noqa: S307

    def inject(self, var_dict: Dict[str, Any]) -> None:
        self._global_vars.update(var_dict)


def extract_code(generation: str) -> List[str]:
    """The first fenced python block of the generation (or the raw text when the model did not fence it)."""
    m = re.search(r"```(?:python)?\n(.*?)```", generation, re.S)
    body = m.group(1) if m else generation.split("```")[0]
    return body.rstrip().split("\n")


def _worker(code: List[str], answer_expr: str, q):
    try:
        rt = GenericRuntime()
        rt.exec_code("\n".join(code))
        q.put(("ok", rt.eval_code(answer_expr)))
    except BaseException as e:  # noqa: B902 - the generated program may raise anything, including SystemExit
        q.put(("err", repr(e)))


def run_with_timeout(code: List[str], answer_expr: str = "solution()", time_out: float = 10.0):
    """Run the program in a child process; returns ``(status, value)`` with status ``ok | err | timeout``."""
    ctx = mp.get_context("fork" if sys.platform != "win32" else "spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(code, answer_expr, q))
    p.start()
    p.join(time_out)
    if p.is_alive():
        p.terminate()
        p.join()
        return "timeout", None
    return q.get() if not q.empty() else ("err", "no result")


def gold_answer(answer_field: str) -> Optional[float]:
    m = re.search(r"####\s*(-?[\d,\.]+)", answer_field)
    return float(m.group(1).replace(",", "")) if m else None


def is_correct(pred: Any, gold: Optional[float], tol: float = 1e-3) -> bool:
    try:
        return gold is not None and abs(float(pred) - gold) < tol
    except (TypeError, ValueError):
        return False


class PALInterface:
    """prompt → generation → program → answer.  ``generate_fn(prompt) -> str`` abstracts the model."""

    def __init__(self, generate_fn, answer_expr: str = "solution()", time_out: float = 10.0, verbose: bool = False):
        self.generate_fn, self.answer_expr, self.time_out, self.verbose = generate_fn, answer_expr, time_out, verbose
        self.history: List[str] = []

    def run(self, question: str):
        gen = self.generate_fn(PROMPT_HEAD.format(question=question))
        self.history.append(gen)
        code = extract_code(gen)
        status, value = run_with_timeout(code, self.answer_expr, self.time_out)
        if self.verbose:
            print(gen, "\n->", status, value)
        return (value if status == "ok" else None), gen, status


def load_model(path: str, max_length: int):
    import torch
    from transformers import AutoModelForCausalLM, AutoTokenizer

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from interface import GenerationConfig, generate

    tok = AutoTokenizer.from_pretrained(path, trust_remote_code=True)
    model = AutoModelForCausalLM.from_pretrained(path, trust_remote_code=True, torch_dtype=torch.bfloat16)
    model = model.cuda().eval() if torch.cuda.is_available() else model.float().eval()
    cfg = GenerationConfig(max_length=max_length, do_sample=False)
    return lambda prompt: generate(model, tok, prompt, cfg, stop_fn=lambda t: t.count("```") >= 2)


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("model", help="HF folder written by tools/convert2hf.py")
    p.add_argument("out", help="result file (JSON lines)")
    p.add_argument("--dataset", required=True, help="local JSONL with `question` / `answer` (GSM8K format)")
    p.add_argument("--max_length", type=int, default=2048)
    p.add_argument("--time_out", type=float, default=10.0)
    p.add_argument("--verbose", "-v", action="store_true")
    p.add_argument("--append", "-a", action="store_true", help="continue an existing result file")
    a = p.parse_args(argv)
    data = [json.loads(line) for line in open(a.dataset) if line.strip()]
    done = sum(1 for _ in open(a.out)) if a.append and os.path.exists(a.out) else 0
    pal = PALInterface(load_model(a.model, a.max_length), time_out=a.time_out, verbose=a.verbose)
    correct = total = 0
    with open(a.out, "a" if a.append else "w") as f:
        for ex in data[done:]:
            pred, gen, status = pal.run(ex["question"])
            ok = is_correct(pred, gold_answer(ex["answer"]))
            correct, total = correct + int(ok), total + 1
            f.write(json.dumps({"question": ex["question"], "generation": gen, "status": status, "prediction": pred,
                                "gold": gold_answer(ex["answer"]), "correct": ok}, default=str) + "\n")
            f.flush()
    print(f"accuracy {correct}/{total} = {correct / max(total, 1):.4f}")


if __name__ == "__main__":
    main()
