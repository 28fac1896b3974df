#!/usr/bin/env python
"""Re-shard the MODEL files of a training checkpoint to another tensor x pipeline (or ISP weight-parallel) layout.

    python tools/reshard_ckpt.py --src llm_ckpts/1000 --tgt resharded/1000 --tp 4 --pp 2
    python tools/reshard_ckpt.py --src llm_ckpts/1000 --tgt resharded_isp --tp 2 --wp 4            # ISP files (single stage)

Reads every ``model_tp*_pp*.pt`` (or ``model_tp*_wp*_pp*.pt``) of ``--src`` - any layout - merges them into one state dict
with global layer indices (``tools/ckpt_io.py``) and writes the files of the requested layout, plus ``model_config.pt``.  Start
the new run with ``ckpt = dict(load_ckpt_info=dict(path=<tgt>, content=("model",), ckpt_type="internevo"))``: the optimizer's
fp32 master copy is rebuilt from the loaded weights.  (The reference has no such tool: a checkpoint there is tied to the
tensor / pipeline sizes it was written with, ``internlm/checkpoint/components.py:146-158``.)  Optimizer files are re-sharded over
another ZeRO / data-parallel size by the trainer itself when it loads them; they cannot follow a tensor / pipeline change,
because the reference-compatible file format keeps Adam moments per (tp, pp) coordinate.
"""
import argparse
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ckpt_io  # noqa: E402


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--src", required=True, help="checkpoint folder (one step) with model_tp*_pp*.pt files")
    p.add_argument("--tgt", required=True)
    p.add_argument("--tp", type=int, default=1, help="tensor parallel size of the new layout")
    p.add_argument("--pp", type=int, default=1, help="pipeline parallel size of the new layout")
    p.add_argument("--wp", type=int, default=0, help="> 0: write ISP (weight parallel) files, --tp is then the sequence size")
    p.add_argument("--no_embed_split", action="store_true", help="model.embed_split_hidden=False (vocabulary-parallel embedding)")
    a = p.parse_args(argv)
    cfg = ckpt_io.load_model_config(a.src)
    split_hidden = not a.no_embed_split and bool(cfg.get("embed_split_hidden", True))
    full = ckpt_io.load_full_state(a.src, split_hidden)
    if a.wp > 0:
        assert a.pp == 1, "ISP files are written single-stage"
        ckpt_io.save_sharded_isp(full, a.tgt, a.tp, a.wp)
    else:
        ckpt_io.save_sharded(full, a.tgt, a.tp, split_hidden, pp_size=a.pp)
    for extra in ("model_config.pt", "config_file.pt"):
        if os.path.exists(os.path.join(a.src, extra)):
            shutil.copy(os.path.join(a.src, extra), os.path.join(a.tgt, extra))
    n = sum(v.numel() for v in full.values())
    print(f"{a.src} -> {a.tgt}: {n / 1e6:.1f} M parameters as " + (f"tp{a.tp} x wp{a.wp} (isp)" if a.wp else f"tp{a.tp} x pp{a.pp}"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
