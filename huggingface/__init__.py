"""HF ``transformers`` model / tokenizer code for checkpoints converted with ``tools/convert2hf.py`` (the reference ships
the same under ``transformers/``; the directory is named ``huggingface`` here so that it can never shadow the installed
``transformers`` package when the repo root is on ``sys.path``)."""
