"""The user journey of ``tests/test_demo_flow_cpu.py`` on 2 B200s: bf16, native kernels, fused Hybrid-ZeRO over peer memory,
checkpoint save → auto-resume → HF conversion → ``AutoModelForCausalLM`` load and generate."""
import pytest
import torch

from test_demo_flow_cpu import run_flow

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]


def test_tokenize_train_resume_convert_load_on_gpus(tmp_path):
    run_flow(tmp_path, gpu=True)
