"""Builds a small model of one family with the UNMODIFIED reference on CPU (``use_flash_attn=False``: its torch attention / rotary
/ norm code), runs one forward and saves weights, inputs and logits (see ``test_reference_differential_cpu.py``).  The reference
has no CPU mode: the accelerator's RNG / device hooks are pointed at the CPU generator, a one-rank gloo group stands in for
every parallel mode, nothing in its model or op code is touched.

    python differential_model_probe.py <reference root> <model type> <output .pt> [<.pt with "state" and "ids" to use instead>]
"""
import os
import sys

root, family, dst = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, root)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import internlm  # noqa: F401,E402
import internlm.utils.common as common  # noqa: E402
from internlm.accelerator import get_accelerator  # noqa: E402
from internlm.core.context import ParallelMode  # noqa: E402
from internlm.core.context import global_context as gpc  # noqa: E402
from internlm.core.context.parallel_context import Config  # noqa: E402

cpu, orig = torch.device("cpu"), common.get_current_device
for mod in list(sys.modules.values()):
    if mod is not None and getattr(mod, "get_current_device", None) is orig:
        setattr(mod, "get_current_device", lambda: cpu)
acc = get_accelerator()
acc.get_rng_state = lambda *a, **k: torch.get_rng_state()
acc.set_rng_state = lambda st, *a, **k: torch.set_rng_state(st)
acc.manual_seed = acc.manual_seed_all = lambda s: torch.manual_seed(s)
acc.synchronize = acc.empty_cache = lambda *a, **k: None
acc.current_device = lambda: 0
acc.is_available = lambda: True
dist.init_process_group("gloo", rank=0, world_size=1, init_method="tcp://127.0.0.1:%d" % (20000 + os.getpid() % 20000))
for mode in ParallelMode:
    gpc._world_sizes[mode], gpc._local_ranks[mode], gpc._global_ranks[mode] = 1, 0, 0
    gpc._groups[mode], gpc._ranks_in_group[mode] = dist.group.WORLD, [0]

model_cfg = dict(checkpoint=False, num_chunks=1, num_attention_heads=4, embed_split_hidden=True, vocab_size=64, embed_grad_scale=1,
                 parallel_output=False, hidden_size=32, num_layers=2, mlp_ratio=float(os.environ.get("PROBE_MLP_RATIO", "2")),
                 apply_post_layer_norm=False,
                 dtype=torch.float32, norm_type="rmsnorm", layer_norm_epsilon=1e-5, use_flash_attn=False)
extra = {}
if family in ("INTERNLM2_PUBLIC", "LLAMA2"):
    model_cfg.update(no_bias=True, num_kv_attention_heads=2)
if family == "INTERNLM_MoE":
    model_cfg.update(num_experts=4, moe_use_residual=False, moe_type="GShard")
    extra = dict(moe=dict(top_k=1, capacity_factor=4.0, eval_capacity_factor=4.0, min_capacity=4, noisy_gate_policy=None,
                          drop_tokens=True, use_rts=False),
                 loss=dict(label_smoothing=0, moe_loss_coeff=0.1))
gpc._config = Config(dict(
    model=model_cfg, model_type=family, use_fp32_norm=False,
    data=dict(seq_len=16, micro_bsz=2, micro_num=1, use_packed_dataset=False),
    parallel=dict(zero1=dict(size=1, fsdp=False), tensor=dict(size=1, mode="mtp"), pipeline=dict(size=1, interleaved_overlap=False),
                  weight=dict(size=1, overlap=False, memory_pool=False), sequence_parallel=False), **extra))
gpc.expert_parallel_size = 1
gpc.set_seed(1024)

import internlm.model  # noqa: F401,E402  (registers the families)
from internlm.utils.registry import MODEL_INITIALIZER  # noqa: E402

torch.manual_seed(0)
model = MODEL_INITIALIZER.get_module(module_name=family)(**model_cfg).float().eval()
for p in model.parameters():            # biases / norm weights start at 0 / 1: give every parameter a value that matters
    if p.dim() == 1:
        p.data.add_(0.1 * torch.randn_like(p))
given = torch.load(sys.argv[4], weights_only=False) if len(sys.argv) > 4 else None
if given is not None:       # weights that came from somewhere else (a reverted HF model): they must fit key for key
    missing, unexpected = model.load_state_dict(given["state"], strict=False)
    assert not missing and not unexpected, (missing, unexpected)
state = {k: v.clone() for k, v in model.state_dict().items()}
torch.manual_seed(1)
ids = given["ids"] if given is not None else torch.randint(1, 64, (2, 16))
with torch.no_grad():
    out = model(input_ids=ids)
moe_losses = None
if family == "INTERNLM_MoE":
    out, moe_losses = out
    moe_losses = [float(x) for x in moe_losses]
torch.save({"state": state, "ids": ids, "logits": out.float(), "moe_losses": moe_losses}, dst)
print("PROBE_OK", tuple(out.shape), flush=True)
os._exit(0)      # the reference leaves helper threads behind whose teardown may abort the interpreter; the work is done
