"""The reference's ``evaluate_on_val_dls`` on CPU: a small fp32 InternLM2 built by its ``initialize_model`` (torch attention path), the
validation loaders of a data folder built by its ``build_valid_loader_with_data_type``; saves the weights and every scalar it reports
(see ``test_reference_differential_cpu.py``).

    python differential_eval_probe.py <reference root> <output .pt> <data folder with valid/<type>/*.bin>
"""
import os, sys, contextlib
root, dst = sys.argv[1], sys.argv[2]
family = "INTERNLM2_PUBLIC"
data_folder = sys.argv[3]
ckpt_folder = None
resume = False
sys.path.insert(0, root)
import torch, torch.distributed as dist
import internlm
import internlm.utils.common as common
from internlm.accelerator import get_accelerator
from internlm.core.context import ParallelMode, global_context as gpc
from internlm.core.context.parallel_context import Config
cpu, orig = torch.device("cpu"), common.get_current_device
for mod in list(sys.modules.values()):
    if mod is not None and getattr(mod, "get_current_device", None) is orig:
        setattr(mod, "get_current_device", lambda: cpu)
acc = get_accelerator()
class _S:
    def wait_stream(self, *a): pass
    def synchronize(self): pass
    def wait_event(self, *a): pass
    def record_event(self, *a): return _E()
class _E:
    def record(self, *a): pass
    def wait(self, *a): pass
    def synchronize(self): pass
    def query(self): return True
acc.get_rng_state = lambda *a, **k: torch.get_rng_state()
acc.set_rng_state = lambda st, *a, **k: torch.set_rng_state(st)
acc.manual_seed = acc.manual_seed_all = lambda s: torch.manual_seed(s)
acc.synchronize = acc.empty_cache = lambda *a, **k: None
acc.current_device = lambda: 0
acc.is_available = lambda: True
type(acc).Stream = property(lambda self: (lambda *a, **k: _S()))
type(acc).Event = property(lambda self: (lambda *a, **k: _E()))
acc.current_stream = lambda *a, **k: _S()
acc.default_stream = lambda *a, **k: _S()
acc.stream = lambda s: contextlib.nullcontext()
acc.memory_allocated = acc.max_memory_allocated = acc.memory_reserved = acc.max_memory_reserved = lambda *a, **k: 0
acc.reset_peak_memory_stats = lambda *a, **k: None
dist.init_process_group("gloo", rank=0, world_size=1, init_method="tcp://127.0.0.1:%d" % (20000 + os.getpid() % 20000))
for mode in ParallelMode:
    gpc._world_sizes[mode], gpc._local_ranks[mode], gpc._global_ranks[mode] = 1, 0, 0
    gpc._groups[mode], gpc._ranks_in_group[mode] = dist.group.WORLD, [0]
S, MB, MN = 16, 2, 2
cfg = dict(
    JOB_NAME="diff", model_type=family, use_fp32_norm=False,
    model=dict(checkpoint=False, num_chunks=1, num_attention_heads=4, embed_split_hidden=True, vocab_size=64, embed_grad_scale=1,
               parallel_output=False, hidden_size=32, num_layers=2, no_bias=True, mlp_ratio=2, apply_post_layer_norm=False,
               dtype=torch.float32, norm_type="rmsnorm", layer_norm_epsilon=1e-5, num_kv_attention_heads=2, use_flash_attn=False),
    data=dict(seq_len=S, micro_bsz=MB, micro_num=MN, use_packed_dataset=False, gradient_accumulation=MN, total_steps=10, valid_every=1, valid_micro_num=2,
              valid_folder=os.path.join(data_folder, "valid"), train_folder=None, type="tokenized"),
    parallel=dict(zero1=dict(size=1, fsdp=False), tensor=dict(size=1, mode="mtp"), pipeline=dict(size=1, interleaved_overlap=False),
                  weight=dict(size=1, overlap=False, memory_pool=False), sequence_parallel=False),
    grad_scaler=dict(fp16=dict(initial_scale=2**16, min_scale=1, growth_interval=1000), growth_factor=2, backoff_factor=0.5, max_scale=2**24, hysteresis=2),
    hybrid_zero_optimizer=dict(overlap_sync_grad=False, overlap_sync_param=False, reduce_bucket_size=512*1024*1024, clip_grad_norm=100.0),
    loss=dict(label_smoothing=0.0),
    adam=dict(lr=3e-3, adam_beta1=0.9, adam_beta2=0.95, adam_beta2_c=0, adam_eps=1e-4, weight_decay=0.01),
    lr_scheduler=dict(total_steps=2000, init_steps=0, warmup_ratio=0.001, eta_min=1e-4, last_epoch=-1),
    beta2_scheduler=dict(init_beta2=0.95, c=0, cur_iter=-1),
    ckpt=dict(enable_save_ckpt=ckpt_folder is not None, save_ckpt_folder=f"local:{ckpt_folder}", checkpoint_every=4, oss_snapshot_freq=0,
              async_upload=False, async_upload_tmp_folder=None, auto_resume=False, stop_file_path=None,
              load_ckpt_info=dict(path=f"local:{ckpt_folder}/4", content=("model", "optimizer", "scheduler"),
                                  ckpt_type="internevo") if resume else None,
              snapshot_ckpt_folder=f"local:{ckpt_folder}/snapshot", is_async_upload=False),
    monitor=dict(alert=dict(enable_feishu_alert=False)), resume_tb_folder=None, tensorboard_folder=None,
)
if family in ("INTERNLM", "INTERNLM_MoE"):
    cfg["model"].pop("no_bias"); cfg["model"].pop("num_kv_attention_heads")
if family == "INTERNLM_MoE":
    cfg["model"].update(num_experts=4, moe_use_residual=False, moe_type="GShard")
    cfg["moe"] = dict(top_k=1, capacity_factor=4.0, eval_capacity_factor=4.0, min_capacity=4, noisy_gate_policy=None, drop_tokens=True, use_rts=False)
    cfg["loss"]["moe_loss_coeff"] = 0.1
gpc._config = Config(cfg)
gpc.expert_parallel_size = 1
gpc.zero1_parallel_size = 1; gpc.data_parallel_size = 1; gpc.tensor_parallel_size = 1; gpc.pipeline_parallel_size = 1; gpc.weight_parallel_size=1
gpc.set_seed(1024)
from internlm.train import initialize_model, initialize_optimizer, get_scheduler_hooks
from internlm.model.losses import FlashGPTLMLoss
from internlm.data.utils import unpack_data
torch.manual_seed(0)
model = initialize_model()
state = {k: v.clone() for k, v in model.model.state_dict().items()}
optimizer, beta2_scheduler, lr_scheduler = initialize_optimizer(model)
criterion = FlashGPTLMLoss(parallel_output=False, label_smoothing=0.0)
trainer, _, _, _ = internlm.initialize_trainer(model=model, optimizer=optimizer, criterion=criterion, lr_scheduler=lr_scheduler, beta2_scheduler=beta2_scheduler,
                                               scheduler_hooks=get_scheduler_hooks(None, optimizer, None))
trainer.train()
from internlm.accelerator.abstract_accelerator import AcceleratorType

acc.get_accelerator_backend = lambda: AcceleratorType.OTHER      # no torch_scatter here: the metric's plain scatter path
from internlm.data import build_valid_loader_with_data_type
from internlm.eval.evaluation import evaluate_on_val_dls

gpc.is_rank_for_log = lambda: True
gpc.get_global_rank = lambda: 0
scalars, lines = {}, []


class _Writer:
    def add_scalar(self, key, value, step):
        scalars[key] = float(value)


class _Logger:
    def info(self, msg, *a, **k):
        lines.append(str(msg))

    warning = error = info


val_dls = build_valid_loader_with_data_type()
evaluate_on_val_dls(trainer, val_dls, _Writer(), _Logger(), step_count=3)
torch.save({"state": state, "scalars": scalars, "sizes": {k: len(v) for k, v in val_dls.items()}}, dst)
print("PROBE_OK", scalars, flush=True)
os._exit(0)
