"""The whole data path of ``train.py`` on a folder of tokenized files (several sub-folders = dataset types, several files each):
``args_sanity_check`` -> ``build_train_loader_with_data_type`` / ``build_valid_loader_with_data_type`` -> the first batches, run against
the reference and against this repository (``internlm`` alias); see ``test_reference_differential_cpu.py``.

    PROBE_DP_RANK=<r> python differential_loader_probe.py <root that provides `internlm`> <data folder> <output json> [one]
"""
import json, os, sys
root, work, dst = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, root)
import torch
from internlm.core.context import ParallelMode, global_context as gpc
from internlm.core.context.parallel_context import Config
pack_one = len(sys.argv) > 4 and sys.argv[4] == "one"
gpc._config = Config(dict(
    JOB_NAME="p", model_type="INTERNLM2_PUBLIC",
    data=dict(seq_len=32, micro_num=2, micro_bsz=2, valid_micro_num=1, valid_every=0, pack_sample_into_one=pack_one, total_steps=20, skip_batches="",
              rampup_batch_size="", min_length=6, train_folder=os.path.join(work, "train"), valid_folder=os.path.join(work, "valid"), num_worker=0),
    model=dict(dtype="torch.bfloat16", use_flash_attn=True, num_layers=2, hidden_size=32, num_attention_heads=4, vocab_size=64, mlp_ratio=2, checkpoint=False),
    parallel=dict(zero1=dict(size=-1), tensor=dict(size=1, mode="mtp"), pipeline=dict(size=1, interleaved_overlap=True), weight=dict(size=1, overlap=True, memory_pool=True)),
    ckpt=dict(enable_save_ckpt=False), adam=dict(lr=1e-3), hybrid_zero_optimizer=dict(overlap_sync_grad=False, overlap_sync_param=False, reduce_bucket_size=1, clip_grad_norm=1.0),
    grad_scaler=dict(fp16=dict(initial_scale=2**16, min_scale=1, growth_interval=1000), growth_factor=2, backoff_factor=0.5, max_scale=2**24, hysteresis=2),
    loss=dict(label_smoothing=0), lr_scheduler=dict(total_steps=20, init_steps=0, warmup_ratio=0.1, eta_min=1e-5, last_epoch=-1),
    beta2_scheduler=dict(init_beta2=0.95, c=0, cur_iter=-1), monitor=dict(alert=dict(enable_feishu_alert=False, feishu_alert_address=None, light_monitor_address=None, alert_file_path=None)),
))
RANK = int(os.environ.get("PROBE_DP_RANK", "0"))
gpc.is_rank_for_log = lambda: True
gpc.get_world_size = lambda mode: 2 if mode in (ParallelMode.GLOBAL, ParallelMode.DATA) else 1
gpc.get_local_rank = lambda mode: RANK if mode in (ParallelMode.GLOBAL, ParallelMode.DATA) else 0
gpc.is_initialized = lambda mode: True
gpc.is_using_parallel_mode = lambda mode: mode in (ParallelMode.DATA,)
import torch.distributed as dist
dist.init_process_group("gloo", rank=0, world_size=1, init_method="tcp://127.0.0.1:%d" % (20000 + os.getpid() % 20000))
gpc.get_group = lambda mode: dist.group.WORLD
gpc.get_global_rank = lambda: 0          # every probe process walks the folder itself
from internlm.initialize.launch import args_sanity_check
args_sanity_check()
from internlm.data import build_train_loader_with_data_type, build_valid_loader_with_data_type
def tl(v):
    if torch.is_tensor(v): return v.tolist()
    if isinstance(v, (list, tuple)): return [tl(x) for x in v]
    if isinstance(v, dict): return {k: tl(x) for k, x in v.items()}
    return v
train_dl, types = build_train_loader_with_data_type()
out = {"types": list(types), "len": len(train_dl), "batches": []}
it = iter(train_dl)
for _ in range(6):
    b = next(it)
    out["batches"].append([tl(b[0]), tl(b[1])])
vals = build_valid_loader_with_data_type()
out["valid"] = {k: [[tl(b[0]), tl(b[1])] for i, b in zip(range(2), dl)] for k, dl in sorted(vals.items())}
json.dump(out, open(dst, "w"))
print("PROBE_OK", flush=True)
os._exit(0)
