"""Tiny InternLM2 for plumbing runs: world_size=2 data-parallel on CPU/gloo (or GPUs when present).

    torchrun --nproc_per_node=2 --master-addr 127.0.0.1 train.py --config configs/demo.py --launcher torch
"""
JOB_NAME = "demo_tiny_internlm2"
model_type = "INTERNLM2_PUBLIC"

VOCAB_SIZE = 512
SEQ_LEN = 128
HIDDEN_SIZE = 128
NUM_ATTENTION_HEAD = 4
NUM_KV_ATTENTION_HEAD = 2
MLP_RATIO = 2
NUM_LAYER = 2

ckpt = dict(enable_save_ckpt=False, auto_resume=False)

data = dict(
    seq_len=SEQ_LEN,
    micro_num=2,
    micro_bsz=2,
    valid_micro_num=2,
    valid_every=0,
    pack_sample_into_one=False,
    total_steps=20,
    skip_batches="",
    rampup_batch_size="",
    min_length=0,
    train_folder=None,
    valid_folder=None,
    empty_cache_and_diag_interval=200,
    diag_outlier_ratio=1.1,
    num_random_samples=2000,
)

grad_scaler = dict(
    fp16=dict(initial_scale=2**16, min_scale=1, growth_interval=1000),
    growth_factor=2,
    backoff_factor=0.5,
    max_scale=2**24,
    hysteresis=2,
)
hybrid_zero_optimizer = dict(overlap_sync_grad=False, overlap_sync_param=False, reduce_bucket_size=512 * 1024 * 1024,
                             clip_grad_norm=1.0)
loss = dict(label_smoothing=0)
adam = dict(lr=3e-3, adam_beta1=0.9, adam_beta2=0.95, adam_beta2_c=0, adam_eps=1e-8, weight_decay=0.01)
lr_scheduler = dict(total_steps=data["total_steps"], init_steps=0, warmup_ratio=0.1, eta_min=1e-4, last_epoch=-1)
beta2_scheduler = dict(init_beta2=adam["adam_beta2"], c=adam["adam_beta2_c"], cur_iter=-1)

use_fp32_norm = False
model = dict(
    checkpoint=False,
    num_chunks=1,
    num_attention_heads=NUM_ATTENTION_HEAD,
    embed_split_hidden=True,
    vocab_size=VOCAB_SIZE,
    embed_grad_scale=1,
    parallel_output=True,
    hidden_size=HIDDEN_SIZE,
    num_layers=NUM_LAYER,
    no_bias=True,
    mlp_ratio=MLP_RATIO,
    apply_post_layer_norm=False,
    dtype="torch.float32",
    norm_type="rmsnorm",
    layer_norm_epsilon=1e-5,
    num_kv_attention_heads=NUM_KV_ATTENTION_HEAD,
    use_flash_attn=True,
)
parallel = dict(
    zero1=dict(size=-1),
    tensor=dict(size=1, mode="mtp"),
    pipeline=dict(size=1, interleaved_overlap=True),
    weight=dict(size=1, overlap=True, memory_pool=True),
)
cudnn_deterministic = False
cudnn_benchmark = False
enable_tb = False
monitor = dict(alert=dict(enable_feishu_alert=False, feishu_alert_address=None, light_monitor_address=None,
                          alert_file_path=f"llm_alter/{JOB_NAME}_alert.log"),
               tensorboard=dict(queue_max_length=10))
