JOB_NAME = "7b_moe_train"
model_type = "INTERNLM_MoE"
DO_ALERT = False

VOCAB_SIZE = 103168
SEQ_LEN = 2048
HIDDEN_SIZE = 4096
NUM_ATTENTION_HEAD = 32
NUM_KV_ATTENTION_HEAD = 32
MLP_RATIO = 4 / 3
NUM_LAYER = 32

# Ckpt folder format: "local:/path" | "boto3:s3://bucket.endpoint/path" | "volc:..." | "oss2:..."
SAVE_CKPT_FOLDER = "local:llm_ckpts"
CHECKPOINT_EVERY = 50
ckpt = dict(
    enable_save_ckpt=False,  # enable ckpt save.
    save_ckpt_folder=SAVE_CKPT_FOLDER,
    # load_ckpt_info=dict(path="local:llm_ckpts/50", content=("model",), ckpt_type="internevo"),
    auto_resume=False,  # resume from the newest checkpoint / snapshot under save_ckpt_folder when restarted
    checkpoint_every=CHECKPOINT_EVERY,
    async_upload=True,  # only for object-store backends
    async_upload_tmp_folder="/dev/shm/internlm_tmp_ckpt/",
    oss_snapshot_freq=int(CHECKPOINT_EVERY / 2),
)

TRAIN_FOLDER = None  # "/path/to/dataset"; None -> synthetic RandomDataset
VALID_FOLDER = None
data = dict(
    seq_len=SEQ_LEN,
    micro_num=4,   # micro-batches per optimizer step (gradient accumulation)
    micro_bsz=2,   # packed_length = micro_bsz * SEQ_LEN
    valid_micro_num=4,
    valid_every=0,
    pack_sample_into_one=False,
    total_steps=20,
    skip_batches="",
    rampup_batch_size="",
    min_length=50,
    train_folder=TRAIN_FOLDER,
    valid_folder=VALID_FOLDER,
    empty_cache_and_diag_interval=200,
    diag_outlier_ratio=1.1,
)

grad_scaler = dict(
    fp16=dict(initial_scale=2**16, min_scale=1, growth_interval=1000),
    growth_factor=2,
    backoff_factor=0.5,
    max_scale=2**24,
    hysteresis=2,
)
hybrid_zero_optimizer = dict(
    overlap_sync_grad=False,
    overlap_sync_param=False,
    reduce_bucket_size=512 * 1024 * 1024,
    clip_grad_norm=1.0,
)
loss = dict(label_smoothing=0, moe_loss_coeff=0.1)
adam = dict(lr=0.0001, adam_beta1=0.9, adam_beta2=0.95, adam_beta2_c=0, adam_eps=1e-8, weight_decay=0.01)
lr_scheduler = dict(total_steps=data["total_steps"], init_steps=0, warmup_ratio=0.01, eta_min=1e-5, last_epoch=-1)
beta2_scheduler = dict(init_beta2=adam["adam_beta2"], c=adam["adam_beta2_c"], cur_iter=-1)

use_fp32_norm = False
model = dict(
    checkpoint=False,  # activation checkpointing: True/False or a fraction of layers in [0, 1]
    num_chunks=1,  # virtual pipeline chunks per rank (interleaved 1F1B when > 1)
    num_attention_heads=NUM_ATTENTION_HEAD,
    embed_split_hidden=True,
    vocab_size=VOCAB_SIZE,
    embed_grad_scale=1,
    parallel_output=True,
    hidden_size=HIDDEN_SIZE,
    num_layers=NUM_LAYER,
    mlp_ratio=MLP_RATIO,
    apply_post_layer_norm=False,
    dtype="torch.bfloat16",
    norm_type="rmsnorm",
    layer_norm_epsilon=1e-5,
    use_flash_attn=True,
    num_experts=4,
    moe_use_residual=False,
    moe_type="GShard",
)
"""
zero1:    size <= 0 -> ZeRO over the whole data-parallel group; 1 -> off; k -> sub-group of k ranks (ZeRO-1.5)
tensor:   size + mode in ['mtp', 'msp', 'fsp', 'isp']
pipeline: size, interleaved_overlap
weight:   size, overlap, memory_pool   (isp only)
fused_comm (top level): run the tensor-parallel linears and the ZeRO step as fused peer-memory kernels over NVLink
"""
parallel = dict(
    zero1=dict(size=-1, fsdp=False),
    tensor=dict(size=1, mode="mtp"),
    pipeline=dict(size=1, interleaved_overlap=True),
    weight=dict(size=1, overlap=True, memory_pool=True),
)
fused_comm = True
cudnn_deterministic = False
cudnn_benchmark = False
# moe layer kwargs (moe_type selects the implementation: GShard | MegaBlock | MegaBlock-D)
moe = dict(
    top_k=2,
    capacity_factor=1.0,
    eval_capacity_factor=1.0,
    min_capacity=4,
    noisy_gate_policy=None,
    drop_tokens=True,
    use_rts=True,
)
monitor = dict(
    alert=dict(
        enable_feishu_alert=DO_ALERT,
        feishu_alert_address=None,  # webhook (Feishu/Lark "post" JSON) for alerts
        light_monitor_address=None,  # heartbeat endpoint
        alert_file_path=f"llm_alter/{JOB_NAME}_alert.log",
    ),
    tensorboard=dict(queue_max_length=10),
)
