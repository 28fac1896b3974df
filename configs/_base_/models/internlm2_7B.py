# model / parallel fragment, composed with `with read_base(): from configs._base_.models.internlm2_7B import *`
model_type = "INTERNLM2_PUBLIC"
VOCAB_SIZE = 92544
HIDDEN_SIZE = 4096
NUM_ATTENTION_HEAD = 32
NUM_KV_ATTENTION_HEAD = 8
MLP_RATIO = 3.5
NUM_LAYER = 32
model = dict(
    num_chunks=1, checkpoint=False, dtype="torch.bfloat16", embed_split_hidden=True, num_layers=NUM_LAYER,
    hidden_size=HIDDEN_SIZE, vocab_size=VOCAB_SIZE, embed_grad_scale=1, parallel_output=True,
    num_attention_heads=NUM_ATTENTION_HEAD, num_kv_attention_heads=NUM_KV_ATTENTION_HEAD, mlp_ratio=MLP_RATIO,
    norm_type="rmsnorm", layer_norm_epsilon=1e-5, no_bias=True, apply_post_layer_norm=False, use_flash_attn=True,
)
parallel = dict(zero1=dict(size=8), tensor=dict(size=1, mode="mtp"), pipeline=dict(size=1, interleaved_overlap=True),
                weight=dict(size=1, overlap=True, memory_pool=True))
