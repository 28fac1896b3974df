"""Run by ``test_reference_differential_cpu.py`` in two sub-processes: once with the UNMODIFIED reference first on ``sys.path`` and
once with this repository (whose ``internlm`` package is an alias of ``internevo_b200``).  The script only uses the reference's
import paths and public signatures; what it prints is compared value by value.

    python differential_probe.py <root that provides `internlm`> <work dir with en/part0.bin> <output json>
"""
import json
import os
import sys

root, work, dst = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, root)

import torch  # noqa: E402

from internlm.core.context import global_context as gpc  # noqa: E402
from internlm.core.context.parallel_context import Config  # noqa: E402

out = {}
gpc._config = Config(dict(data=dict(seq_len=64, micro_bsz=2, micro_num=2, pack_sample_into_one=False, min_length=0,
                                    use_packed_dataset=True, break_mode="cut"),
                          model=dict(dtype=torch.float32), parallel=dict(sequence_parallel=False)))


def tl(v):
    if torch.is_tensor(v):
        return v.tolist()
    if hasattr(v, "tolist"):
        return v.tolist()
    if isinstance(v, (list, tuple)):
        return [tl(x) for x in v]
    if isinstance(v, dict):
        return {k: tl(x) for k, x in v.items()}
    return v


# ---- StaticBatchSampler: batches, epoch roll-over, ramp-up, resume from a state dict -----------------------------------------
from internlm.data.tokenized.batch_sampler import StaticBatchSampler  # noqa: E402


class Sized:
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


def take(sampler, k):
    it, res = iter(sampler), []
    while len(res) < k:
        try:
            res.append([int(i) for i in next(it)])
        except StopIteration:           # the training loop re-iterates at the end of an epoch
            res.append("epoch")
            it = iter(sampler)
    return res


CASES = [("plain", dict(batch_size=8, rampup_batch_size="", micro_bsz=2, seed=3, data_rank=0, data_world_size=2), 50, 9),
         ("rank1", dict(batch_size=8, rampup_batch_size="", micro_bsz=2, seed=3, data_rank=1, data_world_size=2), 50, 9),
         ("rampup", dict(batch_size=12, rampup_batch_size="4 4 2", micro_bsz=2, seed=5, data_rank=0, data_world_size=1), 100, 12)]
for name, kw, n, k in CASES:
    out["sampler_" + name] = take(StaticBatchSampler([Sized(n)], **kw), k)
half = StaticBatchSampler([Sized(50)], **CASES[0][1])
it = iter(half)
next(it), next(it)
state = half.state_dict()
resumed = StaticBatchSampler([Sized(50)], **CASES[0][1])
resumed.load_state_dict(state)
out["sampler_resume"] = take(resumed, 4)
out["sampler_state_keys"] = sorted(state.keys())

# ---- tokenized file -> JsonlDataset -> packed datasets -> collate -------------------------------------------------------------
from internlm.data.tokenized.collaters import jsonl_ds_collate_fn, packed_collate_fn  # noqa: E402
from internlm.data.tokenized.dataset import JsonlDataset  # noqa: E402
from internlm.data.tokenized.packed_dataset import PackedDatasetWithCut, PackedDatasetWithoutCuSeqlen  # noqa: E402

ds = JsonlDataset(os.path.join(work, "en", "part0.bin"), 1, min_length=0)
out["jsonl_len"] = len(ds)
out["jsonl_items"] = [dict(tokens=[int(t) for t in ds[i]["tokens"]], type_id=int(ds[i]["type_id"])) for i in range(len(ds))]
one = PackedDatasetWithoutCuSeqlen(ds, 64, 128)
out["pack_into_one"] = [tl(one[i]) for i in range(len(one))]
cut = PackedDatasetWithCut(ds, 64, 128)
out["pack_with_cut"] = [tl(cut[i]) for i in range(len(cut))]
batch = packed_collate_fn([cut[0], cut[1]], 128)
out["packed_collate"] = [tl(batch[0]), tl(batch[1])]
batch = jsonl_ds_collate_fn([ds[0], ds[1], ds[2]], 32)
out["jsonl_collate"] = [tl(batch[0]), tl(batch[1])]

# ---- schedules --------------------------------------------------------------------------------------------------------------
from internlm.solver.schedulers.beta2_scheduler import Beta2Scheduler  # noqa: E402
from internlm.solver.schedulers.lr_scheduler import FineTuneCosineAnnealingWarmupLR  # noqa: E402

p = torch.nn.Parameter(torch.zeros(1))
for name, kw in [("cos", dict(total_steps=400, init_steps=0, warmup_ratio=0.05, eta_min=1e-5, last_epoch=-1)),
                 ("cos_init", dict(total_steps=400, init_steps=7, warmup_ratio=0.1, eta_min=1e-4, last_epoch=-1))]:
    opt = torch.optim.SGD([p], lr=1e-3)
    sch = FineTuneCosineAnnealingWarmupLR(opt, **kw)
    lrs = []
    for _ in range(420):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    out["lr_" + name] = lrs
opt = torch.optim.AdamW([p], lr=1e-3, betas=(0.9, 0.95))
b2 = Beta2Scheduler(opt, init_beta2=0.95, c=0.8, cur_iter=-1)
seq = []
for _ in range(40):
    b2.step()
    seq.append(opt.param_groups[0]["betas"][1])
out["beta2"] = seq

# ---- loss scaler: growth / back-off / hysteresis ------------------------------------------------------------------------------
from internlm.solver.optimizer.utils import DynamicGradScaler  # noqa: E402

scaler = DynamicGradScaler(initial_scale=2**10, min_scale=1, growth_factor=2, backoff_factor=0.5, growth_interval=3,
                           max_scale=2**14, hysteresis=2)
seq = []
for overflow in [0, 0, 0, 1, 0, 1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]:
    scaler.update(bool(overflow))
    seq.append(float(scaler.scale))
out["scaler"] = seq

# ---- reported TFLOPS, layer partition -----------------------------------------------------------------------------------------
from internlm.solver.pipeline_utils import partition_uniform  # noqa: E402
from internlm.utils.common import get_megatron_flops  # noqa: E402

out["flops"] = [float(get_megatron_flops(1.5, checkpoint=c, seq_len=4096, hidden_size=4096, num_layers=32, vocab_size=92544,
                                         global_batch_size=32, global_world_size=8, mlp_ratio=3.5, use_swiglu=s))
                for c in (False, True) for s in (True, False)]
out["partition"] = {f"{L}_{P}_{C}": tl(partition_uniform(L, P, C))
                    for L, P, C in [(32, 4, 1), (30, 4, 1), (32, 4, 2), (48, 8, 1), (14, 4, 1), (24, 3, 2), (60, 8, 1)]}

# ---- skip_batches parser, un-packing of a packed row ----------------------------------------------------------------------------
from internlm.data.utils import unpack_data  # noqa: E402
from internlm.utils.common import parse_args  # noqa: F401,E402  (import path exists on both sides)

ids = torch.arange(1, 25).reshape(2, 12)
cu = torch.tensor([[0, 5, 9, 12], [0, 3, 12, 12]])
gpc._config.data.micro_bsz = 3
try:
    out["unpack"] = tl(unpack_data(ids, cu))
except Exception as e:   # noqa: BLE001
    out["unpack"] = "error: " + type(e).__name__

# ---- skip_batches ranges, GShard gating ------------------------------------------------------------------------------------------
from internlm.model.moe.gshard_layer import top1gating, top2gating  # noqa: E402
from internlm.utils.common import BatchSkipper  # noqa: E402

skip = BatchSkipper("2-4,7,10-11")
out["skipper"] = [bool(skip(i)) for i in range(14)]
torch.manual_seed(0)
logits = torch.randn(32, 4)
for name, fn, kw in [("top2", top2gating, dict(capacity_factor=1.0, min_capacity=2)),
                     ("top2_drop", top2gating, dict(capacity_factor=0.5, min_capacity=1)),
                     ("top1", top1gating, dict(capacity_factor=1.0, min_capacity=2, use_rts=False)),
                     ("top1_drop", top1gating, dict(capacity_factor=0.5, min_capacity=1, use_rts=False))]:
    l_aux, combine, dispatch, counts = fn(logits, **kw)
    out["gate_" + name] = dict(l_aux=float(l_aux), combine=tl(combine.float()), dispatch=tl(dispatch.int()), counts=tl(counts))
out["gate_probs"] = tl(torch.softmax(logits, dim=1))

json.dump(out, open(dst, "w"))
print("PROBE_OK")
