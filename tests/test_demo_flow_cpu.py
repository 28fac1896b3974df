"""The whole user journey on CPU / gloo (the reference checks the same chain in its `demo_in_readme` CI job,
`ci_scripts/data/tokenizer_*.sh` → `ci_scripts/train/torchrun.sh` → `ci_scripts/model/convert_to_hf.sh` → load):

raw text → `tools/tokenizer.py` shards → `train.py` (2 ranks, real data folder, checkpoint every 4 steps) → second launch
auto-resumes from the newest checkpoint → `tools/convert2hf.py` → `AutoModelForCausalLM.from_pretrained(trust_remote_code)`
gives the same logits as the training framework's own model on the checkpoint weights.
"""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PY = sys.executable
TRAIN_PY = os.path.join(ROOT, "train.py")      # launched from the test's tmp dir: logs / traces / tensorboards land there

CONFIG = '''
JOB_NAME = "demo_flow"
model_type = "{model_type}"
ckpt = dict(enable_save_ckpt=True, save_ckpt_folder="local:{ckpt}", checkpoint_every=4, auto_resume=True,
            async_upload=False, oss_snapshot_freq=0)
data = dict(seq_len=64, micro_num=2, micro_bsz=2, valid_micro_num=1, valid_every=4, pack_sample_into_one=False,
            total_steps={steps}, skip_batches="", rampup_batch_size="", min_length=4, train_folder="{train}",
            valid_folder="{valid}", valid_min_length=0, empty_cache_and_diag_interval=200, diag_outlier_ratio=1.1)
grad_scaler = dict(fp16=dict(initial_scale=2**16, min_scale=1, growth_interval=1000), growth_factor=2, backoff_factor=0.5,
                   max_scale=2**24, hysteresis=2)
hybrid_zero_optimizer = dict(overlap_sync_grad=False, overlap_sync_param=False, reduce_bucket_size=512 * 1024 * 1024,
                             clip_grad_norm=1.0)
loss = dict(label_smoothing=0)
adam = dict(lr=3e-3, adam_beta1=0.9, adam_beta2=0.95, adam_beta2_c=0, adam_eps=1e-8, weight_decay=0.01)
lr_scheduler = dict(total_steps=12, init_steps=0, warmup_ratio=0.1, eta_min=1e-4, last_epoch=-1)
beta2_scheduler = dict(init_beta2=0.95, c=0, cur_iter=-1)
use_fp32_norm = False
model = dict(checkpoint=False, num_chunks=1, num_attention_heads=4, embed_split_hidden=True, vocab_size=64,
             embed_grad_scale=1, parallel_output=True, hidden_size={hidden}, num_layers=2, no_bias=True, mlp_ratio=2,
             apply_post_layer_norm=False, dtype="{dtype}", norm_type="rmsnorm", layer_norm_epsilon=1e-5,
             num_kv_attention_heads=2, use_flash_attn=True{model_extra})
{moe_section}
parallel = dict(zero1=dict(size=-1), tensor=dict(size={tp}, mode="mtp"), pipeline=dict(size={pp}, interleaved_overlap=True),
                weight=dict(size=1, overlap=True, memory_pool=True))
cudnn_deterministic = False
cudnn_benchmark = False
enable_tb = False
monitor = dict(alert=dict(enable_feishu_alert=False, feishu_alert_address=None, light_monitor_address=None,
                          alert_file_path="{ckpt}/alert.log"), tensorboard=dict(queue_max_length=10))
'''


def _run(cmd, cwd, timeout=600, env=None):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, f"{' '.join(cmd)}\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    return r.stdout + r.stderr


def run_flow(tmp_path, gpu: bool, tp: int = 1, pp: int = 1, moe: bool = False):
    """``gpu=False``: 2 gloo ranks, fp32, plain-PyTorch ops.  ``gpu=True``: 2 GPUs, bf16, the sm_100a kernels and the fused
    Hybrid-ZeRO step over peer memory (``tests/test_demo_flow_gpu.py``)."""
    import sentencepiece as spm

    # ---- 1. corpus + tokenizer + shards
    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa"]
    rng = np.random.RandomState(0)
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("\n".join(" ".join(rng.choice(words, 12)) for _ in range(600)))
    spm.SentencePieceTrainer.Train(input=str(corpus), model_prefix=str(tmp_path / "tok"), vocab_size=64, bos_id=1, eos_id=2,
                                   unk_id=0, pad_id=-1, model_type="bpe", minloglevel=2)
    tok_model = str(tmp_path / "tok.model")
    for split in ("train", "valid"):
        for lang in ("en", "cn"):          # two dataset types: exercises the per-type loss / accuracy bookkeeping
            os.makedirs(tmp_path / "data" / split / lang)
            _run([PY, "tools/tokenizer.py", "--text_input_path", str(corpus), "--bin_output_path",
                  str(tmp_path / "data" / split / lang / "part0.bin"), "--tokenizer_model", tok_model], ROOT)
    assert os.path.exists(tmp_path / "data" / "train" / "en" / "part0.bin.meta")

    # ---- 2. train 8 steps on 2 ranks (gloo), checkpoints at 4 and 8
    ckpt = tmp_path / "ckpts"
    world = 2 * tp * pp          # data parallel 2 on top of the model-parallel layout

    def launch(steps, port):
        cfg = tmp_path / f"cfg_{steps}.py"
        text = CONFIG.format(ckpt=ckpt, steps=steps, train=tmp_path / "data" / "train", valid=tmp_path / "data" / "valid",
                             hidden=512 if gpu else 64, dtype="torch.bfloat16" if gpu else "torch.float32", tp=tp, pp=pp,
                             model_type="INTERNLM_MoE" if moe else "INTERNLM2_PUBLIC",
                             model_extra=", num_experts=4, moe_type='GShard'" if moe else "",
                             moe_section="moe = dict(top_k=2)" if moe else "")
        if moe:   # v1 block: multi-head attention with biases
            text = text.replace("num_kv_attention_heads=2, ", "").replace("no_bias=True, ", "")
            text = text.replace("loss = dict(label_smoothing=0)", "loss = dict(label_smoothing=0, moe_loss_coeff=0.1)")
        cfg.write_text(text + ("fused_comm = True\n" if gpu else ""))
        env = dict(os.environ) if gpu else dict(os.environ, CUDA_VISIBLE_DEVICES="")
        return _run([PY, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                     "--master-port", str(port), TRAIN_PY, "--config", str(cfg), "--launcher", "torch", "--backend",
                     "nccl" if gpu else "gloo"], str(tmp_path), timeout=900, env=env)

    from common import find_free_port

    launch(8, find_free_port())
    saved = sorted(int(d) for d in os.listdir(ckpt) if d.isdigit())
    assert saved == [4, 8], os.listdir(ckpt)
    files = set(os.listdir(ckpt / "8"))
    want = {"context.pt", "sampler.pt", "schedulder.pt", "model_config.pt"}
    for t in range(tp):
        for q in range(pp):
            want |= {f"model_tp{t}_pp{q}.pt", f"optimizer_tp{t}_pp{q}_zo0.pt", f"optimizer_tp{t}_pp{q}_zo1.pt"}
    assert want <= files, (sorted(want - files), sorted(files))

    # ---- 3. second launch: auto-resume from step 8, run to 12
    log = launch(12, find_free_port())
    assert "12" in [d for d in os.listdir(ckpt)], os.listdir(ckpt)
    assert "resum" in log.lower() or "load" in log.lower()
    assert "Validation on en" in log and "Validation on cn" in log, "one validation set per sub-folder of valid_folder"

    if moe:   # expert parallel (ep = dp = 2): every expert-parallel rank wrote its own experts' file next to the dense part
        files = set(os.listdir(ckpt / "12"))
        assert any("expert" in f or "moe" in f for f in files), sorted(files)
        return
    # ---- 4. convert the last checkpoint to HF and load it with the Auto classes
    hf = tmp_path / "hf"
    _run([PY, "tools/convert2hf.py", "--src", str(ckpt / "12"), "--tgt", str(hf), "--dtype", "float32", "--tokenizer", tok_model,
          "--max_pos", "128"], ROOT)
    from transformers import AutoModelForCausalLM, AutoTokenizer

    model = AutoModelForCausalLM.from_pretrained(str(hf), trust_remote_code=True, torch_dtype=torch.float32).eval()
    tok = AutoTokenizer.from_pretrained(str(hf), trust_remote_code=True)
    ids = tok("alpha beta gamma", return_tensors="pt")["input_ids"]
    with torch.no_grad():
        logits = model(input_ids=ids).logits
        out = model.generate(ids, max_new_tokens=5, do_sample=False)
    assert logits.shape == (1, ids.shape[1], 64) and torch.isfinite(logits).all() and out.shape[1] == ids.shape[1] + 5
    # the trained model must have learned the corpus a little: loss on a corpus line well below ln(64) = 4.16
    line = tok(corpus.read_text().split("\n")[0], return_tensors="pt")["input_ids"]
    with torch.no_grad():
        loss = float(model(input_ids=line, labels=line).loss)
    assert loss < 3.9, loss

    if gpu or tp * pp > 1:
        return
    # ---- 5. the way back (SFT workflow): HF folder -> tools/revert_hf.py -> tp = 2 training shards -> `load_ckpt_info`
    #         (model only) in a tensor-parallel run: the very first step already sees a trained model
    back = tmp_path / "from_hf"
    _run([PY, "tools/revert_hf.py", "--src", str(hf), "--tgt", str(back), "--tp_size", "2", "--embed_split"], ROOT)
    assert {"model_tp0_pp0.pt", "model_tp1_pp0.pt", "model_config.pt"} <= set(os.listdir(back))
    text = CONFIG.format(ckpt=tmp_path / "ckpts_sft", steps=2, train=tmp_path / "data" / "train",
                         valid=tmp_path / "data" / "valid", hidden=64, dtype="torch.float32", tp=2, pp=1,
                         model_type="INTERNLM2_PUBLIC", model_extra="", moe_section="")
    text = text.replace("auto_resume=True", f"auto_resume=False, load_ckpt_info=dict(path='local:{back}', content=('model',), "
                                            "ckpt_type='internevo')")
    text = text.replace("lr=3e-3", "lr=1e-5")
    cfg = tmp_path / "cfg_sft.py"
    cfg.write_text(text)
    log = _run([PY, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                "--master-port", str(find_free_port()), TRAIN_PY, "--config", str(cfg), "--launcher", "torch", "--backend",
                "gloo"], str(tmp_path), timeout=900, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    import re

    first = re.search(r"step=0 loss=([0-9.]+)", log)
    assert first is not None, log[-2000:]
    assert float(first.group(1)) < 3.9, first.group(0)       # ln(64) = 4.16 for an untrained model


def test_tokenize_train_resume_convert_load(tmp_path):
    run_flow(tmp_path, gpu=False)


def test_flow_with_tensor_and_pipeline_parallel_checkpoints(tmp_path):
    """Same journey on 8 gloo ranks (dp2 x tp2 x pp2): sharded checkpoint files per (tp, pp, zero) rank, auto-resume with the
    same layout, and `convert2hf` merging the tp / pp shards back into one HF model."""
    run_flow(tmp_path, gpu=False, tp=2, pp=2)


def test_flow_with_moe_expert_parallel_checkpoints(tmp_path):
    """MoE (4 experts, expert parallel over 2 ranks): train, checkpoint (dense part + per-expert files), auto-resume."""
    run_flow(tmp_path, gpu=False, moe=True)


def test_alpaca_sft_shards_train(tmp_path):
    """SFT journey: Alpaca-format JSON -> `tools/alpaca_tokenizer.py` (prompt tokens negated = no loss) -> `train.py` on the
    resulting train / valid folders: the loss falls."""
    import json
    import re

    import sentencepiece as spm
    from common import find_free_port

    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa"]
    rng = np.random.RandomState(0)
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("\n".join(" ".join(rng.choice(words, 12)) for _ in range(300)))
    spm.SentencePieceTrainer.Train(input=str(corpus), model_prefix=str(tmp_path / "tok"), vocab_size=64, bos_id=1, eos_id=2,
                                   unk_id=0, pad_id=-1, model_type="bpe", minloglevel=2)
    samples = [{"instruction": " ".join(rng.choice(words, 5)), "input": "", "output": " ".join(rng.choice(words, 8))}
               for _ in range(300)]
    (tmp_path / "alpaca.json").write_text(json.dumps(samples))
    _run([PY, "tools/alpaca_tokenizer.py", str(tmp_path / "alpaca.json"), str(tmp_path / "data"), str(tmp_path / "tok.model"),
          "--split_ratio", "0.1", "--eoh_id", "3", "--eoa_id", "4", "--nl_id", "5"], ROOT)
    assert os.path.exists(tmp_path / "data" / "train" / "en" / "dataset.bin.meta")
    text = CONFIG.format(ckpt=tmp_path / "ckpts", steps=6, train=tmp_path / "data" / "train", valid=tmp_path / "data" / "valid",
                         hidden=64, dtype="torch.float32", tp=1, pp=1, model_type="INTERNLM2_PUBLIC", model_extra="",
                         moe_section="").replace("enable_save_ckpt=True", "enable_save_ckpt=False")
    cfg = tmp_path / "cfg.py"
    cfg.write_text(text)
    log = _run([PY, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                "--master-port", str(find_free_port()), TRAIN_PY, "--config", str(cfg), "--launcher", "torch", "--backend",
                "gloo"], str(tmp_path), timeout=900, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    losses = [float(x) for x in re.findall(r"step=\d+ loss=([0-9.]+)", log)]
    assert len(losses) >= 6 and losses[-1] < losses[0] - 0.5, losses


def test_flow_with_weight_parallel_checkpoints(tmp_path):
    """ISP (sequence parallel 2 x weight parallel 2): train on real shards, checkpoint in the ``model_tp{t}_wp{w}_pp{p}.pt`` layout
    (embedding split along hidden, head rows over the tensor group), `convert2hf` merges it and the HF model has learned the corpus."""
    import sentencepiece as spm

    from common import find_free_port

    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa"]
    rng = np.random.RandomState(0)
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("\n".join(" ".join(rng.choice(words, 12)) for _ in range(600)))
    spm.SentencePieceTrainer.Train(input=str(corpus), model_prefix=str(tmp_path / "tok"), vocab_size=64, bos_id=1, eos_id=2,
                                   unk_id=0, pad_id=-1, model_type="bpe", minloglevel=2)
    tok_model = str(tmp_path / "tok.model")
    for split in ("train", "valid"):
        os.makedirs(tmp_path / "data" / split / "en")
        _run([PY, "tools/tokenizer.py", "--text_input_path", str(corpus), "--bin_output_path",
              str(tmp_path / "data" / split / "en" / "part0.bin"), "--tokenizer_model", tok_model], ROOT)
    ckpt = tmp_path / "ckpts"
    text = CONFIG.format(ckpt=ckpt, steps=8, train=tmp_path / "data" / "train", valid=tmp_path / "data" / "valid", hidden=64,
                         dtype="torch.float32", tp=2, pp=1, model_type="INTERNLM2_PUBLIC", model_extra="", moe_section="")
    text = text.replace('mode="mtp"', 'mode="isp"').replace("weight=dict(size=1", "weight=dict(size=2")
    cfg = tmp_path / "cfg_isp.py"
    cfg.write_text(text)
    _run([PY, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
          str(find_free_port()), TRAIN_PY, "--config", str(cfg), "--launcher", "torch", "--backend", "gloo"], str(tmp_path), timeout=900,
         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    files = set(os.listdir(ckpt / "8"))
    assert {"model_tp0_wp0_pp0.pt", "model_tp1_wp1_pp0.pt", "8.step"} <= files, sorted(files)
    st = torch.load(ckpt / "8" / "model_tp1_wp1_pp0.pt", weights_only=False)
    emb = next(v for k, v in st.items() if k.endswith("tok_embeddings.weight"))
    head = next(v for k, v in st.items() if k.endswith("output.weight"))
    assert tuple(emb.shape) == (64, 32) and tuple(head.shape) == (32, 64), (emb.shape, head.shape)
    hf = tmp_path / "hf"
    _run([PY, "tools/convert2hf.py", "--src", str(ckpt / "8"), "--tgt", str(hf), "--dtype", "float32", "--tokenizer", tok_model,
          "--max_pos", "128"], ROOT)
    from transformers import AutoModelForCausalLM, AutoTokenizer

    model = AutoModelForCausalLM.from_pretrained(str(hf), trust_remote_code=True, dtype=torch.float32).eval()
    tok = AutoTokenizer.from_pretrained(str(hf), trust_remote_code=True)
    line = tok(corpus.read_text().split("\n")[0], return_tensors="pt")["input_ids"]
    with torch.no_grad():
        loss = float(model(input_ids=line, labels=line).loss)
    assert loss < 3.9, loss        # ln(64) = 4.16 untrained
