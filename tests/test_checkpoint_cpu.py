"""Checkpoint compatibility: the reference's directory layout, tracker file and dictionary keys
(utils.py:158-166, :175-176, :188-232, :254-380; generate_samples.py:55-61; data_utils/vqvae_tokenizer.py:38-47)."""
import math
import os
import random

import numpy as np
import pytest
import torch

from oracle import recipes

CFG = dict(num_layers=2, vocab_size=512, hidden_size=64, max_sequence_length=32)


def _model():
    from cogview_b200.model import GPT2Model
    return GPT2Model(num_layers=CFG["num_layers"], vocab_size=CFG["vocab_size"], hidden_size=CFG["hidden_size"],
                     num_attention_heads=1, embedding_dropout_prob=0.0, attention_dropout_prob=0.0,
                     output_dropout_prob=0.0, max_sequence_length=CFG["max_sequence_length"], max_memory_length=0,
                     checkpoint_activations=False)


def test_load_published_layout_release_checkpoint(tmp_path):
    """What generate_samples.py:55-61 reads: <dir>/<tag>/mp_rank_00_model_states.pt with fp16 tensors under 'module'."""
    from cogview_b200 import checkpoint as ck
    sd16 = {k: v.half() for k, v in recipes.gpt2_state_dict(**CFG).items()}
    name = ck.get_checkpoint_name(str(tmp_path), 0, release=True)
    assert name.endswith(os.path.join("release", "mp_rank_00_model_states.pt"))
    os.makedirs(os.path.dirname(name))
    torch.save({"module": sd16, "iteration": 5000}, name)
    assert ck.get_checkpoint_iteration(str(tmp_path)) == (0, False, False)          # no tracker file yet
    with open(ck.get_checkpoint_tracker_filename(str(tmp_path)), "w") as f:
        f.write("release")
    assert ck.get_checkpoint_iteration(str(tmp_path)) == (0, True, True)
    m = _model().bfloat16()
    assert ck.load_checkpoint(m, None, None, str(tmp_path)) == 0                    # release -> restart the count
    got = m.state_dict()
    assert set(got) == set(sd16)
    for k, v in sd16.items():
        assert torch.equal(got[k], v.to(torch.bfloat16)), k
    m2 = _model()
    assert ck.load_checkpoint(m2, None, None, name) == 5000                         # the .pt file itself also works
    assert torch.equal(m2.state_dict()["word_embeddings.weight"], sd16["word_embeddings.weight"].float())


def test_save_then_load_restores_model_optimizer_scheduler_and_rng(tmp_path):
    from cogview_b200 import checkpoint as ck
    from cogview_b200 import mpu
    m = _model()
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 1.0 / (1 + it))
    for p in m.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    sched.step()
    random.seed(3)
    np.random.seed(4)
    torch.manual_seed(5)
    mpu.random.set_dropout_site_counter(77)

    class Wrapped(torch.nn.Module):                                                 # torch DDP exposes .module
        def __init__(self, module):
            super().__init__()
            self.module = module
    name = ck.save_checkpoint(1200, Wrapped(m), opt, sched, str(tmp_path))
    assert open(ck.get_checkpoint_tracker_filename(str(tmp_path))).read() == "1200"
    assert name == ck.get_checkpoint_name(str(tmp_path), 1200)
    sd = torch.load(name, map_location="cpu", weights_only=False)
    for key in ("iteration", "module", "optimizer", "lr_scheduler", "random_rng_state", "np_rng_state",
                "torch_rng_state", "cuda_rng_state", "rng_tracker_states"):          # the reference's keys
        assert key in sd, key
    expect = (random.random(), float(np.random.rand()), torch.rand(3))
    want = {k: v.clone() for k, v in m.state_dict().items()}
    mom = [opt.state[p]["momentum_buffer"].clone() for p in m.parameters()]

    m2 = _model()
    opt2 = torch.optim.SGD(m2.parameters(), lr=0.5, momentum=0.9)
    sched2 = torch.optim.lr_scheduler.LambdaLR(opt2, lambda it: 1.0 / (1 + it))
    mpu.random.set_dropout_site_counter(0)
    assert ck.load_checkpoint(Wrapped(m2), opt2, sched2, str(tmp_path)) == 1200
    for k, v in m2.state_dict().items():
        assert torch.equal(v, want[k]), k
    for p, b in zip(m2.parameters(), mom):
        assert torch.equal(opt2.state[p]["momentum_buffer"], b)
    assert sched2.last_epoch == sched.last_epoch and opt2.param_groups[0]["lr"] == opt.param_groups[0]["lr"]
    got = (random.random(), float(np.random.rand()), torch.rand(3))
    assert got[0] == expect[0] and got[1] == expect[1] and torch.equal(got[2], expect[2])
    assert mpu.random.get_dropout_site_counter() == 77
    # finetune: weights only, iteration restarts
    m3 = _model()
    assert ck.load_checkpoint(m3, None, None, str(tmp_path), finetune=True) == 0
    assert torch.equal(m3.state_dict()["word_embeddings.weight"], want["word_embeddings.weight"])


def test_tracker_file_errors_and_missing_module(tmp_path):
    from cogview_b200 import checkpoint as ck
    with open(ck.get_checkpoint_tracker_filename(str(tmp_path)), "w") as f:
        f.write("garbage")
    with pytest.raises(ValueError):
        ck.get_checkpoint_iteration(str(tmp_path))
    bad = tmp_path / "x.pt"
    torch.save({"iteration": 1}, str(bad))
    with pytest.raises(KeyError):
        ck.load_checkpoint(_model(), None, None, str(bad))
    w = torch.arange(12.0).view(4, 3)
    assert torch.equal(ck.extend_position_embedding(w, 8), torch.cat((w, w)))


def test_vqvae_checkpoint_with_dataparallel_prefix(tmp_path):
    from cogview_b200 import checkpoint as ck
    from cogview_b200 import vqvae
    torch.manual_seed(0)
    src = vqvae.new_model()
    path = str(tmp_path / "vqvae_hard_biggerset_011.pt")
    torch.save({"module." + k: v for k, v in src.state_dict().items()}, path)
    dst = vqvae.new_model()
    with torch.no_grad():
        for p in dst.parameters():
            p.zero_()
    ck.load_vqvae_checkpoint(dst, path)
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k


def test_fused_adamw_state_dict_round_trip_keeps_fp32_state(tmp_path):
    """torch.optim.Optimizer.load_state_dict casts state tensors to the (bf16) parameter dtype; FusedAdamW must
    restore master / exp_avg / exp_avg_sq as contiguous fp32 — the kernels address them as float* — bit for bit."""
    from cogview_b200.optim import FusedAdamW
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(33).bfloat16()), torch.nn.Parameter(torch.randn(4, 8).bfloat16())]
    opt = FusedAdamW([{'params': [ps[0]]}, {'params': [ps[1]], 'weight_decay': 0.0}], lr=1e-3, max_grad_norm=1.0)
    for i, p in enumerate(ps):
        st = opt._state_for(p)
        st['step'] = 5
        st['master'] += 1e-5 * (i + 1)          # not representable in bf16: a bf16 round trip would lose it
        st['exp_avg'].normal_()
        st['exp_avg_sq'].uniform_()
    path = os.path.join(tmp_path, 'opt.pt')
    torch.save(opt.state_dict(), path)
    qs = [torch.nn.Parameter(torch.zeros(33).bfloat16()), torch.nn.Parameter(torch.zeros(4, 8).bfloat16())]
    opt2 = FusedAdamW([{'params': [qs[0]]}, {'params': [qs[1]], 'weight_decay': 0.0}], lr=1e-3, max_grad_norm=1.0)
    opt2.load_state_dict(torch.load(path, weights_only=False))
    for p, q in zip(ps, qs):
        a, b = opt.state[p], opt2.state[q]
        assert b['step'] == 5
        for k in ('master', 'exp_avg', 'exp_avg_sq'):
            assert b[k].dtype == torch.float32 and b[k].is_contiguous() and torch.equal(a[k], b[k]), k
    assert opt2._applied == 5
    # a reference-format optimizer entry (FP16_Optimizer dict) is refused with a clear message
    with pytest.raises(ValueError):
        opt2.load_state_dict({'loss_scaler': None, 'optimizer_state_dict': {}, 'fp32_from_fp16': []})


def test_annealing_lr_matches_reference_schedule_and_checkpoint_keys(tmp_path):
    """learning_rates.py:22-88: warm-up + linear / cosine decay, and the state_dict keys a reference checkpoint holds."""
    from cogview_b200 import checkpoint
    from cogview_b200.learning_rates import AnnealingLR
    p = [torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.SGD(p, lr=1.0)
    sch = AnnealingLR(opt, start_lr=4e-4, warmup_iter=10, num_iters=100, decay_style='cosine', last_iter=-1,
                      decay_ratio=0.1)
    assert opt.param_groups[0]['lr'] == 0.0
    for _ in range(5):
        sch.step()
    assert abs(opt.param_groups[0]['lr'] - 4e-4 * 5 / 10) < 1e-12
    for _ in range(55):
        sch.step()
    frac = (60 - 10) / 100
    want = 4e-4 / 10 * ((math.cos(math.pi * frac) + 1) * (10 - 1) / 2 + 1)
    assert abs(opt.param_groups[0]['lr'] - want) < 1e-12
    sd = sch.state_dict()
    assert set(sd) == {'warmup_iter', 'num_iters', 'decay_style', 'end_iter', 'decay_ratio'} and sd['num_iters'] == 60
    # resume through load_checkpoint with a reference-style optimizer entry: no_load_optim keeps weights + schedule
    model = torch.nn.Linear(2, 2)
    name = checkpoint.get_checkpoint_name(str(tmp_path), 60)
    os.makedirs(os.path.dirname(name), exist_ok=True)
    torch.save({'iteration': 60, 'module': model.state_dict(), 'optimizer': {'loss_scaler': 1}, 'lr_scheduler': sd}, name)
    open(checkpoint.get_checkpoint_tracker_filename(str(tmp_path)), 'w').write('60')
    opt2 = torch.optim.SGD(model.parameters(), lr=1.0)
    sch2 = AnnealingLR(opt2, 4e-4, 10, 100, 'cosine', -1, 0.1)
    with pytest.raises(RuntimeError):
        checkpoint.load_checkpoint(model, opt2, sch2, str(tmp_path))
    it = checkpoint.load_checkpoint(model, opt2, sch2, str(tmp_path), no_load_optim=True)
    assert it == 60 and sch2.num_iters == 60 and abs(opt2.param_groups[0]['lr'] - want) < 1e-12
