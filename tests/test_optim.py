"""N4 training glue: parameter arena + fused AdamW/clip kernel + arena all-reduce (aicity_action_b200/optim.py) against
torch.optim.AdamW + torch.nn.utils.clip_grad_norm_ (what the reference runs: tools/train_net.py:229-246)."""
import os
import socket

import pytest
import torch
import torch.nn.functional as F


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _toy():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.LayerNorm(40), torch.nn.GELU(), torch.nn.Linear(40, 7))


@pytest.mark.gpu
@pytest.mark.parametrize("max_norm", [None, 0.05])
def test_fused_adamw_matches_torch(max_norm):
    from aicity_action_b200.optim import FusedAdamW
    ref, ours = _toy().cuda(), _toy().cuda()
    decay = [p for p in ref.parameters() if p.ndim > 1]
    plain = [p for p in ref.parameters() if p.ndim <= 1]
    o_ref = torch.optim.AdamW([{"params": decay, "weight_decay": 0.05}, {"params": plain, "weight_decay": 0.0}],
                              lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    o_ours = FusedAdamW(ours, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05, max_grad_norm=max_norm)
    assert [n for n in o_ours.arena.names[:2]] == ["0.weight", "3.weight"]          # decayed tensors first
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(5):
        x = torch.randn(16, 24, device="cuda", generator=g)
        y = torch.randint(0, 7, (16,), device="cuda", generator=g)
        lr = 1e-2 * (1 + it) / 5                                                   # per-iteration lr, as optim.set_lr does
        for opt in (o_ref, o_ours):
            for grp in opt.param_groups:
                grp["lr"] = lr
        o_ref.zero_grad()
        F.cross_entropy(ref(x), y).backward()
        norm_ref = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm) if max_norm else None
        o_ref.step()
        o_ours.zero_grad()
        F.cross_entropy(ours(x), y).backward()
        o_ours.step()
        if max_norm:
            assert abs(o_ours.last_grad_norm() - float(norm_ref)) < 1e-5 * max(1.0, float(norm_ref))
        for (n, a), b in zip(ours.named_parameters(), ref.parameters()):
            assert torch.allclose(a, b, rtol=2e-5, atol=2e-7), (it, n, float((a - b).abs().max()))
    assert o_ours.steps_done() == 5
    # the bf16 shadow the kernel maintains is the next forward's tensor-core operand
    from aicity_action_b200.weights import cached_weight
    w = ours[0].weight
    assert torch.equal(cached_weight(w, torch.bfloat16), w.detach().bfloat16())
    # state_dict round trip
    sd = o_ours.state_dict()
    o2 = FusedAdamW(_toy().cuda(), lr=1e-2)
    o2.load_state_dict(sd)
    assert o2.steps_done() == 5 and torch.equal(o2.exp_avg, o_ours.exp_avg)


@pytest.mark.gpu
def test_fused_training_step_on_mvit_and_graph_capture():
    """The real model: arena-backed MViT trained by FusedAdamW follows torch AdamW + clip_grad_norm_ on a twin; then the same
    step captured as ONE CUDA graph (GraphedTrainStep) replays to the same parameters as eager stepping."""
    from aicity_action_b200.config import aicity_cfg
    from aicity_action_b200.mvit import MViT
    from aicity_action_b200.optim import FusedAdamW, GraphedTrainStep
    from tests.golden.cases import MODEL_CASES, tiny_cfg_overrides
    from tests.golden.synth import synth_clip, synth_state_dict
    c = MODEL_CASES[0]
    cfg = aicity_cfg(c["yaml"], tiny_cfg_overrides(c) + ["MVIT.DROPPATH_RATE", 0.0, "MODEL.DROPOUT_RATE", 0.0])

    def build():
        m = MViT(cfg).train()
        m.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 5))
        return m.cuda()

    x = synth_clip(5, 4, cfg.DATA.NUM_FRAMES, cfg.DATA.TRAIN_CROP_SIZE).cuda()
    y = torch.tensor([1, 4, 9, 16]).cuda()
    loss_fn = lambda out, t: F.cross_entropy(out.float(), t)
    ref, a, b = build(), build(), build()
    init = [p.detach().clone() for p in ref.parameters()]

    def update_error(model, other):
        """|| (model - other) || / || (other - init) || over all parameters: disagreement relative to the UPDATE itself.
        (Adam turns the ~1e-7 summation-order noise of atomically reduced gradients into +-lr moves on elements whose true
        gradient is ~0, e.g. norm_k.bias, so element-wise relative errors are meaningless here.)"""
        num = sum(float((p.detach() - q.detach()).pow(2).sum()) for p, q in zip(model.parameters(), other.parameters()))
        den = sum(float((q.detach() - i).pow(2).sum()) for q, i in zip(other.parameters(), init))
        return (num / den) ** 0.5
    o_ref = torch.optim.AdamW([{"params": [p for p in ref.parameters() if p.ndim > 1], "weight_decay": 1e-4},
                               {"params": [p for p in ref.parameters() if p.ndim <= 1], "weight_decay": 0.0}], lr=1e-3)
    o_a = FusedAdamW(a, lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0)
    keys = list(ref.state_dict())
    assert list(a.state_dict()) == keys                           # re-homing the parameters leaves the state_dict alone
    for _ in range(3):                                            # fp32 kernels: tight comparison
        o_ref.zero_grad()
        loss_fn(ref([x]), y).backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        o_ref.step()
        o_a.zero_grad()
        loss_fn(a([x]), y).backward()
        o_a.step()
    err = update_error(a, ref)
    assert err < 5e-2, err
    # graph capture of the whole bf16 step == eager bf16 steps from the same start
    o_b = FusedAdamW(b, lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0)
    c_model = build()
    o_c = FusedAdamW(c_model, lr=1e-3, weight_decay=1e-4, max_grad_norm=1.0)
    xb = x.bfloat16()
    step = GraphedTrainStep(c_model, o_c, loss_fn, xb.clone(), y.clone(), warmup=3)      # 3 eager warm-up steps (capture runs nothing)
    for _ in range(3):
        o_b.zero_grad()
        loss_fn(b([xb]), y).backward()
        o_b.step()
    for _ in range(2):
        o_b.zero_grad()
        loss_fn(b([xb]), y).backward()
        o_b.step()
        loss = step()
    assert o_c.steps_done() == o_b.steps_done() == 5
    assert torch.isfinite(loss)
    err = update_error(c_model, b)
    assert err < 0.25, err                # bf16 + atomics: run-to-run summation order differs


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aicity_action_b200.optim import ArenaDataParallel

    class CpuArena:                      # the arena contract on CPU tensors (the CUDA kernels are not involved here)
        def __init__(self, m):
            ps = list(m.parameters())
            self.flat = torch.cat([p.detach().reshape(-1) for p in ps]).clone()
            self.grad = torch.zeros_like(self.flat)
            o = 0
            for p in ps:
                p.data = self.flat[o:o + p.numel()].view_as(p)
                p.grad = self.grad[o:o + p.numel()].view_as(p)
                o += p.numel()

        def sync_shadow(self):
            pass

    torch.manual_seed(100 + rank)                     # different initial weights per rank: the broadcast must fix that
    m = torch.nn.Linear(6, 3)
    arena = CpuArena(m)
    dp = ArenaDataParallel(m, arena)
    torch.manual_seed(7)
    x_all, y_all = torch.randn(4 * world, 6), torch.randn(4 * world, 3)
    F.mse_loss(dp(x_all[4 * rank:4 * rank + 4]), y_all[4 * rank:4 * rank + 4]).backward()
    dp.reduce_gradients()
    if rank == 0:
        q.put((arena.flat.clone(), arena.grad.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_arena_data_parallel_gloo_world2():
    """world_size-2 gloo: broadcast of rank 0's parameters, one all-reduce of the flat gradient buffer == the gradient of
    one process on the concatenated batch."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat, grad = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(100)
    m = torch.nn.Linear(6, 3)
    assert torch.equal(flat, torch.cat([p.detach().reshape(-1) for p in m.parameters()]))
    torch.manual_seed(7)
    x_all, y_all = torch.randn(8, 6), torch.randn(8, 3)
    F.mse_loss(m(x_all), y_all).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    assert torch.allclose(grad, ref, rtol=1e-5, atol=1e-7)
