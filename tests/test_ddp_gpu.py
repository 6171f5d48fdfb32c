"""Data-parallel training of the B200 path over NCCL (build.py:44-53 wraps the model in DistributedDataParallel):
averaged gradients of two ranks == gradients of one process on the concatenated batch.  Needs 2 GPUs (skipped otherwise;
run with `gpurun --gpus 2 -- python -m pytest tests/test_ddp_gpu.py -m gpu`)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _build(seed):
    from aicity_action_b200.config import aicity_cfg
    from aicity_action_b200.mvit import MViT
    from tests.golden.cases import MODEL_CASES, tiny_cfg_overrides
    from tests.golden.synth import synth_state_dict
    c = MODEL_CASES[0]
    cfg = aicity_cfg(c["yaml"], tiny_cfg_overrides(c) + ["MVIT.DROPPATH_RATE", 0.0, "MODEL.DROPOUT_RATE", 0.0])
    m = MViT(cfg).train()
    m.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed))
    return cfg, m


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    torch.distributed.init_process_group("nccl", rank=rank, world_size=world)
    from tests.golden.synth import synth_clip
    cfg, m = _build(7)
    m = m.cuda()
    ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[rank])
    x = synth_clip(7, 2 * world, cfg.DATA.NUM_FRAMES, cfg.DATA.TRAIN_CROP_SIZE)[2 * rank:2 * rank + 2].cuda()
    y = (torch.arange(2 * world) % cfg.MODEL.NUM_CLASSES)[2 * rank:2 * rank + 2].cuda()
    F.cross_entropy(ddp([x]), y).backward()
    if rank == 0:
        torch.save({k: p.grad.cpu() for k, p in m.named_parameters()}, os.path.join(out_dir, "ddp.pt"))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_gradients_match_single_process(tmp_path):
    import torch.multiprocessing as mp
    from tests.golden.synth import synth_clip
    world = 2
    mp.spawn(_worker, args=(world, 29533, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(str(tmp_path), "ddp.pt"))
    cfg, m = _build(7)
    m = m.cuda()
    x = synth_clip(7, 2 * world, cfg.DATA.NUM_FRAMES, cfg.DATA.TRAIN_CROP_SIZE).cuda()
    y = (torch.arange(2 * world) % cfg.MODEL.NUM_CLASSES).cuda()
    F.cross_entropy(m([x]), y).backward()       # mean over 4 clips == average of the two ranks' 2-clip means
    gmax = max(p.grad.abs().max().item() for p in m.parameters())
    for k, p in m.named_parameters():
        err = (got[k] - p.grad.cpu()).abs().max().item()
        assert err <= 1e-4 * max(p.grad.abs().max().item(), 1e-3 * gmax), (k, err)
