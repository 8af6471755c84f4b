"""World-size-2 `gloo` tests (CPU) of the data-parallel host logic: flat gradient bucket all-reduce, the fused
normaliser all-reduce of the distillation step and the torch_dist helpers (reference utils/torch_dist.py:33-43).
The CUDA kernels are not involved; the N>1 GPU path uses the same code with backend `nccl`."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cvpr2023-unidistill_b200")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from unidistill.utils import torch_dist
        from unidistill_b200.step import AsyncNormalisers, FlatGradBucket, clip_and_average_, fused_normalisers

        assert torch_dist.get_world_size() == world and torch_dist.get_rank() == rank
        # reduce_mean == sum / world, input untouched
        t = torch.tensor([float(rank + 1), 10.0 * (rank + 1)])
        r = torch_dist.reduce_mean(t)
        assert torch.allclose(r, torch.tensor([1.5, 15.0])) and t[0] == rank + 1

        # flat bucket: gradients of a tiny model, different data per rank -> identical averaged grads everywhere
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        bucket = FlatGradBucket(net.parameters(), overlap=False)
        assert bucket.numel == sum(p.numel() for p in net.parameters())
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn(7, 6, generator=g)
        bucket.zero()
        assert all(p.grad is None for p in net.parameters())      # steal mode: autograd keeps its own gradient tensors
        net(x).square().sum().backward()
        # the flat buffer is laid out back to front
        local = torch.cat([p.grad.reshape(-1) for p in reversed(list(net.parameters()))]).clone()
        bucket.all_reduce_mean()
        lo, hi = bucket.flat.data_ptr(), bucket.flat.data_ptr() + 4 * bucket.numel
        for p in net.parameters():                       # finish() pointed every .grad at its view of the flat buffer
            assert lo <= p.grad.data_ptr() < hi
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(bucket.flat, sum(gathered) / world, atol=1e-6)
        # legacy mode: .grad ARE the views and backward accumulates into them
        torch.manual_seed(0)
        net_l = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        b_l = FlatGradBucket(net_l.parameters(), overlap=False, steal=False)
        b_l.zero()
        net_l(x).square().sum().backward()
        assert torch.allclose(b_l.flat, local, atol=1e-7)
        for p in net_l.parameters():
            assert lo != p.grad.data_ptr() and b_l.flat.data_ptr() <= p.grad.data_ptr() < b_l.flat.data_ptr() + 4 * b_l.numel

        # overlapped path: 3 buckets (tiny bucket size), all-reduces fired from the gradient hooks during backward,
        # then average + global-norm clip in place == clip_grad_norm_ of the averaged gradients
        torch.manual_seed(0)
        net2 = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        b2 = FlatGradBucket(net2.parameters(), bucket_bytes=16, overlap=True)
        assert len(b2.buckets) >= 3 and sum(hi - lo for lo, hi, _ in b2.buckets) == b2.numel
        for _ in range(2):                                   # two steps: the per-step hook state resets
            b2.zero()
            net2(x).square().sum().backward()
            assert all(b2._fired), "every bucket's all-reduce was issued from the hooks"
            summed = b2.finish()
            assert torch.allclose(summed, sum(gathered), atol=1e-5)
        ref = (sum(gathered) / world).clone()
        ref_norm = ref.norm()
        norm = clip_and_average_(b2.flat, world, 0.1)
        assert torch.allclose(norm, ref_norm, rtol=1e-5)
        assert torch.allclose(b2.flat, ref * min(1.0, 0.1 / (float(ref_norm) + 1e-6)), atol=1e-6)
        assert float(b2.flat.norm()) <= 0.1 + 1e-5
        for p_ in net2.parameters():                         # param.grad still views the (clipped) flat buffer
            assert float(p_.grad.abs().sum()) > 0
        # a parameter that receives no gradient: its bucket is reduced by finish()
        net3 = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(6, 5))
        b3 = FlatGradBucket(net3.parameters(), bucket_bytes=64, overlap=True)
        b3.zero()
        net3[0](x).sum().backward()
        assert not all(b3._fired)
        b3.finish()
        assert all(b3._fired)

        # asynchronous normalisers: issued early, waited for at first use
        pend = AsyncNormalisers({"mask": [torch.ones(2, 3) > 0], "heatmap": [torch.ones(1, 1, 2, 2)]},
                                torch.ones(1, 4, dtype=torch.bool), torch.tensor(float(rank)))
        np_, nh_, nv_, ms_ = pend.get()
        assert float(np_[0]) == 6.0 and float(nh_[0]) == 4.0 and float(nv_) == 4.0 and float(ms_) == pytest.approx(0.5)

        # fused normalisers: one all-reduce for 2*T + 2 scalars
        ret = {"mask": [torch.full((2, 4), float(rank + t)) > 0.5 for t in range(3)],
               "heatmap": [(torch.arange(8).reshape(1, 1, 2, 4) % (t + 2 + rank) == 0).float() for t in range(3)]}
        valid = torch.ones(2, 5, dtype=torch.bool)
        valid[rank, 3:] = False
        num_pos, num_hm, n_valid, mask_sum = fused_normalisers(ret, valid, torch.tensor(2.0 + rank))
        exp_pos = [sum(float((torch.full((2, 4), float(r_ + t)) > 0.5).sum()) for r_ in range(world)) / world
                   for t in range(3)]
        assert torch.allclose(num_pos, torch.tensor(exp_pos))
        assert float(n_valid) == pytest.approx((10 - 2) * 2 / 2 / 1.0 if False else 8.0)
        assert float(mask_sum) == pytest.approx(2.5)
        assert num_hm.shape == (3,)
        open(os.path.join(out_dir, "ok_%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok_0") and os.path.exists(tmp_path / "ok_1")
