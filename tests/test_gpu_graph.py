"""CUDA-graph replay of the dense trunk + target assignment == the eager path (same weights, same inputs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def fp32_cudnn():
    """cuDNN may pick different TF32 algorithms for the eager and the captured launch of the same convolution
    (~3e-4 apart); the comparison below is about the graph plumbing, so both run in plain fp32."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def test_graphed_forward_equals_eager(cuda, native_lib, fp32_cudnn):
    from unidistill_b200 import models, synthetic as S

    torch.manual_seed(0)
    cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
    model = models.BEVFusionCenterHead(cfg).to(cuda).train().use_channels_last(True)
    batch = S.lidar_batch(2, seed=9, points_per_sweep=8000, sweeps=2)
    pts = [batch["points"][b].to(cuda) for b in range(2)]
    gt = torch.cat([batch["gt_boxes"], (batch["gt_labels"].float() + 1).unsqueeze(-1)], dim=-1).to(cuda)
    sd = {k: v.clone() for k, v in model.state_dict().items()}

    def run():
        with torch.no_grad():
            bev, feat, heads = model(pts, None, None, gt, return_feature=True)
        return bev.clone(), feat.clone(), [{k: v.clone() for k, v in h.items()} for h in heads]

    eager = run()
    bufs_eager = {k: v.clone() for k, v in model.state_dict().items()}
    model.load_state_dict(sd)
    model.use_cuda_graph(True)
    first = run()            # captures
    bufs_graph = {k: v.clone() for k, v in model.state_dict().items()}
    second = run()           # replays
    for got in (first, second):
        # ~30 conv + train-mode BN layers: round-off differences between conv algorithms grow to ~1e-4
        assert torch.allclose(got[1], eager[1], rtol=2e-3, atol=2e-4)
        for hg, he in zip(got[2], eager[2]):
            for k in he:
                assert torch.allclose(hg[k], he[k], rtol=5e-3, atol=5e-4), k
    # BatchNorm running statistics advanced exactly once by the first graphed call (warm-up runs left no trace)
    for k in bufs_eager:
        if "running_" in k or "num_batches" in k:
            assert torch.allclose(bufs_graph[k].float(), bufs_eager[k].float(), rtol=1e-4, atol=1e-6), k
