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


def _assert_close(a, b, rtol, atol, what):
    bad = (a - b).abs() > atol + rtol * b.abs()
    if bool(bad.any()):
        d = (a - b).abs()
        i = int(d.argmax())
        raise AssertionError("%s: %d of %d elements differ, max |d| %.3e at flat index %d (%.6g vs %.6g), mean |d| %.3e"
                             % (what, int(bad.sum()), bad.numel(), float(d.max()), i, float(a.flatten()[i]),
                                float(b.flatten()[i]), float(d.mean())))


def test_graphed_forward_equals_eager(cuda, native_lib, fp32_cudnn):
    """The sparse encoder runs once; the dense trunk (BEV backbone + head towers + target assignment) then runs eagerly
    and through the CUDA graph on that same BEV tensor.  (Re-running the sparse encoder would not give the same input:
    its BatchNorm statistics are reduced with atomics, and at random init the 21 train-mode BN layers amplify that
    1-ulp run-to-run difference to ~1e-3 -- see tools/determinism_debug.py.)"""
    from unidistill_b200 import models, synthetic as S

    torch.manual_seed(0)
    cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
    model = models.BEVFusionCenterHead(cfg).to(cuda).train().use_channels_last(True)
    batch = S.lidar_batch(2, seed=9, points_per_sweep=8000, sweeps=2)
    pts = [batch["points"][b].to(cuda) for b in range(2)]
    gt = torch.cat([batch["gt_boxes"], (batch["gt_labels"].float() + 1).unsqueeze(-1)], dim=-1).to(cuda)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        bev = model.extract_bev(pts, None, None).contiguous(memory_format=torch.channels_last)

    def snap(feats, ret):
        return feats.clone(), [{k: v.clone() for k, v in h.items()} for h in ret["multi_head_features"]], \
            {k: v.clone() for k, v in ret["heatmap"].items()} if isinstance(ret["heatmap"], dict) else \
            [v.clone() for v in ret["heatmap"]]

    with torch.no_grad():
        x = model.bev_encoder(bev)
        eager = snap(x[0], model.det_head(x[0], gt))
    bufs_eager = {k: v.clone() for k, v in model.state_dict().items()}
    model.load_state_dict(sd)
    trunk = models.GraphedDenseTrunk(model)
    with torch.no_grad():
        feats, _, ret = trunk(bev, gt)          # captures
        first = snap(feats, ret)
        bufs_graph = {k: v.clone() for k, v in model.state_dict().items()}
        feats, _, ret = trunk(bev, gt)          # replays
        second = snap(feats, ret)
    for got in (first, second):
        _assert_close(got[0], eager[0], 2e-3, 2e-4, "BEV features")
        for hg, he in zip(got[1], eager[1]):
            for k in he:
                _assert_close(hg[k], he[k], 5e-3, 5e-4, k)
        hms_g = got[2].values() if isinstance(got[2], dict) else got[2]
        hms_e = eager[2].values() if isinstance(eager[2], dict) else eager[2]
        for a, b in zip(hms_g, hms_e):
            assert torch.equal(a, b), "target heat-maps differ"
    # BatchNorm running statistics advanced exactly once by the first graphed call (warm-up runs left no trace)
    for k in bufs_eager:
        if ("running_" in k or "num_batches" in k) and k.startswith(("bev_encoder.", "det_head.")):
            assert torch.allclose(bufs_graph[k].float(), bufs_eager[k].float(), rtol=1e-4, atol=1e-6), k


def test_graphed_model_forward_runs(cuda, native_lib):
    """The public switch: model.use_cuda_graph(True) routes gradient-free forwards through the graph."""
    from unidistill_b200 import models, synthetic as S

    torch.manual_seed(0)
    cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
    model = models.BEVFusionCenterHead(cfg).to(cuda).train().use_channels_last(True).use_cuda_graph(True)
    batch = S.lidar_batch(2, seed=9, points_per_sweep=8000, sweeps=2)
    pts = [batch["points"][b].to(cuda) for b in range(2)]
    gt = torch.cat([batch["gt_boxes"], (batch["gt_labels"].float() + 1).unsqueeze(-1)], dim=-1).to(cuda)
    with torch.no_grad():
        for _ in range(2):
            bev, feat, heads = model(pts, None, None, gt, return_feature=True)
    assert model._dense_graph.graph is not None
    assert bev.shape == (2, 256, 128, 128) and feat.shape == (2, 512, 128, 128) and len(heads) == 6
    assert all(bool(torch.isfinite(v).all()) for h in heads for v in h.values())
