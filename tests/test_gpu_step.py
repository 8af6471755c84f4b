"""Distillation training step (unidistill_b200.step.DistillStep == Exp.training_step of the reference's distill
scripts) on the GPU: every loss term against the CPU oracle evaluated on the same student/teacher tensors, the
fused-normaliser path against the per-loss reduce_mean path, and one optimiser step."""
import numpy as np
import pytest
import torch

from oracle import distill as OD

pytestmark = pytest.mark.gpu


def _batch(dev, bs=1):
    from unidistill_b200 import synthetic as S

    lid = S.lidar_batch(bs, seed=5, points_per_sweep=6000, sweeps=2)
    cam = S.camera_batch(bs, seed=5)
    b = {"points": [lid["points"][i].to(dev) for i in range(bs)], "imgs": cam["imgs"].to(dev),
         "mats_dict": {k: v.to(dev) for k, v in cam["mats_dict"].items()},
         "gt_boxes": lid["gt_boxes"].to(dev), "gt_labels": lid["gt_labels"].to(dev)}
    return b


@pytest.fixture(scope="module")
def step_and_batch(cuda, native_lib):
    from unidistill_b200 import models
    from unidistill_b200.step import DistillStep

    torch.manual_seed(0)
    t_cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
    s_cfg = models.make_model_cfg(bev=128, lidar=False, camera=True, img_depth=18)
    teacher = models.BEVFusionCenterHead(t_cfg).to(cuda)
    student = models.BEVFusionCenterHead(s_cfg).to(cuda)
    step = DistillStep(student, teacher, t_cfg["geometry"], lr=1e-4)
    step.keep_features = True
    return step, _batch(cuda), t_cfg["geometry"]


def test_loss_terms_match_oracle(step_and_batch):
    step, batch, geo = step_and_batch
    step.bucket.zero()
    terms = step.losses(batch)
    L = step.last
    cpu = lambda t: t.detach().float().cpu().contiguous()  # noqa: E731
    coords = OD.bev_corners(cpu(L["gt"]), geo["pc_range"], geo["voxel_size"], geo["out_size_factor"])
    assert np.allclose(cpu(L["corners"]).numpy(), coords.numpy(), rtol=1e-6, atol=5e-5)
    valid = OD.valid_prefix(cpu(batch["gt_boxes"]))
    assert torch.equal(valid, L["valid"].cpu())
    lf = OD.feature_distill_loss(cpu(L["feat_s"]), cpu(L["feat_t"]), coords, valid)
    lr = OD.bev_distill_loss(cpu(L["bev_s"]), cpu(L["bev_t"]), coords, valid)
    rs = [{k: cpu(v) for k, v in t.items()} for t in L["resp_s"]]
    rt = [{k: cpu(v) for k, v in t.items()} for t in L["resp_t"]]
    lc, lg = OD.response_distill_loss(rs, rt, cpu(L["gt"]), geo["pc_range"], geo["voxel_size"],
                                      geo["out_size_factor"])
    for name, ref in (("loss_feature", lf), ("loss_bev_rel", lr), ("loss_resp_cls", lc), ("loss_resp_reg", lg)):
        got, ref = float(terms[name]), float(ref)
        assert abs(got - ref) <= 1e-3 * abs(ref) + 1e-7, (name, got, ref)
    total = float(terms["loss_det"]) + 100 * float(lf) + 40 * float(lr) + 10 * (float(lc) + float(lg))
    assert abs(float(terms["loss"]) - total) <= 1e-3 * abs(total)


def test_fused_normalisers_equal_per_loss_reduce_mean(step_and_batch):
    """World size 1: the single fused normaliser vector must give exactly the reference's per-loss behaviour."""
    step, batch, geo = step_and_batch
    from unidistill_b200.step import gt_with_labels

    gt = gt_with_labels(batch["gt_boxes"], batch["gt_labels"])
    torch.manual_seed(1)
    with torch.no_grad():
        a = step.student(batch["points"], batch["imgs"], batch["mats_dict"], gt)[0]["loss"]
        step.bucket.zero()
        b = step.losses(batch)["loss_det"]
    # BN batch statistics are identical (same batch), running stats do not enter a train-mode forward
    assert abs(float(a) - float(b)) <= 1e-5 * abs(float(a))


def test_one_optimizer_step_changes_student_only(step_and_batch):
    step, batch, _ = step_and_batch
    s0 = [p.detach().clone() for p in step.student.parameters()]
    t0 = [p.detach().clone() for p in step.teacher.parameters()]
    tb0 = [b.detach().clone() for b in step.teacher.buffers()]
    terms = step(batch)
    assert np.isfinite(float(terms["loss"]))
    assert float(step.bucket.flat.abs().sum()) > 0
    changed = sum(int(not torch.equal(a, p.detach())) for a, p in zip(s0, step.student.parameters()))
    # the 42 tower-convolution biases sit in front of a train-mode BatchNorm: their gradient is exactly zero (the reference
    # gets round-off noise there), so one step does not move them (center_head.batched_towers)
    assert changed > 0.85 * len(s0)
    assert all(torch.equal(a, p.detach()) for a, p in zip(t0, step.teacher.parameters()))
    step._restore_teacher()
    assert all(torch.equal(a, b) for a, b in zip(tb0, step.teacher.buffers()))


def test_folded_graphed_teacher_equals_plain_teacher(cuda, native_lib):
    """DistillStep folds the frozen teacher's BatchNorm into its convolutions and replays its dense trunk from a CUDA
    graph: same teacher outputs as the untouched eval-mode module (plain fp32 convolutions for the comparison)."""
    import copy

    from unidistill_b200 import models
    from unidistill_b200.step import DistillStep, gt_with_labels

    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(0)
        cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
        teacher = models.BEVFusionCenterHead(cfg).to(cuda)
        for m in teacher.modules():                      # non-trivial running statistics
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
        plain = copy.deepcopy(teacher).eval()
        plain.det_head.dense_head.distill = True
        student = models.BEVFusionCenterHead(cfg).to(cuda)
        step = DistillStep(student, teacher, cfg["geometry"], optimizer=False)
        assert step.folded_bn > 20 and teacher._dense_graph is not None
        batch = _batch(cuda)
        gt = gt_with_labels(batch["gt_boxes"], batch["gt_labels"])
        with torch.no_grad():
            ref = plain(batch["points"], None, None, gt, return_feature=True)
            for _ in range(2):                           # capture, then replay
                got = teacher(batch["points"], None, None, gt, return_feature=True)
        assert torch.allclose(got[0], ref[0], rtol=1e-4, atol=1e-5)          # sparse encoder: eval-mode BN, deterministic
        assert torch.allclose(got[1], ref[1], rtol=1e-3, atol=1e-4)
        for hg, he in zip(got[2], ref[2]):
            for k in he:
                assert torch.allclose(hg[k], he[k], rtol=2e-3, atol=2e-4), k
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_graphed_student_trunk_matches_eager(cuda, native_lib):
    """The student's dense trunk + loss + backward replayed from CUDA graphs (GraphedTrainTrunk) against the same
    student run eagerly: loss terms, every gradient and the BatchNorm running statistics after two steps' worth of
    forward/backward (the first call captures, the second replays)."""
    import copy

    from unidistill_b200 import models
    from unidistill_b200.step import DistillStep

    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(0)
        cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
        teacher = models.BEVFusionCenterHead(cfg).to(cuda)
        student_g = models.BEVFusionCenterHead(cfg).to(cuda)
        student_e = copy.deepcopy(student_g)
        step_g = DistillStep(student_g, teacher, cfg["geometry"], optimizer=False, preset="lidar_exp_distill_fusion")
        step_e = DistillStep(student_e, copy.deepcopy(teacher), cfg["geometry"], optimizer=False,
                             preset="lidar_exp_distill_fusion", student_cuda_graph=False)
        assert student_g._train_graph is not None and student_e._train_graph is None
        batch = _batch(cuda, bs=2)
        for it in range(2):
            out = []
            for st in (step_g, step_e):
                st.bucket.zero()
                terms = st.losses(batch)
                terms["loss"].backward()
                st.bucket.finish()
                out.append(({k: float(v) for k, v in terms.items()}, st.bucket.flat.clone()))
            (tg, gg), (te, ge) = out
            for k in te:
                assert abs(tg[k] - te[k]) <= 1e-4 * abs(te[k]) + 1e-6, (it, k, tg[k], te[k])
            assert float(ge.abs().max()) > 0
            cos = float((gg * ge).sum() / (gg.norm() * ge.norm()))
            assert cos > 0.9999, (it, cos)
            assert float((gg - ge).abs().max()) <= 2e-3 * float(ge.abs().max()), it
        assert student_g._train_graph.fn is not None
        bg = dict(student_g.named_buffers())
        for name, b in student_e.named_buffers():
            assert torch.allclose(bg[name].float(), b.float(), rtol=1e-4, atol=1e-5), name
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_camera_branch_on_a_side_stream_gives_the_same_teacher(cuda, native_lib):
    """start_camera() runs the gradient-free camera branch on a side stream (its scratch buffers are per stream);
    forward(camera_future=...) must see exactly what the in-line branch produces."""
    from unidistill_b200 import models
    from unidistill_b200.step import gt_with_labels

    torch.manual_seed(0)
    cfg = models.make_model_cfg(bev=128, lidar=True, camera=True, img_depth=18)
    teacher = models.BEVFusionCenterHead(cfg).to(cuda).eval()
    teacher.det_head.dense_head.distill = True
    batch = _batch(cuda, bs=2)
    gt = gt_with_labels(batch["gt_boxes"], batch["gt_labels"])
    with torch.no_grad():
        ref = teacher(batch["points"], batch["imgs"], batch["mats_dict"], gt, return_feature=True)
        ref = [ref[0].clone(), ref[1].clone(), [{k: v.clone() for k, v in t.items()} for t in ref[2]]]
        for _ in range(2):
            fut = teacher.start_camera(batch["imgs"], batch["mats_dict"])
            got = teacher(batch["points"], batch["imgs"], batch["mats_dict"], gt, return_feature=True, camera_future=fut)
            torch.cuda.synchronize()
            assert torch.allclose(got[0], ref[0], rtol=1e-5, atol=1e-6)
            assert torch.allclose(got[1], ref[1], rtol=1e-4, atol=1e-5)
            for tg, tr in zip(got[2], ref[2]):
                for k in tr:
                    assert torch.allclose(tg[k], tr[k], rtol=1e-4, atol=1e-5), k


def test_prefetched_geometry_gives_identical_losses(cuda, native_lib):
    """DistillStep.prefetch_geometry builds the next step's voxels and rulebooks on a side stream; the step that consumes
    them must compute exactly what a step that builds them in line computes (LiDAR student and LiDAR teacher share them)."""
    from unidistill_b200 import models
    from unidistill_b200.step import DistillStep

    torch.manual_seed(0)
    cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
    teacher = models.BEVFusionCenterHead(cfg).to(cuda)
    student = models.BEVFusionCenterHead(cfg).to(cuda)
    step = DistillStep(student, teacher, cfg["geometry"], optimizer=False, student_cuda_graph=False)
    batch = _batch(cuda, bs=2)

    def forward_terms():
        step.bucket.zero()
        with torch.no_grad():
            t = step.losses(batch)
        return {k: float(v) for k, v in t.items()}

    ref = forward_terms()
    assert step._prefetched is None
    step.prefetch_geometry(batch["points"])
    assert step._prefetched is not None and "lidar" in step._prefetched[1]
    n_rulebooks = len([k for k in step._prefetched[1]["lidar"]["indice_dict"] if isinstance(k, str)])
    assert n_rulebooks >= 8                       # 4 SubM keys + 4 strided keys of VoxelResBackBone8x
    got = forward_terms()
    assert step._prefetched is None               # consumed
    assert got == ref, (got, ref)
    # a prefetch for OTHER clouds is ignored (identity of the points object decides)
    other = _batch(cuda, bs=2)
    step.prefetch_geometry(other["points"])
    assert forward_terms() == ref
