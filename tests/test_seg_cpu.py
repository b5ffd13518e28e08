"""CPU tests of the callers around the hot path: layer-table builder, loss,
optimizer groups, batch sharding and the world_size-2 data-parallel step over gloo.  The DCN slots
are built with dcn='none' here (the DCNv3 op has no CPU path, as in the reference)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from yolo_dual_b200.seg import (CAMVID_CLASS_WEIGHTS, YOLOV5_SEG, YOLOV8_SEG, SegModel, SegmentationLoss,
                                shard_batch, smart_optimizer, train_step, wrap_ddp)


@pytest.mark.parametrize("cfg", [YOLOV5_SEG, YOLOV8_SEG], ids=["yolov5seg", "yolov8seg"])
def test_layer_table_builds_and_runs(cfg):
    torch.manual_seed(0)
    m = SegModel(cfg, dcn="none", img_size=(64, 64)).eval()
    with torch.no_grad():
        y = m(torch.randn(2, 3, 64, 64))
    assert y.shape == (2, 12, 64, 64)
    torch.testing.assert_close(y.sum(1), torch.ones(2, 64, 64), rtol=1e-4, atol=1e-4)  # ends in Softmax


def test_dcn_slots_become_dcnv3_blocks():
    from yolo_dual_b200.blocks import C3_DCNV3
    m = SegModel(YOLOV5_SEG, dcn="dcnv3")
    sites = m.dcn_sites()
    # three sites, C = 128 / 256 / 512 with group_channels 16
    assert [(s.channels, s.group, s.group_channels) for _, s in sites] == [(128, 8, 16), (256, 16, 16), (512, 32, 16)]
    assert sum(isinstance(x, C3_DCNV3) for x in m.modules()) == 3
    # checkpoint-visible names of the reference module survive inside the model
    keys = [k for k in m.state_dict() if "dcnv3." in k]
    for leaf in ("dw_conv.conv.weight", "offset.weight", "mask.bias", "input_proj.weight", "output_proj.bias"):
        assert any(k.endswith(leaf) for k in keys), leaf


def test_loss_matches_formula():
    torch.manual_seed(1)
    pred = torch.randn(2, 12, 8, 8)
    tgt = torch.randint(0, 12, (2, 8, 8))
    w = torch.tensor(CAMVID_CLASS_WEIGHTS)
    crit = SegmentationLoss(12, class_weights=CAMVID_CLASS_WEIGHTS)
    total, (t, ce, dice) = crit(pred, tgt)
    ce_want = F.cross_entropy(pred, tgt, weight=w)
    prob = pred.softmax(1) * w.view(1, -1, 1, 1)
    oh = F.one_hot(tgt, 12).permute(0, 3, 1, 2).float()
    d = (2 * (prob * oh).sum((2, 3)) + 1e-6) / (prob.sum((2, 3)) + oh.sum((2, 3)) + 1e-6)
    torch.testing.assert_close(ce, ce_want)
    torch.testing.assert_close(dice, 1 - d.mean())
    torch.testing.assert_close(total, ce_want + 0.5 * (1 - d.mean()))
    # label maps of another size are resized with 'nearest' (seg_diceloss_yolov5.py:722-727)
    total2, _ = crit(pred, tgt[:, ::2, ::2].contiguous())
    assert torch.isfinite(total2)


def test_optimizer_groups():
    m = SegModel(YOLOV5_SEG, dcn="dcnv3")
    opt = smart_optimizer(m)
    n_b, n_w, n_bn = (len(g["params"]) for g in opt.param_groups)
    assert n_b + n_w + n_bn == sum(1 for p in m.parameters() if p.requires_grad)
    assert opt.param_groups[0].get("weight_decay", 0) == 0 and opt.param_groups[1]["weight_decay"] == 5e-4
    assert opt.param_groups[2]["weight_decay"] == 0 and opt.defaults["nesterov"]
    assert n_b > 0 and n_bn > 0  # DCNv3 Linear biases and its dw_conv BN are there


def test_shard_batch_tiles_the_global_batch():
    for world in (1, 2, 4, 8):
        idx = []
        for r in range(world):
            s = shard_batch(64, r, world)
            idx += list(range(64))[s]
        assert idx == list(range(64))
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)


def test_op_is_batch_shardable_without_exchange(pixel_oracle):
    """The hot path needs no data-path collective: op(batch) == concat(op(shard))."""
    from oracle.dcnv3_oracle import make_inputs
    args = (3, 3, 1, 1, 1, 1, 1, 1, 2, 8, 1.0)
    x, off, m, go = make_inputs(4, 6, 6, 2, 8, dist="unit", seed=1)
    full = pixel_oracle.forward(x, off, m, *args)
    gfull = pixel_oracle.backward(x, off, m, go, *args)
    for world in (2, 4):
        outs, grads = [], [[], [], []]
        for r in range(world):
            s = shard_batch(4, r, world)
            a = [t[s].contiguous() for t in (x, off, m, go)]
            outs.append(pixel_oracle.forward(*a[:3], *args))
            for k, g in enumerate(pixel_oracle.backward(*a, *args)):
                grads[k].append(g)
        assert torch.equal(torch.cat(outs), full)
        for k in range(3):
            assert torch.equal(torch.cat(grads[k]), gfull[k])


def _dp_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        torch.set_num_threads(1)
        model = SegModel(YOLOV5_SEG, dcn="none", img_size=(32, 32))
        crit = SegmentationLoss(12, class_weights=CAMVID_CLASS_WEIGHTS)
        gen = torch.Generator().manual_seed(5)
        imgs = torch.randn(4, 3, 32, 32, generator=gen)
        labels = torch.randint(0, 12, (4, 32, 32), generator=gen)
        s = shard_batch(4, rank, world)
        ddp = wrap_ddp(model)
        opt = smart_optimizer(ddp, lr=0.0)  # lr 0: keep the weights, inspect the reduced grads
        # BatchNorm in eval so per-shard statistics do not differ from the full batch
        ddp.eval()
        loss, _ = train_step(ddp, crit, opt, imgs[s], labels[s])
        grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        torch.save({"loss": loss, "grads": grads}, os.path.join(tmp, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_data_parallel_step_world2_gloo(tmp_path):
    """world=2 over gloo: DDP-averaged grads on the batch shards == grads of world=1 on the full batch
    (CE and Dice are batch means, so shard means average to the full mean for equal shards)."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    for k in r0["grads"]:
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k  # all-reduced: identical on both ranks

    torch.manual_seed(0)
    torch.set_num_threads(1)
    model = SegModel(YOLOV5_SEG, dcn="none", img_size=(32, 32)).eval()
    crit = SegmentationLoss(12, class_weights=CAMVID_CLASS_WEIGHTS)
    gen = torch.Generator().manual_seed(5)
    imgs = torch.randn(4, 3, 32, 32, generator=gen)
    labels = torch.randint(0, 12, (4, 32, 32), generator=gen)
    # the weighted CE normalises by the sum of target weights per shard; average the two shard losses
    losses = []
    for r in range(2):
        sl = shard_batch(4, r, 2)
        loss, _ = crit(model(imgs[sl]), labels[sl])
        losses.append(loss)
    (sum(losses) / 2).backward()
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        torch.testing.assert_close(r0["grads"][k], p.grad, rtol=1e-4, atol=1e-6, msg=lambda m: f"{k}: {m}")
    torch.testing.assert_close((r0["loss"] + r1["loss"]) / 2, (sum(losses) / 2).detach(), rtol=1e-5, atol=1e-6)


def _trainer_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from yolo_dual_b200.trainer import Trainer
        torch.manual_seed(0)
        torch.set_num_threads(1)
        model = SegModel(YOLOV5_SEG, dcn="none", img_size=(32, 32))
        crit = SegmentationLoss(12, class_weights=CAMVID_CLASS_WEIGHTS)
        t = Trainer(model, crit, batch_size=32, epochs=2, ema=True)   # nominal 64 -> accumulate = 2
        assert t.accumulate == 2 and hasattr(t.model, "no_sync")
        gen = torch.Generator().manual_seed(10 + rank)
        stepped = []
        for _ in range(4):  # micro-steps 1 and 3 run inside no_sync(): no all-reduce, gradients accumulate locally
            x = torch.randn(2, 3, 32, 32, generator=gen)
            y = torch.randint(0, 12, (2, 32, 32), generator=gen)
            stepped.append(t.micro_step(x, y)[2])
        torch.save({"stepped": stepped, "w": {k: v.clone() for k, v in model.state_dict().items()}, "ema": t.ema.updates},
                   os.path.join(tmp, f"t{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_trainer_accumulates_under_ddp_world2_gloo(tmp_path):
    """Trainer + DDP + gradient accumulation (the first backward runs inside no_sync(): DDP(static_graph=True) asserts
    there — found on the GPU box under NCCL, pinned here over gloo): optimizer steps every second micro-step, weights
    identical on both ranks afterwards although the ranks saw different data."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_trainer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "t0.pt"), torch.load(tmp_path / "t1.pt")
    assert r0["stepped"] == [False, True, False, True] == r1["stepped"] and r0["ema"] == 2
    for k, v in r0["w"].items():
        if v.dtype.is_floating_point and "running_" not in k:  # BatchNorm statistics are per rank (no SyncBN)
            assert torch.equal(v, r1["w"][k]), k


@pytest.mark.parametrize("cfg", [YOLOV5_SEG, YOLOV8_SEG], ids=["v5", "v8"])
def test_deferred_upsample_is_the_reference_order(cfg):
    """The pointwise tail (1x1 Conv + BN(train) + SiLU, channel Softmax) commutes with the last nearest Upsample:
    outputs and gradients equal the reference's layer order; `state_dict` keys are the same."""
    torch.manual_seed(0)
    a = SegModel(cfg, dcn="none", img_size=(64, 64), defer_upsample=True).double().train()
    b = SegModel(cfg, dcn="none", img_size=(64, 64), defer_upsample=False).double().train()
    assert a._deferred is not None and b._deferred is None
    assert list(a.state_dict()) == list(b.state_dict())
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 3, 64, 64, dtype=torch.double)
    ya, yb = a(x), b(x)
    assert ya.shape == yb.shape == (2, 12, 64, 64)
    torch.testing.assert_close(ya, yb, rtol=1e-10, atol=1e-12)
    w = torch.randn_like(ya)
    (ya * w).sum().backward()
    (yb * w).sum().backward()
    for (n, p), q in zip(a.named_parameters(), b.parameters()):
        assert (p.grad is None) == (q.grad is None), n
        if p.grad is not None:
            torch.testing.assert_close(p.grad, q.grad, rtol=1e-8, atol=1e-10 * float(q.grad.abs().max()) + 1e-14)
    # running_mean identical; running_var differs only by BatchNorm's unbiased factor n/(n-1) in the tail's BN
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if k.endswith("running_mean"):
            torch.testing.assert_close(sa[k], sb[k], rtol=1e-9, atol=1e-12)
    a.eval(), b.eval()
    torch.testing.assert_close(a(x), b(x), rtol=2e-3, atol=1e-6)


def test_lowres_output_and_scaled_loss_on_cpu():
    """lowres=True hands out the map before the deferred Upsample; the unfused loss replicates it itself."""
    torch.manual_seed(0)
    m = SegModel(YOLOV5_SEG, dcn="none", img_size=(64, 64)).eval()
    x = torch.randn(1, 3, 64, 64)
    full = m(x)
    low, scale = m(x, lowres=True)
    assert scale == 4 and low.shape == (1, 12, 16, 16)
    torch.testing.assert_close(F.interpolate(low, scale_factor=4.0, mode="nearest"), full)
    lab = torch.randint(0, 12, (1, 64, 64))
    crit = SegmentationLoss(12, class_weights=CAMVID_CLASS_WEIGHTS)
    torch.testing.assert_close(crit(low, lab, scale)[0], crit(full, lab)[0])
    m8 = SegModel(YOLOV8_SEG, dcn="none", img_size=(64, 64)).eval()
    low8, s8 = m8(x, lowres=True)
    assert s8 == 2 and low8.shape == (1, 12, 32, 32)
    torch.testing.assert_close(F.interpolate(low8, scale_factor=2.0, mode="nearest"), m8(x))
    m8.img_size = [96, 96]                      # output needs the final bilinear resize: nothing to skip
    low8, s8 = m8(x, lowres=True)
    assert s8 == 1 and low8.shape == (1, 12, 96, 96)
    torch.testing.assert_close(low8, m8(x))
