"""CPU unit tests of the framework's components (no GPU, no process group)."""
import os
import pickle

import numpy as np
import pytest
import torch

from theanompi_b200 import ops
from theanompi_b200.models import layers2
from theanompi_b200.models.layers2 import (FC, Constant, Conv, ConvPoolLRN, Crop, Dropout, Flatten, LRN, Normal, Pool, Softmax,
                                           Subtract, extract_weight_types, get_layers, get_params)
from theanompi_b200.ops import reference as ref
from theanompi_b200.parallel.arena import BLOCK, FlatArena
from theanompi_b200.utils.opt import FlatSGD
from theanompi_b200.utils.recorder import Recorder


# ------------------------------------------------------------------ layers / ops
def test_layer_shapes_alexnet_chain():
    layers2.reseed()
    c1 = ConvPoolLRN(input=None, input_shape=(2, 227, 227, 3), filter_shape=(3, 11, 11, 96), convstride=4, padsize=0, group=1,
                     poolsize=3, poolstride=2, b=0.0, lrn=True, printinfo=False)
    c2 = ConvPoolLRN(input=c1, filter_shape=(96, 5, 5, 256), convstride=1, padsize=2, group=2, poolsize=3, poolstride=2, b=0.1,
                     lrn=True, printinfo=False)
    assert c1.output_shape == (2, 27, 27, 96) and c2.output_shape == (2, 13, 13, 256)
    assert len(c2.params) == 4 and c2.weight_type == ["W", "b", "W", "b"]
    fl = Flatten(input=c2, axis=2, printinfo=False)
    fc = FC(input=fl, n_out=32, printinfo=False)
    sm = Softmax(input=fc, n_out=10, printinfo=False)
    ls = get_layers(sm)
    assert [type(l).__name__ for l in ls] == ["ConvPoolLRN", "ConvPoolLRN", "Flatten", "FC", "Softmax"]
    params, wt = get_params(ls)
    assert len(params) == 2 + 4 + 2 + 2 and extract_weight_types(params) == wt


def test_same_seed_same_weights():
    layers2.reseed()
    a = Normal((4, 5)).val.clone()
    layers2.reseed()
    b = Normal((4, 5)).val.clone()
    assert torch.equal(a, b)


def test_conv_layer_matches_torch_and_grads_land_in_arena():
    layers2.reseed()
    conv = Conv(input=None, input_shape=(2, 9, 9, 8), convstride=1, padsize=1, W=Normal((16, 3, 3, 8), std=0.1), b=Constant((16,), 0.1),
                printinfo=False)
    arena = FlatArena(conv.params, conv.weight_type, "cpu", weight_decay=0.0)
    x = torch.randn(2, 9, 9, 8, requires_grad=True)
    y = conv.forward(x)
    want = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), conv.W.val.permute(0, 3, 1, 2), conv.b.val, padding=1)).permute(0, 2, 3, 1)
    assert torch.allclose(y, want, atol=1e-5)
    y.sum().backward()
    assert conv.W.val.grad is None                      # gradients bypass AccumulateGrad …
    gw = arena.views("G")[0]
    assert float(gw.abs().sum()) > 0                    # … and land in the flat G region
    xr = x.detach().clone().requires_grad_(True)
    wr = conv.W.val.detach().clone().requires_grad_(True)
    torch.relu(torch.nn.functional.conv2d(xr.permute(0, 3, 1, 2), wr.permute(0, 3, 1, 2), conv.b.val.detach(), padding=1)).sum().backward()
    assert torch.allclose(gw, wr.grad, atol=1e-4)
    assert torch.allclose(x.grad, xr.grad, atol=1e-4)


def test_lrn_matches_reference_formula():
    x = torch.randn(2, 3, 3, 16) * 5
    y = ops.lrn(x)
    sq = (x ** 2)
    pad = torch.nn.functional.pad(sq, (2, 2))
    s = sum(pad[..., i:i + 16] for i in range(5))
    assert torch.allclose(y, x / (2 + 1e-4 * s) ** 0.75, atol=1e-6)


def test_dropout_reference_semantics():
    Dropout.layers.clear()
    d = Dropout(input=None, input_shape=(4, 1000), prob_drop=0.5, printinfo=False)
    x = torch.ones(4, 1000)
    y = d.forward(x)
    assert set(np.unique(y.numpy()).tolist()) <= {0.0, 1.0}          # train: mask * x (no inverted scaling)
    Dropout.SetDropoutOff()
    assert torch.allclose(d.forward(x), 0.5 * x)                     # eval: (1-p) * x
    Dropout.SetDropoutOn()


def test_crop_and_subtract_layers():
    Crop.layers.clear()
    sub = Subtract(input=None, input_shape=(2, 6, 6, 3), subtract_arr=np.ones((6, 6, 3), np.float32), printinfo=False)
    crop = Crop(input=sub, output_shape=(2, 4, 4, 3), flag_batch=False, printinfo=False)
    x = torch.arange(2 * 6 * 6 * 3, dtype=torch.float32).view(2, 6, 6, 3)
    Crop.SetRandCropOff()
    y = crop.forward(sub.forward(x))
    assert torch.equal(y, (x - 1)[:, 1:5, 1:5, :])                    # centre crop when random cropping is off
    Crop.SetRandCropOn()
    assert crop.forward(sub.forward(x)).shape == (2, 4, 4, 3)


def test_softmax_layer_errors():
    sm = Softmax(input=None, input_shape=(6, 8), n_out=7, printinfo=False)
    x = torch.randn(6, 8)
    y = torch.randint(0, 7, (6,))
    sm.forward(x)
    nll = sm.negative_log_likelihood(y)
    want = torch.nn.functional.cross_entropy(x @ sm.W.val.t() + sm.b.val, y)
    assert abs(float(nll) - float(want)) < 1e-5
    assert 0.0 <= float(sm.errors(y)) <= 1.0 and float(sm.errors_top_x(y)) <= float(sm.errors(y)) + 1e-6


# ------------------------------------------------------------------ arena / optimizer
def test_arena_layout_groups_and_buckets():
    ps = [torch.randn(300, 7), torch.randn(300), torch.randn(2000), torch.randn(5)]
    ps[2].pname, ps[3].pname = "gamma", "beta"
    a = FlatArena(ps, ["W", "b", "b", "b"], "cpu", weight_decay=1e-3)
    assert a.numel % BLOCK == 0 and all(o % BLOCK == 0 for o in a.offsets)
    assert a.exchanged_mask() == [True, True, False, False]
    assert ps[0].data_ptr() == a.W.data_ptr()                           # params are views into W
    bk = a.make_buckets(4096 * 4)
    assert bk[0]["hi"] == a.numel and bk[-1]["lo"] == 0                 # bucket 0 = last params (ready first in backward)
    assert sum(b["hi"] - b["lo"] for b in bk) == a.numel
    sd = a.state_dict()
    a.W.zero_()
    a.load_state_dict(sd)
    assert float(a.W.abs().sum()) > 0


def test_flat_sgd_equals_per_tensor_momentum_sgd():
    torch.manual_seed(0)
    ps = [torch.randn(40, 9), torch.randn(40)]
    ref_w = [p.clone() for p in ps]
    a = FlatArena(ps, ["W", "b"], "cpu", weight_decay=5e-4)
    sgd = FlatSGD(a, 0.9)
    u = [torch.zeros_like(p) for p in ref_w]
    for step in range(3):
        g = [torch.randn_like(p) for p in ref_w]
        for v, gi in zip(a.views("G"), g):
            v.copy_(gi)
        sgd.step(0.01, 1)
        for i, (w, gi) in enumerate(zip(ref_w, g)):                      # opt.py:229-251
            real_grad = gi + 5e-4 * w if i == 0 else gi
            real_lr = 0.01 if i == 0 else 0.02
            u[i] = 0.9 * u[i] + real_grad
            w -= real_lr * u[i]
    for p, w in zip(ps, ref_w):
        assert torch.allclose(p, w, atol=1e-6)


def test_three_bsp_optimizers_agree_on_one_rank_equivalents():
    """BSP_MSGD (aggregate momentum), _BSP_MSGD (aggregate gradient) and the fused single pass
    give the same weights when the 'allreduce' is simulated for k = 2 (ref test-cdd-train idea)."""
    from theanompi_b200.utils import opt

    class M(object):
        use_momentum, use_nesterov_momentum, mu = True, False, 0.9

    def make():
        torch.manual_seed(3)
        ps = [torch.randn(30, 5), torch.randn(30)]
        m = M()
        m.arena = FlatArena(ps, ["W", "b"], "cpu", weight_decay=1e-3, with_recv=True)
        m.shared_lr = opt.SharedScalar(m.arena.hyper, 0, 0.05)
        return m

    torch.manual_seed(9)
    grads = [[torch.randn(2048) for _ in range(2)] for _ in range(3)]       # 3 steps x 2 ranks
    results = []
    for builder in (opt.BSP_MSGD, opt._BSP_MSGD):
        ranks = [make(), make()]
        fns = [builder(m, False, k=2) for m in ranks]
        for step in range(3):
            for m, (pre, post), g in zip(ranks, fns, grads[step]):
                m.arena.G.copy_(g[:m.arena.numel]); pre()
            tot = sum(getattr(m.arena, m._send_region) for m in ranks)
            for m, (pre, post) in zip(ranks, fns):
                m.arena.R.copy_(tot); post()
        assert torch.allclose(ranks[0].arena.W, ranks[1].arena.W)
        results.append(ranks[0].arena.W.clone())
    one = make()
    sgd = FlatSGD(one.arena, 0.9)
    for step in range(3):
        one.arena.G.copy_((grads[step][0] + grads[step][1])[:one.arena.numel])
        sgd.step(0.05, k=2)
    assert torch.allclose(results[0], results[1], atol=1e-5)
    assert torch.allclose(results[0], one.arena.W, atol=1e-5)


def test_easgd_and_gosgd_algebra():
    w, c = torch.randn(100), torch.randn(100)
    w0, c0 = w.clone(), c.clone()
    ref.easgd_elastic(w, c, 0.5)
    assert torch.allclose(w + c, w0 + c0, atol=1e-6)                     # the elastic move conserves w + c
    assert torch.allclose(w, w0 - 0.5 * (w0 - c0))
    a, b = torch.randn(50), torch.randn(50)
    a0 = a.clone()
    ref.gosgd_merge(a, b, 0.25, 0.25)
    assert torch.allclose(a, 0.5 * (a0 + b), atol=1e-6)                  # equal push-sum weights → plain average


# ------------------------------------------------------------------ recorder / checkpoint
def test_recorder_roundtrip_and_cut(tmp_path):
    r = Recorder(None, printFreq=2, modelname="m", verbose=False, device="cpu")
    for i in range(1, 5):
        r.start(); r.end("calc"); r.start(); r.end("comm")
        r.train_error(i, 1.0 / i, 0.5)
        r.print_train_info(i)
    r.val_error(4, 0.3, 0.2, 0.1); r.gather_val_info(); r.print_val_info(4)
    r.save(4, 0.01, filepath=str(tmp_path) + "/")
    with open(tmp_path / "inforec.pkl", "rb") as f:
        d = pickle.load(f)
    assert len(d["train_info"]) == 2 and len(d["all_time"]) == 2 and d["val_info"][0][1:] == [0.3, 0.2, 0.1]
    r2 = Recorder(None, 2, "m", False, device="cpu")
    r2.load(str(tmp_path / "inforec.pkl"))
    r2.cut(1)
    assert len(r2.info_dict["train_info"]) == 1
    assert "mean_time_per_period" in r.summary()


def test_checkpoint_resume_restores_weights_momentum_lr_epoch(tmp_path):
    from theanompi_b200.models.cifar10 import Cifar10_model
    from theanompi_b200.utils.helper_funcs import latest_checkpoint, load_checkpoint, save_model
    cfg = dict(verbose=False, rank=0, size=1, device="cpu", batch_size=16, file_batch_size=16, data_kwargs=dict(n_synthetic=160))
    layers2.reseed()
    m = Cifar10_model(cfg); m.compile_iter_fns("avg")
    rec = Recorder(None, 1000, "c", False, device="cpu")
    for i in range(3):
        m.train_iter(i, rec)
    m.epoch = 4
    m.shared_lr.set_value(0.123)
    save_model(m, str(tmp_path) + "/", verbose=False)
    assert os.path.exists(tmp_path / "W_2_4.npy") and os.path.exists(tmp_path / "ckpt_4.pt")
    assert float(np.load(tmp_path / "lr_4.npy")) == pytest.approx(0.123)       # reference wrote a constant 0 here
    layers2.reseed(999)
    m2 = Cifar10_model(cfg); m2.compile_iter_fns("avg")
    assert not torch.equal(m2.arena.W, m.arena.W)
    nxt = load_checkpoint(m2, latest_checkpoint(str(tmp_path)))
    assert nxt == 5 and m2.shared_lr.get_value() == pytest.approx(0.123)
    assert torch.equal(m2.arena.W, m.arena.W) and torch.equal(m2.arena.U, m.arena.U)


# ------------------------------------------------------------------ launcher / data / loader
def test_tmlauncher_cfg_and_flags(tmp_path):
    from theanompi_b200 import launcher
    cfg = tmp_path / "s.cfg"
    cfg.write_text("# c\nRULE=EASGD\nMODELFILE=theanompi_b200.models.cifar10\nMODELCLASS=Cifar10_model\nDEVICES=cuda0,cuda1,cuda2\n")
    plan = launcher.resolve(launcher.parse_args(["-cfg=%s" % cfg]))
    assert plan["rule"] == "EASGD" and plan["devices"] == ["cuda0", "cuda1", "cuda2"]
    plan = launcher.resolve(launcher.parse_args(["-file=a.b", "-class=C", "-r=BSP", "-s=2", "-bsp_exch_strategy=nccl16", "-bsp_sync_type=avg"]))
    assert plan["devices"] == ["cuda0", "cuda1"] and plan["exch_strategy"] == "nccl16" and plan["sync_type"] == "avg"
    multi = launcher.expand_devices(["node0:cuda0", "node1:cuda0,cuda1"])
    assert multi == ["node0:cuda0", "node1:cuda0", "node1:cuda1"]
    with pytest.raises(SystemExit):
        launcher.parse_args(["-nonsense"])
    with pytest.raises(SystemExit):
        launcher.expand_devices("gpu:x")


def test_hwloc_utils():
    from theanompi_b200.parallel import hwloc_utils as h
    assert h.range_to_list("0-3,8,10-11") == [0, 1, 2, 3, 8, 10, 11]
    topo = "\tGPU0\tGPU1\tCPU Affinity\tNUMA Affinity\nGPU0\t X \tNV18\t0-55,112-167\t0\nGPU1\tNV18\t X \t56-111\t1\n"
    aff = h.parse_topo(topo)
    assert aff[0].startswith("0-55") and aff[1] == "56-111"
    used = h.bind_to_socket_mem(",".join(str(c) for c in sorted(os.sched_getaffinity(0))), label="t")
    assert used and os.environ["CPULIST_t"]
    # explicit memory binding (the reference's set_membind): policy visible in /proc/self/numa_maps-independent way —
    # get_mempolicy is not wrapped, so check the syscall result and that resetting works
    nodes = h.nodes_of_cpus(used)
    if nodes and h.set_membind(nodes[:1]):
        assert h.set_membind([], h.MPOL_DEFAULT)
    assert h.set_membind([], h.MPOL_BIND) is False           # an empty node set is refused, nothing changes


def test_data_extend_shuffle_shard():
    from theanompi_b200.models.data.imagenet import ImageNet_data
    from theanompi_b200.models.data.utils import extend_data
    img, lab = extend_data(0, 4, list(range(10)), list(range(10)))
    assert len(img) == 12 and img[-2:] == [8, 9]
    ds = [ImageNet_data(synthetic=True, n_train_files=10, n_val_files=4, file_batch_size=8) for _ in range(2)]
    for r, d in enumerate(ds):
        d.batch_data(8); d.extend_data(r, 2); d.shuffle_data("train", common_seed=7); d.shard_data("train", r, 2)
    assert ds[0].train_img_shuffle == ds[1].train_img_shuffle                # same permutation on every rank
    assert set(ds[0].train_img_shard).isdisjoint(ds[1].train_img_shard)      # disjoint shards
    assert len(ds[0].train_img_shard) == len(ds[1].train_img_shard) == 5


def test_loader_handoff_identity_cpu():
    """What the trainer reads equals normalise+crop of what the 'file' holds (ref test-paraload-cnmem)."""
    from theanompi_b200.models.data.imagenet import ImageNet_data
    d = ImageNet_data(synthetic=True, n_train_files=3, n_val_files=1, file_batch_size=4, size_hw=32)
    d.batch_data(4)
    ld = d.para_load_init("cpu", 24, 24, rand_crop=False, batch_crop_mirror=False)
    ld.request(d.train_img[0], "val"); ld.request(d.train_img[1], "val")
    b0 = ld.get()
    raw = np.empty((4, 32, 32, 3), np.uint8)
    src = d.read(d.train_img[0], raw)
    raw = src.numpy() if src is not None else raw
    std = np.array([0.229, 0.224, 0.225], np.float32)          # (x - mean) / 255 / img_std  (ref proc_load_mpi.py:99)
    want = ((raw.astype(np.float32) - 127.5) / 255.0 / std)[:, 4:28, 4:28, :]
    assert np.allclose(b0.x.numpy(), want, atol=1e-5)
    b1 = ld.get()
    assert b1.item == d.train_img[1]
    ld.drain(); d.para_load_close()


def test_loader_process_mode_cpu(tmp_path):
    """``TMPI_LOADER=process``: a separate loader process fills the shared-memory ring from real .npy batch files; what the
    trainer reads equals normalise+crop of the files (ref proc_load_mpi.py:16-133, test-paraload-cnmem)."""
    from theanompi_b200.models.data.loader import ParaLoader
    from theanompi_b200.models.data.proc_loader import ProcReader
    arrs = []
    for i in range(3):
        a = np.random.RandomState(i).randint(0, 256, (4, 32, 32, 3), dtype=np.uint8)
        np.save(str(tmp_path / ("b%d.npy" % i)), a)
        arrs.append(a)
    pr = ProcReader((4, 32, 32, 3), depth=2, pin=False)
    ld = ParaLoader(pr.read, "cpu", (4, 32, 32, 3), (24, 24), mean=np.full((32, 32, 3), 127.5, np.float32), std_scale=1 / 255.0,
                    depth=2, rand_crop=False, host_buffers=pr.tensors, on_close=pr.close)
    try:
        for i in range(3):
            ld.request(str(tmp_path / ("b%d.npy" % i)), "val")
            b = ld.get()
            want = ((arrs[i].astype(np.float32) - 127.5) / 255.0)[:, 4:28, 4:28, :]
            assert np.allclose(b.x.numpy(), want, atol=1e-6), i
        with pytest.raises(RuntimeError):
            ld.request(str(tmp_path / "missing.npy"), "val")
    finally:
        ld.close()
    assert pr.proc.poll() is not None                    # child gone


def test_gemm_wave_planning_host_side():
    """The launcher's wave arithmetic (host code of the native extension, no GPU needed): split-K factors must not spill a
    few tiles into an extra wave, and 256-row tiles are only chosen when they do not waste most of a wave."""
    from theanompi_b200.ops import native
    L = native.lib()
    if L is None:
        pytest.skip("native extension not built")
    sms = 148
    # AlexNet conv2 wgrad: 13 tiles, 1458 k-blocks.  ceil(148 / 13) = 12 slices would be 156 CTAs → a 2nd wave for 8 tiles.
    s = L.gemm_plan_splits(13, 1458, sms)
    assert 13 * s <= sms and s == 11
    # conv1 (space-to-depth) wgrad: 5 tiles, 6050 k-blocks → one full wave
    s = L.gemm_plan_splits(5, 6050, sms)
    assert 5 * s <= sms and 5 * (s + 1) > sms
    assert L.gemm_plan_splits(200, 100, sms) == 1            # already more tiles than SMs
    assert L.gemm_plan_splits(10, 4, sms) == 1               # too little K to split
    # conv3 fprop (M = 21632, 3 n-tiles): 255 tall tiles = 2 waves beats 507 = 4 short waves
    assert L.gemm_plan_tall(21632, 3, 1, sms) == 1
    # conv3 dgrad (2 n-tiles): 170 tall tiles = 2 waves loses to 338 = 3 short waves
    assert L.gemm_plan_tall(21632, 2, 1, sms) == 0
    assert L.gemm_plan_tall(128, 32, 1, sms) == 0            # a single m-tile cannot be tall
    assert L.gemm_plan_tall(8192, 64, 0, sms) == 0           # fp32 / split-K outputs never use tall tiles


@pytest.mark.parametrize("groups", [1, 2])
def test_conv_pool_block_single_node_equals_composition_cpu(groups):
    """conv(+ReLU) with the pooling layer folded into the same autograd node (what ``ConvPoolLRN`` builds, so the CUDA
    backward can fuse pool scatter + ReLU mask + bias gradient) == conv followed by ``pool2d`` — CPU reference path."""
    from theanompi_b200 import ops
    pool = (3, 2, 0, "max")

    def make():
        torch.manual_seed(3)
        x = torch.randn(2, 13, 13, 8, requires_grad=True)
        ws = [(torch.randn(6, 3, 3, 8 // groups) * 0.2).requires_grad_(True) for _ in range(groups)]
        bs = [(torch.randn(6) * 0.1).requires_grad_(True) for _ in range(groups)]
        return x, ws, bs

    def run(fold):
        x, ws, bs = make()
        pl = pool if fold else None
        if groups == 1:
            y = ops.conv2d_bias_act(x, ws[0], bs[0], 1, 1, 1, True, pl)
        else:
            y = ops.conv2d_group2_bias_act(x, ws[0], bs[0], ws[1], bs[1], 1, 1, True, pl)
        if not fold:
            y = ops.pool2d(y, *pool)
        torch.manual_seed(4)
        y.backward(torch.randn_like(y))
        return [y.detach(), x.grad] + [w.grad for w in ws] + [b.grad for b in bs]

    for a, b in zip(run(True), run(False)):
        assert torch.allclose(a, b, atol=1e-5), float((a - b).abs().max())


def test_arena_buckets_with_solo_parameters():
    """``make_buckets(solo=…)`` (used by the reduce-scatter strategy): a solo parameter is a bucket of its own, buckets stay
    contiguous, cover the arena exactly once and are ordered last-parameter-first."""
    from theanompi_b200.parallel.arena import FlatArena
    shapes = [(16, 3, 3, 8), (16,), (2048, 1024), (2048,), (1000, 2048), (1000,)]
    params = [torch.randn(s) * 0.1 for s in shapes]
    arena = FlatArena(params, ["W" if len(s) > 1 else "b" for s in shapes], torch.device("cpu"), weight_decay=0.0)
    plain = arena.make_buckets(1 << 30)
    assert len(plain) == 1 and plain[0]["lo"] == 0 and plain[0]["hi"] == arena.numel
    bs = arena.make_buckets(1 << 30, solo={2, 4})
    assert [b["params"] for b in bs] == [[5], [4], [3], [2], [0, 1]]
    assert bs[0]["hi"] == arena.numel and bs[-1]["lo"] == 0
    for a, b in zip(bs, bs[1:]):
        assert a["lo"] == b["hi"]
    assert all(b["lo"] % 1024 == 0 and b["hi"] % 1024 == 0 for b in bs)


def test_parallel_copyto_matches_copyto(tmp_path, monkeypatch):
    from theanompi_b200.models.data.utils import parallel_copyto
    rs = np.random.RandomState(3)
    src = rs.randint(0, 256, (37, 64, 64, 3), dtype=np.uint8)
    f = tmp_path / "b.npy"
    np.save(f, src)
    for threads in (1, 3, 8):
        out = np.zeros_like(src)
        parallel_copyto(out, np.load(f, mmap_mode="r"), threads=threads, min_bytes=0)
        assert np.array_equal(out, src)
    monkeypatch.setenv("TMPI_LOADER_THREADS", "2")
    out = np.zeros_like(src)
    parallel_copyto(out, src, min_bytes=0)
    assert np.array_equal(out, src)


@pytest.mark.parametrize("modelfile,modelclass,cfg,steps", [
    ("theanompi_b200.models.keras_model_zoo.wresnet", "Wide_ResNet",
     dict(batch_size=8, file_batch_size=8, depth=10, widen=1, data_kwargs=dict(n_synthetic=64, synthetic=True)), 2),
    ("theanompi_b200.models.lstm", "LSTM", dict(dim_proj=32, batch_size=8, data_kwargs=dict(n_synthetic=64, n_words=200)), 2),
    ("theanompi_b200.models.lasagne_model_zoo.resnet50", "ResNet50",
     dict(batch_size=2, file_batch_size=2, n_class=8, data_kwargs=dict(n_train_files=2, n_val_files=1, synthetic=True)), 1),
    ("theanompi_b200.models.lasagne_model_zoo.wgan", "WGAN", dict(critic_runs=1, data_kwargs=dict(n_synthetic=128)), 2),
    ("theanompi_b200.models.lasagne_model_zoo.lsgan", "LSGAN", dict(data_kwargs=dict(n_synthetic=128)), 2),
])
def test_zoo_models_step_on_the_cpu_reference_path(modelfile, modelclass, cfg, steps):
    """The native residual / recurrent models (BN, residual add, flat Adam, LSTM sequence node) and the GAN adapters run the same
    model code on the fp32 PyTorch reference ops when there is no GPU: a few steps train and validate."""
    import importlib
    import math
    from theanompi_b200.models import layers2
    from theanompi_b200.utils.recorder import Recorder
    layers2.reseed(); layers2.Dropout.layers.clear(); layers2.Crop.layers.clear()
    base = dict(verbose=False, rank=0, size=1, device="cpu")
    base.update(cfg)
    m = getattr(importlib.import_module(modelfile), modelclass)(base)
    m.compile_iter_fns("avg")
    rec = Recorder(None, 10 ** 6, modelclass, False, device="cpu")
    w0 = m.arena.W.clone()
    c = 0
    for _ in range(steps):
        out = m.train_iter(c, rec)
        c = out if isinstance(out, int) else c + 1
    m.val_iter(c, rec)
    assert math.isfinite(float(rec.train_info["cost"][-1]))
    assert not torch.equal(w0, m.arena.W), "weights did not move"
    m.cleanup()


def test_arena_buckets_partition_property():
    """Whatever the tensor sizes, bucket size, tail size and solo set: the buckets tile the arena exactly once, in reverse layer
    order (bucket 0 = last parameters), contiguous, with every parameter in exactly one bucket — the fused exchange launches one
    kernel per bucket over [lo, hi) and counts grad-ready callbacks per bucket, so a gap or an overlap would silently skip or
    double-apply an update."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=120, deadline=None)
    @given(sizes=st.lists(st.integers(1, 9000), min_size=1, max_size=14),
           bucket_kb=st.integers(1, 96), tail_kb=st.integers(0, 64), solo_seed=st.integers(0, 2 ** 16))
    def prop(sizes, bucket_kb, tail_kb, solo_seed):
        params = [torch.zeros(s) for s in sizes]
        a = FlatArena(params, ["W"] * len(params), "cpu")
        rs = np.random.RandomState(solo_seed)
        solo = set(int(i) for i in np.nonzero(rs.rand(len(sizes)) < 0.2)[0])
        buckets = a.make_buckets(bucket_kb << 10, solo=solo, tail_bytes=tail_kb << 10)
        assert buckets[0]["hi"] == a.numel and buckets[-1]["lo"] == 0
        seen = []
        for prev, b in zip([None] + buckets[:-1], buckets):
            assert b["lo"] < b["hi"]
            if prev is not None:
                assert b["hi"] == prev["lo"]                                  # contiguous, descending
            assert b["lo"] == a.offsets[b["params"][0]]
            assert b["params"] == list(range(b["params"][0], b["params"][-1] + 1))
            end = a.offsets[b["params"][-1] + 1] if b["params"][-1] + 1 < len(sizes) else a.numel
            assert b["hi"] == end
            if set(b["params"]) & solo:
                assert len(b["params"]) == 1
            seen.extend(b["params"])
        assert sorted(seen) == list(range(len(sizes)))

    prop()
