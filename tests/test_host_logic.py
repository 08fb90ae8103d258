"""Host-side mirror of the reference API (no GPU): constants, schedules, RNG order, key layout, errors, sharding."""

import math
import os
import socket

import pytest
import numpy as np
import torch

import helpers as H
from bsi_b200 import BSI, Discretization, LogUniform, broadcast_right
from bsi_b200._lib import BsiNativeError
from bsi_b200.distributed import gather_rows, shard_range
from bsi_b200.models import DenoisingDiT, NyquistPositionalEmbedding
from bsi_b200.nn import MLP, FourierFeatures

O = H.O
HYPER = dict(lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, preconditioning="edm")
C32 = O.make_consts(1e-2, 1e6, 2e6)


def test_discretization_host_properties():
    d = Discretization.image_8bit()
    assert (d.min, d.max, d.k) == (-1.0, 1.0, 256)
    assert d.dx == O.GRID_8BIT.width and d.range == O.GRID_8BIT.span
    g = H.load_golden("disc.pt")
    assert torch.equal(d.bin_boundaries(torch.device("cpu"), torch.float32), g["edges32"])
    assert torch.equal(d.to_unit_interval(g["x"]), g["to_unit"])
    assert torch.equal(d.to_8bit_image(g["x"]), g["to_u8"])
    assert torch.equal(Discretization(-1.0, 1.0, 3).bin_boundaries(torch.device("cpu"), torch.float32), g["t_edges3"])


def test_loguniform_and_schedule_tables_match_reference():
    bsi = BSI(torch.nn.Identity(), data_shape=(3, 32, 32), k=256, **HYPER)
    g = H.load_golden("schedule.pt")
    assert bsi.p_lambda.ln_low == g["ln_low"] and bsi.p_lambda.ln_high == g["ln_high"]
    r = g["k256"]
    assert torch.equal(bsi.default_schedule, r["t"])
    assert torch.equal(bsi.p_lambda.icdf(r["t"]), r["lam"])
    assert torch.equal(bsi.p_lambda.cdf(r["lam"]), r["t_back"])
    assert torch.equal(bsi.p_lambda.reciprocal_pdf(r["lam"]), r["inv_pdf"])
    cs, co, ci = bsi._edm_preconditioning(r["t"])
    assert torch.equal(cs, r["c_skip"]) and torch.equal(co, r["c_out"]) and torch.equal(ci, r["c_in"])
    k, lam, coef, c_in, t_rows = bsi._step_table(bsi.default_schedule)
    assert k == 256 and coef.shape == (257, 8)
    assert torch.equal(coef[:256, 3], r["alpha"]) and torch.equal(coef[:256, 4], r["lam"][:256]) and torch.equal(coef[:256, 5], r["lam"][1:])
    assert torch.equal(coef[:256, 0], r["c_skip"][:256]) and float(t_rows[-1]) == 1.0
    assert torch.equal(coef[:256, 2], torch.rsqrt(r["alpha"]))
    assert list(bsi.state_dict()) == [] and bsi.tensor_args == {"device": torch.device("cpu"), "dtype": torch.float32}
    assert isinstance(bsi.p_lambda, LogUniform)


def test_sample_lambda_rng_order_matches_oracle():
    bsi = BSI(torch.nn.Identity(), data_shape=(3, 32, 32), k=8, **HYPER)
    lam = bsi._sample_lambda(3, 5, torch.Generator().manual_seed(9))
    ref = O.draw_ld_lambda(C32, 3, 5, torch.Generator().manual_seed(9), torch.float32)
    assert torch.equal(lam, ref)
    bsi.low_discrepancy_sampling = False
    assert bsi._sample_lambda(3, 5, torch.Generator().manual_seed(9)).shape == (5, 3)  # the reference's transposed branch


def test_elbo_assembly_matches_oracle():
    bsi = BSI(torch.nn.Identity(), data_shape=(3, 32, 32), k=8, **HYPER)
    l_r, l_m = H.det_uniform("a.r", (3, 6)).abs() * 100, H.det_uniform("a.m", (4, 6)).abs() * 1000
    e, b, ex = bsi._assemble_elbo(l_r, l_m, True)
    e2, b2, ex2 = O.combine_elbo(l_r, l_m, 3072, True)
    assert torch.equal(e, e2) and torch.equal(b, b2) and torch.equal(ex["bpd_var"], ex2["bpd_var"])
    with pytest.raises(AssertionError):
        bsi._assemble_elbo(l_r[:1], l_m, True)


def test_error_behaviour_without_gpu():
    bsi = BSI(torch.nn.Identity(), data_shape=(3, 32, 32), k=8, **HYPER)
    with pytest.raises(BsiNativeError):
        bsi.sample(2)
    bsi.preconditioning = "vp"
    with pytest.raises(RuntimeError, match="Unknown preconditioning"):
        bsi._predict_x(torch.zeros(1, 3, 32, 32), torch.zeros(1))
    with pytest.raises(AssertionError):
        broadcast_right(torch.zeros(2, 3), torch.zeros(2))
    with pytest.raises(AssertionError):
        DenoisingDiT((3, 64), 4, 128, 1, 2)
    with pytest.raises(AssertionError):
        FourierFeatures(n_min=1, n_max=2)(torch.zeros(2, 3), dim=-1)
    m = DenoisingDiT((3, 64, 64), 4, 128, 1, 2)
    with pytest.raises(BsiNativeError):
        m.eval().requires_grad_(False)(torch.zeros(1, 3, 64, 64), torch.zeros(1))


def test_transposed_lambda_grid_is_rejected_instead_of_read_out_of_bounds():
    """low_discrepancy_sampling=False draws the reference's [batch, n_samples] grid (bsi/bsi.py:441-445): the kernels index rows
    as (sample, data point), so anything but n_samples == batch must raise before a launch (it used to read out of bounds)."""
    bsi = BSI(torch.nn.Identity(), data_shape=(3, 32, 32), k=8, low_discrepancy_sampling=False, **HYPER)
    x = H.det_images("tl.x", 4, (3, 32, 32))
    with pytest.raises(ValueError, match="transposed"):
        bsi.train_loss(x)
    with pytest.raises(ValueError, match="transposed"):
        bsi.inf_measurement_loss(x, 3)
    ld = BSI(torch.nn.Identity(), data_shape=(3, 32, 32), k=8, **HYPER)
    with pytest.raises((ValueError, BsiNativeError)):  # lambda not ending in the batch axis
        ld._sample_q_mu_lambda(x.cuda() if torch.cuda.is_available() else x, torch.ones(3))


def test_replay_noise_hands_out_recorded_draws_in_order():
    from bsi_b200.bsi import ReplayNoise

    bsi = BSI(torch.nn.Identity(), data_shape=(3, 32, 32), k=8, **HYPER)
    off, perm = torch.tensor(0.25), torch.tensor([3, 1, 0, 2, 5, 4])
    lam = bsi._sample_lambda(2, 3, ReplayNoise([off, perm]))
    assert torch.equal(lam, O.lam_of_t(O.make_consts(1e-2, 1e6, 2e6), O.ld_times(2, 3, off, perm)))
    with pytest.raises(RuntimeError, match="shape"):
        bsi._sample_lambda(2, 3, ReplayNoise([off, perm[:4]]))
    with pytest.raises(RuntimeError, match="exhausted"):
        bsi._sample_lambda(2, 3, ReplayNoise([off]))


def test_dit_state_dict_layout_and_buffers():
    spec = O.DiTSpec((3, 64, 64), 4, 128, 2, 2)
    m = DenoisingDiT(spec.data_shape, spec.patch, spec.dim, spec.depth, spec.heads, dropout=0.05, fourier_features=FourierFeatures(n_min=6, n_max=8, name="fourier"), name="dit")
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == H.dit_shapes(spec)
    assert torch.equal(m.dit.patch_pos_embedding, O.dit_pos_table(spec))
    sc, bi = O.nyquist_tables(128, 1000)
    assert torch.equal(m.dit.t_embedding.scale, sc) and torch.equal(m.dit.t_embedding.bias, bi)
    # adaLN-Zero initialisation like the reference (dit.py:83-85)
    assert float(m.dit.blocks[0].adaLN_modulation[2].weight.abs().max()) == 0.0
    if H.have_reference():
        _, ref_dit, _, _, ref_nn = H.import_reference()
        torch.manual_seed(0)
        r = ref_dit.DenoisingDiT(spec.data_shape, spec.patch, spec.dim, spec.depth, spec.heads, dropout=0.05, fourier_features=ref_nn.FourierFeatures(n_min=6, n_max=8))
        torch.manual_seed(0)
        mine = DenoisingDiT(spec.data_shape, spec.patch, spec.dim, spec.depth, spec.heads, dropout=0.05, fourier_features=FourierFeatures(n_min=6, n_max=8))
        rs, ms = r.state_dict(), mine.state_dict()
        assert list(rs) == list(ms)
        assert all(torch.equal(rs[k], ms[k]) for k in rs), "same seed must give the reference's initial weights"
        mine.load_state_dict(rs)


def test_standalone_modules_match_oracle():
    g = H.load_golden("embed.pt")
    t = torch.tensor([0.0, 1 / 256, 0.5, 1.0])
    assert torch.equal(NyquistPositionalEmbedding(1024, 1000)(t), g["nyq1024_1000"])
    assert torch.equal(NyquistPositionalEmbedding.from_config(32, 100, name="nyquist")(t), g["nyq32_100"])
    x = 1.5 * H.det_uniform("ff.x", (2, 3, 4, 4))
    ff = FourierFeatures(n_min=6, n_max=8)
    assert ff.n_features() == 6 and torch.equal(ff(x, dim=1), g["fourier_6_8"])
    mlp = MLP(8, 4, hidden_features=[16], actfn=torch.nn.GELU)
    assert list(mlp.state_dict()) == ["0.weight", "0.bias", "2.weight", "2.bias"]


def test_shard_range_partition():
    for n in (0, 1, 7, 256, 1025):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert sum(c for _, c in spans) == n
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
            # the reference's per-rank batch rule (bsi/data/h5image.py:312)
            assert [c for _, c in spans] == [n // w + int(r < n % w) for r in range(w)]


def _gloo_worker(rank, world, port, n_total, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = shard_range(n_total, rank, world)
    local = torch.arange(start, start + count, dtype=torch.float32)[:, None] * torch.ones(1, 3)
    full = gather_rows(local, n_total)
    q.put((rank, full[:, 0].tolist()))
    dist.destroy_process_group()


def test_gather_rows_gloo_world2():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert res[0] == res[1] == [float(i) for i in range(7)]


def test_sharded_lambda_grid_is_a_column_slice_of_the_single_gpu_grid():
    """Host logic of sharded_elbo: every rank draws offset + permutation for the WHOLE batch from the same seed and
    keeps its columns; the concatenation over any world size is the single-GPU grid (reference bsi/bsi.py:422-440)."""
    bsi = BSI(torch.nn.Identity(), data_shape=(3, 32, 32), k=8, **HYPER)
    B, n = 11, 3
    full = bsi._sample_lambda(n, B, torch.Generator().manual_seed(5))
    for world in (1, 2, 3, 4, 16):
        cols = []
        for rank in range(world):
            start, count = shard_range(B, rank, world)
            x_local = torch.zeros(count, 3, 32, 32)
            grid = bsi._sample_lambda(n, B, torch.Generator().manual_seed(5))
            cols.append(bsi._shard_columns(grid, x_local, (start, B)))
        assert torch.equal(torch.cat(cols, dim=1), full)


def test_reference_checkpoint_ingestion():
    """A Lightning checkpoint of the reference (state_dict keys model.* / ema_model.ema_model.*, config under "config") loads unchanged."""
    from bsi_b200.checkpoint import build_denoiser, denoiser_state_dict, from_reference_checkpoint

    spec = O.DiTSpec((3, 32, 32), 2, 128, 1, 2)
    sd = H.det_state_dict(H.dit_shapes(spec), seed=3)
    ema = {k: v + 1 for k, v in sd.items()}
    ckpt = {
        "state_dict": {**{"model." + k: v for k, v in sd.items()}, **{"ema_model.ema_model." + k: v for k, v in ema.items()}, "ema_model.step": torch.tensor(5)},
        "config": {
            # resolved config/data/imagenet32.yaml as ConfigInCheckpoint stores it (bsi/lightning/callbacks.py:15-16)
            "data": {"_target_": "bsi.data.imagenet.ImageNetDataModule", "name": "imagenet32", "root": "data/imagenet32", "n": 32},
            "task": {
                "bsi": {"_target_": "bsi.bsi.BSI", "lambda_0": 1e-2, "alpha_M": 1e6, "alpha_R": 2e6, "k": 50, "preconditioning": "edm", "low_discrepancy_sampling": True},
                "model": {"_target_": "bsi.models.dit.DenoisingDiT", "name": "DiT", "patch_size": 2, "dim": 128, "depth": 1, "heads": 2, "dropout": None,
                          "fourier_features": {"_target_": "bsi.nn.FourierFeatures", "name": "fourier", "n_min": 6, "n_max": 8}},
            },
        },
    }
    assert set(denoiser_state_dict(ckpt, "online")) == set(sd)
    assert torch.equal(denoiser_state_dict(ckpt, "ema")["dit.patch_encoder.bias"], ema["dit.patch_encoder.bias"])
    bsi, model = from_reference_checkpoint(ckpt, which="ema", device="cpu")
    assert bsi.k == 50 and bsi.data_shape == (3, 32, 32) and torch.equal(model.state_dict()["dit.patch_encoder.bias"], ema["dit.patch_encoder.bias"])
    from bsi_b200.checkpoint import _data_shape

    assert _data_shape({"data": {"_target_": "bsi.data.imagenet.ImageNetDataModule", "name": "imagenet64", "n": 64}}) == (3, 64, 64)
    assert _data_shape({"data": {"_target_": "bsi.data.cifar10.CIFAR10DataModule", "name": "cifar10", "width": 32, "height": 32}}) == (3, 32, 32)
    with pytest.raises(KeyError):
        _data_shape({"data": {"_target_": "somewhere.Else", "name": "set64"}})  # no guessing from substrings of the name
    unet = build_denoiser({"_target_": "bsi.models.vdm_unet.DenoisingVDMUNet", "name": "unet", "actfn": "silu", "dim": 128, "levels": 1, "dropout": 0.1,
                           "pos_emb_mult": 4, "downsampling_attention": False, "n_attention_heads": 1, "padding_mode": "zeros",
                           "pos_emb": {"name": "nyquist", "size": 32, "expected_rate": 100}, "fourier_features": {"n_min": 6, "n_max": 8}}, (3, 32, 32))
    assert {k: tuple(v.shape) for k, v in unet.state_dict().items()} == H.unet_shapes(O.UNetSpec((3, 32, 32), dim=128, levels=1))


def test_unet_state_dict_layout_matches_reference():
    from bsi_b200.models import DenoisingVDMUNet

    spec = O.UNetSpec((3, 32, 32), dim=128, levels=2)
    mk = lambda: DenoisingVDMUNet(spec.data_shape, NyquistPositionalEmbedding(32, 100), "silu", 128, 2, 4, n_attention_heads=1, dropout=0.1,
                                  fourier_features=FourierFeatures(n_min=6, n_max=8), name="unet")
    m = mk()
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == H.unet_shapes(spec)
    if H.have_reference():
        _, _, ref_unet, ref_pos, ref_nn = H.import_reference()
        torch.manual_seed(0)
        r = ref_unet.DenoisingVDMUNet(spec.data_shape, ref_pos.NyquistPositionalEmbedding(32, 100), "silu", 128, 2, 4, n_attention_heads=1, dropout=0.1,
                                      fourier_features=ref_nn.FourierFeatures(n_min=6, n_max=8))
        torch.manual_seed(0)
        rs, ms = r.state_dict(), mk().state_dict()
        assert list(rs) == list(ms) and all(torch.equal(rs[k], ms[k]) for k in rs)
    with pytest.raises(NotImplementedError):
        DenoisingVDMUNet(spec.data_shape, NyquistPositionalEmbedding(32, 100), "gelu", 128, 2, 4)


# ----- optimizer side: host logic of bsi_b200.optim (no GPU needed) -----
def test_ema_schedule_host_logic_matches_oracle_and_reference_fixture():
    from bsi_b200 import optim as NO
    from oracle import optim_oracle as OO

    g = H.load_golden("optim.pt")
    ema = NO.EMA(torch.nn.Linear(3, 2), include_online_model=False, **H.OPTIM_EMA)
    sched = OO.EMASchedule(**H.OPTIM_EMA)
    step, initted = 0, False
    names = {0: "none", 1: "copy", 2: "lerp"}
    for i in range(H.OPTIM_STEPS):
        mode, w = ema._next_action()
        action, w_ref, step, initted = OO.ema_action(step, initted, sched)
        assert names[mode] == action and w == w_ref
        ema.step += 1  # what update() does to the counters (the kernel call itself needs a GPU)
        ema.initted = True
        assert ema.get_current_decay() == g["decay"][i] == OO.ema_current_decay(step, sched)
    assert ema.get_extra_state() == {"initted": True, "step": H.OPTIM_STEPS}
    created = NO.create_ema(torch.nn.Linear(3, 2), beta=0.9999, update_after_step=1000, update_every=1, power=0.5, inv_gamma=7.0, name="ema")
    assert (created.power, created.inv_gamma, created.include_online_model) == (2 / 3, 1.0, False)  # yaml extras are swallowed (bsi/tasks/bsi.py:73-81)


def test_flat_arena_layout_and_cpu_rejection():
    from bsi_b200 import optim as NO

    arena = NO.FlatArena([torch.Size(s) for s in H.OPTIM_SHAPES], torch.device("cpu"))
    assert arena.offsets == [0, 36, 72, 120] and arena.numel == 124 and arena.numel % 4 == 0
    v = arena.view(2)
    v.fill_(3.0)
    assert v.shape == (4, 4, 3) and float(arena.flat[72:120].sum()) == 144.0 and float(arena.flat.sum()) == 144.0
    with pytest.raises(RuntimeError, match="no CPU path"):
        NO.FlatArena.adopt([torch.nn.Parameter(torch.zeros(4))])


def test_native_denoiser_copies_and_pickles_without_its_native_handles():
    """create_ema deep-copies the model (bsi/tasks/ema_pytorch.py:203-236); a copy must not share (or choke on) the engine handle."""
    import copy
    import ctypes
    import pickle

    from bsi_b200.models import DenoisingDiT

    m = DenoisingDiT((3, 32, 32), 2, 128, 1, 2)
    m._engine, m._packed_sig, m._scratch = ctypes.c_void_p(0), ("stale",), {"workspace": torch.zeros(3)}  # as after a first forward (null handle)
    for c in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert c._engine is None and c._arena is None and c._packed_sig is None and c._scratch == {}
        assert all(torch.equal(a, b) and a.data_ptr() != b.data_ptr() for a, b in zip(m.state_dict().values(), c.state_dict().values()))
    m._engine = None


def test_native_denoiser_is_found_behind_ema_and_ddp_style_wrappers():
    from bsi_b200 import BSI
    from bsi_b200.models import DenoisingDiT

    class Wrapper(torch.nn.Module):
        def __init__(self, attr, inner):
            super().__init__()
            setattr(self, attr, inner)

    m = DenoisingDiT((3, 32, 32), 2, 128, 1, 2)
    hyper = dict(data_shape=(3, 32, 32), lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, k=4, preconditioning="edm")
    assert BSI(m, **hyper)._native() is m
    assert BSI(Wrapper("ema_model", m), **hyper)._native() is m
    assert BSI(Wrapper("module", Wrapper("ema_model", m)), **hyper)._native() is m
    assert BSI(torch.nn.Identity(), **hyper)._native() is None and BSI(Wrapper("module", torch.nn.Identity()), **hyper)._native() is None


def _gloo_exchange_worker(rank, world, port, q):
    """The gradient exchange of AdamW.all_reduce_grads on a CPU arena: some ranges were reduced "during the backward" (here: up
    front), the remaining ones afterwards -- every element must be summed over ranks exactly once."""
    import torch.distributed as dist

    from bsi_b200.optim import FlatArena, unreduced_ranges

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arena = FlatArena([torch.Size(s) for s in H.OPTIM_SHAPES], torch.device("cpu"))
    for i in range(len(H.OPTIM_SHAPES)):
        arena.view(i).copy_(H.optim_grad(i, rank))
    early = [(arena.offsets[1], arena.offsets[2]), (arena.offsets[3], arena.numel)]  # tensors 1 and 3 "finished early"
    for a, b in early:
        dist.all_reduce(arena.flat[a:b])
    for a, b in unreduced_ranges(early, arena.numel):
        dist.all_reduce(arena.flat[a:b])
    q.put((rank, arena.flat.numpy().copy()))
    dist.destroy_process_group()


def test_gradient_exchange_ranges_gloo_world2():
    import torch.multiprocessing as mp

    from bsi_b200.optim import FlatArena, unreduced_ranges

    assert unreduced_ranges([], 10) == [(0, 10)]
    assert unreduced_ranges([(4, 6), (0, 2)], 10) == [(2, 4), (6, 10)]
    assert unreduced_ranges([(0, 10)], 10) == [] and unreduced_ranges([(2, 5), (4, 8)], 8) == [(0, 2)]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    ref = FlatArena([torch.Size(s) for s in H.OPTIM_SHAPES], torch.device("cpu"))
    for i in range(len(H.OPTIM_SHAPES)):
        ref.view(i).copy_(H.optim_grad(i, 0) + H.optim_grad(i, 1))
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], ref.flat.numpy())


def test_training_path_covers_every_parameter_exactly_once():
    """dit_train differentiates the GEMM / LayerNorm parameters itself and leaves the adaLN chain to autograd through `mods`:
    together they must cover every parameter of the model exactly once, in the arena order AdamW.grads_ready relies on."""
    from bsi_b200.models import DenoisingDiT
    from bsi_b200.models.dit_train import _layer_seed, trainable_parameters

    m = DenoisingDiT((3, 32, 32), 2, 128, 3, 2)
    direct = trainable_parameters(m)
    ada = [p for blk in m.dit.blocks for p in blk.adaLN_modulation.parameters()]
    everything = list(m.parameters())
    assert len({id(p) for p in direct}) == len(direct) and not ({id(p) for p in direct} & {id(p) for p in ada})
    assert {id(p) for p in direct} | {id(p) for p in ada} == {id(p) for p in everything}
    index = {id(p): i for i, p in enumerate(everything)}
    for l in range(3):  # a block's eight GEMM tensors are one contiguous run of the parameter list (one all-reduce range)
        run = [index[id(p)] for p in direct[2 + 8 * l : 10 + 8 * l]]
        assert run == list(range(run[0], run[0] + 8))
    assert [index[id(p)] for p in direct[-4:]] == list(range(len(everything) - 4, len(everything)))
    seeds = {_layer_seed(s, k) for s in (0, 1, 12345) for k in range(48)}
    assert len(seeds) == 3 * 48 and all(0 <= x < 2**32 for x in seeds)  # distinct dropout streams per (call, layer, site)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU oracle port; here with a 2-block model to stay fast) prints ONE JSON line with the keys
    the driver reads: impl, metric/unit, value, cpu_baseline describing the run, an e2e object with zero transfer bytes."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--depth", "2"],
                         capture_output=True, text=True, timeout=300, cwd=H.ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "BSI.sample samples/sec" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0 and "workload" in d["config"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a machine without a GPU")
def test_native_arm_fails_loudly_without_a_gpu():
    """No CPU fallback anywhere on the product path: bench.py's native arm and the BSI API refuse to run without CUDA."""
    import subprocess
    import sys

    from bsi_b200 import BSI, Discretization
    from bsi_b200._lib import BsiNativeError

    out = subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300,
                         cwd=H.ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
    bsi = BSI(torch.nn.Identity(), data_shape=(3, 32, 32), lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, k=4, preconditioning="edm",
              discretization=Discretization.image_8bit())
    with pytest.raises((BsiNativeError, RuntimeError)):
        bsi.sample(2)


def test_adaln_chain_backward_matches_autograd():
    """The hand-written backward of the batched adaLN chain (bsi_b200/models/dit_train.py::_AdaLNChain, no optimizer arena attached:
    pure torch) against autograd of the same bf16 computation, block by block (dit.py:79-81,90-92)."""
    from bsi_b200.models.dit_train import _AdaLNChain

    torch.manual_seed(3)
    nl, B, d = 3, 8, 32
    cond = torch.randn(B, d, requires_grad=True)
    ada = []
    for _ in range(nl):
        ada += [torch.randn(d, d, requires_grad=True) * 0.2, torch.randn(d, requires_grad=True) * 0.1,
                torch.randn(6 * d, d, requires_grad=True) * 0.2, torch.randn(6 * d, requires_grad=True) * 0.1]
    ada = [a.detach().requires_grad_(True) for a in ada]
    dmods = torch.randn(nl, B, 6 * d)

    class _NoSink:
        pass

    mods = _AdaLNChain.apply(_NoSink(), cond, *ada)
    assert mods.shape == (nl, B, 6 * d) and mods.dtype == torch.bfloat16
    mods.backward(dmods.to(mods.dtype))
    got = [cond.grad.clone()] + [a.grad.clone() for a in ada]

    cond2 = cond.detach().clone().requires_grad_(True)
    ada2 = [a.detach().clone().requires_grad_(True) for a in ada]
    outs = []
    for l in range(nl):
        w0, b0, w2, b2 = (t.to(torch.bfloat16) for t in ada2[4 * l : 4 * l + 4])
        h = torch.nn.functional.silu(torch.nn.functional.linear(cond2.to(torch.bfloat16), w0, b0))
        outs.append(torch.nn.functional.linear(h, w2, b2))
    ref = torch.stack(outs)
    torch.testing.assert_close(mods.float(), ref.float(), rtol=2e-2, atol=2e-2)
    ref.backward(dmods.to(ref.dtype))
    want = [cond2.grad] + [a.grad for a in ada2]
    for g, w in zip(got, want):
        assert g.shape == w.shape
        rel = float((g.float() - w.float()).norm() / w.float().norm().clamp_min(1e-6))
        assert rel < 3e-2, rel


def test_dropout_mask_restatement_matches_the_header(tmp_path):
    """tests/helpers.py restates csrc/common.cuh's stateless dropout decision (16 random bits per element, elements 2i and 2i + 1 share
    one hash) for the GPU parity tests; here the header's own __host__ functions are compiled with nvcc and compared on the CPU."""
    import ctypes
    import shutil
    import subprocess

    import helpers as H

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "drop.cu"
    src.write_text('#include "common.cuh"\n'
                   'extern "C" int keep_host(unsigned seed, unsigned idx, float p) { return bsi::dropout_keep(seed, idx, bsi::dropout_thresh(p)) ? 1 : 0; }\n'
                   'extern "C" unsigned mix_host(unsigned x) { return bsi::mix32(x); }\n')
    so = tmp_path / "drop.so"
    subprocess.run([nvcc, "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-I", os.path.join(root, "bsi_b200", "csrc"), "-I", os.path.join(root, "include"),
                    "-o", str(so), str(src)], check=True, capture_output=True, timeout=300)
    lib = ctypes.CDLL(str(so))
    lib.keep_host.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float]
    lib.mix_host.argtypes = [ctypes.c_uint32]
    lib.mix_host.restype = ctypes.c_uint32
    xs = torch.tensor([0, 1, 2, 12345, 0xFFFFFFFF, 0x9E3779B9], dtype=torch.int64)
    assert H._mix32(xs).tolist() == [lib.mix_host(int(v)) for v in xs.tolist()]
    idx = torch.cat([torch.arange(0, 4096, dtype=torch.int64), torch.tensor([2**31 - 1, 2**31, 2**32 - 2, 2**32 - 1], dtype=torch.int64)])
    for seed, p in ((77, 0.2), (0xDEADBEEF, 0.05), (5, 0.999)):
        want = [lib.keep_host(seed, int(i), p) for i in idx.tolist()]
        got = H.dropout_keep(seed, idx, p).to(torch.int64).tolist()
        assert got == want, (seed, p)
    frac = 1.0 - sum(lib.keep_host(123, i, 0.05) for i in range(200000)) / 200000.0
    assert abs(frac - 0.05) < 2e-3, frac  # the drop rate is what was asked for


def _wgrad_schedule(n_tiles, k_tiles, total_kb, pairs):
    """Python restatement of the weight-gradient GEMM's work distribution (csrc/wgrad_sm100.cu, splits == 0): per CTA pair the list of
    segments (tile, kb0, kb1).  Host side: workers = min(pairs, steps / 8), H = steps / workers; device side: heads + tails."""
    tiles = n_tiles * k_tiles
    all_steps = tiles * total_kb
    workers = int(min(max(all_steps // 8, 1), pairs))
    H_ = all_steps // workers
    f = workers // tiles
    heads, left = tiles * f, workers - tiles * f
    tail0 = total_kb if left == 0 else f * H_
    tail_len = total_kb - tail0
    tail_total = tiles * tail_len
    out = []
    for w in range(workers):
        segs = []
        if w < heads:
            seg, t = divmod(w, tiles)
            if left == 0:
                kb0, kb1 = total_kb * seg // f, total_kb * (seg + 1) // f
            else:
                kb0, kb1 = seg * H_, seg * H_ + H_
            if kb1 > kb0:
                segs.append((t, kb0, kb1))
        else:
            tw = w - heads
            pos, hi = tail_total * tw // left, tail_total * (tw + 1) // left
            while pos < hi:
                t, off = divmod(pos, tail_len)
                ln = min(hi - pos, tail_len - off)
                segs.append((t, tail0 + off, tail0 + off + ln))
                pos += ln
        out.append(segs)
    return out


def test_wgrad_equal_work_schedule_covers_every_block_step_exactly_once():
    """Every (output tile, k-block) of dW += dY^T X is computed by exactly one CTA pair, for the DiT-L shapes and for awkward ones;
    the busiest pair carries at most 6 % more than the mean on the big shapes (H is rounded down, so the tail pairs get the remainder:
    +4.8 % on the out-projection shape, under 1.5 % on the others) -- whole (tile, split) items on 74 pairs were 16-39 % off."""
    cases = [(12, 4, 512, 74), (4, 4, 512, 74), (16, 4, 512, 74), (4, 16, 512, 74), (12, 4, 1024, 74),  # DiT-L at batch 128 / 256
             (24, 4, 2, 74), (4, 4, 2, 74), (1, 1, 5, 74), (1, 1, 16, 74), (37, 2, 9, 74), (74, 1, 64, 74), (75, 1, 64, 74), (5, 3, 1000, 74),
             (2, 2, 7, 3), (3, 5, 11, 8), (148, 1, 33, 74)]
    for n_tiles, k_tiles, total_kb, pairs in cases:
        sched = _wgrad_schedule(n_tiles, k_tiles, total_kb, pairs)
        seen = np.zeros((n_tiles * k_tiles, total_kb), dtype=np.int32)
        for segs in sched:
            for t, kb0, kb1 in segs:
                assert 0 <= t < n_tiles * k_tiles and 0 <= kb0 < kb1 <= total_kb, (n_tiles, k_tiles, total_kb, pairs, t, kb0, kb1)
                seen[t, kb0:kb1] += 1
        assert (seen == 1).all(), (n_tiles, k_tiles, total_kb, pairs)
        loads = [sum(b - a for _, a, b in segs) for segs in sched]
        if n_tiles * k_tiles * total_kb >= 8192:
            assert max(loads) <= 1.06 * (sum(loads) / len(loads)) + 1, (n_tiles, k_tiles, total_kb, loads)
