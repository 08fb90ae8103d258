"""BSI public API with a caller-owned PyTorch denoiser (config 1: README toy Conv2d) vs the CPU oracle and goldens."""

import math

import pytest
import torch

import helpers as H
from gpu_util import dev, report, sync
from bsi_b200 import BSI, Discretization

O = H.O
pytestmark = pytest.mark.gpu
C32 = O.make_consts(1e-2, 1e6, 2e6)
HYPER = dict(lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, preconditioning="edm")


class ToyModel(torch.nn.Module):
    """User-side denoiser of the reference README (README.md:24-31)."""

    def __init__(self):
        super().__init__()
        self.layer = torch.nn.Conv2d(4, 3, 3, padding=1)

    def forward(self, mu, t):
        plane = t.reshape(-1, 1, 1, 1).expand(-1, 1, *mu.shape[-2:])
        return self.layer(torch.cat((mu, plane), dim=1))


def make(k=128, noise="torch"):
    model = ToyModel()
    sd = H.det_state_dict(H.TOY_SHAPES, seed=1, bf16_exact=False)
    model.load_state_dict(sd)
    torch.backends.cudnn.allow_tf32 = False
    bsi = BSI(model.to(dev()), data_shape=(3, 32, 32), k=k, discretization=Discretization.image_8bit(), **HYPER).to(dev())
    bsi.noise_source = noise
    return bsi, sd


def cuda_noise_like_reference(seed, n, k):
    gen = torch.Generator(device=dev()).manual_seed(seed)
    return torch.stack([torch.randn((n, 3, 32, 32), device=dev(), generator=gen) for _ in range(k + 1)]).cpu()


def test_sample_free_running_vs_oracle_config1():
    """Config 1 is not chaotic: the whole k=128 trajectory is compared free-running at the fp32 tier."""
    bsi, sd = make()
    with torch.inference_mode():
        out = bsi.sample(16, torch.Generator(device=dev()).manual_seed(7))
        eps = cuda_noise_like_reference(7, 16, 128)
        ref = O.sample_with_noise(lambda mu, t: O.toy_conv_forward(sd, mu, t), C32, torch.linspace(0.0, 1.0, 129), eps)
    assert out.shape == (16, 3, 32, 32)
    report("config-1 sample (free running, fp32 tier 1e-5)", out, ref, 1e-5, 1e-5)


def test_sample_history_vs_golden_teacher_forced():
    bsi, sd = make()
    g = H.load_golden("toy.pt")["sample"]
    t = torch.linspace(0.0, 1.0, 129)
    with torch.inference_mode():
        for j, i in enumerate(g["steps"].tolist()):
            x_hat = bsi._predict_x(g["mu"][j].to(dev()), t[i].expand(2).to(dev()))
            report(f"_predict_x step {i} vs reference", x_hat, g["x_hat"][j], 1e-5, 1e-5)
        mus, xs, ys = bsi.sample_history(2, torch.Generator(device=dev()).manual_seed(7))
    assert mus.shape == (129, 2, 3, 32, 32) and ys.shape == (128, 2, 3, 32, 32)


def test_losses_vs_oracle_config1():
    bsi, sd = make()
    f = lambda mu, t: O.toy_conv_forward(sd, mu, t)
    x = H.det_images("toy.x", 32, (3, 32, 32), seed=2)
    with torch.inference_mode():
        e, b, ex = bsi.elbo(x.to(dev()), 1, 10, torch.Generator(device=dev()).manual_seed(4))
        gen = torch.Generator(device=dev()).manual_seed(4)
        eps_r = torch.randn((1, 32, 3, 32, 32), device=dev(), generator=gen).cpu()
        off, perm = torch.rand((), device=dev(), generator=gen).cpu(), torch.randperm(320, device=dev(), generator=gen).cpu()
        eps_m = torch.randn((10, 32, 3, 32, 32), device=dev(), generator=gen).cpu()
        l_r = O.recon_loss(f, C32, x, 1, eps_r, O.GRID_8BIT)
        l_m = O.inf_measure_loss(f, C32, x, O.lam_of_t(C32, O.ld_times(10, 32, off, perm)), eps_m)
        e_ref, b_ref, _ = O.combine_elbo(l_r, l_m, 3072)
    report("l_recon", ex["l_recon"], l_r, 1e-4, 1e-2)
    report("l_measure", ex["l_measure"], l_m, 1e-4, 1e-3)
    assert float((b.cpu() - b_ref).abs().max()) < 1e-3
    # estimate_var and finite_elbo shapes / semantics
    with torch.inference_mode():
        e2, b2, ex2 = bsi.elbo(x.to(dev()), 2, 3, torch.Generator(device=dev()).manual_seed(5), estimate_var=True)
        e3, b3, ex3 = bsi.finite_elbo(x.to(dev()), 2, 3, torch.Generator(device=dev()).manual_seed(6))
    assert ex2["bpd_var"].shape == (32,) and bool((ex2["bpd_var"] >= 0).all())
    assert ex3["l_measure"].shape == (3, 32) and torch.isfinite(b3).all()
    bsi.preconditioning = "vp"
    with pytest.raises(RuntimeError):
        bsi.train_loss(x.to(dev()))


def test_train_loss_backward_matches_oracle_autograd():
    bsi, sd = make()
    x = H.det_images("toy.x", 32, (3, 32, 32), seed=2)
    loss = bsi.train_loss(x.to(dev()), torch.Generator(device=dev()).manual_seed(3))
    assert loss.shape == (32,) and loss.requires_grad
    loss.mean().backward()
    gen = torch.Generator(device=dev()).manual_seed(3)
    off, perm = torch.rand((), device=dev(), generator=gen).cpu(), torch.randperm(32, device=dev(), generator=gen).cpu()
    eps = torch.randn((1, 32, 3, 32, 32), device=dev(), generator=gen).cpu()[0]
    w = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.train_loss_with(lambda mu, t: O.toy_conv_forward(w, mu, t), C32, x, O.lam_of_t(C32, O.ld_times(1, 32, off, perm))[0], eps)
    ref.mean().backward()
    report("train_loss values", loss, ref, 1e-4, 1e-6)
    report("grad conv weight", bsi.model.layer.weight.grad, w["layer.weight"].grad, 1e-3, 1e-5)
    report("grad conv bias", bsi.model.layer.bias.grad, w["layer.bias"].grad, 1e-3, 1e-5)


def test_philox_mode_statistics_and_seed_control():
    bsi, sd = make(k=16, noise="philox")
    with torch.inference_mode():
        a = bsi.sample(8, seed=5)
        b = bsi.sample(8, seed=5)
        c = bsi.sample(8, seed=6)
        torch.manual_seed(0)
        d = bsi.sample(8)
        torch.manual_seed(0)
        e = bsi.sample(8)
    assert torch.equal(a, b) and not torch.equal(a, c) and torch.equal(d, e)
    assert torch.isfinite(a).all()


def test_cpu_tensors_are_rejected_loudly():
    from bsi_b200._lib import BsiNativeError

    model = ToyModel()
    bsi = BSI(model, data_shape=(3, 32, 32), k=4, **HYPER)
    with pytest.raises(BsiNativeError):
        bsi.sample(2)
    with pytest.raises(BsiNativeError):
        Discretization.image_8bit().bucketize(torch.zeros(4))


def test_finite_elbo_vs_oracle_with_custom_schedule():
    """finite_elbo (bsi/bsi.py:184-215,249-274; eval_elbo.py -k <int>): same RNG consumption as the reference
    (randn recon -> randint steps -> randn measure) replayed into the oracle, default and cosine-like custom schedule."""
    bsi, sd = make(k=32)
    f = lambda mu, t: O.toy_conv_forward(sd, mu, t)
    x = H.det_images("toy.x", 8, (3, 32, 32), seed=2)
    for name, t_grid in (("default", None), ("custom", torch.sin(torch.linspace(0.0, 1.0, 21) * (math.pi / 2)) ** 2)):
        k = 32 if t_grid is None else len(t_grid) - 1
        with torch.inference_mode():
            e, b, ex = bsi.finite_elbo(x.to(dev()), 2, 5, torch.Generator(device=dev()).manual_seed(11), t=None if t_grid is None else t_grid.to(dev()))
            gen = torch.Generator(device=dev()).manual_seed(11)
            eps_r = torch.randn((2, 8, 3, 32, 32), device=dev(), generator=gen).cpu()
            idx = torch.randint(0, k, (5, 8), device=dev(), generator=gen).cpu()
            eps_m = torch.randn((5, 8, 3, 32, 32), device=dev(), generator=gen).cpu()
            grid = torch.linspace(0.0, 1.0, 33) if t_grid is None else t_grid
            l_r = O.recon_loss(f, C32, x, 2, eps_r, O.GRID_8BIT)
            l_m = O.finite_measure_loss(f, C32, x, grid, idx, eps_m)
            e_ref, b_ref, _ = O.combine_elbo(l_r, l_m, 3072)
        report(f"finite l_measure ({name})", ex["l_measure"], l_m, 1e-4, 1e-3)
        report(f"finite l_recon ({name})", ex["l_recon"], l_r, 1e-4, 1e-2)
        assert float((b.cpu() - b_ref).abs().max()) < 1e-3, name


def test_edge_cases_empty_single_and_ragged_batches():
    """Empty and ragged inputs behave like the reference: sample(0) is an empty tensor of the data shape, odd batch sizes and
    k = 1 work, and a data shape whose element count the vectorised kernels cannot handle fails loudly instead of silently."""
    from bsi_b200._lib import BsiNativeError

    bsi, sd = make(k=4)
    f = lambda mu, t: O.toy_conv_forward(sd, mu, t)
    with torch.inference_mode():
        empty = bsi.sample(0, torch.Generator(device=dev()).manual_seed(1))
        assert empty.shape == (0, 3, 32, 32) and empty.dtype == torch.float32
        for n in (1, 3, 7):
            out = bsi.sample(n, torch.Generator(device=dev()).manual_seed(2))
            gen = torch.Generator(device=dev()).manual_seed(2)
            eps = torch.stack([torch.randn((n, 3, 32, 32), device=dev(), generator=gen) for _ in range(5)]).cpu()
            ref = O.sample_with_noise(f, C32, torch.linspace(0.0, 1.0, 5), eps)
            report(f"sample n={n}", out, ref, 1e-5, 1e-5)
        x = H.det_images("toy.x", 5, (3, 32, 32), seed=2)
        e, b, ex = bsi.elbo(x.to(dev()), 1, 3, torch.Generator(device=dev()).manual_seed(4))
        assert b.shape == (5,) and ex["l_measure"].shape == (3, 5) and torch.isfinite(b).all()
        e0, b0, ex0 = bsi.elbo(x[:0].to(dev()), 1, 2, torch.Generator(device=dev()).manual_seed(4))
        assert b0.shape == (0,) and ex0["l_recon"].shape == (1, 0)
    one, _ = make(k=1)
    with torch.inference_mode():
        assert torch.isfinite(one.sample(2, torch.Generator(device=dev()).manual_seed(3))).all()
    # element counts that are not a multiple of 4 are served by the scalar kernels (test_data_shapes_not_divisible_by_four_...)
    odd = BSI(lambda mu, t: 0.5 * mu, data_shape=(1, 5, 5), k=4, discretization=Discretization.image_8bit(), **HYPER).to(dev())
    with torch.inference_mode():
        s_odd = odd.sample(2, torch.Generator(device=dev()).manual_seed(1))
    assert s_odd.shape == (2, 1, 5, 5) and torch.isfinite(s_odd).all()


@pytest.mark.parametrize("noise", ["philox", "torch"])
def test_elbo_is_invariant_to_how_the_batch_is_sharded(noise):
    """the low-discrepancy grid (offset + permutation over n_m * B points, bsi/bsi.py:422-440) and the noise are drawn
    once for the whole batch and sliced, so a column shard evaluated alone (as a rank of bsi_b200.distributed.sharded_elbo would)
    reproduces the corresponding entries of the single-GPU call bit for bit -- for ragged splits and for finite_elbo too."""
    bsi, sd = make(k=16, noise=noise)
    B = 7
    x = H.det_images("shard.x", B, (3, 32, 32), seed=4).to(dev())
    gen = lambda: torch.Generator(device=dev()).manual_seed(77)
    with torch.inference_mode():
        e, b, ex = bsi.elbo(x, 2, 3, gen(), estimate_var=True)
        fe, fb, fex = bsi.finite_elbo(x, 1, 3, gen())
        for splits in ([(0, 4), (4, 3)], [(0, 3), (3, 2), (5, 2)], [(0, 7)]):
            parts = [bsi.elbo(x[a : a + c], 2, 3, gen(), _shard=(a, B)) for a, c in splits]
            assert torch.equal(torch.cat([p[2]["l_recon"] for p in parts], dim=1), ex["l_recon"])
            assert torch.equal(torch.cat([p[2]["l_measure"] for p in parts], dim=1), ex["l_measure"])
            assert torch.equal(torch.cat([p[1] for p in parts]), b)
            fparts = [bsi.finite_elbo(x[a : a + c], 1, 3, gen(), _shard=(a, B)) for a, c in splits]
            assert torch.equal(torch.cat([p[2]["l_measure"] for p in fparts], dim=1), fex["l_measure"])
        # without a process group sharded_elbo is the single-GPU call
        from bsi_b200.distributed import sharded_elbo

        e2, b2, ex2 = sharded_elbo(bsi, x, 2, 3, 77, estimate_var=True)
        assert torch.equal(b2, b) and torch.equal(ex2["bpd_var"], ex["bpd_var"])


def test_data_shapes_not_divisible_by_four_take_the_scalar_kernels():
    """The reference accepts any data_shape; 3 x 5 x 5 = 75 elements per sample leaves rows unaligned for 16-byte accesses, so every
    row kernel has a scalar variant.  Sampler (free running, config-1 model), ELBO and train_loss gradient vs the oracle; Philox noise of
    an element is the same whichever variant draws it."""
    shape = (3, 5, 5)
    model = ToyModel()
    sd = H.det_state_dict(H.TOY_SHAPES, seed=1, bf16_exact=False)
    model.load_state_dict(sd)
    bsi = BSI(model.to(dev()), data_shape=shape, k=32, discretization=Discretization.image_8bit(), **HYPER).to(dev())
    bsi.noise_source = "torch"
    f = lambda mu, t: O.toy_conv_forward(sd, mu, t)
    with torch.inference_mode():
        out = bsi.sample(6, torch.Generator(device=dev()).manual_seed(7))
        gen = torch.Generator(device=dev()).manual_seed(7)
        eps = torch.stack([torch.randn((6, *shape), device=dev(), generator=gen) for _ in range(33)]).cpu()
        ref = O.sample_with_noise(f, C32, torch.linspace(0.0, 1.0, 33), eps)
    report("odd shape: sample (free running)", out, ref, 1e-5, 1e-5)
    x = H.det_images("odd.x", 5, shape, seed=2)
    with torch.inference_mode():
        e, b, ex = bsi.elbo(x.to(dev()), 2, 3, torch.Generator(device=dev()).manual_seed(4))
        gen = torch.Generator(device=dev()).manual_seed(4)
        eps_r = torch.randn((2, 5, *shape), device=dev(), generator=gen).cpu()
        off, perm = torch.rand((), device=dev(), generator=gen).cpu(), torch.randperm(15, device=dev(), generator=gen).cpu()
        eps_m = torch.randn((3, 5, *shape), device=dev(), generator=gen).cpu()
        l_r = O.recon_loss(f, C32, x, 2, eps_r, O.GRID_8BIT)
        l_m = O.inf_measure_loss(f, C32, x, O.lam_of_t(C32, O.ld_times(3, 5, off, perm)), eps_m)
    report("odd shape: l_recon", ex["l_recon"], l_r, 1e-4, 1e-2)
    report("odd shape: l_measure", ex["l_measure"], l_m, 1e-4, 1e-3)
    # gradient of the training loss through the scalar backward kernel
    gen = torch.Generator(device=dev()).manual_seed(9)
    loss = bsi.train_loss(x.to(dev()), gen).mean()
    grads = torch.autograd.grad(loss, list(bsi.model.parameters()))
    gen = torch.Generator(device=dev()).manual_seed(9)
    off, perm = torch.rand((), device=dev(), generator=gen).cpu(), torch.randperm(5, device=dev(), generator=gen).cpu()
    eps = torch.randn((1, 5, *shape), device=dev(), generator=gen).cpu()[0]
    w, bias = sd["layer.weight"].clone().requires_grad_(True), sd["layer.bias"].clone().requires_grad_(True)
    ref_loss = O.train_loss_with(lambda mu, t: O.toy_conv_forward({"layer.weight": w, "layer.bias": bias}, mu, t), C32, x,
                                 O.lam_of_t(C32, O.ld_times(1, 5, off, perm))[0], eps).mean()
    gw, gb = torch.autograd.grad(ref_loss, [w, bias])
    report("odd shape: train_loss", loss.detach().reshape(1), ref_loss.detach().reshape(1), 1e-4, 1e-5)
    report("odd shape: dL/dw", grads[0], gw, 1e-3, 1e-4 * float(gw.abs().max()))
    # in-kernel Philox: element e of a row is component e & 3 of quad e >> 2 -- the scalar kernel draws what the vector kernel would
    from bsi_b200 import _lib as L

    n, D = 3, 75
    mu = torch.empty(n, D, device=dev())
    s0 = torch.ones(1, device=dev())
    L.check(L.load().bsi_sample_init(L.ptr(mu), L.ptr(s0), L.noise(seed=11, sample_base=2, draw=0), n, D, L.stream_ptr()))
    sync()
    report("odd shape: philox", mu, torch.from_numpy(O.philox_normal(11, 2, n, D, 0)), 1e-4, 1e-4)
