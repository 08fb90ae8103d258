#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the REAL reference.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every fixture stores outputs of the unmodified reference (`bsi.bsi`, `bsi.models.*`,
`bsi.nn.*` imported from /root/reference) on deterministic inputs that the tests can
rebuild bit-identically from tests/helpers.py (numpy Philox), so only small tensors are
committed.  Reference entry points exercised (file:line in the reference):
  Discretization.bucketize / bin_boundaries      bsi/bsi.py:29-35
  LogUniform.icdf/cdf, BSI._edm_preconditioning  bsi/bsi.py:76-84, 390-403
  BSI.sample_history / sample                    bsi/bsi.py:312-373
  BSI.train_loss / elbo / finite_elbo            bsi/bsi.py:152-310
  DenoisingDiT.forward, DenoisingVDMUNet.forward bsi/models/dit.py:225-233, bsi/models/vdm_unet.py:92-100
  NyquistPositionalEmbedding, FourierFeatures    bsi/models/pos_emb.py:77-84, bsi/nn/fourier_features.py:24-36
  EMA.update (+ clip_grad_norm_, torch AdamW)     bsi/tasks/ema_pytorch.py:316-341, config/train.yaml:40, config/task/optimizer/adamw.yaml
"""

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402

O = H.O
ref_bsi, ref_dit, ref_unet, ref_pos, ref_nn = H.import_reference()

torch.set_num_threads(8)
torch.manual_seed(0)

HYPER = dict(lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, preconditioning="edm")


def _compact(obj):
    """Clone tensors so views do not drag their whole base storage into the file."""
    if isinstance(obj, torch.Tensor):
        return obj.detach().clone().contiguous()
    if isinstance(obj, dict):
        return {k: _compact(v) for k, v in obj.items()}
    return obj


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(_compact(obj), path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def load_into(module, sd):
    missing, unexpected = module.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    # only non-persistent buffers may be "missing" (they are not in state_dict at all)
    assert not missing, missing
    return module.eval()


# ---------------------------------------------------------------- A. discretisation
def gen_disc():
    D8 = ref_bsi.Discretization.image_8bit()
    u8 = torch.arange(256, dtype=torch.float32)
    on_grid = u8 * (2 / 255) - 1
    edges32 = D8.bin_boundaries(torch.device("cpu"), torch.float32)
    off = H.det_uniform("disc.off", (4096,)) * 1.2
    near = torch.cat([edges32, torch.nextafter(edges32, torch.tensor(2.0)), torch.nextafter(edges32, torch.tensor(-2.0))])
    x = torch.cat([on_grid, off, near, torch.tensor([-5.0, 5.0, -1.0, 1.0, 0.0])])
    out = {
        "x": x,
        "idx": D8.bucketize(x).to(torch.int16),
        "edges32": edges32,
        "edges64": D8.bin_boundaries(torch.device("cpu"), torch.float64),
        "to_unit": D8.to_unit_interval(x),
        "to_u8": D8.to_8bit_image(x),
        # the reference's own unit tests (tests/test_bsi.py:7-34), evaluated in float64 as conftest sets
        "t_rgb_x": torch.tensor([-0.1, 0.0, 1.0, 1.0 - 1 / 256], dtype=torch.float64),
        "t_rgb_idx": ref_bsi.Discretization(0.0, 1.0, 256).bucketize(
            torch.tensor([-0.1, 0.0, 1.0, 1.0 - 1 / 256], dtype=torch.float64)
        ),
        "t_edges5": ref_bsi.Discretization(-1.0, 1.0, 5).bin_boundaries(torch.device("cpu"), torch.float64),
        "t_edges3": ref_bsi.Discretization(-1.0, 1.0, 3).bin_boundaries(torch.device("cpu"), torch.float32),
    }
    b5 = out["t_edges5"]
    out["t_idx5_at"] = ref_bsi.Discretization(-1.0, 1.0, 5).bucketize(b5)
    out["t_idx5_below"] = ref_bsi.Discretization(-1.0, 1.0, 5).bucketize(b5 - 1e-8)
    save("disc.pt", out)


# ---------------------------------------------------------------- B. schedule / coefficients
def gen_schedule():
    out = {}
    for k in (50, 128, 256):
        bsi = ref_bsi.BSI(torch.nn.Identity(), data_shape=(3, 32, 32), k=k, **HYPER)
        t = bsi.default_schedule
        lam = bsi.p_lambda.icdf(t)
        cs, co, ci = bsi._edm_preconditioning(t)
        out[f"k{k}"] = dict(t=t, lam=lam, alpha=lam.diff(), c_skip=cs, c_out=co, c_in=ci, t_back=bsi.p_lambda.cdf(lam), inv_pdf=bsi.p_lambda.reciprocal_pdf(lam))
        out["ln_low"], out["ln_high"] = bsi.p_lambda.ln_low, bsi.p_lambda.ln_high
    save("schedule.pt", out)


# ---------------------------------------------------------------- C. config 1: README toy conv denoiser
class ToyModel(torch.nn.Module):  # README.md:24-31 (user-side example model, restated)
    def __init__(self):
        super().__init__()
        self.layer = torch.nn.Conv2d(in_channels=4, out_channels=3, kernel_size=3, padding=1)

    def forward(self, mu, t):
        t = torch.movedim(t.expand((1, *mu.shape[-2:], len(t))), -1, 0)
        return self.layer(torch.cat((mu, t), dim=-3))


def gen_toy():
    model = load_into(ToyModel(), H.det_state_dict(H.TOY_SHAPES, seed=1, bf16_exact=False))
    k = 128
    bsi = ref_bsi.BSI(model, data_shape=(3, 32, 32), k=k, discretization=ref_bsi.Discretization.image_8bit(), **HYPER)
    out = {"k": k}
    with torch.inference_mode():
        g = torch.Generator().manual_seed(7)
        mus, xs, ys = bsi.sample_history(2, g)
        g = torch.Generator().manual_seed(7)
        final = bsi.sample(2, g)
        assert torch.equal(final, xs[-1])
        g = torch.Generator().manual_seed(7)
        eps = O.draw_sample_noise(2, (3, 32, 32), k, g)
        steps = [0, 1, 2, 64, 126, 127]
        out["sample"] = dict(
            seed=7, n=2, final=final, steps=torch.tensor(steps),
            mu=mus[steps], x_hat=xs[steps], y=ys[steps], mu_next=mus[[s + 1 for s in steps]],
            eps=eps[[s + 1 for s in steps]], eps0=eps[0], eps_sum=eps.double().sum(), eps_abs_sum=eps.double().abs().sum(),
        )
        # custom (cosine-like) schedule through the t= argument (scripts/generate_samples.py:120-152 style)
        tt = torch.sin(torch.linspace(0, 1, 33) * torch.pi / 2) ** 2
        g = torch.Generator().manual_seed(11)
        out["sample_custom_t"] = dict(seed=11, n=2, t=tt, final=bsi.sample(2, g, t=tt))

        x = H.det_images("toy.x", 8, (3, 32, 32), seed=2)
        g = torch.Generator().manual_seed(3)
        out["train_loss"] = dict(seed=3, loss=bsi.train_loss(x, g))
        g = torch.Generator().manual_seed(4)
        e, b, ex = bsi.elbo(x, 1, 10, g)
        out["elbo_1_10"] = dict(seed=4, elbo=e, bpd=b, l_recon=ex["l_recon"], l_measure=ex["l_measure"])
        g = torch.Generator().manual_seed(5)
        e, b, ex = bsi.elbo(x, 2, 3, g, estimate_var=True)
        out["elbo_2_3_var"] = dict(seed=5, elbo=e, bpd=b, l_recon=ex["l_recon"], l_measure=ex["l_measure"], bpd_var=ex["bpd_var"])
        g = torch.Generator().manual_seed(6)
        e, b, ex = bsi.finite_elbo(x, 2, 3, g)
        out["finite_elbo_2_3"] = dict(seed=6, elbo=e, bpd=b, l_recon=ex["l_recon"], l_measure=ex["l_measure"])
        # the raw loss terms on fixed x_hat so kernels can be pinned without any RNG
        xh = (x[None] + 0.002 * H.det_uniform("toy.xh", (2, *x.shape))).contiguous()
        p = torch.distributions.Normal(xh, torch.full_like(xh, torch.rsqrt(bsi.alpha_R)), validate_args=False)
        D8 = bsi.discretization
        edges = D8.bin_boundaries(x.device, x.dtype)
        idx = D8.bucketize(x)
        cl = torch.where(idx == 0, 0, p.cdf(edges[idx]))
        cr = torch.where(idx == 255, 1, p.cdf(edges[idx + 1]))
        out["recon_terms"] = dict(value=(-torch.log(torch.clamp(cr - cl, min=1e-20))).flatten(2).sum(2))
    save("toy.pt", out)


# ---------------------------------------------------------------- D. DiT
def make_ref_dit(spec):
    ff = None if spec.fourier is None else ref_nn.FourierFeatures(n_min=spec.fourier[0], n_max=spec.fourier[1])
    m = ref_dit.DenoisingDiT(spec.data_shape, spec.patch, spec.dim, spec.depth, spec.heads, dropout=0.05, fourier_features=ff)
    sd = H.det_state_dict(H.dit_shapes(spec), seed=1)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    return load_into(m, sd), sd


def gen_dit():
    out = {}
    specs = {
        "small64": O.DiTSpec((3, 64, 64), 4, 128, 2, 2),
        "small32": O.DiTSpec((3, 32, 32), 2, 128, 2, 2),
        "L2x64": O.DiTSpec((3, 64, 64), 4, 1024, 2, 16),
        "L2x32": O.DiTSpec((3, 32, 32), 2, 1024, 2, 16),
        "nofourier": O.DiTSpec((3, 32, 32), 2, 128, 1, 2, fourier=None),
    }
    with torch.inference_mode():
        for name, spec in specs.items():
            m, _ = make_ref_dit(spec)
            B = 2
            mu = 1.5 * H.det_uniform(f"dit.{name}.mu", (B, *spec.data_shape))
            t = torch.tensor([0.3, 0.9])
            rec = dict(y=m(mu, t))
            if name.startswith("small"):
                x = mu if spec.fourier is None else torch.cat((mu, m.fourier_features(mu, dim=1)), dim=1)
                emb = m.dit.patch_encoder(m.dit.patchify(x)) + m.dit.patch_pos_embedding
                h = emb
                c = m.dit.t_embedding(t)
                for blk in m.dit.blocks:
                    h = blk(h, c)
                rec.update(embed=emb, blocks=h, cond=c, pos=m.dit.patch_pos_embedding)
            out[name] = rec

        # BSI on the small DiT: losses + one trajectory prefix
        spec = specs["small64"]
        m, _ = make_ref_dit(spec)
        bsi = ref_bsi.BSI(m, data_shape=spec.data_shape, k=16, discretization=ref_bsi.Discretization.image_8bit(), **HYPER)
        x = H.det_images("dit.x", 4, spec.data_shape, seed=2)
        g = torch.Generator().manual_seed(21)
        e, b, ex = bsi.elbo(x, 1, 2, g)
        out["bsi_small64"] = dict(elbo_seed=21, elbo=e, bpd=b, l_recon=ex["l_recon"], l_measure=ex["l_measure"])
        g = torch.Generator().manual_seed(22)
        out["bsi_small64"]["train_seed"] = 22
        out["bsi_small64"]["train_loss"] = bsi.train_loss(x, g)
        g = torch.Generator().manual_seed(23)
        mus, xs, ys = bsi.sample_history(2, g)
        g = torch.Generator().manual_seed(23)
        eps = O.draw_sample_noise(2, spec.data_shape, 16, g)
        steps = [0, 1, 8, 15]
        out["bsi_small64"]["traj"] = dict(
            seed=23, steps=torch.tensor(steps), mu=mus[steps], x_hat=xs[steps], mu_next=mus[[s + 1 for s in steps]],
            eps=eps[[s + 1 for s in steps]], last_mu=mus[-1], final=xs[-1],
        )
    save("dit.pt", out)


# ---------------------------------------------------------------- E. U-Net
def gen_unet():
    out = {}
    for name, spec in {"dim64": O.UNetSpec((3, 32, 32), dim=64, levels=2), "dim128": O.UNetSpec((3, 32, 32), dim=128, levels=2)}.items():
        pe = ref_pos.NyquistPositionalEmbedding(spec.pos_size, spec.pos_rate)
        ff = ref_nn.FourierFeatures(n_min=6, n_max=8)
        m = ref_unet.DenoisingVDMUNet(spec.data_shape, pe, "silu", spec.dim, spec.levels, spec.pos_mult, n_attention_heads=1, dropout=0.1, fourier_features=ff)
        sd = H.det_state_dict(H.unet_shapes(spec), seed=1)
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
        load_into(m, sd)
        with torch.inference_mode():
            mu = 1.5 * H.det_uniform("unet.mu", (2, *spec.data_shape))
            t = torch.tensor([0.2, 0.95])
            out[name] = dict(y=m(mu, t))
            if name == "dim128":
                bsi = ref_bsi.BSI(m, data_shape=spec.data_shape, k=8, discretization=ref_bsi.Discretization.image_8bit(), **HYPER)
                x = H.det_images("unet.x", 4, spec.data_shape, seed=2)
                e, b, ex = bsi.elbo(x, 1, 2, torch.Generator().manual_seed(31))
                out["bsi_dim128"] = dict(elbo_seed=31, bpd=b, l_recon=ex["l_recon"], l_measure=ex["l_measure"])
    out["y"] = out["dim64"]["y"]
    save("unet.pt", out)


# ---------------------------------------------------------------- F. embeddings
def gen_embed():
    out = {}
    with torch.inference_mode():
        t = torch.tensor([0.0, 1 / 256, 0.5, 1.0])
        out["nyq1024_1000"] = ref_pos.NyquistPositionalEmbedding(1024, 1000)(t)
        out["nyq32_100"] = ref_pos.NyquistPositionalEmbedding(32, 100)(t)
        x = 1.5 * H.det_uniform("ff.x", (2, 3, 4, 4))
        out["fourier_6_8"] = ref_nn.FourierFeatures(n_min=6, n_max=8)(x, dim=1)
        # reference unit test (tests/models/components/test_fourier_features.py:9-28), float64
        xt = torch.tensor([1.333, -2.718281828459045 / 7], dtype=torch.float64)[None, :, None].expand(2, 2, 3)
        ffm = ref_nn.FourierFeatures(n_min=5, n_max=6).double()
        out["fourier_test_y"] = ffm(xt, dim=1)
    save("embed.pt", out)


def gen_optim():
    """Optimizer side of the training step: clip_grad_norm_(1.0) -> torch.optim.AdamW -> the reference's EMA.update()."""
    from bsi.tasks.ema_pytorch import EMA  # the reference's own class (create_ema's arguments, bsi/tasks/bsi.py:73-81)

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(H.det_uniform(f"optim.p{i}", shp)) for i, shp in enumerate(H.OPTIM_SHAPES)])

    model = Holder()
    sched = H.OPTIM_EMA
    ema = EMA(model, beta=sched["beta"], update_after_step=sched["update_after_step"], update_every=sched["update_every"],
              include_online_model=False, use_foreach=True)
    opt = torch.optim.AdamW(model.parameters(), foreach=False, fused=False, **H.OPTIM_HYPER)
    out = {"params": [], "ema": [], "decay": [], "coef": []}
    for step in range(H.OPTIM_STEPS):
        for i, p in enumerate(model.ps):
            p.grad = H.optim_grad(i, step)
        total = torch.nn.utils.clip_grad_norm_(model.parameters(), H.OPTIM_MAX_NORM)
        out["coef"].append(torch.clamp(H.OPTIM_MAX_NORM / (total + 1e-6), max=1.0))
        opt.step()
        ema.update()
        out["decay"].append(ema.get_current_decay())
        out["params"].append([p.detach().clone() for p in model.ps])
        out["ema"].append([p.detach().clone() for p in ema.ema_model.ps])
    out["exp_avg"] = [opt.state[p]["exp_avg"].clone() for p in model.ps]
    out["exp_avg_sq"] = [opt.state[p]["exp_avg_sq"].clone() for p in model.ps]
    save("optim.pt", out)


# ---------------------------------------------------------------- G. the BASELINE.json sizes at full depth (B = 2..4)
def gen_full():
    """DiT-L/4 x24, DiT-L/2 x24 and the 32-level U-Net of BASELINE.json configs 2-4 through the real reference in fp32:
    forwards, one teacher-forced sampler step and elbo(x[4], 1, 2) with the drawn noise and lambda grid stored, so the GPU
    tests can inject exactly what the reference consumed (RNG order of bsi/bsi.py:224-226,283-284,430-434)."""
    out = {}
    with torch.inference_mode():
        spec64 = O.DiTSpec((3, 64, 64), 4, 1024, 24, 16)
        m64, _ = make_ref_dit(spec64)
        mu = 1.5 * H.det_uniform("full.dit64.mu", (2, *spec64.data_shape))
        t = torch.tensor([0.3, 0.9])
        out["dit64"] = dict(y=m64(mu, t))
        print("dit64 forward done", flush=True)

        bsi = ref_bsi.BSI(m64, data_shape=spec64.data_shape, k=256, discretization=ref_bsi.Discretization.image_8bit(), **HYPER)
        x = H.det_images("full.x", 4, spec64.data_shape, seed=2)
        seed = 41
        e, b, ex = bsi.elbo(x, 1, 2, torch.Generator().manual_seed(seed))
        g = torch.Generator().manual_seed(seed)  # the same draws again, in the reference's order
        eps_r = torch.randn((1, 4, *spec64.data_shape), generator=g)
        off = torch.rand((), generator=g)
        perm = torch.randperm(8, generator=g)
        eps_m = torch.randn((2, 4, *spec64.data_shape), generator=g)
        lam = bsi.p_lambda.icdf(torch.remainder((perm / 9).view(2, 4) + off, 1))
        # the replayed draws reproduce the reference's result bit for bit
        l_m = O.inf_measure_loss(lambda a, b_: m64(a, b_), O.make_consts(1e-2, 1e6, 2e6), x, lam, eps_m)
        assert torch.allclose(l_m, ex["l_measure"], rtol=1e-5), (l_m, ex["l_measure"])
        out["elbo64"] = dict(seed=seed, elbo=e, bpd=b, l_recon=ex["l_recon"], l_measure=ex["l_measure"], eps_r=eps_r, offset=off, perm=perm,
                             eps_m=eps_m, lam=lam)
        print("elbo64 done", b.tolist(), flush=True)

        # one teacher-forced sampler step in the middle of the k = 256 schedule (bsi/bsi.py:331-335)
        tt = bsi.default_schedule
        lam_t = bsi.p_lambda.icdf(tt)
        alpha = lam_t.diff()
        steps = [1, 128, 255]
        xs = H.det_images("full.step.x", 2, spec64.data_shape, seed=3)
        rec = dict(steps=torch.tensor(steps), x_hat=[], mu_next=[])
        for i in steps:
            mu_i = bsi._sample_q_mu_lambda(xs, lam_t[i].expand(2), torch.Generator().manual_seed(100 + i))
            eps = H.det_uniform(f"full.step.eps{i}", (2, *spec64.data_shape)) * 1.7
            x_hat = bsi._predict_x(mu_i, tt[i].repeat(2))
            y = x_hat + torch.rsqrt(alpha[i]) * eps
            mu_next = (alpha[i] * y + lam_t[i] * mu_i) / lam_t[i + 1]
            rec["x_hat"].append(x_hat), rec["mu_next"].append(mu_next)
            rec.setdefault("mu", []).append(mu_i)
        out["step64"] = {k_: (torch.stack(v) if isinstance(v, list) else v) for k_, v in rec.items()}
        print("step64 done", flush=True)
        del m64, bsi

        spec32 = O.DiTSpec((3, 32, 32), 2, 1024, 24, 16)
        m32, _ = make_ref_dit(spec32)
        mu = 1.5 * H.det_uniform("full.dit32.mu", (2, *spec32.data_shape))
        out["dit32"] = dict(y=m32(mu, t))
        del m32
        print("dit32 forward done", flush=True)

        us = O.UNetSpec((3, 32, 32), dim=128, levels=32)
        pe = ref_pos.NyquistPositionalEmbedding(us.pos_size, us.pos_rate)
        mu_ = ref_unet.DenoisingVDMUNet(us.data_shape, pe, "silu", us.dim, us.levels, us.pos_mult, n_attention_heads=1, dropout=0.1,
                                        fourier_features=ref_nn.FourierFeatures(n_min=6, n_max=8))
        load_into(mu_, H.det_state_dict(H.unet_shapes(us), seed=1))
        mu = 1.5 * H.det_uniform("full.unet.mu", (2, *us.data_shape))
        out["unet32"] = dict(y=mu_(mu, torch.tensor([0.2, 0.95])))
        xu = H.det_images("full.unet.x", 4, us.data_shape, seed=2)
        ub = ref_bsi.BSI(mu_, data_shape=us.data_shape, k=256, discretization=ref_bsi.Discretization.image_8bit(), **HYPER)
        e, b, ex = ub.elbo(xu, 1, 2, torch.Generator().manual_seed(43))
        g = torch.Generator().manual_seed(43)
        eps_r = torch.randn((1, 4, *us.data_shape), generator=g)
        off = torch.rand((), generator=g)
        perm = torch.randperm(8, generator=g)
        eps_m = torch.randn((2, 4, *us.data_shape), generator=g)
        out["elbo_unet32"] = dict(seed=43, elbo=e, bpd=b, l_recon=ex["l_recon"], l_measure=ex["l_measure"], eps_r=eps_r, offset=off, perm=perm,
                                  eps_m=eps_m, lam=ub.p_lambda.icdf(torch.remainder((perm / 9).view(2, 4) + off, 1)))
        print("unet done", b.tolist(), flush=True)
    save("full.pt", out)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            globals()["gen_" + name]()
        sys.exit(0)
    gen_disc()
    gen_schedule()
    gen_toy()
    gen_dit()
    gen_unet()
    gen_embed()
    gen_optim()
    gen_full()
