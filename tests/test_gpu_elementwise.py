"""CUDA elementwise / reduction kernels vs the CPU oracle and the reference goldens (through the C ABI)."""

import numpy as np
import pytest
import torch

import helpers as H
from gpu_util import call, dev, report, sync
from bsi_b200 import _lib as L
from bsi_b200 import Discretization

O = H.O
pytestmark = pytest.mark.gpu
C32 = O.make_consts(1e-2, 1e6, 2e6)


def step_table(k):
    t = torch.linspace(0.0, 1.0, k + 1)
    lam, alpha = O.schedule(C32, t)
    cs, co, ci = O.edm_coeffs(C32, t)
    coef = torch.zeros(k + 1, 8)
    coef[:, 0], coef[:, 1] = cs, co
    coef[:k, 2], coef[:k, 3], coef[:k, 4], coef[:k, 5] = torch.rsqrt(alpha), alpha, lam[:k], lam[1:]
    return t, lam, alpha, coef


def ulp_tol(*operands):
    """torch's CPU addcmul is FMA-fused, the CUDA kernels (torch's and ours) round the product first: results may differ
    by one ulp of the largest operand, which cancellation turns into a larger error relative to a small result."""
    return 2.5e-7 * max(float(o.abs().max()) for o in operands)


def test_step_fused_bit_exact_vs_oracle_on_cuda():
    """The oracle's op sequence executed by torch on the SAME GPU (the reference as it runs in production) must be
    reproduced bit for bit by the fused kernel: same fp32 ops, same rounding points, injected noise."""
    k, n, shape = 128, 8, (3, 64, 64)
    D = int(np.prod(shape))
    # tabulate the schedule with torch ops on the device, exactly as bsi_b200.BSI._step_table (and the reference) would
    t_d = torch.linspace(0.0, 1.0, k + 1, device=dev())
    lam_d, alpha_d = O.schedule(C32, t_d)
    cs_d, co_d, _ = O.edm_coeffs(C32, t_d)
    coef_d = torch.zeros(k + 1, 8, device=dev())
    coef_d[:, 0], coef_d[:, 1] = cs_d, co_d
    coef_d[:k, 2], coef_d[:k, 3], coef_d[:k, 4], coef_d[:k, 5] = torch.rsqrt(alpha_d), alpha_d, lam_d[:k], lam_d[1:]
    exact = []
    for i in (0, 3, 64, 127):
        mu = (3.0 * H.det_uniform(f"sx.mu{i}", (n, *shape))).to(dev())
        f = H.det_uniform(f"sx.f{i}", (n, *shape)).to(dev())
        eps = (2.0 * H.det_uniform(f"sx.e{i}", (n, *shape))).to(dev())
        cs, co = coef_d[i, 0].expand(n), coef_d[i, 1].expand(n)
        x_hat = torch.addcmul(O.rpad(cs, mu) * mu, O.rpad(co, mu), f)
        y, mu_next = O.posterior_step(mu, x_hat, eps, alpha_d[i], lam_d[i], lam_d[i + 1])
        mu_k = mu.clone()
        xh_k, y_k = torch.empty_like(mu), torch.empty_like(mu)
        call("bsi_step_fused", L.ptr(mu_k), L.ptr(f), L.ptr(coef_d), None, i, 1, L.noise(eps=eps), L.ptr(xh_k), L.ptr(y_k), n, D, L.stream_ptr())
        sync()
        exact.append((float((xh_k == x_hat).float().mean()), float((y_k == y).float().mean()), float((mu_k == mu_next).float().mean())))
        report(f"cuda-oracle mu' step {i}", mu_k, mu_next, 0.0, ulp_tol(mu, y))
    assert all(min(e) == 1.0 for e in exact), f"fraction of bit-identical elements (x_hat, y, mu') per step: {exact}"


def test_step_fused_matches_oracle_bit_exact():
    k, n, shape = 128, 4, (3, 32, 32)
    t, lam, alpha, coef = step_table(k)
    D = int(np.prod(shape))
    coef_d = coef.to(dev())
    for i in (0, 1, 17, 64, 127):
        mu = 3.0 * H.det_uniform(f"st.mu{i}", (n, *shape))
        f = H.det_uniform(f"st.f{i}", (n, *shape))
        eps = 2.0 * H.det_uniform(f"st.e{i}", (n, *shape))
        cs, co, _ = O.edm_coeffs(C32, t[i].expand(n))
        x_hat = torch.addcmul(O.rpad(cs, mu) * mu, O.rpad(co, mu), f)
        y, mu_next = O.posterior_step(mu, x_hat, eps, alpha[i], lam[i], lam[i + 1])
        mu_d, f_d, eps_d = mu.to(dev()), f.to(dev()), eps.to(dev())
        xh_d, y_d = torch.empty_like(mu_d), torch.empty_like(mu_d)
        call("bsi_step_fused", L.ptr(mu_d), L.ptr(f_d), L.ptr(coef_d), None, i, 1, L.noise(eps=eps_d), L.ptr(xh_d), L.ptr(y_d), n, D, L.stream_ptr())
        sync()
        # same op order with separate roundings -> expected to be bit-identical; 1-ulp slack for CPU FMA contraction
        report(f"x_hat step {i}", xh_d, x_hat, 0.0, ulp_tol(mu, f))
        report(f"y step {i}", y_d, y, 0.0, ulp_tol(x_hat, eps * coef[i, 2]))
        # a 1-ulp difference in y is amplified by the cancellation in alpha*y + lam*mu
        report(f"mu' step {i}", mu_d, mu_next, 0.0, 4 * ulp_tol(mu, y))
        assert (mu_d.cpu() == mu_next).float().mean() > 0.6, "posterior update deviates from the CPU oracle on too many elements"


def test_step_fused_golden_teacher_forced():
    g = H.load_golden("toy.pt")["sample"]
    k = 128
    _, _, _, coef = step_table(k)
    coef_d = coef.to(dev())
    sd = H.det_state_dict(H.TOY_SHAPES, seed=1, bf16_exact=False)
    for j, i in enumerate(g["steps"].tolist()):
        mu = g["mu"][j]
        ci = O.edm_coeffs(C32, torch.linspace(0.0, 1.0, k + 1)[i].expand(2))[2]
        f = O.toy_conv_forward(sd, O.rpad(ci, mu) * mu, torch.linspace(0.0, 1.0, k + 1)[i].expand(2))
        mu_d, f_d, eps_d = mu.to(dev()), f.to(dev()), g["eps"][j].to(dev())
        xh_d = torch.empty_like(mu_d)
        call("bsi_step_fused", L.ptr(mu_d), L.ptr(f_d), L.ptr(coef_d), None, i, 1, L.noise(eps=eps_d), L.ptr(xh_d), None, 2, 3072, L.stream_ptr())
        sync()
        report(f"golden x_hat {i}", xh_d, g["x_hat"][j], 1e-6, 1e-6)
        report(f"golden mu_next {i}", mu_d, g["mu_next"][j], 1e-6, 1e-6)


def test_step_counter_and_advance():
    k, n, D = 8, 2, 64
    _, lam, alpha, coef = step_table(k)
    coef_d = coef.to(dev())
    step = torch.zeros(1, dtype=torch.int32, device=dev())
    mu = H.det_uniform("sc.mu", (n, D))
    f = H.det_uniform("sc.f", (n, D))
    eps = H.det_uniform("sc.e", (n, D))
    mu_d, f_d, eps_d = mu.to(dev()), f.to(dev()), eps.to(dev())
    ref = mu.clone()
    for i in range(3):
        call("bsi_step_fused", L.ptr(mu_d), L.ptr(f_d), L.ptr(coef_d), L.ptr(step), 0, 1, L.noise(eps=eps_d), None, None, n, D, L.stream_ptr())
        call("bsi_step_advance", L.ptr(step), L.stream_ptr())
        xh = coef[i, 0] * ref + coef[i, 1] * f
        _, ref = O.posterior_step(ref, xh, eps, alpha[i], lam[i], lam[i + 1])
    sync()
    assert int(step.item()) == 3
    report("graph-style stepping", mu_d, ref, 1e-6, 1e-6)


def test_sample_init_and_philox_vs_oracle():
    n, D, seed = 8, 3072, 1234567
    mu = torch.empty(n, D, device=dev())
    s0 = torch.tensor([10.0], device=dev())
    call("bsi_sample_init", L.ptr(mu), L.ptr(s0), L.noise(seed=seed, sample_base=5, draw=0), n, D, L.stream_ptr())
    sync()
    z = torch.from_numpy(O.philox_normal(seed, 5, n, D, 0))
    report("philox normal", mu.cpu() / 10.0, z, 1e-4, 1e-4)
    # sharding invariance: rows 4.. of a batch starting at 5 == rows 0.. of a batch starting at 9
    mu2 = torch.empty(4, D, device=dev())
    call("bsi_sample_init", L.ptr(mu2), L.ptr(s0), L.noise(seed=seed, sample_base=9, draw=0), 4, D, L.stream_ptr())
    sync()
    assert torch.equal(mu[4:], mu2)
    # the step kernel uses draw = noise.draw + step
    coef = torch.zeros(4, 8)
    coef[:, 1], coef[:, 2], coef[:, 3], coef[:, 5] = 1.0, 1.0, 1.0, 1.0  # x_hat = f; mu' = x_hat + eps
    zero = torch.zeros(n, D, device=dev())
    m = torch.zeros(n, D, device=dev())
    coef_d = coef.to(dev())
    call("bsi_step_fused", L.ptr(m), L.ptr(zero), L.ptr(coef_d), None, 2, 1, L.noise(seed=seed, sample_base=5, draw=1), None, None, n, D, L.stream_ptr())
    sync()
    report("philox in step", m, torch.from_numpy(O.philox_normal(seed, 5, n, D, 3)), 1e-4, 1e-4)
    big = torch.empty(256, 12288, device=dev())
    one = torch.ones(1, device=dev())
    call("bsi_sample_init", L.ptr(big), L.ptr(one), L.noise(seed=7, sample_base=0, draw=0), 256, 12288, L.stream_ptr())
    sync()
    assert abs(float(big.mean())) < 2e-3 and abs(float(big.std()) - 1) < 2e-3
    assert abs(float((big**4).mean()) - 3.0) < 0.02


def test_q_sample_and_scale_combine():
    n, B, shape = 3, 4, (3, 32, 32)
    D = 3072
    x = H.det_images("q.x", B, shape)
    lam = (10.0 ** (H.det_uniform("q.lam", (n, B)) * 4 + 1)).contiguous()
    eps = H.det_uniform("q.eps", (n, B, *shape)) * 2
    ref = O.q_sample(C32, x, lam, eps)
    t = O.t_of_lam(C32, lam).flatten()
    cs, co, ci = O.edm_coeffs(C32, t)
    flat = lam.flatten()
    gamma, sigma = (flat - C32.lambda_0) / flat, torch.rsqrt(flat)
    mu_d = torch.empty(n * B, D, device=dev())
    in_d = torch.empty_like(mu_d)
    # named device tensors: a temporary would be recycled by the caching allocator before the kernel reads it
    x_d, gamma_d, sigma_d, ci_d, eps_d = x.to(dev()), gamma.to(dev()), sigma.to(dev()), ci.to(dev()), eps.to(dev())
    call("bsi_q_sample", L.ptr(mu_d), L.ptr(in_d), L.ptr(x_d), L.ptr(gamma_d), L.ptr(sigma_d), L.ptr(ci_d),
         L.noise(eps=eps_d), n * B, B, D, L.stream_ptr())
    sync()
    report("q_sample mu", mu_d.reshape(ref.shape), ref, 0.0, ulp_tol(ref, eps * sigma.max()))
    report("q_sample model_in", in_d.reshape(ref.shape), O.rpad(ci.reshape(n, B), ref) * ref, 0.0, ulp_tol(ref, eps * sigma.max()))
    f = H.det_uniform("q.f", (n * B, D))
    xh = torch.empty_like(mu_d)
    cs_d, co_d, f_d = cs.to(dev()), co.to(dev()), f.to(dev())
    call("bsi_edm_combine", L.ptr(xh), L.ptr(mu_d), L.ptr(f_d), L.rowref(cs_d, 1), L.rowref(co_d, 1), None, n * B, D, L.stream_ptr())
    sync()
    mu_c = mu_d.cpu()
    report("edm_combine", xh, torch.addcmul(cs[:, None] * mu_c, co[:, None], f), 0.0, ulp_tol(mu_c, f))
    out = torch.empty_like(mu_d)
    call("bsi_scale_rows", L.ptr(out), L.ptr(mu_d), L.rowref(ci_d, 1), None, n * B, D, L.stream_ptr())
    sync()
    assert torch.equal(out.cpu(), ci[:, None] * mu_c)


def test_bucketize_bit_exact_golden():
    g = H.load_golden("disc.pt")
    disc = Discretization.image_8bit()
    x = g["x"].to(dev())
    idx = disc.bucketize(x)
    assert idx.dtype == torch.int64
    assert torch.equal(idx.cpu().to(torch.int16), g["idx"])
    assert torch.equal(disc.bucketize_u8(x).cpu().to(torch.int16), g["idx"])
    assert torch.equal(disc.bin_boundaries(dev(), torch.float32).cpu(), g["edges32"])
    # full-size property: every 8-bit grid value maps to its own index, on a cfg-4 sized batch
    u8 = torch.randint(0, 256, (256, 3, 64, 64), device=dev(), generator=torch.Generator(device=dev()).manual_seed(2))
    xg = u8.float() * (2 / 255) - 1
    assert torch.equal(disc.bucketize(xg), u8.to(torch.int64))
    # reference unit tests restated on the CUDA path (fp32)
    d01 = Discretization(0.0, 1.0, 256)
    assert d01.bucketize(torch.tensor([-0.1, 0.0, 1.0, 1.0 - 1 / 256], device=dev())).tolist() == [0, 0, 255, 254]
    d5 = Discretization(-1.0, 1.0, 5)
    b = d5.bin_boundaries(dev(), torch.float32)
    assert d5.bucketize(b)[:-1].tolist() == list(range(5))
    assert d5.bucketize(b - 1e-6)[1:].tolist() == list(range(5))
    assert torch.equal(d5.bucketize(torch.empty(0, device=dev())), torch.empty(0, dtype=torch.int64, device=dev()))


def test_to_8bit_image_bit_exact():
    g = H.load_golden("disc.pt")
    disc = Discretization.image_8bit()
    assert torch.equal(disc.to_8bit_image(g["x"].to(dev())).cpu(), g["to_u8"])
    x = 1.3 * H.det_uniform("u8.x", (4, 3, 64, 64))
    assert torch.equal(disc.to_8bit_image(x.to(dev())).cpu(), O.to_uint8(x, O.GRID_8BIT))


def test_recon_and_sqerr_reduce_vs_oracle_and_golden():
    g = H.load_golden("toy.pt")["recon_terms"]
    B, shape, D = 8, (3, 32, 32), 3072
    x = H.det_images("toy.x", B, shape, seed=2)
    xh = (x[None] + 0.002 * H.det_uniform("toy.xh", (2, *x.shape))).contiguous()
    disc = Discretization.image_8bit()
    edges = disc.bin_boundaries(dev(), torch.float32)
    inv_scale = float(torch.rsqrt(C32.alpha_R).reciprocal())
    out = torch.empty(2 * B, device=dev())
    x_d, xh_d = x.to(dev()), xh.reshape(2 * B, D).to(dev())
    call("bsi_recon_reduce", L.ptr(out), L.ptr(x_d), None, L.ptr(xh_d), None, None, L.ptr(edges), 256,
         disc.range[0], disc.dx, inv_scale, 2 * B, B, D, L.stream_ptr())
    sync()
    report("recon golden", out.reshape(2, B), g["value"], 2e-5, 1e-3)
    # with the EDM combine fused in: x_hat = c_skip*mu + c_out*f
    R = 2 * B
    mu = xh.reshape(R, D) + 0.001 * H.det_uniform("r.mu", (R, D))
    cs = 0.99 + 0.01 * H.det_uniform("r.cs", (R,)).abs()
    co = 0.01 * (1 + H.det_uniform("r.co", (R,)).abs())
    f = (xh.reshape(R, D) - cs[:, None] * mu) / co[:, None]
    xh2 = torch.addcmul(cs[:, None] * mu, co[:, None], f)
    ref = O.recon_terms(C32, x, xh2.reshape(2, B, *shape), O.GRID_8BIT)
    mu_d, f_d, cs_d, co_d = mu.to(dev()), f.to(dev()), cs.to(dev()), co.to(dev())
    call("bsi_recon_reduce", L.ptr(out), L.ptr(x_d), L.ptr(mu_d), L.ptr(f_d), L.ptr(cs_d), L.ptr(co_d),
         L.ptr(edges), 256, disc.range[0], disc.dx, inv_scale, R, B, D, L.stream_ptr())
    sync()
    report("recon fused combine", out.reshape(2, B), ref, 2e-5, 1e-3)
    call("bsi_sqerr_reduce", L.ptr(out), L.ptr(x_d), L.ptr(mu_d), L.ptr(f_d), L.ptr(cs_d), L.ptr(co_d), R, B, D, L.stream_ptr())
    sync()
    ref_sq = (x[None] - xh2.reshape(2, B, *shape)).square().flatten(2).sum(2)
    report("sqerr", out.reshape(2, B), ref_sq, 1e-5, 1e-9)
    w = H.det_uniform("r.w", (R,)).abs() + 0.5
    gf = torch.empty(R, D, device=dev())
    w_d = w.to(dev())
    call("bsi_sqerr_backward", L.ptr(gf), L.ptr(w_d), L.ptr(x_d), L.ptr(mu_d), L.ptr(f_d), L.ptr(cs_d), L.ptr(co_d), R, B, D, L.stream_ptr())
    sync()
    fr = f.clone().requires_grad_(True)
    loss = (w * (x.repeat(2, 1, 1, 1).reshape(R, D) - (cs[:, None] * mu + co[:, None] * fr)).square().sum(1)).sum()
    loss.backward()
    report("sqerr backward", gf, fr.grad, 1e-5, 1e-9)


def test_abi_rejects_bad_arguments():
    lib = L.load()
    t = torch.zeros(8, device=dev())
    assert lib.bsi_step_fused(L.ptr(t), L.ptr(t), L.ptr(t), None, 0, 1, L.noise(eps=t), None, None, 1, 0, L.stream_ptr()) == -1
    assert b"must be positive" in lib.bsi_last_error()
    assert lib.bsi_step_fused(L.ptr(t), None, L.ptr(t), None, 0, 1, L.noise(eps=t), None, None, 1, 8, L.stream_ptr()) == -1
    assert lib.bsi_bucketize(L.ptr(t), None, None, 0.0, 1.0, 4, 8, L.stream_ptr()) == -1
    assert lib.bsi_device_arch() == 100
