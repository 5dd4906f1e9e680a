"""Size-independent properties of the CPU oracle (hypothesis, small random cases): what the GPU tests at full size rely
on when no reference output exists -- weights sum to acc and stay in [0, 1], inverse-CDF samples stay inside their
bins and are sorted for sorted u, ray independence and range invariants of the multi-field compositing,
and the restated Adam recurrence tracks torch.optim.Adam step for step."""
import torch
from hypothesis import given, settings, strategies as st

from oracle import star_oracle as so, train_oracle as to

SET = dict(max_examples=20, deadline=None, derandomize=True, database=None)   # same examples on every run


def gen(seed):
    return torch.Generator().manual_seed(seed)


@settings(**SET)
@given(seed=st.integers(0, 10_000), R=st.integers(1, 7), S=st.integers(2, 40), white=st.booleans())
def test_raw2outputs_weights_are_a_sub_probability(seed, R, S, white):
    g = gen(seed)
    raw_a, raw_c = 3.0 * torch.randn(R, S, generator=g), torch.randn(R, S, 3, generator=g)
    z = 2.0 + 4.0 * torch.sort(torch.rand(R, S, generator=g), dim=1).values
    rd = torch.randn(R, 3, generator=g)
    o = so.raw2outputs(raw_a, raw_c, z, rd, 0.0, white, 1e10)
    w = o["weights"]
    assert float(w.min()) >= 0.0 and float(w.max()) <= 1.0 + 1e-6
    assert torch.allclose(w.sum(-1), o["acc"], atol=1e-6)
    assert float(o["acc"].max()) <= 1.0 + 1e-5
    assert torch.isfinite(o["rgb"]).all() and torch.isfinite(o["depth"]).all()
    if white:       # rgb + (1 - acc): never below the composited colour
        assert float(o["rgb"].min()) >= -1e-6


@settings(**SET)
@given(seed=st.integers(0, 10_000), R=st.integers(1, 6), nb=st.integers(3, 33), Ni=st.integers(1, 48), det=st.booleans())
def test_sample_pdf_stays_in_range_and_is_monotone_in_u(seed, R, nb, Ni, det):
    g = gen(seed)
    bins = torch.sort(torch.rand(R, nb, generator=g), dim=1).values
    weights = torch.rand(R, nb - 1, generator=g) ** 3
    u = None if det else torch.sort(torch.rand(R, Ni, generator=g), dim=1).values
    s = so.sample_pdf(bins, weights, Ni, det=det, u=u)
    assert s.shape == (R, Ni)
    assert (s >= bins[:, :1] - 1e-6).all() and (s <= bins[:, -1:] + 1e-6).all()
    assert (s[:, 1:] >= s[:, :-1] - 1e-6).all()        # sorted u -> sorted samples (the merge in the kernel relies on it)
    s2 = so.sample_pdf(bins, weights, Ni, det=det, u=u, exact_sum=True)
    # the defined arithmetic differs from torch.sum only in rounding, but (u - cdf_lo) / denom with a small denom
    # amplifies a last-bit difference of the cdf (sample_pdf conditioning): the free-running tolerance
    assert torch.allclose(s, s2, atol=5e-3)


@settings(**SET)
@given(seed=st.integers(0, 10_000), R=st.integers(1, 5), S=st.integers(2, 24), V=st.integers(1, 3))
def test_multi_field_compositing_invariants(seed, R, S, V):
    """(The reference rectifies the SUM of the raw densities, rendering__.py:412-418, so a very negative object raw
    does not give an "empty object" limit: it empties the whole sample.  What does hold for any input:)"""
    g = gen(seed)
    ra_s, rc_s = 2.0 * torch.randn(R, S, generator=g), torch.randn(R, S, 3, generator=g)
    ra_d, rc_d = 2.0 * torch.randn(R, V, S, generator=g), torch.randn(R, V, S, 3, generator=g)
    z = 0.03 + 0.77 * torch.sort(torch.rand(R, S, generator=g), dim=1).values
    rd = torch.randn(R, 3, generator=g)
    o = so.raw2outputs_star(ra_s, rc_s, ra_d, rc_d, z, rd, white_bkgd=False, far_dist=1e10, test=True)
    assert torch.allclose(o["weights"].sum(-1), o["acc"], atol=1e-6)
    assert float(o["weights"].min()) >= 0.0 and float(o["acc"].max()) <= 1.0 + 1e-5
    T = o["dynamic_transmittance"]
    assert T.shape == (R, V) and float(T.min()) >= 0.0 and float(T.max()) <= 1.0 + 1e-6
    for k in ("loss_alpha_entropy", "loss_dynamic_vs_static_reg", "loss_ray_reg", "loss_static_reg", "loss_dynamic_reg"):
        assert torch.isfinite(torch.as_tensor(o[k])).all(), k
    # rays are independent: compositing a subset of the rays gives the same per-ray outputs
    sub = so.raw2outputs_star(ra_s[:1], rc_s[:1], ra_d[:1], rc_d[:1], z[:1], rd[:1], white_bkgd=False, far_dist=1e10,
                              test=True)
    for k in ("rgb", "acc", "depth", "weights", "dynamic_transmittance"):
        assert torch.allclose(sub[k], o[k][:1], atol=1e-6), k
    # the quirk itself: an object raw of -1e4 empties the sample instead of removing the object
    dead = so.raw2outputs_star(ra_s, rc_s, torch.full_like(ra_d, -1e4), rc_d, z, rd, white_bkgd=False, far_dist=1e10,
                               test=True)
    assert float(dead["acc"].abs().max()) == 0.0


@settings(max_examples=10, deadline=None, derandomize=True, database=None)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 50), steps=st.integers(1, 5),
       lr=st.sampled_from([1e-4, 5e-4, 1e-2]), scale=st.sampled_from([1e-6, 1.0, 30.0]))
def test_adam_recurrence_tracks_torch_adam(seed, n, steps, lr, scale):
    g = gen(seed)
    p0 = torch.randn(n, generator=g)
    grads = [scale * torch.randn(n, generator=g) for _ in range(steps)]
    ref_p, ref_m, ref_v, _ = to.clip_and_adam([[p0]], [[x] for x in grads], [lr])
    p, m, v = to.adam_restated(p0, grads, lr)
    assert torch.allclose(p, ref_p[0], rtol=2e-6, atol=2e-7 * max(1.0, lr * 1e3))
    assert torch.allclose(m, ref_m[0], rtol=2e-6, atol=1e-7 * scale)
    assert torch.allclose(v, ref_v[0], rtol=2e-6, atol=1e-12 * scale * scale)
